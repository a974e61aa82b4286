"""Parity properties at BASELINE.json's FULL sizes (C2: 30 blocks, R=D=S=256, global conditioning, T=8000, 8 sequences),
where the CPU oracle is too slow to be the checker: size-independent properties of the path instead."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def c2():
  from wavenets_b200 import CONFIGS, WaveNet, model_kwargs, synth
  cfg = dict(CONFIGS['c2'])
  kw = model_kwargs(cfg)
  B, T = cfg['batch_size'], cfg['recording_length']
  m = WaveNet(**kw, precision='bf16', max_batch=B, max_time=T)
  m.build(((B, T, 1), (B, 109)))
  m.handle.glorot_init(seed=1, bias_std=0.02)
  x = torch.from_numpy(synth.frames(B, T, seed=0)).cuda()
  c = torch.from_numpy(synth.speakers_onehot(B, 109, seed=0)).cuda()
  return m, x, c, B, T


def test_fullsize_determinism_and_loss_scaling(c2):
  m, x, c, B, T = c2
  m.n_replicas = 1
  l1 = m.train_step((x, c))['loss']
  g1 = m.handle.flat_grads.clone()
  for _ in range(2):                                    # eager, then CUDA-graph replays
    assert m.train_step((x, c))['loss'] == l1
    assert torch.equal(g1, m.handle.flat_grads)         # no atomics anywhere: bit-identical gradients
  assert np.isfinite(l1) and bool(torch.isfinite(g1).all())
  # compute_average_loss divides by B * n_replicas (model.py:328): a power of two -> every gradient scales EXACTLY
  m.n_replicas = 2
  l2 = m.train_step((x, c))['loss']
  assert l2 == 0.5 * l1
  assert torch.equal(m.handle.flat_grads, 0.5 * g1)
  m.n_replicas = 1


def test_fullsize_causality_receptive_field_and_batch_isolation(c2):
  m, x, c, B, T = c2
  rf = m.receptive_field
  assert rf == 3071
  xin = x[:, :-1].clone()
  y0 = m((xin, c))
  t0 = 3500
  x2 = xin.clone()
  x2[2, t0, 0] += 0.25
  y2 = m((x2, c))
  d = (y2 != y0).any(dim=-1)                             # (B, T) which outputs changed at all
  assert not bool(d[[0, 1, 3, 4, 5, 6, 7]].any())       # other sequences: bit-identical
  assert not bool(d[2, :t0].any())                      # the past: bit-identical (causal)
  assert bool(d[2, t0])                                 # the present depends on the input
  assert not bool(d[2, t0 + rf:].any())                 # beyond the receptive field: bit-identical
  assert bool(d[2, t0 + rf - 1])                        # and the last sample inside it does change
  p = y0[0, 100:110].sum(dim=-1)
  assert torch.allclose(p, torch.ones_like(p), atol=1e-4)    # softmax rows


def test_fullsize_batch_permutation(c2):
  m, x, c, B, T = c2
  m.n_replicas = 1
  l1 = m.train_step((x, c))['loss']
  g1 = m.handle.flat_grads.clone()
  perm = torch.tensor([3, 0, 7, 1, 6, 2, 5, 4], device=x.device)
  l2 = m.train_step((x[perm].contiguous(), c[perm].contiguous()))['loss']
  g2 = m.handle.flat_grads
  # sequences are independent; only the fp32 summation order of the split reductions changes
  assert abs(l2 - l1) <= 1e-5 * abs(l1)
  num = float((g2 - g1).norm()); den = float(g1.norm())
  assert num <= 2e-3 * den, num / den


def test_fullsize_schedules_agree(c2):
  """C2 at full size through the default schedules (stack forward as one persistent launch, grouped weight gradients with
  side launches) against one launch per block for everything (WN_TC_STACK_FWD=0, WN_TC_GROUP_WGRAD=0): the forward is the
  same tile arithmetic (loss bit-equal), the gradients differ by fp32 summation order only."""
  import os
  from wavenets_b200 import CONFIGS, WaveNet, model_kwargs
  m, x, c, B, T = c2
  m.n_replicas = 1
  for _ in range(3):                                     # plan built, side launches, graph replay
    l_new = m.train_step((x, c))['loss']
  g_new = m.get_grads()
  old = {k: os.environ.get(k) for k in ('WN_TC_STACK_FWD', 'WN_TC_GROUP_WGRAD')}
  os.environ['WN_TC_STACK_FWD'] = '0'
  os.environ['WN_TC_GROUP_WGRAD'] = '0'
  try:
    kw = model_kwargs(dict(CONFIGS['c2']))
    m2 = WaveNet(**kw, precision='bf16', max_batch=B, max_time=T)     # the switches are read at wn_create
    m2.build(((B, T, 1), (B, 109)))
    m2.set_weights(m.get_weights())
    l_old = m2.train_step((x, c))['loss']
    g_old = m2.get_grads()
  finally:
    for k, v in old.items():
      if v is None:
        os.environ.pop(k, None)
      else:
        os.environ[k] = v
  assert int(m.handle.lib.wn_stack_forward_layers(m.handle.h)) == 30 and int(m2.handle.lib.wn_stack_forward_layers(m2.handle.h)) == 0
  assert int(m.handle.lib.wn_grouped_wgrad_tiles(m.handle.h, None)) > 0 and int(m2.handle.lib.wn_grouped_wgrad_tiles(m2.handle.h, None)) == 0
  assert l_new == l_old
  for k in g_old:
    ref = g_old[k]
    err = float(np.abs(g_new[k] - ref).max() / (np.abs(ref).max() + 1e-30))
    assert err < 3e-4, (k, err)       # 64,000 rows per gradient entry, fp32 chains of different lengths
