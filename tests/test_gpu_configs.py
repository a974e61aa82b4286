"""Oracle parity at the FIVE BASELINE.json topologies (`wavenets_b200.config.CONFIGS`: real block counts, widths and
dilation bounds — dilations up to 512, 30-40 blocks, K = L*D skip sum), through the C ABI, on seeded synthetic audio.

Checkers (both independent of the CUDA code; oracle/ is test infrastructure):
  * `oracle/faithful.py`, faithful=False — the reference path in float64 (model.py:213-239,309-335,505-551;
    layers.py:178-224).  fp32 tier: <= 1e-4 relative (north star).  bf16 tier: the STATED bf16 tolerance
    (loss 1e-2, gradients rel-L2 3e-2 per tensor): this bounds the effect of bf16 storage, not the kernels.
  * `oracle/faithful.py`, faithful=True — the same model rounding to bf16 exactly where the kernels do.  What is left is
    fp32-vs-fp64 accumulation order and MUFU.TANH (2^-11 relative): gradients rel-L2 <= TOL_FAITHFUL per tensor for every
    activation, leaky_relu included (a kernel bug of a few percent cannot hide behind bf16 storage error here).

Shapes: B x T chosen so that a layer has MORE 256-row tiles than the GPU has CTA pairs (74): the persistent stack-forward
launch with its inter-layer flags, the grouped weight gradients with side launches and CUDA-graph replay are all on
the checked path (asserted below).  T > receptive field + 512.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import faithful
from oracle import wavenet_oracle as wo
from tests.util import oracle_config, rel_l2

pytestmark = pytest.mark.gpu

TOL_LOSS_BF16, TOL_GRAD_BF16 = 1e-2, 3e-2      # vs float64: stated bf16 tolerance
TOL_LOSS_FAITHFUL, TOL_GRAD_FAITHFUL = 2e-4, 4e-3   # vs the bf16-faithful oracle
TOL_FP32 = 1e-4

# config -> (B, T, expects the stack-forward launch, expects grouped weight gradients)
SHAPES = {
  'c1': (1, 8000, False, False),      # BASELINE configs[0] exactly: defaults.yaml topology, batch 1, fp32
  'c2': (3, 6400, True, True),
  'c3': (3, 2048, False, True),
  'c4': (3, 6400, True, True),
  'c5': (3, 6400, True, True),
}


def _setup(name):
  from wavenets_b200 import CONFIGS, WaveNet, model_kwargs, synth
  cfg = dict(CONFIGS[name])
  kw = model_kwargs(cfg)
  B, T, want_stack, want_group = SHAPES[name]
  cond_in = cfg.get('n_speakers', 109) if kw['conditioning'] == 'global' else 0
  ocfg = oracle_config(kw, cond_in)
  _, rf = wo.dilation_schedule(ocfg)
  assert T >= rf + 512
  p = wo.init_params(ocfg, seed=1)              # glorot-uniform kernels, N(0, 0.02) biases (zero biases would hide bias bugs)
  x = synth.frames(B, T, seed=3, apply_mulaw=cfg.get('apply_mulaw', True))
  cond = synth.speakers_onehot(B, cond_in, seed=3) if cond_in else None
  m = WaveNet(**kw, precision=cfg['precision'], max_batch=B, max_time=T)
  m.build(((B, T, 1), (B, cond_in)) if cond_in else (B, T, 1))
  assert m.receptive_field == rf
  m.set_weights({k: v.astype(np.float32) for k, v in p.items()})
  # the kernels see the fp32 weights: the oracle gets the same values
  p32 = {k: v.astype(np.float32).astype(np.float64) for k, v in p.items()}
  return m, ocfg, p32, x, cond, cfg['precision'], want_stack, want_group


def _worst(g, ref):
  scale = max(float(np.linalg.norm(v)) for v in ref.values())
  worst, who = 0.0, None
  for k, r in ref.items():
    n = float(np.linalg.norm(r))
    if n <= 1e-9 * scale:
      # a conv whose output nothing reads (conv1 of the last block under use_skip): zero in the reference, zero here
      assert float(np.abs(g[k]).max()) <= 1e-6 * scale, k
      continue
    e = rel_l2(g[k], r)
    if e > worst:
      worst, who = e, k
  return worst, who


@pytest.mark.parametrize('name', ['c1', 'c2', 'c3', 'c4', 'c5'])
def test_config_topology_vs_oracle(name):
  m, ocfg, p, x, cond, precision, want_stack, want_group = _setup(name)
  data = (x, cond) if cond is not None else x
  h = m.handle
  # three steps: eager; plans built -> side launches; CUDA-graph replay.  The third is the one compared.
  for _ in range(3):
    out = m.train_step(data)
  g = m.get_grads()
  stack_layers = int(h.lib.wn_stack_forward_layers(h.h))
  side = C.c_int(0)
  tiles = int(h.lib.wn_grouped_wgrad_tiles(h.h, C.byref(side)))
  if want_stack:
    assert stack_layers == m.blocks, 'the persistent stack-forward launch must be on the checked path'
  if want_group:
    assert tiles > 0, 'the grouped weight-gradient launch must be on the checked path'
  l_exact, g_exact = faithful.train_step(p, ocfg, x, cond, faithful=False)
  e_exact, who_exact = _worst(g, g_exact)
  msg = f'{name}: loss cuda {out["loss"]:.6f} fp64 {l_exact:.6f}; worst grad rel-L2 vs fp64 {e_exact:.2e} ({who_exact})'
  if precision == 'fp32':
    print(msg)
    assert abs(out['loss'] - l_exact) <= TOL_FP32 * abs(l_exact), msg
    assert e_exact <= TOL_FP32, msg
    return
  l_f, g_f = faithful.train_step(p, ocfg, x, cond, faithful=True)
  e_f, who_f = _worst(g, g_f)
  msg += f'; vs bf16-faithful: loss {abs(out["loss"] - l_f) / abs(l_f):.2e}, worst grad {e_f:.2e} ({who_f}); stack layers {stack_layers}, grouped tiles {tiles}, side launches {side.value}'
  print(msg)
  assert abs(out['loss'] - l_exact) <= TOL_LOSS_BF16 * abs(l_exact), msg
  assert e_exact <= TOL_GRAD_BF16, msg
  assert abs(out['loss'] - l_f) <= TOL_LOSS_FAITHFUL * abs(l_f), msg
  assert e_f <= TOL_GRAD_FAITHFUL, msg


def test_c2_forward_vs_faithful():
  """WaveNet.call at the C2 topology: probabilities against the bf16-faithful oracle (and softmax rows sum to one)."""
  m, ocfg, p, x, cond, precision, _, _ = _setup('c2')
  y = m((x[:, :-1], cond)).cpu().numpy()
  ref = faithful.forward(p, ocfg, x[:, :-1], cond, faithful=True)
  ref64 = faithful.forward(p, ocfg, x[:, :-1], cond, faithful=False)
  e_f, e_64 = rel_l2(y, ref), rel_l2(y, ref64)
  print(f'c2 forward: rel-L2 vs bf16-faithful {e_f:.2e}, vs fp64 {e_64:.2e}')
  assert e_f <= 2e-3 and e_64 <= 1e-2
  assert np.allclose(y.sum(-1), 1.0, atol=1e-4)
