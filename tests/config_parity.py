"""One BASELINE topology through the C ABI against the float64 and the bf16-faithful oracle (shared by
tests/test_gpu_configs.py; also a script: `python -m tests.config_parity c2` prints one JSON line — the parity tests run it
in a subprocess with WN_LIB pointing at the precise-math flavour of the library)."""
import ctypes as C
import json
import sys

import numpy as np

from oracle import faithful
from oracle import wavenet_oracle as wo
from tests.util import device_slope_masks, oracle_config, rel_l2

# config -> (B, T, expects the stack-forward launch, expects grouped weight gradients)
SHAPES = {
  'c1': (1, 8000, False, False),      # BASELINE configs[0] exactly: defaults.yaml topology, batch 1, fp32
  'c2': (3, 6400, True, True),
  'c3': (3, 6400, True, True),        # 5 x 5 dilations per block: 20 plain convs + 5 gated convs in the one stack-forward launch
  'c4': (3, 6400, True, True),
  'c5': (3, 6400, True, True),
}


def setup(name, dropout=0.0):
  from wavenets_b200 import CONFIGS, WaveNet, model_kwargs, synth
  precision = None
  if name.endswith('_bf16'):       # 'c1_bf16': the defaults.yaml topology (32 channels) on the tcgen05 tier
    name, precision = name[:-5], 'bf16'
  cfg = dict(CONFIGS[name])
  if precision:
    cfg['precision'] = precision
  cfg['dropout'] = dropout
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
  return m, ocfg, p32, x, cond, cfg['precision'], want_stack, want_group, kw


def errors(g, ref):
  """per-tensor relative L2 errors, worst first; tensors whose reference gradient is zero must be zero"""
  scale = max(float(np.linalg.norm(v)) for v in ref.values())
  out = []
  for k, r in ref.items():
    n = float(np.linalg.norm(r))
    if n <= 1e-9 * scale:
      # a conv whose output nothing reads (conv1 of the last block under use_skip): zero in the reference, zero here
      assert float(np.abs(g[k]).max()) <= 1e-6 * scale, k
      continue
    out.append((rel_l2(g[k], r), k))
  return sorted(out, reverse=True)


def run(name, dropout=0.0):
  """dropout > 0: a TRAINING pass with injected keep-masks (TF's RNG stream cannot be reproduced): the same masks go to the oracle"""
  m, ocfg, p, x, cond, precision, want_stack, want_group, kw = setup(name, dropout)
  name = name[:-5] if name.endswith('_bf16') else name
  data = (x, cond) if cond is not None else x
  h = m.handle
  keep = None
  if dropout > 0:
    B, T = SHAPES[name][:2]
    rng = np.random.default_rng(17)
    keep = [(rng.random((B, T, kw['channels'])) >= dropout).astype(np.uint8) for _ in range(kw['blocks'])]
    m.set_dropout_masks(keep)
  # three steps: eager; plans built -> side launches; CUDA-graph replay.  The third is the one compared.
  for _ in range(3):
    out = m.train_step(data)
  g = m.get_grads()
  side = C.c_int(0)
  res = {'config': name, 'precision': precision, 'loss': out['loss'], 'stack_layers': int(h.lib.wn_stack_forward_layers(h.h)),
         'grouped_tiles': int(h.lib.wn_grouped_wgrad_tiles(h.h, C.byref(side))), 'side_launches': side.value, 'blocks': m.blocks,
         'want_stack': want_stack, 'want_group': want_group, 'stack_bwd_layers': int(h.lib.wn_stack_backward_layers(h.h)), 'build': h.lib.wn_build_info().decode()}
  l64, g64 = faithful.train_step(p, ocfg, x, cond, faithful=False, keep_masks=keep)
  e64 = errors(g, g64)
  res.update(loss_fp64=l64, worst_fp64=e64[0][0], worst_fp64_tensor=e64[0][1], top_fp64=e64[:6])
  if precision != 'fp32':
    # (the derivative masks of relu / leaky_relu come from the device: tests/util.py:device_slope_masks)
    B, T = SHAPES[name][:2]
    masks = device_slope_masks(m, kw, B, T)
    lf, gf = faithful.train_step(p, ocfg, x, cond, faithful=True, slope_masks=masks, keep_masks=keep)
    ef = errors(g, gf)
    res['slope_mask_sites'] = 0 if not masks else len(masks)
    glob = float(np.sqrt(sum(np.sum((g[k].astype(np.float64) - gf[k]) ** 2) for k in gf) / sum(np.sum(gf[k] ** 2) for k in gf)))
    res.update(loss_faithful=lf, worst_faithful=ef[0][0], worst_faithful_tensor=ef[0][1], top_faithful=ef[:6], global_faithful=glob,
               median_faithful=float(np.median([e for e, _ in ef])))
  return res


if __name__ == '__main__':
  print(json.dumps(run(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.0)))
