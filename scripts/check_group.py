"""Grouped weight-gradient launch (gemm_tc_wgroup.cuh) against the per-block launches (WN_TC_GROUP_WGRAD=0): same model,
weights and inputs, gradients must agree to fp32 summation order."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from wavenets_b200 import WaveNet
from tests.util import make_inputs

CASES = {
  'cond_skip_l2': (dict(channels=256, blocks=3, layers_per_block=1, dilation_bound=4, skip_channels=256, final_layers_channels=[256],
                        activation='tanh', conditioning='global', mapping_layers=[8], mapping_activation='tanh', l2_reg_factor=0.01), 3, 333),
  'noskip_k3': (dict(channels=256, blocks=2, layers_per_block=1, dilation_bound=9, kernel_size=3, use_skip=False, final_layers_channels=[128]), 2, 200),
  'multidil': (dict(channels=256, blocks=2, layers_per_block=2, dilation_bound=8, skip_channels=256, final_layers_channels=[128], activation='leaky_relu'), 2, 257),
  'wide512': (dict(channels=256, dilation_channels=512, blocks=2, layers_per_block=1, dilation_bound=4, skip_channels=512, final_layers_channels=[128]), 5, 130),
  'dropout': (dict(channels=256, blocks=2, layers_per_block=1, dilation_bound=4, skip_channels=256, final_layers_channels=[128], dropout=0.2), 2, 300),
}

def run(kw, B, T, group):
  os.environ['WN_TC_GROUP_WGRAD'] = '1' if group else '0'
  cond_in = 9 if kw.get('conditioning') else 0
  m = WaveNet(**kw, precision='bf16')
  x, cond = make_inputs(B, T, cond_in)
  m.build((x[:, :-1].shape, cond.shape) if cond is not None else x[:, :-1].shape)
  m.handle.glorot_init(seed=3, bias_std=0.05)
  if kw.get('dropout'):
    rng = np.random.default_rng(5)
    m.set_dropout_masks([(rng.random((B, T, kw['channels'])) > kw['dropout']).astype(np.uint8) for _ in range(kw['blocks'])])
  outs = []
  for _ in range(3):   # eager, eager, graph replay
    out = m.train_step((x, cond) if cond is not None else x)
    outs.append((out['loss'], m.get_grads()))
  return outs

bad = 0
for name, (kw, B, T) in CASES.items():
  try:
    a = run(kw, B, T, True); b = run(kw, B, T, False)
  except Exception as e:
    print(name, 'EXC', repr(e)[:300]); bad += 1; continue
  worst, wk = 0.0, None
  for k in b[0][1]:
    e = float(np.abs(a[0][1][k] - b[0][1][k]).max() / (np.abs(b[0][1][k]).max() + 1e-30))
    if e > worst: worst, wk = e, k
  rep = all(np.array_equal(a[0][1][k], a[i][1][k]) for k in a[0][1] for i in (1, 2))
  ok = worst < 2e-5 and rep and a[0][0] == b[0][0]
  bad += not ok
  print(f'{name}: loss {a[0][0]:.6f} vs {b[0][0]:.6f}  worst grad rel diff {worst:.2e} ({wk})  replay bit-equal {rep}  {"ok" if ok else "FAIL"}')
sys.exit(1 if bad else 0)
