"""Where does the CUDA bf16 tier leave the bf16-faithful oracle?  Per stored tensor (forward values and backward gradients)
relative L2 error against oracle/faithful.py on a small C2-shaped model: python scripts/diag_faithful.py [ce|mol] [blocks]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import faithful, wavenet_oracle as wo
from tests.util import make_inputs, oracle_config, rel_l2
from wavenets_b200 import WaveNet

head = sys.argv[1] if len(sys.argv) > 1 else 'ce'
L = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kw = dict(channels=256, blocks=L, layers_per_block=1, dilation_bound=4, skip_channels=256, final_layers_channels=[256], activation='leaky_relu')
if head == 'mol':
  kw.update(num_mixtures=10, sampling_function='logistic', bits=16)
B, T = 2, 300
cfg = oracle_config(kw, 0)
p = wo.init_params(cfg, seed=1)
x, _ = make_inputs(B, T, 0)
m = WaveNet(**kw, precision='bf16')
m.build(x[:, :-1].shape)
m.set_weights({k: v.astype(np.float32) for k, v in p.items()})
out = m.train_step(x)
g = m.get_grads()
p32 = {k: v.astype(np.float32).astype(np.float64) for k, v in p.items()}
h = m.handle
def dev(name, idx, width_hint):
  buf = np.empty(B * T * width_hint, dtype=np.float32)
  w = h.lib.wn_debug_tensor(h.h, name.encode(), idx, buf.ctypes.data_as(C.c_void_p), buf.size)
  if w < 0:
    return None
  return buf[:B * T * w].reshape(B, T, w)
masks = {('hact', 0): dev('hact', 0, 512) >= 0} if os.environ.get('DIAG_MASKS', '1') == '1' else None
tap = {}
lf, gf = faithful.train_step(p32, cfg, x, None, faithful=True, tap=tap, slope_masks=masks)
print(f'head {head} blocks {L}: loss cuda {out["loss"]:.6f} faithful {lf:.6f} build {h.lib.wn_build_info().decode()[-8:]}')
def cmp(label, a, ref):
  if a is None:
    print(f'  {label:14s} n/a'); return
  ref = ref.detach().numpy() if hasattr(ref, 'detach') else ref
  d = np.abs(a - ref)
  print(f'  {label:14s} rel-L2 {rel_l2(a, ref):.2e}  max|d| {d.max():.2e}  frac(|d|>0) {np.mean(d > 0):.3f}  |ref| rms {np.sqrt(np.mean(ref**2)):.2e}')
Wd = 512
cmp('h0', dev('h0', 0, Wd), tap['h0'])
for l in range(L):
  zz = tap[('z', l)].detach(); Dh = zz.shape[-1] // 2      # the device caches the gate's derivative coefficients [P | Q], not z
  th_, sg_ = torch.tanh(zz[..., :Dh]), torch.sigmoid(zz[..., Dh:])
  cmp(f'PQ[{l}]', dev('z', l, Wd), faithful._bf16(torch.cat([sg_ * (1 - th_ * th_), th_ * sg_ * (1 - sg_)], -1)))
  cmp(f'g[{l}]', dev('g', l, Wd), tap[('g', l)])
  cmp(f'xout[{l}]', dev('xout', l, Wd), tap[('xout', l)])
cmp('skipsum', dev('skipsum', 0, Wd), tap['skipsum'])
cmp('hact[0]', dev('hact', 0, Wd), tap[('hact', 0)])
nl = len(kw['final_layers_channels'])
cmp('logits', dev('logits', 0, Wd), tap[('lin', nl)])
print(' backward')
cmp('dlogits', dev('dlogits', 0, Wd), tap[('lin', nl)].grad)
cmp('dskip', dev('dskip', 0, Wd), tap['skipsum'].grad)
for l in range(L - 1, -1, -1):
  cmp(f'dz[{l}]', dev('dz', l, Wd), tap[('z', l)].grad)
  if l < L - 1:
    cmp(f'dx[{l}]', dev('dx', l, Wd), tap[('xout', l)].grad)
errs = sorted(((rel_l2(g[k], gf[k]), k) for k in gf if np.linalg.norm(gf[k]) > 0), reverse=True)
print(' grads:', [(f'{e:.1e}', k) for e, k in errs])
