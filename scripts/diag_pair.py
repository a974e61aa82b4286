"""Row-range error map of one block's forward in the bf16 tier (CTA-pair debugging)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import wavenet_oracle as wo
from wavenets_b200 import WaveNetLayer
R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 512
B = 1
rng = np.random.default_rng(5)
lay = WaveNetLayer(dilation_rate=[2], channels=R, skip_channels=None, precision='bf16')
x = rng.standard_normal((B, T, R)).astype(np.float32)
lay.build(x.shape)
w = {n: (rng.standard_normal(s) * 0.1).astype(np.float32) for n, s in zip(lay.weight_names, lay._handle.shapes)}
lay.set_weights(w)
for rep in range(2):
  xo, sk = lay(x)
  p = {'block0/' + k: v.astype(np.float64) for k, v in w.items()}
  lc = dict(dilations=[2], activation=None, residual=True, has_skip=False, condition=False)
  xo_o, sk_o, cache = wo.layer_forward(p, 'block0', lc, x.astype(np.float64), None)
  for name, a, b in (('x_out', xo.cpu().numpy(), xo_o), ('skip', sk.cpu().numpy(), sk_o)):
    for r0 in range(0, T, 64):
      e = np.abs(a[0, r0:r0 + 64] - b[0, r0:r0 + 64]).max(axis=0)
      cols = [float(e[c0:c0 + 32].max()) for c0 in range(0, R, 32)]
      print(rep, name, 'rows', r0, 'max err per 32-col group', ' '.join('%.3f' % c for c in cols), 'zero frac %.2f' % float((a[0, r0:r0 + 64] == 0).mean()))
