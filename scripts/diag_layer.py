import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import wavenet_oracle as wo
from tests.util import rel_l2
from wavenets_b200 import WaveNetLayer
for prec in ['fp32', 'bf16']:
  for (dils, act, S, T) in [([4], None, None, 200), ([4], None, 128, 200), ([1, 4], 'leaky_relu', 128, 200), ([1, 4], 'leaky_relu', 128, 2000), ([4], None, None, 2000)]:
    rng = np.random.default_rng(5)
    B, R = 2, 64
    lay = WaveNetLayer(dilation_rate=dils, activation=act, channels=R, skip_channels=S, precision=prec)
    x = rng.standard_normal((B, T, R)).astype(np.float32)
    lay.build(x.shape)
    w = {n: (rng.standard_normal(s) * 0.1).astype(np.float32) for n, s in zip(lay.weight_names, lay._handle.shapes)}
    lay.set_weights(w)
    xo, sk = lay(x)
    p = {'block0/' + k: v.astype(np.float64) for k, v in w.items()}
    lc = dict(dilations=dils, activation=act, residual=True, has_skip=S is not None, condition=False)
    xo_o, sk_o, cache = wo.layer_forward(p, 'block0', lc, x.astype(np.float64), None)
    dxo = rng.standard_normal(xo_o.shape).astype(np.float32)
    dsk = rng.standard_normal(sk_o.shape).astype(np.float32)
    dx, _ = lay.backward(dxo, dsk)
    dx_o, _, g_o = wo.layer_backward(p, 'block0', lc, cache, dxo.astype(np.float64), dsk.astype(np.float64))
    g = lay.get_grads()
    print(prec, dils, act, S, T, 'xo %.2e sk %.2e dx %.2e |' % (rel_l2(xo.cpu().numpy(), xo_o), rel_l2(sk.cpu().numpy(), sk_o), rel_l2(dx.cpu().numpy(), dx_o)),
          ' '.join('%s=%.2e' % (k[len('block0/'):], rel_l2(g[k[len('block0/'):]], v)) for k, v in g_o.items()))
