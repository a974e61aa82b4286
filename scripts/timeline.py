"""Per-tile timeline (clock64) of the conv GEMM kernels at the C2 block shape; needs a -DTC_TIMELINE build."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wavenets_b200 import WaveNetLayer
B, T, R = 8, 8000, 256
rng = np.random.default_rng(5)
lay = WaveNetLayer(dilation_rate=[64], channels=R, skip_channels=R, precision='bf16')
x = rng.standard_normal((B, T, R)).astype(np.float32)
lay.build(x.shape)
w = {n: (rng.standard_normal(s) * 0.05).astype(np.float32) for n, s in zip(lay.weight_names, lay._handle.shapes)}
lay.set_weights(w)
print('=== forward (gate, conv1[, skip])', flush=True)
xo, sk = lay(x)
torch.cuda.synchronize()
print('=== backward (wgrads, gate adjoint, dgrad)', flush=True)
dx, _ = lay.backward(rng.standard_normal((B, T, R)).astype(np.float32), rng.standard_normal((B, T, R)).astype(np.float32))
torch.cuda.synchronize()
