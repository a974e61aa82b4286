"""Micro-benchmarks of the two tcgen05 mainloops at the C2 shapes (run on a GPU box)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavenets_b200 import _lib
lib = _lib.load()
B, T = 8, 8000
def run(which, nseg, shifts, K, N, lda, ldg, reps=20):
  A = (torch.randn(B, T, lda, device='cuda') * 0.5).to(torch.bfloat16)
  if which == 0:
    G = (torch.randn(N, nseg * K, device='cuda') * 0.05).to(torch.bfloat16)
    out = torch.empty(B, T, N, device='cuda', dtype=torch.bfloat16); ldo = N
  else:
    G = (torch.randn(B, T, ldg, device='cuda') * 0.5).to(torch.bfloat16)
    out = torch.empty(nseg * K, N, device='cuda', dtype=torch.float32); ldo = N
  ms = C.c_float()
  sh = (C.c_int * nseg)(*shifts)
  _lib.check(lib.wn_debug_bench(which, reps, C.c_void_p(A.data_ptr()), lda, C.c_void_p(G.data_ptr()), ldg, B, T, nseg, sh, K, N,
                                C.c_void_p(out.data_ptr()), ldo, C.byref(ms)))
  fl = 2.0 * B * T * nseg * K * N
  return ms.value * 1e3, fl / (ms.value * 1e-3) / 1e12
for name, args in [
  ('conv gemm  K=2x256 N=512 (gate shape)', (0, 2, [-64, 0], 256, 512, 256, 0)),
  ('conv gemm  K=1x256 N=256 (conv1 shape)', (0, 1, [0], 256, 256, 256, 0)),
  ('conv gemm  K=2x512 N=256 (dgrad shape)', (0, 2, [64, 0], 512, 256, 512, 0)),
  ('wgrad+fin  K=2x256 N=512 (dilated)', (1, 2, [-64, 0], 256, 512, 256, 512)),
  ('wgrad only K=2x256 N=512 (dilated)', (2, 2, [-64, 0], 256, 512, 256, 512)),
  ('wgrad only K=1x256 N=512 (conv1|skip)', (2, 1, [0], 256, 512, 256, 512)),
]:
  us, tf = run(*args)
  print(f'{name:45s} {us:8.1f} us  {tf:7.1f} TFLOP/s')
