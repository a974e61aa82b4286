"""GPU tests of the two tcgen05 mainloops (gemm_tc.cuh) in isolation, through the C ABI test
hooks, against a plain PyTorch fp32 reference of the same contraction on the same bf16 inputs.
fp32 accumulation order differs, nothing else: tolerance 2e-5 relative to the output scale.
The shifted rows (causal padding, batch isolation) must be exact zeros."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
  from wavenets_b200 import _lib
  return _lib, _lib.load()


def _ref_conv(A, W, shifts, K):
  """A (B,T,lda) bf16, W (N16, nseg*K) bf16 -> (B,T,N16) fp32."""
  B, T, _ = A.shape
  out = torch.zeros(B, T, W.shape[0], dtype=torch.float32, device=A.device)
  Af = A.float()
  for s, sh in enumerate(shifts):
    sl = torch.zeros(B, T, K, dtype=torch.float32, device=A.device)
    lo, hi = max(0, -sh), min(T, T - sh)
    if hi > lo:
      sl[:, lo:hi] = Af[:, lo + sh:hi + sh, :K]
    out += sl @ W[:, s * K:(s + 1) * K].float().T
  return out


@pytest.mark.parametrize('B,T,K,N,tile,shifts', [
  (1, 128, 64, 64, 64, [0]),
  (2, 300, 64, 128, 128, [-3, 0]),
  (3, 517, 128, 256, 256, [-64, 0]),
  (2, 1000, 256, 512, 256, [-512, 0]),
  (2, 200, 64, 64, 64, [5, 0]),          # dgrad-style anti-causal shift
  (1, 90, 192, 128, 128, [-2, -1, 0]),   # kernel_size 3
  (2, 131, 64, 30, 64, [0]),             # ragged N (mixture head)
])
def test_conv_gemm_matches_torch(B, T, K, N, tile, shifts):
  _l, lib = _lib()
  g = torch.Generator(device='cuda').manual_seed(0)
  lda = K + 64
  A = (torch.randn(B, T, lda, device='cuda', generator=g) * 0.5).to(torch.bfloat16)
  N16 = (N + tile - 1) // tile * tile
  W = torch.zeros(N16, len(shifts) * K, device='cuda', dtype=torch.bfloat16)
  W[:N] = (torch.randn(N, len(shifts) * K, device='cuda', generator=g) * 0.1).to(torch.bfloat16)
  out = torch.full((B, T, N), float('nan'), device='cuda', dtype=torch.float32)
  sh = (C.c_int * len(shifts))(*shifts)
  _l.check(lib.wn_debug_conv_gemm(C.c_void_p(A.data_ptr()), lda, B, T, len(shifts), sh, K, C.c_void_p(W.data_ptr()), N, N16, tile,
                                  C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
  ref = _ref_conv(A, W, shifts, K)[:, :, :N]
  err = (out - ref).abs().max().item() / (ref.abs().max().item() + 1e-30)
  assert err < 2e-5, err


@pytest.mark.parametrize('B,T,K,N,shifts', [
  (1, 64, 64, 64, [0]),
  (2, 300, 64, 128, [-3, 0]),
  (3, 517, 128, 256, [-64, 0]),
  (2, 2000, 256, 512, [-512, 0]),
  (2, 131, 64, 30, [0]),
  (1, 90, 192, 128, [-2, -1, 0]),
])
def test_wgrad_matches_torch(B, T, K, N, shifts):
  _l, lib = _lib()
  g = torch.Generator(device='cuda').manual_seed(1)
  lda, ldg = K + 64, (N + 63) // 64 * 64 + 64
  A = (torch.randn(B, T, lda, device='cuda', generator=g) * 0.5).to(torch.bfloat16)
  G = (torch.randn(B, T, ldg, device='cuda', generator=g) * 0.5).to(torch.bfloat16)
  out = torch.full((len(shifts) * K, N), float('nan'), device='cuda', dtype=torch.float32)
  sh = (C.c_int * len(shifts))(*shifts)
  _l.check(lib.wn_debug_wgrad(C.c_void_p(A.data_ptr()), lda, C.c_void_p(G.data_ptr()), ldg, B, T, len(shifts), sh, K, N,
                              C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
  Af, Gf = A.float(), G.float()[:, :, :N]
  ref = torch.zeros(len(shifts) * K, N, device='cuda')
  for s, shv in enumerate(shifts):
    sl = torch.zeros(B, T, K, device='cuda')
    lo, hi = max(0, -shv), min(T, T - shv)
    if hi > lo:
      sl[:, lo:hi] = Af[:, lo + shv:hi + shv, :K]
    ref[s * K:(s + 1) * K] = torch.einsum('btk,btn->kn', sl, Gf)
  err = (out - ref).abs().max().item() / (ref.abs().max().item() + 1e-30)
  assert err < 2e-5, err
