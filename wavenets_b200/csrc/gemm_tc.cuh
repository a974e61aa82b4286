// gemm_tc.cuh — bf16 tcgen05/TMEM/TMA mainloops (throughput tier).  STUB: filled in next.
#pragma once
#include "common.cuh"

#define WN_MAX_WGRAD_SPLITS 64

struct TmapCache { int unused = 0; };
struct TcSeg { const bf16* A; int lda; int shift; int K; };
struct TcGemmDesc {
  int B, T, nseg; TcSeg seg[4]; int n_outer; long long outer_stride;
  const bf16* W; int ktot; int N16; int tileN;
};
struct TcWgradDesc {
  int B, T, N; const bf16* G; int ldg; int nseg; TcSeg seg[4]; int ktot; float* partial;
};
static inline const char* tc_last_error() { return "tcgen05 tier not built"; }
static inline int tc_init() { return -1; }
static inline int tc_check_config(int R, int D, int S, int K) { return -1; }
static inline void tc_pick_tile(int n, bool gate, int* n16, int* tile) { *n16 = (n + 63) / 64 * 64; *tile = 64; }
static inline void tc_pack_gate_T(cudaStream_t, const float*, int, int, bf16*, int, int, int, int) {}
template <class Epi> static int tc_conv_gemm(TmapCache&, cudaStream_t, const TcGemmDesc&, const typename Epi::Params&) { return -1; }
static inline int tc_wgrad(TmapCache&, cudaStream_t, const TcWgradDesc&, int* nsplit) { *nsplit = 1; return -1; }
