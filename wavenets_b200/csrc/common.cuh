// common.cuh — shared device helpers for libwavenet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

enum { ACT_LINEAR = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_TANH = 3, ACT_SIGMOID = 4 };
// Keras 3 activation='leaky_relu' => negative_slope 0.2 (reference defaults.yaml:13,16)
#define WN_LEAKY_SLOPE 0.2f

// ---------------------------------------------------------------- math
// FAST=false: accurate libm-grade functions (fp32 parity tier, <=1e-4 vs the fp64 oracle).
// FAST=true : MUFU tanh.approx (1 SFU op each) for the bf16 throughput tier.
// -DWN_PRECISE_MATH builds the "precise" flavour of the library (libwavenet_b200_precise.so, test infrastructure): the bf16
// tier's gates use the accurate functions too, everything else is the same code.  MUFU.TANH's 2^-11 relative error is as
// large as half a bf16 ulp, so against the bf16-faithful oracle (oracle/faithful.py) it flips ~10 % of the stored bf16
// values by one ulp; the precise flavour shows that this, and nothing else, separates the shipped kernels from that oracle.
#ifdef WN_PRECISE_MATH
#define WN_FAST_MATH(F) false
#else
#define WN_FAST_MATH(F) (F)
#endif
template <bool FAST> __device__ __forceinline__ float wn_tanh(float x) {
  if constexpr (WN_FAST_MATH(FAST)) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  } else {
    return tanhf(x);
  }
}
template <bool FAST> __device__ __forceinline__ float wn_sigmoid(float x) {
  if constexpr (WN_FAST_MATH(FAST)) {
    return fmaf(wn_tanh<true>(0.5f * x), 0.5f, 0.5f);
  } else {
    return 1.0f / (1.0f + expf(-x));
  }
}
template <bool FAST> __device__ __forceinline__ float wn_act(int act, float x) {
  switch (act) {
    case ACT_RELU: return fmaxf(x, 0.0f);
    case ACT_LEAKY: return x >= 0.0f ? x : WN_LEAKY_SLOPE * x;
    case ACT_TANH: return wn_tanh<FAST>(x);
    case ACT_SIGMOID: return wn_sigmoid<FAST>(x);
    default: return x;
  }
}
// 16 values at once with ONE dispatch on the activation.  (A per-element `switch (act)` inside an unrolled loop compiles to a
// branch chain per element: measured 21 k cycles per 256 x 256 epilogue tile of a leaky_relu conv in the stack-forward kernel
// against 4 k for the same tile without activation — in-kernel clock64 accounting, profiles/r2.)
template <bool FAST> __device__ __forceinline__ void wn_act16(int act, float* v) {
  if (act == ACT_LEAKY) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = v[i] >= 0.0f ? v[i] : WN_LEAKY_SLOPE * v[i];
  } else if (act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
  } else if (act == ACT_TANH) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = wn_tanh<FAST>(v[i]);
  } else if (act == ACT_SIGMOID) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = wn_sigmoid<FAST>(v[i]);
  }
}
// derivative expressed through the activation OUTPUT y
__device__ __forceinline__ float wn_act_grad_from_out(int act, float y) {
  switch (act) {
    case ACT_RELU: return y > 0.0f ? 1.0f : 0.0f;
    case ACT_LEAKY: return y >= 0.0f ? 1.0f : WN_LEAKY_SLOPE;
    case ACT_TANH: return 1.0f - y * y;
    case ACT_SIGMOID: return y * (1.0f - y);
    default: return 1.0f;
  }
}
// v[i] *= act'(y[i]) for 16 values, one dispatch (see wn_act16)
__device__ __forceinline__ void wn_act_grad16(int act, const float* y, float* v) {
  if (act == ACT_LEAKY) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = y[i] >= 0.0f ? v[i] : WN_LEAKY_SLOPE * v[i];
  } else if (act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = y[i] > 0.0f ? v[i] : 0.0f;
  } else if (act == ACT_TANH) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] *= 1.0f - y[i] * y[i];
  } else if (act == ACT_SIGMOID) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] *= y[i] * (1.0f - y[i]);
  }
}

// ---------------------------------------------------------------- typed element access
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <class T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 16 consecutive elements; `vec` => pointer is 16-byte aligned and all 16 are in range.
template <class T> __device__ __forceinline__ void load16(const T* p, float* v, int n_valid, bool vec);
template <> __device__ __forceinline__ void load16<float>(const float* p, float* v, int n_valid, bool vec) {
  if (vec && n_valid >= 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 q = reinterpret_cast<const float4*>(p)[i];
      v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = i < n_valid ? p[i] : 0.0f;
  }
}
template <> __device__ __forceinline__ void load16<bf16>(const bf16* p, float* v, int n_valid, bool vec) {
  if (vec && n_valid >= 16) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint4 q = reinterpret_cast<const uint4*>(p)[i];
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[8 * i + 2 * j] = __uint_as_float(w[j] << 16);
        v[8 * i + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = i < n_valid ? __bfloat162float(p[i]) : 0.0f;
  }
}
template <class T> __device__ __forceinline__ void store16(T* p, const float* v, int n_valid, bool vec);
template <> __device__ __forceinline__ void store16<float>(float* p, const float* v, int n_valid, bool vec) {
  if (vec && n_valid >= 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < n_valid) p[i] = v[i];
  }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
template <> __device__ __forceinline__ void store16<bf16>(bf16* p, const float* v, int n_valid, bool vec) {
  if (vec && n_valid >= 16) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint4 q;
      q.x = pack_bf16x2(v[8 * i], v[8 * i + 1]);
      q.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
      q.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
      q.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
      reinterpret_cast<uint4*>(p)[i] = q;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < n_valid) p[i] = __float2bfloat16_rn(v[i]);
  }
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- quantiser (model.py:151-155)
// index = #{j in 1..2^bits-1 : -1 + j*2^(1-bits) <= x}.  Every boundary is exactly
// representable in fp32 for bits <= 16, and the comparisons are done on the fp32 boundary
// values themselves (floor((x+1)*2^(bits-1)) alone is wrong for tiny negative x).
__device__ __forceinline__ int wn_quantize_idx(float x, int bits) {
  const int nb = 1 << bits;
  const float half = (float)(1 << (bits - 1));
  const float inv = 1.0f / half;
  float g = floorf((x + 1.0f) * half);
  int k = g < 0.0f ? 0 : (g > (float)(nb - 1) ? nb - 1 : (int)g);
  while (k < nb - 1 && fmaf((float)(k + 1), inv, -1.0f) <= x) ++k;
  while (k > 0 && fmaf((float)k, inv, -1.0f) > x) --k;
  return k;
}
