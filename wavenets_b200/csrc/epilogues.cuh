// epilogues.cuh — GEMM epilogues shared by the FFMA (fp32) and tcgen05 (bf16) mainloops.
//
// Contract: after a mainloop has produced a BM x BN fp32 accumulator tile, every epilogue
// thread owns ONE output row (batch b, time t; flat row index grow = b*T + t) and calls
//     Epi::row(params, acc, b, grow, n0, bn, q, nq)
// where acc.load16(c, v) yields accumulator columns [c, c+16) of that row (c relative to the
// tile), n0 is the tile's first output column, bn the tile width, and the thread handles the
// 16-column chunks q, q+nq, q+2nq, ...  (FFMA path: 4 threads per row; tcgen05 path: the
// TMEM lane's owner thread(s)).
//
// T  = storage type of activations in HBM (float | bf16); FAST selects MUFU approximations.
#pragma once
#include "common.cuh"

// out = act(acc + bias[n] + cbias[b][n]) + res[row][n]
// Used by: pre-stack dilated convs (layers.py:66-74), conv1 + residual add (layers.py:213,222-223),
// skip-sum GEMM (model.py:236), head 1x1 convs (model.py:105-119).
template <class T, class TO, bool FAST> struct EpiBiasActRes {
  static constexpr const char* kLabel = "bias_act_res";
  static constexpr bool kHeavy = false;   // tcgen05 path: 4 epilogue warps are enough
  struct Params {
    TO* out; int ldo;
    const float* bias;                 // [N] or null
    const float* cbias; int ldcb;      // [B][ldcb] or null (global-conditioning bias)
    int act;
    const T* res; int ldr;             // [rows][ldr] or null
    int N; int vec;
  };
  template <class Acc>
  static __device__ __forceinline__ void row(const Params& p, Acc& acc, int b, long long grow, int n0, int bn, int q, int nq) {
    for (int c = q * 16; c < bn; c += nq * 16) {
      const int n = n0 + c;
      if (n >= p.N) break;
      const int nv = acc.clip(min(16, p.N - n));
      float v[16];
      acc.load16(c, v);
      if (p.bias) {
#pragma unroll
        for (int i = 0; i < 16; ++i) if (i < nv) v[i] += p.bias[n + i];
      }
      if (p.cbias) {
        const float* cb = p.cbias + (long long)b * p.ldcb + n;
#pragma unroll
        for (int i = 0; i < 16; ++i) if (i < nv) v[i] += cb[i];
      }
      if (p.act != ACT_LINEAR) {
        wn_act16<FAST>(p.act, v);
      }
      if (p.res) {
        float r[16];
        load16<T>(p.res + grow * p.ldr + n, r, nv, p.vec);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += r[i];
      }
      store16<TO>(p.out + grow * p.ldo + n, v, nv, p.vec);
    }
  }
};

// Gated activation (layers.py:203-210).  The gated conv's weight columns are pre-interleaved at
// import so that a tile of width bn holds [filter ch0..ch0+bn/2 | gate ch0..ch0+bn/2), ch0 = n0/2.
//   z = acc + bias + cbias[b]   (Keras column order [filter D | gate D] in memory)
//   g = tanh(z_f) * sigmoid(z_s)
template <class T, bool FAST> struct EpiGate {
  static constexpr const char* kLabel = "gate";
  static constexpr bool kHeavy = true;    // 2 MUFU per output: 8 epilogue warps
  struct Params {
    T* z; T* g;                        // z [rows][2D], g [rows][ldg]
    int ldg;
    const float* bias;                 // [2D] Keras order
    const float* cbias;                // [B][2D] or null
    int D; int vec;
  };
  template <class Acc>
  static __device__ __forceinline__ void row(const Params& p, Acc& acc, int b, long long grow, int n0, int bn, int q, int nq) {
    const int half = bn >> 1;
    for (int c = q * 16; c < half; c += nq * 16) {
      const int ch = (n0 >> 1) + c;
      if (ch >= p.D) break;
      const int nv = acc.clip(min(16, p.D - ch));
      float f[16], s[16], g[16];
      acc.load16(c, f);
      acc.load16(half + c, s);
      const float* cb = p.cbias ? p.cbias + (long long)b * 2 * p.D : nullptr;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i < nv) {
          float zf = f[i] + p.bias[ch + i], zs = s[i] + p.bias[p.D + ch + i];
          if (cb) { zf += cb[ch + i]; zs += cb[p.D + ch + i]; }
          if constexpr (sizeof(T) == 2) {
            // bf16 tier: cache the gate's derivative coefficients P, Q instead of z (csrc/tc_epilogues.cuh: TcEpiGate)
            const float th = wn_tanh<FAST>(zf), sg = wn_sigmoid<FAST>(zs);
            g[i] = th * sg;
            f[i] = fmaf(-sg * th, th, sg);
            s[i] = fmaf(-g[i], sg, g[i]);
          } else {
            f[i] = zf; s[i] = zs;
            g[i] = wn_tanh<FAST>(zf) * wn_sigmoid<FAST>(zs);
          }
        } else {
          g[i] = 0.f;
        }
      }
      T* zrow = p.z + grow * 2 * p.D;
      store16<T>(zrow + ch, f, nv, p.vec);
      store16<T>(zrow + p.D + ch, s, nv, p.vec);
      store16<T>(p.g + grow * p.ldg + ch, g, nv, p.vec);
    }
  }
};

// Adjoint of the gate: acc = dg for channels [n0, n0+bn) ;
//   dz_f = dg * sig(z_s) * (1 - tanh(z_f)^2),  dz_s = dg * tanh(z_f) * sig(z_s) * (1 - sig(z_s))
template <class T, bool FAST> struct EpiGateBwd {
  static constexpr const char* kLabel = "gate_bwd";
  static constexpr bool kHeavy = true;
  struct Params {
    const T* z;                        // [rows][2D] cached pre-activations
    T* dz;                             // [rows][2D]
    int D; int vec;
  };
  template <class Acc>
  static __device__ __forceinline__ void row(const Params& p, Acc& acc, int b, long long grow, int n0, int bn, int q, int nq) {
    for (int c = q * 16; c < bn; c += nq * 16) {
      const int ch = n0 + c;
      if (ch >= p.D) break;
      const int nv = acc.clip(min(16, p.D - ch));
      float dg[16], f[16], s[16];
      acc.load16(c, dg);
      const T* zrow = p.z + grow * 2 * p.D;
      load16<T>(zrow + ch, f, nv, p.vec);
      load16<T>(zrow + p.D + ch, s, nv, p.vec);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if constexpr (sizeof(T) == 2) {
          f[i] = dg[i] * f[i];      // cached P, Q
          s[i] = dg[i] * s[i];
        } else {
          const float th = wn_tanh<FAST>(f[i]), sg = wn_sigmoid<FAST>(s[i]);
          f[i] = dg[i] * sg * (1.0f - th * th);
          s[i] = dg[i] * th * sg * (1.0f - sg);
        }
      }
      T* drow = p.dz + grow * 2 * p.D;
      store16<T>(drow + ch, f, nv, p.vec);
      store16<T>(drow + p.D + ch, s, nv, p.vec);
    }
  }
};

// Generic dgrad epilogue: out = (acc + add[row][n]) * act'(y[row][n])
// (residual pass-through of d x_out, and activation adjoint from the cached activation output)
template <class T, class TO> struct EpiActBwd {
  static constexpr const char* kLabel = "dgrad";
  static constexpr bool kHeavy = false;
  struct Params {
    TO* out; int ldo;
    const T* add; int lda;             // or null
    const T* y; int ldy; int act;      // or null / ACT_LINEAR
    int N; int vec;
  };
  template <class Acc>
  static __device__ __forceinline__ void row(const Params& p, Acc& acc, int b, long long grow, int n0, int bn, int q, int nq) {
    for (int c = q * 16; c < bn; c += nq * 16) {
      const int n = n0 + c;
      if (n >= p.N) break;
      const int nv = acc.clip(min(16, p.N - n));
      float v[16];
      acc.load16(c, v);
      if (p.add) {
        float a[16];
        load16<T>(p.add + grow * p.lda + n, a, nv, p.vec);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += a[i];
      }
      if (p.y && p.act != ACT_LINEAR) {
        float y[16];
        load16<T>(p.y + grow * p.ldy + n, y, nv, p.vec);
        wn_act_grad16(p.act, y, v);
      }
      store16<TO>(p.out + grow * p.ldo + n, v, nv, p.vec);
    }
  }
};
