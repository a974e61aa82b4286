// generate.cuh — autoregressive generation with per-conv input histories ("fast WaveNet" single-step form,
// layers.py:226-290; driver loop model.py:258-307).  fp32 arithmetic on the fp32 master weights (Keras layouts).
//
// Every dilated conv keeps the history of ITS input, hist[b][time][c]; producing output sample t+1 costs one K-tap
// matrix-vector product per conv (taps at t - (K-1-k)*d) instead of a forward pass over the whole receptive field.
// One step = a short chain of small kernels replayed as a CUDA graph; the time index lives in device memory so the same
// graph serves every step.  Equivalent to the reference's sliding-window loop: the window is exactly one receptive
// field long, so its causal zero padding never reaches the last position (model.py:122).
#pragma once
#include "common.cuh"

struct GenVec {           // one K-tap dense layer evaluated at time t for every batch row
  const float* in; long long in_bstride; int in_tstride;   // input history (B, cap, Cin) (in_tstride = Cin) or a plain (B, Cin) vector (in_tstride = 0)
  int Cin, K, dil;
  int in_gate;            // 1: the input vector is the gate of a (B, 2*Cin) pre-activation: tanh(z[c]) * sigmoid(z[Cin + c])
  const float* W;         // [K][Cin][N] (Keras kernel layout)
  const float* bias;      // [N] or null
  const float* cbias; int ldcb;   // per-batch additive term [B][ldcb] or null
  int N, act;
  float* out; long long out_bstride; int out_tstride;      // destination: history (written at time t) or plain vector
  const float* res; long long res_bstride; int res_tstride; int res_cols;   // residual added to columns [0, res_cols)
  float* acc; int acc_ld; int acc_col0;                    // columns [acc_col0, N) are ACCUMULATED into acc[b][n - acc_col0] (skip sum)
};

// grid (ceil(N / 32), ceil(B / GEN_ROWS)), 256 threads: a CTA computes 32 consecutive outputs for GEN_ROWS batch rows (every
// weight fetched once per GEN_ROWS rows; measured: 1 row per CTA is fastest, the step is latency- not traffic-bound);
// warp w sums the (k, c) pairs congruent to w mod 8, then warp r finishes row r
#define GEN_ROWS 1
__global__ void __launch_bounds__(256) gen_dense_kernel(const GenVec v, const int B, const int* __restrict__ t_dev) {
  extern __shared__ float xs[];              // [GEN_ROWS][K*Cin] staged input taps
  __shared__ float red[8][GEN_ROWS][33];
  const int t = *t_dev;
  const int b0 = blockIdx.y * GEN_ROWS;
  const int kc = v.K * v.Cin;
  for (int i = threadIdx.x; i < GEN_ROWS * kc; i += 256) {
    const int r = i / kc, j = i % kc;
    const int k = j / v.Cin, c = j % v.Cin;
    const int ts = t - (v.K - 1 - k) * v.dil;
    const int b = b0 + r;
    float x = 0.f;
    if (b < B && (ts >= 0 || v.in_tstride == 0)) {
      const float* p = v.in + (long long)b * v.in_bstride + (v.in_tstride ? (long long)ts * v.in_tstride : 0);   // gate inputs are plain vectors
      if (v.in_gate) x = tanhf(p[c]) * (1.0f / (1.0f + expf(-p[v.Cin + c])));
      else x = p[c];
    }
    xs[i] = x;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 32 + lane;
  float s[GEN_ROWS];
#pragma unroll
  for (int r = 0; r < GEN_ROWS; ++r) s[r] = 0.f;
  if (n < v.N) {
    for (int i = warp; i < kc; i += 8) {
      const float w = v.W[(long long)i * v.N + n];
#pragma unroll
      for (int r = 0; r < GEN_ROWS; ++r) s[r] = fmaf(w, xs[r * kc + i], s[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < GEN_ROWS; ++r) red[warp][r][lane] = s[r];
  __syncthreads();
  const int b = b0 + warp;                   // warp r finishes batch row b0 + r
  if (warp < GEN_ROWS && b < B && n < v.N) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) r += red[w][warp][lane];
    if (v.bias) r += v.bias[n];
    if (v.cbias) r += v.cbias[(long long)b * v.ldcb + n];
    r = wn_act<false>(v.act, r);
    if (v.acc && v.acc_col0 >= 0 && n >= v.acc_col0) {
      v.acc[(long long)b * v.acc_ld + (n - v.acc_col0)] += r;
    } else {
      float o = r;
      if (v.res && n < v.res_cols) o += v.res[(long long)b * v.res_bstride + (long long)t * v.res_tstride + n];
      // aliased skip (skip_channels=None): the skip sum takes conv1's output BEFORE the residual add (layers.py:216-223)
      if (v.acc && v.acc_col0 < 0) v.acc[(long long)b * v.acc_ld + n] += r;
      v.out[(long long)b * v.out_bstride + (long long)t * v.out_tstride + n] = o;
    }
  }
}

// h0[b][t][c] = sum_k W[k][0][c] * audio[b][t - (K-1-k)] + bias[c]   (model.py:84-88)
__global__ void gen_input_conv_kernel(const float* __restrict__ audio, int cap, const float* __restrict__ W, const float* __restrict__ bias,
                                      float* __restrict__ h0, int R, int K, int B, const int* __restrict__ t_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * R) return;
  const int t = *t_dev;
  const int b = i / R, c = i % R;
  float s = bias[c];
  for (int k = 0; k < K; ++k) {
    const int ts = t - (K - 1 - k);
    if (ts >= 0) s = fmaf(W[k * R + c], audio[(long long)b * cap + ts], s);
  }
  h0[((long long)b * cap + t) * R + c] = s;
}
__global__ void gen_zero_kernel(float* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}
// end of a step: the sampled value becomes audio[t+1] (unless teacher forced), predictions are optionally kept, t advances
__global__ void gen_advance_kernel(float* __restrict__ audio, int cap, const float* __restrict__ sampled, int B, int write_audio, int* t_dev) {
  const int b = threadIdx.x;
  const int t = *t_dev;
  if (b < B && write_audio && t + 1 < cap) audio[(long long)b * cap + t + 1] = sampled[b];
  __syncthreads();
  if (threadIdx.x == 0) *t_dev = t + 1;
}
// softmax over C logits per row (B rows), or a plain copy for mixture parameters: pred_out[b][step][:]
__global__ void gen_pred_kernel(const float* __restrict__ logits, int C, int softmax, float* __restrict__ pred, long long pred_bstride, int t0,
                                const int* __restrict__ t_dev) {
  __shared__ float sh[32];
  const int b = blockIdx.x;
  const int step = *t_dev - t0;
  const float* lp = logits + (long long)b * C;
  float* o = pred + (long long)b * pred_bstride + (long long)step * C;
  if (!softmax) { for (int c = threadIdx.x; c < C; c += blockDim.x) o[c] = lp[c]; return; }
  float m = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) m = fmaxf(m, lp[c]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  m = sh[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, sh[w]);
  __syncthreads();
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s += expf(lp[c] - m);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
  const float inv = 1.0f / s;
  for (int c = threadIdx.x; c < C; c += blockDim.x) o[c] = expf(lp[c] - m) * inv;
}
