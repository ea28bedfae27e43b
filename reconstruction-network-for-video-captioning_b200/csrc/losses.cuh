// Loss kernels of the RecNet step drivers.
//  * masked cross-entropy over the stacked teacher-forced logits (train.py:54-60,68): per row r=(t,b)
//      loss += weight[r] * (logsumexp(z_r) - z_r[target_r]),   weight[r] = mask / (n_t * sum_t n_t)
//    with the reference's train-mode dropout on the LOGITS (models/decoder.py:69) applied in-kernel.
//  * reconstruction MSE, local (train.py:125-128) and global (train.py:96-100) flavours.
#pragma once
#include "common.cuh"

namespace loss {

constexpr int CE_THREADS = 256;

// forward: lse[r], row_loss[r] (unweighted NLL); loss reduced deterministically by ce_reduce_kernel.
__global__ void __launch_bounds__(CE_THREADS) ce_fwd_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ target,
                                                             const float* __restrict__ weight, int V, float p_drop,
                                                             const unsigned long long* rng, unsigned int site,
                                                             float* __restrict__ lse, float* __restrict__ row_loss) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  const float w = weight[r];
  if (w == 0.f) { if (threadIdx.x == 0) { lse[r] = 0.f; row_loss[r] = 0.f; } return; }
  const float* z = logits + (long long)r * ld;
  float m = -INFINITY;
  for (int v = threadIdx.x; v < V; v += CE_THREADS) {
    float x = z[v];
    if (p_drop > 0.f) x *= dropout_scale(rng, site, (uint64_t)r * V + v, p_drop);
    m = fmaxf(m, x);
  }
  m = block_max(m, red);
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += CE_THREADS) {
    float x = z[v];
    if (p_drop > 0.f) x *= dropout_scale(rng, site, (uint64_t)r * V + v, p_drop);
    s += expf(x - m);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const long long t = target[r];
    float zt = z[t];
    if (p_drop > 0.f) zt *= dropout_scale(rng, site, (uint64_t)r * V + t, p_drop);
    const float l = m + logf(s);
    lse[r] = l;
    row_loss[r] = w * (l - zt);
  }
}
__global__ void sum_kernel(const float* __restrict__ x, int n, float* __restrict__ out, float scale) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s * scale;
}
// backward: dlogits[r,v] = g * weight[r] * (softmax(z_r)[v] - 1[v==target]) (x dropout scale), cols [V,Vp) zeroed
template <typename TO>
__global__ void __launch_bounds__(CE_THREADS) ce_bwd_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ target,
                                                             const float* __restrict__ weight, const float* __restrict__ lse,
                                                             const float* __restrict__ gscale, int V, int Vp, float p_drop,
                                                             const unsigned long long* rng, unsigned int site,
                                                             TO* __restrict__ dlogits, long long ldd) {
  const int r = blockIdx.x;
  const float w = weight[r] * (gscale ? *gscale : 1.f);
  TO* d = dlogits + (long long)r * ldd;
  if (weight[r] == 0.f) {
    for (int v = threadIdx.x; v < Vp; v += CE_THREADS) d[v] = from_f32<TO>(0.f);
    return;
  }
  const float* z = logits + (long long)r * ld;
  const float l = lse[r];
  const long long t = target[r];
  for (int v = threadIdx.x; v < Vp; v += CE_THREADS) {
    float g = 0.f;
    if (v < V) {
      float x = z[v], ds = 1.f;
      if (p_drop > 0.f) { ds = dropout_scale(rng, site, (uint64_t)r * V + v, p_drop); x *= ds; }
      g = w * (expf(x - l) - (v == t ? 1.f : 0.f)) * ds;
    }
    d[v] = from_f32<TO>(g);
  }
}

// ---- MSE -------------------------------------------------------------------------------------------------------
// local: sum over (t,b,r) (out[t,b,r] - feats[b,t,r])^2   -> per-block partials
__global__ void mse_local_fwd_kernel(const float* __restrict__ out, const float* __restrict__ feats, int Tn, int B, int R,
                                     float* __restrict__ partial) {
  __shared__ float red[32];
  const long long total = (long long)Tn * B * R;
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i % R); const long long tb = i / R; const int b = (int)(tb % B); const int t = (int)(tb / B);
    const float d = out[i] - feats[((long long)b * Tn + t) * R + r];
    s += d * d;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// dOut[t,b,r] = g * 2/(N) * (out - feats)    (TO operand for the out-projection backward GEMMs)
template <typename TO>
__global__ void mse_local_bwd_kernel(const float* __restrict__ out, const float* __restrict__ feats, int Tn, int B, int R,
                                     const float* __restrict__ gscale, float k, TO* __restrict__ dout) {
  const long long total = (long long)Tn * B * R;
  const float g = k * (gscale ? *gscale : 1.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i % R); const long long tb = i / R; const int b = (int)(tb % B); const int t = (int)(tb / B);
    dout[i] = from_f32<TO>(g * (out[i] - feats[((long long)b * Tn + t) * R + r]));
  }
}
// global: diff[b,r] = mean_t out[t,b,r] - mean_tau feats[b,tau,r] ; partial sums of diff^2
__global__ void mse_global_diff_kernel(const float* __restrict__ out, int L, const float* __restrict__ feats, int Tn, int B, int R,
                                       float* __restrict__ diff, float* __restrict__ partial) {
  __shared__ float red[32];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float sq = 0.f;
  if (i < (long long)B * R) {
    const int b = (int)(i / R), r = (int)(i % R);
    float so = 0.f, sf = 0.f;
    for (int t = 0; t < L; ++t) so += out[((long long)t * B + b) * R + r];
    for (int t = 0; t < Tn; ++t) sf += feats[((long long)b * Tn + t) * R + r];
    const float d = so / L - sf / Tn;
    diff[i] = d;
    sq = d * d;
  }
  sq = block_sum(sq, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = sq;
}
// dOut[t,b,r] = g * k * diff[b,r]    for every t
template <typename TO>
__global__ void mse_global_bwd_kernel(const float* __restrict__ diff, int L, int B, int R, const float* __restrict__ gscale, float k,
                                      TO* __restrict__ dout) {
  const long long total = (long long)L * B * R;
  const float g = k * (gscale ? *gscale : 1.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    dout[i] = from_f32<TO>(g * diff[i % ((long long)B * R)]);
}
}  // namespace loss
