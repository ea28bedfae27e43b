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
// Vectorised path (V % 4 == 0, ld % 4 == 0, V <= 4 * CE_THREADS * CE_MAXG): every thread keeps its <= CE_MAXG float4 groups of
// the (dropout-scaled) row in registers -- one pass over memory, one Philox call per 4 logits (r1: the scalar three-pass
// version re-derived the mask per element in both passes and took 71 us for 52 MB).
constexpr int CE_MAXG = 6;
__global__ void __launch_bounds__(CE_THREADS) ce_fwd_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ target,
                                                             const float* __restrict__ weight, int V, float p_drop,
                                                             const unsigned long long* rng, unsigned int site,
                                                             float* __restrict__ lse, float* __restrict__ row_loss) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  const float w = weight[r];
  if (w == 0.f) { if (threadIdx.x == 0) { lse[r] = 0.f; row_loss[r] = 0.f; } return; }
  const float* z = logits + (long long)r * ld;
  float m = -INFINITY, s = 0.f;
  const int ng = V >> 2;
  if (!(V & 3) && !(ld & 3) && ng <= CE_THREADS * CE_MAXG) {
    float4 x[CE_MAXG];
#pragma unroll
    for (int k = 0; k < CE_MAXG; ++k) {
      const int g = min((int)threadIdx.x + k * CE_THREADS, ng - 1);
      x[k] = reinterpret_cast<const float4*>(z)[g];
    }
#pragma unroll
    for (int k = 0; k < CE_MAXG; ++k) {
      const int g = threadIdx.x + k * CE_THREADS;
      if (g < ng) {
        if (p_drop > 0.f) {
          const float4 d = dropout_scale4(rng, site, (uint64_t)r * V + 4 * g, p_drop);
          x[k].x *= d.x; x[k].y *= d.y; x[k].z *= d.z; x[k].w *= d.w;
        }
        m = fmaxf(m, fmaxf(fmaxf(x[k].x, x[k].y), fmaxf(x[k].z, x[k].w)));
      }
    }
    m = block_max(m, red);
#pragma unroll
    for (int k = 0; k < CE_MAXG; ++k) {
      const int g = threadIdx.x + k * CE_THREADS;
      if (g < ng) s += (expf(x[k].x - m) + expf(x[k].y - m)) + (expf(x[k].z - m) + expf(x[k].w - m));
    }
  } else {
    for (int v = threadIdx.x; v < V; v += CE_THREADS) {
      float x = z[v];
      if (p_drop > 0.f) x *= dropout_scale(rng, site, (uint64_t)r * V + v, p_drop);
      m = fmaxf(m, x);
    }
    m = block_max(m, red);
    for (int v = threadIdx.x; v < V; v += CE_THREADS) {
      float x = z[v];
      if (p_drop > 0.f) x *= dropout_scale(rng, site, (uint64_t)r * V + v, p_drop);
      s += expf(x - m);
    }
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
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(bf16* p, float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
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
  if (!(V & 3) && !(ld & 3) && !(ldd & 3) && !(Vp & 3)) {
    const int ng = V >> 2, ngp = Vp >> 2;
    for (int g4 = threadIdx.x; g4 < ngp; g4 += CE_THREADS) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g4 < ng) {
        const float4 x = reinterpret_cast<const float4*>(z)[g4];
        const float4 ds = dropout_scale4(rng, site, (uint64_t)r * V + 4 * g4, p_drop);
        const int v0 = 4 * g4;
        o.x = w * (expf(x.x * ds.x - l) - (v0 == t ? 1.f : 0.f)) * ds.x;
        o.y = w * (expf(x.y * ds.y - l) - (v0 + 1 == t ? 1.f : 0.f)) * ds.y;
        o.z = w * (expf(x.z * ds.z - l) - (v0 + 2 == t ? 1.f : 0.f)) * ds.z;
        o.w = w * (expf(x.w * ds.w - l) - (v0 + 3 == t ? 1.f : 0.f)) * ds.w;
      }
      store4(d + 4 * g4, o);
    }
    return;
  }
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
// local: sum over (t,b,r) (out[t,b,r] - feats[b,t,r])^2   -> per-block partials.  A block walks whole (t,b) rows, threads take
// float4 groups of r (no per-element index division; r1: the flat-index version ran at 1.3 TB/s).
__global__ void mse_local_fwd_kernel(const float* __restrict__ out, const float* __restrict__ feats, int Tn, int B, int R,
                                     float* __restrict__ partial) {
  __shared__ float red[32];
  float s = 0.f;
  const bool vec = !(R & 3);
  for (int tb = blockIdx.x; tb < Tn * B; tb += gridDim.x) {
    const int t = tb / B, b = tb - t * B;
    const float* o = out + (size_t)tb * R;
    const float* f = feats + ((size_t)b * Tn + t) * R;
    if (vec) {
      for (int r4 = threadIdx.x; r4 < (R >> 2); r4 += blockDim.x) {
        const float4 x = reinterpret_cast<const float4*>(o)[r4], y = reinterpret_cast<const float4*>(f)[r4];
        const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
        s += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
      }
    } else {
      for (int r = threadIdx.x; r < R; r += blockDim.x) { const float d = o[r] - f[r]; s += d * d; }
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// dOut[t,b,r] = g * 2/(N) * (out - feats)    (TO operand for the out-projection backward GEMMs)
template <typename TO>
__global__ void mse_local_bwd_kernel(const float* __restrict__ out, const float* __restrict__ feats, int Tn, int B, int R,
                                     const float* __restrict__ gscale, float k, TO* __restrict__ dout) {
  const float g = k * (gscale ? *gscale : 1.f);
  const bool vec = !(R & 3);
  for (int tb = blockIdx.x; tb < Tn * B; tb += gridDim.x) {
    const int t = tb / B, b = tb - t * B;
    const float* o = out + (size_t)tb * R;
    const float* f = feats + ((size_t)b * Tn + t) * R;
    TO* d = dout + (size_t)tb * R;
    if (vec) {
      for (int r4 = threadIdx.x; r4 < (R >> 2); r4 += blockDim.x) {
        const float4 x = reinterpret_cast<const float4*>(o)[r4], y = reinterpret_cast<const float4*>(f)[r4];
        store4(d + 4 * r4, make_float4(g * (x.x - y.x), g * (x.y - y.y), g * (x.z - y.z), g * (x.w - y.w)));
      }
    } else {
      for (int r = threadIdx.x; r < R; r += blockDim.x) d[r] = from_f32<TO>(g * (o[r] - f[r]));
    }
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
