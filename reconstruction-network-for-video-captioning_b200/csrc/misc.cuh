// Small data-movement kernels around the hot loops: operand staging (fp32 -> bf16 with K padding),
// embedding gather / scatter-add (models/decoder.py:46-48), bias-gradient column sums, the global
// reconstructor's mean-pool (models/global_reconstructor.py:33-37) and the multi-tensor L2-norm
// regulariser (train.py:69,101,127).  All HBM-bound, coalesced along the contiguous dimension.
#pragma once
#include "common.cuh"

namespace misc {

// dst[r, 0:cols] = (TO) src[r, 0:cols] ; dst[r, cols:cols_pad] = 0
template <typename TO>
__global__ void cast_pad_kernel(const float* __restrict__ src, long long ld_src, TO* __restrict__ dst, long long ld_dst,
                                long long rows, int cols, int cols_pad) {
  const long long total = rows * cols_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols_pad; const int c = (int)(i % cols_pad);
    dst[r * ld_dst + c] = from_f32<TO>(c < cols ? src[r * ld_src + c] : 0.f);
  }
}
template <typename TO>
static int cast_pad(const float* src, long long ld_src, TO* dst, long long ld_dst, long long rows, int cols, int cols_pad,
                    cudaStream_t st) {
  const long long total = rows * cols_pad;
  if (total == 0) return 0;
  int blocks = (int)min((long long)148 * 16, (total + 255) / 256);
  cast_pad_kernel<TO><<<blocks, 256, 0, st>>>(src, ld_src, dst, ld_dst, rows, cols, cols_pad);
  RN_LAUNCH_OK();
  return 0;
}

// Xe[r, c] = emb[tok[r], c] * scale * dropmask   (c < EMB), 0 for the K padding
template <typename TO>
__global__ void embed_gather_kernel(const float* __restrict__ emb, const long long* __restrict__ tok, TO* __restrict__ out,
                                    long long ld_out, int rows, int EMB, int EMBp, int V, float scale, float p_drop,
                                    const unsigned long long* rng, unsigned int site) {
  const int r = blockIdx.x;
  long long t = tok[r];
  if (t < 0 || t >= V) t = 0;
  for (int c = threadIdx.x; c < EMBp; c += blockDim.x) {
    float v = 0.f;
    if (c < EMB) {
      v = emb[t * EMB + c] * scale;
      if (p_drop > 0.f) v *= dropout_scale(rng, site, (uint64_t)r * EMB + c, p_drop);
    }
    out[(long long)r * ld_out + c] = from_f32<TO>(v);
  }
}
// dEmb[tok[r], c] += dXe[r, c] * scale * dropmask
__global__ void embed_scatter_kernel(float* __restrict__ demb, const long long* __restrict__ tok, const float* __restrict__ dxe,
                                     long long ld, int rows, int EMB, int V, float scale, float p_drop,
                                     const unsigned long long* rng, unsigned int site) {
  const int r = blockIdx.x;
  long long t = tok[r];
  if (t < 0 || t >= V) return;
  for (int c = threadIdx.x; c < EMB; c += blockDim.x) {
    float v = dxe[(long long)r * ld + c] * scale;
    if (p_drop > 0.f) v *= dropout_scale(rng, site, (uint64_t)r * EMB + c, p_drop);
    atomicAdd(demb + t * EMB + c, v);
  }
}

// out[n] (+)= sum_m X[m*ld + n]      block = 32 columns x 8 row-lanes
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ X, long long ld, int M, int N, float* __restrict__ out, int accumulate) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (n < N)
    for (int m = ty; m < M; m += 8) s += to_f32<T>(X[(long long)m * ld + n]);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][tx];
    out[n] = accumulate ? out[n] + t : t;
  }
}
template <typename T>
static int colsum(const T* X, long long ld, int M, int N, float* out, int accumulate, cudaStream_t st) {
  colsum_kernel<T><<<rn_cdiv(N, 32), 256, 0, st>>>(X, ld, M, N, out, accumulate);
  RN_LAUNCH_OK();
  return 0;
}

// mp[b,j] = (sum_l Hd[l,b,j]) * scale          (global reconstructor mean-pool, scale = cap / L^2 for one decoder layer)
__global__ void pool_time_kernel(const float* __restrict__ Hd, int L, long long n, float scale, float* __restrict__ mp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int l = 0; l < L; ++l) s += Hd[(long long)l * n + i];
  mp[i] = s * scale;
}
// dHd[l,b,j] (+)= dmp[b,j] * scale   for every l
__global__ void pool_time_bwd_kernel(const float* __restrict__ dmp, int L, long long n, float scale, float* __restrict__ dHd,
                                     int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = dmp[i] * scale;
  for (int l = 0; l < L; ++l) {
    float* p = dHd + (long long)l * n + i;
    *p = accumulate ? *p + g : g;
  }
}

// Global reconstructor operand rows: Xg[t,b,:] = [Hd[t,b,:], mp[b,:] * dropmask(t,b,:)]   (TO)
template <typename TO>
__global__ void global_x_kernel(const float* __restrict__ Hd, const float* __restrict__ mp, TO* __restrict__ X, int L, int B,
                                int H, float p_drop, const unsigned long long* rng, unsigned int site) {
  const long long total = (long long)L * B * 2 * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % (2 * H));
    const long long tb = i / (2 * H);
    const int b = (int)(tb % B);
    float v;
    if (c < H) v = Hd[tb * H + c];
    else {
      v = mp[(long long)b * H + (c - H)];
      if (p_drop > 0.f) v *= dropout_scale(rng, site, (uint64_t)(tb * H + (c - H)), p_drop);
    }
    X[i] = from_f32<TO>(v);
  }
}
// dmp[b,j] = sum_t dXg[t,b,H+j] * dropmask ; dHd[t,b,j] (+)= dXg[t,b,j]
__global__ void global_x_bwd_kernel(const float* __restrict__ dX, float* __restrict__ dHd, float* __restrict__ dmp, int L, int B,
                                    int H, int accumulate, float p_drop, const unsigned long long* rng, unsigned int site) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * H) return;
  const int b = (int)(i / H), j = (int)(i % H);
  float s = 0.f;
  for (int t = 0; t < L; ++t) {
    const long long tb = (long long)t * B + b;
    float g = dX[tb * 2 * H + H + j];
    if (p_drop > 0.f) g *= dropout_scale(rng, site, (uint64_t)(tb * H + j), p_drop);
    s += g;
    float* p = dHd + tb * H + j;
    const float d = dX[tb * 2 * H + j];
    *p = accumulate ? *p + d : d;
  }
  dmp[i] = s;
}

// ---- multi-tensor L2 norm regulariser: reg = sum_p ||p||_2 ; grad_p (+)= g * p / ||p|| -------------------------
constexpr int MT_CHUNK = 16384;
// table: ptrs[n] (device addresses), sizes[n]; blk_tensor[nb], blk_chunk[nb] map a block to a chunk of one tensor
__global__ void mt_sumsq_kernel(const long long* __restrict__ ptrs, const long long* __restrict__ sizes,
                                const int* __restrict__ blk_tensor, const int* __restrict__ blk_chunk, float* __restrict__ partial) {
  __shared__ float red[32];
  const int t = blk_tensor[blockIdx.x];
  const float* p = reinterpret_cast<const float*>(ptrs[t]);
  const long long n = sizes[t], lo = (long long)blk_chunk[blockIdx.x] * MT_CHUNK, hi = min(n, lo + MT_CHUNK);
  float s = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) { const float v = p[i]; s += v * v; }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;          // fixed-order two-stage reduction: bitwise reproducible
}
// one thread per tensor sums that tensor's block partials in block order, then reg = sum_t sqrt(sumsq[t])
__global__ void mt_norm_finalize_kernel(const float* __restrict__ partial, const int* __restrict__ blk_tensor, int n_blocks,
                                        float* __restrict__ sumsq, int n, float* __restrict__ out) {
  __shared__ float red[32];
  float r = 0.f;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < n_blocks; ++b) if (blk_tensor[b] == t) s += partial[b];
    sumsq[t] = s;
    r += sqrtf(s);
  }
  r = block_sum(r, red);
  if (threadIdx.x == 0) out[0] = r;
}
__global__ void mt_reg_grad_kernel(const long long* __restrict__ ptrs, const long long* __restrict__ gptrs,
                                   const long long* __restrict__ sizes, const int* __restrict__ blk_tensor,
                                   const int* __restrict__ blk_chunk, const float* __restrict__ sumsq,
                                   const float* __restrict__ gscale, float lambda, int accumulate) {
  const int t = blk_tensor[blockIdx.x];
  const float* p = reinterpret_cast<const float*>(ptrs[t]);
  float* g = reinterpret_cast<float*>(gptrs[t]);
  const long long n = sizes[t], lo = (long long)blk_chunk[blockIdx.x] * MT_CHUNK, hi = min(n, lo + MT_CHUNK);
  const float nrm = sqrtf(sumsq[t]);
  const float k = (nrm > 0.f) ? lambda * (gscale ? *gscale : 1.f) / nrm : 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) g[i] = (accumulate ? g[i] : 0.f) + k * p[i];
}
}  // namespace misc
