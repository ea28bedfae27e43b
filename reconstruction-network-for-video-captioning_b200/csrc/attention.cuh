// Fused additive temporal "attention" of RecNet (reference models/decoder.py:50-62 and
// models/local_reconstructor.py:38-50):
//     e[b,tau] = w . tanh(Wh[b] + Uv[b,tau] + bias)          (NO softmax in the reference)
//     ctx[b,:] = (1/Tn) * sum_tau e[b,tau] * V[b,tau,:]       (mean over frames)
// with Uv = V U^T hoisted out of the time loop and Wh = h W^T arriving as split-K partials of the
// tensor-core GEMM.  HBM/L2-bound: V is streamed once with 16-byte coalesced loads, scores live in
// shared memory, the weighted sum accumulates in registers.  normalize = 1 adds a warp-shuffle softmax
// over frames (the paper's variant; not what the reference computes, never used for parity).
#pragma once
#include "common.cuh"

namespace attn {

constexpr int FWD_THREADS = 256;
constexpr int BWD_THREADS = 256;
constexpr int MAX_T = 64;       // frames (28 / 40) or decoder steps (<= 31)

// All three kernels are latency-bound (a few MB that live in L2, ~100 CTAs): every phase issues its loads in
// batches before the first use so that a thread has 8-28 independent requests in flight (r1 profile: the first
// version with serial per-frame loops took 13-20 us per launch).

struct FwdArgs {
  const float* WhP; int n_whp; long long whp_stride;   // Wh partials [n_whp][B,A]
  const float* Uv; long long uv_bs, uv_ts;             // Uv[b*uv_bs + tau*uv_ts + a]
  const float* attn_b; const float* attn_w;            // [A]
  const void* V; long long v_bs, v_ts;                 // V[b*v_bs + tau*v_ts + d]   (TV)
  int B, Tn, A, D; float inv_T; int normalize; int d_slice;   // d_slice: columns of D per blockIdx.x (set by launch_fwd)
  float* Wh_out;                                       // [B,A]   saved for backward (nullable)
  float* e_out;                                        // [B,Tn]  saved for backward (nullable)
  void* ctx_out; long long ctx_ld;                     // ctx[b*ctx_ld + d]   (TO) GEMM operand slot
  float p_drop; const unsigned long long* rng; unsigned int site; long long drop_base;  // dropout on ctx (train)
};

template <typename TV, typename TO>
__global__ void __launch_bounds__(FWD_THREADS) attn_fwd_kernel(FwdArgs a) {
  extern __shared__ float sm[];
  float* Wh = sm;              // [A] (bias folded in)
  float* wv = sm + a.A;        // [A] attn_w
  float* e = sm + 2 * a.A;     // [Tn]
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = FWD_THREADS / 32;
  for (int i = tid; i < a.A; i += FWD_THREADS) {
    float s = 0.f;
    for (int p0 = 0; p0 < a.n_whp; p0 += 8) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (p0 + k < a.n_whp) ? __ldg(a.WhP + (long long)(p0 + k) * a.whp_stride + (long long)b * a.A + i) : 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[k];
    }
    Wh[i] = s + a.attn_b[i];
    wv[i] = a.attn_w[i];
    if (a.Wh_out && blockIdx.x == 0) a.Wh_out[(long long)b * a.A + i] = s;
  }
  __syncthreads();
  // scores: warp w handles frames w, w+NW, ... ; up to 4 frames' Uv rows are loaded before any tanh
  for (int t0 = warp; t0 < a.Tn; t0 += 4 * NW) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = lane; i < a.A; i += 32) {
      float u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int tau = t0 + k * NW;
        u[k] = (tau < a.Tn) ? __ldg(a.Uv + (long long)b * a.uv_bs + (long long)tau * a.uv_ts + i) : 0.f;
      }
      const float whi = Wh[i], wi = wv[i];
#pragma unroll
      for (int k = 0; k < 4; ++k) s[k] += wi * tanhf(whi + u[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float r = warp_sum(s[k]);
      if (lane == 0 && t0 + k * NW < a.Tn) e[t0 + k * NW] = r;
    }
  }
  __syncthreads();
  if (a.normalize) {   // optional softmax over frames (paper variant)
    if (warp == 0) {
      float m = -INFINITY;
      for (int t = lane; t < a.Tn; t += 32) m = fmaxf(m, e[t]);
      m = warp_max(m);
      float z = 0.f;
      for (int t = lane; t < a.Tn; t += 32) z += __expf(e[t] - m);
      z = warp_sum(z);
      for (int t = lane; t < a.Tn; t += 32) e[t] = __expf(e[t] - m) / z;
    }
    __syncthreads();
  }
  if (a.e_out && blockIdx.x == 0)
    for (int t = tid; t < a.Tn; t += FWD_THREADS) a.e_out[(long long)b * a.Tn + t] = e[t];

  constexpr int VN = Vec16<TV>::N;
  const int d_lo = blockIdx.x * a.d_slice, d_hi = min(a.D, d_lo + a.d_slice);
  const TV* Vb = reinterpret_cast<const TV*>(a.V) + (long long)b * a.v_bs;
  TO* out = reinterpret_cast<TO*>(a.ctx_out) + (long long)b * a.ctx_ld;
  for (int d = d_lo + tid * VN; d < d_hi; d += FWD_THREADS * VN) {
    float acc[VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) acc[j] = 0.f;
    for (int t0 = 0; t0 < a.Tn; t0 += 8) {          // 8 frames (8 x 16 B) in flight per thread
      Vec16<TV> v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (t0 + k < a.Tn) v[k].load(Vb + (long long)(t0 + k) * a.v_ts + d);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (t0 + k < a.Tn) {
          float f[VN]; v[k].get(f);
          const float ek = e[t0 + k];
#pragma unroll
          for (int j = 0; j < VN; ++j) acc[j] += ek * f[j];
        }
      }
    }
    const float sc = a.normalize ? 1.f : a.inv_T;
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      acc[j] *= sc;
      if (a.p_drop > 0.f) acc[j] *= dropout_scale(a.rng, a.site, (uint64_t)(a.drop_base + (long long)b * a.D + d + j), a.p_drop);
      out[d + j] = from_f32<TO>(acc[j]);
    }
  }
}

// Backward.  One CTA per sample:
//   dctx = sum of split-K partials of d[ctx;h] (first D columns)      (x dropout mask in train)
//   de[tau] = inv_T * dctx . V[b,tau,:]
//   ds[tau,a] = de[tau] * w[a] * (1 - s^2),  s = tanh(Wh + Uv + b)   (recomputed, not stored)
//   dWh[b,a] = sum_tau ds ; dUv[b,tau,a] += ds ; dw_acc[b,a] += sum_tau de[tau]*s
struct BwdArgs {
  const float* dXp; int n_p; long long p_stride; long long p_ld;      // partials [n_p][B, p_ld], cols [0,D)
  const void* V; long long v_bs, v_ts;
  const float* Wh; const float* Uv; long long uv_bs, uv_ts;
  const float* attn_b; const float* attn_w;
  int B, Tn, A, D; float inv_T;
  float* dWh_out;                      // [B,A] fp32
  void* dWh_op;                        // [B,A] operand type (nullable): A-operand of the dWh @ attn_W GEMM
  float* dUv_acc; int uv_first;        // same strides as Uv; first => overwrite instead of +=
  float* dw_acc;                       // [B,A] += ; first => overwrite
  float* dctx_out;                     // [B,D] fp32 (nullable): summed/masked dctx, for the deferred dV pass
  float* de_out;                       // [B,Tn] (nullable)
  float p_drop; const unsigned long long* rng; unsigned int site; long long drop_base;
};

template <typename TV, typename TO>
__global__ void __launch_bounds__(BWD_THREADS) attn_bwd_kernel(BwdArgs a) {
  extern __shared__ float sm[];
  float* dctx = sm;                 // [D]
  float* de = sm + a.D;             // [Tn]
  float* red = de + a.Tn;           // [2 * BWD_THREADS]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = BWD_THREADS / 32;
  for (int d = tid; d < a.D; d += BWD_THREADS) {
    float s = 0.f;
    for (int p0 = 0; p0 < a.n_p; p0 += 8) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (p0 + k < a.n_p) ? __ldg(a.dXp + (long long)(p0 + k) * a.p_stride + (long long)b * a.p_ld + d) : 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[k];
    }
    if (a.p_drop > 0.f) s *= dropout_scale(a.rng, a.site, (uint64_t)(a.drop_base + (long long)b * a.D + d), a.p_drop);
    dctx[d] = s;
    if (a.dctx_out) a.dctx_out[(long long)b * a.D + d] = s;
  }
  __syncthreads();
  constexpr int VN = Vec16<TV>::N;
  const TV* Vb = reinterpret_cast<const TV*>(a.V) + (long long)b * a.v_bs;
  // de: warp w handles frames w, w+NW, ...; two frames' rows in flight
  for (int t0 = warp; t0 < a.Tn; t0 += 2 * NW) {
    float s[2] = {0.f, 0.f};
    for (int d = lane * VN; d < a.D; d += 32 * VN * 2) {
      Vec16<TV> v[2][2];
      bool ok[2][2];
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int tau = t0 + k * NW, dd = d + u * 32 * VN;
          ok[k][u] = tau < a.Tn && dd < a.D;
          if (ok[k][u]) v[k][u].load(Vb + (long long)tau * a.v_ts + dd);
        }
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (ok[k][u]) {
            float f[VN]; v[k][u].get(f);
            const int dd = d + u * 32 * VN;
#pragma unroll
            for (int j = 0; j < VN; ++j) s[k] += f[j] * dctx[dd + j];
          }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float r = warp_sum(s[k]) * a.inv_T;
      const int tau = t0 + k * NW;
      if (lane == 0 && tau < a.Tn) { de[tau] = r; if (a.de_out) a.de_out[(long long)b * a.Tn + tau] = r; }
    }
  }
  __syncthreads();
  // ds: thread (i, part) handles attention unit i and frames part, part+nparts, ... (4 frames in flight)
  const int nparts = (a.A <= BWD_THREADS && BWD_THREADS % a.A == 0) ? BWD_THREADS / a.A : 1;
  for (int i0 = 0; i0 < a.A; i0 += BWD_THREADS / nparts) {
    const int i = i0 + tid % (BWD_THREADS / nparts), part = tid / (BWD_THREADS / nparts);
    float dwh = 0.f, dw = 0.f;
    if (i < a.A) {
      const float wh = a.Wh[(long long)b * a.A + i] + a.attn_b[i];
      const float w = a.attn_w[i];
      for (int t0 = part; t0 < a.Tn; t0 += 4 * nparts) {
        float u[4], o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int tau = t0 + k * nparts;
          const long long off = (long long)b * a.uv_bs + (long long)tau * a.uv_ts + i;
          u[k] = (tau < a.Tn) ? __ldg(a.Uv + off) : 0.f;
          o[k] = (tau < a.Tn && !a.uv_first) ? a.dUv_acc[off] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int tau = t0 + k * nparts;
          if (tau < a.Tn) {
            const float s = tanhf(wh + u[k]);
            const float g = de[tau] * w * (1.f - s * s);
            dwh += g;
            dw += de[tau] * s;
            a.dUv_acc[(long long)b * a.uv_bs + (long long)tau * a.uv_ts + i] = o[k] + g;
          }
        }
      }
    }
    if (nparts > 1) {     // combine the frame partitions
      red[tid] = dwh; red[BWD_THREADS + tid] = dw;
      __syncthreads();
      if (part == 0 && i < a.A) {
        for (int q = 1; q < nparts; ++q) { dwh += red[q * (BWD_THREADS / nparts) + tid]; dw += red[BWD_THREADS + q * (BWD_THREADS / nparts) + tid]; }
      }
      __syncthreads();
    }
    if (part == 0 && i < a.A) {
      a.dWh_out[(long long)b * a.A + i] = dwh;
      if (a.dWh_op) reinterpret_cast<TO*>(a.dWh_op)[(long long)b * a.A + i] = from_f32<TO>(dwh);
      float* pw = a.dw_acc + (long long)b * a.A + i;
      *pw = a.uv_first ? dw : *pw + dw;
    }
  }
}

// Deferred value-gradient of the local reconstructor's attention (values = decoder hiddens, which need grad):
//   dV[l,b,:] (+)= inv_T * sum_t beta_t[b,l] * dx_t[b,:]        one pass after the time loop instead of Tsteps RMWs
// beta [S,B,Tn] fp32, dx [S,B,D] fp32, dV[b*dv_bs + l*dv_ts + d] fp32.
__global__ void attn_dv_kernel(const float* __restrict__ beta, const float* __restrict__ dx, float* __restrict__ dV,
                               long long dv_bs, long long dv_ts, int S, int B, int Tn, int D, float inv_T, int accumulate) {
  const int b = blockIdx.y, l = blockIdx.x;
  extern __shared__ float bt[];   // [S]
  for (int t = threadIdx.x; t < S; t += blockDim.x) bt[t] = beta[((long long)t * B + b) * Tn + l];
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < S; ++t) s += bt[t] * dx[((long long)t * B + b) * D + d];
    float* p = dV + (long long)b * dv_bs + (long long)l * dv_ts + d;
    s *= inv_T;
    *p = accumulate ? *p + s : s;
  }
}

template <typename TV, typename TO>
static int launch_fwd(FwdArgs a, cudaStream_t st) {
  if (a.Tn > MAX_T || a.Tn < 1) return RECNET_ERR_BAD_SHAPE;
  constexpr int VN = Vec16<TV>::N;
  if (a.D % VN || a.v_ts % VN || a.v_bs % VN) return RECNET_ERR_ALIGNMENT;
  // enough CTAs to cover the SMs (~150), each thread owning at most one 16-byte column group when possible
  int slices = rn_cdiv(150, a.B);
  const int min_slices = rn_cdiv(a.D, FWD_THREADS * VN);
  if (slices < min_slices) slices = min_slices;
  if (slices > a.D / VN) slices = a.D / VN;
  int slice = rn_cdiv(rn_cdiv(a.D, slices), VN) * VN;
  slices = rn_cdiv(a.D, slice);
  a.d_slice = slice;
  dim3 grid(slices, a.B);
  const size_t smem = (size_t)(2 * a.A + a.Tn) * sizeof(float);
  ProfScope prof(KC_ATTN_FWD, a.B, a.Tn, a.D, st);
  attn_fwd_kernel<TV, TO><<<grid, FWD_THREADS, smem, st>>>(a);
  RN_LAUNCH_OK();
  return 0;
}

template <typename TV, typename TO>
static int launch_bwd(const BwdArgs& a, cudaStream_t st) {
  if (a.Tn > MAX_T || a.Tn < 1) return RECNET_ERR_BAD_SHAPE;
  constexpr int VN = Vec16<TV>::N;
  if (a.D % VN || a.v_ts % VN || a.v_bs % VN) return RECNET_ERR_ALIGNMENT;
  const size_t smem = (size_t)(a.D + a.Tn + 2 * BWD_THREADS) * sizeof(float);
  ProfScope prof(KC_ATTN_BWD, a.B, a.Tn, a.D, st);
  attn_bwd_kernel<TV, TO><<<a.B, BWD_THREADS, smem, st>>>(a);
  RN_LAUNCH_OK();
  return 0;
}
}  // namespace attn
