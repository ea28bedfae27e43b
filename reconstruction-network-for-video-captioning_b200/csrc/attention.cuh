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

constexpr int FWD_THREADS = 128;
constexpr int BWD_THREADS = 256;
constexpr int MAX_T = 64;       // frames (28 / 40) or decoder steps (<= 31)

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
  float* e = sm + a.A;         // [Tn]
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < a.A; i += FWD_THREADS) {
    float s = 0.f;
    for (int p = 0; p < a.n_whp; ++p) s += a.WhP[p * a.whp_stride + (long long)b * a.A + i];
    Wh[i] = s + a.attn_b[i];
    if (a.Wh_out && blockIdx.x == 0) a.Wh_out[(long long)b * a.A + i] = s;
  }
  __syncthreads();
  for (int tau = warp; tau < a.Tn; tau += FWD_THREADS / 32) {
    const float* uv = a.Uv + (long long)b * a.uv_bs + (long long)tau * a.uv_ts;
    float s = 0.f;
    for (int i = lane; i < a.A; i += 32) s += a.attn_w[i] * tanhf(Wh[i] + uv[i]);
    s = warp_sum(s);
    if (lane == 0) e[tau] = s;
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
    int tau = 0;
    for (; tau + 4 <= a.Tn; tau += 4) {
      Vec16<TV> v0, v1, v2, v3;
      v0.load(Vb + (long long)(tau + 0) * a.v_ts + d);
      v1.load(Vb + (long long)(tau + 1) * a.v_ts + d);
      v2.load(Vb + (long long)(tau + 2) * a.v_ts + d);
      v3.load(Vb + (long long)(tau + 3) * a.v_ts + d);
      float f0[VN], f1[VN], f2[VN], f3[VN];
      v0.get(f0); v1.get(f1); v2.get(f2); v3.get(f3);
      const float e0 = e[tau], e1 = e[tau + 1], e2 = e[tau + 2], e3 = e[tau + 3];
#pragma unroll
      for (int j = 0; j < VN; ++j) acc[j] += e0 * f0[j] + e1 * f1[j] + e2 * f2[j] + e3 * f3[j];
    }
    for (; tau < a.Tn; ++tau) {
      Vec16<TV> v0; v0.load(Vb + (long long)tau * a.v_ts + d);
      float f0[VN]; v0.get(f0);
      const float e0 = e[tau];
#pragma unroll
      for (int j = 0; j < VN; ++j) acc[j] += e0 * f0[j];
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
  float* dWh_out;                      // [B,A]
  float* dUv_acc; int uv_first;        // same strides as Uv; first => overwrite instead of +=
  float* dw_acc;                       // [B,A] += ; first => overwrite
  float* dctx_out;                     // [B,D] fp32 (nullable): summed/masked dctx, for the deferred dV pass
  float* de_out;                       // [B,Tn] (nullable)
  float p_drop; const unsigned long long* rng; unsigned int site; long long drop_base;
};

template <typename TV>
__global__ void __launch_bounds__(BWD_THREADS) attn_bwd_kernel(BwdArgs a) {
  extern __shared__ float sm[];
  float* dctx = sm;                 // [D]
  float* de = sm + a.D;             // [Tn]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int d = tid; d < a.D; d += BWD_THREADS) {
    float s = 0.f;
    for (int p = 0; p < a.n_p; ++p) s += a.dXp[p * a.p_stride + (long long)b * a.p_ld + d];
    if (a.p_drop > 0.f) s *= dropout_scale(a.rng, a.site, (uint64_t)(a.drop_base + (long long)b * a.D + d), a.p_drop);
    dctx[d] = s;
    if (a.dctx_out) a.dctx_out[(long long)b * a.D + d] = s;
  }
  __syncthreads();
  constexpr int VN = Vec16<TV>::N;
  const TV* Vb = reinterpret_cast<const TV*>(a.V) + (long long)b * a.v_bs;
  for (int tau = warp; tau < a.Tn; tau += BWD_THREADS / 32) {
    const TV* vr = Vb + (long long)tau * a.v_ts;
    float s = 0.f;
    for (int d = lane * VN; d < a.D; d += 32 * VN) {
      Vec16<TV> v; v.load(vr + d);
      float f[VN]; v.get(f);
#pragma unroll
      for (int j = 0; j < VN; ++j) s += f[j] * dctx[d + j];
    }
    s = warp_sum(s);
    if (lane == 0) { de[tau] = s * a.inv_T; if (a.de_out) a.de_out[(long long)b * a.Tn + tau] = s * a.inv_T; }
  }
  __syncthreads();
  for (int i = tid; i < a.A; i += BWD_THREADS) {
    const float wh = a.Wh[(long long)b * a.A + i] + a.attn_b[i];
    const float w = a.attn_w[i];
    float dwh = 0.f, dw = 0.f;
    for (int tau = 0; tau < a.Tn; ++tau) {
      const long long off = (long long)b * a.uv_bs + (long long)tau * a.uv_ts + i;
      const float s = tanhf(wh + a.Uv[off]);
      const float g = de[tau] * w * (1.f - s * s);
      dwh += g;
      dw += de[tau] * s;
      a.dUv_acc[off] = a.uv_first ? g : a.dUv_acc[off] + g;
    }
    a.dWh_out[(long long)b * a.A + i] = dwh;
    float* pw = a.dw_acc + (long long)b * a.A + i;
    *pw = a.uv_first ? dw : *pw + dw;
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
  if (a.D % VN || a.v_ts % VN || a.v_bs % VN || (sizeof(TO) == 2 && (a.ctx_ld % 2))) return RECNET_ERR_ALIGNMENT;
  int slices = rn_cdiv(a.D, FWD_THREADS * VN);
  // keep slice boundaries vector aligned
  int slice = rn_cdiv(rn_cdiv(a.D, slices), VN) * VN;
  slices = rn_cdiv(a.D, slice);
  a.d_slice = slice;
  dim3 grid(slices, a.B);
  const size_t smem = (size_t)(a.A + a.Tn) * sizeof(float);
  ProfScope prof(KC_ATTN_FWD, a.B, a.Tn, a.D, st);
  attn_fwd_kernel<TV, TO><<<grid, FWD_THREADS, smem, st>>>(a);
  RN_LAUNCH_OK();
  return 0;
}

template <typename TV>
static int launch_bwd(const BwdArgs& a, cudaStream_t st) {
  if (a.Tn > MAX_T || a.Tn < 1) return RECNET_ERR_BAD_SHAPE;
  constexpr int VN = Vec16<TV>::N;
  if (a.D % VN || a.v_ts % VN || a.v_bs % VN) return RECNET_ERR_ALIGNMENT;
  const size_t smem = (size_t)(a.D + a.Tn) * sizeof(float);
  ProfScope prof(KC_ATTN_BWD, a.B, a.Tn, a.D, st);
  attn_bwd_kernel<TV><<<a.B, BWD_THREADS, smem, st>>>(a);
  RN_LAUNCH_OK();
  return 0;
}
}  // namespace attn
