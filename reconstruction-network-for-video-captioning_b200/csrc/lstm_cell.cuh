// Fused LSTM gate-activation / cell-update kernels (reference: nn.LSTM step at models/decoder.py:66,
// models/global_reconstructor.py:43, models/local_reconstructor.py:52; PyTorch gate order i,f,g,o).
//
// Forward:  pre = sum_s P[s] (split-K partials of [x;h] W_rec^T) + Gx (hoisted input projection) + b_ih + b_hh
//           i,f,o = sigmoid, g = tanh ; c' = f*c + i*g ; h' = o*tanh(c')
// The activated gates are stashed (TS = float or bf16) for BPTT; h' is written both as fp32 (returned
// hiddens) and in the GEMM operand type straight into next step's [x ; h] operand row.
//
// These kernels move a few MB that sit in L2, so they are latency-bound, not bandwidth-bound: one hidden unit
// per thread (B*H threads, coalesced 4-byte accesses) and all split-K partial loads of a thread are issued
// before the first use (16 independent loads in flight per thread) -- r1 profile: the previous 4-units-per-thread
// version with a serial partial loop took 13 us for 6.5 MB.
#pragma once
#include "common.cuh"

namespace cell {
constexpr int THREADS = 256;

struct FwdArgs {
  const float* P; int n_p; long long p_stride; long long p_ld;   // partials [n_p][B, p_ld(=4H)]
  const float* Gx; long long gx_ld;                               // [B,4H] nullable
  const float* b1; const float* b2;                               // [4H] nullable each
  const float* c_prev;                                            // [B,H]
  int B, H;
  void* gates_out;                                                // [B,4H] TS stash (nullable in inference)
  float* c_out;                                                   // [B,H]
  float* h_out; long long h_ld;                                   // fp32 h' (nullable)
  void* h_op; long long hop_ld;                                   // operand-typed h' (nullable)
  void* h_op2; long long hop2_ld;                                 // second operand-typed copy (nullable)
};

// sum of n partial buffers at `off`, 4 loads in flight per call site x 4 gates
__device__ __forceinline__ void sum_partials4(const float* __restrict__ P, int n_p, long long stride, long long off, int H,
                                              float& s0, float& s1, float& s2, float& s3) {
  for (int p0 = 0; p0 < n_p; p0 += 4) {
    float v[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool ok = p0 + k < n_p;
      const float* q = P + (long long)(p0 + k) * stride + off;
      v[k][0] = ok ? __ldg(q) : 0.f;
      v[k][1] = ok ? __ldg(q + H) : 0.f;
      v[k][2] = ok ? __ldg(q + 2 * H) : 0.f;
      v[k][3] = ok ? __ldg(q + 3 * H) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { s0 += v[k][0]; s1 += v[k][1]; s2 += v[k][2]; s3 += v[k][3]; }
  }
}

template <typename TS, typename TO>
__global__ void __launch_bounds__(THREADS) lstm_cell_fwd_kernel(FwdArgs a) {
  const long long idx = (long long)blockIdx.x * THREADS + threadIdx.x;
  if (idx >= (long long)a.B * a.H) return;
  const int b = (int)(idx / a.H), j = (int)(idx % a.H);
  const int H = a.H;
  float pi = 0.f, pf = 0.f, pg = 0.f, po = 0.f;
  const float cp = a.c_prev[(long long)b * H + j];
  if (a.Gx) { const float* g = a.Gx + (long long)b * a.gx_ld + j; pi += g[0]; pf += g[H]; pg += g[2 * H]; po += g[3 * H]; }
  if (a.b1) { pi += a.b1[j]; pf += a.b1[H + j]; pg += a.b1[2 * H + j]; po += a.b1[3 * H + j]; }
  if (a.b2) { pi += a.b2[j]; pf += a.b2[H + j]; pg += a.b2[2 * H + j]; po += a.b2[3 * H + j]; }
  sum_partials4(a.P, a.n_p, a.p_stride, (long long)b * a.p_ld + j, H, pi, pf, pg, po);
  const float gi = sigmoidf_(pi), gf = sigmoidf_(pf), gg = tanhf(pg), go = sigmoidf_(po);
  const float cn = gf * cp + gi * gg;
  const float hn = go * tanhf(cn);
  a.c_out[(long long)b * H + j] = cn;
  if (a.h_out) a.h_out[(long long)b * a.h_ld + j] = hn;
  if (a.gates_out) {
    TS* g = reinterpret_cast<TS*>(a.gates_out) + (long long)b * 4 * H + j;
    g[0] = from_f32<TS>(gi); g[H] = from_f32<TS>(gf); g[2 * H] = from_f32<TS>(gg); g[3 * H] = from_f32<TS>(go);
  }
  if (a.h_op) reinterpret_cast<TO*>(a.h_op)[(long long)b * a.hop_ld + j] = from_f32<TO>(hn);
  if (a.h_op2) reinterpret_cast<TO*>(a.h_op2)[(long long)b * a.hop2_ld + j] = from_f32<TO>(hn);
}

// Backward of one step.
//   dh = dh_scale*dh_ext + dh_ext2 + sum_s dXp[s][:, col0 + j] + sum_s dQp[s][:, j]
// dXp = split-K partials of d[x;h] = dG_{t+1} [W_x | W_hh] (recurrent path); dQp = split-K partials of
// dWh_{t+1} @ attn_W (attention-query path, its own small tensor-core GEMM).
struct BwdArgs {
  const float* dh_ext; long long dh_ld; const float* dh_scale;    // [B,H] nullable; optional device scalar multiplier
  const float* dh_ext2; long long dh2_ld;                          // second external term (nullable)
  const float* dXp; int n_p; long long p_stride; long long p_ld; int col0;   // nullable
  const float* dQp; int n_q; long long q_stride; long long q_ld;   // nullable
  float* dc;                                                       // [B,H] in/out (dc_next -> dc_prev); first => treated as 0
  int first;
  const void* gates;                                               // [B,4H] TS
  const float* c_prev; const float* c_new;                         // [B,H]
  int B, H;
  void* dG; long long dg_ld;                                       // [B,4H] TO (GEMM operand)
};

__device__ __forceinline__ float sum_partials1(const float* __restrict__ P, int n_p, long long stride, long long off) {
  float s = 0.f;
  for (int p0 = 0; p0 < n_p; p0 += 8) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (p0 + k < n_p) ? __ldg(P + (long long)(p0 + k) * stride + off) : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
  }
  return s;
}

template <typename TS, typename TO>
__global__ void __launch_bounds__(THREADS) lstm_cell_bwd_kernel(BwdArgs a) {
  const long long idx = (long long)blockIdx.x * THREADS + threadIdx.x;
  if (idx >= (long long)a.B * a.H) return;
  const int b = (int)(idx / a.H), j = (int)(idx % a.H);
  const int H = a.H;
  const TS* g = reinterpret_cast<const TS*>(a.gates) + (long long)b * 4 * H + j;
  const float gi = to_f32<TS>(g[0]), gf = to_f32<TS>(g[H]), gg = to_f32<TS>(g[2 * H]), go = to_f32<TS>(g[3 * H]);
  const float cp = a.c_prev[(long long)b * H + j], cn = a.c_new[(long long)b * H + j];
  const float dcn = a.first ? 0.f : a.dc[(long long)b * H + j];
  float dh = 0.f;
  if (a.dh_ext) dh += (a.dh_scale ? *a.dh_scale : 1.f) * a.dh_ext[(long long)b * a.dh_ld + j];
  if (a.dh_ext2) dh += a.dh_ext2[(long long)b * a.dh2_ld + j];
  if (a.dXp) dh += sum_partials1(a.dXp, a.n_p, a.p_stride, (long long)b * a.p_ld + a.col0 + j);
  if (a.dQp) dh += sum_partials1(a.dQp, a.n_q, a.q_stride, (long long)b * a.q_ld + j);
  const float tc = tanhf(cn);
  const float dc = dcn + dh * go * (1.f - tc * tc);
  a.dc[(long long)b * H + j] = dc * gf;
  TO* o = reinterpret_cast<TO*>(a.dG) + (long long)b * a.dg_ld + j;
  o[0] = from_f32<TO>(dc * gg * gi * (1.f - gi));
  o[H] = from_f32<TO>(dc * cp * gf * (1.f - gf));
  o[2 * H] = from_f32<TO>(dc * gi * (1.f - gg * gg));
  o[3 * H] = from_f32<TO>(dh * tc * go * (1.f - go));
}

template <typename TS, typename TO>
static int launch_fwd(const FwdArgs& a, cudaStream_t st) {
  const long long n = (long long)a.B * a.H;
  ProfScope prof(KC_CELL_FWD, a.B, a.H, a.n_p, st);
  lstm_cell_fwd_kernel<TS, TO><<<rn_cdiv(n, THREADS), THREADS, 0, st>>>(a);
  RN_LAUNCH_OK();
  return 0;
}
template <typename TS, typename TO>
static int launch_bwd(const BwdArgs& a, cudaStream_t st) {
  const long long n = (long long)a.B * a.H;
  ProfScope prof(KC_CELL_BWD, a.B, a.H, (a.dXp ? a.n_p : 0) + (a.dQp ? a.n_q : 0), st);
  lstm_cell_bwd_kernel<TS, TO><<<rn_cdiv(n, THREADS), THREADS, 0, st>>>(a);
  RN_LAUNCH_OK();
  return 0;
}
}  // namespace cell
