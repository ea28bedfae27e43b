// Fused GRU gate-activation / state-update kernels (reference: nn.GRU step at models/decoder.py:32-40,66 when
// model_name != "LSTM" -- the reference's DEFAULT decoder, config.py:31; PyTorch gate order r,z,n):
//     r = sigmoid(gi_r + gh_r)   z = sigmoid(gi_z + gh_z)   n = tanh(gi_n + r * gh_n)   h' = (1 - z) * n + z * h
// gi = x W_ih^T + b_ih and gh = h W_hh^T + b_hh must stay SEPARATE for the n gate, so the step feeds two split-K
// partial sets (x-part GEMM over the [ctx] columns, h-part GEMM over the [h] columns of the same operand rows).
// Stash for BPTT: r, z, n and gh_n (4 slots per unit, same footprint as the LSTM stash).
#pragma once
#include "common.cuh"

namespace gru {
constexpr int THREADS = 256;

struct FwdArgs {
  const float* Px; int n_px; long long px_stride; long long px_ld;   // x-part partials [n_px][B, 3H] (nullable / 0)
  const float* Ph; int n_ph; long long ph_stride; long long ph_ld;   // h-part partials [n_ph][B, 3H] (nullable / 0)
  const float* Gx; long long gx_ld;                                   // hoisted input projection [B,3H] (nullable)
  const float* b_ih; const float* b_hh;                               // [3H]; b_ih nullable when folded into Gx
  const float* h_prev; long long hp_ld;                               // [B,H] fp32
  int B, H;
  void* stash;                                                        // [B,4H] TS: r, z, n, gh_n (nullable in inference)
  float* h_out; long long h_ld;                                       // fp32 h'
  void* h_op; long long hop_ld;                                       // operand-typed h' (nullable)
};

__device__ __forceinline__ void sum3(const float* __restrict__ q, int n, long long stride, int H, float& s0, float& s1, float& s2) {
  int p = 0;
  for (; p + 4 <= n; p += 4) {
    const float* q1 = q + stride; const float* q2 = q1 + stride; const float* q3 = q2 + stride;
    const float a0 = q[0], a1 = q[H], a2 = q[2 * H], b0 = q1[0], b1 = q1[H], b2 = q1[2 * H];
    const float c0 = q2[0], c1 = q2[H], c2 = q2[2 * H], d0 = q3[0], d1 = q3[H], d2 = q3[2 * H];
    s0 += (a0 + b0) + (c0 + d0); s1 += (a1 + b1) + (c1 + d1); s2 += (a2 + b2) + (c2 + d2);
    q = q3 + stride;
  }
  for (; p < n; ++p) { s0 += q[0]; s1 += q[H]; s2 += q[2 * H]; q += stride; }
}

template <typename TS, typename TO>
__device__ __forceinline__ void gru_cell_fwd_body(const FwdArgs& a, int bx, int b, int tid) {
  constexpr bool FAST = FastMath<TO>::value;
  const int H = a.H, j = bx * THREADS + tid;
  if (j >= H) return;
  float ir = 0.f, iz = 0.f, in_ = 0.f, hr = 0.f, hz = 0.f, hn = 0.f;
  if (a.Gx) { const float* g = a.Gx + b * a.gx_ld + j; ir = g[0]; iz = g[H]; in_ = g[2 * H]; }
  if (a.b_ih) { ir += a.b_ih[j]; iz += a.b_ih[H + j]; in_ += a.b_ih[2 * H + j]; }
  if (a.b_hh) { hr = a.b_hh[j]; hz = a.b_hh[H + j]; hn = a.b_hh[2 * H + j]; }
  if (a.Px) sum3(a.Px + b * a.px_ld + j, a.n_px, a.px_stride, H, ir, iz, in_);
  if (a.Ph) sum3(a.Ph + b * a.ph_ld + j, a.n_ph, a.ph_stride, H, hr, hz, hn);
  const float hp = a.h_prev[b * a.hp_ld + j];
  const float r = act_sigmoid<FAST>(ir + hr), z = act_sigmoid<FAST>(iz + hz);
  const float n = act_tanh<FAST>(fmaf(r, hn, in_));
  const float hnew = fmaf(z, hp - n, n);                       // (1 - z) * n + z * h
  a.h_out[b * a.h_ld + j] = hnew;
  if (a.stash) {
    TS* s = reinterpret_cast<TS*>(a.stash) + b * 4 * H + j;
    s[0] = from_f32<TS>(r); s[H] = from_f32<TS>(z); s[2 * H] = from_f32<TS>(n); s[3 * H] = from_f32<TS>(hn);
  }
  if (a.h_op) reinterpret_cast<TO*>(a.h_op)[b * a.hop_ld + j] = from_f32<TO>(hnew);
}
template <typename TS, typename TO>
__global__ void __launch_bounds__(THREADS) gru_cell_fwd_kernel(FwdArgs a) {
  pdl_wait();            // prerequisites complete ...
  pdl_launch_next();     // ... only then let the NEXT kernel be scheduled (depth-1 look-ahead, no cascade of resident waiters)
  gru_cell_fwd_body<TS, TO>(a, blockIdx.x, blockIdx.y, threadIdx.x);
}

// Backward of one step.  dh' = dh_ext + dh_ext2 + carry (z_{t+1} * dh'_{t+1}, the direct path) + sum_s dHp[s] (h-part
// dgrad partials of step t+1) + sum_s dQp[s] (attention-query path).  Writes dGi / dGh [B,3H] (operand type) and the
// new carry z * dh'.
struct BwdArgs {
  const float* dh_ext; long long dh_ld;
  const float* dh_ext2; long long dh2_ld;
  const float* dHp; int n_p; long long p_stride; long long p_ld;   // nullable
  const float* dQp; int n_q; long long q_stride; long long q_ld;   // nullable
  float* carry; int first;                                          // [B,H] in/out; first => treated as 0
  const void* stash;                                                // [B,4H] TS
  const float* h_prev; long long hp_ld;                             // [B,H] fp32
  int B, H;
  void* dGi; void* dGh; long long dg_ld;                            // [B,3H] TO each
};

__device__ __forceinline__ float sum1(const float* __restrict__ q, int n, long long stride) {
  float s = 0.f;
  int p = 0;
  for (; p + 4 <= n; p += 4) { const float v0 = q[0], v1 = q[stride], v2 = q[2 * stride], v3 = q[3 * stride]; s += (v0 + v1) + (v2 + v3); q += 4 * stride; }
  for (; p < n; ++p) { s += q[0]; q += stride; }
  return s;
}

template <typename TS, typename TO>
__device__ __forceinline__ void gru_cell_bwd_body(const BwdArgs& a, int bx, int b, int tid) {
  const int H = a.H, j = bx * THREADS + tid;
  if (j >= H) return;
  const TS* s = reinterpret_cast<const TS*>(a.stash) + b * 4 * H + j;
  const float r = to_f32<TS>(s[0]), z = to_f32<TS>(s[H]), n = to_f32<TS>(s[2 * H]), hn = to_f32<TS>(s[3 * H]);
  const float hp = a.h_prev[b * a.hp_ld + j];
  float dh = a.first ? 0.f : a.carry[b * H + j];
  if (a.dh_ext) dh += a.dh_ext[b * a.dh_ld + j];
  if (a.dh_ext2) dh += a.dh_ext2[b * a.dh2_ld + j];
  if (a.dHp) dh += sum1(a.dHp + b * a.p_ld + j, a.n_p, a.p_stride);
  if (a.dQp) dh += sum1(a.dQp + b * a.q_ld + j, a.n_q, a.q_stride);
  const float dn_pre = dh * (1.f - z) * (1.f - n * n);
  const float dz_pre = dh * (hp - n) * z * (1.f - z);
  const float dr_pre = dn_pre * hn * r * (1.f - r);
  a.carry[b * H + j] = dh * z;
  TO* gi = reinterpret_cast<TO*>(a.dGi) + b * a.dg_ld + j;
  TO* gh = reinterpret_cast<TO*>(a.dGh) + b * a.dg_ld + j;
  gi[0] = from_f32<TO>(dr_pre); gi[H] = from_f32<TO>(dz_pre); gi[2 * H] = from_f32<TO>(dn_pre);
  gh[0] = from_f32<TO>(dr_pre); gh[H] = from_f32<TO>(dz_pre); gh[2 * H] = from_f32<TO>(dn_pre * r);
}
template <typename TS, typename TO>
__global__ void __launch_bounds__(THREADS) gru_cell_bwd_kernel(BwdArgs a) {
  pdl_wait();            // prerequisites complete ...
  pdl_launch_next();     // ... only then let the NEXT kernel be scheduled (depth-1 look-ahead, no cascade of resident waiters)
  gru_cell_bwd_body<TS, TO>(a, blockIdx.x, blockIdx.y, threadIdx.x);
}

template <typename TS, typename TO>
static int launch_fwd(const FwdArgs& a, cudaStream_t st) {
  ProfScope prof(KC_CELL_FWD, a.B, a.H, a.n_px + a.n_ph, st);
  RN_CUDA_OK(launch_pdl(gru_cell_fwd_kernel<TS, TO>, dim3(rn_cdiv(a.H, THREADS), a.B), dim3(THREADS), 0, st, a));
  RN_LAUNCH_OK();
  return 0;
}
template <typename TS, typename TO>
static int launch_bwd(const BwdArgs& a, cudaStream_t st) {
  ProfScope prof(KC_CELL_BWD, a.B, a.H, (a.dHp ? a.n_p : 0) + (a.dQp ? a.n_q : 0), st);
  RN_CUDA_OK(launch_pdl(gru_cell_bwd_kernel<TS, TO>, dim3(rn_cdiv(a.H, THREADS), a.B), dim3(THREADS), 0, st, a));
  RN_LAUNCH_OK();
  return 0;
}
}  // namespace gru
