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
  // inter-layer dropout (nn.LSTM(dropout=p), train mode): applied to the h_op2 copy only (the next layer's input)
  float op2_drop; const unsigned long long* rng; unsigned int site; long long drop_base;
};

// body: virtual block (bx = 256-wide slice of H, b = sample): no integer division, pointer-bump partial loop
// returns h' of this thread's unit as the GEMM operand type sees it (0 outside H)
template <typename TS, typename TO>
__device__ __forceinline__ float lstm_cell_fwd_body(const FwdArgs& a, int bx, int b, int tid) {
  constexpr bool FAST = FastMath<TO>::value;
  const int H = a.H, j = bx * THREADS + tid;
  if (j >= H) return 0.f;
  float pi = 0.f, pf = 0.f, pg = 0.f, po = 0.f;
  const float cp = a.c_prev[b * H + j];
  if (a.Gx) { const float* g = a.Gx + b * a.gx_ld + j; pi = g[0]; pf = g[H]; pg = g[2 * H]; po = g[3 * H]; }
  if (a.b1) { pi += a.b1[j]; pf += a.b1[H + j]; pg += a.b1[2 * H + j]; po += a.b1[3 * H + j]; }
  if (a.b2) { pi += a.b2[j]; pf += a.b2[H + j]; pg += a.b2[2 * H + j]; po += a.b2[3 * H + j]; }
  const float* q = a.P + b * a.p_ld + j;
  int p = 0;
  for (; p + 4 <= a.n_p; p += 4) {              // 16 independent loads in flight
    const float* q1 = q + a.p_stride; const float* q2 = q1 + a.p_stride; const float* q3 = q2 + a.p_stride;
    const float v00 = q[0], v01 = q[H], v02 = q[2 * H], v03 = q[3 * H];
    const float v10 = q1[0], v11 = q1[H], v12 = q1[2 * H], v13 = q1[3 * H];
    const float v20 = q2[0], v21 = q2[H], v22 = q2[2 * H], v23 = q2[3 * H];
    const float v30 = q3[0], v31 = q3[H], v32 = q3[2 * H], v33 = q3[3 * H];
    pi += (v00 + v10) + (v20 + v30); pf += (v01 + v11) + (v21 + v31);
    pg += (v02 + v12) + (v22 + v32); po += (v03 + v13) + (v23 + v33);
    q = q3 + a.p_stride;
  }
  for (; p < a.n_p; ++p) { pi += q[0]; pf += q[H]; pg += q[2 * H]; po += q[3 * H]; q += a.p_stride; }
  const float gi = act_sigmoid<FAST>(pi), gf = act_sigmoid<FAST>(pf), gg = act_tanh<FAST>(pg), go = act_sigmoid<FAST>(po);
  const float cn = fmaf(gf, cp, gi * gg);
  const float hn = go * act_tanh<FAST>(cn);
  a.c_out[b * H + j] = cn;
  if (a.h_out) a.h_out[b * a.h_ld + j] = hn;
  if (a.gates_out) {
    TS* g = reinterpret_cast<TS*>(a.gates_out) + b * 4 * H + j;
    g[0] = from_f32<TS>(gi); g[H] = from_f32<TS>(gf); g[2 * H] = from_f32<TS>(gg); g[3 * H] = from_f32<TS>(go);
  }
  if (a.h_op) reinterpret_cast<TO*>(a.h_op)[b * a.hop_ld + j] = from_f32<TO>(hn);
  if (a.h_op2) {
    float hd = hn;
    if (a.op2_drop > 0.f) hd *= dropout_scale(a.rng, a.site, (uint64_t)(a.drop_base + (long long)b * H + j), a.op2_drop);
    reinterpret_cast<TO*>(a.h_op2)[b * a.hop2_ld + j] = from_f32<TO>(hd);
  }
  return to_f32<TO>(from_f32<TO>(hn));
}

template <typename TS, typename TO>
__global__ void __launch_bounds__(THREADS) lstm_cell_fwd_kernel(FwdArgs a) {
  pdl_wait();            // prerequisites complete ...
  pdl_launch_next();     // ... only then let the NEXT kernel be scheduled (depth-1 look-ahead, no cascade of resident waiters)
  lstm_cell_fwd_body<TS, TO>(a, blockIdx.x, blockIdx.y, threadIdx.x);
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
  float q_drop; const unsigned long long* rng; unsigned int q_site; long long q_base;   // dropout mask on the dQp term (inter-layer path)
  float* dc;                                                       // [B,H] in/out (dc_next -> dc_prev); first => treated as 0
  int first;
  const void* gates;                                               // [B,4H] TS
  const float* c_prev; const float* c_new;                         // [B,H]
  int B, H;
  void* dG; long long dg_ld;                                       // [B,4H] TO (GEMM operand)
};

__device__ __forceinline__ float sum_partials1(const float* __restrict__ q, int n_p, long long stride) {
  float s = 0.f;
  int p = 0;
  for (; p + 4 <= n_p; p += 4) {
    const float v0 = q[0], v1 = q[stride], v2 = q[2 * stride], v3 = q[3 * stride];
    s += (v0 + v1) + (v2 + v3);
    q += 4 * stride;
  }
  for (; p < n_p; ++p) { s += q[0]; q += stride; }
  return s;
}

template <typename TS, typename TO>
__device__ __forceinline__ void lstm_cell_bwd_body(const BwdArgs& a, int bx, int b, int tid) {
  constexpr bool FAST = FastMath<TO>::value;
  const int H = a.H, j = bx * THREADS + tid;
  if (j >= H) return;
  const TS* g = reinterpret_cast<const TS*>(a.gates) + b * 4 * H + j;
  const float gi = to_f32<TS>(g[0]), gf = to_f32<TS>(g[H]), gg = to_f32<TS>(g[2 * H]), go = to_f32<TS>(g[3 * H]);
  const float cp = a.c_prev[b * H + j], cn = a.c_new[b * H + j];
  const float dcn = a.first ? 0.f : a.dc[b * H + j];
  float dh = 0.f;
  if (a.dh_ext) dh = (a.dh_scale ? *a.dh_scale : 1.f) * a.dh_ext[b * a.dh_ld + j];
  if (a.dh_ext2) dh += a.dh_ext2[b * a.dh2_ld + j];
  if (a.dXp) dh += sum_partials1(a.dXp + b * a.p_ld + a.col0 + j, a.n_p, a.p_stride);
  if (a.dQp) {
    float q = sum_partials1(a.dQp + b * a.q_ld + j, a.n_q, a.q_stride);
    if (a.q_drop > 0.f) q *= dropout_scale(a.rng, a.q_site, (uint64_t)(a.q_base + (long long)b * H + j), a.q_drop);
    dh += q;
  }
  const float tc = act_tanh<FAST>(cn);
  const float dc = fmaf(dh * go, 1.f - tc * tc, dcn);
  a.dc[b * H + j] = dc * gf;
  TO* o = reinterpret_cast<TO*>(a.dG) + b * a.dg_ld + j;
  o[0] = from_f32<TO>(dc * gg * gi * (1.f - gi));
  o[H] = from_f32<TO>(dc * cp * gf * (1.f - gf));
  o[2 * H] = from_f32<TO>(dc * gi * (1.f - gg * gg));
  o[3 * H] = from_f32<TO>(dh * tc * go * (1.f - go));
}

template <typename TS, typename TO>
__global__ void __launch_bounds__(THREADS) lstm_cell_bwd_kernel(BwdArgs a) {
  pdl_wait();            // prerequisites complete ...
  pdl_launch_next();     // ... only then let the NEXT kernel be scheduled (depth-1 look-ahead, no cascade of resident waiters)
  lstm_cell_bwd_body<TS, TO>(a, blockIdx.x, blockIdx.y, threadIdx.x);
}

template <typename TS, typename TO>
static int launch_fwd(const FwdArgs& a, cudaStream_t st) {
  ProfScope prof(KC_CELL_FWD, a.B, a.H, a.n_p, st);
  RN_CUDA_OK(launch_pdl(lstm_cell_fwd_kernel<TS, TO>, dim3(rn_cdiv(a.H, THREADS), a.B), dim3(THREADS), 0, st, a));
  RN_LAUNCH_OK();
  return 0;
}
template <typename TS, typename TO>
static int launch_bwd(const BwdArgs& a, cudaStream_t st) {
  ProfScope prof(KC_CELL_BWD, a.B, a.H, (a.dXp ? a.n_p : 0) + (a.dQp ? a.n_q : 0), st);
  RN_CUDA_OK(launch_pdl(lstm_cell_bwd_kernel<TS, TO>, dim3(rn_cdiv(a.H, THREADS), a.B), dim3(THREADS), 0, st, a));
  RN_LAUNCH_OK();
  return 0;
}
}  // namespace cell
