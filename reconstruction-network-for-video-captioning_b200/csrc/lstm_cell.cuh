// Fused LSTM gate-activation / cell-update kernels (reference: nn.LSTM step at models/decoder.py:66,
// models/global_reconstructor.py:43, models/local_reconstructor.py:52; PyTorch gate order i,f,g,o).
//
// Forward:  pre = sum_s P[s] (split-K partials of [x;h] W_rec^T) + Gx (hoisted input projection) + b_ih + b_hh
//           i,f,o = sigmoid, g = tanh ; c' = f*c + i*g ; h' = o*tanh(c')
// The activated gates are stashed (TS = float or bf16) for BPTT; h' is written both as fp32 (returned
// hiddens) and in the GEMM operand type straight into next step's [x ; h] operand row.
// HBM/L2-bound elementwise work: 4 consecutive hidden units per thread, 16-byte loads.
#pragma once
#include "common.cuh"

namespace cell {
constexpr int THREADS = 128;

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

template <typename TS, typename TO>
__global__ void __launch_bounds__(THREADS) lstm_cell_fwd_kernel(FwdArgs a) {
  const int H4 = a.H >> 2;
  const long long idx = (long long)blockIdx.x * THREADS + threadIdx.x;
  if (idx >= (long long)a.B * H4) return;
  const int b = (int)(idx / H4), j = (int)(idx % H4) * 4;
  float pre[4][4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int col = g * a.H + j;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < a.n_p; ++p) {
      const float4 v = *reinterpret_cast<const float4*>(a.P + p * a.p_stride + (long long)b * a.p_ld + col);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    if (a.Gx) { const float4 v = *reinterpret_cast<const float4*>(a.Gx + (long long)b * a.gx_ld + col); s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
    if (a.b1) { const float4 v = *reinterpret_cast<const float4*>(a.b1 + col); s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
    if (a.b2) { const float4 v = *reinterpret_cast<const float4*>(a.b2 + col); s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
    pre[g][0] = s.x; pre[g][1] = s.y; pre[g][2] = s.z; pre[g][3] = s.w;
  }
  const float4 cp4 = *reinterpret_cast<const float4*>(a.c_prev + (long long)b * a.H + j);
  const float cp[4] = {cp4.x, cp4.y, cp4.z, cp4.w};
  float gi[4], gf[4], gg[4], go[4], cn[4], hn[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    gi[k] = sigmoidf_(pre[0][k]);
    gf[k] = sigmoidf_(pre[1][k]);
    gg[k] = tanhf(pre[2][k]);
    go[k] = sigmoidf_(pre[3][k]);
    cn[k] = gf[k] * cp[k] + gi[k] * gg[k];
    hn[k] = go[k] * tanhf(cn[k]);
  }
  *reinterpret_cast<float4*>(a.c_out + (long long)b * a.H + j) = make_float4(cn[0], cn[1], cn[2], cn[3]);
  if (a.h_out) *reinterpret_cast<float4*>(a.h_out + (long long)b * a.h_ld + j) = make_float4(hn[0], hn[1], hn[2], hn[3]);
  if (a.gates_out) {
    TS* g = reinterpret_cast<TS*>(a.gates_out) + (long long)b * 4 * a.H + j;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      g[0 * a.H + k] = from_f32<TS>(gi[k]);
      g[1 * a.H + k] = from_f32<TS>(gf[k]);
      g[2 * a.H + k] = from_f32<TS>(gg[k]);
      g[3 * a.H + k] = from_f32<TS>(go[k]);
    }
  }
  if (a.h_op) {
    TO* o = reinterpret_cast<TO*>(a.h_op) + (long long)b * a.hop_ld + j;
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = from_f32<TO>(hn[k]);
  }
  if (a.h_op2) {
    TO* o = reinterpret_cast<TO*>(a.h_op2) + (long long)b * a.hop2_ld + j;
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = from_f32<TO>(hn[k]);
  }
}

// Backward of one step.  dh = dh_ext + sum_s dXp[s][:, col0:col0+H] + scale_q * (dQ @ Wq)  (attention-query path,
// dQ [B,A], Wq [A,H] = attn_W, fused here so the BPTT chain has no extra GEMM launch).
struct BwdArgs {
  const float* dh_ext; long long dh_ld; const float* dh_scale;    // [B,H] nullable; optional device scalar multiplier
  const float* dh_ext2; long long dh2_ld;                          // second external term (nullable)
  const float* dXp; int n_p; long long p_stride; long long p_ld; int col0;   // partials of d[x;h] (nullable)
  const float* dQ; const float* Wq; int A;                         // nullable
  float* dc;                                                       // [B,H] in/out (dc_next -> dc_prev); first => treated as 0
  int first;
  const void* gates;                                               // [B,4H] TS
  const float* c_prev; const float* c_new;                         // [B,H]
  int B, H;
  void* dG; long long dg_ld;                                       // [B,4H] TO (GEMM operand)
};

template <typename TS, typename TO>
__global__ void __launch_bounds__(THREADS) lstm_cell_bwd_kernel(BwdArgs a) {
  const int H4 = a.H >> 2;
  const long long idx = (long long)blockIdx.x * THREADS + threadIdx.x;
  if (idx >= (long long)a.B * H4) return;
  const int b = (int)(idx / H4), j = (int)(idx % H4) * 4;
  float dh[4] = {0.f, 0.f, 0.f, 0.f};
  if (a.dh_ext) {
    const float4 v = *reinterpret_cast<const float4*>(a.dh_ext + (long long)b * a.dh_ld + j);
    const float sc = a.dh_scale ? *a.dh_scale : 1.f;
    dh[0] += sc * v.x; dh[1] += sc * v.y; dh[2] += sc * v.z; dh[3] += sc * v.w;
  }
  if (a.dh_ext2) {
    const float4 v = *reinterpret_cast<const float4*>(a.dh_ext2 + (long long)b * a.dh2_ld + j);
    dh[0] += v.x; dh[1] += v.y; dh[2] += v.z; dh[3] += v.w;
  }
  if (a.dXp) {
    for (int p = 0; p < a.n_p; ++p) {
      const float4 v = *reinterpret_cast<const float4*>(a.dXp + p * a.p_stride + (long long)b * a.p_ld + a.col0 + j);
      dh[0] += v.x; dh[1] += v.y; dh[2] += v.z; dh[3] += v.w;
    }
  }
  if (a.dQ) {
    const float* q = a.dQ + (long long)b * a.A;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
    for (int k = 0; k < a.A; ++k) {
      const float qk = __ldg(q + k);
      const float4 w = __ldg(reinterpret_cast<const float4*>(a.Wq + (long long)k * a.H + j));
      s0 = fmaf(qk, w.x, s0); s1 = fmaf(qk, w.y, s1); s2 = fmaf(qk, w.z, s2); s3 = fmaf(qk, w.w, s3);
    }
    dh[0] += s0; dh[1] += s1; dh[2] += s2; dh[3] += s3;
  }
  const TS* g = reinterpret_cast<const TS*>(a.gates) + (long long)b * 4 * a.H + j;
  const float4 cp4 = *reinterpret_cast<const float4*>(a.c_prev + (long long)b * a.H + j);
  const float4 cn4 = *reinterpret_cast<const float4*>(a.c_new + (long long)b * a.H + j);
  const float cp[4] = {cp4.x, cp4.y, cp4.z, cp4.w}, cn[4] = {cn4.x, cn4.y, cn4.z, cn4.w};
  float dcn[4] = {0.f, 0.f, 0.f, 0.f};
  if (!a.first) {
    const float4 v = *reinterpret_cast<const float4*>(a.dc + (long long)b * a.H + j);
    dcn[0] = v.x; dcn[1] = v.y; dcn[2] = v.z; dcn[3] = v.w;
  }
  float dgi[4], dgf[4], dgg[4], dgo[4], dcp[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float gi = to_f32<TS>(g[0 * a.H + k]), gf = to_f32<TS>(g[1 * a.H + k]);
    const float gg = to_f32<TS>(g[2 * a.H + k]), go = to_f32<TS>(g[3 * a.H + k]);
    const float tc = tanhf(cn[k]);
    const float dc = dcn[k] + dh[k] * go * (1.f - tc * tc);
    dgo[k] = dh[k] * tc * go * (1.f - go);
    dgi[k] = dc * gg * gi * (1.f - gi);
    dgf[k] = dc * cp[k] * gf * (1.f - gf);
    dgg[k] = dc * gi * (1.f - gg * gg);
    dcp[k] = dc * gf;
  }
  *reinterpret_cast<float4*>(a.dc + (long long)b * a.H + j) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
  TO* o = reinterpret_cast<TO*>(a.dG) + (long long)b * a.dg_ld + j;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    o[0 * a.H + k] = from_f32<TO>(dgi[k]);
    o[1 * a.H + k] = from_f32<TO>(dgf[k]);
    o[2 * a.H + k] = from_f32<TO>(dgg[k]);
    o[3 * a.H + k] = from_f32<TO>(dgo[k]);
  }
}

template <typename TS, typename TO>
static int launch_fwd(const FwdArgs& a, cudaStream_t st) {
  if (a.H % 4 || a.p_ld % 4 || a.p_stride % 4 || (a.Gx && a.gx_ld % 4) || (a.h_out && a.h_ld % 4)) return RECNET_ERR_ALIGNMENT;
  const long long n = (long long)a.B * (a.H / 4);
  ProfScope prof(KC_CELL_FWD, a.B, a.H, a.n_p, st);
  lstm_cell_fwd_kernel<TS, TO><<<rn_cdiv(n, THREADS), THREADS, 0, st>>>(a);
  RN_LAUNCH_OK();
  return 0;
}
template <typename TS, typename TO>
static int launch_bwd(const BwdArgs& a, cudaStream_t st) {
  if (a.H % 4 || (a.dXp && (a.p_ld % 4 || a.p_stride % 4 || a.col0 % 4)) || (a.dh_ext && a.dh_ld % 4) ||
      (a.dh_ext2 && a.dh2_ld % 4))
    return RECNET_ERR_ALIGNMENT;
  const long long n = (long long)a.B * (a.H / 4);
  ProfScope prof(KC_CELL_BWD, a.B, a.H, a.dXp ? a.n_p : 0, st);
  lstm_cell_bwd_kernel<TS, TO><<<rn_cdiv(n, THREADS), THREADS, 0, st>>>(a);
  RN_LAUNCH_OK();
  return 0;
}
}  // namespace cell
