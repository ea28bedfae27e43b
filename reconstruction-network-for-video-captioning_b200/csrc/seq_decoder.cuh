// Decoder sequence driver: the whole teacher-forced loop of train.forward_decoder (train.py:17-75) over
// Decoder.forward (models/decoder.py:45-70), forward and BPTT, restructured for B200:
//   * U.v hoisted out of the loop (the reference recomputes it 31x, decoder.py:54)        -> 1 batched GEMM
//   * W_ih split into [W_emb | W_ctx]: the embedding half is time-batched (M = L*B)        -> 1 batched GEMM
//   * per step: [ctx_t ; h_{t-1}] @ [W_ctx | W_hh]^T as ONE K-concatenated tensor-core GEMM (no torch.cat,
//     decoder.py:64), split-K partials summed inside the fused cell kernel
//   * vocabulary projection batched over all steps (M = L*B), fused CE forward/backward over stacked logits
//   * BPTT mirrors it; all weight gradients are batched GEMMs over the stashed operands after the loop.
#pragma once
#include "gru_cell.cuh"
#include "mega.cuh"
#include "runtime.cuh"

namespace dec {
using namespace rt;

enum : unsigned { SITE_EMB = 1, SITE_LOGITS = 2 };

template <typename T>
struct Ws {
  // geometry
  int EMBp, KX, Vp, Vld;
  int nch, Bc;                       // concurrent sample chains, max rows per chain
  int G;                             // gates per unit: 4 (LSTM) or 3 (GRU)
  GemmPlan pl_wh, pl_gate, pl_dx, pl_dq;
  GemmPlan pl_gx, pl_gh, pl_dxx, pl_dxh;   // GRU: x-part / h-part kept separate (n gate)
  // operand copies of the weights / inputs (rebuilt every forward: the optimiser changes the masters)
  T *Wemb, *Wrec, *U, *Wa, *Wout, *feats;
  // forward state kept for BPTT
  float* Uv; T* Xe; float* Gx; T* X; float* WhP; float* Wh; float* e; float* P; float* P2; T* gates; float* c;
  float *logits, *lse, *row_loss;
  // backward scratch
  T* dlogits; float* dHext; T* dG; T* dG2; float* dXp; float* dXp2; float* dQp; float* dWh; T* dWh_op; float* dUv; T* dUv_op; float* dw_acc;
  float* dc; float* dXe; float* splitk;
  uint8_t* table; size_t table_bytes; unsigned* bar; int* err;     // loop-kernel phase table, grid-barrier counter, error flag
  // layers 1 .. NL-1 of a stacked decoder (index l-1): operand rows X_l[t] = [h^{l-1}_t ; h^l_{t-1}], K = 2H
  int NL;
  GemmPlan pl_gate_x, pl_dx_x;
  T* Wrec_x[RECNET_MAX_LAYERS - 1]; T* X_x[RECNET_MAX_LAYERS - 1]; T* gates_x[RECNET_MAX_LAYERS - 1]; float* c_x[RECNET_MAX_LAYERS - 1];
  T* dG_x[RECNET_MAX_LAYERS - 1]; float* dXp_x[RECNET_MAX_LAYERS - 1]; float* dc_x[RECNET_MAX_LAYERS - 1];
  size_t bytes;
};

template <typename T>
static Ws<T> plan(const recnet_decoder_desc& d, void* base) {
  Ws<T> w;
  const int B = d.B, L = d.L, H = d.H, E = d.E, A = d.A, V = d.V, Tn = d.T;
  w.EMBp = round_up(d.EMB, Prec<T>::kpad);
  w.KX = E + H;
  w.Vp = round_up(V, 8);
  w.Vld = round_up(V, 4);
  w.nch = num_chains(B);
  w.Bc = chain_rows_max(B, w.nch);
  const int tgt = w.nch > 1 ? NUM_SMS / 2 : NUM_SMS;
  w.pl_wh = plan_gemm<T>(w.Bc, A, H, tgt);
  w.pl_gate = plan_gemm<T>(w.Bc, 4 * H, w.KX, tgt);
  w.pl_dx = plan_gemm<T>(w.Bc, w.KX, 4 * H, tgt);
  w.pl_dq = plan_gemm<T>(w.Bc, H, A, tgt);
  w.G = d.cell == RECNET_CELL_GRU ? 3 : 4;
  w.pl_gx = plan_gemm<T>(w.Bc, 3 * H, E, tgt);
  w.pl_gh = plan_gemm<T>(w.Bc, 3 * H, H, tgt);
  w.pl_dxx = plan_gemm<T>(w.Bc, E, 3 * H, tgt);
  w.pl_dxh = plan_gemm<T>(w.Bc, H, 3 * H, tgt);
  const bool gru_ = d.cell == RECNET_CELL_GRU;
  Bump m(base);
  w.Wemb = m.take<T>((size_t)4 * H * w.EMBp);
  w.Wrec = m.take<T>((size_t)4 * H * w.KX);
  w.U = m.take<T>((size_t)A * E);
  w.Wa = m.take<T>((size_t)A * H);
  w.Wout = m.take<T>((size_t)V * H);
  w.feats = m.take<T>((size_t)B * Tn * E);
  w.Uv = m.take<float>((size_t)B * Tn * A);
  w.Xe = m.take<T>((size_t)L * B * w.EMBp);
  w.Gx = m.take<float>((size_t)L * B * 4 * H);
  w.X = m.take<T>((size_t)(L + 1) * B * w.KX);
  w.WhP = m.take<float>((size_t)w.nch * w.pl_wh.splits * w.Bc * A);
  w.Wh = m.take<float>((size_t)L * B * A);
  w.e = m.take<float>((size_t)L * B * Tn);
  w.P = m.take<float>((size_t)w.nch * (gru_ ? w.pl_gx.splits : w.pl_gate.splits) * w.Bc * 4 * H + (size_t)16 * B * 4 * H);
  w.P2 = m.take<float>(gru_ ? (size_t)w.nch * w.pl_gh.splits * w.Bc * 3 * H : 1);
  w.gates = m.take<T>((size_t)L * B * 4 * H);
  w.c = m.take<float>((size_t)(L + 1) * B * H);
  w.logits = m.take<float>((size_t)L * B * w.Vld);
  w.lse = m.take<float>((size_t)L * B);
  w.row_loss = m.take<float>((size_t)L * B);
  w.dlogits = m.take<T>((size_t)L * B * w.Vp);
  w.dHext = m.take<float>((size_t)L * B * H);
  w.dG = m.take<T>((size_t)L * B * 4 * H);
  w.dXp = m.take<float>((size_t)w.nch * (gru_ ? w.pl_dxx.splits : w.pl_dx.splits) * w.Bc * w.KX);
  w.dXp2 = m.take<float>(gru_ ? (size_t)w.nch * w.pl_dxh.splits * w.Bc * H : 1);
  w.dG2 = m.take<T>(gru_ ? (size_t)L * B * 3 * H : 1);
  w.dQp = m.take<float>((size_t)w.nch * w.pl_dq.splits * w.Bc * H);
  w.dWh = m.take<float>((size_t)L * B * A);
  w.dWh_op = m.take<T>((size_t)L * B * A);
  w.dUv = m.take<float>((size_t)B * Tn * A);
  w.dUv_op = m.take<T>((size_t)B * Tn * A);
  w.dw_acc = m.take<float>((size_t)B * A);
  w.dc = m.take<float>((size_t)B * H);
  w.dXe = m.take<float>((size_t)L * B * w.EMBp);
  w.splitk = m.take<float>(SPLITK_SCRATCH_FLOATS);
  w.table_bytes = mega::table_bytes(L);
  w.table = m.take<uint8_t>(w.table_bytes);
  w.bar = m.take<unsigned>(64);
  w.err = m.take<int>(64);
  w.NL = d.n_layers < 1 ? 1 : d.n_layers;
  w.pl_gate_x = plan_gemm<T>(w.Bc, 4 * H, 2 * H, tgt);
  w.pl_dx_x = plan_gemm<T>(w.Bc, 2 * H, 4 * H, tgt);
  for (int l = 1; l < w.NL && l < RECNET_MAX_LAYERS; ++l) {
    w.Wrec_x[l - 1] = m.take<T>((size_t)4 * H * 2 * H);
    w.X_x[l - 1] = m.take<T>((size_t)(L + 1) * B * 2 * H);
    w.gates_x[l - 1] = m.take<T>((size_t)L * B * 4 * H);
    w.c_x[l - 1] = m.take<float>((size_t)(L + 1) * B * H);
    w.dG_x[l - 1] = m.take<T>((size_t)L * B * 4 * H);
    w.dXp_x[l - 1] = m.take<float>((size_t)w.pl_dx_x.splits * B * 2 * H);
    w.dc_x[l - 1] = m.take<float>((size_t)B * H);
  }
  w.bytes = m.off + 256;
  return w;
}

static inline int check(const recnet_decoder_desc& d) {
  if (d.B < 1 || d.T < 1 || d.L < 1 || d.V < 3 || d.EMB < 1 || d.T > attn::MAX_T) return RECNET_ERR_BAD_SHAPE;
  if (d.cell != RECNET_CELL_LSTM && d.cell != RECNET_CELL_GRU) return RECNET_ERR_UNSUPPORTED;
  if (d.n_layers > RECNET_MAX_LAYERS) return RECNET_ERR_UNSUPPORTED;
  if (d.n_layers > 1 && d.cell != RECNET_CELL_LSTM) return RECNET_ERR_UNSUPPORTED;      // stacked GRU: not built
  const int al = d.precision == RECNET_PREC_BF16 ? 8 : 4;
  if (d.E % al || d.H % al || d.A % 4) return RECNET_ERR_ALIGNMENT;
  return 0;
}

// Build the operand-typed copies of weights and features.
template <typename T>
static int prepare(const recnet_decoder_desc& d, const recnet_decoder_tensors& p, const float* feats, Ws<T>& w, cudaStream_t st) {
  const int H = d.H, E = d.E, A = d.A, V = d.V, EMB = d.EMB;
  const long long ldih = EMB + E;
  const int GH = w.G * H;
  RN_TRY(misc::cast_pad<T>(p.w_ih, ldih, w.Wemb, w.EMBp, GH, EMB, w.EMBp, st));
  RN_TRY(misc::cast_pad<T>(p.w_ih + EMB, ldih, w.Wrec, w.KX, GH, E, E, st));
  RN_TRY(misc::cast_pad<T>(p.w_hh, H, w.Wrec + E, w.KX, GH, H, H, st));
  RN_TRY(misc::cast_pad<T>(p.attn_U, E, w.U, E, A, E, E, st));
  RN_TRY(misc::cast_pad<T>(p.attn_W, H, w.Wa, H, A, H, H, st));
  RN_TRY(misc::cast_pad<T>(p.out_w, H, w.Wout, H, V, H, H, st));
  if (feats) RN_TRY(misc::cast_pad<T>(feats, E, w.feats, E, (long long)d.B * d.T, E, E, st));
  return 0;
}

// one decoder step for the sample rows [b0, b0+nb) of chain `ch`: attention -> gate GEMM -> cell.
// Pointer arguments are the FULL-batch row-0 addresses of step t; the row offset is applied here.
template <typename T>
static int step(mega::Emitter<T>& em, const recnet_decoder_desc& d, const recnet_decoder_tensors& p, Ws<T>& w, int t, int ch,
                int b0, int nb, const float* gx_t, T* x_t, T* x_next, float* Wh_t, float* e_t, T* gates_t, const float* c_prev,
                float* c_next, float* h_out) {
  const int H = d.H, E = d.E, A = d.A, Tn = d.T;
  float* WhP = w.WhP + (size_t)ch * w.pl_wh.splits * w.Bc * A;
  float* P = w.P + (size_t)ch * w.pl_gate.splits * w.Bc * 4 * H;
  T* xr = x_t + (size_t)b0 * w.KX;
  int n_whp = 0;
  if (t > 0) {   // h_{-1} = 0 -> W h = 0, skip the GEMM
    RN_TRY(em.gemm_partials(xr + E, w.KX, 0, w.Wa, H, 0, WhP, nb, A, H, w.pl_wh));
    n_whp = w.pl_wh.splits;
  }
  attn::FwdArgs fa{};
  fa.WhP = WhP; fa.n_whp = n_whp; fa.whp_stride = (long long)nb * A;
  fa.Uv = w.Uv + (size_t)b0 * Tn * A; fa.uv_bs = (long long)Tn * A; fa.uv_ts = A;
  fa.attn_b = p.attn_b; fa.attn_w = p.attn_w;
  fa.V = w.feats + (size_t)b0 * Tn * E; fa.v_bs = (long long)Tn * E; fa.v_ts = E;
  fa.B = nb; fa.Tn = Tn; fa.A = A; fa.D = E; fa.inv_T = 1.f / Tn; fa.normalize = 0;
  fa.Wh_out = Wh_t ? Wh_t + (size_t)b0 * A : nullptr; fa.e_out = e_t ? e_t + (size_t)b0 * Tn : nullptr;
  fa.ctx_out = xr; fa.ctx_ld = w.KX; fa.p_drop = 0.f;
  RN_TRY(em.attn_fwd(fa));
  if (d.cell == RECNET_CELL_GRU) {
    // x-part: ctx_t @ W_ctx^T (K = E) ; h-part: h_{t-1} @ W_hh^T (K = H): same operand rows, same weight copy, two column ranges
    float* Px = w.P + (size_t)ch * w.pl_gx.splits * w.Bc * 3 * H;
    float* Ph = w.P2 + (size_t)ch * w.pl_gh.splits * w.Bc * 3 * H;
    RN_TRY(em.gemm_partials(xr, w.KX, 0, w.Wrec, w.KX, 0, Px, nb, 3 * H, E, w.pl_gx));
    if (t > 0) RN_TRY(em.gemm_partials(xr + E, w.KX, 0, w.Wrec + E, w.KX, 0, Ph, nb, 3 * H, H, w.pl_gh));
    gru::FwdArgs ga{};
    ga.Px = Px; ga.n_px = w.pl_gx.splits; ga.px_stride = (long long)nb * 3 * H; ga.px_ld = 3 * H;
    ga.Ph = t > 0 ? Ph : nullptr; ga.n_ph = t > 0 ? w.pl_gh.splits : 0; ga.ph_stride = (long long)nb * 3 * H; ga.ph_ld = 3 * H;
    ga.Gx = gx_t + (size_t)b0 * 3 * H; ga.gx_ld = 3 * H; ga.b_ih = nullptr; ga.b_hh = p.b_hh;
    ga.h_prev = c_prev + (size_t)b0 * H; ga.hp_ld = H; ga.B = nb; ga.H = H;
    ga.stash = gates_t ? gates_t + (size_t)b0 * 4 * H : nullptr;
    ga.h_out = h_out + (size_t)b0 * H; ga.h_ld = H;
    ga.h_op = x_next + (size_t)b0 * w.KX + E; ga.hop_ld = w.KX;
    return gru::launch_fwd<T, T>(ga, em.st);
  }
  RN_TRY(em.gemm_partials(xr, w.KX, 0, w.Wrec, w.KX, 0, P, nb, 4 * H, w.KX, w.pl_gate));
  cell::FwdArgs ca{};
  ca.P = P; ca.n_p = w.pl_gate.splits; ca.p_stride = (long long)nb * 4 * H; ca.p_ld = 4 * H;
  ca.Gx = gx_t + (size_t)b0 * 4 * H; ca.gx_ld = 4 * H; ca.b1 = nullptr; ca.b2 = p.b_hh;
  ca.c_prev = c_prev + (size_t)b0 * H; ca.B = nb; ca.H = H;
  ca.gates_out = gates_t ? gates_t + (size_t)b0 * 4 * H : nullptr; ca.c_out = c_next + (size_t)b0 * H;
  ca.h_out = h_out + (size_t)b0 * H; ca.h_ld = H;
  ca.h_op = x_next + (size_t)b0 * w.KX + E; ca.hop_ld = w.KX; ca.h_op2 = nullptr;
  RN_TRY(em.cell_fwd(ca));
  return 0;
}

template <typename T>
static int forward(const recnet_decoder_desc& d, const recnet_decoder_tensors& p, const float* feats, const long long* tokens_in,
                   const long long* targets, const float* ce_weight, const unsigned long long* rng, void* ws, long long ws_bytes,
                   float* hiddens, float* ce_out, cudaStream_t st) {
  RN_TRY(check(d));
  Ws<T> w = plan<T>(d, ws);
  if ((long long)w.bytes > ws_bytes) return RECNET_ERR_WORKSPACE;
  const int B = d.B, L = d.L, H = d.H, E = d.E, A = d.A, V = d.V, Tn = d.T;
  const float p_emb = d.train ? d.p_emb_drop : 0.f, p_out = d.train ? d.p_out_drop : 0.f;
  RN_TRY(prepare<T>(d, p, feats, w, st));
  // hoisted projections
  RN_TRY(gemm_full<T>(w.feats, E, 0, w.U, E, 0, w.Uv, A, nullptr, B * Tn, A, E, 0, w.splitk, st));
  misc::embed_gather_kernel<T><<<L * B, 128, 0, st>>>(p.embedding, tokens_in, w.Xe, w.EMBp, L * B, d.EMB, w.EMBp, V,
                                                      d.embedding_scale, p_emb, rng, SITE_EMB);
  RN_LAUNCH_OK();
  const int GH = w.G * H;
  const bool is_gru = d.cell == RECNET_CELL_GRU;
  RN_TRY(gemm_full<T>(w.Xe, w.EMBp, 0, w.Wemb, w.EMBp, 0, w.Gx, GH, p.b_ih, L * B, GH, w.EMBp, 0, w.splitk, st));
  // initial state: h_{-1} = 0 (operand slot of X[0]), c_{-1} = 0
  RN_CUDA_OK(cudaMemsetAsync(w.X, 0, (size_t)B * w.KX * sizeof(T), st));
  RN_CUDA_OK(cudaMemsetAsync(w.c, 0, (size_t)B * H * sizeof(float), st));
  // time loop.  One chain (default): the whole loop is ONE persistent loop-kernel launch (bf16 build) or one kernel
  // per phase (fp32 build / RECNET_MEGA=0).  Several chains (RECNET_CHAINS, experimental): forked streams, eager.
  RN_CUDA_OK(cudaMemsetAsync(w.err, 0, sizeof(int), st));
  Chains& cs = chains();
  if (w.nch > 1) RN_TRY(cs.fork(st, w.nch));
  {
    mega::Emitter<T> em0(w.nch == 1 && !is_gru, st, (size_t)E + Tn + 8 * attn::BWD_THREADS + 2 * A);
    for (int t = 0; t < L; ++t) {
      T* x_t = w.X + (size_t)t * B * w.KX;
      for (int ch = 0; ch < w.nch; ++ch) {
        int b0, nb;
        chain_rows(B, w.nch, ch, &b0, &nb);
        mega::Emitter<T> emc(false, w.nch > 1 ? cs.s[ch] : st);
        // LSTM: c_{t-1} -> c_t live in w.c; GRU: the fp32 state is h itself (previous row of `hiddens`, zeros = w.c at t = 0)
        const float* s_prev = is_gru ? (t == 0 ? w.c : hiddens + (size_t)(t - 1) * B * H) : w.c + (size_t)t * B * H;
        RN_TRY(step<T>(w.nch == 1 ? em0 : emc, d, p, w, t, ch, b0, nb, w.Gx + (size_t)t * B * GH, x_t, x_t + (size_t)B * w.KX,
                       w.Wh + (size_t)t * B * A, w.e + (size_t)t * B * Tn, w.gates + (size_t)t * B * 4 * H,
                       s_prev, w.c + (size_t)(t + 1) * B * H, hiddens + (size_t)t * B * H));
      }
    }
    RN_TRY(em0.flush(w.table, w.table_bytes, w.bar, w.err, 1));
  }
  if (w.nch > 1) RN_TRY(cs.join(st, w.nch));
  // vocabulary projection over all steps, then the masked CE (train.py:54-60,68)
  RN_TRY(gemm_full<T>(w.X + (size_t)B * w.KX + E, w.KX, 0, w.Wout, H, 0, w.logits, w.Vld, p.out_b, L * B, V, H, 0,
                      w.splitk, st));
  if (targets && ce_weight && ce_out) {
    ProfScope prof(KC_CE, L * B, V, 0, st);
    loss::ce_fwd_kernel<<<L * B, loss::CE_THREADS, 0, st>>>(w.logits, w.Vld, targets, ce_weight, V, p_out, rng, SITE_LOGITS,
                                                            w.lse, w.row_loss);
    RN_LAUNCH_OK();
    loss::sum_kernel<<<1, 1024, 0, st>>>(w.row_loss, L * B, ce_out, 1.f);
    RN_LAUNCH_OK();
  }
  return 0;
}

template <typename T>
static int backward(const recnet_decoder_desc& d, const recnet_decoder_tensors& p, const float* feats, const long long* tokens_in,
                    const long long* targets, const float* ce_weight, const unsigned long long* rng, void* ws, long long ws_bytes,
                    const float* g_ce, const float* g_hiddens, const float* hiddens_fp32, const recnet_decoder_tensors& g,
                    cudaStream_t st) {
  RN_TRY(check(d));
  Ws<T> w = plan<T>(d, ws);
  if ((long long)w.bytes > ws_bytes) return RECNET_ERR_WORKSPACE;
  const int B = d.B, L = d.L, H = d.H, E = d.E, A = d.A, V = d.V, Tn = d.T, EMB = d.EMB;
  const float p_emb = d.train ? d.p_emb_drop : 0.f, p_out = d.train ? d.p_out_drop : 0.f;
  const int LB = L * B;
  const T* Hall = w.X + (size_t)B * w.KX + E;     // h_t rows, ld = KX
  // ---- CE backward and the vocabulary projection --------------------------------------------------------
  {
    ProfScope prof(KC_CE, LB, V, 1, st);
    loss::ce_bwd_kernel<T><<<LB, loss::CE_THREADS, 0, st>>>(w.logits, w.Vld, targets, ce_weight, w.lse, g_ce, V, w.Vp, p_out, rng,
                                                            SITE_LOGITS, w.dlogits, w.Vp);
  }
  RN_LAUNCH_OK();
  RN_TRY(gemm_full<T>(w.dlogits, w.Vp, 0, w.Wout, H, 1, w.dHext, H, nullptr, LB, H, V, 0, w.splitk, st));
  RN_TRY(gemm_full<T>(w.dlogits, w.Vp, 1, Hall, w.KX, 1, g.out_w, H, nullptr, V, H, LB, 0, w.splitk, st));
  RN_TRY(misc::colsum<T>(w.dlogits, w.Vp, LB, V, g.out_b, 0, w.splitk, st));
  // ---- BPTT: the same sample chains, each with private split-K scratch ------------------------------------------
  Chains& cs = chains();
  if (w.nch > 1) RN_TRY(cs.fork(st, w.nch));
  const bool is_gru = d.cell == RECNET_CELL_GRU;
  const int GH = w.G * H;
  mega::Emitter<T> em0(w.nch == 1 && !is_gru, st, (size_t)E + Tn + 8 * attn::BWD_THREADS + 2 * A);
  for (int t = L - 1; t >= 0; --t) {
    const bool last = (t == L - 1);
    for (int ch = 0; ch < w.nch; ++ch) {
      int b0, nb;
      chain_rows(B, w.nch, ch, &b0, &nb);
      mega::Emitter<T> emc(false, w.nch > 1 ? cs.s[ch] : st);
      mega::Emitter<T>& em = w.nch == 1 ? em0 : emc;
      float* dXp = w.dXp + (size_t)ch * w.pl_dx.splits * w.Bc * w.KX;
      float* dQp = w.dQp + (size_t)ch * w.pl_dq.splits * w.Bc * H;
      const size_t r = (size_t)t * B + b0;           // first (t, b) row of this chain
      if (is_gru) {
        float* dXx = w.dXp + (size_t)ch * w.pl_dxx.splits * w.Bc * w.KX;      // x-part dgrad partials [splits][nb, E]
        float* dXh = w.dXp2 + (size_t)ch * w.pl_dxh.splits * w.Bc * H;         // h-part dgrad partials [splits][nb, H]
        gru::BwdArgs gb{};
        gb.dh_ext = w.dHext + r * H; gb.dh_ld = H;
        gb.dh_ext2 = g_hiddens ? g_hiddens + r * H : nullptr; gb.dh2_ld = H;
        gb.dHp = last ? nullptr : dXh; gb.n_p = w.pl_dxh.splits; gb.p_stride = (long long)nb * H; gb.p_ld = H;
        gb.dQp = last ? nullptr : dQp; gb.n_q = w.pl_dq.splits; gb.q_stride = (long long)nb * H; gb.q_ld = H;
        gb.carry = w.dc + (size_t)b0 * H; gb.first = last ? 1 : 0;
        gb.stash = w.gates + r * 4 * H;
        gb.h_prev = t == 0 ? w.c + (size_t)b0 * H : hiddens_fp32 + ((size_t)(t - 1) * B + b0) * H; gb.hp_ld = H;
        gb.B = nb; gb.H = H; gb.dGi = w.dG + r * 3 * H; gb.dGh = w.dG2 + r * 3 * H; gb.dg_ld = 3 * H;
        RN_TRY((gru::launch_bwd<T, T>(gb, em.st)));
        RN_TRY(em.gemm_partials(w.dG + r * 3 * H, 3 * H, 0, w.Wrec, w.KX, 1, dXx, nb, E, 3 * H, w.pl_dxx));           // dctx = dGi @ W_ctx
        if (t > 0) RN_TRY(em.gemm_partials(w.dG2 + r * 3 * H, 3 * H, 0, w.Wrec + E, w.KX, 1, dXh, nb, H, 3 * H, w.pl_dxh));  // dh += dGh @ W_hh
        attn::BwdArgs ab{};
        ab.dXp = dXx; ab.n_p = w.pl_dxx.splits; ab.p_stride = (long long)nb * E; ab.p_ld = E;
        ab.V = w.feats + (size_t)b0 * Tn * E; ab.v_bs = (long long)Tn * E; ab.v_ts = E;
        ab.Wh = w.Wh + r * A; ab.Uv = w.Uv + (size_t)b0 * Tn * A; ab.uv_bs = (long long)Tn * A; ab.uv_ts = A;
        ab.attn_b = p.attn_b; ab.attn_w = p.attn_w; ab.B = nb; ab.Tn = Tn; ab.A = A; ab.D = E; ab.inv_T = 1.f / Tn;
        ab.dWh_out = w.dWh + r * A; ab.dWh_op = w.dWh_op + r * A; ab.dUv_acc = w.dUv + (size_t)b0 * Tn * A;
        ab.uv_first = last ? 1 : 0; ab.dw_first = last ? 1 : 0; ab.dw_acc = w.dw_acc + (size_t)b0 * A;
        ab.dctx_out = nullptr; ab.de_out = nullptr; ab.p_drop = 0.f;
        RN_TRY(em.attn_bwd(ab));
        if (t > 0) RN_TRY(em.gemm_partials(w.dWh_op + r * A, A, 0, w.Wa, H, 1, dQp, nb, H, A, w.pl_dq));
        continue;
      }
      cell::BwdArgs cb{};
      cb.dh_ext = w.dHext + r * H; cb.dh_ld = H; cb.dh_scale = nullptr;
      cb.dh_ext2 = g_hiddens ? g_hiddens + r * H : nullptr; cb.dh2_ld = H;
      cb.dXp = last ? nullptr : dXp; cb.n_p = w.pl_dx.splits; cb.p_stride = (long long)nb * w.KX; cb.p_ld = w.KX; cb.col0 = E;
      cb.dQp = last ? nullptr : dQp; cb.n_q = w.pl_dq.splits; cb.q_stride = (long long)nb * H; cb.q_ld = H;
      cb.dc = w.dc + (size_t)b0 * H; cb.first = last ? 1 : 0;
      cb.gates = w.gates + r * 4 * H;
      cb.c_prev = w.c + r * H; cb.c_new = w.c + ((size_t)(t + 1) * B + b0) * H;
      cb.B = nb; cb.H = H; cb.dG = w.dG + r * 4 * H; cb.dg_ld = 4 * H;
      RN_TRY(em.cell_bwd(cb));
      // d[ctx ; h_{t-1}] = dG_t @ [W_ctx | W_hh]
      RN_TRY(em.gemm_partials(w.dG + r * 4 * H, 4 * H, 0, w.Wrec, w.KX, 1, dXp, nb, w.KX, 4 * H, w.pl_dx));
      attn::BwdArgs ab{};
      ab.dXp = dXp; ab.n_p = w.pl_dx.splits; ab.p_stride = (long long)nb * w.KX; ab.p_ld = w.KX;
      ab.V = w.feats + (size_t)b0 * Tn * E; ab.v_bs = (long long)Tn * E; ab.v_ts = E;
      ab.Wh = w.Wh + r * A; ab.Uv = w.Uv + (size_t)b0 * Tn * A; ab.uv_bs = (long long)Tn * A; ab.uv_ts = A;
      ab.attn_b = p.attn_b; ab.attn_w = p.attn_w; ab.B = nb; ab.Tn = Tn; ab.A = A; ab.D = E; ab.inv_T = 1.f / Tn;
      ab.dWh_out = w.dWh + r * A; ab.dWh_op = w.dWh_op + r * A; ab.dUv_acc = w.dUv + (size_t)b0 * Tn * A;
      ab.uv_first = last ? 1 : 0; ab.dw_first = last ? 1 : 0; ab.dw_acc = w.dw_acc + (size_t)b0 * A;
      ab.dctx_out = nullptr; ab.de_out = nullptr; ab.p_drop = 0.f;
      RN_TRY(em.attn_bwd(ab));
      // attention-query path into h_{t-1}: dWh_t @ attn_W, consumed (as split-K partials) by the next cell backward
      if (t > 0) RN_TRY(em.gemm_partials(w.dWh_op + r * A, A, 0, w.Wa, H, 1, dQp, nb, H, A, w.pl_dq));
    }
  }
  RN_TRY(em0.flush(w.table, w.table_bytes, w.bar, w.err, 2));
  if (w.nch > 1) RN_TRY(cs.join(st, w.nch));
  // ---- batched weight gradients over the stashed operands ----------------------------------------------------
  const long long ldih = EMB + E;
  // LSTM: one gate-gradient matrix dG [LB,4H]; GRU: dGi (input side) in w.dG and dGh (hidden side, n column scaled by r) in w.dG2
  const T* dGh = is_gru ? w.dG2 : w.dG;
  RN_TRY(misc::colsum<T>(w.dG, GH, LB, GH, g.b_ih, 0, w.splitk, st));
  if (is_gru) { RN_TRY(misc::colsum<T>(dGh, GH, LB, GH, g.b_hh, 0, w.splitk, st)); }
  else RN_CUDA_OK(cudaMemcpyAsync(g.b_hh, g.b_ih, (size_t)GH * sizeof(float), cudaMemcpyDeviceToDevice, st));
  RN_TRY(gemm_full<T>(w.dG, GH, 1, w.X, w.KX, 1, g.w_ih + EMB, ldih, nullptr, GH, E, LB, 0, w.splitk, st));            // dW_ctx
  RN_TRY(gemm_full<T>(dGh, GH, 1, w.X + E, w.KX, 1, g.w_hh, H, nullptr, GH, H, LB, 0, w.splitk, st));                   // dW_hh
  RN_TRY(gemm_full<T>(w.dG, GH, 1, w.Xe, w.EMBp, 1, g.w_ih, ldih, nullptr, GH, EMB, LB, 0, w.splitk, st));              // dW_emb
  RN_TRY(gemm_full<T>(w.dG, GH, 0, w.Wemb, w.EMBp, 1, w.dXe, w.EMBp, nullptr, LB, EMB, GH, 0, w.splitk, st));           // dXe
  RN_CUDA_OK(cudaMemsetAsync(g.embedding, 0, (size_t)V * EMB * sizeof(float), st));
  misc::embed_scatter_kernel<<<LB, 128, 0, st>>>(g.embedding, tokens_in, w.dXe, w.EMBp, LB, EMB, V, d.embedding_scale, p_emb, rng,
                                                 SITE_EMB);
  RN_LAUNCH_OK();
  // attention parameters
  RN_TRY(gemm_full<T>(w.dWh_op, A, 1, w.X + E, w.KX, 1, g.attn_W, H, nullptr, A, H, LB, 0, w.splitk, st));              // dW_a = dWh^T h_{t-1}
  RN_TRY(misc::cast_pad<T>(w.dUv, A, w.dUv_op, A, (long long)B * Tn, A, A, st));
  RN_TRY(gemm_full<T>(w.dUv_op, A, 1, w.feats, E, 1, g.attn_U, E, nullptr, A, E, B * Tn, 0, w.splitk, st));             // dU = dUv^T v
  RN_TRY(misc::colsum<float>(w.dWh, A, LB, A, g.attn_b, 0, w.splitk, st));
  RN_TRY(misc::colsum<float>(w.dw_acc, A, B, A, g.attn_w, 0, w.splitk, st));
  return 0;
}

// ---- greedy decoding (eval.py:19-33) ---------------------------------------------------------------------------
// argmax over the vocabulary (lowest index wins ties, like topk(1)), feeds the id back as next token, all on device.
__global__ void argmax_feedback_kernel(const float* __restrict__ logits, long long ld, int V, long long* __restrict__ ids_out,
                                       long long* __restrict__ next_tok, int* __restrict__ nonpad_count) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const int b = blockIdx.x;
  const float* z = logits + (long long)b * ld;
  float best = -INFINITY; int bi = 0x7fffffff;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float x = z[v];
    if (x > best || (x == best && v < bi)) { best = x; bi = v; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) { sv[wp] = best; si[wp] = bi; }
  __syncthreads();
  if (wp == 0) {
    best = lane < nw ? sv[lane] : -INFINITY; bi = lane < nw ? si[lane] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
      ids_out[b] = bi; next_tok[b] = bi;
      if (bi != 0) atomicAdd(nonpad_count, 1);
    }
  }
}
// n_steps = first step (1-based) after which every fed-back token was <PAD> (eval.py:30), else max_steps
__global__ void greedy_finalize_kernel(const int* __restrict__ nonpad, int max_steps, int* __restrict__ n_steps) {
  int n = max_steps;
  for (int t = 0; t < max_steps; ++t) if (nonpad[t] == 0) { n = t + 1; break; }
  *n_steps = n;
}
__global__ void fill_tokens_kernel(long long* tok, int n, long long v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tok[i] = v;
}

template <typename T>
struct GreedyWs { Ws<T> w; long long* tok; int* nonpad; float* h_scratch; size_t bytes; };

template <typename T>
static GreedyWs<T> plan_greedy(const recnet_decoder_desc& d, void* base, int max_steps) {
  GreedyWs<T> g;
  recnet_decoder_desc d1 = d; d1.L = 1;
  g.w = plan<T>(d1, base);
  Bump m(base); m.off = g.w.bytes;
  g.tok = m.take<long long>(d.B);
  g.nonpad = m.take<int>(max_steps);
  g.h_scratch = m.take<float>((size_t)d.B * d.H);
  g.bytes = m.off + 256;
  return g;
}

template <typename T>
static int greedy(const recnet_decoder_desc& d0, const recnet_decoder_tensors& p, const float* feats, int max_steps, void* ws,
                  long long ws_bytes, long long* ids_out, int* n_steps_out, cudaStream_t st) {
  recnet_decoder_desc d = d0; d.L = 1; d.train = 0;
  RN_TRY(check(d));
  GreedyWs<T> g = plan_greedy<T>(d0, ws, max_steps);
  if ((long long)g.bytes > ws_bytes) return RECNET_ERR_WORKSPACE;
  Ws<T>& w = g.w;
  const int B = d.B, H = d.H, E = d.E, A = d.A, V = d.V, Tn = d.T;
  RN_TRY(prepare<T>(d, p, feats, w, st));
  RN_TRY(gemm_full<T>(w.feats, E, 0, w.U, E, 0, w.Uv, A, nullptr, B * Tn, A, E, 0, w.splitk, st));
  RN_CUDA_OK(cudaMemsetAsync(w.X, 0, (size_t)2 * B * w.KX * sizeof(T), st));
  RN_CUDA_OK(cudaMemsetAsync(w.c, 0, (size_t)2 * B * H * sizeof(float), st));
  RN_CUDA_OK(cudaMemsetAsync(g.nonpad, 0, (size_t)max_steps * sizeof(int), st));
  fill_tokens_kernel<<<rn_cdiv(B, 128), 128, 0, st>>>(g.tok, B, 1 /* <SOS> */);
  RN_LAUNCH_OK();
  // ping-pong the two X rows / c rows of the L=1 plan
  for (int t = 0; t < max_steps; ++t) {
    T* x_t = w.X + (size_t)(t & 1) * B * w.KX;
    T* x_n = w.X + (size_t)((t + 1) & 1) * B * w.KX;
    float* c_p = w.c + (size_t)(t & 1) * B * H;
    float* c_n = w.c + (size_t)((t + 1) & 1) * B * H;
    misc::embed_gather_kernel<T><<<B, 128, 0, st>>>(p.embedding, g.tok, w.Xe, w.EMBp, B, d.EMB, w.EMBp, V, d.embedding_scale, 0.f,
                                                    nullptr, SITE_EMB);
    RN_LAUNCH_OK();
    RN_TRY(gemm_full<T>(w.Xe, w.EMBp, 0, w.Wemb, w.EMBp, 0, w.Gx, w.G * H, p.b_ih, B, w.G * H, w.EMBp, 0, w.splitk, st));
    mega::Emitter<T> em(false, st);
    // GRU keeps its fp32 state h in the c ping-pong rows (h_prev = c_p, h' -> c_n)
    RN_TRY(step<T>(em, d, p, w, t, 0, 0, B, w.Gx, x_t, x_n, nullptr, nullptr, nullptr, c_p, c_n,
                   d.cell == RECNET_CELL_GRU ? c_n : g.h_scratch));
    RN_TRY(gemm_full<T>(x_n + E, w.KX, 0, w.Wout, H, 0, w.logits, w.Vld, p.out_b, B, V, H, 0, w.splitk, st));
    argmax_feedback_kernel<<<B, 256, 0, st>>>(w.logits, w.Vld, V, ids_out + (size_t)t * B, g.tok, g.nonpad + t);
    RN_LAUNCH_OK();
  }
  greedy_finalize_kernel<<<1, 1, 0, st>>>(g.nonpad, max_steps, n_steps_out);
  RN_LAUNCH_OK();
  return 0;
}
// ---- beam search (eval.beam_search, eval.py:36-120) as a device loop: zero host syncs ------------------------------------------
// Rows of every per-step tensor are (beam k, sample b) -> k * B0 + b; the caller passes the features tiled K times.  Per step: the
// decoder step on all K * B0 rows (same kernels as greedy), then ONE selection kernel per sample -- scores log(sigmoid(logit)) +
// cum / len^0.7 (eval.py:53-61: the running score is divided by the length at EVERY step; len = position of the last <EOS> + 1 once
// the beam has emitted one, else t + 1), top-K over the K_cur * V candidates, sequence / <EOS> / score bookkeeping -- and one gather
// kernel that moves the recurrent state of the source beams into place.  The loop freezes on the device once every fed-back token of a
// step is <PAD> (eval.py:116); the host never reads anything back until the end.
constexpr int BEAM_MAX_K = 8;
constexpr int BEAM_THREADS = 256;

__global__ void beam_select_kernel(const float* __restrict__ logits, long long ld, int V, int B0, int K, int t, int max_steps, long long eos,
                                   const float* __restrict__ cum_in, float* __restrict__ cum_out, const int* __restrict__ eos_in,
                                   int* __restrict__ eos_out, const int* __restrict__ seq_in, int* __restrict__ seq_out,
                                   long long* __restrict__ tok_out, int* __restrict__ src_out, int* __restrict__ nonpad) {
  if (t > 0 && nonpad[t - 1] == 0) return;                  // frozen: every token fed back at the previous step was <PAD>
  __shared__ float sv[BEAM_THREADS / 32];
  __shared__ int si[BEAM_THREADS / 32];
  __shared__ int chosen[BEAM_MAX_K];
  __shared__ float base[BEAM_MAX_K];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Kc = t == 0 ? 1 : K;                            // one live beam before the first expansion
  if (tid < Kc) {
    const int le = eos_in[b * K + tid];
    const float len = le >= 0 ? (float)(le + 1) : (float)(t + 1);
    base[tid] = cum_in[tid * B0 + b] / powf(len, 0.7f);
  }
  __syncthreads();
  const int total = Kc * V;
  int n_nonpad = 0;
  for (int k = 0; k < K; ++k) {
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int f = tid; f < total; f += BEAM_THREADS) {
      bool taken = false;
      for (int q = 0; q < k; ++q) taken |= (chosen[q] == f);
      if (taken) continue;
      const int i = f / V, v = f - i * V;
      const float x = logits[(long long)(i * B0 + b) * ld + v];
      const float sc = logf(1.f / (1.f + expf(-x))) + base[i];
      if (sc > best || (sc == best && f < bi)) { best = sc; bi = f; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { sv[warp] = best; si[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      best = lane < BEAM_THREADS / 32 ? sv[lane] : -INFINITY; bi = lane < BEAM_THREADS / 32 ? si[lane] : 0x7fffffff;
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) {
        chosen[k] = bi;
        const int i = bi / V, v = bi - i * V;
        tok_out[k * B0 + b] = v;
        src_out[b * K + k] = i;
        cum_out[k * B0 + b] = best;
        eos_out[b * K + k] = (v == (int)eos) ? t : eos_in[b * K + i];
        if (v != 0) ++n_nonpad;
      }
    }
    __syncthreads();
  }
  // sequences: beam k continues beam src[k]
  for (int x = tid; x < K * (t + 1); x += BEAM_THREADS) {
    const int k = x / (t + 1), pos = x - k * (t + 1);
    const int f = chosen[k], i = f / V, v = f - i * V;
    seq_out[(b * K + k) * max_steps + pos] = pos < t ? seq_in[(b * K + i) * max_steps + pos] : v;
  }
  if (tid == 0 && n_nonpad) atomicAdd(nonpad + t, n_nonpad);
}

// recurrent state of row (k, b) <- state of row (src[b][k], b): h operand rows (ld elements apart) and the fp32 state rows
template <typename T>
__global__ void beam_gather_kernel(const T* __restrict__ h_src, T* __restrict__ h_dst, long long ld, const float* __restrict__ c_src,
                                   float* __restrict__ c_dst, int H, int B0, int K, const int* __restrict__ src, const int* __restrict__ nonpad, int t) {
  if (t > 0 && nonpad[t - 1] == 0) return;                  // frozen (see beam_select_kernel)
  const int row = blockIdx.x, k = row / B0, b = row - k * B0;
  const int srow = src[b * K + k] * B0 + b;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    h_dst[(long long)row * ld + j] = h_src[(long long)srow * ld + j];
    c_dst[(long long)row * H + j] = c_src[(long long)srow * H + j];
  }
}
__global__ void beam_finalize_kernel(const int* __restrict__ nonpad, int max_steps, int B0, int K, const int* __restrict__ seq0,
                                     const int* __restrict__ seq1, long long* __restrict__ out, int* __restrict__ n_steps) {
  int n = max_steps;
  for (int t = 0; t < max_steps; ++t) if (nonpad[t] == 0) { n = t + 1; break; }
  const int* seq = (n & 1) ? seq1 : seq0;                   // step t wrote buffer (t + 1) & 1; the last executed step is n - 1
  for (int x = threadIdx.x; x < B0 * max_steps; x += blockDim.x) {
    const int b = x / max_steps, pos = x - b * max_steps;
    out[x] = pos < n ? seq[(b * K) * max_steps + pos] : -1;   // top-1 beam (eval.py:118-119)
  }
  if (threadIdx.x == 0) *n_steps = n;
}
__global__ void beam_init_kernel(float* cum, int* eos, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { cum[i] = 0.f; eos[i] = -1; }                  // log(1.) ; no <EOS> yet
}

template <typename T>
struct BeamWs { GreedyWs<T> g; float* cum[2]; int* eos[2]; int* seq[2]; int* src; size_t bytes; };
template <typename T>
static BeamWs<T> plan_beam(const recnet_decoder_desc& d, void* base, int K, int max_steps) {
  BeamWs<T> w;
  w.g = plan_greedy<T>(d, base, max_steps);
  Bump m(base); m.off = w.g.bytes;
  const int B0 = d.B / (K > 0 ? K : 1);
  for (int i = 0; i < 2; ++i) {
    w.cum[i] = m.take<float>((size_t)d.B);
    w.eos[i] = m.take<int>((size_t)d.B);
    w.seq[i] = m.take<int>((size_t)d.B * max_steps);
  }
  w.src = m.take<int>((size_t)d.B);
  (void)B0;
  w.bytes = m.off + 256;
  return w;
}

// d0.B = K * B0 rows; feats [K * B0, T, E] = the B0 samples tiled K times; seq_out [B0, max_steps] int64 (-1 beyond n_steps)
template <typename T>
static int beam(const recnet_decoder_desc& d0, const recnet_decoder_tensors& p, const float* feats, int K, int max_steps, long long eos,
                void* ws, long long ws_bytes, long long* seq_out, int* n_steps_out, cudaStream_t st) {
  recnet_decoder_desc d = d0; d.L = 1; d.train = 0;
  RN_TRY(check(d));
  if (K < 1 || K > BEAM_MAX_K || d.B % K || max_steps < 1 || max_steps > 64 || K > d.V) return RECNET_ERR_BAD_SHAPE;
  BeamWs<T> bw = plan_beam<T>(d0, ws, K, max_steps);
  if ((long long)bw.bytes > ws_bytes) return RECNET_ERR_WORKSPACE;
  GreedyWs<T>& g = bw.g;
  Ws<T>& w = g.w;
  const int B = d.B, B0 = B / K, H = d.H, E = d.E, A = d.A, V = d.V, Tn = d.T;
  RN_TRY(prepare<T>(d, p, feats, w, st));
  RN_TRY(gemm_full<T>(w.feats, E, 0, w.U, E, 0, w.Uv, A, nullptr, B * Tn, A, E, 0, w.splitk, st));
  RN_CUDA_OK(cudaMemsetAsync(w.X, 0, (size_t)2 * B * w.KX * sizeof(T), st));
  RN_CUDA_OK(cudaMemsetAsync(w.c, 0, (size_t)2 * B * H * sizeof(float), st));
  RN_CUDA_OK(cudaMemsetAsync(g.nonpad, 0, (size_t)max_steps * sizeof(int), st));
  fill_tokens_kernel<<<rn_cdiv(B, 128), 128, 0, st>>>(g.tok, B, 1 /* <SOS> */);
  RN_LAUNCH_OK();
  beam_init_kernel<<<rn_cdiv(B, 128), 128, 0, st>>>(bw.cum[0], bw.eos[0], B);
  RN_LAUNCH_OK();
  // state buffers: step reads (x0, c0) and writes (x1, c1); the gather moves the selected beams' state back into (x0, c0)
  T* x0 = w.X; T* x1 = w.X + (size_t)B * w.KX;
  float* c0 = w.c; float* c1 = w.c + (size_t)B * H;
  for (int t = 0; t < max_steps; ++t) {
    misc::embed_gather_kernel<T><<<B, 128, 0, st>>>(p.embedding, g.tok, w.Xe, w.EMBp, B, d.EMB, w.EMBp, V, d.embedding_scale, 0.f,
                                                    nullptr, SITE_EMB);
    RN_LAUNCH_OK();
    RN_TRY(gemm_full<T>(w.Xe, w.EMBp, 0, w.Wemb, w.EMBp, 0, w.Gx, w.G * H, p.b_ih, B, w.G * H, w.EMBp, 0, w.splitk, st));
    mega::Emitter<T> em(false, st);
    RN_TRY(step<T>(em, d, p, w, t, 0, 0, B, w.Gx, x0, x1, nullptr, nullptr, nullptr, c0, c1, d.cell == RECNET_CELL_GRU ? c1 : g.h_scratch));
    RN_TRY(gemm_full<T>(x1 + E, w.KX, 0, w.Wout, H, 0, w.logits, w.Vld, p.out_b, B, V, H, 0, w.splitk, st));
    const int in = t & 1, out = (t + 1) & 1;
    beam_select_kernel<<<B0, BEAM_THREADS, 0, st>>>(w.logits, w.Vld, V, B0, K, t, max_steps, eos, bw.cum[in], bw.cum[out], bw.eos[in], bw.eos[out],
                                                    bw.seq[in], bw.seq[out], g.tok, bw.src, g.nonpad);
    RN_LAUNCH_OK();
    beam_gather_kernel<T><<<B, 128, 0, st>>>(x1 + E, x0 + E, w.KX, c1, c0, H, B0, K, bw.src, g.nonpad, t);
    RN_LAUNCH_OK();
  }
  beam_finalize_kernel<<<1, 256, 0, st>>>(g.nonpad, max_steps, B0, K, bw.seq[0], bw.seq[1], seq_out, n_steps_out);
  RN_LAUNCH_OK();
  return 0;
}
}  // namespace dec
