// Weight-resident persistent time loop of the LOCAL reconstructor (forward), one cooperative launch for all S steps.
//   reference: models/local_reconstructor.py:37-55 called S times from train.py:122-124
//
// Why (profiles/r2_a_segments.md): as a chain of 4 kernels per step (query GEMM, attention, gate GEMM, cell) the loop costs
// ~22 us per step, of which ~1.3 us is tensor-pipe time; the rest is kernel boundaries and the 25 MB [W_ih | W_hh] operand being
// re-pulled from L2 every step.  Here the weights are loaded ONCE into the shared memory of UG x KS CTAs and stay there:
//
//   CTA (ug, ks):  32 hidden units x 4 gates (128 accumulator columns)  x  K-slice ks of [x_t ; h_{t-1}]   (<= 11 k-blocks = 176 KB)
//
// Per step:  [h_{t-1} visible] -> recurrent half of the gate GEMM (tcgen05, accumulates in TMEM) runs while 8 warps of the
// first B CTAs do the attention of one sample each -> [x_t visible] -> input half of the GEMM into the same accumulator ->
// fp32 K-slice partials exchanged through L2 inside the unit group (KS CTAs, one flag) -> cell update, h_t written as next
// step's operand row, and the partial attention query  Wh_{t+1}[b, :] += h_t[b, units] . W_a[:, units]^T  of these 32 units
// (CUDA cores, W_a slice in registers) -> [h_t visible].  Two grid-wide flag waits + one unit-group flag per step; nothing
// else leaves the SM.  The stash (x_t / h_t operand rows, Wh, scores, activated gates, c) is laid out exactly as the
// kernel-per-phase path leaves it, so the batched GEMMs around the loop and the BPTT run unchanged behind it.
//
// Every spin is bounded (clock64 timeout -> error flag in the workspace, all later waits fall through), so a protocol bug
// surfaces as RECNET_ERR_* instead of a hung GPU.
#pragma once
#include "runtime.cuh"

namespace rp {
using namespace tc;

constexpr int THREADS = 448;          // warps 0-3 epilogue, 4 TMA producer, 5 MMA issuer, 6-13 attention
constexpr int STAGES = 3;             // activation ring: B rows x 64 k per stage
constexpr int UNITS = 32;             // hidden units per CTA
constexpr int NCOL = 4 * UNITS;       // accumulator columns (gate g, unit u) -> column g*32 + u
constexpr int MAX_KB = 11;            // resident k-blocks per CTA (16 KB each)
constexpr int WTILE = NCOL * BK * 2;  // 16384
constexpr int NWORK = 12;             // warps in the cell / query phase
constexpr int MAX_UG = 64;
constexpr int HS_LD = 40;             // bf16 elements per hs row (80 B: conflict-free fragment loads)
constexpr long long TIMEOUT = 600000000LL;      // ~0.3 s of SM clocks

struct FwdParams {
  int B, S, R, H, A, L, KX, UG, KS, NHB, NXB;
  float inv_L, p_drop;
  bf16* X;                            // [(S+1)*B, KX] operand rows [x_t | h_{t-1}]
  const bf16* Hd;                     // [L*B, H] decoder states (attention values)
  const float* Uv;                    // [L, B, A]
  const bf16* Wa;                     // [A, R]
  const float *attn_b, *attn_w, *b_ih, *b_hh;
  float* Wh;                          // [S, B, A] stash (query without the bias)
  float* beta;                        // [S, B, L] stash (scores)
  bf16* gates;                        // [S*B, 4R] activated gates
  float* c;                           // [(S+1)*B, R]
  float* XP;                          // [UG][KS][B][128] K-slice partials of the pre-activations
  float* WhP;                         // [UG][B][A] unit-group partials of the next step's query
  unsigned* sync;                     // [0] h-ready counter, [32] x-ready counter, [64 + 32*ug] exchange flags (zeroed before launch)
  int* err;
  const unsigned long long* rng;
  unsigned site;
};

__device__ __forceinline__ unsigned ld_acq(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_rel(unsigned* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void nbar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// bounded waits; once *abort is set every later wait falls through (all bar.sync's are still executed -> no CTA-level hang)
__device__ __forceinline__ void wait_flag(const unsigned* f, unsigned target, int* err, volatile int* abort_) {
  if (*abort_) return;
  const long long t0 = clock64();
  unsigned spins = 0;
  while (ld_acq(f) < target) {
    if ((++spins & 31u) == 0) {
      if (*reinterpret_cast<volatile int*>(err) != 0) { *abort_ = 1; return; }
      if (clock64() - t0 > TIMEOUT) { atomicExch(err, 3); *abort_ = 1; return; }
    }
  }
}
__device__ __forceinline__ void wait_mbar(uint32_t bar, uint32_t parity, int* err, volatile int* abort_) {
  if (*abort_) return;
  const long long t0 = clock64();
  uint32_t ok;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > TIMEOUT) { atomicExch(err, 2); *abort_ = 1; return; }
    if (*abort_) return;
  }
}
// developer timeline: any single thread of block 0 (atomic slot)
__device__ __forceinline__ void stamp(int tag) {
  if (g_timeline != nullptr && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned int i = atomicAdd(&g_timeline_n, 1u);
    if (i < 4000) g_timeline[i] = (t & 0x00FFFFFFFFFFFFFFull) | ((unsigned long long)tag << 56);
  }
}

// ---- TMEM as a per-warp register file extension: 32 lanes x n columns (32-bit) of the warp's own lane quarter ----
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// loads WITHOUT the wait (several can be in flight); tmem_ld_wait() before the registers are read
__device__ __forceinline__ void tmem_ld32_nw(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_nw(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
constexpr int TMEM_COLS = 256;        // [0,128) gate accumulator, [128, 224) attention operands parked by the attention warps

struct Smem {                // offsets from the 1024-aligned base
  int ring, wres, bars, e, whred, hs, total;
};
__host__ __device__ inline Smem smem_layout(int B, int KS) {
  Smem s;
  const int stage = ((B + 7) & ~7) * 128;
  s.ring = 0;
  s.wres = (STAGES * stage + 1023) & ~1023;
  s.bars = s.wres + MAX_KB * WTILE;               // full[3] empty[3] tmem wload | tmem slot | abort
  s.e = s.bars + 128;
  s.whred = s.e + 128;                            // [8][128] floats
  s.hs = s.whred + 8 * 128 * 4;                   // [48][HS_LD] bf16: h_t of this CTA's cells, A operand of the query mma
  s.total = s.hs + 4 * NWORK * HS_LD * 2 + 1024;  // + alignment slack
  return s;
}

// per-thread constants of the cell / query phase
struct CellCtx {
  int ww, lane, ug, ks, ns, j, tid, cta;
  float bi, bf_, bg, bo;
  uint32_t wb[8];                     // W_a fragments (B operand of mma.m16n8k16) of a-tiles 2ww, 2ww+1 over this CTA's 32 units
};

// cell update of this CTA's (sample, unit) cells + partial attention query of the next step + stash; all NWORK warps
template <int KS>
__device__ __forceinline__ void cell_query_phase(const FwdParams& p, const CellCtx& cq, int t, float (&creg)[4], bf16* hs) {
  const int ww = cq.ww, lane = cq.lane, ug = cq.ug, ks = cq.ks, ns = cq.ns, j = cq.j, tid = cq.tid, cta = cq.cta;
  const int B = p.B, S = p.S, R = p.R, H = p.H, A = p.A, KX = p.KX;
  const float bi = cq.bi, bf_ = cq.bf_, bg = cq.bg, bo = cq.bo;
  unsigned* barA = p.sync;
      nbar(2, NWORK * 32);
      if (cta == 0 && tid == 0) stamp(7);
      // ---------------- cell update of (sample, unit): lane = unit, warp ww takes sample slots i = ww + 12 r ----------------
      uint32_t gp_if[4], gp_go[4];                        // activated gates packed as bf16 pairs (i,f) (g,o): what the stash keeps
#pragma unroll
      for (int r0 = 0; r0 < 4; r0 += 2) {
        float pre[2][4];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int r3 = r0 + rr;
          const int i = min(ww + r3 * NWORK, ns - 1), b = ks + KS * i;
          const float* q = p.XP + ((long long)ug * KS * B + b) * NCOL + lane;
          float part[KS][4];
#pragma unroll
          for (int k2 = 0; k2 < KS; ++k2)
#pragma unroll
            for (int g = 0; g < 4; ++g) part[k2][g] = q[(long long)k2 * B * NCOL + g * UNITS];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float s = part[0][g];
#pragma unroll
            for (int k2 = 1; k2 < KS; ++k2) s += part[k2][g];
            pre[rr][g] = s;
          }
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int r3 = r0 + rr;
          const int i = ww + r3 * NWORK;
          const float gi = act_sigmoid<true>(pre[rr][0] + bi), gf = act_sigmoid<true>(pre[rr][1] + bf_);
          const float gg = act_tanh<true>(pre[rr][2] + bg), go = act_sigmoid<true>(pre[rr][3] + bo);
          { const __nv_bfloat162 x = __floats2bfloat162_rn(gi, gf), y = __floats2bfloat162_rn(gg, go);
            gp_if[r3] = *reinterpret_cast<const uint32_t*>(&x); gp_go[r3] = *reinterpret_cast<const uint32_t*>(&y); }
          if (i < ns) {
            const int b = ks + KS * i;
            creg[r3] = fmaf(gf, creg[r3], gi * gg);
            const bf16 hb = __float2bfloat16_rn(go * act_tanh<true>(creg[r3]));
            p.X[((long long)(t + 1) * B + b) * KX + H + j] = hb;          // next step's operand row
            hs[i * HS_LD + lane] = hb;
          }
        }
        if (ww + (r0 + 2) * NWORK >= ns) break;                          // warp-uniform: no further sample slots
      }
      if (cta == 0 && tid == 0) stamp(8);
      nbar(3, NWORK * 32);
      // ---------------- partial attention query of the next step over this CTA's 32 units ----------------
      // WhP[ug][b, a] = sum_u h_t[b, u] W_a[a, u]: [ns x 32] . [32 x 128] on mma.sync m16n8k16 (bf16 in, fp32 out); warp ww < 8 owns
      // a-tiles 2ww, 2ww+1 (its W_a fragments live in registers), loops over the sample tiles; rows >= ns are never stored
      if (t + 1 < S && ww < 8) {
        const int gid = lane >> 2, tig = lane & 3;
        for (int mt = 0; mt * 16 < ns; ++mt) {
          uint32_t af[2][4];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const bf16* r0 = hs + (mt * 16 + gid) * HS_LD + 16 * k + tig * 2;
            af[k][0] = *reinterpret_cast<const uint32_t*>(r0);
            af[k][1] = *reinterpret_cast<const uint32_t*>(r0 + 8 * HS_LD);
            af[k][2] = *reinterpret_cast<const uint32_t*>(r0 + 8);
            af[k][3] = *reinterpret_cast<const uint32_t*>(r0 + 8 * HS_LD + 8);
          }
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
            for (int k = 0; k < 2; ++k)
              asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                           : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3)
                           : "r"(af[k][0]), "r"(af[k][1]), "r"(af[k][2]), "r"(af[k][3]), "r"(cq.wb[4 * nt + 2 * k]), "r"(cq.wb[4 * nt + 2 * k + 1]));
            const int a = 16 * ww + 8 * nt + tig * 2;
            const int i0 = mt * 16 + gid, i1 = i0 + 8;
            if (i0 < ns) *reinterpret_cast<float2*>(p.WhP + ((long long)ug * B + (ks + KS * i0)) * A + a) = make_float2(d0, d1);
            if (i1 < ns) *reinterpret_cast<float2*>(p.WhP + ((long long)ug * B + (ks + KS * i1)) * A + a) = make_float2(d2, d3);
          }
        }
      }
      if (cta == 0 && tid == 0) stamp(9);
      nbar(4, NWORK * 32);
      if (tid == 0) { if (cta == 0) stamp(15); red_rel(barA); if (cta == 0) stamp(6); }
      // ---------------- stash for the BPTT (activated gates, c): only has to land by the end of the kernel ----------------
#pragma unroll
      for (int r3 = 0; r3 < 4; ++r3) {
        const int i = ww + r3 * NWORK;
        if (i < ns) {
          const int b = ks + KS * i;
          p.c[((long long)(t + 1) * B + b) * R + j] = creg[r3];
          unsigned short* gs = reinterpret_cast<unsigned short*>(p.gates + ((long long)t * B + b) * 4 * R + j);
          gs[0] = (unsigned short)(gp_if[r3] & 0xFFFFu); gs[R] = (unsigned short)(gp_if[r3] >> 16);
          gs[2 * R] = (unsigned short)(gp_go[r3] & 0xFFFFu); gs[3 * R] = (unsigned short)(gp_go[r3] >> 16);
        }
      }
}

template <int KS>
__global__ void __launch_bounds__(THREADS, 1)
local_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const Smem L_ = smem_layout(p.B, KS);
  const int stage_bytes = ((p.B + 7) & ~7) * 128;
  const uint32_t ring = base + L_.ring, wres = base + L_.wres;
  const uint32_t bar_full = base + L_.bars, bar_empty = bar_full + 8 * STAGES, bar_tmem = bar_empty + 8 * STAGES, bar_w = bar_tmem + 8;
  const uint32_t tmem_slot = bar_w + 8;
  volatile int* abort_ = reinterpret_cast<volatile int*>(gen + L_.bars + 96);
  float* e_s = reinterpret_cast<float*>(gen + L_.e);
  float* whred = reinterpret_cast<float*>(gen + L_.whred);
  bf16* hs = reinterpret_cast<bf16*>(gen + L_.hs);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x;
  const int ug = cta / KS, ks = cta % KS;
  const int B = p.B, S = p.S, R = p.R, H = p.H, A = p.A, Ln = p.L, KX = p.KX;
  const int hb0 = (p.NHB * ks) / KS, hb1 = (p.NHB * (ks + 1)) / KS, nh = hb1 - hb0;
  const int xb0 = (p.NXB * ks) / KS, xb1 = (p.NXB * (ks + 1)) / KS, nx = xb1 - xb0;
  const int nkb = nh + nx;
  const unsigned ncta = gridDim.x;
  unsigned* barA = p.sync;
  unsigned* barB = p.sync + 32;
  unsigned* flagX = p.sync + 64 + 32 * ug;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_tmem, 1);
    mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *abort_ = 0;
  }
  if (warp == 5) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 4) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
      // resident weights: k-block i of this CTA = 4 boxes {64 k, 32 rows} (gate g rows g*R + ug*32 ..) -> tile rows g*32 + u
      mbar_expect_tx(bar_w, (uint32_t)nkb * WTILE);
      for (int i = 0; i < nkb; ++i) {
        const int kcol = (i < nh ? p.NXB + hb0 + i : xb0 + (i - nh)) * BK;
#pragma unroll
        for (int g = 0; g < 4; ++g) tma_load_2d(wres + i * WTILE + g * (UNITS * 128), &tmW, bar_w, kcol, g * R + ug * UNITS);
      }
      uint32_t it = 0;
      const uint32_t tx = (uint32_t)B * 128u;
      for (int t = 0; t < S; ++t) {
        if (t > 0) wait_flag(barA, ncta * (unsigned)t, p.err, abort_);          // h_{t-1} rows of every unit group are visible
        proxy_fence();
        if (cta == 0) stamp(1);
        for (int i = 0; i < nh; ++i, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
          wait_mbar(bar_empty + 8 * s, ph ^ 1u, p.err, abort_);
          if (*abort_) break;                                                     // never issue bulk copies into a CTA that is draining
          mbar_expect_tx(bar_full + 8 * s, tx);
          tma_load_2d(ring + s * stage_bytes, &tmX, bar_full + 8 * s, (p.NXB + hb0 + i) * BK, t * B);
        }
        wait_flag(barB, (unsigned)B * (unsigned)(t + 1), p.err, abort_);       // x_t of every sample is visible
        proxy_fence();
        if (cta == 0) stamp(2);
        for (int i = 0; i < nx; ++i, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
          wait_mbar(bar_empty + 8 * s, ph ^ 1u, p.err, abort_);
          if (*abort_) break;
          mbar_expect_tx(bar_full + 8 * s, tx);
          tma_load_2d(ring + s * stage_bytes, &tmX, bar_full + 8 * s, (xb0 + i) * BK, t * B);
        }
      }
    }
  } else if (warp == 5) {
    // ================================ MMA issuer ================================
    // D[n = gate row g*32 + u (M = 128), b = sample (N = NB)] += W[n, k] . act[b, k]: weights are the M-side operand, so every accumulator
    // row is useful and an epilogue warp (TMEM lane quarter) holds ONE gate of 32 consecutive units for all samples -> its stores
    // to the exchange buffer are 128-byte rows (the sample-on-M orientation needed 512-byte-strided stores: 2.2 us per step)
    const int NB = (B + 15) & ~15;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    wait_mbar(bar_w, 0, p.err, abort_);
    uint32_t it = 0;
    for (int t = 0; t < S; ++t) {
      for (int i = 0; i < nkb; ++i, ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
        wait_mbar(bar_full + 8 * s, ph, p.err, abort_);
        tc_fence_after();
        if (lane == 0 && !*abort_) {
          const uint32_t sact = ring + s * stage_bytes, sw = wres + i * WTILE;
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk)
            umma_bf16(tmem_base, umma_smem_desc(sw + kk * (UMMA_K * 2), 16, 1024), umma_smem_desc(sact + kk * (UMMA_K * 2), 16, 1024), idesc,
                      (i > 0 || kk > 0) ? 1u : 0u);
          umma_commit(bar_empty + 8 * s);
          if (i == nkb - 1) umma_commit(bar_tmem);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================ worker warps ================================
    // two separate loops (epilogue warps 0-3 / attention warps 6-13) that meet in cell_query_phase: the attention warps keep
    // their time-invariant operands in registers for the whole kernel, the epilogue warps need theirs for tcgen05.ld
    const bool is_epi = warp < 4;
    CellCtx cq;
    cq.ww = is_epi ? warp : warp - 2;                      // 0..11 in the cell / query phase
    cq.lane = lane; cq.ug = ug; cq.ks = ks; cq.tid = tid; cq.cta = cta;
    cq.ns = (B - ks + KS - 1) / KS;                        // samples b = ks + KS*i owned by this CTA in the cell phase (<= 4 * NWORK)
    cq.j = ug * UNITS + lane;                              // this lane's hidden unit
    {
      const int j = cq.j;
      cq.bi = p.b_ih[j] + p.b_hh[j]; cq.bf_ = p.b_ih[R + j] + p.b_hh[R + j];
      cq.bg = p.b_ih[2 * R + j] + p.b_hh[2 * R + j]; cq.bo = p.b_ih[3 * R + j] + p.b_hh[3 * R + j];
      // W_a fragments of the query mma (B operand, col-major k x n): a = 16 ww + 8 nt + lane/4, units 16 k + 2 (lane%4) (+8)
      if (cq.ww < 8) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const bf16* src = p.Wa + (long long)min(16 * cq.ww + 8 * nt + (lane >> 2), A - 1) * R + ug * UNITS + 16 * k + (lane & 3) * 2;
            cq.wb[4 * nt + 2 * k] = *reinterpret_cast<const uint32_t*>(src);
            cq.wb[4 * nt + 2 * k + 1] = *reinterpret_cast<const uint32_t*>(src + 8);
          }
      }
    }
    float creg[4] = {0.f, 0.f, 0.f, 0.f};                  // cell state of this thread's (sample slot r, unit) cells: never re-read from memory

    if (is_epi) {
      for (int t = 0; t < S; ++t) {
        // ---------------- accumulator -> K-slice partials (warp q holds gate q of the 32 units, all samples) ----------------
        wait_mbar(bar_tmem, (uint32_t)t & 1u, p.err, abort_);
        tc_fence_after();
        if (cta == 0 && tid == 0) stamp(4);
        float* dst = p.XP + ((long long)ug * KS + ks) * B * NCOL + warp * UNITS + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < B; c0 += 64) {                // two 32-column chunks in flight per wait
          uint32_t r[64];
          tmem_ld32_nw(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
          tmem_ld32_nw(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c0 + 32), r + 32);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 64; ++q)
            if (c0 + q < B) dst[(long long)(c0 + q) * NCOL] = __uint_as_float(r[q]);
        }
        tc_fence_before();
        if (cta == 0 && tid == 0) stamp(10);
        nbar(1, 128);
        if (tid == 0) {
          red_rel(flagX);
          if (cta == 0) stamp(11);
          wait_flag(flagX, (unsigned)KS * (unsigned)(t + 1), p.err, abort_);     // all KS slices of this unit group have landed
          if (cta == 0) stamp(5);
        }
        cell_query_phase<KS>(p, cq, t, creg, hs);
      }
    } else if (cta >= B) {
      for (int t = 0; t < S; ++t) cell_query_phase<KS>(p, cq, t, creg, hs);
    } else {
      // ---------------- attention of sample b_att = cta (warps 6-13 of the first B CTAs) ----------------
      const int aw = warp - 6, b_att = cta;
      // operands that do not change over the S steps stay in registers: the L value rows of this lane's column pair, U.v of this
      // warp's frames, attn_b / attn_w chunks (lane owns a-chunk [4 lane, 4 lane + 4) of the score dot product)
      const int col = 2 * (aw * 32 + lane);
      const bool col_ok = col < H;
      const float4 ab4 = reinterpret_cast<const float4*>(p.attn_b)[lane], aw4 = reinterpret_cast<const float4*>(p.attn_w)[lane];
      // ... parked in tensor memory (this warp's lane quarter, columns 128 + 48*(aw/4) ..): 48 registers per thread that would
      // otherwise be live across the whole loop; tcgen05.ld brings them back each step (64 B/clk, off the critical loads)
      const uint32_t t_park = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 128u + 48u * (uint32_t)(aw >> 2);
      {
        uint32_t v[32];
        const bf16* vb = p.Hd + (long long)b_att * H + (col_ok ? col : 0);
#pragma unroll
        for (int l = 0; l < 32; ++l) v[l] = *reinterpret_cast<const uint32_t*>(vb + (long long)min(l, Ln - 1) * B * H);
        tmem_st32(t_park, v);
        uint32_t u[16];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          float4 x = reinterpret_cast<const float4*>(p.Uv + ((long long)min(aw + 8 * f, Ln - 1) * B + b_att) * A)[lane];
          x = attn::l4_add(x, ab4);                           // attn_b folded in once
          u[4 * f] = __float_as_uint(x.x); u[4 * f + 1] = __float_as_uint(x.y); u[4 * f + 2] = __float_as_uint(x.z); u[4 * f + 3] = __float_as_uint(x.w);
        }
        tmem_st16(t_park + 32u, u);
        tmem_st_wait();
      }
      float dr0 = 1.f, dr1 = 1.f;                          // dropout scales of this lane's two x_t columns for the coming step
      auto draw_dropout = [&](int t) {
        if (p.p_drop > 0.f && col_ok) {
          const uint64_t idx = (uint64_t)((long long)t * B + b_att) * (uint64_t)H + (uint64_t)col;
          const float4 d4 = dropout_scale4(p.rng, p.site, idx & ~3ull, p.p_drop);
          if ((idx & 3ull) == 0) { dr0 = d4.x; dr1 = d4.y; } else { dr0 = d4.z; dr1 = d4.w; }
        }
      };
      draw_dropout(0);
      for (int t = 0; t < S; ++t) {
        float4 wh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t > 0) {
          if (aw == 0 && lane == 0) { wait_flag(barA, ncta * (unsigned)t, p.err, abort_); if (cta == 0) stamp(12); }
          nbar(5, 256);
          // query = sum of the unit-group partials the cell phase of step t-1 left: warp aw takes groups aw, aw + 8, ..
#pragma unroll
          for (int i0 = 0; i0 < 8; i0 += 4) {
            float4 q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = reinterpret_cast<const float4*>(p.WhP + ((long long)min(aw + 8 * (i0 + i), p.UG - 1) * B + b_att) * A)[lane];
#pragma unroll
            for (int i = 0; i < 4; ++i) if (aw + 8 * (i0 + i) < p.UG) wh = attn::l4_add(wh, q[i]);
          }
        }
        reinterpret_cast<float4*>(whred + aw * 128)[lane] = wh;
        nbar(5, 256);
        wh = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) wh = attn::l4_add(wh, reinterpret_cast<const float4*>(whred + i * 128)[lane]);
        if (aw == 0) reinterpret_cast<float4*>(p.Wh + ((long long)t * B + b_att) * A)[lane] = wh;
        if (cta == 0 && aw == 0 && lane == 0) stamp(13);
        float sc[4];
        uint32_t v[32];
        {
          uint32_t u[16];
          tmem_ld16_nw(t_park + 32u, u);
          tmem_ld32_nw(t_park, v);
          tmem_ld_wait();
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            float s = aw4.x * act_tanh<true>(wh.x + __uint_as_float(u[4 * f]));
            s = fmaf(aw4.y, act_tanh<true>(wh.y + __uint_as_float(u[4 * f + 1])), s);
            s = fmaf(aw4.z, act_tanh<true>(wh.z + __uint_as_float(u[4 * f + 2])), s);
            sc[f] = fmaf(aw4.w, act_tanh<true>(wh.w + __uint_as_float(u[4 * f + 3])), s);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int f = 0; f < 4; ++f) sc[f] += __shfl_xor_sync(0xffffffffu, sc[f], o);
        if (lane < 4) {
          const float s = lane == 0 ? sc[0] : lane == 1 ? sc[1] : lane == 2 ? sc[2] : sc[3];
          const int l = aw + 8 * lane;
          if (l < Ln) { e_s[l] = s; p.beta[((long long)t * B + b_att) * Ln + l] = s; }
        }
        nbar(5, 256);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int l = 0; l < 32; l += 2) {
          if (l < Ln) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[l]));
            const float e = e_s[l];
            a0 = fmaf(e, f.x, a0); a1 = fmaf(e, f.y, a1);
          }
          if (l + 1 < Ln) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[l + 1]));
            const float e = e_s[l + 1];
            a2 = fmaf(e, f.x, a2); a3 = fmaf(e, f.y, a3);
          }
        }
        if (col_ok)
          *reinterpret_cast<__nv_bfloat162*>(p.X + ((long long)t * B + b_att) * KX + col) =
              __floats2bfloat162_rn((a0 + a2) * p.inv_L * dr0, (a1 + a3) * p.inv_L * dr1);
        if (cta == 0 && aw == 0 && lane == 0) stamp(14);
        nbar(5, 256);                                       // all of x_t[b] written (also protects e_s / whred for the next step)
        if (aw == 0 && lane == 0) { red_rel(barB); if (cta == 0) stamp(3); }
        if (t + 1 < S) draw_dropout(t + 1);                 // Philox for the next step, off the critical path
        cell_query_phase<KS>(p, cq, t, creg, hs);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---- host side ---------------------------------------------------------------------------------------------------
static inline bool persist_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("RECNET_PERSIST"); v = e ? atoi(e) : 1; }
  return v != 0;
}
static inline int sm_count() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = rt::NUM_SMS; }
  return n;
}
// K-split that keeps every CTA's weight slice resident (<= MAX_KB k-blocks) with UG*KS CTAs co-resident; 0 = shape not covered
static inline int pick_ks(int R, int H) {
  if (R % 64 || H % 64) return 0;
  const int UG = R / UNITS, NHB = R / BK, NXB = H / BK;
  for (int ks = 1; ks <= 4; ++ks) {
    int worst = 0;
    for (int k = 0; k < ks; ++k) {
      const int n = ((NHB * (k + 1)) / ks - (NHB * k) / ks) + ((NXB * (k + 1)) / ks - (NXB * k) / ks);
      worst = n > worst ? n : worst;
    }
    if (worst <= MAX_KB && UG * ks <= sm_count()) return ks;
  }
  return 0;
}
struct Shape { int B, S, R, H, A, L; };
static inline bool local_fwd_ok(const Shape& s) {
  if (!persist_enabled()) return false;
  const int ks = pick_ks(s.R, s.H);
  if (!ks) return false;
  const int UG = s.R / UNITS;
  return s.A == 128 && s.B >= 1 && s.B <= 128 && s.B <= UG * ks && s.L >= 1 && s.L <= 32 && s.H <= 512 && UG <= MAX_UG && s.S >= 1 &&
         (s.B + ks - 1) / ks <= 4 * NWORK && smem_layout(s.B, ks).total <= 227 * 1024;
}
static inline size_t xp_floats(const Shape& s) { const int ks = pick_ks(s.R, s.H); return ks ? (size_t)(s.R / UNITS) * ks * s.B * NCOL : 4; }
static inline size_t whp_floats(const Shape& s) { return (size_t)(s.R / UNITS + 1) * s.B * s.A; }
constexpr int SYNC_WORDS = 64 + 32 * MAX_UG;

template <int KS>
static int launch_ks(const CUtensorMap& mx, const CUtensorMap& mw, const FwdParams& p, cudaStream_t st) {
  auto kern = local_fwd_kernel<KS>;
  const int smem = smem_layout(p.B, KS).total;
  static int attr_smem = 0;
  if (attr_smem < smem) {
    RN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  void* args[] = {(void*)&mx, (void*)&mw, (void*)&p};
  ProfScope prof(KC_LOOP, p.S, p.UG * KS, 0, st);
  RN_CUDA_OK(cudaLaunchCooperativeKernel((const void*)kern, dim3(p.UG * KS), dim3(THREADS), args, (size_t)smem, st));
  RN_LAUNCH_OK();
  return 0;
}

// X / Wrec as the kernel-per-phase path stages them; sync must be zeroed (SYNC_WORDS words) before the launch
static int launch_local_fwd(FwdParams p, const bf16* Wrec, cudaStream_t st) {
  p.KS = pick_ks(p.R, p.H);
  p.UG = p.R / UNITS; p.NHB = p.R / BK; p.NXB = p.H / BK; p.KX = p.H + p.R;
  CUtensorMap mx, mw;
  RN_TRY(make_map(&mx, p.X, (long long)(p.S + 1) * p.B, p.KX, p.KX, BK, p.B));
  RN_TRY(make_map(&mw, Wrec, 4LL * p.R, p.KX, p.KX, BK, UNITS));
  switch (p.KS) {
    case 1: return launch_ks<1>(mx, mw, p, st);
    case 2: return launch_ks<2>(mx, mw, p, st);
    case 3: return launch_ks<3>(mx, mw, p, st);
    case 4: return launch_ks<4>(mx, mw, p, st);
  }
  return RECNET_ERR_UNSUPPORTED;
}
}  // namespace rp
