// EXPERIMENTAL, OPT-IN (RECNET_DEC_CLUSTER=1), written at the end of round 1 and NOT YET RUN ON A GPU (it compiles for sm_100a;
// the round's GPU budget was spent).  The default decoder path never reaches this file.  DESIGN.md section 9 item 3 is the plan.
//
// Decoder forward loop (projected-feature form, seq_decoder_pf.cuh) as ONE persistent kernel of independent 16-CTA clusters with
// the per-step weight [W_a ; W_hh] RESIDENT IN SHARED MEMORY for the whole sequence and no grid-wide synchronisation:
//   * samples are independent through the loop (train.py:41-66): cluster q owns samples [q MS, (q+1) MS), MS = ceil(B / #clusters) <= 16
//   * CTA r of a cluster owns hidden units [32 r, 32 r + 32) -> the 4 x 32 gate rows of W_hh of those units + attention rows
//     [A r / 16, A (r+1) / 16) of W_a: 136 rows x 512 bf16 = 139 KB, loaded once
//   * per step:  P1  out[16 x 136] = h_{t-1}[16 x H] . Wslice^T           mma.sync m16n8k16 (bf16, fp32 accumulate), no split-K
//                --- barrier.cluster ---
//                P2  CTA s (< MS) gathers Wh[s, 0:A] from the 16 CTAs (DSMEM), computes the Tn scores of sample s
//                --- barrier.cluster ---
//                P3  every CTA: context sum over Tn frames of VW for its 32 units x MS samples (scores read through DSMEM),
//                    gates, c_t, h_t; stashes exactly what pf_fwd_kernel stashes; h_t (operand type) is written into EVERY
//                    CTA's copy of the operand rows (DSMEM stores)
//                --- barrier.cluster ---
// Same arithmetic as pf_fwd_kernel + the per-step tcgen05 GEMM (same bf16 operands, fp32 accumulation; only the summation order
// of the K loop differs).  Requirements checked by the launcher: bf16 build, H % 16 == 0 and H / 16 a multiple of 8, A % 16 == 0,
// A / 16 <= 8... (see cluster_ok).  Anything else falls back to the per-step path.
#pragma once
#include <cooperative_groups.h>

#include "proj_attn.cuh"

namespace dcl {
namespace cg = cooperative_groups;

constexpr int CS = 16;            // CTAs per cluster (non-portable size: one cluster per GPC)
constexpr int THREADS = 256;
constexpr int MT = 16;            // samples per cluster (one m16 tile)
constexpr int KPAD = 8;           // bf16 elements of padding per smem row: row pitch = (H + 8) * 2 bytes -> conflict-free fragment loads

struct Args {
  const bf16* Wcat;               // [A + 4H, H]  rows [0, A) = W_a, rows A + g H + j = W_hh gate g of unit j
  const float* Uv;                // [B, Tn, A]   U v + attn_b
  const float* attn_w;            // [A]
  const bf16* VW;                 // [B, Tn, H, 4] unit-interleaved projected features
  const float* Gx;                // [L B, 4H]    embedding projection + b_ih (gate-block order)
  const float* b_hh;              // [4H]
  float* c;                       // [(L+1) B, H] fp32 cell states, row block 0 = zeros
  float* hiddens;                 // [L B, H] fp32
  bf16* Hop;                      // [(L+1) B, H] operand rows, row block 0 = zeros
  float* Wh; float* e;            // [L B, A], [L B, Tn] stash
  bf16* gates;                    // [L B, H, 4] stash
  int B, L, Tn, A, H, MS;         // MS = samples per cluster
  float inv_T;
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// dynamic shared memory layout (bytes): Ws [NR][H + KPAD] bf16 | hs [MT][H + KPAD] bf16 | og [MT][NR] f32 | es [32] f32 | whs [A] f32
__global__ void __launch_bounds__(THREADS, 1) decoder_fwd_cluster_kernel(Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int r = (int)cluster.block_rank();                   // CTA rank in the cluster = owner of units [UPC r, UPC r + UPC)
  const int q = blockIdx.x / CS;                             // cluster index = sample group
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = a.H, A = a.A, Tn = a.Tn, B = a.B;
  const int UPC = H / CS, APC = A / CS, NR = APC + 4 * UPC;  // units / attention rows / weight rows per CTA
  const int pitch = H + KPAD;
  bf16* Ws = reinterpret_cast<bf16*>(smem_raw);
  bf16* hs = Ws + (size_t)NR * pitch;
  float* og = reinterpret_cast<float*>(hs + (size_t)MT * pitch);
  float* es = og + MT * NR;                                  // scores of sample r (this CTA's P2 result), Tn <= 32 per pass
  float* whs = es + 64;                                      // gathered W h of sample r
  const int b0 = q * a.MS, nb = min(a.MS, B - b0);           // this cluster's samples
  if (nb <= 0) return;                                       // whole cluster exits together (uniform condition)

  // ---- one-time: this CTA's weight rows -> shared memory; h_{-1} = 0
  for (int i = tid; i < NR * (H / 8); i += THREADS) {
    const int lr = i / (H / 8), c8 = i - lr * (H / 8);
    const int grow = lr < APC ? r * APC + lr : A + ((lr - APC) / UPC) * H + r * UPC + (lr - APC) % UPC;
    *reinterpret_cast<uint4*>(Ws + (size_t)lr * pitch + 8 * c8) = *reinterpret_cast<const uint4*>(a.Wcat + (size_t)grow * H + 8 * c8);
  }
  for (int i = tid; i < MT * pitch / 2; i += THREADS) reinterpret_cast<uint32_t*>(hs)[i] = 0u;
  // fixed ownership in P3: warp w handles samples w, w + 8; lane = unit within the CTA's slice (UPC == 32 required)
  const int j = r * UPC + lane;                              // global hidden unit of this lane
  float cstate[2] = {0.f, 0.f};
  __syncthreads();
  cluster.sync();

  for (int t = 0; t < a.L; ++t) {
    // ---- P1: og[m][n] = sum_k hs[m][k] Ws[n][k]   (skipped at t = 0: h_{-1} = 0)
    if (t > 0) {
      const int g = lane >> 2, tq = lane & 3;
      for (int nt = warp; nt * 8 < NR; nt += THREADS / 32) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const bf16* arow0 = hs + (size_t)g * pitch + 2 * tq;
        const bf16* arow1 = hs + (size_t)(g + 8) * pitch + 2 * tq;
        const bf16* brow = Ws + (size_t)(nt * 8 + g) * pitch + 2 * tq;
#pragma unroll 4
        for (int k0 = 0; k0 < H; k0 += 16) {
          uint32_t af[4], bfr[2];
          af[0] = *reinterpret_cast<const uint32_t*>(arow0 + k0);
          af[1] = *reinterpret_cast<const uint32_t*>(arow1 + k0);
          af[2] = *reinterpret_cast<const uint32_t*>(arow0 + k0 + 8);
          af[3] = *reinterpret_cast<const uint32_t*>(arow1 + k0 + 8);
          bfr[0] = *reinterpret_cast<const uint32_t*>(brow + k0);
          bfr[1] = *reinterpret_cast<const uint32_t*>(brow + k0 + 8);
          mma_bf16_16816(acc, af, bfr);
        }
        const int n = nt * 8 + 2 * tq;
        og[g * NR + n] = acc[0]; og[g * NR + n + 1] = acc[1];
        og[(g + 8) * NR + n] = acc[2]; og[(g + 8) * NR + n + 1] = acc[3];
      }
    } else {
      for (int i = tid; i < MT * NR; i += THREADS) og[i] = 0.f;
    }
    __syncthreads();
    cluster.sync();                                          // #1: every CTA's og (its columns of [Wh | gates_h]) is complete

    // ---- P2: CTA s < nb computes the Tn scores of sample b0 + s
    if (r < nb) {
      const int b = b0 + r;
      for (int i = tid; i < A; i += THREADS) {               // gather W h_{t-1} [b, 0:A]: column i lives in CTA i / APC
        const float* rog = cluster.map_shared_rank(og, i / APC);
        const float v = rog[r * NR + (i % APC)];
        whs[i] = v;
        a.Wh[((size_t)t * B + b) * A + i] = v;
      }
      __syncthreads();
      const float* uvb = a.Uv + (size_t)b * Tn * A;
      for (int tau = warp; tau < Tn; tau += THREADS / 32) {
        float s = 0.f;
        for (int i = lane; i < A; i += 32) s = fmaf(a.attn_w[i], act_tanh<true>(whs[i] + uvb[(size_t)tau * A + i]), s);
        s = warp_sum(s);
        if (lane == 0) { es[tau] = s; a.e[((size_t)t * B + b) * Tn + tau] = s; }
      }
    }
    __syncthreads();
    cluster.sync();                                          // #2: scores of every sample are in their owner CTA's es

    // ---- P3: context + cell for (samples warp, warp + 8) x (unit j).  Latency-bound like pf_fwd_kernel, so the loads of BOTH samples
    // are issued unconditionally on clamped indices, 16 frames per sample in flight (32 x 8-byte loads per lane), before any use.
    {
      const int s0 = warp, s1 = warp + 8;
      const bool ok0 = s0 < nb, ok1 = s1 < nb;                 // warp-uniform
      const int c0 = min(s0, nb - 1), c1 = min(s1, nb - 1);
      const float* res0 = cluster.map_shared_rank(es, c0);     // scores of sample s live in CTA s
      const float* res1 = cluster.map_shared_rank(es, c1);
      const bf16* vw0 = a.VW + ((size_t)(b0 + c0) * Tn * H + j) * 4;
      const bf16* vw1 = a.VW + ((size_t)(b0 + c1) * Tn * H + j) * 4;
      float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
      for (int tau0 = 0; tau0 < Tn; tau0 += 16) {
        pf::Quad<bf16> v0[16], v1[16];
        float e0[16], e1[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int tau = min(tau0 + k, Tn - 1);
          v0[k].load(vw0 + (size_t)tau * 4 * H);
          v1[k].load(vw1 + (size_t)tau * 4 * H);
          e0[k] = res0[tau];
          e1[k] = res1[tau];
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float w0 = tau0 + k < Tn ? e0[k] : 0.f, w1 = tau0 + k < Tn ? e1[k] : 0.f;
          float f0[4], f1[4];
          v0[k].get(f0);
          v1[k].get(f1);
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) { acc0[gg] = fmaf(w0, f0[gg], acc0[gg]); acc1[gg] = fmaf(w1, f1[gg], acc1[gg]); }
        }
      }
#pragma unroll
      for (int sidx = 0; sidx < 2; ++sidx) {
        const int s = sidx ? s1 : s0;
        if (sidx ? ok1 : ok0) {                                // warp-uniform
          const int b = b0 + s;
          const float* acc = sidx ? acc1 : acc0;
          const float* gx = a.Gx + ((size_t)t * B + b) * 4 * H + j;
          float pre[4];
#pragma unroll
          for (int gg = 0; gg < 4; ++gg)
            pre[gg] = gx[gg * H] + a.b_hh[gg * H + j] + og[s * NR + APC + gg * UPC + lane] + acc[gg] * a.inv_T;
          const float gi = act_sigmoid<true>(pre[0]), gf = act_sigmoid<true>(pre[1]), gt = act_tanh<true>(pre[2]), go = act_sigmoid<true>(pre[3]);
          const float cn = fmaf(gf, cstate[sidx], gi * gt);
          const float hn = go * act_tanh<true>(cn);
          cstate[sidx] = cn;
          const size_t o1 = ((size_t)t * B + b) * H + j;
          a.c[o1 + (size_t)B * H] = cn;                        // row block t + 1
          a.hiddens[o1] = hn;
          const bf16 hb = __float2bfloat16_rn(hn);
          a.Hop[o1 + (size_t)B * H] = hb;
          pf::Quad<bf16>::store(a.gates + o1 * 4, gi, gf, gt, go);
          // h_t (operand type) into every CTA's operand rows: pairs of units packed by lane pairs, one 4-byte DSMEM store each
          const uint32_t lo = (uint32_t)__bfloat16_as_ushort(hb);
          const uint32_t hi = __shfl_down_sync(0xffffffffu, lo, 1);
          if (!(lane & 1)) {
            const uint32_t packed = lo | (hi << 16);
            for (int dst = 0; dst < CS; ++dst) {
              bf16* rhs = cluster.map_shared_rank(hs, dst);
              *reinterpret_cast<uint32_t*>(rhs + (size_t)s * pitch + j) = packed;
            }
          }
        }
      }
    }
    __syncthreads();
    cluster.sync();                                          // #3: h_t is everywhere, og / es may be overwritten
  }
}

static inline bool cluster_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("RECNET_DEC_CLUSTER"); v = e ? atoi(e) : 0; }
  return v != 0;
}
static inline size_t smem_bytes(int H, int A) {
  const int UPC = H / CS, APC = A / CS, NR = APC + 4 * UPC, pitch = H + KPAD;
  return (size_t)NR * pitch * 2 + (size_t)MT * pitch * 2 + (size_t)MT * NR * 4 + 64 * 4 + (size_t)A * 4 + 16;
}
// shapes this kernel handles (everything else: per-step path)
static inline bool cluster_ok(int B, int Tn, int A, int H, int n_clusters) {
  if (n_clusters < 1) return false;
  const int MS = (B + n_clusters - 1) / n_clusters;
  return H % CS == 0 && H / CS == 32 && H % 16 == 0 && A % CS == 0 && A >= CS && Tn >= 1 && Tn <= 64 && MS <= MT && MS <= CS &&
         smem_bytes(H, A) <= 227 * 1024 && ((A / CS + 4 * (H / CS)) % 8) == 0;
}

// Runs the whole forward loop; returns RECNET_ERR_UNSUPPORTED (before launching anything) if the device cannot host the cluster.
static int launch(const Args& a0, int n_clusters_wanted, cudaStream_t st) {
  Args a = a0;
  auto kern = decoder_fwd_cluster_kernel;
  const size_t smem = smem_bytes(a.H, a.A);
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return RECNET_ERR_UNSUPPORTED; }
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return RECNET_ERR_UNSUPPORTED; }
    configured = true;
  }
  int ncl = n_clusters_wanted;
  a.MS = (a.B + ncl - 1) / ncl;
  ncl = (a.B + a.MS - 1) / a.MS;                               // drop empty clusters
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ncl * CS, 1, 1);
  cfg.blockDim = dim3(THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  RN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, a));
  RN_LAUNCH_OK();
  return 0;
}
}  // namespace dcl
