// Persistent cooperative LOOP KERNEL: one launch runs a whole time loop (decoder forward, decoder BPTT,
// reconstructor forward, reconstructor BPTT).
//
// Why: with one kernel per phase the step is bound by per-kernel launch/dependency latency (r1 profile: 538 kernels,
// 4.47 ms, every kernel 5-10 us for ~1 us of work).  Here one CTA per SM stays resident for the whole loop; the host
// uploads a PHASE TABLE (one entry per former kernel launch: a tensor-core GEMM tile set, the fused attention, the
// fused cell update, or their backward counterparts) and the CTAs walk it, separated by grid barriers only where
// a phase consumes what the previous one produced.  The GEMM phase is the same TMA -> smem ring -> tcgen05.mma ->
// TMEM -> tcgen05.ld pipeline as gemm_tc.cuh, with the ring, the mbarriers and the TMEM allocation kept alive across
// phases (runtime tile width / operand majors).  Warp roles inside a GEMM phase: 0-3 epilogue (TMEM lane quarters),
// 4 TMA producer, 5 MMA issuer.
//
// Safety: every spin (mbarrier wait, grid barrier) has a clock64() timeout that raises a device error flag and makes
// all CTAs drain, so a protocol bug surfaces as RECNET_ERR_* on the next status check instead of a hung GPU.
#pragma once
#include <map>
#include <type_traits>
#include <vector>

#include "runtime.cuh"

namespace mega {

constexpr int THREADS = 256;         // NSUB 256-thread sub-blocks (1024 threads = 4 sub-blocks spills at 64 regs and measured slower)
constexpr int SUB = 256;
constexpr int NSUB = THREADS / SUB;
constexpr int STAGES = 5;
constexpr int A_BYTES = tc::BM * tc::BK * 2;          // 16 KB
constexpr int B_BYTES_MAX = 128 * tc::BK * 2;         // BN <= 128
constexpr int STAGE_BYTES = A_BYTES + B_BYTES_MAX;    // 32 KB
constexpr int SCRATCH_BYTES = 20 * 1024;              // attention scratch per sub-block (D + Tn + 512 floats)
constexpr long long SPIN_TIMEOUT = 400000000LL;       // ~0.2 s of SM clocks

enum PhaseType { PH_GEMM = 1, PH_ATTN_FWD = 2, PH_CELL_FWD = 3, PH_CELL_BWD = 4, PH_ATTN_BWD = 5 };

struct GemmArgs {
  CUtensorMap tmA, tmB;
  tc::EpiArgs ep;
  int M, N, K, kb_per, n_tiles, m_tiles, splits, BN, TA, TB;
};

struct alignas(128) Phase {
  union U {
    GemmArgs g; attn::FwdArgs af; attn::BwdArgs ab; cell::FwdArgs cf; cell::BwdArgs cb;
    __host__ __device__ U() {}
  } u;
  int type, nvb, nvb_x, sync_after;
};

struct SmemLayout {
  static constexpr int RING = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = RING;                               // full[STAGES], empty[STAGES], tmem, slot
  static constexpr int PHASE_OFF = BAR_OFF + 128;
  static constexpr int PHASE_BYTES = ((int)sizeof(Phase) + 127) / 128 * 128;
  static constexpr int SCRATCH_OFF = PHASE_OFF + 2 * PHASE_BYTES;      // double-buffered phase descriptor
  static constexpr int TOTAL = SCRATCH_OFF + NSUB * SCRATCH_BYTES + 1024;   // + alignment slack
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool mbar_wait_to(uint32_t bar, uint32_t parity, int* err) {
  const long long t0 = clock64();
  uint32_t ok;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
    if (clock64() - t0 > SPIN_TIMEOUT) { atomicExch(err, 2); return false; }
  }
}

// grid-wide barrier: all threads' prior writes are ordered by bar.sync before thread 0's gpu-scope release; thread 0
// arrives with red.release (no round trip) and polls with ld.acquire.  `target` = arrivals expected in total so far.
// On timeout (or when another CTA already failed) sets *s_abort so the whole CTA leaves the phase loop.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target, int* err, int* s_abort) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    const long long t0 = clock64();
    unsigned spins = 0;
    while (ld_acquire_u32(counter) < target) {
      if ((++spins & 0x3FFu) == 0) {
        if (clock64() - t0 > SPIN_TIMEOUT) { atomicExch(err, 3); *s_abort = 1; break; }
        if (*reinterpret_cast<volatile int*>(err) != 0) { *s_abort = 1; break; }
      }
    }
  }
  __syncthreads();
}

// One 128 x BN output tile (x K-slice) of a GEMM phase.  `it` counts k-blocks pushed through the ring so far,
// `tiles` counts accumulator hand-offs; both advance identically in every thread of the CTA.
__device__ __forceinline__ void gemm_tile(const GemmArgs& g, const Phase* gph, int vb, uint32_t base, uint32_t bar_full,
                                          uint32_t bar_empty, uint32_t bar_tmem, uint32_t tmem_base, uint32_t& it,
                                          uint32_t& tiles, int* err) {
  using namespace tc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = vb % g.n_tiles, rest = vb / g.n_tiles;
  const int mt = rest % g.m_tiles, z = rest / g.m_tiles;
  const int m0 = mt * BM, n0 = nt * g.BN;
  const int nkb_total = (g.K + BK - 1) / BK;
  const int kb0 = z * g.kb_per;
  const int nkb = max(0, min(nkb_total, kb0 + g.kb_per) - kb0);
  const uint32_t stage_tx = A_BYTES + (uint32_t)g.BN * BK * 2;

  if (warp == 4) {
    if (lane == 0) {
      asm volatile("fence.proxy.async;" ::: "memory");     // generic-proxy writes of earlier phases -> TMA reads
      for (int i = 0; i < nkb; ++i) {
        const uint32_t s = (it + i) % STAGES, ph = ((it + i) / STAGES) & 1u;
        if (!mbar_wait_to(bar_empty + 8 * s, ph ^ 1u, err)) break;
        mbar_expect_tx(bar_full + 8 * s, stage_tx);
        const uint32_t sa = base + s * STAGE_BYTES, sb = sa + A_BYTES;
        const int k = (kb0 + i) * BK;
        if (!g.TA) tma_load_2d(sa, &gph->u.g.tmA, bar_full + 8 * s, k, m0);
        else
          for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * (BK * 128), &gph->u.g.tmA, bar_full + 8 * s, m0 + j * 64, k);
        if (!g.TB) tma_load_2d(sb, &gph->u.g.tmB, bar_full + 8 * s, k, n0);
        else
          for (int j = 0; j < g.BN / 64; ++j) tma_load_2d(sb + j * (BK * 128), &gph->u.g.tmB, bar_full + 8 * s, n0 + j * 64, k);
      }
    }
  } else if (warp == 5) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((g.TA ? 1u : 0u) << 15) | ((g.TB ? 1u : 0u) << 16) |
                           ((uint32_t)(g.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    for (int i = 0; i < nkb; ++i) {
      const uint32_t s = (it + i) % STAGES, ph = ((it + i) / STAGES) & 1u;
      if (!mbar_wait_to(bar_full + 8 * s, ph, err)) break;
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = base + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
          const uint64_t ad = g.TA ? umma_smem_desc(sa + kk * (UMMA_K * 128), BK * 128, 1024)
                                   : umma_smem_desc(sa + kk * (UMMA_K * 2), 16, 1024);
          const uint64_t bd = g.TB ? umma_smem_desc(sb + kk * (UMMA_K * 128), BK * 128, 1024)
                                   : umma_smem_desc(sb + kk * (UMMA_K * 2), 16, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (i > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(bar_empty + 8 * s);
        if (i == nkb - 1) umma_commit(bar_tmem);
      }
      __syncwarp();
    }
  } else if (warp < 4) {
    const int q = warp;
    const int m = m0 + q * 32 + lane;
    bool ok = true;
    if (nkb > 0) {
      ok = mbar_wait_to(bar_tmem, tiles & 1u, err);
      tc_fence_after();
    }
    const bool z0 = (z == 0);
    const EpiArgs& ep = g.ep;
    float* crow = ep.Cf ? ep.Cf + (long long)z * ep.split_stride + (long long)m * ep.ldc : nullptr;
    for (int c0 = 0; c0 < g.BN && ok; c0 += 32) {
      uint32_t r[32];
      if (nkb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (m < g.M && crow) {
        const int n = n0 + c0;
        if (ep.vec_ok && n + 32 <= g.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            if (ep.bias && z0) {
              const float4 b4 = *reinterpret_cast<const float4*>(ep.bias + n + j);
              v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
            }
            float4* p = reinterpret_cast<float4*>(crow + n + j);
            if (ep.accumulate) { const float4 o = *p; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
            *p = v;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n + j < g.N) {
              float v = __uint_as_float(r[j]);
              if (ep.bias && z0) v += ep.bias[n + j];
              if (ep.accumulate) v += crow[n + j];
              crow[n + j] = v;
            }
        }
      }
    }
    tc_fence_before();
  }
  it += (uint32_t)nkb;
  if (nkb > 0) tiles += 1u;
}

__global__ void __launch_bounds__(THREADS, 1)
loop_kernel(const Phase* __restrict__ table, int n_phases, unsigned* bar_counter, int* err) {
  using L = SmemLayout;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - tc::smem_u32(smem_raw));       // generic pointer to the aligned base
  const uint32_t bar_full = base + L::BAR_OFF;
  const uint32_t bar_empty = bar_full + 8 * STAGES;
  const uint32_t bar_tmem = bar_empty + 8 * STAGES;
  const uint32_t tmem_slot = bar_tmem + 8;
  int* s_abort = reinterpret_cast<int*>(gen + L::BAR_OFF + 8 * (2 * STAGES + 1) + 8);
  const int warp = threadIdx.x >> 5;
  const int sub = threadIdx.x / SUB, tid = threadIdx.x % SUB;          // sub-block index / thread index inside it
  float* scratch = reinterpret_cast<float*>(gen + L::SCRATCH_OFF + sub * SCRATCH_BYTES);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(bar_full + 8 * s, 1); tc::mbar_init(bar_empty + 8 * s, 1); }
    tc::mbar_init(bar_tmem, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *s_abort = 0;
  }
  if (warp == 5) tc::tmem_alloc(tmem_slot, 128);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  uint32_t it = 0, tiles = 0;
  unsigned n_bar = 0;
  // descriptor p lives in buffer p & 1; descriptor p+1 is prefetched with cp.async while phase p runs
  auto prefetch = [&](int p) {
    if (p < n_phases) {
      const uint32_t dst = base + L::PHASE_OFF + (p & 1) * L::PHASE_BYTES;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(table + p);
      for (int i = threadIdx.x; i < (int)(sizeof(Phase) / 16); i += THREADS)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * i), "l"(src + 16 * i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(0);
  for (int p = 0; p < n_phases; ++p) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                       // descriptor p visible to all; everyone is done with buffer (p+1)&1
    if (*s_abort) break;
    prefetch(p + 1);
    const Phase* ph = reinterpret_cast<const Phase*>(gen + L::PHASE_OFF + (p & 1) * L::PHASE_BYTES);
    const Phase* gph = table + p;
    probe(ph->type, threadIdx.x);
    const int nvs = gridDim.x * NSUB, vs = blockIdx.x * NSUB + sub;      // virtual sub-blocks of the whole grid
    switch (ph->type) {
      case PH_GEMM:
        for (int vb = blockIdx.x; vb < ph->nvb; vb += gridDim.x) {
          if (sub == 0) gemm_tile(ph->u.g, gph, vb, base, bar_full, bar_empty, bar_tmem, tmem_base, it, tiles, err);
          __syncthreads();
        }
        break;
      case PH_ATTN_FWD:
        for (int vb = vs; vb < ph->nvb; vb += nvs) {
          attn::attn_fwd_body<bf16, bf16>(ph->u.af, vb % ph->nvb_x, vb / ph->nvb_x, scratch, tid, 1 + sub);
          blk_sync(1 + sub);
        }
        break;
      case PH_ATTN_BWD:
        for (int vb = vs; vb < ph->nvb; vb += nvs) {
          attn::attn_bwd_body<bf16, bf16>(ph->u.ab, vb, scratch, tid, 1 + sub);
          blk_sync(1 + sub);
        }
        break;
      case PH_CELL_FWD:
        for (int vb = vs; vb < ph->nvb; vb += nvs) cell::lstm_cell_fwd_body<bf16, bf16>(ph->u.cf, vb % ph->nvb_x, vb / ph->nvb_x, tid);
        break;
      case PH_CELL_BWD:
        for (int vb = vs; vb < ph->nvb; vb += nvs) cell::lstm_cell_bwd_body<bf16, bf16>(ph->u.cb, vb % ph->nvb_x, vb / ph->nvb_x, tid);
        break;
      default: break;
    }
    if (ph->sync_after) {
      probe(100, threadIdx.x);
      n_bar += 1;
      grid_barrier(bar_counter, n_bar * gridDim.x, err, s_abort);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  probe(0, threadIdx.x);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc(tmem_base, 128);
}

// ---- host side: phase emitter ------------------------------------------------------------------------------------
struct Pinned { Phase* host = nullptr; size_t cap = 0; };
inline std::map<std::pair<const void*, int>, Pinned>& pinned_cache() { static std::map<std::pair<const void*, int>, Pinned> m; return m; }

static inline bool mega_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("RECNET_MEGA"); v = e ? atoi(e) : 0; }   // measured slower than the graph of kernels: opt-in
  return v != 0;
}
static inline int sm_count() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = rt::NUM_SMS; }
  return n;
}

// Either launches each phase as its own kernel (eager: fp32 build, profiling, RECNET_MEGA=0) or records it into the
// phase table of one loop-kernel launch (bf16 build).
template <typename T>
struct Emitter {
  bool mega;
  cudaStream_t st;
  std::vector<Phase> phases;
  // scratch_floats: shared-memory floats the attention bodies of this loop need per sub-block (0 = no attention)
  Emitter(bool want_mega, cudaStream_t s, size_t scratch_floats = 0)
      : mega(want_mega && std::is_same<T, bf16>::value && mega_enabled() && scratch_floats * sizeof(float) <= (size_t)SCRATCH_BYTES), st(s) {}

  int gemm_partials(const T* A, long long lda, int tA, const T* B, long long ldb, int tB, float* P, int M, int N, int K,
                    rt::GemmPlan p) {
    if (!mega) return rt::gemm_partials<T>(A, lda, tA, B, ldb, tB, P, M, N, K, p, st);
    if constexpr (std::is_same<T, bf16>::value) {
      Phase ph;
      memset(&ph, 0, sizeof(ph));
      GemmArgs& g = ph.u.g;
      const int BN = (p.bn == 128) ? 128 : 64;
      const int nkb = rn_cdiv(K, tc::BK);
      int splits = p.splits < 1 ? 1 : (p.splits > nkb ? nkb : p.splits);
      const int kb_per = rn_cdiv(nkb, splits);
      splits = rn_cdiv(nkb, kb_per);
      if (!tA) { RN_TRY(tc::make_map(&g.tmA, A, M, K, lda, tc::BK, tc::BM)); } else { RN_TRY(tc::make_map(&g.tmA, A, K, M, lda, 64, tc::BK)); }
      if (!tB) { RN_TRY(tc::make_map(&g.tmB, B, N, K, ldb, tc::BK, BN)); } else { RN_TRY(tc::make_map(&g.tmB, B, K, N, ldb, 64, tc::BK)); }
      g.ep.Cf = P; g.ep.ldc = N; g.ep.split_stride = (long long)M * N; g.ep.Cb = nullptr; g.ep.ldcb = 0; g.ep.bias = nullptr;
      g.ep.accumulate = 0;
      g.ep.vec_ok = ((reinterpret_cast<uintptr_t>(P) & 15) == 0 && (N & 3) == 0 && (((long long)M * N) & 3) == 0) ? 1 : 0;
      g.M = M; g.N = N; g.K = K; g.kb_per = kb_per; g.n_tiles = rn_cdiv(N, BN); g.m_tiles = rn_cdiv(M, tc::BM);
      g.splits = splits; g.BN = BN; g.TA = tA; g.TB = tB;
      ph.type = PH_GEMM; ph.nvb = g.n_tiles * g.m_tiles * splits; ph.nvb_x = g.n_tiles; ph.sync_after = 1;
      phases.push_back(ph);
    }
    return 0;
  }
  int attn_fwd(attn::FwdArgs a) {
    if (!mega) return attn::launch_fwd<T, T>(a, st);
    int slices = 1;
    RN_TRY(attn::prepare_fwd<T>(a, &slices));
    if (attn::fwd_smem_bytes(a) > SCRATCH_BYTES) return RECNET_ERR_BAD_SHAPE;
    Phase ph; memset(&ph, 0, sizeof(ph));
    ph.u.af = a; ph.type = PH_ATTN_FWD; ph.nvb = slices * a.B; ph.nvb_x = slices; ph.sync_after = 1;
    phases.push_back(ph);
    return 0;
  }
  int attn_bwd(const attn::BwdArgs& a) {
    if (!mega) return attn::launch_bwd<T, T>(a, st);
    if (a.Tn > attn::MAX_T || a.Tn < 1 || attn::bwd_smem_bytes(a) > SCRATCH_BYTES) return RECNET_ERR_BAD_SHAPE;
    Phase ph; memset(&ph, 0, sizeof(ph));
    ph.u.ab = a; ph.type = PH_ATTN_BWD; ph.nvb = a.B; ph.nvb_x = 1; ph.sync_after = 1;
    phases.push_back(ph);
    return 0;
  }
  int cell_fwd(const cell::FwdArgs& a) {
    if (!mega) return cell::launch_fwd<T, T>(a, st);
    Phase ph; memset(&ph, 0, sizeof(ph));
    ph.u.cf = a; ph.type = PH_CELL_FWD; ph.nvb_x = rn_cdiv(a.H, cell::THREADS); ph.nvb = ph.nvb_x * a.B; ph.sync_after = 1;
    phases.push_back(ph);
    return 0;
  }
  int cell_bwd(const cell::BwdArgs& a) {
    if (!mega) return cell::launch_bwd<T, T>(a, st);
    Phase ph; memset(&ph, 0, sizeof(ph));
    ph.u.cb = a; ph.type = PH_CELL_BWD; ph.nvb_x = rn_cdiv(a.H, cell::THREADS); ph.nvb = ph.nvb_x * a.B; ph.sync_after = 1;
    phases.push_back(ph);
    return 0;
  }
  // the previous phase's output is not consumed by the next one: drop the barrier between them
  void no_sync_after_last() { if (mega && !phases.empty()) phases.back().sync_after = 0; }

  // upload the table (pinned staging, cached per destination) and launch the loop kernel once
  int flush(void* table_dev, size_t table_cap_bytes, unsigned* bar_dev, int* err_dev, int kind) {
    if (!mega || phases.empty()) return 0;
    const size_t bytes = phases.size() * sizeof(Phase);
    if (bytes > table_cap_bytes) return RECNET_ERR_WORKSPACE;
    // Pinned staging buffer, one per (destination, loop kind): a captured memcpy node re-reads it at every graph
    // replay, so it must stay untouched for the lifetime of that workspace address.  cudaHostAlloc is not allowed
    // under the default capture mode -> switch this thread to relaxed mode around it (what PyTorch's allocator does).
    Pinned& pin = pinned_cache()[std::make_pair((const void*)table_dev, kind)];
    if (pin.cap < bytes) {
      cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
      RN_CUDA_OK(cudaThreadExchangeStreamCaptureMode(&mode));
      cudaError_t e = cudaSuccess;
      if (pin.host) e = cudaFreeHost(pin.host);
      if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&pin.host), bytes, cudaHostAllocDefault);
      cudaThreadExchangeStreamCaptureMode(&mode);
      if (e != cudaSuccess) return (int)e;
      pin.cap = bytes;
    }
    memcpy(pin.host, phases.data(), bytes);
    RN_CUDA_OK(cudaMemcpyAsync(table_dev, pin.host, bytes, cudaMemcpyHostToDevice, st));
    RN_CUDA_OK(cudaMemsetAsync(bar_dev, 0, sizeof(unsigned), st));
    static bool attr = false;
    if (!attr) {
      RN_CUDA_OK(cudaFuncSetAttribute(loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout::TOTAL));
      attr = true;
    }
    const Phase* tab = reinterpret_cast<const Phase*>(table_dev);
    int n = (int)phases.size();
    void* args[] = {(void*)&tab, (void*)&n, (void*)&bar_dev, (void*)&err_dev};
    ProfScope prof(KC_LOOP, n, kind, 0, st);
    RN_CUDA_OK(cudaLaunchCooperativeKernel((const void*)loop_kernel, dim3(sm_count()), dim3(THREADS), args, SmemLayout::TOTAL, st));
    prof_state().launches++;
    return 0;
  }
};

constexpr size_t table_bytes(int steps) { return (size_t)(steps * 4 + 8) * sizeof(Phase); }
}  // namespace mega
