// Gradient all-reduce (average) over NVLink / NVSwitch written as ONE kernel of our own -- the data-parallel exchange step of the
// path (the reference has no distributed code).  No NCCL collective sits in the captured step graph: the flat gradient
// buffers live in symmetric memory (every rank maps every peer's copy, plus the NVSwitch multicast alias of all of them), and a few
// CTAs do a two-shot all-reduce in place:
//
//   barrier (every rank's gradients are complete)                                   per-CTA flags in symmetric memory, monotonic epochs
//   rank r owns slice r:  v = multimem.ld_reduce.add(slice)  -> the switch sums the W copies (NVLS in-switch reduction)
//                         multimem.st(slice, v / W)          -> the switch writes all W copies
//   barrier (every slice has landed everywhere)
//
// Per GPU and direction ~n bytes cross NVLink (vs 2 (W-1)/W n for a ring), and the kernel needs no proxy thread, no work FIFO and no
// second stream.  Without multicast support the same kernel sums the W peer pointers with volatile 16-byte loads and stores the result
// to every peer (two-shot over peer memory).  Epochs only grow, so CUDA-graph replays need no reset.  All spins are bounded.
#pragma once
#include "common.cuh"

namespace ar {
constexpr int THREADS = 512;
constexpr int MAX_CTAS = 64;
constexpr int MAX_WORLD = 16;
constexpr int UNROLL = 8;
constexpr long long TIMEOUT = 4000000000LL;      // ~2 s of SM clocks: a rank that never arrives is a launch-order bug, not a hang

struct Args {
  float* local;                 // this rank's symmetric buffer
  float* mc;                    // multicast alias (nullptr: peer-pointer path)
  const long long* peers;       // [world] device table: every rank's buffer as mapped into this process
  unsigned* flags;              // this rank's flag block [2][MAX_CTAS][MAX_WORLD] (symmetric, zeroed once)
  const long long* peer_flags;  // [world] device table: every rank's flag block
  unsigned* epochs;             // [MAX_CTAS] local per-CTA epoch counters (zeroed once)
  int* err;                     // local error flag (0 ok, 4 = a peer did not arrive in time)
  long long off4, n4;           // range of the buffer to reduce, in float4 units
  int rank, world;
  float scale;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 mm_ld_reduce(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_volatile4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// all ranks' CTA `c` meet: thread i < world tells rank i "rank `rank` reached epoch e of phase ph", then waits for rank i's word
__device__ __forceinline__ void rank_barrier(const Args& a, int ph, unsigned e) {
  __syncthreads();                                         // everything this CTA wrote before is ordered before the releases below
  const int i = threadIdx.x, c = blockIdx.x;
  if (i < a.world) {
    unsigned* theirs = reinterpret_cast<unsigned*>(a.peer_flags[i]) + ((size_t)ph * MAX_CTAS + c) * MAX_WORLD + a.rank;
    st_release_sys(theirs, e);
    const unsigned* mine = a.flags + ((size_t)ph * MAX_CTAS + c) * MAX_WORLD + i;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(mine) - e) < 0) {
      if (clock64() - t0 > TIMEOUT) { atomicExch(a.err, 4); break; }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(THREADS) allreduce_avg_kernel(const Args a) {
  const unsigned e = a.epochs[blockIdx.x] + 1u;
  rank_barrier(a, 0, e);
  const long long per = (a.n4 + a.world - 1) / a.world;
  const long long lo = a.off4 + (long long)a.rank * per, hi = min(a.off4 + a.n4, lo + per);
  const long long stride = (long long)gridDim.x * THREADS;
  if (a.mc) {
    for (long long i0 = lo + (long long)blockIdx.x * THREADS + threadIdx.x; i0 < hi; i0 += stride * UNROLL) {
      float4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) { const long long i = i0 + u * stride; if (i < hi) v[u] = mm_ld_reduce(a.mc + 4 * i); }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const long long i = i0 + u * stride;
        if (i < hi) mm_st(a.mc + 4 * i, make_float4(v[u].x * a.scale, v[u].y * a.scale, v[u].z * a.scale, v[u].w * a.scale));
      }
    }
  } else {
    for (long long i = lo + (long long)blockIdx.x * THREADS + threadIdx.x; i < hi; i += stride) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p = 0; p < a.world; ++p) {                  // same order on every rank: bitwise identical results everywhere
        const float4 v = ld_volatile4(reinterpret_cast<const float*>(a.peers[p]) + 4 * i);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      s.x *= a.scale; s.y *= a.scale; s.z *= a.scale; s.w *= a.scale;
      for (int p = 0; p < a.world; ++p) *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.peers[p]) + 4 * i) = s;
    }
  }
  rank_barrier(a, 1, e);
  if (threadIdx.x == 0) a.epochs[blockIdx.x] = e;
}

static int launch(const Args& a, int ctas, cudaStream_t st) {
  if (a.n4 <= 0) return 0;
  if (a.world < 1 || a.world > MAX_WORLD || a.rank < 0 || a.rank >= a.world || !a.local || !a.flags || !a.peer_flags || !a.epochs || !a.err ||
      (!a.mc && !a.peers))
    return RECNET_ERR_BAD_SHAPE;
  if ((reinterpret_cast<uintptr_t>(a.local) & 15) || (a.mc && (reinterpret_cast<uintptr_t>(a.mc) & 15))) return RECNET_ERR_ALIGNMENT;
  if (ctas < 1) ctas = 1;
  if (ctas > MAX_CTAS) ctas = MAX_CTAS;
  allreduce_avg_kernel<<<ctas, THREADS, 0, st>>>(a);
  RN_LAUNCH_OK();
  return 0;
}
}  // namespace ar
