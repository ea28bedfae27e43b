// Phase emitter of the kernel-per-phase time loops (GRU / stacked / unaligned variants, the fp32 build, and every shape the
// weight-resident persistent loops of seq_recon_persist.cuh do not cover): each phase -- a tensor-core GEMM leaving split-K
// partials, the fused attention, the fused cell update, or their backward counterparts -- is launched as its own kernel on the
// emitter's stream.  (Round 1 also recorded the phases into a table walked by one cooperative "loop kernel" per time loop; with one
// CTA per SM re-pulling the weights from L2 in every phase it measured 39 % slower than the graph of kernels,
// profiles/r1_c_loop_kernel.md, and was removed in round 2 in favour of the weight-resident design.)
#pragma once
#include "runtime.cuh"

namespace mega {

template <typename T>
struct Emitter {
  cudaStream_t st;
  Emitter(bool, cudaStream_t s, size_t = 0) : st(s) {}
  int gemm_partials(const T* A, long long lda, int tA, const T* B, long long ldb, int tB, float* P, int M, int N, int K, rt::GemmPlan p) {
    return rt::gemm_partials<T>(A, lda, tA, B, ldb, tB, P, M, N, K, p, st);
  }
  int attn_fwd(attn::FwdArgs a) { return attn::launch_fwd<T, T>(a, st); }
  int attn_bwd(const attn::BwdArgs& a) { return attn::launch_bwd<T, T>(a, st); }
  int cell_fwd(const cell::FwdArgs& a) { return cell::launch_fwd<T, T>(a, st); }
  int cell_bwd(const cell::BwdArgs& a) { return cell::launch_bwd<T, T>(a, st); }
  void no_sync_after_last() {}
  int flush(void*, size_t, unsigned*, int*, int) { return 0; }
};

constexpr size_t table_bytes(int) { return 256; }
}  // namespace mega
