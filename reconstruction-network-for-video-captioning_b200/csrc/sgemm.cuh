// fp32 CUDA-core GEMM (FFMA, fp32 accumulate): the "fp32 parity" provider.
//   C[m,n] (+)= sum_k A(m,k) * B(n,k) (+ bias[n])
// A(m,k) = A[m*a_rs + k*a_cs], B(n,k) = B[n*b_rs + k*b_cs]  -> any of NN/NT/TN/TT without copies.
// blockIdx.z = split-K slice; slice z writes C + z*split_stride (bias only in slice 0).
// This is the precision=fp32 path (1e-3 parity against the fp32/fp64 oracle and bit-exact greedy ids);
// the throughput path is the tcgen05 kernel in gemm_tc.cuh.
#pragma once
#include "common.cuh"

namespace sg {
constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;

__global__ void __launch_bounds__(THREADS) sgemm_kernel(
    const float* __restrict__ A, long long a_rs, long long a_cs,
    const float* __restrict__ B, long long b_rs, long long b_cs,
    float* __restrict__ C, long long ldc, const float* __restrict__ bias,
    int M, int N, int K, int k_per_split, long long split_stride, int accumulate) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb = blockIdx.z * k_per_split;
  const int ke = min(K, kb + k_per_split);
  const int tx = tid & 15, ty = tid >> 4;   // 16x16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool a_kfast = (a_cs == 1);
  const bool b_kfast = (b_cs == 1);
  for (int k0 = kb; k0 < ke; k0 += BK) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / THREADS; ++i) {
      const int idx = tid + i * THREADS;
      int m, k;
      if (a_kfast) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < ke) ? __ldg(A + (long long)gm * a_rs + (long long)gk * a_cs) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / THREADS; ++i) {
      const int idx = tid + i * THREADS;
      int n, k;
      if (b_kfast) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
      const int gn = n0 + n, gk = k0 + k;
      Bs[k][n] = (gn < N && gk < ke) ? __ldg(B + (long long)gn * b_rs + (long long)gk * b_cs) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* Cz = C + (long long)blockIdx.z * split_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr && blockIdx.z == 0) v += bias[gn];
      float* p = Cz + (long long)gm * ldc + gn;
      if (accumulate) v += *p;
      *p = v;
    }
  }
}

static inline int launch(const float* A, long long lda, int transA, const float* B, long long ldb, int transB,
                         float* C, long long ldc, const float* bias, int M, int N, int K, int splits,
                         long long split_stride, int accumulate, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  if (splits < 1) splits = 1;
  int k_per = rn_cdiv(rn_cdiv(K, splits), BK) * BK;
  if (k_per < BK) k_per = BK;
  dim3 grid(rn_cdiv(N, BN), rn_cdiv(M, BM), splits);
  const long long a_rs = transA ? 1 : lda, a_cs = transA ? lda : 1;
  const long long b_rs = transB ? 1 : ldb, b_cs = transB ? ldb : 1;
  ProfScope prof(KC_SGEMM, M, N, K, st);
  sgemm_kernel<<<grid, THREADS, 0, st>>>(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, bias, M, N, K, k_per, split_stride,
                                         accumulate);
  RN_LAUNCH_OK();
  return 0;
}
}  // namespace sg
