// Fused tail of the training iteration (reference train.py:269-273): gradient-norm clip of a parameter list
// (torch.nn.utils.clip_grad_norm_, train.py:269-270) + one Adam step with L2 weight decay and optional amsgrad
// (torch.optim.Adam as configured at train.py:149-150,186-187), over the same (pointer, size, block map) tables the
// L2-norm regulariser uses.  Three launches per module: squared-norm partials of the gradients (only when clipping),
// a one-block prologue (fixed-order reduction of the partials -> clip coefficient; step += 1; bias corrections) and
// one streaming update kernel that reads p, g, m, v (, vmax) once and writes p, m, v (, vmax) once.
// The step counter lives on the device so that a captured CUDA graph advances it on every replay.
#pragma once
#include "common.cuh"
#include "misc.cuh"

namespace optim {

// state (device float[8]): [0] step count, [1] last total gradient norm (0 when not clipping), [2] clip coefficient,
//                          [3] lr / (1 - beta1^step), [4] 1 / sqrt(1 - beta2^step), [5..7] reserved
enum { ST_STEP = 0, ST_GNORM = 1, ST_CLIP = 2, ST_STEPSIZE = 3, ST_RSQRT_BC2 = 4, ST_WORDS = 8 };

__global__ void adam_prologue_kernel(const float* __restrict__ partial, int n_partial, float max_norm, double lr, double beta1,
                                     double beta2, float* __restrict__ state) {
  __shared__ float red[32];
  float s = 0.f;
  if (partial)
    for (int i = threadIdx.x; i < n_partial; i += blockDim.x) s += partial[i];        // thread-strided, fixed order
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    float gn = 0.f, coef = 1.f;
    if (partial && max_norm > 0.f) {
      gn = sqrtf(s);
      coef = fminf(1.f, max_norm / (gn + 1e-6f));                                      // clip_grad_norm_: max_norm / (norm + 1e-6), clamped to 1
    }
    const float step = state[ST_STEP] + 1.f;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    state[ST_STEP] = step;
    state[ST_GNORM] = gn;
    state[ST_CLIP] = coef;
    state[ST_STEPSIZE] = (float)(lr / bc1);
    state[ST_RSQRT_BC2] = (float)(1.0 / sqrt(bc2));
  }
}

struct AdamHyper { float beta1, beta2, omb1, omb2, eps, weight_decay; };   // omb = 1 - beta, rounded once from double

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float* vmax, const AdamHyper& h, float clip,
                                         float step_size, float rsqrt_bc2) {
  g *= clip;
  g = fmaf(h.weight_decay, p, g);                     // L2 (coupled) weight decay: grad += wd * p
  m = fmaf(h.omb1, g - m, m);                         // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(h.beta2, v, h.omb2 * g * g);
  float vv = v;
  if (vmax) { vv = fmaxf(*vmax, v); *vmax = vv; }
  const float denom = fmaf(sqrtf(vv), rsqrt_bc2, h.eps);
  p -= step_size * (m / denom);
}

template <bool AMSGRAD>
__global__ void adam_mt_kernel(const long long* __restrict__ pptrs, const long long* __restrict__ gptrs,
                               const long long* __restrict__ mptrs, const long long* __restrict__ vptrs,
                               const long long* __restrict__ xptrs, const long long* __restrict__ sizes,
                               const int* __restrict__ blk_tensor, const int* __restrict__ blk_chunk, AdamHyper h,
                               const float* __restrict__ state, int write_grads) {
  const int t = blk_tensor[blockIdx.x];
  float* p = reinterpret_cast<float*>(pptrs[t]);
  float* g = reinterpret_cast<float*>(gptrs[t]);
  float* m = reinterpret_cast<float*>(mptrs[t]);
  float* v = reinterpret_cast<float*>(vptrs[t]);
  float* x = AMSGRAD ? reinterpret_cast<float*>(xptrs[t]) : nullptr;
  const long long n = sizes[t], lo = (long long)blk_chunk[blockIdx.x] * misc::MT_CHUNK, hi = min(n, lo + misc::MT_CHUNK);
  const float clip = state[ST_CLIP], step_size = state[ST_STEPSIZE], rs = state[ST_RSQRT_BC2];
  uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v);
  if (AMSGRAD) al |= reinterpret_cast<uintptr_t>(x);
  long long tail = lo;
  if (!(al & 15)) {
    const long long hi4 = lo + ((hi - lo) & ~3ll);
    for (long long i = lo + 4 * threadIdx.x; i < hi4; i += 4 * blockDim.x) {
      float4 P = *reinterpret_cast<const float4*>(p + i), G = *reinterpret_cast<const float4*>(g + i);
      float4 M = *reinterpret_cast<const float4*>(m + i), V = *reinterpret_cast<const float4*>(v + i);
      float4 X = make_float4(0.f, 0.f, 0.f, 0.f);
      if (AMSGRAD) X = *reinterpret_cast<const float4*>(x + i);
      adam_one(P.x, G.x, M.x, V.x, AMSGRAD ? &X.x : nullptr, h, clip, step_size, rs);
      adam_one(P.y, G.y, M.y, V.y, AMSGRAD ? &X.y : nullptr, h, clip, step_size, rs);
      adam_one(P.z, G.z, M.z, V.z, AMSGRAD ? &X.z : nullptr, h, clip, step_size, rs);
      adam_one(P.w, G.w, M.w, V.w, AMSGRAD ? &X.w : nullptr, h, clip, step_size, rs);
      *reinterpret_cast<float4*>(p + i) = P;
      *reinterpret_cast<float4*>(m + i) = M;
      *reinterpret_cast<float4*>(v + i) = V;
      if (AMSGRAD) *reinterpret_cast<float4*>(x + i) = X;
      if (write_grads) *reinterpret_cast<float4*>(g + i) = make_float4(G.x * clip, G.y * clip, G.z * clip, G.w * clip);
    }
    tail = hi4;
  }
  for (long long i = tail + threadIdx.x; i < hi; i += blockDim.x) {
    float P = p[i], M = m[i], V = v[i], X = AMSGRAD ? x[i] : 0.f;
    const float G = g[i];
    adam_one(P, G, M, V, AMSGRAD ? &X : nullptr, h, clip, step_size, rs);
    p[i] = P; m[i] = M; v[i] = V;
    if (AMSGRAD) x[i] = X;
    if (write_grads) g[i] = G * clip;
  }
}

static int adam_step(const long long* pptrs, const long long* gptrs, const long long* mptrs, const long long* vptrs,
                     const long long* xptrs, const long long* sizes, int n, const int* blk_tensor, const int* blk_chunk,
                     int n_blocks, double lr, double beta1, double beta2, double eps, double weight_decay, double max_grad_norm,
                     float* partial, float* state, int write_clipped_grads, cudaStream_t st) {
  if (n <= 0 || n_blocks <= 0) return 0;
  if (!pptrs || !gptrs || !mptrs || !vptrs || !sizes || !blk_tensor || !blk_chunk || !state) return RECNET_ERR_BAD_SHAPE;
  if (!(lr >= 0.0) || !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.0)) return RECNET_ERR_BAD_SHAPE;
  const bool clip = max_grad_norm > 0.0;
  if (clip) {
    if (!partial) return RECNET_ERR_BAD_SHAPE;
    misc::mt_sumsq_kernel<<<n_blocks, 256, 0, st>>>(gptrs, sizes, blk_tensor, blk_chunk, partial);
    RN_LAUNCH_OK();
  }
  adam_prologue_kernel<<<1, 512, 0, st>>>(clip ? partial : nullptr, n_blocks, (float)max_grad_norm, lr, beta1, beta2, state);
  RN_LAUNCH_OK();
  AdamHyper h{(float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, (float)weight_decay};
  const int wg = (clip && write_clipped_grads) ? 1 : 0;
  if (xptrs) adam_mt_kernel<true><<<n_blocks, 256, 0, st>>>(pptrs, gptrs, mptrs, vptrs, xptrs, sizes, blk_tensor, blk_chunk, h, state, wg);
  else adam_mt_kernel<false><<<n_blocks, 256, 0, st>>>(pptrs, gptrs, mptrs, vptrs, nullptr, sizes, blk_tensor, blk_chunk, h, state, wg);
  RN_LAUNCH_OK();
  return 0;
}
}  // namespace optim
