// Training-step tail: global grad-norm clipping + SGD(momentum, weight decay) fused
// over a flat fp32 parameter buffer; replaces clip_grad_norm_ / optimizer.step of
// var_updown/scripts/train.py:173-176 (~100 small launches in eager PyTorch) with three launches.
#include "../../include/sscvae.h"
#include "common.cuh"

namespace sscvae {
extern unsigned long long g_launch_count_pw;
#define LAUNCHED() do { CUDA_TRY(cudaGetLastError()); ++g_launch_count_pw; } while (0)

static constexpr int NORM_BLOCKS = 1024;

__global__ void sqnorm_partial_kernel(const float* __restrict__ g, size_t n, float* __restrict__ partial) {
  __shared__ float red[32];
  float s = 0.f;
  const size_t n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (size_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s += g[i] * g[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}
__global__ void sqnorm_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) *out = v;
  }
}

__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, size_t n,
                           const float* __restrict__ sqnorm, float max_norm, float lr, float momentum, float wd,
                           int first) {
  // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), applied only when < 1
  float coef = 1.f;
  if (sqnorm && max_norm > 0.f) coef = fminf(1.f, max_norm / (sqrtf(*sqnorm) + 1e-6f));
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float w = p[i];
    float d = g[i] * coef + wd * w;
    if (momentum != 0.f) {
      const float b = first ? d : momentum * m[i] + d;
      m[i] = b;
      d = b;
    }
    p[i] = w - lr * d;
  }
}
// ---- multi-tensor form: every parameter of the model in three launches ---------------------------------
static constexpr int MT_MAX = 32;
static constexpr unsigned MT_CHUNK = 16384;            // elements per block
struct MultiTensorTable {
  float* p[MT_MAX]; const float* g[MT_MAX]; float* m[MT_MAX];
  unsigned long long n[MT_MAX];
  unsigned chunk0[MT_MAX + 1];                         // first chunk (block) of tensor i
  int first[MT_MAX];
  int count;
};

__device__ __forceinline__ int mt_find(const MultiTensorTable& t, unsigned block) {
  int i = 0;
  while (i + 1 < t.count && block >= t.chunk0[i + 1]) ++i;
  return i;
}

__global__ void __launch_bounds__(256) mt_sqnorm_kernel(const __grid_constant__ MultiTensorTable t, float gscale, float* __restrict__ partial) {
  __shared__ float red[8];
  const int ti = mt_find(t, blockIdx.x);
  const size_t lo = (size_t)(blockIdx.x - t.chunk0[ti]) * MT_CHUNK;
  const size_t hi = min((size_t)t.n[ti], lo + MT_CHUNK);
  const float* g = t.g[ti];
  float s = 0.f;
  if (g)                                               // NULL gradient = all zeros (see sscvae_sgd_step_multi)
    for (size_t i = lo + threadIdx.x; i < hi; i += 256) { const float v = g[i] * gscale; s += v * v; }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w];
    partial[blockIdx.x] = v;                           // fixed chunking + ordered sums: deterministic
  }
}

__global__ void __launch_bounds__(256) mt_sgd_kernel(const __grid_constant__ MultiTensorTable t, const float* __restrict__ sqnorm,
                                                     float max_norm, float lr, float momentum, float wd, float gscale) {
  const int ti = mt_find(t, blockIdx.x);
  const size_t lo = (size_t)(blockIdx.x - t.chunk0[ti]) * MT_CHUNK;
  const size_t hi = min((size_t)t.n[ti], lo + MT_CHUNK);
  float coef = 1.f;
  if (max_norm > 0.f) coef = fminf(1.f, max_norm / (sqrtf(*sqnorm) + 1e-6f));   // torch clip_grad_norm_
  coef *= gscale;                                      // gradients arrive as a SUM over ranks: the mean is taken here
  float* p = t.p[ti]; const float* g = t.g[ti]; float* m = t.m[ti];
  const bool first = t.first[ti] != 0;
  for (size_t i = lo + threadIdx.x; i < hi; i += 256) {
    const float w = p[i];
    float d = (g ? g[i] * coef : 0.f) + wd * w;
    if (momentum != 0.f) {
      const float b = first ? d : momentum * m[i] + d;
      m[i] = b;
      d = b;
    }
    p[i] = w - lr * d;
  }
}
}  // namespace sscvae

using namespace sscvae;
extern "C" {
int sscvae_sgd_step_multi(int count, void* const* params, const void* const* grads, void* const* momentum_bufs,
                          const uint64_t* sizes, const int32_t* first_step, float max_norm, float lr, float momentum,
                          float weight_decay, float grad_scale, float* scratch, size_t scratch_floats, void* stream) {
  REQUIRE(count > 0 && count <= MT_MAX && params && grads && sizes && first_step && scratch, "bad argument (at most %d tensors)", MT_MAX);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MultiTensorTable t;
  unsigned chunks = 0;
  for (int i = 0; i < count; ++i) {
    // grads[i] == NULL: a zero gradient (torch 1.1's zero_grad() leaves ZERO tensors behind, so a parameter frozen by
    // train.py:156-161 keeps decaying and coasting on its momentum; see FusedClipSGD(legacy_zero_grad=True))
    REQUIRE(params[i] && (momentum == 0.f || (momentum_bufs && momentum_bufs[i])), "NULL tensor %d", i);
    t.p[i] = reinterpret_cast<float*>(params[i]); t.g[i] = reinterpret_cast<const float*>(grads[i]);
    t.m[i] = momentum_bufs ? reinterpret_cast<float*>(momentum_bufs[i]) : nullptr;
    t.n[i] = sizes[i]; t.first[i] = first_step[i]; t.chunk0[i] = chunks;
    chunks += (unsigned)((sizes[i] + MT_CHUNK - 1) / MT_CHUNK);
  }
  t.chunk0[count] = chunks; t.count = count;
  REQUIRE(scratch_floats >= (size_t)chunks + 1, "scratch too small: need %u floats", chunks + 1);
  mt_sqnorm_kernel<<<chunks, 256, 0, st>>>(t, grad_scale, scratch);
  LAUNCHED();
  sqnorm_final_kernel<<<1, 1024, 0, st>>>(scratch, (int)chunks, scratch + chunks);
  LAUNCHED();
  mt_sgd_kernel<<<chunks, 256, 0, st>>>(t, scratch + chunks, max_norm, lr, momentum, weight_decay, grad_scale);
  LAUNCHED();
  return 0;
}

int sscvae_grad_sqnorm(const float* grads, size_t n, float* partial, float* sqnorm_out, void* stream) {
  REQUIRE(grads && partial && sqnorm_out, "NULL argument");
  REQUIRE((reinterpret_cast<uintptr_t>(grads) & 15) == 0, "grads must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  sqnorm_partial_kernel<<<NORM_BLOCKS, 256, 0, st>>>(grads, n, partial);
  LAUNCHED();
  sqnorm_final_kernel<<<1, 1024, 0, st>>>(partial, NORM_BLOCKS, sqnorm_out);
  LAUNCHED();
  return 0;
}
int sscvae_sgd_step(float* params, const float* grads, float* momentum_buf, size_t n, const float* sqnorm, float max_norm,
                    float lr, float momentum, float weight_decay, int first_step, void* stream) {
  REQUIRE(params && grads && (momentum == 0.f || momentum_buf), "NULL argument");
  sgd_kernel<<<1184, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(params, grads, momentum_buf, n, sqnorm, max_norm,
                                                                      lr, momentum, weight_decay, first_step);
  LAUNCHED();
  return 0;
}
}
