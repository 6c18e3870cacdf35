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
}  // namespace sscvae

using namespace sscvae;
extern "C" {
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
