// Fused bottom-up/top-down region attention (north_star kernel #2), forward and backward.
//
// forward, per row r (a caption in training, an (image,state,beam) row in decode):
//   u_n   = w_a . tanh(q_r + P[img(r), n, :])             (attention.py:69-88)
//   alpha = masked_softmax(u, mask[img(r)])                (allennlp: softmax(u*m)*m / (sum + 1e-13))
//   xhat  = sum_n alpha_n * x[img(r), n, :]                (updown_cell.py:156-158)
// One 4-CTA thread-block cluster per row (boxes / feature slices / projection columns split across the CTAs,
// the N scores exchanged through distributed shared memory); nothing intermediate goes to global memory.
// HBM/L2 traffic per row: N*Ap + N*Fp bf16 elements read once (16-byte coalesced vectors), F written.
#include "kernels.cuh"
#include "prof.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace sscvae {

#define LAUNCHED() do { CUDA_TRY(cudaGetLastError()); ++g_launch_count_pw; } while (0)

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

static constexpr int ATT_THREADS = 256;
static constexpr int CL = 4;      // CTAs per row: a thread-block cluster splits boxes, features and projection columns

// scores u_n for boxes [n_lo, n_hi) of this row: warp per box, lanes over the projection axis
__device__ __forceinline__ void attn_scores(const AttnArgs& a, const bf16* __restrict__ proj_img,
                                            const float* __restrict__ mask_img, const float* q_s, const float* wa_s,
                                            float* u_s, int n_lo, int n_hi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int nvec = a.Ap >> 3;
  for (int n = n_lo + warp; n < n_hi; n += nwarp) {
    float s = 0.f;
    if (mask_img[n] != 0.f) {                      // masked boxes enter the softmax as u*m = 0
      const bf16x8* p = reinterpret_cast<const bf16x8*>(proj_img + (size_t)n * a.Ap);
#pragma unroll 4
      for (int i = lane; i < nvec; i += 32) {
        const bf16x8 v = p[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __bfloat1622float2(v.v[k]);
          const int idx = i * 8 + 2 * k;
          s += wa_s[idx] * tanh_approx(q_s[idx] + f.x) + wa_s[idx + 1] * tanh_approx(q_s[idx + 1] + f.y);
        }
      }
      s = warp_sum(s);
    }
    if (lane == 0) u_s[n] = s;
  }
}

// After every CTA of the cluster has filled its slice [rank*nper, ...) of a per-box array in its own shared
// memory, copy the other slices over distributed shared memory so each CTA holds all N values.
__device__ __forceinline__ void cluster_gather(cg::cluster_group& cluster, float* arr, int N, int nper, int rank) {
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const int owner = n / nper;
    if (owner != rank) arr[n] = *cluster.map_shared_rank(arr + n, owner);
  }
}

// masked softmax by warp 0. On exit: s_s[n] = softmax(u*m)[n], al_s[n] = alpha[n]; returns R = sum r + 1e-13
__device__ __forceinline__ float attn_softmax(int N, const float* mask_img, const float* u_s, float* s_s, float* al_s) {
  const int lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int n = lane; n < N; n += 32) mx = fmaxf(mx, u_s[n] * mask_img[n]);
  mx = warp_max(mx);
  float se = 0.f;
  for (int n = lane; n < N; n += 32) { const float e = __expf(u_s[n] * mask_img[n] - mx); s_s[n] = e; se += e; }
  se = warp_sum(se);
  float sr = 0.f;
  for (int n = lane; n < N; n += 32) { const float sv = s_s[n] / se; s_s[n] = sv; sr += sv * mask_img[n]; }
  sr = warp_sum(sr);
  const float Rn = sr + 1e-13f;
  for (int n = lane; n < N; n += 32) al_s[n] = s_s[n] * mask_img[n] / Rn;
  return Rn;
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(ATT_THREADS)
attention_fwd_kernel(AttnArgs a, float* __restrict__ alpha, bf16* __restrict__ xhat, int ld_x) {
  extern __shared__ float sm[];
  float* q_s = sm;                      // Ap
  float* wa_s = q_s + a.Ap;             // Ap
  float* u_s = wa_s + a.Ap;             // N
  float* s_s = u_s + a.N;               // N
  float* al_s = s_s + a.N;              // N
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int r = blockIdx.x / CL;
  const int img = a.rowmap ? a.rowmap[r] : r;
  const bf16* proj_img = a.proj + (size_t)img * a.N * a.Ap;
  const bf16* feat_img = a.feats + (size_t)img * a.N * a.Fp;
  const float* mask_img = a.mask + (size_t)img * a.N;
  const int nper = (a.N + CL - 1) / CL;
  for (int i = threadIdx.x; i < a.Ap; i += blockDim.x) {
    q_s[i] = (i < a.A) ? a.q[(size_t)r * a.ld_q + i] : 0.f;
    wa_s[i] = (i < a.A) ? a.w_a[i] : 0.f;
  }
  __syncthreads();
  attn_scores(a, proj_img, mask_img, q_s, wa_s, u_s, rank * nper, min(a.N, (rank + 1) * nper));
  cluster.sync();
  cluster_gather(cluster, u_s, a.N, nper, rank);
  cluster.sync();                       // nobody's shared memory is read remotely after this point
  if (threadIdx.x < 32) attn_softmax(a.N, mask_img, u_s, s_s, al_s);
  __syncthreads();
  if (rank == 0)
    for (int n = threadIdx.x; n < a.N; n += blockDim.x) alpha[(size_t)r * a.N + n] = al_s[n];
  // weighted sum over this CTA's quarter of the feature axis; a thread owns 2 consecutive features (a warp
  // reads 128 contiguous bytes per box), NB independent loads in flight
  const int npair = a.Fp >> 1;
  const int pper = (npair + CL - 1) / CL;
  const int p_hi = min(npair, (rank + 1) * pper);
  constexpr int NB = 12;
  for (int i = rank * pper + threadIdx.x; i < p_hi; i += blockDim.x) {
    float acc0 = 0.f, acc1 = 0.f;
    const __nv_bfloat162* col = reinterpret_cast<const __nv_bfloat162*>(feat_img) + i;
    for (int n0 = 0; n0 < a.N; n0 += NB) {
      __nv_bfloat162 v[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) v[j] = col[(size_t)min(n0 + j, a.N - 1) * npair];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const float w = (n0 + j < a.N) ? al_s[n0 + j] : 0.f;
        const float2 f = __bfloat1622float2(v[j]);
        acc0 += w * f.x; acc1 += w * f.y;
      }
    }
    *reinterpret_cast<__nv_bfloat162*>(xhat + (size_t)r * ld_x + 2 * i) = __floats2bfloat162_rn(acc0, acc1);
  }
}

int attention_forward(cudaStream_t s, const AttnArgs& a, float* alpha, bf16* xhat, int ld_x) {
  PROF_SCOPE(s, "attention_fwd", 0, (double)a.R*((double)a.N*(a.Ap+a.Fp)*2.0 + a.A*4.0 + a.Fp*2.0 + a.N*4.0));
  const size_t smem = (size_t)(2 * a.Ap + 3 * a.N) * sizeof(float);
  attention_fwd_kernel<<<a.R * CL, ATT_THREADS, smem, s>>>(a, alpha, xhat, ld_x);
  LAUNCHED();
  return 0;
}

// backward of the same three fused ops. dproj_acc (images,N,A) and dwa_acc (R,A) are accumulated
// across timesteps by the owning CTA (row r == image r in training), so no atomics are needed.
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(ATT_THREADS)
attention_bwd_kernel(AttnArgs a, const float* __restrict__ alpha_in, const float* __restrict__ dxhat, int ld_dx,
                     bf16* __restrict__ dq, int ld_dq, float* __restrict__ dproj_acc, float* __restrict__ dwa_acc) {
  extern __shared__ float sm[];
  float* q_s = sm;                      // Ap
  float* wa_s = q_s + a.Ap;             // Ap
  float* dx_s = wa_s + a.Ap;            // Fp
  float* u_s = dx_s + a.Fp;             // N
  float* s_s = u_s + a.N;               // N
  float* al_s = s_s + a.N;              // N
  float* da_s = al_s + a.N;             // N  (d alpha, then d u)
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int r = blockIdx.x / CL;
  const int img = a.rowmap ? a.rowmap[r] : r;
  const bf16* proj_img = a.proj + (size_t)img * a.N * a.Ap;
  const bf16* feat_img = a.feats + (size_t)img * a.N * a.Fp;
  const float* mask_img = a.mask + (size_t)img * a.N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int nper = (a.N + CL - 1) / CL;
  const int n_lo = rank * nper, n_hi = min(a.N, (rank + 1) * nper);
  for (int i = threadIdx.x; i < a.Ap; i += blockDim.x) {
    q_s[i] = (i < a.A) ? a.q[(size_t)r * a.ld_q + i] : 0.f;
    wa_s[i] = (i < a.A) ? a.w_a[i] : 0.f;
  }
  for (int i = threadIdx.x; i < a.Fp; i += blockDim.x) dx_s[i] = (i < a.F) ? dxhat[(size_t)r * ld_dx + i] : 0.f;
  __syncthreads();
  // d alpha_n = dxhat . x_n for this CTA's boxes
  const int fvec = a.Fp >> 3;
  for (int n = n_lo + warp; n < n_hi; n += nwarp) {
    float s = 0.f;
    if (mask_img[n] != 0.f) {
      const bf16x8* p = reinterpret_cast<const bf16x8*>(feat_img + (size_t)n * a.Fp);
#pragma unroll 8
      for (int i = lane; i < fvec; i += 32) {
        const bf16x8 v = p[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __bfloat1622float2(v.v[k]);
          s += dx_s[i * 8 + 2 * k] * f.x + dx_s[i * 8 + 2 * k + 1] * f.y;
        }
      }
      s = warp_sum(s);
    }
    if (lane == 0) da_s[n] = s;
  }
  attn_scores(a, proj_img, mask_img, q_s, wa_s, u_s, n_lo, n_hi);
  cluster.sync();
  cluster_gather(cluster, u_s, a.N, nper, rank);
  cluster_gather(cluster, da_s, a.N, nper, rank);
  cluster.sync();
  if (threadIdx.x < 32) {
    const float Rn = attn_softmax(a.N, mask_img, u_s, s_s, al_s);
    // alpha = r / R with r = s*m :  dr = (dalpha - sum_k dalpha_k alpha_k) / R ; ds = dr*m
    float dot = 0.f;
    for (int n = lane; n < a.N; n += 32) dot += da_s[n] * al_s[n];
    dot = warp_sum(dot);
    float dss = 0.f;
    for (int n = lane; n < a.N; n += 32) {
      const float ds = (da_s[n] - dot) / Rn * mask_img[n];
      da_s[n] = ds;
      dss += ds * s_s[n];
    }
    dss = warp_sum(dss);
    // softmax backward on x = u*m, then du = dx * m
    for (int n = lane; n < a.N; n += 32) da_s[n] = s_s[n] * (da_s[n] - dss) * mask_img[n];
  }
  __syncthreads();
  // this CTA's quarter of the projection columns: dq_a = sum_n du_n w_a (1 - th^2), dP_na += du_n w_a (1 - th^2),
  // dw_a += du_n th. Boxes are processed NB at a time with all loads issued before any use.
  constexpr int NB = 6;
  const int cper = (ld_dq + CL - 1) / CL;
  const int c_hi = min(ld_dq, (rank + 1) * cper);
  for (int i = rank * cper + threadIdx.x; i < c_hi; i += blockDim.x) {
    float dqa = 0.f, dwa = 0.f;
    if (i < a.A) {
      const float qa = q_s[i], wa = wa_s[i];
      float* acc = dproj_acc + (size_t)img * a.N * a.A + i;
      const bf16* pj = proj_img + i;
      for (int n0 = 0; n0 < a.N; n0 += NB) {
        float pv[NB], av[NB];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          const int n = n0 + k;
          const bool ok = n < a.N;
          pv[k] = ok ? __bfloat162float(pj[(size_t)n * a.Ap]) : 0.f;
          av[k] = ok ? acc[(size_t)n * a.A] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          const int n = n0 + k;
          if (n < a.N) {
            const float du = da_s[n];
            const float th = tanh_approx(qa + pv[k]);
            const float g = du * wa * (1.f - th * th);
            dqa += g;
            dwa += du * th;
            acc[(size_t)n * a.A] = av[k] + g;
          }
        }
      }
      dwa_acc[(size_t)r * a.A + i] += dwa;
    }
    dq[(size_t)r * ld_dq + i] = __float2bfloat16_rn(dqa);
  }
}

int attention_backward(cudaStream_t s, const AttnArgs& a, const float* alpha, const float* dxhat, int ld_dx, bf16* dq,
                       int ld_dq, float* dproj_acc, float* dwa_acc) {
  PROF_SCOPE(s, "attention_bwd", 0, (double)a.R*((double)a.N*(2.0*a.Ap+a.Fp)*2.0 + (double)a.N*a.A*8.0 + a.Fp*4.0 + a.A*6.0));
  const size_t smem = (size_t)(2 * a.Ap + a.Fp + 4 * a.N) * sizeof(float);
  attention_bwd_kernel<<<a.R * CL, ATT_THREADS, smem, s>>>(a, alpha, dxhat, ld_dx, dq, ld_dq, dproj_acc, dwa_acc);
  LAUNCHED();
  return 0;
}

}  // namespace sscvae
