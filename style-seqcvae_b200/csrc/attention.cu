// Fused bottom-up/top-down region attention (north_star kernel #2), forward and backward.
//
// forward, per row r (a caption in training, an (image,state,beam) row in decode):
//   u_n   = w_a . tanh(q_r + P[img(r), n, :])             (attention.py:69-88)
//   alpha = masked_softmax(u, mask[img(r)])                (allennlp: softmax(u*m)*m / (sum + 1e-13))
//   xhat  = sum_n alpha_n * x[img(r), n, :]                (updown_cell.py:156-158)
//
// The op is a stream over the image's projected features P (N*Ap bf16) and region features x (N*Fp bf16),
// ~200 KB per row at the shipped dims, with almost no arithmetic: it is bound by bytes in flight. So the
// kernels are persistent (<= one CTA per SM, a contiguous block of rows each) and warp-specialised:
//   warp 8     : producer. One lane issues 1-D bulk async copies (cp.async.bulk -> the TMA engine) of whole
//                box-row chunks into a 4 x 32 KB shared-memory ring, completion on mbarriers. It runs ahead of
//                the consumers across chunk, phase and ROW boundaries, so the next row's data is already in
//                flight while the current row's softmax runs.
//   warps 0..7 : consumers. Scores (warp per box), masked softmax (redundantly per warp, in registers), weighted
//                sum (a thread owns 8 consecutive features). Per-row vectors (q, d xhat, saved softmax) are
//                prefetched one row ahead with cp.async into double buffers.
// Nothing intermediate goes to global memory. HBM/L2 traffic per row: N*(Ap+Fp) bf16 read once, F written.
//
// backward is split in two:
//   per step  : d alpha_n = d xhat . x_n, softmax backward -> d u (saved, (R,N) fp32), and
//               d q_a = w_a sum_n d u_n (1 - tanh^2(q_a + P_na))                        [on the BPTT critical path]
//   deferred  : d P[b,n,a] = w_a sum_t d u[t,b,n] (1 - tanh^2(q[t,b,a] + P[b,n,a])) and d w_a, ONCE after the
//               time loop, instead of a 2 x 28 MB fp32 read-modify-write of d P at every step.
#include "kernels.cuh"
#include "prof.cuh"
#include "ptx.cuh"
#include "attention_dev.cuh"

namespace sscvae {

#define LAUNCHED() do { CUDA_TRY(cudaGetLastError()); ++g_launch_count_pw; } while (0)

using namespace attn;


__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_fwd_kernel(AttnArgs a, AttnPlan pl, float* __restrict__ alpha, float* __restrict__ smx, bf16* __restrict__ xhat,
                     int ld_x) {
  extern __shared__ __align__(128) uint8_t att_smem_raw[];
  const AttnSmem sm = carve(att_smem_raw, a, false);
  attn_prologue(sm, a, false);
  pdl_wait();                                        // the prologue overlaps the previous kernel's tail
  pdl_launch_dependents(2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r_begin = blockIdx.x * pl.rows_per_cta;
  const int r_end = min(a.R, r_begin + pl.rows_per_cta);
  Ring ring;

  if (warp == ATT_CWARPS) {                          // ---- producer
    if (lane == 0) {
      const uint64_t pol = a.l2_policy == 1 ? ptx::l2_policy_evict_first() : a.l2_policy == 2 ? ptx::l2_policy_evict_last() : 0;
      for (int r = r_begin; r < r_end; ++r) {
        const int img = a.rowmap ? a.rowmap[r] : r;
        produce_block(sm, ring, reinterpret_cast<const uint8_t*>(a.proj + (size_t)img * a.N * a.Ap), a.N, a.Ap * 2, pl.nP, pl.bP, pol);
        produce_block(sm, ring, reinterpret_cast<const uint8_t*>(a.feats + (size_t)img * a.N * a.Fp), a.N, a.Fp * 2, pl.nF, pl.bF, pol);
      }
    }
    return;
  }
  // ---- consumers
  if (r_begin < r_end) {
    prefetch_vec(sm.q(0), a.q + (size_t)r_begin * a.ld_q, a.A);
    ptx::cp_async_commit();
  }
  int cur = 0;
  for (int r = r_begin; r < r_end; ++r, cur ^= 1) {
    const int img = a.rowmap ? a.rowmap[r] : r;
    attn_fwd_row(a, pl, sm, ring, cur, a.mask + (size_t)img * a.N, r + 1 < r_end ? a.q + (size_t)(r + 1) * a.ld_q : nullptr,
                 alpha + (size_t)r * a.N, smx ? smx + (size_t)r * a.N : nullptr, xhat + (size_t)r * ld_x);
  }
}

static int make_plan(const AttnArgs& a, AttnPlan& pl) {
  if (a.N > 32 * ATT_NREG || a.Fp > 8 * ATT_CONSUMERS * ATT_FV || a.Ap > 2 * ATT_CONSUMERS * ATT_PV ||
      a.Ap * 2 > ATT_STAGE_BYTES || a.Fp * 2 > ATT_STAGE_BYTES || attn_smem_bytes(a, true) > 112 * 1024) {
    set_error("attention: unsupported shape N=%d (<= %d) Fp=%d (<= %d) Ap=%d (<= %d)", a.N, 32 * ATT_NREG, a.Fp,
              8 * ATT_CONSUMERS * ATT_FV, a.Ap, 2 * ATT_CONSUMERS * ATT_PV);
    return SSCVAE_ERR_UNSUPPORTED;
  }
  REQUIRE((a.Ap % 8) == 0 && (a.Fp % 8) == 0, "attention: Ap/Fp must be multiples of 8");
  auto boxes_per_chunk = [](int row_bytes) {
    int b = 1;
    while (b * 2 <= ATT_MAXB && b * 2 * row_bytes <= ATT_STAGE_BYTES) b *= 2;
    return b;
  };
  pl.bP = boxes_per_chunk(a.Ap * 2); pl.nP = ceil_div(a.N, pl.bP);
  pl.bF = boxes_per_chunk(a.Fp * 2); pl.nF = ceil_div(a.N, pl.bF);
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  pl.rows_per_cta = ceil_div(a.R, 2 * n_sm);
  return 0;
}

static int attn_l2_policy() {
  // SSCVAE_ATT_POLICY: 0 = no hint, 1 = evict_first (default: the stream must not push the recurrent weights out of L2),
  // 2 = evict_last (with SSCVAE_W_EVICT_FIRST=1: keep the features resident, let the weights stream)
  static const int v = [] { const char* e = getenv("SSCVAE_ATT_POLICY"); return e ? atoi(e) : 1; }();
  return v;
}

int attention_forward(cudaStream_t s, const AttnArgs& a_in, float* alpha, float* smx, bf16* xhat, int ld_x) {
  AttnArgs a = a_in;
  a.l2_policy = attn_l2_policy();
  PROF_SCOPE(s, "attention_fwd", 0, (double)a.R*((double)a.N*(a.Ap+a.Fp)*2.0 + a.A*4.0 + a.Fp*2.0 + a.N*4.0));
  REQUIRE(a.R > 0 && (ld_x % 8) == 0 && (reinterpret_cast<uintptr_t>(xhat) & 15) == 0, "attention_forward: bad xhat layout");
  AttnPlan pl;
  TRY(make_plan(a, pl));
  const size_t smem = attn_smem_bytes(a, false);
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    configured = true;
  }
  CUDA_TRY(launch_pdl(attention_fwd_kernel, dim3(ceil_div(a.R, pl.rows_per_cta)), dim3(ATT_THREADS), smem, s, a, pl, alpha, smx, xhat, ld_x));
  LAUNCHED();
  return 0;
}

// ---- per-step backward: d u (saved) and d q ------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_bwd_kernel(AttnArgs a, AttnPlan pl, const float* __restrict__ smx, const float* __restrict__ dxhat, int ld_dx,
                     bf16* __restrict__ dq, int ld_dq, float* __restrict__ du) {
  extern __shared__ __align__(128) uint8_t att_smem_raw[];
  const AttnSmem sm = carve(att_smem_raw, a, true);
  attn_prologue(sm, a, true);
  pdl_wait();                                        // the prologue overlaps the previous kernel's tail
  pdl_launch_dependents(2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r_begin = blockIdx.x * pl.rows_per_cta;
  const int r_end = min(a.R, r_begin + pl.rows_per_cta);
  Ring ring;

  if (warp == ATT_CWARPS) {                          // ---- producer: region features first, then projections
    if (lane == 0) {
      const uint64_t pol = a.l2_policy == 1 ? ptx::l2_policy_evict_first() : a.l2_policy == 2 ? ptx::l2_policy_evict_last() : 0;
      for (int r = r_begin; r < r_end; ++r) {
        const int img = a.rowmap ? a.rowmap[r] : r;
        produce_block(sm, ring, reinterpret_cast<const uint8_t*>(a.feats + (size_t)img * a.N * a.Fp), a.N, a.Fp * 2, pl.nF, pl.bF, pol);
        produce_block(sm, ring, reinterpret_cast<const uint8_t*>(a.proj + (size_t)img * a.N * a.Ap), a.N, a.Ap * 2, pl.nP, pl.bP, pol);
      }
    }
    return;
  }
  // ---- consumers
  auto prefetch_row = [&](int r, int slot) {
    prefetch_vec(sm.q(slot), a.q + (size_t)r * a.ld_q, a.A);
    prefetch_vec(sm.dx(slot), dxhat + (size_t)r * ld_dx, a.F);
    prefetch_vec(sm.sv(slot), smx + (size_t)r * a.N, a.N);
    ptx::cp_async_commit();
  };
  if (r_begin < r_end) prefetch_row(r_begin, 0);
  const int nfv = a.Fp >> 3, npair = a.Ap >> 1;
  int cur = 0;
  for (int r = r_begin; r < r_end; ++r, cur ^= 1) {
    const int img = a.rowmap ? a.rowmap[r] : r;
    const float* mask_img = a.mask + (size_t)img * a.N;
    ptx::cp_async_wait_all();
    ptx::bar_sync(1, ATT_CONSUMERS);
    if (r + 1 < r_end) prefetch_row(r + 1, cur ^ 1);
    const float* dx_s = sm.dx(cur);
    // d alpha_n = d xhat . x_n : warp per box
    for (int c = 0; c < pl.nF; ++c) {
      const int n0 = c * pl.bF, nb = min(pl.bF, a.N - n0);
      ptx::mbar_wait(&sm.full[ring.stage], ring.phase);
      const bf16x8* buf = reinterpret_cast<const bf16x8*>(sm.stage + (size_t)ring.stage * ATT_STAGE_BYTES);
      const int wpb = warps_per_box(nb);
      const int j = warp / wpb, sub = warp % wpb;
      if (j < nb) {
        const int n = n0 + j;
        float s = 0.f;
        if (mask_img[n] != 0.f) {
          const bf16x8* p = buf + (size_t)j * nfv;
          for (int i = sub * 32 + lane; i < nfv; i += 32 * wpb) {
            const bf16x8 v = p[i];
            const float4 da = *reinterpret_cast<const float4*>(dx_s + i * 8);
            const float4 db = *reinterpret_cast<const float4*>(dx_s + i * 8 + 4);
            const float2 f0 = __bfloat1622float2(v.v[0]), f1 = __bfloat1622float2(v.v[1]);
            const float2 f2 = __bfloat1622float2(v.v[2]), f3 = __bfloat1622float2(v.v[3]);
            s += da.x * f0.x + da.y * f0.y + da.z * f1.x + da.w * f1.y;
            s += db.x * f2.x + db.y * f2.y + db.z * f3.x + db.w * f3.y;
          }
          s = warp_sum(s);
        }
        if (lane == 0) sm.u[sub * sm.N4 + n] = s;
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm.empty[ring.stage]);
      ring.advance();
    }
    ptx::bar_sync(1, ATT_CONSUMERS);
    // softmax backward, redundantly per warp. alpha = r/R with r = s*m:  dr = (dalpha - sum_k dalpha_k alpha_k)/R,
    // ds = dr*m; softmax backward on x = u*m; du = dx*m
    float duv[ATT_NREG];
    {
      float m[ATT_NREG], sv[ATT_NREG], da[ATT_NREG];
      float sr = 0.f;
#pragma unroll
      for (int k = 0; k < ATT_NREG; ++k) {
        const int n = lane + 32 * k;
        const bool ok = n < a.N;
        m[k] = ok ? mask_img[n] : 0.f;
        sv[k] = ok ? sm.sv(cur)[n] : 0.f;
        da[k] = ok ? gather_partial(sm.u, sm.N4, a.N, pl.bF, n) : 0.f;
        sr += sv[k] * m[k];
      }
      const float Rn = warp_sum(sr) + 1e-13f;
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < ATT_NREG; ++k) dot += da[k] * (sv[k] * m[k] / Rn);
      dot = warp_sum(dot);
      float dss = 0.f;
#pragma unroll
      for (int k = 0; k < ATT_NREG; ++k) {
        da[k] = (da[k] - dot) / Rn * m[k];
        dss += da[k] * sv[k];
      }
      dss = warp_sum(dss);
#pragma unroll
      for (int k = 0; k < ATT_NREG; ++k) duv[k] = sv[k] * (da[k] - dss) * m[k];
    }
    if (warp == 0) {
#pragma unroll
      for (int k = 0; k < ATT_NREG; ++k) {
        const int n = lane + 32 * k;
        if (n < a.N) du[(size_t)r * a.N + n] = duv[k];
      }
    }
    // d q_a = w_a sum_n du_n (1 - tanh^2(q_a + P_na)): a thread owns pairs of projection columns
    const float* q_s = sm.q(cur);
    float g[ATT_PV][2];
#pragma unroll
    for (int v = 0; v < ATT_PV; ++v) g[v][0] = g[v][1] = 0.f;
    for (int c = 0; c < pl.nP; ++c) {
      const int n0 = c * pl.bP, nb = min(pl.bP, a.N - n0);
      ptx::mbar_wait(&sm.full[ring.stage], ring.phase);
      const __nv_bfloat162* buf = reinterpret_cast<const __nv_bfloat162*>(sm.stage + (size_t)ring.stage * ATT_STAGE_BYTES);
      for (int j = 0; j < nb; ++j) {
        const int n = n0 + j;
        const float d = __shfl_sync(0xffffffffu, pick(duv, n >> 5), n & 31);
        if (d != 0.f) {                              // warp-uniform (masked boxes, padded timesteps)
#pragma unroll
          for (int v = 0; v < ATT_PV; ++v) {
            const int cp = threadIdx.x + ATT_CONSUMERS * v;
            if (cp < npair) {
              const float2 f = __bfloat1622float2(buf[(size_t)j * npair + cp]);
              const float t0 = tanh_approx(q_s[2 * cp] + f.x), t1 = tanh_approx(q_s[2 * cp + 1] + f.y);
              g[v][0] += d * (1.f - t0 * t0);
              g[v][1] += d * (1.f - t1 * t1);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm.empty[ring.stage]);
      ring.advance();
    }
#pragma unroll
    for (int v = 0; v < ATT_PV; ++v) {
      const int cp = threadIdx.x + ATT_CONSUMERS * v;
      if (cp < npair && 2 * cp < ld_dq)
        *reinterpret_cast<__nv_bfloat162*>(dq + (size_t)r * ld_dq + 2 * cp) =
            __floats2bfloat162_rn(sm.wa[2 * cp] * g[v][0], sm.wa[2 * cp + 1] * g[v][1]);
    }
  }
}

int attention_backward(cudaStream_t s, const AttnArgs& a_in, const float* smx, const float* dxhat, int ld_dx, bf16* dq,
                       int ld_dq, float* du) {
  AttnArgs a = a_in;
  a.l2_policy = attn_l2_policy();
  PROF_SCOPE(s, "attention_bwd", 0, (double)a.R*((double)a.N*(a.Ap+a.Fp)*2.0 + a.Fp*4.0 + a.A*6.0 + a.N*8.0));
  REQUIRE(a.R > 0 && (ld_dq % 2) == 0 && (reinterpret_cast<uintptr_t>(dq) & 3) == 0, "attention_backward: bad dq layout");
  AttnPlan pl;
  TRY(make_plan(a, pl));
  const size_t smem = attn_smem_bytes(a, true);
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    configured = true;
  }
  CUDA_TRY(launch_pdl(attention_bwd_kernel, dim3(ceil_div(a.R, pl.rows_per_cta)), dim3(ATT_THREADS), smem, s, a, pl, smx, dxhat, ld_dx, dq,
                      ld_dq, du));
  LAUNCHED();
  return 0;
}

// ---- deferred backward: d P and d w_a over all timesteps at once (training layout: row = t*B + b, image b) ----
constexpr int DEF_COLS = 256;
__global__ void __launch_bounds__(DEF_COLS)
attention_bwd_deferred_kernel(int T, int B, int N, int A, int Ap, const float* __restrict__ q_all,
                              const float* __restrict__ du_all, const bf16* __restrict__ proj, const float* __restrict__ w_a,
                              float* __restrict__ dproj, float* __restrict__ dwa_rows) {
  extern __shared__ float dsm[];
  float* q_s = dsm;                         // T x DEF_COLS
  float* du_s = q_s + (size_t)T * DEF_COLS; // T x N
  int* act = reinterpret_cast<int*>(du_s + (size_t)T * N);   // T
  const int b = blockIdx.x;
  const int a0 = blockIdx.y * DEF_COLS;
  const int col = a0 + threadIdx.x;
  const bool ok = col < A;
  for (int t = threadIdx.x; t < T; t += blockDim.x) act[t] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < T * N; i += blockDim.x) {
    const int t = i / N, n = i - t * N;
    const float v = du_all[((size_t)t * B + b) * N + n];
    du_s[i] = v;
    if (v != 0.f) act[t] = 1;               // benign race: all writers store 1
  }
  for (int t = 0; t < T; ++t) q_s[t * DEF_COLS + threadIdx.x] = ok ? q_all[((size_t)t * B + b) * A + col] : 0.f;
  __syncthreads();
  const float wa = ok ? w_a[col] : 0.f;
  float dw = 0.f;
  for (int n = 0; n < N; ++n) {
    const float p = ok ? __bfloat162float(proj[((size_t)b * N + n) * Ap + col]) : 0.f;
    float acc = 0.f;
    for (int t = 0; t < T; ++t) {
      if (!act[t]) continue;                // padded timesteps carry no gradient
      const float d = du_s[t * N + n];
      const float th = tanh_approx(q_s[t * DEF_COLS + threadIdx.x] + p);
      acc += d * (1.f - th * th);
      dw += d * th;
    }
    if (ok) dproj[((size_t)b * N + n) * A + col] = wa * acc;
  }
  if (ok) dwa_rows[(size_t)b * A + col] = dw;
}

int attention_backward_deferred(cudaStream_t s, const AttnArgs& a, int T, const float* q_all, const float* du_all,
                                float* dproj, float* dwa_rows) {
  PROF_SCOPE(s, "attention_bwd_deferred", 0,
             (double)a.R * ((double)T * (a.A + a.N) * 4.0 + (double)a.N * a.Ap * 2.0 + (double)a.N * a.A * 4.0));
  REQUIRE(a.rowmap == nullptr && a.ld_q == a.A, "attention_backward_deferred: training layout only");
  const size_t smem = ((size_t)T * DEF_COLS + (size_t)T * a.N + T) * 4;
  REQUIRE(smem <= 200 * 1024, "attention_backward_deferred: T=%d too long", T);
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(attention_bwd_deferred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  dim3 grid(a.R, ceil_div(a.A, DEF_COLS));
  attention_bwd_deferred_kernel<<<grid, DEF_COLS, smem, s>>>(T, a.R, a.N, a.A, a.Ap, q_all, du_all, a.proj, a.w_a, dproj,
                                                            dwa_rows);
  LAUNCHED();
  return 0;
}

}  // namespace sscvae
