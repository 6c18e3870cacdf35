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

namespace sscvae {

#define LAUNCHED() do { CUDA_TRY(cudaGetLastError()); ++g_launch_count_pw; } while (0)

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

namespace {

constexpr int ATT_CONSUMERS = 256;                 // 8 consumer warps
constexpr int ATT_CWARPS = ATT_CONSUMERS / 32;
constexpr int ATT_THREADS = ATT_CONSUMERS + 32;    // + 1 producer warp
constexpr int ATT_STAGES = 4;
constexpr int ATT_STAGE_BYTES = 16384;             // two CTAs per SM: their phases overlap
constexpr int ATT_MAXB = 8;                        // boxes per chunk (<= consumer warps: one box per warp and chunk)
constexpr int ATT_NREG = 4;                        // boxes per lane in the softmax: N <= 128
constexpr int ATT_FV = 2;                          // 16-byte feature vectors per thread: Fp <= 4096
constexpr int ATT_PV = 2;                          // projection column pairs per thread: Ap <= 1024

struct AttnPlan {
  int nP, bP;            // chunks / boxes per chunk of the projected features
  int nF, bF;            // same for the region features
  int rows_per_cta;
};

struct AttnSmem {
  uint8_t* stage;        // ATT_STAGES x ATT_STAGE_BYTES
  uint64_t* full;        // [ATT_STAGES]
  uint64_t* empty;       // [ATT_STAGES]
  float* wa;             // Ap
  float* q0;             // 2 x Ap (double buffer, prefetched one row ahead)
  float* u;              // ATT_CWARPS x N4 partial scores (d alpha in backward): [sub-warp][box]
  float* alw;            // ATT_CWARPS x N4: every consumer warp's own copy of alpha (broadcast reads in the weighted sum)
  float* dx0;            // 2 x Fp (backward only)
  float* sv0;            // 2 x N4 (backward only: saved softmax)
  int Ap, Fp, N4;
  __device__ __forceinline__ float* q(int slot) const { return q0 + slot * Ap; }
  __device__ __forceinline__ float* dx(int slot) const { return dx0 + slot * Fp; }
  __device__ __forceinline__ float* sv(int slot) const { return sv0 + slot * N4; }
};

__device__ __forceinline__ AttnSmem carve(uint8_t* raw, const AttnArgs& a, bool bwd) {
  AttnSmem s;
  // no integer round trip on the pointer: it would lose the shared address space and turn every access below into a
  // generic LD/ST with 64-bit address arithmetic (a third of the instructions of the first version of these kernels)
  uint8_t* p = raw;
  s.stage = p; p += ATT_STAGES * ATT_STAGE_BYTES;
  s.full = reinterpret_cast<uint64_t*>(p); p += ATT_STAGES * 8;
  s.empty = reinterpret_cast<uint64_t*>(p); p += ATT_STAGES * 8;
  float* f = reinterpret_cast<float*>(p);
  s.Ap = a.Ap; s.Fp = a.Fp; s.N4 = (a.N + 3) & ~3;
  s.wa = f; f += a.Ap;
  s.q0 = f; f += 2 * a.Ap;
  s.u = f; f += ATT_CWARPS * s.N4;
  s.alw = f; f += ATT_CWARPS * s.N4;
  s.dx0 = s.sv0 = nullptr;
  if (bwd) {
    s.dx0 = f; f += 2 * a.Fp;
    s.sv0 = f; f += 2 * s.N4;
  }
  return s;
}
size_t attn_smem_bytes(const AttnArgs& a, bool bwd) {
  size_t n = 128 + (size_t)ATT_STAGES * ATT_STAGE_BYTES + 2 * ATT_STAGES * 8;
  n += (size_t)(3 * a.Ap + 2 * ATT_CWARPS * ((a.N + 3) & ~3)) * 4;
  if (bwd) n += (size_t)(2 * a.Fp + 2 * ((a.N + 3) & ~3)) * 4;
  return n;
}

// ring bookkeeping shared by producer and consumers (both walk the same chunk sequence)
struct Ring {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance() {
    if (++stage == ATT_STAGES) { stage = 0; phase ^= 1; }
  }
};

__device__ __forceinline__ void produce_block(const AttnSmem& sm, Ring& ring, const uint8_t* base, int N, int row_bytes,
                                              int nchunks, int bper, uint64_t policy) {
  for (int c = 0; c < nchunks; ++c) {
    const int n0 = c * bper;
    const int nb = min(bper, N - n0);
    const uint32_t bytes = (uint32_t)nb * (uint32_t)row_bytes;
    ptx::mbar_wait(&sm.empty[ring.stage], ring.phase ^ 1);
    ptx::mbar_expect_tx(&sm.full[ring.stage], bytes);
    // evict_first: the 52 MB of features + projections are read once per timestep and would otherwise push the
    // recurrent weights (76 MB, re-read every step) out of the 126 MB L2
    if (policy) ptx::bulk_g2s_hint(sm.stage + (size_t)ring.stage * ATT_STAGE_BYTES, base + (size_t)n0 * row_bytes, bytes,
                                   &sm.full[ring.stage], policy);
    else ptx::bulk_g2s(sm.stage + (size_t)ring.stage * ATT_STAGE_BYTES, base + (size_t)n0 * row_bytes, bytes, &sm.full[ring.stage]);
    ring.advance();
  }
}

// row-vector prefetch (cp.async, 4 bytes per op: no alignment requirement on the row stride)
__device__ __forceinline__ void prefetch_vec(float* dst, const float* src, int n) {
  for (int i = threadIdx.x; i < n; i += ATT_CONSUMERS) ptx::cp_async4(dst + i, src + i);
}

template <int K>
__device__ __forceinline__ float pick(const float (&v)[K], int k) {
  float r = v[0];
#pragma unroll
  for (int i = 1; i < K; ++i) r = (k == i) ? v[i] : r;
  return r;
}

// A chunk holds nb <= ATT_MAXB boxes; 8/pow2ceil(nb) warps share one box (each a slice of the vectors), so all
// consumer warps stay busy whatever the chunk size. Partial sums land in part[sub][n] and are added by the readers.
__device__ __forceinline__ int warps_per_box(int nb) { return nb > 4 ? 1 : nb > 2 ? 2 : nb > 1 ? 4 : 8; }
__device__ __forceinline__ float gather_partial(const float* part, int N4, int N, int bper, int n) {
  const int n0 = (n / bper) * bper;
  const int wpb = warps_per_box(min(bper, N - n0));
  float s = 0.f;
  for (int k = 0; k < wpb; ++k) s += part[k * N4 + n];
  return s;
}

// scores u_n of the boxes of one P chunk
__device__ __forceinline__ void chunk_scores(const AttnArgs& a, const bf16* buf, int n0, int nb, const float* mask_img,
                                             const float* q_s, const float* wa_s, float* part, int N4) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = a.Ap >> 3;
  const int wpb = warps_per_box(nb);
  const int j = warp / wpb, sub = warp % wpb;
  if (j >= nb) return;
  const int n = n0 + j;
  float s = 0.f;
  if (mask_img[n] != 0.f) {                        // masked boxes enter the softmax as u*m = 0
    const bf16x8* p = reinterpret_cast<const bf16x8*>(buf + (size_t)j * a.Ap);
    for (int i = sub * 32 + lane; i < nvec; i += 32 * wpb) {
      const bf16x8 v = p[i];
      const float4 qa = *reinterpret_cast<const float4*>(q_s + i * 8);
      const float4 qb = *reinterpret_cast<const float4*>(q_s + i * 8 + 4);
      const float4 wa = *reinterpret_cast<const float4*>(wa_s + i * 8);
      const float4 wb = *reinterpret_cast<const float4*>(wa_s + i * 8 + 4);
      const float2 f0 = __bfloat1622float2(v.v[0]), f1 = __bfloat1622float2(v.v[1]);
      const float2 f2 = __bfloat1622float2(v.v[2]), f3 = __bfloat1622float2(v.v[3]);
      s += wa.x * tanh_approx(qa.x + f0.x) + wa.y * tanh_approx(qa.y + f0.y);
      s += wa.z * tanh_approx(qa.z + f1.x) + wa.w * tanh_approx(qa.w + f1.y);
      s += wb.x * tanh_approx(qb.x + f2.x) + wb.y * tanh_approx(qb.y + f2.y);
      s += wb.z * tanh_approx(qb.z + f3.x) + wb.w * tanh_approx(qb.w + f3.y);
    }
    s = warp_sum(s);
  }
  if (lane == 0) part[sub * N4 + n] = s;
}

// masked softmax, redundantly per warp; lane holds boxes lane, lane+32, ...
// sft[k] = softmax(u*m)[n], al[k] = alpha[n]; returns R = sum_n sft*m + 1e-13 (allennlp masked_softmax)
__device__ __forceinline__ float warp_masked_softmax(int N, const float* mask_img, const float* part, int N4, int bper,
                                                     float (&m)[ATT_NREG], float (&sft)[ATT_NREG], float (&al)[ATT_NREG]) {
  const int lane = threadIdx.x & 31;
  float x[ATT_NREG];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < ATT_NREG; ++k) {
    const int n = lane + 32 * k;
    m[k] = (n < N) ? mask_img[n] : 0.f;
    x[k] = (n < N) ? gather_partial(part, N4, N, bper, n) * m[k] : -INFINITY;
    mx = fmaxf(mx, x[k]);
  }
  mx = warp_max(mx);
  float se = 0.f;
#pragma unroll
  for (int k = 0; k < ATT_NREG; ++k) {
    x[k] = (lane + 32 * k < N) ? __expf(x[k] - mx) : 0.f;
    se += x[k];
  }
  se = warp_sum(se);
  float sr = 0.f;
#pragma unroll
  for (int k = 0; k < ATT_NREG; ++k) {
    sft[k] = x[k] / se;
    sr += sft[k] * m[k];
  }
  sr = warp_sum(sr);
  const float Rn = sr + 1e-13f;
#pragma unroll
  for (int k = 0; k < ATT_NREG; ++k) al[k] = sft[k] * m[k] / Rn;
  return Rn;
}

__device__ __forceinline__ void attn_prologue(const AttnSmem& sm, const AttnArgs& a, bool bwd) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < ATT_STAGES; ++s) {
      ptx::mbar_init(&sm.full[s], 1);
      ptx::mbar_init(&sm.empty[s], ATT_CWARPS);
    }
    ptx::mbar_fence_init();
  }
  for (int i = threadIdx.x; i < a.Ap; i += blockDim.x) {
    sm.wa[i] = (i < a.A) ? a.w_a[i] : 0.f;
    sm.q0[i] = 0.f;                                 // padding columns [A, Ap) stay zero: cp.async never writes them
    sm.q0[a.Ap + i] = 0.f;
  }
  if (bwd)
    for (int i = threadIdx.x; i < 2 * a.Fp; i += blockDim.x) sm.dx0[i] = 0.f;
  __syncthreads();
}

}  // namespace

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
  const int nfv = a.Fp >> 3;
  if (r_begin < r_end) {
    prefetch_vec(sm.q(0), a.q + (size_t)r_begin * a.ld_q, a.A);
    ptx::cp_async_commit();
  }
  int cur = 0;
  for (int r = r_begin; r < r_end; ++r, cur ^= 1) {
    const int img = a.rowmap ? a.rowmap[r] : r;
    const float* mask_img = a.mask + (size_t)img * a.N;
    ptx::cp_async_wait_all();
    ptx::bar_sync(1, ATT_CONSUMERS);                 // q[cur] landed; everybody is done with the previous row
    if (r + 1 < r_end) {
      prefetch_vec(sm.q(cur ^ 1), a.q + (size_t)(r + 1) * a.ld_q, a.A);
      ptx::cp_async_commit();
    }
    for (int c = 0; c < pl.nP; ++c) {
      const int n0 = c * pl.bP, nb = min(pl.bP, a.N - n0);
      ptx::mbar_wait(&sm.full[ring.stage], ring.phase);
      chunk_scores(a, reinterpret_cast<const bf16*>(sm.stage + (size_t)ring.stage * ATT_STAGE_BYTES), n0, nb, mask_img,
                   sm.q(cur), sm.wa, sm.u, sm.N4);
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm.empty[ring.stage]);
      ring.advance();
    }
    ptx::bar_sync(1, ATT_CONSUMERS);                 // all N scores are in shared memory
    float m[ATT_NREG], sft[ATT_NREG], al[ATT_NREG];
    warp_masked_softmax(a.N, mask_img, sm.u, sm.N4, pl.bP, m, sft, al);
    if (warp == 0) {
#pragma unroll
      for (int k = 0; k < ATT_NREG; ++k) {
        const int n = lane + 32 * k;
        if (n < a.N) {
          alpha[(size_t)r * a.N + n] = al[k];
          if (smx) smx[(size_t)r * a.N + n] = sft[k];
        }
      }
    }
    // weighted sum: a thread owns 8 consecutive features (one 16-byte vector per box row). The inner loop is the
    // instruction hot spot of the kernel (ncu: 39 % of all issued instructions, 46 per 8 FMAs when alpha came out of
    // registers through a select chain + shuffle): alpha is read as a shared-memory broadcast from the warp's own
    // copy, bf16 -> fp32 is one shift / one mask per element, and the box loop is unrolled.
    {
      float* mine = sm.alw + warp * sm.N4;
#pragma unroll
      for (int k = 0; k < ATT_NREG; ++k)
        if (lane + 32 * k < a.N) mine[lane + 32 * k] = al[k];
      __syncwarp();
    }
    const float* alw = sm.alw + warp * sm.N4;
    const bool two = nfv > ATT_CONSUMERS;            // uniform: a second vector per thread only when Fp > 2048
    float acc[ATT_FV][8];
#pragma unroll
    for (int v = 0; v < ATT_FV; ++v)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[v][k] = 0.f;
    for (int c = 0; c < pl.nF; ++c) {
      const int n0 = c * pl.bF, nb = min(pl.bF, a.N - n0);
      ptx::mbar_wait(&sm.full[ring.stage], ring.phase);
      const uint4* buf = reinterpret_cast<const uint4*>(sm.stage + (size_t)ring.stage * ATT_STAGE_BYTES);
      if (threadIdx.x < nfv) {
#pragma unroll 4
        for (int j = 0; j < nb; ++j) {
          const float w = alw[n0 + j];
          const uint4 x = buf[(size_t)j * nfv + threadIdx.x];
          acc[0][0] = fmaf(w, __uint_as_float(x.x << 16), acc[0][0]);
          acc[0][1] = fmaf(w, __uint_as_float(x.x & 0xffff0000u), acc[0][1]);
          acc[0][2] = fmaf(w, __uint_as_float(x.y << 16), acc[0][2]);
          acc[0][3] = fmaf(w, __uint_as_float(x.y & 0xffff0000u), acc[0][3]);
          acc[0][4] = fmaf(w, __uint_as_float(x.z << 16), acc[0][4]);
          acc[0][5] = fmaf(w, __uint_as_float(x.z & 0xffff0000u), acc[0][5]);
          acc[0][6] = fmaf(w, __uint_as_float(x.w << 16), acc[0][6]);
          acc[0][7] = fmaf(w, __uint_as_float(x.w & 0xffff0000u), acc[0][7]);
        }
      }
      if (two && threadIdx.x + ATT_CONSUMERS < nfv) {
        for (int j = 0; j < nb; ++j) {
          const float w = alw[n0 + j];
          const uint4 x = buf[(size_t)j * nfv + threadIdx.x + ATT_CONSUMERS];
          acc[1][0] = fmaf(w, __uint_as_float(x.x << 16), acc[1][0]);
          acc[1][1] = fmaf(w, __uint_as_float(x.x & 0xffff0000u), acc[1][1]);
          acc[1][2] = fmaf(w, __uint_as_float(x.y << 16), acc[1][2]);
          acc[1][3] = fmaf(w, __uint_as_float(x.y & 0xffff0000u), acc[1][3]);
          acc[1][4] = fmaf(w, __uint_as_float(x.z << 16), acc[1][4]);
          acc[1][5] = fmaf(w, __uint_as_float(x.z & 0xffff0000u), acc[1][5]);
          acc[1][6] = fmaf(w, __uint_as_float(x.w << 16), acc[1][6]);
          acc[1][7] = fmaf(w, __uint_as_float(x.w & 0xffff0000u), acc[1][7]);
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm.empty[ring.stage]);
      ring.advance();
    }
#pragma unroll
    for (int v = 0; v < ATT_FV; ++v) {
      const int vec = threadIdx.x + ATT_CONSUMERS * v;
      if (vec < nfv) {
        bf16x8 o;
#pragma unroll
        for (int k = 0; k < 4; ++k) o.v[k] = __floats2bfloat162_rn(acc[v][2 * k], acc[v][2 * k + 1]);
        st_bf16x8(xhat + (size_t)r * ld_x + vec * 8, o);
      }
    }
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
