// tcgen05 / TMEM / TMA GEMM for sm_100a (see gemm.cuh for the contract).
//
// One CTA computes one 128 x BN output tile:
//   warp 0      : TMA producer  (cp.async.bulk.tensor.2d, 128B-swizzled K-major tiles, mbarrier ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, bf16 in, fp32 accum in TMEM)
//   warps 2..5  : epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) -> registers ->
//                 fused bias / addends / tanh / dtanh -> fp32 and/or bf16 global stores
// The accumulator never touches registers or shared memory until the epilogue.
#include "gemm.cuh"
#include "kernels.cuh"
#include "prof.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <algorithm>

namespace sscvae {

unsigned long long g_launch_count = 0;

static constexpr int BM = 128;
static constexpr int BK = 64;                    // 64 bf16 = 128 bytes = one swizzle-128B row
static constexpr int A_STAGE_BYTES = BM * BK * 2;
static constexpr int GEMM_THREADS = 192;
static constexpr int MAX_SEG = 3;

struct GemmParams {
  CUtensorMap ta[MAX_SEG];
  CUtensorMap tb[MAX_SEG];
  int kblocks[MAX_SEG];
  int nseg;
  int M, N;
  int vec;                                        // 1: every epilogue pointer allows 16-byte access
  int splits;                                     // gridDim.z
  int kb_per_split;                               // split-K: CTA z works on k-blocks [z*kb_per_split, ...) of the
                                                  // concatenated segments and ADDS its partial tile into C32
  int tma_store;                                  // 1: plain fp32 tile (+bias): the epilogue leaves through TMA stores of `tc`
  CUtensorMap tc;                                 // fp32 output, box 32 columns x 32 rows, no swizzle
  GemmEpi epi;
  int w_policy;                                   // pair kernel: 1 = weight tiles are loaded with an L2 evict_first hint
  int fuse_lstm;                                  // pair kernel: run the LSTM cell `lstm` in the epilogue
  int lstm_tma;                                   // ... and its outputs leave through TMA stores of `tl`
  long long* dbg;                                 // SSCVAE_GEMM_DBG=1: per-CTA clock64 stamps of the pair kernel's phases
  LstmFwdArgs lstm;
  CUtensorMap tl[7];                              // gates i,f,g,o (fp32), c (fp32), h -> h1_dst, h2_dst (bf16); box 32 units x 32 rows
  RowStatsEpi rs;                                 // vocabulary head: softmax statistics / CE gradient in the epilogue
};

// ---------------------------------------------------------------------------------------------
// epilogue math for one run of 32 consecutive columns of one row
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tanh_fast(float x) {
  // accurate enough for fp32 parity: tanhf via exp; the MUFU.TANH approximation is reserved for the
  // attention kernel's 27k tanh/row where the operands are bf16 anyway.
  return tanhf(x);
}

// explicit shared-memory accesses (32-bit shared addresses): pointers derived from the aligned dynamic-smem base are
// treated as generic by the compiler, and generic LD/ST that resolve to shared memory are several times slower than
// LDS/STS for a lone epilogue warp
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t a, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds_v4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}

__device__ __forceinline__ void red_add_f32x4(float* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Second half of the vector epilogue: the warp's 32x32 fp32 block sits in `stage` as [row][column] (pitch 36 floats);
// 8 consecutive lanes own 32 consecutive columns of ONE output row (128 contiguous bytes), 4 rows per instruction,
// for the addend loads and for the stores.
// NOTE: GemmEpi is taken BY VALUE everywhere in the epilogue. Through a reference to the __grid_constant__ kernel
// parameter the compiler must assume the output stores alias the struct and re-reads every field after each store
// (measured: ~250 cycles per store, 2-3.4k cycles per 32x32 block instead of ~300).
__device__ __forceinline__ void epilogue_block_vec(const GemmEpi e, int M, int row0, int col0, uint32_t stage,
                                                   int splitk, bool add1_all) {
  constexpr int PITCH = 36;
  const int lane = threadIdx.x & 31;
  const int c4 = (lane & 7) * 4;
  const int n = col0 + c4;
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e.bias) bias4 = *reinterpret_cast<const float4*>(e.bias + n);
  // The epilogue warps run one per SM sub-partition, so the epilogue is bound by its own instruction count
  // (measured: 3.4k cycles per 32x32 block when the tanh / dtanh / accumulate forms were compiled into the one loop).
  // The common case - scaled accumulator + bias + addends -> fp32 and/or bf16 - gets its own lean loop.
  if (!e.act && !e.dtanh && !e.accumulate && !splitk) {
    const float alpha = e.alpha;
    const float* a1 = add1_all ? e.add1 + n : nullptr;
    const float* a2 = e.add2 ? e.add2 + n : nullptr;
    float* c32 = e.C32 ? e.C32 + n : nullptr;
    bf16* c16 = e.C16 ? e.C16 + n : nullptr;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int rr = k * 4 + (lane >> 3);
      const int r = row0 + rr;
      if (r < M) {
        float4 v = lds_v4(stage + (rr * PITCH + c4) * 4);
        v.x = fmaf(v.x, alpha, bias4.x); v.y = fmaf(v.y, alpha, bias4.y);
        v.z = fmaf(v.z, alpha, bias4.z); v.w = fmaf(v.w, alpha, bias4.w);
        if (a1) {
          const float4 b = *reinterpret_cast<const float4*>(a1 + (size_t)r * e.ld1);
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (a2) {
          const float4 b = *reinterpret_cast<const float4*>(a2 + (size_t)r * e.ld2);
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (c32) *reinterpret_cast<float4*>(c32 + (size_t)r * e.ldc32) = v;
        if (c16) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
          uint2 o;
          o.x = *reinterpret_cast<uint32_t*>(&lo);
          o.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(c16 + (size_t)r * e.ldc16) = o;
        }
      }
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int rr = k * 4 + (lane >> 3);
    const int r = row0 + rr;
    if (r < M) {
      float4 v = lds_v4(stage + (rr * PITCH + c4) * 4);
      v.x = v.x * e.alpha + bias4.x; v.y = v.y * e.alpha + bias4.y;
      v.z = v.z * e.alpha + bias4.z; v.w = v.w * e.alpha + bias4.w;
      if (add1_all) {
        const float4 b = *reinterpret_cast<const float4*>(e.add1 + (size_t)r * e.ld1 + n);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      }
      if (e.add2) {
        const float4 b = *reinterpret_cast<const float4*>(e.add2 + (size_t)r * e.ld2 + n);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      }
      if (e.act == 1) { v.x = tanh_fast(v.x); v.y = tanh_fast(v.y); v.z = tanh_fast(v.z); v.w = tanh_fast(v.w); }
      if (e.dtanh) {
        const float4 t = *reinterpret_cast<const float4*>(e.dtanh + (size_t)r * e.ldd + n);
        v.x *= 1.0f - t.x * t.x; v.y *= 1.0f - t.y * t.y; v.z *= 1.0f - t.z * t.z; v.w *= 1.0f - t.w * t.w;
      }
      if (e.C32) {
        float* p = e.C32 + (size_t)r * e.ldc32 + n;
        if (splitk) {
          red_add_f32x4(p, v);
        } else {
          if (e.accumulate) {
            const float4 c = *reinterpret_cast<const float4*>(p);
            v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
          }
          *reinterpret_cast<float4*>(p) = v;
        }
      }
      if (e.C16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&lo);
        o.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(e.C16 + (size_t)r * e.ldc16 + n) = o;
      }
    }
  }
}

// Plain fp32 tile (+ bias): the warp's 32x32 block goes to shared memory as 32 rows of 128 bytes in the TMA
// 128B-swizzle pattern (16-byte chunk c of row r at chunk c ^ (r & 7): 4-way instead of 32-way bank conflicts for
// lane == row) and leaves with ONE asynchronous TMA store; rows / columns beyond the matrix are clipped by the TMA
// unit. Plain STG.128 stores cost a lone epilogue warp ~220 cycles each (1.8k cycles per block, measured).
__device__ __forceinline__ void epilogue_tma_row32(const GemmEpi e, const CUtensorMap* tc, int N, int row0, int col0,
                                                   const uint32_t (&acc)[32], uint32_t stage) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) tma_store_wait_read();                  // the previous store has finished reading this buffer
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e.bias && col0 + 4 * j + 3 < N) b = *reinterpret_cast<const float4*>(e.bias + col0 + 4 * j);
    else if (e.bias) {
      if (col0 + 4 * j < N) b.x = e.bias[col0 + 4 * j];
      if (col0 + 4 * j + 1 < N) b.y = e.bias[col0 + 4 * j + 1];
      if (col0 + 4 * j + 2 < N) b.z = e.bias[col0 + 4 * j + 2];
    }
    const float4 v = make_float4(fmaf(__uint_as_float(acc[4 * j]), e.alpha, b.x), fmaf(__uint_as_float(acc[4 * j + 1]), e.alpha, b.y),
                                 fmaf(__uint_as_float(acc[4 * j + 2]), e.alpha, b.z), fmaf(__uint_as_float(acc[4 * j + 3]), e.alpha, b.w));
    sts_v4(stage + lane * 128 + ((j ^ (lane & 7)) << 4), v);
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) tma_store_2d(tc, stage, col0, row0);
}

// splitk: 0 = plain store; 1 = this CTA holds a partial sum: add it into C32 (split 0 also applies the addends)
__device__ __forceinline__ void epilogue_row32(GemmEpi e, int vec, int M, int N, int row0, int col0, int ncols,
                                               const uint32_t (&acc)[32], uint32_t stage, int splitk, int first_split) {
  const int lane = threadIdx.x & 31;
  if (splitk && !first_split) { e.bias = nullptr; e.add1 = nullptr; e.add2 = nullptr; }
  const int row = row0 + lane;           // TMEM lane == output row of this thread's 32 accumulator columns
  // `ncols` < 32 only for the narrow (BN=16) tiles; add1 may cover just the first add1_cols columns
  const bool add1_all = e.add1 && (col0 + 32 <= e.add1_cols);
  const bool add1_none = !e.add1 || (col0 >= e.add1_cols);
  const bool full = (ncols == 32) && (col0 + 32 <= N) && (add1_all || add1_none);
  if (vec && full) {
    // Transpose the warp's 32x32 fp32 block through shared memory (the pipeline's stage buffers are free by
    // now) so that global traffic is coalesced: afterwards 8 consecutive lanes own 32 consecutive columns of
    // ONE row (128 contiguous bytes), 4 rows per instruction, for the addend loads and for the stores.
    constexpr int PITCH = 36;            // floats; keeps float4 alignment
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts_v4(stage + (lane * PITCH + 4 * j) * 4,
             make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]), __uint_as_float(acc[4 * j + 2]),
                         __uint_as_float(acc[4 * j + 3])));
    __syncwarp();
    epilogue_block_vec(e, M, row0, col0, stage, splitk, add1_all);
    __syncwarp();
    return;
  }
  if (row >= M) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) * e.alpha;
  // scalar path (ragged right edge or unaligned pointers, e.g. gradients written straight into
  // a column block of a reference-layout weight matrix)
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int n = col0 + j;
    if (n < N && j < ncols) {
      float x = v[j];
      if (e.bias) x += e.bias[n];
      if (e.add1 && n < e.add1_cols) x += e.add1[(size_t)row * e.ld1 + n];
      if (e.add2) x += e.add2[(size_t)row * e.ld2 + n];
      if (e.act == 1) x = tanh_fast(x);
      if (e.dtanh) { float t = e.dtanh[(size_t)row * e.ldd + n]; x *= 1.0f - t * t; }
      if (e.C32) {
        float* p = e.C32 + (size_t)row * e.ldc32 + n;
        if (splitk) atomicAdd(p, x);
        else *p = e.accumulate ? (*p + x) : x;
      }
      if (e.C16) e.C16[(size_t)row * e.ldc16 + n] = __float2bfloat16_rn(x);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr uint32_t IDESC = make_idesc(BM, BN);
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;          // allocation granularity: power of two >= 32
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;                             // __align__(1024): the 128B swizzle needs 1024-byte aligned tiles
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;

  int total_kb = 0;
#pragma unroll
  for (int s = 0; s < MAX_SEG; ++s) total_kb += (s < p.nseg) ? p.kblocks[s] : 0;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(total_kb, kb_begin + p.kb_per_split);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nseg; ++s) { prefetch_tmap(&p.ta[s]); prefetch_tmap(&p.tb[s]); }
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  pdl_wait();                                           // everything above overlaps the previous kernel's tail
  pdl_launch_dependents(1);

  if (warp == 0) {
    if (elect_one_sync()) {
      int stage = 0; uint32_t phase = 0;
      int s = 0, base = 0;                              // segment of the current k-block, its first global index
      for (int g = kb_begin; g < kb_end; ++g) {
        while (g >= base + p.kblocks[s]) { base += p.kblocks[s]; ++s; }
        const int kb = g - base;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES + B_STAGE_BYTES);
        tma_load_2d(smem_a + stage * A_STAGE_BYTES, &p.ta[s], &full_bar[stage], kb * BK, m0);
        tma_load_2d(smem_b + stage * B_STAGE_BYTES, &p.tb[s], &full_bar[stage], kb * BK, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem_a + stage * A_STAGE_BYTES);
        const uint32_t b_base = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 elements (32 bytes) along K inside the 128B swizzle atom
          umma_bf16(tmem_acc, make_smem_desc(a_base + k * 32), make_smem_desc(b_base + k * 32), IDESC,
                    (kb > kb_begin || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);          // frees the smem slot once these MMAs have read it
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(accum_bar);                    // accumulator complete -> epilogue
    }
    __syncwarp();
  } else {
    // epilogue warps 2..5; a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int q = warp & 3;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int row = m0 + q * 32 + lane;
    // vocabulary head (RowStatsEpi): running softmax statistics of this thread's row over the tile's columns
    const RowStatsEpi rs = p.rs;
    float rs_m = -INFINITY, rs_s = 0.f, rs_t = 0.f; int rs_a = 0x7fffffff;
    int rs_target = -1; float rs_lse = 0.f, rs_g = 0.f;
    if (rs.mode && row < p.M) {
      if (rs.target) rs_target = rs.target[row];
      if (rs.mode == 2) { rs_lse = rs.lse[row]; rs_g = rs.gcoef[row]; }
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= p.N) break;                 // warp-uniform
      uint32_t acc[32];
      tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + c0, acc);
      tmem_ld_wait();
      if (rs.mode) {
        // One warp per SM sub-partition runs this with nothing to hide latency behind: everything is written as
        // independent operations + tree reductions (a running  s += exp(..)  chain cost ~120 cycles per element).
        const int nb = n0 + c0;
        const bool full = nb + 32 <= p.N;                     // warp-uniform: only the last tile has a ragged edge
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) * p.epi.alpha;
        if (p.epi.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || nb + j < p.N) v[j] += p.epi.bias[nb + j];
        }
        if (!full) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j >= p.N) v[j] = -INFINITY;              // exp -> 0, never the max
        }
        if (rs.mode == 1) {
          float t[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) t[j] = fmaxf(v[j], v[j + 16]);
#pragma unroll
          for (int w = 8; w > 0; w >>= 1)
#pragma unroll
            for (int j = 0; j < w; ++j) t[j] = fmaxf(t[j], t[j + w]);
          const float cm = t[0];
          if (cm > rs_m) {                       // strict: an earlier column / tile keeps the arg max on ties
#pragma unroll
            for (int j = 31; j >= 0; --j)
              if (v[j] == cm) rs_a = nb + j;
            rs_s *= exp2f((rs_m - cm) * 1.4426950408889634f);
            rs_m = cm;
          }
          const float mb = rs_m * 1.4426950408889634f;
          float e[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) e[j] = exp2f(fmaf(v[j], 1.4426950408889634f, -mb));
#pragma unroll
          for (int w = 16; w > 0; w >>= 1)
#pragma unroll
            for (int j = 0; j < w; ++j) e[j] += e[j + w];
          rs_s += e[0];
          if (rs_target >= nb && rs_target < nb + 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j == rs_target) rs_t = v[j];
          }
          if (!p.epi.C32 && !p.epi.C16) continue;      // statistics only: the logits are never written
        } else {
          const float lb = rs_lse * 1.4426950408889634f;
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(rs_g * exp2f(fmaf(v[j], 1.4426950408889634f, -lb)));
          if (rs_target >= nb && rs_target < nb + 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j == rs_target) acc[j] = __float_as_uint(__uint_as_float(acc[j]) - rs_g);
          }
          GemmEpi e2 = p.epi;
          e2.alpha = 1.f; e2.bias = nullptr;
          epilogue_row32(e2, p.vec, p.M, p.N, m0 + q * 32, n0 + c0, BN - c0 < 32 ? BN - c0 : 32, acc,
                         smem_u32(smem_a) + (warp - 2) * 32 * 36 * 4, 0, 1);
          continue;
        }
      }
      if (BN >= 32 && p.tma_store) {
        epilogue_tma_row32(p.epi, &p.tc, p.N, m0 + q * 32, n0 + c0, acc, smem_u32(smem_a) + (warp - 2) * 4096);
      } else {
        epilogue_row32(p.epi, p.vec, p.M, p.N, m0 + q * 32, n0 + c0, BN - c0 < 32 ? BN - c0 : 32, acc,
                       smem_u32(smem_a) + (warp - 2) * 32 * 36 * 4, gridDim.z > 1, blockIdx.z == 0);
      }
    }
    if (rs.mode == 1 && row < p.M) {             // partial statistics of (row, this 128-column tile): lanes = consecutive rows
      const size_t o = (size_t)blockIdx.x * p.M + row;
      rs.st_max[o] = rs_m; rs.st_sum[o] = rs_s; rs.st_arg[o] = rs_a;
      if (rs.tgt_logit && rs_target >= n0 && rs_target < n0 + BN) rs.tgt_logit[row] = rs_t;
    }
  }
  if (p.tma_store && warp >= 2 && (threadIdx.x & 31) == 0) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_acc);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair kernel: a 2-CTA cluster computes one 256 x BN tile with tcgen05.mma.cta_group::2 (M = 256).
// CTA r of the pair owns rows [128r, 128r+128) of the tile: it loads its own 128 x 64 slice of A and HALF of the
// B tile (BN/2 rows), so per MMA each SM reads half the shared-memory bytes of the single-CTA kernel and the pair
// pulls B through L2 once instead of twice. Only the leader (rank 0) issues MMAs; both CTAs' TMA loads complete on
// the leader's `full` barrier; tcgen05.commit multicasts the `empty` / accumulator-ready arrivals to both CTAs.
// ---------------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ GemmParams p) {
  constexpr int BH = BN / 2;                              // B rows loaded by each CTA
  constexpr int B_STAGE_BYTES = BH * BK * 2;
  constexpr uint32_t IDESC = make_idesc(2 * BM, BN);
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;                             // __align__(1024): the 128B swizzle needs 1024-byte aligned tiles
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int m0 = blockIdx.y * (2 * BM) + (int)rank * BM;
  const int n0 = (blockIdx.x >> 1) * BN;

  int total_kb = 0;
#pragma unroll
  for (int s = 0; s < MAX_SEG; ++s) total_kb += (s < p.nseg) ? p.kblocks[s] : 0;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(total_kb, kb_begin + p.kb_per_split);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nseg; ++s) { prefetch_tmap(&p.ta[s]); prefetch_tmap(&p.tb[s]); }
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc_2sm<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();                                     // both CTAs' barriers exist before any remote arrival
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents(1);

  if (warp == 0) {
    if (elect_one_sync()) {
      int stage = 0; uint32_t phase = 0;
      int s = 0, base = 0;
      for (int g = kb_begin; g < kb_end; ++g) {
        while (g >= base + p.kblocks[s]) { base += p.kblocks[s]; ++s; }
        const int kb = g - base;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (A_STAGE_BYTES + B_STAGE_BYTES));
        tma_load_2d_2sm(smem_a + stage * A_STAGE_BYTES, &p.ta[s], &full_bar[stage], kb * BK, m0);
        tma_load_2d_2sm(smem_b + stage * B_STAGE_BYTES, &p.tb[s], &full_bar[stage], kb * BK, n0 + (int)rank * BH);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one_sync()) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem_a + stage * A_STAGE_BYTES);
        const uint32_t b_base = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          umma_bf16_2sm(tmem_acc, make_smem_desc(a_base + k * 32), make_smem_desc(b_base + k * 32), IDESC,
                        (kb > kb_begin || k > 0) ? 1u : 0u);
        }
        umma_commit_2sm(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit_2sm(accum_bar);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= p.N) break;
      uint32_t acc[32];
      tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + c0, acc);
      tmem_ld_wait();
      if (BN >= 32 && p.tma_store) {
        epilogue_tma_row32(p.epi, &p.tc, p.N, m0 + q * 32, n0 + c0, acc, smem_u32(smem_a) + (warp - 2) * 4096);
      } else {
        epilogue_row32(p.epi, p.vec, p.M, p.N, m0 + q * 32, n0 + c0, BN - c0 < 32 ? BN - c0 : 32, acc,
                       smem_u32(smem_a) + (warp - 2) * 32 * 36 * 4, gridDim.z > 1, blockIdx.z == 0);
      }
    }
  }
  if (p.tma_store && warp >= 2 && (threadIdx.x & 31) == 0) tma_store_wait_all();
  tc_fence_before();
  cluster_sync_all();                                     // the peer may still be reading its half of the accumulator
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<TMEM_COLS>(tmem_acc);
  }
}

// ---------------------------------------------------------------------------------------------
// Skinny-M kernel (M <= 256: the per-timestep recurrent GEMMs, M = batch): operands SWAPPED, K split over a
// thread-block cluster, partial tiles reduced through distributed shared memory.
//
// Measured on B200 (tools/gemm_bench.py, profiles/README.md): an M=128 tcgen05.mma takes ~130 cycles whatever its
// N, and one SM ingests TMA operands at ~77 GB/s at most. A skinny GEMM therefore wants (a) the <= 256 batch rows on
// the MMA's N side so every instruction is a full 128x256x16, and (b) as few operand bytes per SM as possible.
// A CTA owns 128 rows of the WEIGHT matrix (UMMA M) x all batch rows (UMMA N = 256) over 1/S of K: 48 KB of
// operands per 128x256x64 k-block (the 128x64 tiling moves 24 KB per 128x64x64). The S CTAs of a cluster (1,1,S)
// work on the same tile; afterwards CTA r owns the batch columns [r*256/S, (r+1)*256/S): every CTA pushes the other
// CTAs' column slices of its partial accumulator into their shared memory (st.shared::cluster, coalesced), a
// cluster barrier makes them visible, and the owner adds them to its own TMEM slice and runs the full epilogue.
// The accumulator is transposed (TMEM lane = output column n, TMEM column = batch row), so for a fixed batch row the
// 32 lanes of an epilogue warp hold 32 consecutive n: every global access is one coalesced 128-byte line.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_shared_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// v[j] = value of output column n (this lane) at batch row b0 + j. The block is staged in shared memory as
// [batch row][column] (lanes write consecutive words: conflict-free, no transpose needed) and leaves through the
// same vector path as the other kernels, 16 bytes per lane. (With 4-byte stores from the 4 epilogue warps alone the
// SM sustained ~5 B/clk: 3k cycles per 32x32 block.) Ragged / unaligned blocks take a scalar loop over the staged
// block; nothing indexes the register array dynamically (that would push it to local memory).
__device__ __forceinline__ void epilogue_swapped32(const GemmEpi e, int vec, int M, int N, int n_warp0, int b0,
                                                   const float (&v)[32], uint32_t stage) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 32; ++j) sts_f32(stage + (j * 36 + lane) * 4, v[j]);
  __syncwarp();
  const bool add1_all = e.add1 && (n_warp0 + 32 <= e.add1_cols);
  const bool add1_none = !e.add1 || (n_warp0 >= e.add1_cols);
  if (vec && n_warp0 + 32 <= N && (add1_all || add1_none)) {
    epilogue_block_vec(e, M, b0, n_warp0, stage, 0, add1_all);
  } else {
    const int n = n_warp0 + lane;
    if (n < N) {
      const float bias = e.bias ? e.bias[n] : 0.f;
      const bool add1 = e.add1 && n < e.add1_cols;
      const int rows = min(32, M - b0);
#pragma unroll 1
      for (int j = 0; j < rows; ++j) {
        const int b = b0 + j;
        float x = lds_f32(stage + (j * 36 + lane) * 4) * e.alpha + bias;
        if (add1) x += e.add1[(size_t)b * e.ld1 + n];
        if (e.add2) x += e.add2[(size_t)b * e.ld2 + n];
        if (e.act == 1) x = tanh_fast(x);
        if (e.dtanh) { const float t = e.dtanh[(size_t)b * e.ldd + n]; x *= 1.0f - t * t; }
        if (e.C32) {
          float* q = e.C32 + (size_t)b * e.ldc32 + n;
          *q = e.accumulate ? (*q + x) : x;
        }
        if (e.C16) e.C16[(size_t)b * e.ldc16 + n] = __float2bfloat16_rn(x);
      }
    }
  }
  __syncwarp();
}

template <int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tcgen05_swapped_kernel(const __grid_constant__ GemmParams p) {
  constexpr int BW = 128, BX = 256;                     // weight rows per CTA (UMMA M), activation rows (UMMA N)
  constexpr int W_STAGE_BYTES = BW * BK * 2, X_STAGE_BYTES = BX * BK * 2;
  constexpr uint32_t IDESC = make_idesc(BW, BX);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;                             // __align__(1024): the 128B swizzle needs 1024-byte aligned tiles
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  uint8_t* smem_w = smem;
  uint8_t* smem_x = smem + STAGES * W_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_x + STAGES * X_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BW;
  const int S = p.splits;                               // cluster size along z; this CTA's rank = blockIdx.z
  const int rank = blockIdx.z;

  int total_kb = 0;
#pragma unroll
  for (int s = 0; s < MAX_SEG; ++s) total_kb += (s < p.nseg) ? p.kblocks[s] : 0;
  const int kb_begin = rank * p.kb_per_split;
  const int kb_end = min(total_kb, kb_begin + p.kb_per_split);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nseg; ++s) { prefetch_tmap(&p.ta[s]); prefetch_tmap(&p.tb[s]); }
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<BX>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents(1);

  if (warp == 0) {
    if (elect_one_sync()) {
      int stage = 0; uint32_t phase = 0;
      int s = 0, base = 0;
      for (int g = kb_begin; g < kb_end; ++g) {
        while (g >= base + p.kblocks[s]) { base += p.kblocks[s]; ++s; }
        const int kb = g - base;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], W_STAGE_BYTES + X_STAGE_BYTES);
        tma_load_2d(smem_w + stage * W_STAGE_BYTES, &p.tb[s], &full_bar[stage], kb * BK, n0);
        tma_load_2d(smem_x + stage * X_STAGE_BYTES, &p.ta[s], &full_bar[stage], kb * BK, 0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one_sync()) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t w_base = smem_u32(smem_w + stage * W_STAGE_BYTES);
        const uint32_t x_base = smem_u32(smem_x + stage * X_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(tmem_acc, make_smem_desc(w_base + k * 32), make_smem_desc(x_base + k * 32), IDESC,
                    (kb > kb_begin || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  } else {
    mbar_wait(accum_bar, 0);                            // this CTA's partial tile is complete in TMEM
    tc_fence_after();
  }

  const int q = warp & 3;                               // epilogue warps 2..5 own TMEM lanes [32q, 32q+32)
  const int tl = q * 32 + lane;                         // TMEM lane = column n0 + tl of the output
  const int CW = BX / S;                                // batch columns owned by each CTA of the cluster
  float* recv = reinterpret_cast<float*>(smem);         // [(S-1) sources][CW columns][128 lanes], over the dead stages
  if (S > 1) {
    // every CTA of the cluster has finished its main loop (its stage buffers are dead) before anyone pushes
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp >= 2 && kb_end > kb_begin) {
      for (int o = 0; o < S; ++o) {
        if (o == rank) continue;
        const int slot = rank < o ? rank : rank - 1;
        const uint32_t rbase = mapa_shared(smem_u32(recv), (uint32_t)o);
        for (int c = 0; c < CW; c += 32) {
          if (o * CW + c >= p.M) break;                 // columns beyond the batch are never read
          uint32_t acc[32];
          tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + o * CW + c, acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            st_shared_cluster_f32(rbase + (uint32_t)(((slot * CW + c + j) * 128 + tl) * 4), __uint_as_float(acc[j]));
        }
      }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    const GemmEpi epi = p.epi;                          // registers (see the note at epilogue_block_vec)
    // which of the other CTAs actually had k-blocks (a trailing split can be empty when K is short)
    for (int c = 0; c < CW; c += 32) {
      const int b0 = rank * CW + c;
      if (b0 >= p.M) break;
      float v[32];
      if (kb_end > kb_begin) {
        uint32_t acc[32];
        tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + b0, acc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      for (int o = 0; o < S; ++o) {
        if (o == rank || o * p.kb_per_split >= total_kb) continue;
        const int slot = o < rank ? o : o - 1;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += lds_f32(smem_u32(recv) + (uint32_t)(((slot * CW + c + j) * 128 + tl) * 4));
      }
      if (p.tma_store) {
        // dense [32 batch rows][32 columns] block per warp -> one TMA store; the async engine writes it out while the
        // warp goes on (plain STG.128 stores cost this lone warp ~220 cycles each: 1.8k cycles per block)
        const uint32_t st = smem_u32(smem) + 96 * 1024 + (warp - 2) * 4096;
        const float bias = (epi.bias && n0 + tl < p.N) ? epi.bias[n0 + tl] : 0.f;
        if (lane == 0) tma_store_wait_read();           // the previous store has finished reading this buffer
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) sts_f32(st + (j * 32 + lane) * 4, fmaf(v[j], epi.alpha, bias));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tma_store_2d(&p.tc, st, n0 + q * 32, b0);
      } else {
        epilogue_swapped32(epi, p.vec, p.M, p.N, n0 + q * 32, b0, v,
                           smem_u32(smem) + 96 * 1024 + (warp - 2) * 32 * 36 * 4);
      }
    }
  }
  if (p.tma_store && warp >= 2 && lane == 0) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BX>(tmem_acc);
  }
}

// ---------------------------------------------------------------------------------------------
// Skinny-M kernel, CTA-PAIR form: the swapped-operand cluster split-K kernel above with every 256 x 256 tile driven
// by tcgen05.mma.cta_group::2. The chip-wide L2 -> SM rate (~6300 B/clk, B300_MICROARCH.md; measured here as
// ~77 GB/s per SM with all SMs pulling) bounds these GEMMs, and in the single-CTA form 2/3 of every k-block's 48 KB
// is the ACTIVATION tile, re-read by every one of the N/128 weight-row CTAs. A pair splits that tile: CTA x of the
// pair loads its own 128 weight rows (16 KB) and HALF of the batch rows (16 KB), the MMA (M = 256 weight rows,
// N = 256 batch rows) reads both halves across the pair. 32 KB instead of 48 KB per SM and k-block: the loads of a
// 128x256x64 block (~415 ns at 77 GB/s) now hide behind its MMAs (~440 ns at the sustained tensor rate).
// Cluster = (2, 1, S): x = the pair, z = the K split (S = 1..4); CTA (x, z) ends up owning the batch columns
// [z*CW, z*CW+CW) of its 128 weight rows, receives the other splits' partial sums for them through distributed
// shared memory and runs the epilogue, exactly as in the single-CTA form.
// ---------------------------------------------------------------------------------------------
// Fused LSTM cell (updown_cell.py:143-148 etc., nn.LSTMCell gate order i,f,g,o). A thread owns hidden unit j for 8
// batch rows [r0, r0+8) of a 32-row block. Everything the cell needs besides the GEMM result - the addends of the four
// gate pre-activations and c_{t-1} - does not depend on the GEMM, so it is loaded into registers one block AHEAD (before
// the accumulator is ready / while the previous block is processed): the epilogue never waits on global memory.
struct LstmCellPre { float pre[8][4]; float cp[8]; };

// Branch-free forms for the epilogue: its four warps run one per SM sub-partition with nothing to hide instruction
// latency behind, and tanhf / IEEE division are ~4x the instructions (measured: 850 cycles per cell with them).
// Absolute error <= 2e-7, far inside the bf16 operand rounding of the next GEMM.
__device__ __forceinline__ float cell_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float cell_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

__device__ __forceinline__ void lstm_cell_prefetch(const LstmFwdArgs& a, int j, int r0, LstmCellPre& P) {
  const int H = a.H;
  if (j >= H) return;
  float bias[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) bias[k] = a.bias ? a.bias[k * H + j] : 0.f;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int r = min(r0 + t, a.R - 1);                 // clamped: rows beyond R are loaded but never stored
    const int r2 = a.rowmap ? a.rowmap[r] : r;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int na = lstm_gate_row(k, j);
      float v = bias[k];
      if (a.add1) v += a.add1[(size_t)r * a.ld1 + na];
      if (a.add2) v += a.add2[(size_t)r2 * a.ld2 + na];
      if (a.sent) v += a.sent[r2] * a.scol[k * H + j];
      P.pre[t][k] = v;
    }
    P.cp[t] = a.c_prev ? a.c_prev[(size_t)r * H + j] : 0.f;
  }
}

// the pre-activation (GEMM part) of gate k, row r0+t sits at xch + (k*32*32 + t*32) floats.
// stg != 0: the results are staged in shared memory as dense [32 rows][32 units] tiles (gates i,f,g,o and c in fp32 at
// stg + o*4096, h in bf16 at stg + 20480) and leave through TMA stores issued by the caller: global stores cost a lone
// epilogue warp ~220 cycles each (measured), and a cell has seven of them. Units beyond H stage zeros: the fp32 maps
// clip them, the bf16 maps write them into the zero padding columns of the operand buffers.
__device__ __forceinline__ void lstm_cell_rows(const LstmFwdArgs& a, int j, int r0, uint32_t xch, const LstmCellPre& P,
                                               uint32_t stg, int rb0) {
  const int H = a.H;
  if (stg) {
    const int lane = threadIdx.x & 31;
    const bool valid = j < H;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      float g[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) g[k] = lds_f32(xch + (uint32_t)((k * 32 * 32 + t * 32) * 4)) + P.pre[t][k];
      float i = cell_sigmoid(g[0]), f = cell_sigmoid(g[1]), gg = cell_tanh(g[2]), o = cell_sigmoid(g[3]);
      float c = f * P.cp[t] + i * gg;
      float h = o * cell_tanh(c);
      if (!valid) { i = f = gg = o = c = h = 0.f; }
      const uint32_t off = (uint32_t)(((rb0 + t) * 32 + lane) * 4);
      sts_f32(stg + off, i); sts_f32(stg + 4096 + off, f); sts_f32(stg + 8192 + off, gg); sts_f32(stg + 12288 + off, o);
      sts_f32(stg + 16384 + off, c);
      const unsigned short hb = __bfloat16_as_ushort(__float2bfloat16_rn(h));
      asm volatile("st.shared.u16 [%0], %1;" ::"r"(stg + 20480 + (off >> 1)), "h"(hb) : "memory");
    }
    return;
  }
  if (j >= H) return;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int r = r0 + t;
    if (r < a.R) {
      float g[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) g[k] = lds_f32(xch + (uint32_t)((k * 32 * 32 + t * 32) * 4)) + P.pre[t][k];
      const float i = cell_sigmoid(g[0]), f = cell_sigmoid(g[1]), gg = cell_tanh(g[2]), o = cell_sigmoid(g[3]);
      const float c = f * P.cp[t] + i * gg;
      const float h = o * cell_tanh(c);
      a.c_out[(size_t)r * H + j] = c;
      if (a.gates_out) {
        float* go = a.gates_out + (size_t)r * 4 * H;
        go[j] = i; go[H + j] = f; go[2 * H + j] = gg; go[3 * H + j] = o;
      }
      const bf16 hb = __float2bfloat16_rn(h);
      if (a.h1_dst) a.h1_dst[(size_t)r * a.ld_h1 + j] = hb;
      if (a.h2_dst) a.h2_dst[(size_t)r * a.ld_h2 + j] = hb;
    }
  }
}

__device__ __forceinline__ void umma_commit_2sm_mask(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

#define DBG_STAMP(i)                                                                                         \
  do {                                                                                                     \
    if (p.dbg && warp == 2 && lane == 0)                                                                   \
      p.dbg[((size_t)blockIdx.z * gridDim.x + blockIdx.x) * 16 + (i)] = clock64();                          \
  } while (0)

template <int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tcgen05_swapped_pair_kernel(const __grid_constant__ GemmParams p) {
  constexpr int BW = 128, BXH = 128;                    // weight rows per CTA (half of UMMA M), batch rows LOADED per CTA
  constexpr int W_STAGE_BYTES = BW * BK * 2, X_STAGE_BYTES = BXH * BK * 2;
  constexpr uint32_t IDESC = make_idesc(2 * BW, 2 * BXH);
  constexpr int TCOLS = 2 * BXH;                        // accumulator: 128 lanes (this CTA's weight rows) x 256 batch rows
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  uint8_t* smem_w = smem;
  uint8_t* smem_x = smem + STAGES * W_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_x + STAGES * X_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int xr = blockIdx.x & 1;                        // position in the pair; cluster rank = xr + 2 * z
  const int S = p.splits;
  const int rank = blockIdx.z;                          // K split of this CTA (cluster z == grid z)
  const int n0 = (blockIdx.x >> 1) * (2 * BW) + xr * BW;   // first weight row (output column) of this CTA
  DBG_STAMP(0);

  int total_kb = 0;
#pragma unroll
  for (int s = 0; s < MAX_SEG; ++s) total_kb += (s < p.nseg) ? p.kblocks[s] : 0;
  const int kb_begin = rank * p.kb_per_split;
  const int kb_end = min(total_kb, kb_begin + p.kb_per_split);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nseg; ++s) { prefetch_tmap(&p.ta[s]); prefetch_tmap(&p.tb[s]); }
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc_2sm<TCOLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();                                   // the peer's barriers exist before any remote arrival
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  DBG_STAMP(1);
  pdl_wait();                                           // everything above overlaps the previous kernel's tail
  pdl_launch_dependents(1);
  DBG_STAMP(2);
  const int CW = S == 1 ? 256 : S == 2 ? 128 : S == 3 ? 96 : 64;   // batch columns owned by each K split
  LstmCellPre cell_pre;                                 // fused LSTM: addends + c_{t-1} of this thread's next 8 cells
  const LstmFwdArgs lstm = p.lstm;                      // registers (see the note at epilogue_block_vec)

  if (warp == 0) {
    if (elect_one_sync()) {
      int stage = 0; uint32_t phase = 0;
      int s = 0, base = 0;
      uint64_t w_pol = 0;
      if (p.w_policy) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(w_pol));
      for (int g = kb_begin; g < kb_end; ++g) {
        while (g >= base + p.kblocks[s]) { base += p.kblocks[s]; ++s; }
        const int kb = g - base;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (xr == 0) mbar_expect_tx(&full_bar[stage], 2 * (W_STAGE_BYTES + X_STAGE_BYTES));
        if (p.w_policy) tma_load_2d_2sm_hint(smem_w + stage * W_STAGE_BYTES, &p.tb[s], &full_bar[stage], kb * BK, n0, w_pol);
        else tma_load_2d_2sm(smem_w + stage * W_STAGE_BYTES, &p.tb[s], &full_bar[stage], kb * BK, n0);
        tma_load_2d_2sm(smem_x + stage * X_STAGE_BYTES, &p.ta[s], &full_bar[stage], kb * BK, xr * BXH);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (xr == 0 && elect_one_sync()) {
      const uint16_t pair_mask = static_cast<uint16_t>(3u << (2 * rank));
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t w_base = smem_u32(smem_w + stage * W_STAGE_BYTES);
        const uint32_t x_base = smem_u32(smem_x + stage * X_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16_2sm(tmem_acc, make_smem_desc(w_base + k * 32), make_smem_desc(x_base + k * 32), IDESC,
                        (kb > kb_begin || k > 0) ? 1u : 0u);
        umma_commit_2sm_mask(&empty_bar[stage], pair_mask);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit_2sm_mask(accum_bar, pair_mask);
    }
    __syncwarp();
  } else {
    if (p.fuse_lstm) lstm_cell_prefetch(lstm, (n0 >> 2) + lane, rank * CW + (warp - 2) * 8, cell_pre);
    DBG_STAMP(3);
    mbar_wait(accum_bar, 0);                            // this pair's partial tile is complete in both CTAs' TMEM
    tc_fence_after();
    DBG_STAMP(4);
  }

  const int q = warp & 3;                               // epilogue warps 2..5 own TMEM lanes [32q, 32q+32)
  const int tl = q * 32 + lane;                         // TMEM lane = column n0 + tl of the output
  float* recv = reinterpret_cast<float*>(smem);         // [(S-1) sources][CW columns][128 lanes], over the dead stages
  // every CTA of the cluster has finished its main loop (stage buffers dead, also the ones the PEER's MMAs read)
  cluster_sync_all();
  DBG_STAMP(5);
  if (S > 1) {
    if (warp >= 2 && kb_end > kb_begin) {
      for (int o = 0; o < S; ++o) {
        if (o == rank) continue;
        const int slot = rank < o ? rank : rank - 1;
        const uint32_t rbase = mapa_shared(smem_u32(recv), (uint32_t)(xr + 2 * o));
        for (int c = 0; c < CW; c += 32) {
          if (o * CW + c >= p.M || o * CW + c >= TCOLS) break;   // columns beyond the batch are never read
          uint32_t acc[32];
          tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + o * CW + c, acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            st_shared_cluster_f32(rbase + (uint32_t)(((slot * CW + c + j) * 128 + tl) * 4), __uint_as_float(acc[j]));
        }
      }
    }
    DBG_STAMP(6);
    cluster_sync_all();
    DBG_STAMP(7);
  }
  if (warp >= 2) {
    const GemmEpi epi = p.epi;
    for (int c = 0; c < CW; c += 32) {
      const int b0 = rank * CW + c;
      if (b0 >= p.M || b0 >= TCOLS) break;
      float v[32];
      if (kb_end > kb_begin) {
        uint32_t acc[32];
        tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + b0, acc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      for (int o = 0; o < S; ++o) {
        if (o == rank || o * p.kb_per_split >= total_kb) continue;
        const int slot = o < rank ? o : o - 1;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += lds_f32(smem_u32(recv) + (uint32_t)(((slot * CW + c + j) * 128 + tl) * 4));
      }
      if (p.fuse_lstm) {
        // The 128 weight rows of this CTA are 4 gates x 32 hidden units (lstm_gate_row): warp q holds gate q of unit
        // `lane` for 32 batch rows. Exchange through shared memory, then warp w runs the cell for 8 of the 32 batch
        // rows with lane = unit, so every global access is 32 consecutive units of one row (128 / 64 bytes).
        const uint32_t xch = smem_u32(smem) + 96 * 1024;          // [gate][batch row][unit] fp32, 16 KB
        const uint32_t stg = p.lstm_tma ? smem_u32(smem) + 112 * 1024 : 0u;   // output tiles for the TMA stores, 22 KB
#pragma unroll
        for (int j = 0; j < 32; ++j) sts_f32(xch + ((q * 32 + j) * 32 + lane) * 4, v[j]);
        if (c == 0) DBG_STAMP(8);
        if (stg && warp == 2 && lane == 0) tma_store_wait_read();   // the previous block's stores have read the staging tiles
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (c == 0) DBG_STAMP(9);
        const LstmCellPre cur = cell_pre;
        if (c + 32 < CW && b0 + 32 < p.M && b0 + 32 < TCOLS)      // the next block's addends fly during this block's math
          lstm_cell_prefetch(lstm, (n0 >> 2) + lane, b0 + 32 + (warp - 2) * 8, cell_pre);
        lstm_cell_rows(lstm, (n0 >> 2) + lane, b0 + (warp - 2) * 8, xch + (uint32_t)((warp - 2) * 8 * 32 + lane) * 4, cur, stg,
                       (warp - 2) * 8);
        if (c == 0) DBG_STAMP(10);
        if (stg) fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (c == 0) DBG_STAMP(11);
        if (stg && warp == 2 && lane == 0 && n0 < p.N) {
          const int j0 = n0 >> 2;
          if (lstm.gates_out) {
#pragma unroll
            for (int k = 0; k < 4; ++k) tma_store_2d(&p.tl[k], stg + k * 4096, j0, b0);
          }
          tma_store_2d(&p.tl[4], stg + 16384, j0, b0);
          if (lstm.h1_dst) tma_store_2d(&p.tl[5], stg + 20480, j0, b0);
          if (lstm.h2_dst) tma_store_2d(&p.tl[6], stg + 20480, j0, b0);
        }
      } else if (p.tma_store) {
        const uint32_t st = smem_u32(smem) + 96 * 1024 + (warp - 2) * 4096;
        const float bias = (epi.bias && n0 + tl < p.N) ? epi.bias[n0 + tl] : 0.f;
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) sts_f32(st + (j * 32 + lane) * 4, fmaf(v[j], epi.alpha, bias));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && n0 + q * 32 < p.N) tma_store_2d(&p.tc, st, n0 + q * 32, b0);
      } else if (n0 + q * 32 < p.N) {
        epilogue_swapped32(epi, p.vec, p.M, p.N, n0 + q * 32, b0, v,
                           smem_u32(smem) + 96 * 1024 + (warp - 2) * 32 * 36 * 4);
      }
    }
  }
  DBG_STAMP(12);
  if ((p.tma_store || p.lstm_tma) && warp >= 2 && lane == 0) tma_store_wait_all();
  DBG_STAMP(13);
  tc_fence_before();
  cluster_sync_all();                                   // nobody deallocates while the pair's MMAs / loads are in flight
  DBG_STAMP(14);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<TCOLS>(tmem_acc);
  }
}

// Bring-up aid only (SSCVAE_GEMM_DEBUG_SIMT=1): same contract on CUDA cores, one thread per output.
__global__ void gemm_simt_debug_kernel(int M, int N, int nseg, GemmSeg s0, GemmSeg s1, GemmSeg s2, GemmEpi e) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N || m >= M) return;
  GemmSeg segs[3] = {s0, s1, s2};
  float acc = 0.f;
  for (int s = 0; s < nseg; ++s) {
    const bf16* a = segs[s].A + (size_t)m * segs[s].lda;
    const bf16* b = segs[s].B + (size_t)n * segs[s].ldb;
    for (int k = 0; k < segs[s].K; ++k) acc += __bfloat162float(a[k]) * __bfloat162float(b[k]);
  }
  float x = acc * e.alpha;
  if (e.bias) x += e.bias[n];
  if (e.add1) x += e.add1[(size_t)m * e.ld1 + n];
  if (e.add2) x += e.add2[(size_t)m * e.ld2 + n];
  if (e.act == 1) x = tanhf(x);
  if (e.dtanh) { float t = e.dtanh[(size_t)m * e.ldd + n]; x *= 1.0f - t * t; }
  if (e.C32) { float* p = e.C32 + (size_t)m * e.ldc32 + n; *p = e.accumulate ? (*p + x) : x; }
  if (e.C16) e.C16[(size_t)m * e.ldc16 + n] = __float2bfloat16_rn(x);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    // resolved through the runtime so the library has no link-time dependency on libcuda
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

void* tma_encode_fn() { return reinterpret_cast<void*>(get_encode_fn()); }

// cuTensorMapEncodeTiled costs microseconds of host time and a training step issues ~500 of them with the same
// few hundred (pointer, shape) combinations every step: memoise the encoded descriptors.
struct TmapKey {
  const void* base; int rows, K, ld, box_rows;
  bool operator==(const TmapKey& o) const { return base == o.base && rows == o.rows && K == o.K && ld == o.ld && box_rows == o.box_rows; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.base) * 0x9E3779B97F4A7C15ull;
    h ^= ((size_t)k.rows << 40) ^ ((size_t)k.K << 20) ^ ((size_t)k.ld << 4) ^ (size_t)k.box_rows;
    return h * 0xD6E8FEB86659FD93ull;
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static std::mutex g_tmap_mutex;

static int encode_tmap_uncached(CUtensorMap* out, const bf16* base, int rows, int K, int ld, int box_rows);
static int encode_tmap(CUtensorMap* out, const bf16* base, int rows, int K, int ld, int box_rows) {
  const TmapKey key{base, rows, K, ld, box_rows};
  std::lock_guard<std::mutex> lock(g_tmap_mutex);
  auto it = g_tmap_cache.find(key);
  if (it != g_tmap_cache.end()) { *out = it->second; return 0; }
  TRY(encode_tmap_uncached(out, base, rows, K, ld, box_rows));
  if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();
  g_tmap_cache.emplace(key, *out);
  return 0;
}

static int encode_tmap_uncached(CUtensorMap* out, const bf16* base, int rows, int K, int ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return SSCVAE_ERR_DRIVER; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 8) != 0) {
    set_error("gemm operand not TMA-compatible: base %p ld %d (need 16B-aligned base, ld %% 8 == 0)", (const void*)base, ld);
    return SSCVAE_ERR_BAD_ARG;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  static const CUtensorMapL2promotion promo = [] {
    const char* e = getenv("SSCVAE_TMA_L2_PROMOTION");      // tuning knob: 0 none, 64, 128, 256
    const int v = e ? atoi(e) : 256;
    return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
         : v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  }();
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%d box_rows=%d", (int)r, rows, K, ld, box_rows);
    return SSCVAE_ERR_DRIVER;
  }
  return 0;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static bool tma_store_eligible(const GemmParams& prm) {
  static const bool off = [] { const char* v = getenv("SSCVAE_GEMM_NO_TMA_STORE"); return v && v[0] == '1'; }();
  const GemmEpi& e = prm.epi;
  return !off && prm.splits == 1 && e.C32 && !e.C16 && !e.act && !e.dtanh && !e.accumulate && !e.add1 && !e.add2 &&
         aligned16(e.C32) && (e.ldc32 % 4) == 0 && (prm.N % 4) == 0 && (!e.bias || aligned16(e.bias));
  // N % 4: the TMA unit clips stores at 16-byte granularity (measured: with N = 77 the columns 77..79 were written)
}

// fp32 output tile map for TMA stores: box = 32 columns x 32 rows, dense in shared memory
static int encode_tmap_c32(CUtensorMap* out, const float* base, int rows, int cols, int ld, bool swizzle128) {
  const TmapKey key{base, rows, cols, ld, swizzle128 ? -33 : -32};
  std::lock_guard<std::mutex> lock(g_tmap_mutex);
  auto it = g_tmap_cache.find(key);
  if (it != g_tmap_cache.end()) { *out = it->second; return 0; }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return SSCVAE_ERR_DRIVER; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32 output) failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld);
    return SSCVAE_ERR_DRIVER;
  }
  if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();
  g_tmap_cache.emplace(key, *out);
  return 0;
}

// bf16 output tile map for TMA stores: box = 32 columns x 32 rows (64-byte rows), dense in shared memory
static int encode_tmap_h16(CUtensorMap* out, const bf16* base, int rows, int cols, int ld) {
  const TmapKey key{base, rows, cols, ld, -34};
  std::lock_guard<std::mutex> lock(g_tmap_mutex);
  auto it = g_tmap_cache.find(key);
  if (it != g_tmap_cache.end()) { *out = it->second; return 0; }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return SSCVAE_ERR_DRIVER; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (bf16 output) failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld);
    return SSCVAE_ERR_DRIVER;
  }
  if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();
  g_tmap_cache.emplace(key, *out);
  return 0;
}

template <int BN, int STAGES>
static int launch_tc(cudaStream_t stream, GemmParams& prm, const GemmSeg* segs) {
  for (int s = 0; s < prm.nseg; ++s) {
    TRY(encode_tmap(&prm.ta[s], segs[s].A, prm.M, segs[s].K, segs[s].lda, BM));
    TRY(encode_tmap(&prm.tb[s], segs[s].B, prm.N, segs[s].K, segs[s].ldb, BN));
  }
  prm.tma_store = BN >= 32 && tma_store_eligible(prm);
  if (prm.tma_store) TRY(encode_tmap_c32(&prm.tc, prm.epi.C32, prm.M, prm.N, prm.epi.ldc32, true));
  constexpr int smem = 1024 + STAGES * (A_STAGE_BYTES + BN * BK * 2) + 256;
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid(ceil_div(prm.N, BN), ceil_div(prm.M, BM), prm.splits);
  CUDA_TRY(launch_pdl(gemm_tcgen05_kernel<BN, STAGES>, grid, dim3(GEMM_THREADS), smem, stream, prm));
  ++g_launch_count;
  return 0;
}

template <int BN, int STAGES>
static int launch_tc2(cudaStream_t stream, GemmParams& prm, const GemmSeg* segs) {
  for (int s = 0; s < prm.nseg; ++s) {
    TRY(encode_tmap(&prm.ta[s], segs[s].A, prm.M, segs[s].K, segs[s].lda, BM));
    TRY(encode_tmap(&prm.tb[s], segs[s].B, prm.N, segs[s].K, segs[s].ldb, BN / 2));
  }
  prm.tma_store = BN >= 32 && tma_store_eligible(prm);
  if (prm.tma_store) TRY(encode_tmap_c32(&prm.tc, prm.epi.C32, prm.M, prm.N, prm.epi.ldc32, true));
  constexpr int smem = 1024 + STAGES * (A_STAGE_BYTES + (BN / 2) * BK * 2) + 256;
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid(2 * ceil_div(prm.N, BN), ceil_div(prm.M, 2 * BM), prm.splits);
  CUDA_TRY(launch_pdl(gemm_tcgen05_2cta_kernel<BN, STAGES>, grid, dim3(GEMM_THREADS), smem, stream, prm));
  ++g_launch_count;
  return 0;
}

template <int STAGES>
static int launch_swapped(cudaStream_t stream, GemmParams& prm, const GemmSeg* segs) {
  for (int s = 0; s < prm.nseg; ++s) {
    TRY(encode_tmap(&prm.ta[s], segs[s].A, prm.M, segs[s].K, segs[s].lda, 256));
    TRY(encode_tmap(&prm.tb[s], segs[s].B, prm.N, segs[s].K, segs[s].ldb, 128));
  }
  const GemmEpi& e = prm.epi;
  static const bool no_tma_store = [] { const char* v = getenv("SSCVAE_GEMM_NO_TMA_STORE"); return v && v[0] == '1'; }();
  prm.tma_store = !no_tma_store && e.C32 && !e.C16 && !e.act && !e.dtanh && !e.accumulate && !e.add1 && !e.add2 &&
                  aligned16(e.C32) && (e.ldc32 % 4) == 0 && (prm.N % 4) == 0;
  if (prm.tma_store) TRY(encode_tmap_c32(&prm.tc, e.C32, prm.M, prm.N, e.ldc32, false));
  constexpr int smem = 1024 + STAGES * (128 + 256) * BK * 2 + 256;
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_swapped_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ceil_div(prm.N, 128), 1, prm.splits);
  cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = prm.splits;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (pdl_enabled()) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_tcgen05_swapped_kernel<STAGES>, prm));
  ++g_launch_count;
  return 0;
}

// max co-resident clusters of the pair kernel for a K split of S (cluster = 2*S CTAs); 0 if the query fails
template <int STAGES>
static int pair_max_clusters(int S, int smem) {
  static int cache[5] = {-1, -1, -1, -1, -1};
  if (cache[S] >= 0) return cache[S];
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * 64, 1, S);
  cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = S;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm_tcgen05_swapped_pair_kernel<STAGES>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  cache[S] = n;
  return n;
}

template <int STAGES>
static int launch_swapped_pair(cudaStream_t stream, GemmParams& prm, const GemmSeg* segs, int forced_S) {
  constexpr int smem = 1024 + STAGES * (128 + 128) * BK * 2 + 256;
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_swapped_pair_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  int total_kb = 0;
  for (int s = 0; s < prm.nseg; ++s) total_kb += prm.kblocks[s];
  const int pairs = ceil_div(prm.N, 256);
  // the largest K split whose clusters are all co-resident (one wave) and that leaves every CTA >= 4 k-blocks
  int S = 1;
  if (forced_S > 0) S = forced_S;
  else
    for (int c = 4; c > 1; --c)
      if (total_kb >= 4 * c && pairs <= pair_max_clusters<STAGES>(c, smem)) { S = c; break; }
  while (S > 1 && total_kb < S) --S;
  prm.kb_per_split = ceil_div(total_kb, S);
  S = ceil_div(total_kb, prm.kb_per_split);             // no empty trailing split
  prm.splits = S;
  for (int s = 0; s < prm.nseg; ++s) {
    TRY(encode_tmap(&prm.ta[s], segs[s].A, prm.M, segs[s].K, segs[s].lda, 128));
    TRY(encode_tmap(&prm.tb[s], segs[s].B, prm.N, segs[s].K, segs[s].ldb, 128));
  }
  const GemmEpi& e = prm.epi;
  static const bool no_tma_store = [] { const char* v = getenv("SSCVAE_GEMM_NO_TMA_STORE"); return v && v[0] == '1'; }();
  prm.tma_store = !no_tma_store && e.C32 && !e.C16 && !e.act && !e.dtanh && !e.accumulate && !e.add1 && !e.add2 &&
                  aligned16(e.C32) && (e.ldc32 % 4) == 0 && (prm.N % 4) == 0;
  // The weights of the skinny GEMMs (76 MB per timestep) are streamed once per launch; marking them evict_first leaves
  // L2 to the 52 MB of region features + projections the attention kernels re-read every step (measured: step
  // 8.21 -> 8.08 ms, attention_fwd 25.1 -> 23.1 us). SSCVAE_W_EVICT_FIRST=0 turns the hint off.
  static const bool w_evict_first = [] { const char* v = getenv("SSCVAE_W_EVICT_FIRST"); return !(v && v[0] == '0'); }();
  prm.w_policy = w_evict_first ? 1 : 0;
  prm.fuse_lstm = 0;
  prm.lstm_tma = 0;
  if (e.lstm) {
    prm.fuse_lstm = 1;
    prm.lstm = *e.lstm;
    prm.tma_store = 0;
    // outputs through TMA stores when every destination is 16-byte tileable (true for all training / decode buffers);
    // the bf16 h destinations must have room for round_up(H, 32) columns (their K padding: Hp = round_up(H, 64))
    const LstmFwdArgs& l = *e.lstm;
    const int H = l.H, Hq = round_up(H, 32);
    static const bool no_lstm_tma = [] { const char* v = getenv("SSCVAE_LSTM_TMA"); return v && v[0] == '0'; }();
    bool ok = !no_lstm_tma && (H % 4) == 0 && aligned16(l.c_out) && (!l.gates_out || aligned16(l.gates_out));
    ok = ok && (!l.h1_dst || (aligned16(l.h1_dst) && (l.ld_h1 % 8) == 0 && l.ld_h1 >= Hq));
    ok = ok && (!l.h2_dst || (aligned16(l.h2_dst) && (l.ld_h2 % 8) == 0 && l.ld_h2 >= Hq));
    if (ok) {
      prm.lstm_tma = 1;
      if (l.gates_out)
        for (int k = 0; k < 4; ++k) TRY(encode_tmap_c32(&prm.tl[k], l.gates_out + (size_t)k * H, l.R, H, 4 * H, false));
      TRY(encode_tmap_c32(&prm.tl[4], l.c_out, l.R, H, H, false));
      if (l.h1_dst) TRY(encode_tmap_h16(&prm.tl[5], l.h1_dst, l.R, Hq, l.ld_h1));
      if (l.h2_dst) TRY(encode_tmap_h16(&prm.tl[6], l.h2_dst, l.R, Hq, l.ld_h2));
    }
  }
  if (prm.tma_store) TRY(encode_tmap_c32(&prm.tc, e.C32, prm.M, prm.N, e.ldc32, false));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs, 1, S);
  cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = S;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (pdl_enabled()) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  static const bool dbg = [] { const char* v = getenv("SSCVAE_GEMM_DBG"); return v && v[0] == '1'; }();
  static long long* dbg_buf = nullptr;
  if (dbg) {                                            // bring-up aid: phase stamps of a few CTAs, printed per launch
    if (!dbg_buf) CUDA_TRY(cudaMalloc(&dbg_buf, 16 * 8 * 1024));
    CUDA_TRY(cudaMemsetAsync(dbg_buf, 0, 16 * 8 * 1024, stream));
    prm.dbg = dbg_buf;
  }
  CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_tcgen05_swapped_pair_kernel<STAGES>, prm));
  ++g_launch_count;
  if (dbg) {
    static int printed = 0;
    CUDA_TRY(cudaStreamSynchronize(stream));
    std::vector<long long> h(16 * 2 * pairs * S);
    CUDA_TRY(cudaMemcpy(h.data(), dbg_buf, h.size() * 8, cudaMemcpyDeviceToHost));
    if (printed++ < 400) {
      for (int cta : {0, 2 * pairs * S - 2}) {
        fprintf(stderr, "[pairdbg] M=%d N=%d kb=%d S=%d fuse=%d tma=%d cta=%d:", prm.M, prm.N, total_kb, S, prm.fuse_lstm,
                prm.fuse_lstm ? prm.lstm_tma : prm.tma_store, cta);
        for (int i = 1; i < 15; ++i) fprintf(stderr, " %lld", h[cta * 16 + i] ? h[cta * 16 + i] - h[cta * 16] : -1);
        fprintf(stderr, "\n");
      }
    }
  }
  return 0;
}

int gemm_suggest_splits(int M, int N, int K_total) {
  if (M > 256) return 1;
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      n_sm = 148;
  }
  const int tiles = ceil_div(N, 128), kb = ceil_div(K_total, BK);
  int s = 1;
  while (s < 4 && tiles * s * 2 <= n_sm && kb >= 4 * s * 2) s *= 2;   // power of two: the cluster splits 256 columns
  return s;
}

static bool pair_kernel_disabled() {
  static const bool off = [] {
    const char* a = getenv("SSCVAE_GEMM_PAIR");
    const char* b = getenv("SSCVAE_GEMM_NO_SWAPPED");
    const char* c = getenv("SSCVAE_GEMM_DEBUG_SIMT");
    return (a && a[0] == '0') || (b && b[0] == '1') || (c && c[0] == '1');
  }();
  return off || getenv("SSCVAE_GEMM_FORCE") != nullptr;
}

int gemm_bf16_tn(cudaStream_t stream, int M, int N, int nseg, const GemmSeg* segs, const GemmEpi& epi) {
  REQUIRE(M > 0 && N > 0 && nseg >= 1 && nseg <= MAX_SEG, "gemm: bad shape M=%d N=%d nseg=%d", M, N, nseg);
  REQUIRE(epi.C32 || epi.C16 || (epi.rs && epi.rs->mode == 1), "gemm: no output");
  if (epi.lstm) {
    REQUIRE(N == lstm_gate_rows(epi.lstm->H) && epi.lstm->R == M, "gemm: fused LSTM shape mismatch (N=%d H=%d M=%d R=%d)", N,
            epi.lstm->H, M, epi.lstm->R);
    REQUIRE(epi.C32 && epi.ldc32 >= N && !epi.C16 && !epi.bias && !epi.add1 && !epi.add2 && !epi.act && !epi.dtanh && !epi.accumulate,
            "gemm: fused LSTM takes a plain fp32 accumulator buffer");
    // SSCVAE_LSTM_FUSE: 0 (default) = cell kernel behind the GEMM, 1 = fuse cells without per-row addend matrices
    // (encoder / decoder LSTM), 2 = fuse every cell. Measured on B200 (bench27, ms per training step): 0: 7.45,
    // 1: 7.49, 2: 7.61. Phase stamps (SSCVAE_GEMM_DBG=1) show why: the four epilogue warps run one per SM
    // sub-partition with nothing to hide latency behind - ~170 cycles per global load, ~400 cycles per cell even with
    // branch-free math - while the separate cell kernel is one wave of 57 600 threads that takes 4 us. The fused
    // form stays as an opt-in for shapes where launches, not the epilogue, dominate.
    static const int fuse_mode = [] { const char* e = getenv("SSCVAE_LSTM_FUSE"); return e ? atoi(e) : 0; }();
    const bool heavy = epi.lstm->add1 || epi.lstm->add2;
    if (M > 256 || fuse_mode == 0 || (fuse_mode == 1 && heavy) || pair_kernel_disabled()) {   // unfused: cell kernel behind
      GemmEpi e2 = epi;
      e2.lstm = nullptr;
      TRY(gemm_bf16_tn(stream, M, N, nseg, segs, e2));
      LstmFwdArgs l = *epi.lstm;
      l.acc = epi.C32; l.ld_acc = epi.ldc32; l.perm = 1;
      return lstm_forward(stream, l);
    }
  }
  double ksum = 0;
  for (int i = 0; i < nseg; ++i) ksum += segs[i].K;
  // instrumentation class: the skinny long-K GEMMs of the recurrence (the dominant kernel of a training step) are
  // reported on their own as "gemm.recurrent"
  int kb_all = 0;
  for (int i = 0; i < nseg; ++i) kb_all += ceil_div(segs[i].K, BK);
  const bool recurrent = M <= 256 && N >= 1024 && kb_all >= 16;
  PROF_SCOPE(stream, recurrent ? "gemm.recurrent" : epi.tag, 2.0 * M * N * ksum,
             2.0 * (M + N) * ksum + (epi.C32 ? 4.0 : 2.0) * M * N);
  static const bool simt = [] { const char* e = getenv("SSCVAE_GEMM_DEBUG_SIMT"); return e && e[0] == '1'; }();
  if (simt) {
    GemmSeg z{nullptr, 0, nullptr, 0, 0};
    dim3 grid(ceil_div(N, 128), M);
    gemm_simt_debug_kernel<<<grid, 128, 0, stream>>>(M, N, nseg, segs[0], nseg > 1 ? segs[1] : z,
                                                    nseg > 2 ? segs[2] : z, epi);
    CUDA_TRY(cudaGetLastError());
    ++g_launch_count;
    return 0;
  }
  GemmParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.nseg = nseg; prm.M = M; prm.N = N; prm.epi = epi;
  if (epi.rs) prm.rs = *epi.rs;
  for (int s = 0; s < nseg; ++s) {
    REQUIRE(segs[s].K > 0, "gemm: empty K segment %d", s);
    prm.kblocks[s] = ceil_div(segs[s].K, BK);
  }
  bool vec = true;
  if (epi.bias) vec &= aligned16(epi.bias);
  if (epi.add1) vec &= aligned16(epi.add1) && (epi.ld1 % 4 == 0);
  if (epi.add2) vec &= aligned16(epi.add2) && (epi.ld2 % 4 == 0);
  if (epi.dtanh) vec &= aligned16(epi.dtanh) && (epi.ldd % 4 == 0);
  if (epi.C32) vec &= aligned16(epi.C32) && (epi.ldc32 % 4 == 0);
  if (epi.C16) vec &= aligned16(epi.C16) && (epi.ldc16 % 8 == 0);
  prm.vec = vec ? 1 : 0;
  int total_kb = 0;
  for (int s = 0; s < nseg; ++s) total_kb += prm.kblocks[s];
  prm.splits = 1; prm.kb_per_split = total_kb;
  // split-K: gridDim.z CTAs share one output tile, each adds its partial sum into C32 with vector reductions
  // (red.global.add.v4.f32). Only for linear epilogues with an fp32 output; C32 is zeroed first unless accumulating.
  auto use_splits = [&](int n) -> int {
    if (n <= 1 || !epi.C32 || epi.C16 || epi.act || epi.dtanh) return 0;
    n = std::min(n, std::max(1, total_kb / 2));
    if (n <= 1) return 0;
    prm.kb_per_split = ceil_div(total_kb, n);
    prm.splits = ceil_div(total_kb, prm.kb_per_split);
    if (!epi.accumulate)
      CUDA_TRY(cudaMemset2DAsync(epi.C32, (size_t)epi.ldc32 * 4, 0, (size_t)N * 4, (size_t)M, stream));
    return 0;
  };
  if (epi.rs && epi.rs->mode) {                           // vocabulary head with softmax statistics / CE gradient in the epilogue
    REQUIRE(!epi.add1 && !epi.add2 && !epi.act && !epi.dtanh && !epi.accumulate && !epi.lstm, "gemm: row statistics take a plain (alpha, bias) epilogue");
    REQUIRE(epi.rs->mode != 1 || (epi.rs->st_max && epi.rs->st_sum && epi.rs->st_arg), "gemm: row statistics need their partial buffers");
    REQUIRE(epi.rs->mode != 2 || (epi.rs->lse && epi.rs->gcoef && epi.rs->target && epi.C16), "gemm: CE gradient needs lse, gcoef, target and a bf16 output");
    prm.tma_store = 0;
    GemmParams& q = prm;
    for (int s2 = 0; s2 < q.nseg; ++s2) {
      TRY(encode_tmap(&q.ta[s2], segs[s2].A, q.M, segs[s2].K, segs[s2].lda, BM));
      TRY(encode_tmap(&q.tb[s2], segs[s2].B, q.N, segs[s2].K, segs[s2].ldb, 128));
    }
    constexpr int smem_rs = 1024 + 3 * (A_STAGE_BYTES + 128 * BK * 2) + 256;
    static bool configured_rs = false;
    if (!configured_rs) {
      CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_rs));
      configured_rs = true;
    }
    dim3 grid(ceil_div(q.N, 128), ceil_div(q.M, BM), 1);
    CUDA_TRY(launch_pdl(gemm_tcgen05_kernel<128, 3>, grid, dim3(GEMM_THREADS), smem_rs, stream, q));
    ++g_launch_count;
    return 0;
  }
  if (const char* e = getenv("SSCVAE_GEMM_SPLITK")) TRY(use_splits(atoi(e)));
  static const bool no_swapped = [] { const char* e = getenv("SSCVAE_GEMM_NO_SWAPPED"); return e && e[0] == '1'; }();
  // skinny and long-K (the per-timestep LSTM and BPTT GEMMs): swapped-operand cluster split-K. Measured on B200
  // (tools/gemm_bench.py, GPU time inside a CUDA graph, us): 256x3600x4928: 17.0 vs 24.9 for the 128x64 tiling;
  // 256x4160x3648: 16.0 vs 19.4; 256x1920x3648: 13.9 vs 18.8; short K (256x3600x1920: 12.6 vs 12.4) and narrow N
  // (256x768x960: 10.2 vs 7.9) stay on the plain kernel.
  const bool skinny = M <= 256 && N >= 1024 && total_kb >= 48;
  // CTA-pair form (see the kernel): every skinny GEMM with enough weight rows, short K included (the attention-LSTM
  // recurrence, K = 2Hp). SSCVAE_GEMM_PAIR=0 falls back to the single-CTA swapped kernel; epi.splits < 0 forces the
  // pair kernel with K split -epi.splits (tools/gemm_bench.py, tests).
  static const bool no_pair = [] { const char* e = getenv("SSCVAE_GEMM_PAIR"); return e && e[0] == '0'; }();
  const bool pair_shape = M <= 256 && N >= 1024 && total_kb >= 16;
  if ((epi.lstm || epi.splits < 0 || (epi.splits == 0 && pair_shape && !no_pair && !no_swapped)) && M <= 256 &&
      !getenv("SSCVAE_GEMM_FORCE")) {
    REQUIRE(epi.splits >= -4, "gemm: pair split count %d (1..4)", -epi.splits);
    return launch_swapped_pair<6>(stream, prm, segs, epi.splits < 0 ? -epi.splits : 0);
  }
  if ((epi.splits > 0 || (skinny && !no_swapped)) && M <= 256 && !getenv("SSCVAE_GEMM_FORCE")) {
    int S = epi.splits > 0 ? epi.splits : gemm_suggest_splits(M, N, total_kb * BK);
    REQUIRE(S == 1 || S == 2 || S == 4, "gemm: split count %d (1, 2 or 4)", S);
    while (S > 1 && total_kb < S) S /= 2;
    prm.splits = S;
    prm.kb_per_split = ceil_div(total_kb, S);
    return launch_swapped<4>(stream, prm, segs);
  }
  // The recurrent GEMMs have M = batch (two 128-row tiles) and stream their weights once: they are bound by
  // bytes in flight per SM, not by the tensor pipe. So: the widest N tile that still yields >= ~100 CTAs,
  // ~100 KB of TMA stages per CTA and two CTAs resident per SM (no wave-quantisation tail at 149..296 CTAs,
  // one CTA's epilogue overlaps the other's main loop).
  if (const char* f = getenv("SSCVAE_GEMM_FORCE")) {      // tuning knob for tools/gemm_bench.py: "BN,STAGES"
    int bn = 0, st = 0;
    if (sscanf(f, "%d,%d", &bn, &st) == 2) {
      if (bn == 256 && st == 4) return launch_tc<256, 4>(stream, prm, segs);
      if (bn == 128 && st == 6) return launch_tc<128, 6>(stream, prm, segs);
      if (bn == 128 && st == 3) return launch_tc<128, 3>(stream, prm, segs);
      if (bn == 64 && st == 8) return launch_tc<64, 8>(stream, prm, segs);
      if (bn == 64 && st == 4) return launch_tc<64, 4>(stream, prm, segs);
      if (bn == 32 && st == 5) return launch_tc<32, 5>(stream, prm, segs);
      if (bn == 32 && st == 10) return launch_tc<32, 10>(stream, prm, segs);
      if (bn == 16 && st == 6) return launch_tc<16, 6>(stream, prm, segs);
      // CTA-pair kernel: "2xBN,STAGES" is spelled with a leading 2 (e.g. 2256,6 = pair tile 256x256, 6 stages)
      if (bn == 2256 && st == 6) return launch_tc2<256, 6>(stream, prm, segs);
      if (bn == 2256 && st == 4) return launch_tc2<256, 4>(stream, prm, segs);
      if (bn == 2128 && st == 8) return launch_tc2<128, 8>(stream, prm, segs);
      if (bn == 2128 && st == 4) return launch_tc2<128, 4>(stream, prm, segs);
      if (bn == 264 && st == 10) return launch_tc2<64, 10>(stream, prm, segs);
      if (bn == 264 && st == 5) return launch_tc2<64, 5>(stream, prm, segs);
      if (bn == 232 && st == 10) return launch_tc2<32, 10>(stream, prm, segs);
      set_error("SSCVAE_GEMM_FORCE=%s: no such instantiation", f);
      return SSCVAE_ERR_BAD_ARG;
    }
  }
  const long mt = ceil_div(M, BM);
  // large and long-K (the batched weight-gradient GEMMs): CTA pairs, 256x128 tiles (3600x2048x5376: 68 vs 76 us)
  if (M >= 1024 && N >= 512 && total_kb >= 32) return launch_tc2<128, 4>(stream, prm, segs);
  if (mt * ceil_div(N, 128) >= 96) return launch_tc<128, 3>(stream, prm, segs);
  // Small problems are latency-bound, not occupancy-bound: narrower tiles only add CTAs whose MMAs cost the same
  // ~130 cycles and lose the TMA-store epilogue (256x768x960: 7.9 us with 64-wide tiles, 10.4 with 16-wide).
  if (N > 32) return launch_tc<64, 4>(stream, prm, segs);
  if (N > 16) return launch_tc<32, 5>(stream, prm, segs);
  return launch_tc<16, 6>(stream, prm, segs);
}

}  // namespace sscvae
