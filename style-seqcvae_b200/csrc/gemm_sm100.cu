// tcgen05 / TMEM / TMA GEMM for sm_100a (see gemm.cuh for the contract).
//
// One CTA computes one 128 x BN output tile:
//   warp 0      : TMA producer  (cp.async.bulk.tensor.2d, 128B-swizzled K-major tiles, mbarrier ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, bf16 in, fp32 accum in TMEM)
//   warps 2..5  : epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) -> registers ->
//                 fused bias / addends / tanh / dtanh -> fp32 and/or bf16 global stores
// The accumulator never touches registers or shared memory until the epilogue.
#include "gemm.cuh"
#include "prof.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <mutex>

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
  GemmEpi epi;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(addr), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; bf16 operands, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled shared-memory matrix descriptor (cf. cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 ; [16,30) leading byte offset >> 4 (unused for swizzled K-major, =1) ;
//   [32,46) stride byte offset >> 4 = 1024 B (8 rows x 128 B) ; [46,48) version = 1 ; [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor (cf. cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=BF16 [7,10)=1, b=BF16 [10,13)=1,
// a_major/b_major = K (0), n_dim [17,23) = N>>3, m_dim [24,29) = M>>4.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// epilogue math for one run of 32 consecutive columns of one row
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tanh_fast(float x) {
  // accurate enough for fp32 parity: tanhf via exp; the MUFU.TANH approximation is reserved for the
  // attention kernel's 27k tanh/row where the operands are bf16 anyway.
  return tanhf(x);
}

__device__ __forceinline__ void epilogue_row32(const GemmEpi& e, int vec, int M, int N, int row0, int col0, int ncols,
                                               const uint32_t (&acc)[32], float* stage) {
  const int lane = threadIdx.x & 31;
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
      *reinterpret_cast<float4*>(stage + lane * PITCH + 4 * j) =
          make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]), __uint_as_float(acc[4 * j + 2]),
                      __uint_as_float(acc[4 * j + 3]));
    __syncwarp();
    const int c4 = (lane & 7) * 4;
    const int n = col0 + c4;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e.bias) bias4 = *reinterpret_cast<const float4*>(e.bias + n);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int rr = k * 4 + (lane >> 3);
      const int r = row0 + rr;
      if (r < M) {
        float4 v = *reinterpret_cast<const float4*>(stage + rr * PITCH + c4);
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
          if (e.accumulate) {
            const float4 c = *reinterpret_cast<const float4*>(p);
            v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
          }
          *reinterpret_cast<float4*>(p) = v;
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
        *p = e.accumulate ? (*p + x) : x;
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
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int s = 0; s < p.nseg; ++s) {
        for (int kb = 0; kb < p.kblocks[s]; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES + B_STAGE_BYTES);
          tma_load_2d(smem_a + stage * A_STAGE_BYTES, &p.ta[s], &full_bar[stage], kb * BK, m0);
          tma_load_2d(smem_b + stage * B_STAGE_BYTES, &p.tb[s], &full_bar[stage], kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem_a + stage * A_STAGE_BYTES);
        const uint32_t b_base = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 elements (32 bytes) along K inside the 128B swizzle atom
          umma_bf16(tmem_acc, make_smem_desc(a_base + k * 32), make_smem_desc(b_base + k * 32), IDESC,
                    (kb > 0 || k > 0) ? 1u : 0u);
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
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= p.N) break;                 // warp-uniform
      uint32_t acc[32];
      tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + c0, acc);
      tmem_ld_wait();
      epilogue_row32(p.epi, p.vec, p.M, p.N, m0 + q * 32, n0 + c0, BN - c0 < 32 ? BN - c0 : 32, acc,
                     reinterpret_cast<float*>(smem_a) + (warp - 2) * 32 * 36);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_acc);
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

static int encode_tmap(CUtensorMap* out, const bf16* base, int rows, int K, int ld, int box_rows) {
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
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%d box_rows=%d", (int)r, rows, K, ld, box_rows);
    return SSCVAE_ERR_DRIVER;
  }
  return 0;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int BN, int STAGES>
static int launch_tc(cudaStream_t stream, GemmParams& prm, const GemmSeg* segs) {
  for (int s = 0; s < prm.nseg; ++s) {
    TRY(encode_tmap(&prm.ta[s], segs[s].A, prm.M, segs[s].K, segs[s].lda, BM));
    TRY(encode_tmap(&prm.tb[s], segs[s].B, prm.N, segs[s].K, segs[s].ldb, BN));
  }
  constexpr int smem = 1024 + STAGES * (A_STAGE_BYTES + BN * BK * 2) + 256;
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid(ceil_div(prm.N, BN), ceil_div(prm.M, BM));
  gemm_tcgen05_kernel<BN, STAGES><<<grid, GEMM_THREADS, smem, stream>>>(prm);
  CUDA_TRY(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

int gemm_bf16_tn(cudaStream_t stream, int M, int N, int nseg, const GemmSeg* segs, const GemmEpi& epi) {
  REQUIRE(M > 0 && N > 0 && nseg >= 1 && nseg <= MAX_SEG, "gemm: bad shape M=%d N=%d nseg=%d", M, N, nseg);
  REQUIRE(epi.C32 || epi.C16, "gemm: no output");
  double ksum = 0;
  for (int i = 0; i < nseg; ++i) ksum += segs[i].K;
  PROF_SCOPE(stream, epi.tag, 2.0 * M * N * ksum, 2.0 * (M + N) * ksum + (epi.C32 ? 4.0 : 2.0) * M * N);
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
      set_error("SSCVAE_GEMM_FORCE=%s: no such instantiation", f);
      return SSCVAE_ERR_BAD_ARG;
    }
  }
  const long mt = ceil_div(M, BM);
  if (mt * ceil_div(N, 128) >= 96) return launch_tc<128, 3>(stream, prm, segs);
  if (mt * ceil_div(N, 64) >= 96) return launch_tc<64, 4>(stream, prm, segs);
  if (mt * ceil_div(N, 32) >= 96) return launch_tc<32, 5>(stream, prm, segs);
  return launch_tc<16, 6>(stream, prm, segs);
}

}  // namespace sscvae
