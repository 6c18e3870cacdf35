// Probe: how fast can one SM RECEIVE TMA tiles from L2 when every SM pulls the same activation tile (the persistent
// recurrent kernel's situation: all CTA pairs read the same 128 x 64 bf16 batch tile per k-block), and does TMA multicast
// across a thread-block cluster raise that rate?
//   mode 0: unicast, all CTAs load the SAME 16 KB tile per iteration
//   mode 1: unicast, every CTA loads its OWN 16 KB tile (weight-like stream, L2 resident)
//   mode 2: multicast: cluster of C CTAs, CTA r loads rows [128/C * r, +128/C) of the tile and multicasts to all C
// No MMA: the consumer frees a stage as soon as it is full, so the number is the pure ingest rate per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mc_probe tools/mc_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int STAGES = 6;
constexpr int TILE_ROWS = 128, TILE_K = 64, TILE_BYTES = TILE_ROWS * TILE_K * 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void remote_arrive(uint64_t* b, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(b)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void tma_load(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}

struct Params { CUtensorMap full_tile, slice, wtile, x3d; int mode, iters, csize, kmax, rows_total, nmma, mma_n; unsigned long long* out; };
constexpr int W_BYTES = 64 * TILE_K * 2;

__device__ __forceinline__ void tma_load_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2sm_hint(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// modes 3..6: the persistent kernel's k-block: X tile (same for everybody) + W tile (own), 24 KB per iteration
//   3: plain 2-D loads, one CTA          4: CTA pair, cta_group::2 loads signalling the leader's barrier, 2-D X
//   5: as 4 with the 3-D X map           6: as 5 with an L2 evict_first hint on W
__global__ void __launch_bounds__(64, 1) probe2_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int SB = TILE_BYTES + W_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * SB);
  uint64_t* empty = full + STAGES;
  const bool pair = p.mode >= 4;
  const uint32_t rank = pair ? cluster_rank() : 0;
  __shared__ uint32_t tmem_slot;
  if (pair && threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (pair) cluster_sync();
  unsigned long long t0 = 0;
  if (threadIdx.x == 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    int stage = 0; uint32_t phase = 0;
    const int wrow = (int)(((blockIdx.x + 1) * 64) % p.rows_total);
    for (int i = 0; i < p.iters; ++i) {
      const int k = (i * TILE_K) % p.kmax;
      mbar_wait(&empty[stage], phase ^ 1);
      uint8_t* dst = smem + stage * SB;
      if (!pair) {
        mbar_expect_tx(&full[stage], SB);
        tma_load(dst, &p.full_tile, &full[stage], k, 0);
        tma_load(dst + TILE_BYTES, &p.wtile, &full[stage], k, wrow);
      } else {
        if (rank == 0) mbar_expect_tx(&full[stage], 2 * SB);
        if (p.mode == 4 || p.mode == 7) tma_load_2sm(dst, &p.full_tile, &full[stage], k, (int)rank * 128);
        else tma_load_3d_2sm(dst, &p.x3d, &full[stage], k, (int)rank * 128, i & 7);
        if (p.mode == 6) tma_load_2sm_hint(dst + TILE_BYTES, &p.wtile, &full[stage], k, wrow, pol);
        else tma_load_2sm(dst + TILE_BYTES, &p.wtile, &full[stage], k, wrow);
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (p.mode == 7 && threadIdx.x >= 32 && rank == 0) {
    // converged warp: every lane runs the loop, one elected lane issues; loop state is warp-uniform (uniform registers)
    int stage = 0; uint32_t phase = 0;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.mma_n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint32_t tm = tmem_slot;
    const uint32_t base = smem_u32(smem);
    for (int i = 0; i < p.iters; ++i) {
      mbar_wait(&full[stage], phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t elected;
      asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(elected));
      if (elected) {
        const uint32_t xb = base + stage * SB, wb = xb + TILE_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < p.nmma) {
            uint64_t ad = ((uint64_t)(((xb + k * 32) & 0x3FFFF) >> 4)) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
            uint64_t bd = ((uint64_t)(((wb + k * 32) & 0x3FFFF) >> 4)) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
          }
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&empty[stage])), "h"((uint16_t)3) : "memory");
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (threadIdx.x == 32) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      p.out[blockIdx.x * 2 + 1] = t1;
    }
  } else if (p.mode != 7 && threadIdx.x == 32 && rank == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < p.iters; ++i) {
      mbar_wait(&full[stage], phase);
      if (pair && p.nmma > 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t xb = smem_u32(smem + stage * SB), wb = xb + TILE_BYTES;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.mma_n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        for (int k = 0; k < p.nmma; ++k) {
          uint64_t ad = ((uint64_t)(((xb + k * 32) & 0x3FFFF) >> 4)) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
          uint64_t bd = ((uint64_t)(((wb + k * 32) & 0x3FFFF) >> 4)) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_slot), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
        }
      }
      if (pair) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&empty[stage])), "h"((uint16_t)3) : "memory");
      else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    p.out[blockIdx.x * 2 + 1] = t1;
  }
  if (threadIdx.x == 0) p.out[blockIdx.x * 2] = t0;
  __syncthreads();
  if (pair) {
    cluster_sync();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 128;" ::"r"(tmem_slot) : "memory");
  }
}

__global__ void __launch_bounds__(64, 1) probe_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * TILE_BYTES);
  uint64_t* empty = full + STAGES;
  const int C = p.mode == 2 ? p.csize : 1;
  const uint32_t rank = p.mode == 2 ? cluster_rank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], C); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (p.mode == 2) cluster_sync();
  unsigned long long t0 = 0;
  if (threadIdx.x == 0) {                       // producer
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    int stage = 0; uint32_t phase = 0;
    const int slice_rows = TILE_ROWS / C;
    for (int i = 0; i < p.iters; ++i) {
      const int k = (i * TILE_K) % p.kmax;
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_expect_tx(&full[stage], TILE_BYTES);
      uint8_t* dst = smem + stage * TILE_BYTES;
      if (p.mode == 0) tma_load(dst, &p.full_tile, &full[stage], k, 0);
      else if (p.mode == 1) tma_load(dst, &p.full_tile, &full[stage], k, (int)((blockIdx.x * TILE_ROWS) % p.rows_total));
      else tma_load_mc(dst + rank * slice_rows * TILE_K * 2, &p.slice, &full[stage], k, (int)rank * slice_rows, (uint16_t)((1u << C) - 1));
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 32) {               // consumer: frees the stage as soon as it is full
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < p.iters; ++i) {
      mbar_wait(&full[stage], phase);
      if (p.mode == 2) { for (int c = 0; c < C; ++c) remote_arrive(&empty[stage], c); }
      else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    p.out[blockIdx.x * 2 + 1] = t1;
  }
  if (threadIdx.x == 0) p.out[blockIdx.x * 2] = t0;
  __syncthreads();
  if (p.mode == 2) cluster_sync();
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int K = 4096, ROWS = 128 * 160;
  __nv_bfloat16* buf;
  CK(cudaMalloc(&buf, (size_t)ROWS * K * 2));
  CK(cudaMemset(buf, 0, (size_t)ROWS * K * 2));
  void* sym; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)sym;
  auto make = [&](CUtensorMap* m, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)ROWS}; cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {TILE_K, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  };
  unsigned long long* out; CK(cudaMalloc(&out, 2 * 148 * 8));
  const int smem = STAGES * (TILE_BYTES + W_BYTES) + 2 * STAGES * 8 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  struct Cfg { int mode, csize, grid; const char* name; int nmma = 0, mma_n = 128; };
  const Cfg cfgs[] = {{0, 1, 148, "unicast same tile, 148 CTAs"}, {0, 1, 64, "unicast same tile, 64 CTAs"}, {0, 1, 8, "unicast same tile, 8 CTAs"},
                      {1, 1, 148, "unicast own tile, 148 CTAs"}, {1, 1, 64, "unicast own tile, 64 CTAs"},
                      {2, 2, 148, "multicast x2, 148 CTAs"}, {2, 4, 148, "multicast x4, 148 CTAs"}, {2, 8, 144, "multicast x8, 144 CTAs"}, {2, 8, 128, "multicast x8, 128 CTAs"},
                      {2, 4, 64, "multicast x4, 64 CTAs"},
                      {3, 1, 148, "X same + W own, plain, 148 CTAs"}, {3, 1, 116, "X same + W own, plain, 116 CTAs"},
                      {4, 2, 148, "X+W, CTA pair 2sm loads, 148"}, {5, 2, 148, "X+W, pair, 3-D X map, 148"}, {6, 2, 148, "X+W, pair, 3-D X, W hint, 148"},
                      {6, 2, 116, "X+W, pair, 3-D X, W hint, 116"}, {4, 2, 32, "X+W, CTA pair 2sm loads, 32"}, {4, 2, 148, "pair + 1 MMA N=128 per k-block", 1, 128}, {4, 2, 148, "pair + 4 MMA N=128 per k-block", 4, 128},
                      {4, 2, 148, "pair + 4 MMA N=48 per k-block", 4, 48}, {4, 2, 32, "pair + 4 MMA N=48, 32 CTAs", 4, 48},
                      {7, 2, 148, "elect-warp + 4 MMA N=128", 4, 128}, {7, 2, 148, "elect-warp + 4 MMA N=256(fake)", 4, 256}, {7, 2, 148, "elect-warp + 0 MMA", 0, 128}, {7, 2, 148, "elect-warp + 2 MMA N=128", 2, 128}};
  for (const Cfg& c : cfgs) {
    Params p; p.mode = c.mode; p.nmma = c.nmma; p.mma_n = c.mma_n; p.iters = 4000; p.csize = c.csize; p.kmax = K; p.rows_total = ROWS; p.out = out;
    make(&p.full_tile, TILE_ROWS); make(&p.slice, c.mode == 2 ? TILE_ROWS / c.csize : TILE_ROWS); make(&p.wtile, 64);
    {
      cuuint64_t dims[3] = {(cuuint64_t)K, 256, 8}; cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * 2 * 256};
      cuuint32_t box[3] = {TILE_K, 128, 1}; cuuint32_t es[3] = {1, 1, 1};
      CUresult r = enc(&p.x3d, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode3d failed %d\n", (int)r); exit(1); }
    }
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(c.grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = c.csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = c.csize > 1 ? 1 : 0;
    CK(cudaMemset(out, 0, 2 * 148 * 8));
    for (int rep = 0; rep < 2; ++rep) {
      cudaError_t e = c.mode >= 3 ? cudaLaunchKernelEx(&cfg, probe2_kernel, p) : cudaLaunchKernelEx(&cfg, probe_kernel, p);
      if (e != cudaSuccess) { printf("%-34s launch failed: %s\n", c.name, cudaGetErrorString(e)); cudaGetLastError(); break; }
      CK(cudaDeviceSynchronize());
      if (rep == 0) continue;
      std::vector<unsigned long long> h(2 * 148);
      CK(cudaMemcpy(h.data(), out, 2 * 148 * 8, cudaMemcpyDeviceToHost));
      unsigned long long t0 = ~0ull, t1 = 0; double sum = 0;
      for (int b = 0; b < c.grid; b += (c.mode >= 4 ? 2 : 1)) { if (h[2 * b] < t0) t0 = h[2 * b]; if (h[2 * b + 1] > t1) t1 = h[2 * b + 1]; sum += (double)(h[2 * b + 1] - h[2 * b]); }
      const double us = (t1 - t0) / 1e3, per_cta_us = sum / (c.mode >= 4 ? c.grid / 2 : c.grid) / 1e3;
      printf("%-34s %8.1f us total, %7.1f GB/s received per SM (mean CTA time %8.1f us), %6.2f TB/s aggregate received, %.3f us per 16 KB tile\n", c.name, us,
             (double)p.iters * (c.mode >= 3 ? TILE_BYTES + W_BYTES : TILE_BYTES) / per_cta_us / 1e3, per_cta_us, (double)p.iters * (c.mode >= 3 ? TILE_BYTES + W_BYTES : TILE_BYTES) * c.grid / us / 1e6, per_cta_us / p.iters);
    }
  }
  return 0;
}
