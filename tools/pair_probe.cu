// Probe for the CTA-pair (cta_group::2) tensor-core path, to be run at the start of the next round:
//   D[256 x N] = A[256 x 64] * B[N x 64]^T,  one cluster of two CTAs.
// CTA r holds rows [128r, 128r+128) of A and rows [r*N/2, (r+1)*N/2) of B in ITS shared memory (same offsets in both
// CTAs); the leader (rank 0) issues tcgen05.mma.cta_group::2 with M = 256; each CTA reads its own 128 TMEM lanes.
// What it validates before the igemm kernel is rewritten around it: descriptor / instruction-descriptor conventions
// for M = 256, that B is consumed half from each CTA, the multicast commit, and alloc/dealloc with cta_group::2.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared -o tools/libpairprobe.so tools/pair_probe.cu
#include "../convnet/pytorch_b200/csrc/common.cuh"
#include <cuda.h>
using namespace b200;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int N, float* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t load_bar, done_bar;
  __shared__ uint32_t tmem_s;
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;            // 128 rows x 128 B
  uint8_t* sB = smem + 16384;    // N/2 rows x 128 B
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(&load_bar, 1); mbar_init(&done_bar, 1); fence_mbar_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // both CTAs: barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  if (threadIdx.x == 32) {
    mbar_arrive_expect_tx(&load_bar, 128 * 128 + (N / 2) * 128);
    tma_load_2d(&tmA, &load_bar, sA, 0, rank * 128);
    tma_load_2d(&tmB, &load_bar, sB, 0, rank * (N / 2));
    mbar_wait(&load_bar, 0);
  }
  __syncthreads();
  cluster_sync_all();            // both halves of A and B are in shared memory
  tc_fence_after();
  if (rank == 0 && threadIdx.x == 32) {
    const uint32_t idesc = make_idesc_bf16(256, N, 0, 0);
    for (int k = 0; k < 4; ++k) {
      const uint64_t da = make_smem_desc(smem_u32(sA) + k * 32, 16, 1024, 2);
      const uint64_t db = make_smem_desc(smem_u32(sB) + k * 32, 16, 1024, 2);
      umma_bf16_pair(tmem, da, db, idesc, k != 0);
    }
    umma_commit_pair(&done_bar);   // arrives on done_bar of BOTH CTAs
  }
  __syncwarp();
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(taddr + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(rank * 128 + warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // nobody deallocates while the peer still reads / the MMA still runs
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int enc2(CUtensorMap* tm, const void* base, int d0, int d1, int b0, int b1) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
  cuuint64_t strides[1] = {(cuuint64_t)d0 * 2};
  cuuint32_t box[2] = {(cuuint32_t)b0, (cuuint32_t)b1};
  cuuint32_t es[2] = {1, 1};
  return (int)((EncodeTiledFn)fn)(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// a [256][64] bf16, b [N][64] bf16, out [256][N] fp32;  N in {64, 128, 256}
extern "C" int pair_probe(const void* a, const void* b, int N, float* out) {
  CUtensorMap tmA, tmB;
  int r;
  if ((r = enc2(&tmA, a, 64, 256, 64, 128))) return 100 + r;
  if ((r = enc2(&tmB, b, 64, N, 64, N / 2))) return 200 + r;
  cudaFuncSetAttribute((const void*)pair_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  pair_probe_kernel<<<2, 128, 64 * 1024>>>(tmA, tmB, N, out);
  cudaError_t e = cudaDeviceSynchronize();
  return e == cudaSuccess ? 0 : -(int)e;
}
