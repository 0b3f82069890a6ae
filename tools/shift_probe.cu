// Probe: do row-shifted (non-1024B-aligned) shared-memory descriptors work with 128B swizzle, and do they need
// the descriptor's base_offset field?  D[128 x N] = A[shift : shift+128, 0:64] * B[0:N, 0:64]^T  (K-major case)
// and the MN-major analogue (A stored [K pixels][128 channels], shift along K).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared -o tools/libshiftprobe.so tools/shift_probe.cu
#include "../convnet/pytorch_b200/csrc/common.cuh"
#include <cuda.h>
using namespace b200;

__device__ __forceinline__ uint64_t desc_bo(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt, uint32_t base_off) {
  uint64_t d = make_smem_desc(saddr, lbo, sbo, lt);
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  return d;
}

// mode 0: K-major A (rows = M), shift rows.  mode 1: MN-major A and B (rows = K pixels), shift rows of K.
__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                       const __grid_constant__ CUtensorMap tmB, int mode, int shift,
                                                       int use_base_off, int N, float* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar, done;
  __shared__ uint32_t tmem_s;
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;             // up to 384 rows x 128 B = 48 KB   (mode 1: two boxes of [K rows x 64 ch])
  uint8_t* sB = smem + 65536;     // up to 256 rows x 128 B
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_s, 256); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_s;
  // simple, precise version: single thread does everything
  if (threadIdx.x == 32) {
    uint32_t bytes = (mode == 0) ? (256 * 128 + N * 128) : (2 * 96 * 128 + (N / 64) * 96 * 128);
    mbar_arrive_expect_tx(&done, bytes);
    if (mode == 0) {
      tma_load_2d(&tmA, &done, sA, 0, 0);
      tma_load_2d(&tmA, &done, sA + 128 * 128, 0, 128);
      tma_load_2d(&tmB, &done, sB, 0, 0);
    } else {
      // A: [K=96 pixel rows] x [128 channels] as two boxes of 64 channels; B likewise N/64 boxes
      tma_load_2d(&tmA, &done, sA, 0, 0);
      tma_load_2d(&tmA, &done, sA + 96 * 128, 64, 0);
      for (int b = 0; b < N / 64; ++b) tma_load_2d(&tmB, &done, sB + b * 96 * 128, b * 64, 0);
    }
    mbar_wait(&done, 0);
    tc_fence_after();
    if (mode == 0) {
      const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
      const uint32_t a0 = smem_u32(sA) + shift * 128;
      const uint32_t bo = use_base_off ? ((a0 >> 7) & 7) : 0;
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = desc_bo(a0 + k * 32, 16, 1024, 2, bo);
        const uint64_t db = make_smem_desc(smem_u32(sB) + k * 32, 16, 1024, 2);
        umma_bf16(tmem, da, db, idesc, k != 0);
      }
    } else {
      // K = 64 pixels starting at pixel row `shift` for A (dy-like operand); B un-shifted from row 0
      const uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
      const uint32_t a0 = smem_u32(sA) + shift * 128;
      const uint32_t bo = use_base_off ? ((a0 >> 7) & 7) : 0;
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = desc_bo(a0 + k * 16 * 128, 96 * 128, 1024, 2, use_base_off ? (((a0 + k * 2048) >> 7) & 7) : 0);
        const uint64_t db = make_smem_desc(smem_u32(sB) + k * 16 * 128, 96 * 128, 1024, 2);
        umma_bf16(tmem, da, db, idesc, k != 0);
      }
      (void)bo;
    }
    umma_commit(&bar);
  }
  __syncwarp();
  __syncthreads();
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(taddr + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
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

// mode 0: a [256][64] bf16, b [N][64] bf16.  mode 1: a [96][128] bf16 (pixel-major), b [96][N] bf16.
extern "C" int shift_probe(const void* a, const void* b, int mode, int shift, int use_base_off, int N, float* out) {
  CUtensorMap tmA, tmB;
  int r;
  if (mode == 0) {
    if ((r = enc2(&tmA, a, 64, 256, 64, 128))) return 100 + r;
    if ((r = enc2(&tmB, b, 64, N, 64, N))) return 200 + r;
  } else {
    if ((r = enc2(&tmA, a, 128, 96, 64, 96))) return 100 + r;
    if ((r = enc2(&tmB, b, N, 96, 64, 96))) return 200 + r;
  }
  cudaFuncSetAttribute((const void*)probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  probe_kernel<<<1, 128, 140 * 1024>>>(tmA, tmB, mode, shift, use_base_off, N, out);
  cudaError_t e = cudaDeviceSynchronize();
  return e == cudaSuccess ? 0 : -(int)e;
}
