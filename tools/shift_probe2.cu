// Probe 2: row-shifted UMMA shared-memory descriptors for NARROW rows (32 B / 64 B / 128 B swizzle modes).
//  mode 0  K-major A, rows of `rb` bytes (K = rb/2 elements), start address shifted by `shift` rows:
//          D[128 x N] = A[shift : shift+128, :] * B[0:N, :]^T
//  mode 1  MN-major B operand with rows of `rb` bytes (N = rb/2), rows = K (pixels), shifted along K:
//          D[128 x N] = sum_{p<64} A[p, 0:128]^T-ish (A MN-major, 128B swizzle, unshifted) * B[shift+p, 0:N]
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared -o tools/libshiftprobe2.so tools/shift_probe2.cu
#include "../convnet/pytorch_b200/csrc/common.cuh"
#include <cuda.h>
using namespace b200;

__global__ void __launch_bounds__(128, 1) probe2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                        const __grid_constant__ CUtensorMap tmB, int mode, int rb,
                                                        int shift, int N, float* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar, done;
  __shared__ uint32_t tmem_s;
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 65536;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_s, 256); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_s;
  const uint32_t lt = layout_type_for_row_bytes(rb);
  if (threadIdx.x == 32) {
    if (mode == 0) {
      mbar_arrive_expect_tx(&done, 256 * rb + N * rb);
      tma_load_2d(&tmA, &done, sA, 0, 0);          // [256 rows][rb bytes]
      tma_load_2d(&tmB, &done, sB, 0, 0);          // [N rows][rb bytes]
      mbar_wait(&done, 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
      const uint32_t a0 = smem_u32(sA) + shift * rb;
      for (int k = 0; k < rb / 32; ++k) {
        const uint64_t da = make_smem_desc(a0 + k * 32, 16, 8 * rb, lt);
        const uint64_t db = make_smem_desc(smem_u32(sB) + k * 32, 16, 8 * rb, lt);
        umma_bf16(tmem, da, db, idesc, k != 0);
      }
    } else {
      mbar_arrive_expect_tx(&done, 2 * 96 * 128 + 160 * rb);
      tma_load_2d(&tmA, &done, sA, 0, 0);
      tma_load_2d(&tmA, &done, sA + 96 * 128, 64, 0);
      tma_load_2d(&tmB, &done, sB, 0, 0);          // [160 pixel rows][rb bytes]
      mbar_wait(&done, 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = make_smem_desc(smem_u32(sA) + k * 16 * 128, 96 * 128, 1024, 2);
        const uint64_t db = make_smem_desc(smem_u32(sB) + (shift + k * 16) * rb, 160 * rb, 8 * rb, lt);
        umma_bf16(tmem, da, db, idesc, k != 0);
      }
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
  const int rbytes = b0 * 2;
  CUtensorMapSwizzle sw = rbytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                        : (rbytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  return (int)((EncodeTiledFn)fn)(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// mode 0: a [256][rb/2], b [N][rb/2].   mode 1: a [96][128], b [160][rb/2] (N = rb/2).
extern "C" int shift_probe2(const void* a, const void* b, int mode, int rb, int shift, int N, float* out) {
  CUtensorMap tmA, tmB;
  int r;
  const int ch = rb / 2;
  if (mode == 0) {
    if ((r = enc2(&tmA, a, ch, 256, ch, 256))) return 100 + r;
    if ((r = enc2(&tmB, b, ch, N, ch, N))) return 200 + r;
  } else {
    if ((r = enc2(&tmA, a, 128, 96, 64, 96))) return 100 + r;
    if ((r = enc2(&tmB, b, ch, 160, ch, 160))) return 200 + r;
  }
  cudaFuncSetAttribute((const void*)probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  probe2_kernel<<<1, 128, 140 * 1024>>>(tmA, tmB, mode, rb, shift, N, out);
  cudaError_t e = cudaDeviceSynchronize();
  return e == cudaSuccess ? 0 : -(int)e;
}
