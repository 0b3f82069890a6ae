// Host-side plumbing shared by the C-ABI translation units: error text, launch counting,
// driver entry points for tensor-map encoding (resolved at run time: no libcuda link dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include "../../../include/b200conv.h"

namespace b200 {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();
EncodeIm2colFn encode_im2col_fn();

// conv3x3.cu: halo / shift-GEMM path for stride-1 convolutions (3x3 pad 1; the 4x4 pad 0 space-to-depth stem).
// H, W are the OUTPUT map dimensions.
bool halo_geometry_ok(int H, int W, int Cs, int R, int S, int pad);
bool halo_eligible(int H, int W, int Cs, int Nout, int R, int S, int pad);
int launch_halo(const void* src, const void* wmat, void* out, const void* res, const float* bias, int N, int H, int W,
                int Cs, int Nout, int R, int S, int pad, int dir, int act, double* stats, cudaStream_t stream,
                int window = 0);

bool halo_wgrad_eligible(int H, int W, int C, int K_out, int R, int S, int pad);
int launch_halo_wgrad(const void* x, const void* dy, float* dw, void* workspace, size_t workspace_bytes, int N, int H,
                      int W, int C, int K_out, int R, int S, int pad, cudaStream_t stream, int window = 0);

// conv_pair.cu: EXPERIMENTAL cta_group::2 GEMM for wide 1x1 / stride-1 layers (B200_IGEMM_PAIR=1)
bool pair_eligible(long long M, int C, int Nout);
int launch_pair(const void* a, const void* w, void* out, const void* res, const float* bias, long long M, int C, int Nout,
                int act, double* stats, cudaStream_t stream);

// Kernel launch with programmatic dependent launch (PDL): the grid may be scheduled while its stream predecessor is
// still draining; every kernel launched through here executes pdl_wait() (griddepcontrol.wait, common.cuh) in all
// threads before it touches global memory, which blocks until the predecessor grid has completed and its writes are
// visible.  What overlaps is the launch latency, block scheduling and the smem/TMEM/barrier prologue -- ~560 kernel
// boundaries per ResNet-50 step.  B200_PDL=0 launches without the attribute (pdl_wait() is then a no-op).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface in B200_CHECK_LAUNCH
}

// warps of a split-K reduce block that share the loop over the splits (1..8)
int wgrad_reduce_warps(int splits);

inline CUtensorMapSwizzle swizzle_for_row_bytes(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

#define B200_CHECK_LAUNCH(name)                                                   \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      b200::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return B200_ERR_CUDA;                                                       \
    }                                                                             \
    b200::count_launch();                                                         \
  } while (0)

#define B200_REQUIRE(cond, code, ...)   \
  do {                                  \
    if (!(cond)) {                      \
      b200::set_error(__VA_ARGS__);     \
      return code;                      \
    }                                   \
  } while (0)

}  // namespace b200
