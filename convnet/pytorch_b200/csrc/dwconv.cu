// Depthwise RxS convolution (groups == C == K) for NHWC bf16: CUDA-core, HBM-bound, 128-bit vectorised
// fprop / dgrad / wgrad.  Arithmetic intensity is ~R*S FLOP per element, far below the tensor-core
// ridge, so these are written as bandwidth kernels (no tensor cores, no im2col).
// Replaces the depthwise nn.Conv2d of the reference's inverted residual (models/mobilenet_v2.py:57-58)
// and its autograd backward.  Weight layout: [R*S][C] (bf16 for fprop/dgrad, fp32 gradient).
#include "common.cuh"
#include "host.h"

namespace b200 {

__device__ __forceinline__ void dld8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void dst8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

__global__ void __launch_bounds__(256) dw_fprop_kernel(const __nv_bfloat16* __restrict__ x,
                                                       const __nv_bfloat16* __restrict__ w, b200_conv_desc d,
                                                       __nv_bfloat16* __restrict__ y) {
  pdl_wait();
  const int cv = d.C >> 3;
  const long long total = (long long)d.N * d.P * d.Q * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % cv);
    long long pix = idx / cv;
    const int q = (int)(pix % d.Q); pix /= d.Q;
    const int p = (int)(pix % d.P);
    const int n = (int)(pix / d.P);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int r = 0; r < d.R; ++r) {
      const int h = p * d.stride - d.pad_h + r;
      if (h < 0 || h >= d.H) continue;
      for (int s = 0; s < d.S; ++s) {
        const int ww = q * d.stride - d.pad_w + s;
        if (ww < 0 || ww >= d.W) continue;
        float xf[8], wf[8];
        dld8(x + (((long long)n * d.H + h) * d.W + ww) * d.C + v * 8, xf);
        dld8(w + (long long)(r * d.S + s) * d.C + v * 8, wf);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(xf[i], wf[i], acc[i]);
      }
    }
    dst8(y + idx * 8, acc);
  }
}

__global__ void __launch_bounds__(256) dw_dgrad_kernel(const __nv_bfloat16* __restrict__ dy,
                                                       const __nv_bfloat16* __restrict__ w, b200_conv_desc d,
                                                       __nv_bfloat16* __restrict__ dx) {
  pdl_wait();
  const int cv = d.C >> 3;
  const long long total = (long long)d.N * d.H * d.W * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % cv);
    long long pix = idx / cv;
    const int ww = (int)(pix % d.W); pix /= d.W;
    const int h = (int)(pix % d.H);
    const int n = (int)(pix / d.H);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int r = 0; r < d.R; ++r) {
      const int th = h + d.pad_h - r;
      if (th < 0 || th % d.stride != 0) continue;
      const int p = th / d.stride;
      if (p >= d.P) continue;
      for (int s = 0; s < d.S; ++s) {
        const int tw = ww + d.pad_w - s;
        if (tw < 0 || tw % d.stride != 0) continue;
        const int q = tw / d.stride;
        if (q >= d.Q) continue;
        float gf[8], wf[8];
        dld8(dy + (((long long)n * d.P + p) * d.Q + q) * d.C + v * 8, gf);
        dld8(w + (long long)(r * d.S + s) * d.C + v * 8, wf);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(gf[i], wf[i], acc[i]);
      }
    }
    dst8(dx + idx * 8, acc);
  }
}

constexpr int kDwThreads = 256;
constexpr int kDwMaxBlocks = 592;
constexpr int kDwMaxTaps = 9;

// partial[block][tap][C]
__global__ void __launch_bounds__(kDwThreads) dw_wgrad_partial_kernel(const __nv_bfloat16* __restrict__ x,
                                                                      const __nv_bfloat16* __restrict__ dy,
                                                                      b200_conv_desc d, int cv, int rows_per_iter,
                                                                      float* __restrict__ partial) {
  pdl_wait();
  __shared__ float red[kDwThreads][9];
  const int t = threadIdx.x;
  const bool active = t < rows_per_iter * cv;
  const int r0 = t / cv, v = t - r0 * cv;
  const long long M = (long long)d.N * d.P * d.Q;
  const long long rows_per_block = (M + gridDim.x - 1) / gridDim.x;
  const long long row_begin = blockIdx.x * rows_per_block;
  const long long row_end = min(M, row_begin + rows_per_block);
  const int ntaps = d.R * d.S;
  float acc[kDwMaxTaps][8];
#pragma unroll
  for (int a = 0; a < kDwMaxTaps; ++a)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[a][i] = 0.f;
  if (active) {
    for (long long row = row_begin + r0; row < row_end; row += rows_per_iter) {
      const int q = (int)(row % d.Q);
      const int p = (int)((row / d.Q) % d.P);
      const int n = (int)(row / ((long long)d.Q * d.P));
      float gf[8];
      dld8(dy + row * d.C + v * 8, gf);
#pragma unroll
      for (int a = 0; a < kDwMaxTaps; ++a) {
        if (a >= ntaps) break;
        const int r = a / d.S, s = a - r * d.S;
        const int h = p * d.stride - d.pad_h + r, ww = q * d.stride - d.pad_w + s;
        if (h < 0 || h >= d.H || ww < 0 || ww >= d.W) continue;
        float xf[8];
        dld8(x + (((long long)n * d.H + h) * d.W + ww) * d.C + v * 8, xf);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[a][i] = fmaf(gf[i], xf[i], acc[a][i]);
      }
    }
  }
  float* out = partial + (long long)blockIdx.x * ntaps * d.C;
#pragma unroll
  for (int a = 0; a < kDwMaxTaps; ++a) {
    if (a >= ntaps) break;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[t][i] = acc[a][i];
    __syncthreads();
    for (int c = t; c < d.C; c += kDwThreads) {
      const int vv = c >> 3, e = c & 7;
      float s = 0.f;
      for (int r = 0; r < rows_per_iter; ++r) s += red[r * cv + vv][e];
      out[(long long)a * d.C + c] = s;
    }
  }
}

__global__ void dw_wgrad_final_kernel(const float* __restrict__ partial, int nblocks, int n, float* __restrict__ dw) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += (double)partial[(long long)b * n + i];
  dw[i] += (float)s;
}

static int dw_check(const b200_conv_desc* d, const char* who) {
  B200_REQUIRE(d && d->C == d->K && d->C % 8 == 0 && d->C <= 2048, B200_ERR_UNSUPPORTED,
               "%s: depthwise needs C == K, C %% 8 == 0, C <= 2048", who);
  B200_REQUIRE(d->R * d->S <= kDwMaxTaps && d->stride >= 1, B200_ERR_UNSUPPORTED, "%s: filter too large", who);
  return B200_OK;
}
static inline int dw_grid(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_dwconv_fprop(const b200_conv_desc* d, const void* x, const void* w, void* y, b200_stream_t stream) {
  int rc = dw_check(d, "dwconv_fprop");
  if (rc) return rc;
  B200_REQUIRE(x && w && y, B200_ERR_INVALID, "dwconv_fprop: null pointer");
  const long long total = (long long)d->N * d->P * d->Q * (d->C / 8);
  b200::launch(dw_fprop_kernel, dw_grid(total), 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w,
                                                                  *d, (__nv_bfloat16*)y);
  B200_CHECK_LAUNCH("dw_fprop_kernel");
  return B200_OK;
}

extern "C" int b200_dwconv_dgrad(const b200_conv_desc* d, const void* dy, const void* w, void* dx,
                                 b200_stream_t stream) {
  int rc = dw_check(d, "dwconv_dgrad");
  if (rc) return rc;
  B200_REQUIRE(dy && w && dx, B200_ERR_INVALID, "dwconv_dgrad: null pointer");
  const long long total = (long long)d->N * d->H * d->W * (d->C / 8);
  b200::launch(dw_dgrad_kernel, dw_grid(total), 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)w,
                                                                  *d, (__nv_bfloat16*)dx);
  B200_CHECK_LAUNCH("dw_dgrad_kernel");
  return B200_OK;
}

extern "C" int b200_dwconv_wgrad(const b200_conv_desc* d, const void* x, const void* dy, float* dw, float* workspace,
                                 size_t workspace_bytes, b200_stream_t stream) {
  int rc = dw_check(d, "dwconv_wgrad");
  if (rc) return rc;
  B200_REQUIRE(x && dy && dw && workspace, B200_ERR_INVALID, "dwconv_wgrad: null pointer");
  const int cv = d->C / 8;
  const int rows_per_iter = kDwThreads / cv;
  const long long M = (long long)d->N * d->P * d->Q;
  long long blocks = (M + rows_per_iter * 8 - 1) / (rows_per_iter * 8);
  if (blocks > kDwMaxBlocks) blocks = kDwMaxBlocks;
  if (blocks < 1) blocks = 1;
  const int n = d->R * d->S * d->C;
  B200_REQUIRE(workspace_bytes >= (size_t)blocks * n * sizeof(float), B200_ERR_INVALID,
               "dwconv_wgrad: workspace too small (%zu < %zu)", workspace_bytes, (size_t)blocks * n * sizeof(float));
  b200::launch(dw_wgrad_partial_kernel, (int)blocks, kDwThreads, 0, (cudaStream_t)stream,
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, *d, cv, rows_per_iter, workspace);
  B200_CHECK_LAUNCH("dw_wgrad_partial_kernel");
  b200::launch(dw_wgrad_final_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, workspace, (int)blocks, n, dw);
  B200_CHECK_LAUNCH("dw_wgrad_final_kernel");
  return B200_OK;
}
