// Layout / precision transforms around the tensor-core kernels:
//  - network input NCHW fp32 -> NHWC bf16 (channel padded) or space-to-depth NHWC bf16 for the 7x7/s2 stem
//    (replaces inputs.to(device, dtype), trainer.py:116-117, plus the relayout cuDNN does internally)
//  - weight relayouts: [K][T][C] -> [C][T][K] for dgrad, 7x7 stem <-> 4x4 space-to-depth form
//  - fp32 -> bf16 cast of flat arrays
#include "common.cuh"
#include "host.h"

namespace b200 {

__global__ void __launch_bounds__(256) input_prep_kernel(const float* __restrict__ x, int N, int C, int H, int W,
                                                         int Cpad, int mode, __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  // one thread per output pixel; Cpad is a multiple of 8
  const int brd = mode == 2 ? 2 : 0;                       // low border of the padded space-to-depth layout
  const int OH = mode == 0 ? H : H / 2 + (mode == 2 ? 3 : 0), OW = mode == 0 ? W : W / 2 + (mode == 2 ? 3 : 0);
  const long long total = (long long)N * OH * OW;
  const long long plane = (long long)H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % OW) - brd;
    const int i = (int)((idx / OW) % OH) - brd;
    const int n = (int)(idx / ((long long)OW * OH));
    __nv_bfloat16* o = out + idx * Cpad;
    const bool inside = mode != 2 || (i >= 0 && j >= 0 && i < H / 2 && j < W / 2);
    const float* xi = x + (long long)n * C * plane;
    for (int c0 = 0; c0 < Cpad; c0 += 8) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int ch = c0 + e;
        float val = 0.f;
        if (mode == 0) {
          if (ch < C) val = __ldg(xi + (long long)ch * plane + (long long)i * W + j);
        } else {
          const int sub = ch / C, c = ch - sub * C;  // sub = dy*2+dx
          if (sub < 4 && inside) {
            const int dy = sub >> 1, dx = sub & 1;
            val = __ldg(xi + (long long)c * plane + (long long)(2 * i + dy) * W + (2 * j + dx));
          }
        }
        f[e] = val;
      }
      uint4 u;
      u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
      u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(o + c0) = u;
    }
  }
}

// uint8 NHWC images (what a decoder produces) -> the same bf16 layouts, normalised on the fly:
// value = u8 * scale[c] + bias[c]  (scale = 1 / (255 * std), bias = -mean / std: ToTensor + Normalize of the reference's
// preprocess.py:20-24).  4x fewer host->device bytes than the fp32 NCHW batch and no separate normalisation pass.
struct U8Norm { float scale[4], bias[4]; };
__global__ void __launch_bounds__(256) input_prep_u8_kernel(const uint8_t* __restrict__ x, int N, int C, int H, int W,
                                                            int Cpad, int mode, U8Norm nm,
                                                            __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  const int brd = mode == 2 ? 2 : 0;
  const int OH = mode == 0 ? H : H / 2 + (mode == 2 ? 3 : 0), OW = mode == 0 ? W : W / 2 + (mode == 2 ? 3 : 0);
  const long long total = (long long)N * OH * OW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % OW) - brd;
    const int i = (int)((idx / OW) % OH) - brd;
    const int n = (int)(idx / ((long long)OW * OH));
    __nv_bfloat16* o = out + idx * Cpad;
    const bool inside = mode != 2 || (i >= 0 && j >= 0 && i < H / 2 && j < W / 2);
    const uint8_t* xi = x + (long long)n * H * W * C;
    for (int c0 = 0; c0 < Cpad; c0 += 8) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int ch = c0 + e;
        float val = 0.f;
        if (mode == 0) {
          if (ch < C) val = fmaf((float)__ldg(xi + ((long long)i * W + j) * C + ch), nm.scale[ch], nm.bias[ch]);
        } else {
          const int sub = ch / C, c = ch - sub * C;  // sub = dy*2+dx
          if (sub < 4 && inside) {
            const int dy = sub >> 1, dx = sub & 1;
            val = fmaf((float)__ldg(xi + ((long long)(2 * i + dy) * W + (2 * j + dx)) * C + c), nm.scale[c], nm.bias[c]);
          }
        }
        f[e] = val;
      }
      uint4 u;
      u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
      u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(o + c0) = u;
    }
  }
}

// bf16 [K][T][C] -> [C][T][K]; one 32x32 tile per block, blockIdx.z = tap
__global__ void __launch_bounds__(256) weight_transpose_kernel(const __nv_bfloat16* __restrict__ src,
                                                               __nv_bfloat16* __restrict__ dst, int K, int T, int C) {
  pdl_wait();
  __shared__ __nv_bfloat16 tile[32][33];
  const int t = blockIdx.z;
  const int c0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, c = c0 + tx;
    tile[r][tx] = (k < K && c < C) ? src[((long long)k * T + t) * C + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, k = k0 + tx;
    if (c < C && k < K) dst[((long long)c * T + t) * K + k] = tile[tx][r];
  }
}

// multi-tensor version: one launch transposes every conv weight of the network (53 launches -> 1 per step).
// jobs[j] = {src_off, dst_off, K, T, C, tile_start} in elements of the two arenas; a block owns one 32x32 tile and
// finds its job by binary search over tile_start.
__global__ void __launch_bounds__(256) weight_transpose_batched_kernel(const __nv_bfloat16* __restrict__ src_base,
                                                                       __nv_bfloat16* __restrict__ dst_base,
                                                                       const int* __restrict__ jobs, int njobs) {
  pdl_wait();
  __shared__ __nv_bfloat16 tile[32][33];
  int lo = 0, hi = njobs - 1;
  const int g = blockIdx.x;
  while (lo < hi) {   // last job with tile_start <= g
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(jobs + mid * 6 + 5) <= g) lo = mid; else hi = mid - 1;
  }
  const int* jb = jobs + lo * 6;
  const __nv_bfloat16* src = src_base + __ldg(jb + 0);
  __nv_bfloat16* dst = dst_base + __ldg(jb + 1);
  const int K = __ldg(jb + 2), T = __ldg(jb + 3), C = __ldg(jb + 4);
  const int local = g - __ldg(jb + 5);
  const int tiles_c = (C + 31) >> 5, tiles_k = (K + 31) >> 5;
  const int t = local / (tiles_c * tiles_k);
  const int rem = local - t * tiles_c * tiles_k;
  const int k0 = (rem / tiles_c) * 32, c0 = (rem % tiles_c) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, c = c0 + tx;
    tile[r][tx] = (k < K && c < C) ? src[((long long)k * T + t) * C + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, k = k0 + tx;
    if (c < C && k < K) dst[((long long)c * T + t) * K + k] = tile[tx][r];
  }
}

// stem: w fp32 [K][7][7][C] -> bf16 [K][16][Cpad], tap (ah,aw) in 4x4, channel (bh*2+bw)*C + c,
// r = 2*ah + bh - 1, s = 2*aw + bw - 1 (out-of-range -> 0)
__global__ void stem_w_to_s2d_kernel(const float* __restrict__ w, int K, int C, int Cpad,
                                     __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  const int total = K * 16 * Cpad;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int ch = idx % Cpad;
    const int tap = (idx / Cpad) % 16;
    const int k = idx / (Cpad * 16);
    const int ah = tap >> 2, aw = tap & 3;
    float val = 0.f;
    const int sub = ch / C, c = ch - sub * C;
    if (sub < 4) {
      const int r = 2 * ah + (sub >> 1) - 1, s = 2 * aw + (sub & 1) - 1;
      if (r >= 0 && r < 7 && s >= 0 && s < 7) val = w[(((long long)k * 7 + r) * 7 + s) * C + c];
    }
    out[idx] = __float2bfloat16(val);
  }
}
// reverse gather for the gradient: dw[K][7][7][C] += dw_s2d[K][16][Cpad]
__global__ void stem_wgrad_from_s2d_kernel(const float* __restrict__ dws, int K, int C, int Cpad,
                                           float* __restrict__ dw) {
  pdl_wait();
  const int total = K * 49 * C;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c = idx % C;
    const int s = (idx / C) % 7;
    const int r = (idx / (C * 7)) % 7;
    const int k = idx / (C * 49);
    const int ah = (r + 1) >> 1, bh = (r + 1) & 1, aw = (s + 1) >> 1, bw = (s + 1) & 1;
    dw[idx] += dws[((long long)k * 16 + ah * 4 + aw) * Cpad + (bh * 2 + bw) * C + c];
  }
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src,
                                                            __nv_bfloat16* __restrict__ dst, long long n) {
  pdl_wait();
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(src) + i);
    uint2 u;
    u.x = pack_bf16x2(f.x, f.y);
    u.y = pack_bf16x2(f.z, f.w);
    reinterpret_cast<uint2*>(dst)[i] = u;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __float2bfloat16(src[i]);
}

static inline int grid_cap(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_input_prep(const float* x, int N, int C, int H, int W, int Cpad, int mode, void* out,
                               b200_stream_t stream) {
  B200_REQUIRE(x && out && N > 0 && C > 0 && H > 0 && W > 0, B200_ERR_INVALID, "input_prep: bad argument");
  B200_REQUIRE(Cpad % 8 == 0, B200_ERR_UNSUPPORTED, "input_prep: Cpad=%d must be a multiple of 8", Cpad);
  if (mode == 0) {
    B200_REQUIRE(Cpad >= C, B200_ERR_INVALID, "input_prep: Cpad < C");
  } else if (mode == 1 || mode == 2) {
    B200_REQUIRE(H % 2 == 0 && W % 2 == 0 && Cpad >= 4 * C, B200_ERR_UNSUPPORTED,
                 "input_prep: space-to-depth needs even H,W and Cpad >= 4C");
  } else {
    B200_REQUIRE(false, B200_ERR_INVALID, "input_prep: unknown mode %d", mode);
  }
  const long long total = (long long)N * (mode == 0 ? (long long)H * W : (long long)(H / 2 + (mode == 2 ? 3 : 0)) * (W / 2 + (mode == 2 ? 3 : 0)));
  b200::launch(input_prep_kernel, grid_cap(total, 256), 256, 0, (cudaStream_t)stream, x, N, C, H, W, Cpad, mode,
                                                                          (__nv_bfloat16*)out);
  B200_CHECK_LAUNCH("input_prep_kernel");
  return B200_OK;
}

extern "C" int b200_input_prep_u8(const uint8_t* x_nhwc, int N, int C, int H, int W, int Cpad, int mode,
                                  const float* scale_host, const float* bias_host, void* out, b200_stream_t stream) {
  B200_REQUIRE(x_nhwc && out && scale_host && bias_host && N > 0 && C > 0 && C <= 4 && H > 0 && W > 0,
               B200_ERR_INVALID, "input_prep_u8: bad argument (C must be 1..4)");
  B200_REQUIRE(Cpad % 8 == 0 && Cpad >= C, B200_ERR_INVALID, "input_prep_u8: Cpad must be a multiple of 8 and >= C");
  B200_REQUIRE(mode >= 0 && mode <= 2, B200_ERR_INVALID, "input_prep_u8: unknown mode %d", mode);
  if (mode != 0)
    B200_REQUIRE(H % 2 == 0 && W % 2 == 0 && 4 * C <= Cpad, B200_ERR_UNSUPPORTED,
                 "input_prep_u8: space-to-depth needs even H, W and 4*C <= Cpad");
  U8Norm nm;
  for (int c = 0; c < 4; ++c) { nm.scale[c] = c < C ? scale_host[c] : 0.f; nm.bias[c] = c < C ? bias_host[c] : 0.f; }
  const int OH = mode == 0 ? H : H / 2 + (mode == 2 ? 3 : 0), OW = mode == 0 ? W : W / 2 + (mode == 2 ? 3 : 0);
  const long long total = (long long)N * OH * OW;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  b200::launch(input_prep_u8_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, x_nhwc, N, C, H, W, Cpad, mode, nm,
                                                                       (__nv_bfloat16*)out);
  B200_CHECK_LAUNCH("input_prep_u8_kernel");
  return B200_OK;
}

extern "C" int b200_weight_transpose(const void* src, void* dst, int K, int T, int C, b200_stream_t stream) {
  B200_REQUIRE(src && dst && K > 0 && T > 0 && C > 0, B200_ERR_INVALID, "weight_transpose: bad argument");
  dim3 grid((C + 31) / 32, (K + 31) / 32, T);
  b200::launch(weight_transpose_kernel, grid, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, K, T,
                                                                C);
  B200_CHECK_LAUNCH("weight_transpose_kernel");
  return B200_OK;
}

extern "C" int b200_weight_transpose_batched(const void* src_base, void* dst_base, const int* jobs, int njobs,
                                             int total_tiles, b200_stream_t stream) {
  B200_REQUIRE(src_base && dst_base && jobs && njobs > 0 && total_tiles > 0, B200_ERR_INVALID,
               "weight_transpose_batched: bad argument");
  b200::launch(weight_transpose_batched_kernel, total_tiles, 256, 0, (cudaStream_t)stream,
      (const __nv_bfloat16*)src_base, (__nv_bfloat16*)dst_base, jobs, njobs);
  B200_CHECK_LAUNCH("weight_transpose_batched_kernel");
  return B200_OK;
}

extern "C" int b200_stem_weight_to_s2d(const float* w, int K, int C, int Cpad, void* w_s2d, b200_stream_t stream) {
  B200_REQUIRE(w && w_s2d && K > 0 && C > 0 && Cpad >= 4 * C, B200_ERR_INVALID, "stem_weight_to_s2d: bad argument");
  b200::launch(stem_w_to_s2d_kernel, (K * 16 * Cpad + 255) / 256, 256, 0, (cudaStream_t)stream, w, K, C, Cpad,
                                                                                    (__nv_bfloat16*)w_s2d);
  B200_CHECK_LAUNCH("stem_w_to_s2d_kernel");
  return B200_OK;
}

extern "C" int b200_stem_wgrad_from_s2d(const float* dw_s2d, int K, int C, int Cpad, float* dw, b200_stream_t stream) {
  B200_REQUIRE(dw_s2d && dw && K > 0 && C > 0 && Cpad >= 4 * C, B200_ERR_INVALID, "stem_wgrad_from_s2d: bad argument");
  b200::launch(stem_wgrad_from_s2d_kernel, (K * 49 * C + 255) / 256, 256, 0, (cudaStream_t)stream, dw_s2d, K, C, Cpad, dw);
  B200_CHECK_LAUNCH("stem_wgrad_from_s2d_kernel");
  return B200_OK;
}

extern "C" int b200_cast_f32_to_bf16(const float* src, void* dst, long long n, b200_stream_t stream) {
  B200_REQUIRE(src && dst && n >= 0, B200_ERR_INVALID, "cast_f32_to_bf16: bad argument");
  if (n == 0) return B200_OK;
  b200::launch(cast_f32_bf16_kernel, grid_cap((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream, src, (__nv_bfloat16*)dst, n);
  B200_CHECK_LAUNCH("cast_f32_bf16_kernel");
  return B200_OK;
}

// ---- grouped convolution support (ResNeXt, nn.Conv2d(groups=32), models/resnext.py:10-16) -------------------------
// A grouped convolution with C == K is block diagonal at ANY granularity that is a multiple of the group width: output
// channels [W*b, W*b+W) only read input channels of the same window.  The tensor-core kernels run it as K/W independent
// W-wide diagonal blocks ("window" mode of the conv entry points, W = 64 for fprop/dgrad, 128 for the k-tile of wgrad),
// so the weight operand is packed per window:
//   pack      : w_g fp32 [K][T][C/g] -> bf16 [K][T][W],   out[k][t][cl] = w_g[k][t][c - first(k)] if channel
//               c = W*(k/W) + cl lies in the group of k, else 0                                   (fprop operand)
//   transposed: bf16 [C][T][W],  out[c][t][kl] = the same weight seen from input channel c, k = W*(c/W) + kl (dgrad operand)
//   unpack    : dw_g[k][t][cl] += dw_win[k][t][first(k) + cl - W*(k/W)]   from the windowed fp32 gradient [K][T][W]
// W == C reproduces the dense block-diagonal expansion.
namespace b200 {
__global__ void __launch_bounds__(256) group_pack_kernel(const float* __restrict__ wg, int K, int T, int C, int groups,
                                                         int Wd, int transpose, __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  const int cg = C / groups, kg = K / groups;
  const int rows = transpose ? C : K;
  const long long total = (long long)rows * T * Wd;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(idx % Wd);
    const int t = (int)((idx / Wd) % T);
    const int row = (int)(idx / ((long long)Wd * T));
    int k, c;
    if (transpose) { c = row; k = (c / Wd) * Wd + l; } else { k = row; c = (k / Wd) * Wd + l; }
    float v = 0.f;
    if (k < K && c < C) {
      const int grp = k / kg;
      if (c / cg == grp) v = wg[((long long)k * T + t) * cg + (c - grp * cg)];
    }
    out[idx] = __float2bfloat16(v);
  }
}
__global__ void __launch_bounds__(256) group_unpack_kernel(const float* __restrict__ dw_win, int K, int T, int C,
                                                           int groups, int Wd, float* __restrict__ dwg) {
  pdl_wait();
  const int cg = C / groups, kg = K / groups;
  const long long total = (long long)K * T * cg;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cl = (int)(idx % cg);
    const int t = (int)((idx / cg) % T);
    const int k = (int)(idx / ((long long)cg * T));
    const int c = (k / kg) * cg + cl;
    dwg[idx] += dw_win[((long long)k * T + t) * Wd + (c - (k / Wd) * Wd)];
  }
}
}  // namespace b200

extern "C" int b200_group_weight_pack(const float* w_grouped, int K, int T, int C, int groups, int window, int transpose,
                                      void* out_bf16, b200_stream_t stream) {
  B200_REQUIRE(w_grouped && out_bf16 && K > 0 && T > 0 && C > 0 && groups > 0 && C % groups == 0 && K % groups == 0,
               B200_ERR_INVALID, "group_weight_pack: bad argument");
  B200_REQUIRE(window > 0 && window % (C / groups) == 0 && window % (K / groups) == 0 && (window == C || (C == K && C % window == 0)),
               B200_ERR_UNSUPPORTED, "group_weight_pack: window %d must be a multiple of the group width and divide C == K",
               window);
  const long long total = (long long)(transpose ? C : K) * T * window;
  b200::launch(b200::group_pack_kernel, b200::grid_cap(total, 256), 256, 0, (cudaStream_t)stream,
      w_grouped, K, T, C, groups, window, transpose, (__nv_bfloat16*)out_bf16);
  B200_CHECK_LAUNCH("group_pack_kernel");
  return B200_OK;
}

extern "C" int b200_group_wgrad_unpack(const float* dw_win, int K, int T, int C, int groups, int window, float* dw_grouped,
                                       b200_stream_t stream) {
  B200_REQUIRE(dw_win && dw_grouped && K > 0 && T > 0 && C > 0 && groups > 0 && C % groups == 0 && K % groups == 0 &&
                   window > 0 && window % (C / groups) == 0,
               B200_ERR_INVALID, "group_wgrad_unpack: bad argument");
  const long long total = (long long)K * T * (C / groups);
  b200::launch(b200::group_unpack_kernel, b200::grid_cap(total, 256), 256, 0, (cudaStream_t)stream, dw_win, K, T, C, groups, window,
                                                                                       dw_grouped);
  B200_CHECK_LAUNCH("group_unpack_kernel");
  return B200_OK;
}
