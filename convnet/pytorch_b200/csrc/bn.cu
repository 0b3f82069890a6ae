// Batch-norm statistics / apply / backward for NHWC bf16 activations: HBM-bound, 128-bit vectorised,
// ReLU/ReLU6 and the residual add fused.  Replaces nn.BatchNorm2d (+ nn.ReLU, + `out += residual`) of the
// reference (models/resnet.py:88-91,115-116,128-134,162-163; models/mobilenet_v2.py:50-63), train and eval.
//
// Thread mapping shared by all kernels: a block owns a contiguous range of rows (pixels); thread t is
// (row_in_iter = t / cv, vec = t % cv) with cv = C/8 channel vectors, so every iteration of a block touches
// one contiguous span of memory and per-channel coefficients are loaded once per thread.
//
// Forward statistics normally come from the convolution epilogue (fp64 atomics into 16 replica rows of the
// workspace, finished by bn_finalize_kernel); bn_stats_kernel is the stand-alone version for outputs the conv kernel
// cannot cover (single kernel: the last block, elected by a ticket counter, finalises and re-zeroes the
// accumulators).  The backward reduction writes one partial row per block and a second tiny kernel sums them in a
// fixed order (no atomics: same-address fp64 atomics at the end of ~600 blocks cost more than the extra launch).
// The workspace must be zero before first use and must not be shared by concurrent streams.
#include "common.cuh"
#include "host.h"
#include <stdlib.h>

namespace b200 {

constexpr int kBnThreads = 256;
constexpr int kBnMaxC = 2048;
constexpr int kBnMaxBlocks = 1184;  // 8 per SM on 148 SMs
constexpr int kReplicas = 16;         // accumulator copies: spreads same-address fp64 atomics over 16 lines
constexpr int kAccumFloats = kReplicas * 2 * kBnMaxC * 2 + 64;  // replicas x 2*C doubles + ticket counter
constexpr int kMaxPartialBlocks = 6 * 148;                         // backward reduce: per-block partial rows
constexpr int kWsFloats = kAccumFloats + kMaxPartialBlocks * 2 * kBnMaxC;

struct RowMap {
  int cv, rows_per_iter;
};
static inline RowMap make_rowmap(int C) {
  RowMap m;
  m.cv = C / 8;
  m.rows_per_iter = kBnThreads / m.cv;
  return m;
}
static inline int reduce_blocks(long long M, int C, const RowMap& rm, int resident_per_sm = 6) {
  long long iters = (M + rm.rows_per_iter - 1) / rm.rows_per_iter;
  long long want = (iters + 7) / 8;          // >= 8 row-iterations per block
  long long cap = 1200000LL / (2 * C);       // bound the number of fp64 atomics (blocks * 2C)
  if (cap < sm_count()) cap = sm_count();
  static const int env_per_sm = getenv("B200_BN_REDUCE_BLOCKS_PER_SM") ? atoi(getenv("B200_BN_REDUCE_BLOCKS_PER_SM")) : 0;
  const int per_sm = env_per_sm > 0 ? env_per_sm : resident_per_sm;   // one full wave of fat blocks, no tail wave
  if (cap > per_sm * sm_count()) cap = per_sm * sm_count();
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}
// backward reduce (two-stage, no atomics): one resident wave of fat blocks, each at least 4 row-iterations
static inline int partial_blocks(long long M, const RowMap& rm, int resident_per_sm) {
  long long iters = (M + rm.rows_per_iter - 1) / rm.rows_per_iter;
  long long want = (iters + 3) / 4;
  static const int env_per_sm = getenv("B200_BN_REDUCE_BLOCKS_PER_SM") ? atoi(getenv("B200_BN_REDUCE_BLOCKS_PER_SM")) : 0;
  long long cap = (long long)(env_per_sm > 0 ? env_per_sm : resident_per_sm) * sm_count();
  if (cap > kMaxPartialBlocks) cap = kMaxPartialBlocks;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}
static inline int stream_blocks(long long M, const RowMap& rm) {
  long long iters = (M + rm.rows_per_iter - 1) / rm.rows_per_iter;
  long long blocks = (iters + 15) / 16;
  if (blocks < 1) blocks = 1;
  if (blocks > 4 * kBnMaxBlocks) blocks = 4 * kBnMaxBlocks;
  return (int)blocks;
}

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
// streaming 128-bit load: read-only path, no L1 allocation (every byte is touched once per kernel)
__device__ __forceinline__ uint4 ld_stream(const __nv_bfloat16* p) {
  uint4 u;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
  return u;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void loadf8(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
// 1 where the activation passes the gradient, evaluated on the PRE-activation value (same test as act_mask on y)
__device__ __forceinline__ int act_mask_value(float pre, int act) {
  if (act == B200_ACT_RELU) return pre > 0.f ? 1 : 0;
  if (act == B200_ACT_RELU6) return (pre > 0.f && pre < 6.f) ? 1 : 0;
  return 1;
}
__device__ __forceinline__ float act_mask(float v, int act) {
  if (act == B200_ACT_RELU) return v > 0.f ? 1.f : 0.f;
  if (act == B200_ACT_RELU6) return (v > 0.f && v < 6.f) ? 1.f : 0.f;
  return 1.f;
}

// Block partials -> fp64 global accumulators; returns true in the LAST block of the grid (after a grid-wide
// happens-before: every other block's atomics are visible).
__device__ __forceinline__ bool accumulate_and_elect(float (&acc)[16], int cv, int rows_per_iter, int C,
                                                     double* accum, unsigned* ticket) {
  __shared__ float red[kBnThreads][17];
  __shared__ bool is_last;
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 16; ++i) red[t][i] = acc[i];
  __syncthreads();
  for (int o = t; o < 2 * C; o += kBnThreads) {
    const int stat = o / C;
    const int c = o - stat * C;
    const int v = c >> 3, e = c & 7;
    float s = 0.f;
    for (int r = 0; r < rows_per_iter; ++r) s += red[r * cv + v][stat * 8 + e];
    atomicAdd(accum + (blockIdx.x % kReplicas) * 2 * C + o, (double)s);
  }
  __threadfence();
  __syncthreads();
  if (t == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// ---- forward statistics -------------------------------------------------------------------------
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(
    const __nv_bfloat16* __restrict__ z, long long M, int C, int cv, int rows_per_iter,
    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum, float* running_mean,
    float* running_var, long long* num_batches_tracked, float* mean, float* invstd, float* scale, float* shift,
    double* accum, unsigned* ticket) {
  pdl_wait();
  const int t = threadIdx.x;
  const bool active = t < rows_per_iter * cv;
  const int r0 = t / cv, v = t - r0 * cv;
  const long long rows_per_block = (M + gridDim.x - 1) / gridDim.x;
  const long long row_begin = blockIdx.x * rows_per_block;
  const long long row_end = min(M, row_begin + rows_per_block);
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  if (active) {
    long long r = row_begin + r0;
    for (; r + 7LL * rows_per_iter < row_end; r += 8LL * rows_per_iter) {
      uint4 raw[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        raw[u] = *reinterpret_cast<const uint4*>(z + (r + (long long)u * rows_per_iter) * C + v * 8);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float f[8];
        unpack8(raw[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i] += f[i]; acc[8 + i] += f[i] * f[i]; }
      }
    }
    for (; r < row_end; r += rows_per_iter) {
      float f[8];
      load8(z + r * C + v * 8, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { acc[i] += f[i]; acc[8 + i] += f[i] * f[i]; }
    }
  }
  if (!accumulate_and_elect(acc, cv, rows_per_iter, C, accum, ticket)) return;
  // ---- last block: finalise ----
  float f = momentum;
  if (momentum < 0.f) {  // cumulative moving average (momentum=None): factor 1/(num_batches_tracked+1)
    const long long nbt = num_batches_tracked ? *num_batches_tracked : 0;
    f = 1.f / (float)(nbt + 1);
  }
  for (int c = t; c < C; c += kBnThreads) {
    double s1 = 0.0, s2 = 0.0;
    for (int rep = 0; rep < kReplicas; ++rep) {
      s1 += __ldcg(accum + rep * 2 * C + c);
      s2 += __ldcg(accum + rep * 2 * C + C + c);
      accum[rep * 2 * C + c] = 0.0;
      accum[rep * 2 * C + C + c] = 0.0;
    }
    const double mu = s1 / (double)M;
    double var = s2 / (double)M - mu * mu;
    if (var < 0.0) var = 0.0;
    const float istd = (float)(1.0 / sqrt(var + (double)eps));
    mean[c] = (float)mu;
    invstd[c] = istd;
    const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
    const float sc = g * istd;
    scale[c] = sc;
    shift[c] = bt - (float)mu * sc;
    if (running_mean != nullptr && running_var != nullptr) {
      const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
      running_mean[c] = (1.f - f) * running_mean[c] + f * (float)mu;
      running_var[c] = (1.f - f) * running_var[c] + f * (float)unbiased;
    }
  }
  __syncthreads();
  if (t == 0) {
    *ticket = 0u;
    if (num_batches_tracked != nullptr && running_mean != nullptr) *num_batches_tracked += 1;
  }
}

// finalisation of statistics accumulated elsewhere (conv epilogue): same math as the last block of bn_stats_kernel
__global__ void __launch_bounds__(kBnThreads) bn_finalize_kernel(
    long long M, int C, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
    float* running_mean, float* running_var, long long* num_batches_tracked, float* mean, float* invstd, float* scale,
    float* shift, double* accum) {
  pdl_wait();
  // 16 lanes per channel, one per accumulator replica (a serial loop over the replicas was a chain of L2 round
  // trips in a kernel that sits on the critical path between the convolution and the BN apply)
  static_assert(kReplicas == 16, "lane mapping below assumes 16 replicas");
  const int rep = threadIdx.x & 15;
  const int c = blockIdx.x * (kBnThreads / 16) + (threadIdx.x >> 4);
  float f = momentum;
  if (momentum < 0.f) {
    const long long nbt = num_batches_tracked ? *num_batches_tracked : 0;
    f = 1.f / (float)(nbt + 1);
  }
  double s1 = 0.0, s2 = 0.0;
  if (c < C) {
    s1 = accum[rep * 2 * C + c];
    s2 = accum[rep * 2 * C + C + c];
    accum[rep * 2 * C + c] = 0.0;
    accum[rep * 2 * C + C + c] = 0.0;
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (c < C && rep == 0) {
    const double mu = s1 / (double)M;
    double var = s2 / (double)M - mu * mu;
    if (var < 0.0) var = 0.0;
    const float istd = (float)(1.0 / sqrt(var + (double)eps));
    mean[c] = (float)mu;
    invstd[c] = istd;
    const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
    const float sc = g * istd;
    scale[c] = sc;
    shift[c] = bt - (float)mu * sc;
    if (running_mean != nullptr && running_var != nullptr) {
      const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
      running_mean[c] = (1.f - f) * running_mean[c] + f * (float)mu;
      running_var[c] = (1.f - f) * running_var[c] + f * (float)unbiased;
    }
  }
  // with a fixed momentum nobody reads the counter, so it can be bumped here; the cumulative mode (momentum < 0)
  // reads it in every block above and gets a separate, stream-ordered bump kernel instead
  if (momentum >= 0.f && blockIdx.x == 0 && threadIdx.x == 0 && num_batches_tracked != nullptr && running_mean != nullptr)
    *num_batches_tracked += 1;
}
__global__ void bn_bump_kernel(long long* nbt) {
  pdl_wait();
  *nbt += 1;
}

__global__ void bn_eval_coeffs_kernel(int C, const float* gamma, const float* beta, const float* rm, const float* rv,
                                      float eps, float* scale, float* shift) {
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float istd = rsqrtf(rv[c] + eps);
  const float sc = (gamma ? gamma[c] : 1.f) * istd;
  scale[c] = sc;
  shift[c] = (beta ? beta[c] : 0.f) - rm[c] * sc;
}

// Activation-mask layout ("row quads"): the bytes of rows 4q..4q+3 for one 8-channel vector form ONE 32-bit word,
// word index q * (C/8) + v8.  A backward thread that owns ROWS consecutive rows fetches their masks with a single
// 16/32-bit load -- with a plain [row][C/8] byte layout every row cost its own load instruction, and the kernels are
// bound by requests in flight, not by bytes (that version was slower than re-reading the bf16 output).
__device__ __forceinline__ long long mask_byte_index(long long row, int cv8, int v8) {
  return (((row >> 2) * cv8 + v8) << 2) + (row & 3);
}
// mask bytes of rows [row0, row0 + ROWS) (row0 % ROWS == 0, ROWS in {2, 4, 8}) for vector v8: byte u = row row0 + u
template <int ROWS>
__device__ __forceinline__ unsigned long long mask_rows(const uint8_t* __restrict__ amask, long long row0, int cv8,
                                                        int v8) {
  const uint8_t* p = amask + mask_byte_index(row0, cv8, v8);
  if (ROWS == 2) return __ldg(reinterpret_cast<const unsigned short*>(p));
  if (ROWS == 4) return __ldg(reinterpret_cast<const unsigned int*>(p));
  const unsigned long long lo = __ldg(reinterpret_cast<const unsigned int*>(p));
  const unsigned long long hi = __ldg(reinterpret_cast<const unsigned int*>(p + 4LL * cv8));
  return lo | (hi << 32);
}

// ---- forward apply ------------------------------------------------------------------------------
template <int MODE>  // 0: none, 1: + residual, 2: + (z2*scale2+shift2)
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(
    const __nv_bfloat16* __restrict__ z, long long M, int C, int cv, int rows_per_iter,
    const float* __restrict__ scale, const float* __restrict__ shift, const __nv_bfloat16* __restrict__ res,
    const float* __restrict__ scale2, const float* __restrict__ shift2, int act, __nv_bfloat16* __restrict__ y,
    uint8_t* __restrict__ act_mask) {
  pdl_wait();
  const int t = threadIdx.x;
  if (t >= rows_per_iter * cv) return;
  const int r0 = t / cv, v = t - r0 * cv;
  const long long rows_per_block = (M + gridDim.x - 1) / gridDim.x;
  const long long row_begin = blockIdx.x * rows_per_block;
  const long long row_end = min(M, row_begin + rows_per_block);
  float sc[8], sh[8], sc2[8], sh2[8];
  loadf8(scale + v * 8, sc);
  loadf8(shift + v * 8, sh);
  if (MODE == 2) { loadf8(scale2 + v * 8, sc2); loadf8(shift2 + v * 8, sh2); }
  for (long long r = row_begin + r0; r < row_end; r += 2LL * rows_per_iter) {
    const long long ra = r, rb = r + rows_per_iter;
    const bool hb = rb < row_end;
    float fa[8], fb[8], ga[8], gb[8];
    uint32_t ma = 0, mb = 0;
    load8(z + ra * C + v * 8, fa);
    if (hb) load8(z + rb * C + v * 8, fb);
    if (MODE != 0) {
      load8(res + ra * C + v * 8, ga);
      if (hb) load8(res + rb * C + v * 8, gb);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = fa[i] * sc[i] + sh[i];
      float b = hb ? fb[i] * sc[i] + sh[i] : 0.f;
      if (MODE == 1) { a += ga[i]; if (hb) b += gb[i]; }
      if (MODE == 2) { a += ga[i] * sc2[i] + sh2[i]; if (hb) b += gb[i] * sc2[i] + sh2[i]; }
      ma |= static_cast<uint32_t>(act_mask_value(a, act)) << i;
      mb |= static_cast<uint32_t>(act_mask_value(b, act)) << i;
      if (act == B200_ACT_RELU) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
      if (act == B200_ACT_RELU6) { a = fminf(fmaxf(a, 0.f), 6.f); b = fminf(fmaxf(b, 0.f), 6.f); }
      fa[i] = a; fb[i] = b;
    }
    store8(y + ra * C + v * 8, fa);
    if (hb) store8(y + rb * C + v * 8, fb);
    if (act_mask != nullptr) {   // one byte per (row, 8-channel vector): bit i = act'(.) of channel v*8+i
      act_mask[mask_byte_index(ra, cv, v)] = static_cast<uint8_t>(ma);
      if (hb) act_mask[mask_byte_index(rb, cv, v)] = static_cast<uint8_t>(mb);
    }
  }
}

// ---- backward kernels -----------------------------------------------------------------------------
// Thread mapping: VEC channels per thread (VEC=4: 64-bit accesses, half the per-channel coefficient registers of
// VEC=8, which is what keeps these 2-read(+1-write) streams at the occupancy of bn_apply; VEC=8 only for C > 1024).
// g = dy * act'(.), where the activation argument is y when given, else recomputed as z*scale+shift (bit-identical
// to the forward's fused multiply-add).
template <int VEC> struct RawVec;
template <> struct RawVec<4> { uint2 u; };
template <> struct RawVec<8> { uint4 u; };
__device__ __forceinline__ RawVec<4> ldv(const __nv_bfloat16* p, RawVec<4>*) {
  RawVec<4> r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.u.x), "=r"(r.u.y) : "l"(p));
  return r;
}
__device__ __forceinline__ RawVec<8> ldv(const __nv_bfloat16* p, RawVec<8>*) {
  RawVec<8> r;
  r.u = ld_stream(p);
  return r;
}
__device__ __forceinline__ void unpackv(const RawVec<4>& r, float (&f)[4]) {
  const float2 a = unpack_bf16x2(r.u.x), b = unpack_bf16x2(r.u.y);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ void unpackv(const RawVec<8>& r, float (&f)[8]) { unpack8(r.u, f); }
__device__ __forceinline__ void storev(__nv_bfloat16* p, const float (&f)[4]) {
  uint2 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void storev(__nv_bfloat16* p, const float (&f)[8]) { store8(p, f); }
template <int VEC>
__device__ __forceinline__ void loadfv(const float* p, float (&f)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC; i += 4) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p + i));
    f[i] = a.x; f[i + 1] = a.y; f[i + 2] = a.z; f[i + 3] = a.w;
  }
}

// Gradient arriving at a pre-pool pixel of a 3x3 / stride 2 / pad 1 max pool, gathered from the POOLED gradient dp
// [N, OH, OW, C] through the argmax bytes (pool.cu) -- the stem's BN backward reads dp + one byte per pooled element
// instead of a materialised [N, H, W, C] gradient tensor (saves the max-pool backward kernel: one write and two reads
// of the largest activation of the network).  A pixel lies in at most 2 x 2 windows; fp32 sum, no rounding.
struct PoolGeom {
  int H, W, OH, OW;
};
template <int VEC>
__device__ __forceinline__ void pool_gather(const __nv_bfloat16* __restrict__ dp, const uint8_t* __restrict__ amax,
                                            long long row, long long col, int C, const PoolGeom& pg, float (&g)[VEC]) {
  const unsigned r32 = (unsigned)row;                 // the host checks N*H*W < 2^31: 32-bit div/mod
  const int w = (int)(r32 % (unsigned)pg.W);
  const unsigned t = r32 / (unsigned)pg.W;
  const int h = (int)(t % (unsigned)pg.H);
  const long long n = t / (unsigned)pg.H;
  const int p_lo = h >> 1, q_lo = w >> 1;
  uint32_t ab[4][VEC / 4];
  uint32_t gr[4][VEC / 2];
  int want[4];
  bool ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = p_lo + (j >> 1), q = q_lo + (j & 1);
    ok[j] = p <= ((h + 1) >> 1) && q <= ((w + 1) >> 1) && p < pg.OH && q < pg.OW;
    want[j] = (h - (2 * p - 1)) * 3 + (w - (2 * q - 1));
    if (ok[j]) {
      const long long o = ((n * pg.OH + p) * pg.OW + q) * C + col;
      if constexpr (VEC == 8) {
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(amax + o));
        const uint4 d = __ldg(reinterpret_cast<const uint4*>(dp + o));
        ab[j][0] = a.x; ab[j][1] = a.y;
        gr[j][0] = d.x; gr[j][1] = d.y; gr[j][2] = d.z; gr[j][3] = d.w;
      } else {
        ab[j][0] = __ldg(reinterpret_cast<const uint32_t*>(amax + o));
        const uint2 d = __ldg(reinterpret_cast<const uint2*>(dp + o));
        gr[j][0] = d.x; gr[j][1] = d.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) g[i] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (!ok[j]) continue;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int b = (ab[j][i >> 2] >> (8 * (i & 3))) & 0xff;
      const float2 g2 = unpack_bf16x2(gr[j][i >> 1]);
      if (b == want[j]) g[i] += (i & 1) ? g2.y : g2.x;
    }
  }
}

// ---- backward reduce: dbeta = sum g, dgamma = sum g * xhat ------------------------------------------
// SRC: activation argument 0 recomputed from z, 1 = y, 2 = mask bits; 3 = like 0, with the gradient gathered from a
// max-pooled gradient (dy = dp, amask = argmax bytes, pool_gather above)
template <int VEC, int ROWS, int MINB, int SRC>
__global__ void __launch_bounds__(kBnThreads, MINB) bn_bwd_reduce_kernel(
    const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y, const uint8_t* __restrict__ amask,
    const __nv_bfloat16* __restrict__ z, long long M, int C, int cv, int rows_per_iter, int act,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* __restrict__ partial, const PoolGeom pg) {
  pdl_wait();
  const int t = threadIdx.x;
  const bool active = t < rows_per_iter * cv;
  const int r0 = t / cv, v = t - r0 * cv;
  // SRC 2 (mask words): a thread owns ROWS CONSECUTIVE rows (one mask load); block ranges are multiples of 8 rows
  const long long rows_per_block = SRC == 2 ? (((M + gridDim.x - 1) / gridDim.x + 7) & ~7LL) : (M + gridDim.x - 1) / gridDim.x;
  const long long row_begin = blockIdx.x * rows_per_block;
  const long long row_end = min(M, row_begin + rows_per_block);
  constexpr long long kRowStep = SRC == 2 ? 1 : 0;       // distance between a thread's rows: 1 or rows_per_iter
  float acc[2 * VEC];
#pragma unroll
  for (int i = 0; i < 2 * VEC; ++i) acc[i] = 0.f;
  if (active) {
    float mu[VEC], sc[VEC], sh[VEC];
    loadfv<VEC>(mean + v * VEC, mu);
    if ((SRC == 0 || SRC == 3) && act != B200_ACT_NONE) {
      float is[VEC];
      loadfv<VEC>(invstd + v * VEC, is);
      if (gamma) loadfv<VEC>(gamma + v * VEC, sc);
      if (beta) loadfv<VEC>(beta + v * VEC, sh);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        sc[i] = (gamma ? sc[i] : 1.f) * is[i];
        sh[i] = (beta ? sh[i] : 0.f) - mu[i] * sc[i];
      }
    }
    const long long col = (long long)v * VEC;
    const long long rstep = kRowStep ? 1 : rows_per_iter;
    for (long long r = row_begin + (kRowStep ? (long long)r0 * ROWS : r0); r < row_end;
         r += (long long)ROWS * rows_per_iter) {
      RawVec<VEC> rd[ROWS], rz[ROWS], ry[ROWS];
      unsigned long long rm = 0;
      bool ok[ROWS];
      if (SRC == 2) rm = mask_rows<ROWS>(amask, r, C >> 3, (int)(col >> 3)) >> (col & 7 & ~(VEC - 1));
#pragma unroll
      for (int u = 0; u < ROWS; ++u) {
        const long long rr = r + (long long)u * rstep;
        ok[u] = rr < row_end;
        if (ok[u]) {
          if (SRC != 3) rd[u] = ldv(dy + rr * C + col, (RawVec<VEC>*)nullptr);
          rz[u] = ldv(z + rr * C + col, (RawVec<VEC>*)nullptr);
          if (SRC == 1) ry[u] = ldv(y + rr * C + col, (RawVec<VEC>*)nullptr);
        }
      }
#pragma unroll
      for (int u = 0; u < ROWS; ++u) {
        if (!ok[u]) continue;
        float da[VEC], za[VEC], ya[VEC];
        if (SRC == 3) pool_gather<VEC>(dy, amask, r + (long long)u * rstep, col, C, pg, da);
        else unpackv(rd[u], da);
        unpackv(rz[u], za);
        if (SRC == 1) unpackv(ry[u], ya);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          float g = da[i];
          if (SRC == 2) g = ((rm >> (8 * u + i)) & 1ull) ? g : 0.f;
          else if (act != B200_ACT_NONE) g *= act_mask(SRC == 1 ? ya[i] : fmaf(za[i], sc[i], sh[i]), act);
          acc[i] = fmaf(g, za[i] - mu[i], acc[i]);   // the 1/std factor is applied once per channel below
          acc[VEC + i] += g;
        }
      }
    }
    float is[VEC];
    loadfv<VEC>(invstd + v * VEC, is);
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] *= is[i];
  }
  // block partial: fold the rows_per_iter row groups in shared memory, one coalesced row of 2C floats per block
  __shared__ float red[kBnThreads][2 * VEC + 1];
#pragma unroll
  for (int i = 0; i < 2 * VEC; ++i) red[t][i] = acc[i];
  __syncthreads();
  float* dst = partial + (size_t)blockIdx.x * 2 * C;
  for (int o = t; o < 2 * C; o += kBnThreads) {
    const int stat = o / C;
    const int c = o - stat * C;
    const int vv = c / VEC, e = c - vv * VEC;
    float sacc = 0.f;
    for (int r = 0; r < rows_per_iter; ++r) sacc += red[r * cv + vv][stat * VEC + e];
    dst[o] = sacc;
  }
}

// second stage: sums[o] = sum over blocks of partial[b][o]  (o in [0, 2C): dgamma then dbeta); a block of 256
// threads owns 4 outputs x 64 interleaved block groups (every load independent: the kernel is pure latency),
// fp64 across the groups.
__global__ void __launch_bounds__(kBnThreads) bn_bwd_reduce_final_kernel(const float* __restrict__ partial, int nblocks,
                                                                          int C, float* __restrict__ sums,
                                                                          float* dgamma_acc, float* dbeta_acc) {
  pdl_wait();
  __shared__ double red[64][5];
  const int lane_o = threadIdx.x & 3, grp = threadIdx.x >> 2;
  const int o = blockIdx.x * 4 + lane_o;
  double acc = 0.0;
  if (o < 2 * C) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int b = grp;
    for (; b + 192 < nblocks; b += 256) {
      a0 += __ldcg(partial + (size_t)b * 2 * C + o);
      a1 += __ldcg(partial + (size_t)(b + 64) * 2 * C + o);
      a2 += __ldcg(partial + (size_t)(b + 128) * 2 * C + o);
      a3 += __ldcg(partial + (size_t)(b + 192) * 2 * C + o);
    }
    for (; b < nblocks; b += 64) a0 += __ldcg(partial + (size_t)b * 2 * C + o);
    acc = (double)a0 + (double)a1 + (double)a2 + (double)a3;
  }
  red[grp][lane_o] = acc;
  __syncthreads();
  if (grp == 0 && o < 2 * C) {
    double tot = 0.0;
#pragma unroll 8
    for (int g = 0; g < 64; ++g) tot += red[g][lane_o];
    const float f = (float)tot;
    sums[o] = f;
    if (o < C) { if (dgamma_acc) dgamma_acc[o] += f; }
    else if (dbeta_acc) dbeta_acc[o - C] += f;
  }
}

// ---- backward dx --------------------------------------------------------------------------------
template <int VEC, int ROWS, int MINB, int SRC>   // SRC as in bn_bwd_reduce_kernel
__global__ void __launch_bounds__(kBnThreads, MINB) bn_bwd_dx_kernel(
    const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y, const uint8_t* __restrict__ amask,
    const __nv_bfloat16* __restrict__ z, long long M, int C, int cv, int rows_per_iter, int act,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ sums, __nv_bfloat16* __restrict__ dz,
    __nv_bfloat16* __restrict__ g_out, const PoolGeom pg) {
  pdl_wait();
  const int t = threadIdx.x;
  if (t >= rows_per_iter * cv) return;
  const int r0 = t / cv, v = t - r0 * cv;
  const long long rows_per_block = SRC == 2 ? (((M + gridDim.x - 1) / gridDim.x + 7) & ~7LL) : (M + gridDim.x - 1) / gridDim.x;
  const long long row_begin = blockIdx.x * rows_per_block;
  const long long row_end = min(M, row_begin + rows_per_block);
  constexpr long long kRowStep = SRC == 2 ? 1 : 0;       // as in bn_bwd_reduce_kernel
  // dz = A*g + B*z + Cc  with A = gamma*istd, B = -gamma*istd^2*dgamma/M, Cc = -A*dbeta/M - B*mean;
  // the activation argument recomputed from z is z*A + sh
  float A[VEC], B[VEC], Cc[VEC], sh[VEC];
  {
    float mu[VEC], is[VEC], dg[VEC], dbt[VEC];
    loadfv<VEC>(mean + v * VEC, mu);
    loadfv<VEC>(invstd + v * VEC, is);
    if (gamma) loadfv<VEC>(gamma + v * VEC, A);
    if (beta) loadfv<VEC>(beta + v * VEC, sh);
    loadfv<VEC>(sums + v * VEC, dg);
    loadfv<VEC>(sums + C + v * VEC, dbt);
    const float invM = 1.f / (float)M;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float gm = gamma ? A[i] : 1.f;
      A[i] = gm * is[i];
      B[i] = -gm * is[i] * is[i] * dg[i] * invM;
      Cc[i] = -A[i] * dbt[i] * invM - B[i] * mu[i];
      sh[i] = (beta ? sh[i] : 0.f) - mu[i] * A[i];
    }
  }
  const long long col = (long long)v * VEC;
  const long long rstep = kRowStep ? 1 : rows_per_iter;
  for (long long r = row_begin + (kRowStep ? (long long)r0 * ROWS : r0); r < row_end;
       r += (long long)ROWS * rows_per_iter) {
    RawVec<VEC> rd[ROWS], rz[ROWS], ry[ROWS];
    unsigned long long rm = 0;
    bool ok[ROWS];
    if (SRC == 2) rm = mask_rows<ROWS>(amask, r, C >> 3, (int)(col >> 3)) >> (col & 7 & ~(VEC - 1));
#pragma unroll
    for (int u = 0; u < ROWS; ++u) {
      const long long rr = r + (long long)u * rstep;
      ok[u] = rr < row_end;
      if (ok[u]) {
        if (SRC != 3) rd[u] = ldv(dy + rr * C + col, (RawVec<VEC>*)nullptr);
        rz[u] = ldv(z + rr * C + col, (RawVec<VEC>*)nullptr);
        if (SRC == 1) ry[u] = ldv(y + rr * C + col, (RawVec<VEC>*)nullptr);
      }
    }
#pragma unroll
    for (int u = 0; u < ROWS; ++u) {
      if (!ok[u]) continue;
      const long long rr = r + (long long)u * rstep;
      float da[VEC], za[VEC], ya[VEC];
      if (SRC == 3) pool_gather<VEC>(dy, amask, rr, col, C, pg, da);
      else unpackv(rd[u], da);
      unpackv(rz[u], za);
      if (SRC == 1) unpackv(ry[u], ya);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float g = da[i];
        if (SRC == 2) g = ((rm >> (8 * u + i)) & 1ull) ? g : 0.f;
        else if (act != B200_ACT_NONE) g *= act_mask(SRC == 1 ? ya[i] : fmaf(za[i], A[i], sh[i]), act);
        da[i] = g;
        za[i] = A[i] * g + B[i] * za[i] + Cc[i];
      }
      storev(dz + rr * C + col, za);
      if (g_out) storev(g_out + rr * C + col, da);
    }
  }
}

template <int VEC>
static inline RowMap make_rowmap_v(int C) {
  RowMap m;
  m.cv = C / VEC;
  m.rows_per_iter = kBnThreads / m.cv;
  return m;
}

// tuning knob (measured with tools/bn_bench.py): rows in flight per thread x resident blocks per SM
static int bwd_variant() {
  static const int v = getenv("B200_BN_BWD_VARIANT") ? atoi(getenv("B200_BN_BWD_VARIANT")) : 0;
  return v;
}
static int check_c(int C, const char* who) {
  B200_REQUIRE(C > 0 && C % 8 == 0 && C <= kBnMaxC, B200_ERR_UNSUPPORTED,
               "%s: C=%d must be a multiple of 8 and <= %d", who, C, kBnMaxC);
  return B200_OK;
}
static inline double* ws_accum(float* ws) { return reinterpret_cast<double*>(ws); }
static inline unsigned* ws_ticket(float* ws) { return reinterpret_cast<unsigned*>(ws + kReplicas * 2 * kBnMaxC * 2); }

}  // namespace b200

using namespace b200;

extern "C" size_t b200_bn_workspace_floats(int C) { (void)C; return (size_t)kWsFloats; }
// rows padded to a multiple of 8: the backward kernels read whole row-quad words (two of them with 8 rows in flight)
extern "C" size_t b200_bn_act_mask_bytes(long long M, int C) { return (size_t)((M + 7) / 8 * 8) * (size_t)(C / 8); }

extern "C" int b200_bn_stats(const void* z, long long M, int C, const float* gamma, const float* beta, float eps,
                             float momentum, float* running_mean, float* running_var, long long* nbt, float* mean,
                             float* invstd, float* scale, float* shift, float* workspace, b200_stream_t stream_) {
  int rc = check_c(C, "bn_stats");
  if (rc) return rc;
  B200_REQUIRE(z && mean && invstd && scale && shift && workspace && M > 0, B200_ERR_INVALID, "bn_stats: bad argument");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, B200_ERR_INVALID, "bn_stats: workspace misaligned");
  const RowMap rm = make_rowmap(C);
  b200::launch(bn_stats_kernel, reduce_blocks(M, C, rm), kBnThreads, 0, (cudaStream_t)stream_,
      (const __nv_bfloat16*)z, M, C, rm.cv, rm.rows_per_iter, gamma, beta, eps, momentum, running_mean, running_var,
      nbt, mean, invstd, scale, shift, ws_accum(workspace), ws_ticket(workspace));
  B200_CHECK_LAUNCH("bn_stats_kernel");
  return B200_OK;
}

extern "C" int b200_bn_finalize(long long M, int C, const float* gamma, const float* beta, float eps, float momentum,
                                float* running_mean, float* running_var, long long* nbt, float* mean, float* invstd,
                                float* scale, float* shift, float* workspace, b200_stream_t stream_) {
  int rc = check_c(C, "bn_finalize");
  if (rc) return rc;
  B200_REQUIRE(mean && invstd && scale && shift && workspace && M > 0, B200_ERR_INVALID, "bn_finalize: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  // NOTE: momentum < 0 reads *nbt before the bump below (same stream => ordered)
  b200::launch(bn_finalize_kernel, (C + 15) / 16, kBnThreads, 0, stream,
      M, C, gamma, beta, eps, momentum, running_mean, running_var, nbt, mean, invstd, scale, shift, ws_accum(workspace));
  B200_CHECK_LAUNCH("bn_finalize_kernel");
  if (momentum < 0.f && nbt != nullptr && running_mean != nullptr) {
    b200::launch(bn_bump_kernel, 1, 1, 0, stream, nbt);
    B200_CHECK_LAUNCH("bn_bump_kernel");
  }
  return B200_OK;
}

extern "C" int b200_bn_eval_coeffs(int C, const float* gamma, const float* beta, const float* running_mean,
                                   const float* running_var, float eps, float* scale, float* shift,
                                   b200_stream_t stream_) {
  B200_REQUIRE(C > 0 && running_mean && running_var && scale && shift, B200_ERR_INVALID, "bn_eval_coeffs: bad argument");
  b200::launch(bn_eval_coeffs_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream_, C, gamma, beta, running_mean, running_var,
                                                                         eps, scale, shift);
  B200_CHECK_LAUNCH("bn_eval_coeffs_kernel");
  return B200_OK;
}

extern "C" int b200_bn_apply(const void* z, long long M, int C, const float* scale, const float* shift,
                             const void* residual, const void* z2, const float* scale2, const float* shift2, int act,
                             void* y, uint8_t* act_mask, b200_stream_t stream_) {
  int rc = check_c(C, "bn_apply");
  if (rc) return rc;
  B200_REQUIRE(z && scale && shift && y && M > 0, B200_ERR_INVALID, "bn_apply: bad argument");
  B200_REQUIRE(!(residual && z2), B200_ERR_INVALID, "bn_apply: residual and z2 are exclusive");
  B200_REQUIRE(!z2 || (scale2 && shift2), B200_ERR_INVALID, "bn_apply: z2 needs scale2/shift2");
  cudaStream_t stream = (cudaStream_t)stream_;
  const RowMap rm = make_rowmap(C);
  const int blocks = stream_blocks(M, rm);
  const __nv_bfloat16* zz = (const __nv_bfloat16*)z;
  __nv_bfloat16* yy = (__nv_bfloat16*)y;
  if (residual)
    b200::launch(bn_apply_kernel<1>, blocks, kBnThreads, 0, stream, zz, M, C, rm.cv, rm.rows_per_iter, scale, shift,
                                                         (const __nv_bfloat16*)residual, nullptr, nullptr, act, yy, act_mask);
  else if (z2)
    b200::launch(bn_apply_kernel<2>, blocks, kBnThreads, 0, stream, zz, M, C, rm.cv, rm.rows_per_iter, scale, shift,
                                                         (const __nv_bfloat16*)z2, scale2, shift2, act, yy, act_mask);
  else
    b200::launch(bn_apply_kernel<0>, blocks, kBnThreads, 0, stream, zz, M, C, rm.cv, rm.rows_per_iter, scale, shift, nullptr,
                                                         nullptr, nullptr, act, yy, act_mask);
  B200_CHECK_LAUNCH("bn_apply_kernel");
  return B200_OK;
}

// pooled: dy is the gradient of a 3x3/s2/p1 max pool's OUTPUT and act_mask holds the argmax bytes (SRC 3)
static int bn_bwd_reduce_impl(const void* dy, const void* y, const uint8_t* act_mask, const void* z, long long M, int C,
                              int act, const float* mean, const float* invstd, const float* gamma, const float* beta,
                              float* sums, float* dgamma_acc, float* dbeta_acc, float* workspace, b200_stream_t stream_,
                              bool pooled, const PoolGeom pg) {
  int rc = check_c(C, "bn_bwd_reduce");
  if (rc) return rc;
  B200_REQUIRE(dy && z && mean && invstd && sums && workspace && M > 0, B200_ERR_INVALID, "bn_bwd_reduce: bad argument");
  // activation-argument source: 2 = mask bits, 1 = y, 0 = recomputed from z (or no activation)
  const int src = pooled ? 3 : ((act == B200_ACT_NONE) ? 0 : (act_mask != nullptr ? 2 : (y != nullptr ? 1 : 0)));
#define B200_RED_ARGS(VEC)                                                                                     \
  (const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, act_mask, (const __nv_bfloat16*)z, M, C, rm.cv,           \
      rm.rows_per_iter, act, mean, invstd, gamma, beta, partial, pg
#define B200_LAUNCH_RED(VEC, ROWS, MINB)                                                                       \
  do {                                                                                                         \
    const RowMap rm = make_rowmap_v<VEC>(C);                                                                   \
    blocks = partial_blocks(M, rm, MINB);                                                                      \
    if (src == 3)                                                                                              \
      b200::launch(bn_bwd_reduce_kernel<VEC, ROWS, MINB, 3>, blocks, kBnThreads, 0, stream, B200_RED_ARGS(VEC));         \
    else if (src == 2)                                                                                         \
      b200::launch(bn_bwd_reduce_kernel<VEC, ROWS, MINB, 2>, blocks, kBnThreads, 0, stream, B200_RED_ARGS(VEC));         \
    else if (src == 1)                                                                                         \
      b200::launch(bn_bwd_reduce_kernel<VEC, ROWS, MINB, 1>, blocks, kBnThreads, 0, stream, B200_RED_ARGS(VEC));         \
    else                                                                                                       \
      b200::launch(bn_bwd_reduce_kernel<VEC, ROWS, MINB, 0>, blocks, kBnThreads, 0, stream, B200_RED_ARGS(VEC));         \
  } while (0)
  cudaStream_t stream = (cudaStream_t)stream_;
  float* partial = workspace + kAccumFloats;
  int blocks = 0;
  if (pooled) {                 // the gather needs registers: two rows in flight instead of four
    B200_LAUNCH_RED(4, 2, 4);
  } else if (C > 1024) {
    B200_LAUNCH_RED(8, 2, 3);
  } else {
    switch (bwd_variant()) {
      case 1: B200_LAUNCH_RED(4, 2, 5); break;
      case 2: B200_LAUNCH_RED(4, 8, 3); break;
      case 3: B200_LAUNCH_RED(8, 2, 3); break;
      default: B200_LAUNCH_RED(4, 4, 4); break;
    }
  }
#undef B200_LAUNCH_RED
#undef B200_RED_ARGS
  B200_CHECK_LAUNCH("bn_bwd_reduce_kernel");
  b200::launch(bn_bwd_reduce_final_kernel, (2 * C + 3) / 4, kBnThreads, 0, stream, partial, blocks, C, sums, dgamma_acc,
                                                                           dbeta_acc);
  B200_CHECK_LAUNCH("bn_bwd_reduce_final_kernel");
  return B200_OK;
}

extern "C" int b200_bn_bwd_reduce(const void* dy, const void* y, const uint8_t* act_mask, const void* z, long long M,
                                  int C, int act,
                                  const float* mean, const float* invstd, const float* gamma, const float* beta,
                                  float* sums, float* dgamma_acc, float* dbeta_acc, float* workspace,
                                  b200_stream_t stream_) {
  return bn_bwd_reduce_impl(dy, y, act_mask, z, M, C, act, mean, invstd, gamma, beta, sums, dgamma_acc, dbeta_acc,
                            workspace, stream_, false, PoolGeom{0, 0, 0, 0});
}

static int pooled_geom(int N, int H, int W, int C, const void* dp, const uint8_t* argmax, PoolGeom* pg) {
  B200_REQUIRE(N > 0 && H > 1 && W > 1 && dp && argmax, B200_ERR_INVALID, "bn_bwd (pooled): bad argument");
  B200_REQUIRE(C % 8 == 0, B200_ERR_UNSUPPORTED, "bn_bwd (pooled): C=%d must be a multiple of 8", C);
  B200_REQUIRE((long long)N * H * W < (1LL << 31), B200_ERR_UNSUPPORTED, "bn_bwd (pooled): tensor too large");
  pg->H = H; pg->W = W; pg->OH = (H - 1) / 2 + 1; pg->OW = (W - 1) / 2 + 1;
  return B200_OK;
}

extern "C" int b200_bn_bwd_reduce_pooled(const void* dp, const uint8_t* argmax, const void* z, int N, int H, int W, int C,
                                         int act, const float* mean, const float* invstd, const float* gamma,
                                         const float* beta, float* sums, float* dgamma_acc, float* dbeta_acc,
                                         float* workspace, b200_stream_t stream_) {
  PoolGeom pg;
  int rc = pooled_geom(N, H, W, C, dp, argmax, &pg);
  if (rc) return rc;
  return bn_bwd_reduce_impl(dp, nullptr, argmax, z, (long long)N * H * W, C, act, mean, invstd, gamma, beta, sums,
                            dgamma_acc, dbeta_acc, workspace, stream_, true, pg);
}

static int bn_bwd_dx_impl(const void* dy, const void* y, const uint8_t* act_mask, const void* z, long long M, int C,
                          int act, const float* mean, const float* invstd, const float* gamma, const float* beta,
                          const float* sums, void* dz, void* g_out, b200_stream_t stream_, bool pooled,
                          const PoolGeom pg) {
  int rc = check_c(C, "bn_bwd_dx");
  if (rc) return rc;
  B200_REQUIRE(dy && z && mean && invstd && sums && dz && M > 0, B200_ERR_INVALID, "bn_bwd_dx: bad argument");
  const int src = pooled ? 3 : ((act == B200_ACT_NONE) ? 0 : (act_mask != nullptr ? 2 : (y != nullptr ? 1 : 0)));
#define B200_DX_ARGS(VEC)                                                                                    \
  (const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, act_mask, (const __nv_bfloat16*)z, M, C, rm.cv,         \
      rm.rows_per_iter, act, mean, invstd, gamma, beta, sums, (__nv_bfloat16*)dz, (__nv_bfloat16*)g_out, pg
#define B200_LAUNCH_DX(VEC, ROWS, MINB)                                                                      \
  do {                                                                                                       \
    const RowMap rm = make_rowmap_v<VEC>(C);                                                                 \
    const int blocks = stream_blocks(M, rm);                                                                 \
    if (src == 3)                                                                                            \
      b200::launch(bn_bwd_dx_kernel<VEC, ROWS, MINB, 3>, blocks, kBnThreads, 0, (cudaStream_t)stream_, B200_DX_ARGS(VEC)); \
    else if (src == 2)                                                                                       \
      b200::launch(bn_bwd_dx_kernel<VEC, ROWS, MINB, 2>, blocks, kBnThreads, 0, (cudaStream_t)stream_, B200_DX_ARGS(VEC)); \
    else if (src == 1)                                                                                       \
      b200::launch(bn_bwd_dx_kernel<VEC, ROWS, MINB, 1>, blocks, kBnThreads, 0, (cudaStream_t)stream_, B200_DX_ARGS(VEC)); \
    else                                                                                                     \
      b200::launch(bn_bwd_dx_kernel<VEC, ROWS, MINB, 0>, blocks, kBnThreads, 0, (cudaStream_t)stream_, B200_DX_ARGS(VEC)); \
  } while (0)
  if (pooled) {
    B200_LAUNCH_DX(4, 2, 4);
  } else if (C > 1024) {
    B200_LAUNCH_DX(8, 2, 3);
  } else {
    switch (bwd_variant()) {
      case 1: B200_LAUNCH_DX(4, 2, 5); break;
      case 2: B200_LAUNCH_DX(4, 8, 3); break;
      case 3: B200_LAUNCH_DX(8, 2, 3); break;
      default: B200_LAUNCH_DX(4, 4, 4); break;
    }
  }
#undef B200_LAUNCH_DX
#undef B200_DX_ARGS
  B200_CHECK_LAUNCH("bn_bwd_dx_kernel");
  return B200_OK;
}

extern "C" int b200_bn_bwd_dx(const void* dy, const void* y, const uint8_t* act_mask, const void* z, long long M, int C,
                              int act,
                              const float* mean, const float* invstd, const float* gamma, const float* beta,
                              const float* sums, void* dz, void* g_out, b200_stream_t stream_) {
  return bn_bwd_dx_impl(dy, y, act_mask, z, M, C, act, mean, invstd, gamma, beta, sums, dz, g_out, stream_, false,
                        PoolGeom{0, 0, 0, 0});
}

extern "C" int b200_bn_bwd_dx_pooled(const void* dp, const uint8_t* argmax, const void* z, int N, int H, int W, int C,
                                     int act, const float* mean, const float* invstd, const float* gamma,
                                     const float* beta, const float* sums, void* dz, b200_stream_t stream_) {
  PoolGeom pg;
  int rc = pooled_geom(N, H, W, C, dp, argmax, &pg);
  if (rc) return rc;
  return bn_bwd_dx_impl(dp, nullptr, argmax, z, (long long)N * H * W, C, act, mean, invstd, gamma, beta, sums, dz,
                        nullptr, stream_, true, pg);
}
