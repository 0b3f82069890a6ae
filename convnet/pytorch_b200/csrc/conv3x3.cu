// Stride-1 RxS convolutions (3x3 pad 1 fprop/dgrad; the 4x4 pad 0 space-to-depth stem) as a "shift GEMM" on tcgen05
// tensor cores.
//
// The im2col kernel (conv.cu) re-fetches the activation tile once per filter tap: 9 TMA loads of the same
// pixels, shifted.  Here a tile is RT full image rows; its zero-padded halo box [(RT+2) x (W+2) pixels x 64 ch]
// is loaded ONCE per 64-channel block (tiled 4D TMA, out-of-bounds = padding), and the nine taps are nine views of
// that one shared-memory buffer: the UMMA descriptor start address is simply advanced by (r*(W+2)+s) pixel rows
// (row-shifted 128B-swizzle descriptors are valid because the swizzle is a function of the absolute smem address;
// measured with tools/shift_probe.cu).  The 128 accumulator rows are "virtual pixels" of the padded row pitch
// W+2; the two halo columns per row are computed and discarded (2/(W+2) of the MMA work).
// Activation traffic from L2 drops ~9x -> ~1.3x; weights stream through a separate ring.
//
// Pixel rows are 128 B (64-channel chunks, 128B swizzle) or, for the 16-channel stem, 32 B (32B swizzle, one K=16
// MMA per tap).  When all weight slices of a CTA fit in shared memory (64x64 3x3, the stem) they are loaded once
// ("stationary") instead of streaming through the ring with every tile.
//
// Same warp roles / TMEM double buffering / TMA-store epilogue / fused BN statistics as conv_igemm_kernel.
#include "common.cuh"
#include "host.h"
#include <stdlib.h>

namespace b200 {

constexpr int kHThreads = 320;
constexpr int kHTileM = 128;
constexpr int kHMaxA = 6, kHMaxB = 8;
constexpr int kHMaxTaps = 16;
constexpr int kHStatReplicas = 16;

struct HaloParams {
  int N, H, W, C, Kout;    // H x W: OUTPUT map (the source is (H+R-1-2*pad) x (W+S-1-2*pad))
  int RT, Wp;              // output rows per tile, halo row pitch W+S-1
  int tiles_per_img, m_tiles, n_tiles, block_n, c_chunks;
  int sa, sb;              // ring depths (sb unused when the weights are stationary)
  uint32_t a_bytes, a_box_bytes, b_bytes;
  int act;
  int has_res;
  double* stats;
  const float* bias;       // [Kout] fp32 added before the residual / activation (BN folded for inference), or nullptr
  int ntaps, pad;
  int row_bytes;           // bytes of one source pixel chunk: 128 (64 channels) or 32 (16 channels)
  int b_stationary;        // 1: all c_chunks*ntaps weight slices of the n-tile stay in shared memory
  int diag;                // 1: block-diagonal ("window") convolution: n-tile b (64 outputs) reads input channels
                           //    [64b, 64b+64) only; a CTA keeps ONE n-tile (its 9 weight slices stay resident) and walks
                           //    the pixel tiles -- how grouped convolutions run (include/b200conv.h, desc.window)
  uint16_t a_off[kHMaxTaps];  // pixel-row offset of each tap's view inside the halo buffer
  uint16_t b_tap[kHMaxTaps];  // weight tap slice used with it
};

// The MMA-issuing thread is latency-bound per instruction: descriptors are built once per operand buffer and
// advanced by adding (byte offset >> 4) to the address field, with the tap / k-step loops fully unrolled.
template <int NTAPS, int KSTEPS>
__global__ void __launch_bounds__(kHThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kHMaxA], a_empty[kHMaxA], b_full[kHMaxB], b_empty[kHMaxB];
  __shared__ __align__(8) uint64_t tmem_full[2], tmem_empty[2], res_bar, bstat_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sA = smem;
  uint8_t* sB = sA + p.sa * p.a_bytes;
  const int b_slots = p.b_stationary ? p.c_chunks * p.ntaps : p.sb;
  uint8_t* epi = sB + b_slots * p.b_bytes;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.sa; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kHMaxB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(&bstat_bar, 1);
    mbar_init(&tmem_full[0], 1); mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], 8); mbar_init(&tmem_empty[1], 8);
    mbar_init(&res_bar, 1);
    fence_mbar_init();
    prefetch_tmap(&tmX); prefetch_tmap(&tmB); prefetch_tmap(&tmC);
    if (p.has_res) prefetch_tmap(&tmR);
  }
  if (warp == 1) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // prologue above (barriers, TMEM, descriptor prefetch) overlaps the predecessor grid
  const uint32_t tmem_base = tmem_base_s;
  // tile walk: dense -- tiles (m, n) round-robin over the CTAs; diagonal -- the CTA owns n-tile blockIdx.x % n_tiles and
  // walks the m-tiles with stride gridDim.x / n_tiles (the grid is a multiple of n_tiles)
  const int t_start = p.diag ? static_cast<int>(blockIdx.x) / p.n_tiles : static_cast<int>(blockIdx.x);
  const int t_step = p.diag ? static_cast<int>(gridDim.x) / p.n_tiles : static_cast<int>(gridDim.x);
  const int total_tiles = p.diag ? p.m_tiles : p.m_tiles * p.n_tiles;
  const int own_n = static_cast<int>(blockIdx.x) % p.n_tiles;

  if (warp == 0) {
    if (lane == 0) {
      int ia = 0, ib = 0; uint32_t pa = 0, pb = 0;
      const int cw = p.row_bytes >> 1;   // channels per chunk
      if (p.b_stationary && t_start < total_tiles) {
        // stationary weights: every tile of this CTA uses the same slices (n_tiles == 1, or the CTA's own n-tile)
        mbar_arrive_expect_tx(&bstat_bar, static_cast<uint32_t>(p.c_chunks * NTAPS) * p.b_bytes);
        for (int cc = 0; cc < p.c_chunks; ++cc)
          for (int t = 0; t < NTAPS; ++t)
            tma_load_3d(&tmB, &bstat_bar, sB + (cc * NTAPS + t) * p.b_bytes, cc * cw, p.b_tap[t],
                        p.diag ? own_n * p.block_n : 0);
      }
      for (int tile = t_start; tile < total_tiles; tile += t_step) {
        const int m_tile = p.diag ? tile : tile / p.n_tiles, n_tile = p.diag ? own_n : tile - m_tile * p.n_tiles;
        const int img = m_tile / p.tiles_per_img;
        const int h0 = (m_tile - img * p.tiles_per_img) * p.RT;
        for (int cc = 0; cc < p.c_chunks; ++cc) {
          mbar_wait(&a_empty[ia], pa ^ 1u);
          mbar_arrive_expect_tx(&a_full[ia], p.a_box_bytes);
          tma_load_4d(&tmX, &a_full[ia], sA + ia * p.a_bytes, (p.diag ? n_tile : cc) * cw, -p.pad, h0 - p.pad, img);
          if (++ia == p.sa) { ia = 0; pa ^= 1u; }
          if (p.b_stationary) continue;
          for (int t = 0; t < NTAPS; ++t) {
            mbar_wait(&b_empty[ib], pb ^ 1u);
            mbar_arrive_expect_tx(&b_full[ib], p.b_bytes);
            tma_load_3d(&tmB, &b_full[ib], sB + ib * p.b_bytes, cc * cw, p.b_tap[t], n_tile * p.block_n);
            if (++ib == p.sb) { ib = 0; pb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int ia = 0, ib = 0; uint32_t pa = 0, pb = 0;
      const uint32_t idesc = make_idesc_bf16(kHTileM, p.block_n, 0, 0);
      const uint64_t proto = make_smem_desc(0, 16, 8u * p.row_bytes, layout_type_for_row_bytes(p.row_bytes));
      uint32_t tap_inc[NTAPS];
#pragma unroll
      for (int t = 0; t < NTAPS; ++t) tap_inc[t] = (static_cast<uint32_t>(p.a_off[t]) * p.row_bytes) >> 4;
      if (p.b_stationary && t_start < total_tiles) {
        mbar_wait(&bstat_bar, 0);
        tc_fence_after();
      }
      int local = 0;
      for (int tile = t_start; tile < total_tiles; tile += t_step, ++local) {
        const int acc = local & 1;
        mbar_wait(&tmem_empty[acc], ((local >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int cc = 0; cc < p.c_chunks; ++cc) {
          mbar_wait(&a_full[ia], pa);
          tc_fence_after();
          const uint64_t da0 = proto + (smem_u32(sA + ia * p.a_bytes) >> 4);
          if (p.b_stationary) {
            const uint64_t db0 = proto + (smem_u32(sB + cc * NTAPS * p.b_bytes) >> 4);
            const uint32_t b_inc = p.b_bytes >> 4;
#pragma unroll
            for (int t = 0; t < NTAPS; ++t) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k)
                umma_bf16(d_tmem, da0 + tap_inc[t] + 2 * k, db0 + t * b_inc + 2 * k, idesc,
                          (t | k) != 0 ? 1u : (cc != 0 ? 1u : 0u));
            }
          } else {
#pragma unroll
            for (int t = 0; t < NTAPS; ++t) {
              mbar_wait(&b_full[ib], pb);
              tc_fence_after();
              const uint64_t db0 = proto + (smem_u32(sB + ib * p.b_bytes) >> 4);
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k)
                umma_bf16(d_tmem, da0 + tap_inc[t] + 2 * k, db0 + 2 * k, idesc, (t | k) != 0 ? 1u : (cc != 0 ? 1u : 0u));
              umma_commit(&b_empty[ib]);
              if (++ib == p.sb) { ib = 0; pb ^= 1u; }
            }
          }
          umma_commit(&a_empty[ia]);
          if (++ia == p.sa) { ia = 0; pa ^= 1u; }
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const bool leader = (warp == 2 && lane == 0);
    const int v = q * 32 + lane;            // virtual pixel (accumulator row)
    const int ry = v / p.Wp, cx = v - ry * p.Wp;
    const int srow = ry * p.W + cx;         // row of the staged output tile [RT][W][64]
    const int nbox = p.block_n >> 6;
    const uint32_t box_bytes = static_cast<uint32_t>(p.RT * p.W) * 128u;
    const uint32_t box_pitch = (box_bytes + 1023u) & ~1023u;
    // fused BN statistics: this thread owns one output column and a row range of every tile
    const int st_tid = threadIdx.x - 64;
    const int st_col = st_tid % p.block_n;
    const int st_rows = kHTileM / (256 / p.block_n);
    const int st_row0 = (st_tid / p.block_n) * st_rows;
    int st_ntile = -1;
    float st_s1 = 0.f, st_s2 = 0.f;
    int local = 0;
    for (int tile = t_start; tile < total_tiles; tile += t_step, ++local) {
      const int acc = local & 1;
      const int m_tile = p.diag ? tile : tile / p.n_tiles, n_tile = p.diag ? own_n : tile - m_tile * p.n_tiles;
      const int img = m_tile / p.tiles_per_img;
      const int h0 = (m_tile - img * p.tiles_per_img) * p.RT;
      const int nbase = n_tile * p.block_n;
      const bool valid = (ry < p.RT) && (cx < p.W) && (h0 + ry < p.H);
      if (leader && local > 0) bulk_wait_group_read0();
      named_bar_sync(1, 256);
      if (p.has_res) {
        if (leader) {
          mbar_arrive_expect_tx(&res_bar, static_cast<uint32_t>(nbox) * box_bytes);
          for (int b = 0; b < nbox; ++b) tma_load_4d(&tmR, &res_bar, epi + b * box_pitch, nbase + b * 64, 0, h0, img);
        }
        mbar_wait(&res_bar, static_cast<uint32_t>(local & 1));
      }
      mbar_wait(&tmem_full[acc], (local >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
      for (int c0 = half * 16; c0 < p.block_n; c0 += 32) {
        uint32_t vv[16];
        tmem_ld16(taddr + c0, vv);
        tmem_ld_wait();
        if (valid) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(vv[i]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] += __ldg(p.bias + nbase + c0 + i);
          }
          uint8_t* row = epi + (c0 >> 6) * box_pitch + srow * 128;
          const int j0 = (c0 & 63) >> 3;
          uint4* p0 = reinterpret_cast<uint4*>(row + ((j0 ^ (srow & 7)) << 4));
          uint4* p1 = reinterpret_cast<uint4*>(row + (((j0 + 1) ^ (srow & 7)) << 4));
          if (p.has_res) {
            const uint4 r0 = *p0, r1 = *p1;
            const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 t2 = unpack_bf16x2(rr[i]);
              f[2 * i] += t2.x;
              f[2 * i + 1] += t2.y;
            }
          }
          if (p.act == B200_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
          } else if (p.act == B200_ACT_RELU6) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fminf(fmaxf(f[i], 0.f), 6.f);
          }
          uint4 a, b;
          a.x = pack_bf16x2(f[0], f[1]);   a.y = pack_bf16x2(f[2], f[3]);
          a.z = pack_bf16x2(f[4], f[5]);   a.w = pack_bf16x2(f[6], f[7]);
          b.x = pack_bf16x2(f[8], f[9]);   b.y = pack_bf16x2(f[10], f[11]);
          b.z = pack_bf16x2(f[12], f[13]); b.w = pack_bf16x2(f[14], f[15]);
          *p0 = a;
          *p1 = b;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      fence_proxy_async();
      named_bar_sync(1, 256);
      if (leader) {
        for (int b = 0; b < nbox; ++b) tma_store_4d(&tmC, epi + b * box_pitch, nbase + b * 64, 0, h0, img);
        bulk_commit_group();
      }
      if (p.stats != nullptr) {
        if (st_ntile != n_tile) {
          if (st_ntile >= 0) {
            double* dst = p.stats + (blockIdx.x % kHStatReplicas) * 2 * p.Kout + st_ntile * p.block_n + st_col;
            atomicAdd(dst, (double)st_s1);
            atomicAdd(dst + p.Kout, (double)st_s2);
          }
          st_ntile = n_tile; st_s1 = 0.f; st_s2 = 0.f;
        }
        const int vrows = min(p.RT, p.H - h0) * p.W;   // staged rows that belong to the image
        const uint8_t* col = epi + (st_col >> 6) * box_pitch + (st_col & 7) * 2;
        const int j = (st_col & 63) >> 3;
        const int r_end = min(st_row0 + st_rows, vrows);
        for (int r = st_row0; r < r_end; ++r) {
          const float x = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(col + r * 128 + ((j ^ (r & 7)) << 4)));
          st_s1 += x;
          st_s2 = fmaf(x, x, st_s2);
        }
      }
    }
    if (p.stats != nullptr && st_ntile >= 0) {
      double* dst = p.stats + (blockIdx.x % kHStatReplicas) * 2 * p.Kout + st_ntile * p.block_n + st_col;
      atomicAdd(dst, (double)st_s1);
      atomicAdd(dst + p.Kout, (double)st_s2);
    }
    if (leader) bulk_wait_group0();
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

static int enc4(CUtensorMap* tm, const void* base, int C, int W, int H, int N, int b0, int b1, int b2) {
  EncodeTiledFn fn = encode_tiled_fn();
  B200_REQUIRE(fn != nullptr, B200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(b0 * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled(4d) failed (%d) dims=(%d,%d,%d,%d) box=(%d,%d,%d)",
               (int)r, C, W, H, N, b0, b1, b2);
  return B200_OK;
}

static bool halo_enabled() {
  static const bool enabled = !(getenv("B200_HALO") && atoi(getenv("B200_HALO")) == 0);
  return enabled;
}

// Geometry shared by the fprop/dgrad and wgrad halo kernels: stride-1 RxS taps over an H x W OUTPUT map.
// Supported: 3x3 pad 1 with 64-channel source chunks, and the 4x4 pad 0 stem with a 16-channel source.
bool halo_geometry_ok(int H, int W, int Cs, int R, int S, int pad) {
  if (!halo_enabled()) return false;
  const bool k3 = (R == 3 && S == 3 && pad == 1 && Cs % 64 == 0);
  const bool k4 = (R == 4 && S == 4 && pad == 0 && Cs == 16);
  if (!k3 && !k4) return false;
  const int Wp = W + S - 1;
  if (Wp > 128 || W < 8) return false;
  int RT = 128 / Wp;
  if (RT > H) RT = H;
  const int tiles = (H + RT - 1) / RT;
  return (double)H * W / (128.0 * tiles) >= 0.6;   // fill of the 128 accumulator rows
}

bool halo_eligible(int H, int W, int Cs, int Nout, int R, int S, int pad) {
  if (!halo_geometry_ok(H, W, Cs, R, S, pad) || Nout % 64 != 0) return false;
  const int n_tiles = (Nout + 255) / 256;
  const int block_n = Nout / n_tiles;
  return Nout % n_tiles == 0 && block_n % 64 == 0 && 256 % block_n == 0;
}

// dir 0: fprop (tap (r,s) reads halo offset (r,s), weight slice r*S+s); dir 1: dgrad (offset (R-1-r, S-1-s))
int launch_halo(const void* src, const void* wmat, void* out, const void* res, const float* bias, int N, int H, int W,
                int Cs, int Nout, int R, int S, int pad, int dir, int act, double* stats, cudaStream_t stream, int window) {
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.diag = window > 0 ? 1 : 0;
  if (p.diag)
    B200_REQUIRE(window == 64 && Cs == Nout && Cs % 64 == 0 && R == 3 && S == 3, B200_ERR_UNSUPPORTED,
                 "conv halo: window mode needs window == 64, C == K, C %% 64 == 0 and a 3x3 filter (C=%d K=%d)", Cs, Nout);
  p.N = N; p.H = H; p.W = W; p.C = Cs; p.Kout = Nout;
  p.ntaps = R * S; p.pad = pad;
  p.row_bytes = (Cs == 16) ? 32 : 128;
  const int cw = p.row_bytes / 2;
  p.Wp = W + S - 1;
  p.RT = 128 / p.Wp;
  if (p.RT > H) p.RT = H;
  p.tiles_per_img = (H + p.RT - 1) / p.RT;
  p.m_tiles = N * p.tiles_per_img;
  p.n_tiles = p.diag ? Nout / 64 : (Nout + 255) / 256;
  p.block_n = Nout / p.n_tiles;
  p.c_chunks = p.diag ? 1 : Cs / cw;
  p.a_box_bytes = (uint32_t)(p.RT + R - 1) * p.Wp * p.row_bytes;
  uint32_t a_need = (uint32_t)(kHTileM + (R - 1) * p.Wp + (S - 1)) * p.row_bytes;
  if (a_need < p.a_box_bytes) a_need = p.a_box_bytes;
  p.a_bytes = (a_need + 1023u) & ~1023u;
  p.b_bytes = (uint32_t)p.block_n * p.row_bytes;
  p.act = act;
  p.has_res = res != nullptr;
  p.stats = stats;
  p.bias = bias;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      const int t = r * S + s;
      p.b_tap[t] = (uint16_t)t;
      p.a_off[t] = (uint16_t)(dir == 0 ? r * p.Wp + s : (R - 1 - r) * p.Wp + (S - 1 - s));
    }
  const uint32_t box_pitch = (((uint32_t)(p.RT * W) * 128u) + 1023u) & ~1023u;
  const uint32_t epi_bytes = (uint32_t)(p.block_n / 64) * box_pitch;
  int budget = 212 * 1024 - (int)epi_bytes;
  const int b_all = p.c_chunks * p.ntaps * (int)p.b_bytes;
  p.b_stationary = ((p.n_tiles == 1 || p.diag) && b_all <= 80 * 1024) ? 1 : 0;
  B200_REQUIRE(!p.diag || p.b_stationary, B200_ERR_UNSUPPORTED, "conv halo: window mode expects resident weights");
  int b_region;
  if (p.b_stationary) {
    b_region = b_all;
    p.sb = 0;
    p.sa = (budget - b_region) / (int)p.a_bytes;
    if (p.sa > kHMaxA) p.sa = kHMaxA;
  } else {
    p.sa = 2;
    p.sb = (budget - p.sa * (int)p.a_bytes) / (int)p.b_bytes;
    if (p.sb > kHMaxB) p.sb = kHMaxB;
    if (p.sb >= 6 && (int)(3 * p.a_bytes + 4 * p.b_bytes) <= budget) {
      p.sa = 3;
      p.sb = (budget - 3 * (int)p.a_bytes) / (int)p.b_bytes;
      if (p.sb > kHMaxB) p.sb = kHMaxB;
    }
    b_region = p.sb * (int)p.b_bytes;
  }
  B200_REQUIRE(p.sa >= 2 && (p.b_stationary || p.sb >= 2), B200_ERR_UNSUPPORTED,
               "conv halo: shared memory budget exceeded (W=%d, block_n=%d)", W, p.block_n);
  CUtensorMap tmX, tmB, tmC, tmR;
  memset(&tmR, 0, sizeof(tmR));
  int rc = enc4(&tmX, src, Cs, W + S - 1 - 2 * pad, H + R - 1 - 2 * pad, N, cw, p.Wp, p.RT + R - 1);
  if (rc) return rc;
  {
    EncodeTiledFn fn = encode_tiled_fn();
    const int wc = p.diag ? window : Cs;      // channel extent of the weight operand: [Nout][taps][wc]
    cuuint64_t dims[3] = {(cuuint64_t)wc, (cuuint64_t)p.ntaps, (cuuint64_t)Nout};
    cuuint64_t strides[2] = {(cuuint64_t)wc * 2, (cuuint64_t)wc * p.ntaps * 2};
    cuuint32_t box[3] = {(cuuint32_t)cw, 1, (cuuint32_t)p.block_n};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wmat), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(p.row_bytes),
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "conv halo: weight tensor map failed (%d)", (int)r);
  }
  rc = enc4(&tmC, out, Nout, W, H, N, 64, W, p.RT);
  if (rc) return rc;
  if (res) {
    rc = enc4(&tmR, res, Nout, W, H, N, 64, W, p.RT);
    if (rc) return rc;
  }
  const int smem_bytes = p.sa * (int)p.a_bytes + b_region + (int)epi_bytes + 1024;
  const void* kfn = (p.ntaps == 9) ? (const void*)conv_halo_kernel<9, 4> : (const void*)conv_halo_kernel<16, 1>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  B200_REQUIRE(e == cudaSuccess, B200_ERR_CUDA, "conv halo: smem attribute (%d bytes): %s", smem_bytes, cudaGetErrorString(e));
  const int total = p.m_tiles * p.n_tiles;
  int grid = total < sm_count() ? total : sm_count();
  if (p.diag) {                       // a multiple of n_tiles: every CTA owns one n-tile
    int per_n = sm_count() / p.n_tiles;
    if (per_n > p.m_tiles) per_n = p.m_tiles;
    if (per_n < 1) per_n = 1;
    grid = per_n * p.n_tiles;
  }
  if (p.ntaps == 9)
    b200::launch(conv_halo_kernel<9, 4>, grid, kHThreads, smem_bytes, stream, tmX, tmB, tmC, tmR, p);
  else
    b200::launch(conv_halo_kernel<16, 1>, grid, kHThreads, smem_bytes, stream, tmX, tmB, tmC, tmR, p);
  B200_CHECK_LAUNCH("conv_halo_kernel");
  return B200_OK;
}


// =================================================================================================
// Halo wgrad: dw[k, (r,s), c] += sum_pixels dy[pixel, k] * x[pixel + (r,s), c]   (stride 1)
//
// Both operands are MN-major (rows of shared memory = pixels = the GEMM K dimension).  A tile is RT output rows
// in the padded pitch Wp: dy is loaded as [RT x Wp] pixels (the S-1 extra columns are out of bounds -> TMA writes
// zeros), x as the [(RT+R-1) x Wp] halo, ONCE; the R*S taps are row-shifted views of that one x buffer, each
// accumulating into its own TMEM column block [128 k x cw c].  With cw = 32 (3x3) the 9 taps use 288 of the 512
// TMEM columns; the stem (16 taps x 16 channels) uses 256.  The im2col kernel re-fetched x once per tap.
// One CTA owns a (k-tile, channel-chunk) unit and a contiguous range of pixel tiles (split-K over pixels); partial
// fp32 tiles go to the workspace and conv_halo_wgrad_reduce_kernel adds them into dw in a fixed order.
constexpr int kWThreads = 192;   // TMA warp, MMA warp, 4 epilogue warps
constexpr int kWMaxStages = 4;

struct HaloWgradParams {
  int N, H, W, K_out, C;
  int RT, Wp, tiles_per_img, m_tiles;
  int ntaps, pad, halo_rows;
  int x_row_bytes, cw, ncols;
  int c_chunks, k_tiles, units, splits, tiles_per_split;
  int stages;
  int window;              // 0 dense; 128: block-diagonal -- k-tile t only pairs with input channels [128t, 128t+128):
                           // units = 4 x k_tiles, dw is [K][taps][128]
  uint32_t a_box_bytes, a_bytes, x_box_bytes, x_bytes;
  float* partial;          // [unit][split][128][ncols]
  uint16_t x_off[kHMaxTaps];
};

// The S taps of one filter row are ONE MMA: their x views start one pixel row apart, so the descriptor's
// leading-dimension stride (distance between swizzle atoms along N) is set to one pixel row and N = S * cw.  The
// dy operand (4 KB per MMA) is then read from shared memory once per filter row instead of once per tap -- with
// N = cw the kernel was bound by shared-memory bandwidth, not by the tensor pipe.
template <int R, int S>
__global__ void __launch_bounds__(kWThreads, 1)
conv_halo_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                       const __grid_constant__ HaloWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kWMaxStages], empty_bar[kWMaxStages], acc_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t stage_bytes = p.a_bytes + p.x_bytes;

  // Rows that TMA never writes (dy rows >= RT*Wp, the second dy box when K_out = 64, x rows past the halo box) are
  // multiplied into the accumulators: they must be zero / finite, so the whole ring is cleared once.
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = static_cast<int>(p.stages * stage_bytes) >> 4;
    for (int i = threadIdx.x; i < n16; i += kWThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    fence_mbar_init();
    prefetch_tmap(&tmDy);
    prefetch_tmap(&tmX);
  }
  if (warp == 1) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // prologue above (barriers, TMEM, descriptor prefetch) overlaps the predecessor grid
  const uint32_t tmem_base = tmem_base_s;

  const int split = blockIdx.x / p.units;
  const int unit = blockIdx.x - split * p.units;
  const int k_tile = unit % p.k_tiles;
  const int cc = unit / p.k_tiles;
  const int k0 = k_tile * kHTileM;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(p.m_tiles, t_begin + p.tiles_per_split);
  const int ntiles = t_end - t_begin;
  const int nA = min(2, (p.K_out - k0 + 63) / 64);

  if (ntiles > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0; uint32_t phase = 0;
        const uint32_t tx = nA * p.a_box_bytes + p.x_box_bytes;
        for (int tile = t_begin; tile < t_end; ++tile) {
          const int img = tile / p.tiles_per_img;
          const int h0 = (tile - img * p.tiles_per_img) * p.RT;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          for (int j = 0; j < nA; ++j) tma_load_4d(&tmDy, &full_bar[stage], sa + j * 16384, k0 + j * 64, 0, h0, img);
          tma_load_4d(&tmX, &full_bar[stage], sa + p.a_bytes, cc * p.cw + (p.window ? k0 : 0), -p.pad, h0 - p.pad, img);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        int stage = 0; uint32_t phase = 0;
        const uint32_t idesc = make_idesc_bf16(kHTileM, S * p.cw, 1, 1);
        const uint64_t protoA = make_smem_desc(0, 16384, 1024, 2);
        const uint64_t protoX = make_smem_desc(0, p.x_row_bytes, 8u * p.x_row_bytes, layout_type_for_row_bytes(p.x_row_bytes));
        uint32_t row_inc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) row_inc[r] = (static_cast<uint32_t>(p.x_off[r * S]) * p.x_row_bytes) >> 4;
        const uint32_t kx_inc = (16u * p.x_row_bytes) >> 4;   // 16 pixel rows per K step
        for (int i = 0; i < ntiles; ++i) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint64_t da0 = protoA + (a_addr >> 4);
          const uint64_t dx0 = protoX + ((a_addr + p.a_bytes) >> 4);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t acc = (k != 0) ? 1u : (i != 0 ? 1u : 0u);
#pragma unroll
            for (int r = 0; r < R; ++r)
              umma_bf16(tmem_base + r * S * p.cw, da0 + k * 128, dx0 + row_inc[r] + k * kx_inc, idesc, acc);
          }
          umma_commit(&empty_bar[stage]);
          if (i == ntiles - 1) umma_commit(&acc_bar);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    } else {
      const int q = warp & 3;
      mbar_wait(&acc_bar, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      float* dst = p.partial + ((static_cast<long long>(unit) * p.splits + split) * kHTileM + (q * 32 + lane)) * p.ncols;
      for (int c0 = 0; c0 < p.ncols; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float4*>(dst + c0 + 4 * i) =
              make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                          __uint_as_float(v[4 * i + 3]));
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// dw[k][tap][c] += sum over splits of partial[unit(k_tile, cc)][split][k % 128][tap * cw + c % cw]
// (same block organisation as conv_wgrad_reduce_kernel: 32 float4 outputs per block, 8 warps over the splits)
__global__ void __launch_bounds__(256) conv_halo_wgrad_reduce_kernel(const float* __restrict__ partial,
                                                                     float* __restrict__ dw, int K_out, int ntaps, int C,
                                                                     int cw, int k_tiles, int splits, int ncols) {
  pdl_wait();
  __shared__ float4 red[8][32];
  const int c4n = C >> 2;
  const long long total = static_cast<long long>(K_out) * ntaps * c4n;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (long long base = static_cast<long long>(blockIdx.x) * 32; base < total; base += static_cast<long long>(gridDim.x) * 32) {
    const long long idx = base + lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int c = 0, tap = 0, k = 0;
    if (idx < total) {
      c = static_cast<int>(idx % c4n) * 4;
      tap = static_cast<int>((idx / c4n) % ntaps);
      k = static_cast<int>(idx / (static_cast<long long>(c4n) * ntaps));
      const int k_tile = k >> 7, row = k & 127;
      const int cc = c / cw;
      const int unit = cc * k_tiles + k_tile;
      const float* src = partial + ((static_cast<long long>(unit) * splits) * kHTileM + row) * ncols + tap * cw + (c - cc * cw);
      for (int s2 = w; s2 < splits; s2 += nw) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(src + static_cast<long long>(s2) * kHTileM * ncols));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    red[w][lane] = acc;
    __syncthreads();
    if (w == 0 && idx < total) {
      for (int j = 1; j < nw; ++j) {
        const float4 v = red[j][lane];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      float4* d = reinterpret_cast<float4*>(dw + (static_cast<long long>(k) * ntaps + tap) * C + c);
      float4 o = *d;
      o.x += acc.x; o.y += acc.y; o.z += acc.z; o.w += acc.w;
      *d = o;
    }
    __syncthreads();
  }
}

bool halo_wgrad_eligible(int H, int W, int C, int K_out, int R, int S, int pad) {
  static const bool enabled = !(getenv("B200_HALO_WGRAD") && atoi(getenv("B200_HALO_WGRAD")) == 0);
  if (!enabled || !halo_geometry_ok(H, W, C, R, S, pad)) return false;
  return K_out % 64 == 0;
}

int launch_halo_wgrad(const void* x, const void* dy, float* dw, void* workspace, size_t workspace_bytes, int N, int H,
                      int W, int C, int K_out, int R, int S, int pad, cudaStream_t stream, int window) {
  HaloWgradParams p;
  memset(&p, 0, sizeof(p));
  p.window = window;
  if (window)
    B200_REQUIRE(window == 128 && C == K_out && C % 128 == 0, B200_ERR_UNSUPPORTED,
                 "conv halo wgrad: window mode needs window == 128 and C == K, C %% 128 == 0 (C=%d K=%d)", C, K_out);
  p.N = N; p.H = H; p.W = W; p.K_out = K_out; p.C = C;
  p.ntaps = R * S; p.pad = pad;
  p.cw = (C == 16) ? 16 : 32;
  p.x_row_bytes = p.cw * 2;
  p.ncols = p.ntaps * p.cw;
  p.Wp = W + S - 1;
  p.RT = 128 / p.Wp;
  if (p.RT > H) p.RT = H;
  p.halo_rows = p.RT + R - 1;
  p.tiles_per_img = (H + p.RT - 1) / p.RT;
  p.m_tiles = N * p.tiles_per_img;
  p.c_chunks = (window ? window : C) / p.cw;       // channel chunks that pair with one k-tile
  p.k_tiles = (K_out + kHTileM - 1) / kHTileM;
  p.units = p.c_chunks * p.k_tiles;
  int splits = sm_count() / p.units;
  if (splits < 1) splits = 1;
  if (splits > p.m_tiles) splits = p.m_tiles;
  p.tiles_per_split = (p.m_tiles + splits - 1) / splits;
  p.splits = (p.m_tiles + p.tiles_per_split - 1) / p.tiles_per_split;   // every split owns at least one tile
  p.a_box_bytes = (uint32_t)(p.RT * p.Wp) * 128u;
  p.a_bytes = 2 * 16384;
  p.x_box_bytes = (uint32_t)(p.halo_rows * p.Wp) * p.x_row_bytes;
  int max_off = 0;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      p.x_off[r * S + s] = (uint16_t)(r * p.Wp + s);
      if (r * p.Wp + s > max_off) max_off = r * p.Wp + s;
    }
  uint32_t x_need = (uint32_t)(kHTileM + max_off) * p.x_row_bytes;
  if (x_need < p.x_box_bytes) x_need = p.x_box_bytes;
  p.x_bytes = (x_need + 1023u) & ~1023u;
  p.stages = (212 * 1024) / (int)(p.a_bytes + p.x_bytes);
  if (p.stages > kWMaxStages) p.stages = kWMaxStages;
  B200_REQUIRE(p.stages >= 2, B200_ERR_UNSUPPORTED, "conv halo wgrad: shared memory budget exceeded (W=%d)", W);
  const size_t need = (size_t)p.units * p.splits * kHTileM * p.ncols * sizeof(float);
  B200_REQUIRE(workspace != nullptr && workspace_bytes >= need, B200_ERR_INVALID,
               "conv halo wgrad: workspace of %zu bytes needed (%zu given)", need, workspace_bytes);
  p.partial = reinterpret_cast<float*>(workspace);
  CUtensorMap tmDy, tmX;
  int rc = enc4(&tmDy, dy, K_out, W, H, N, 64, p.Wp, p.RT);
  if (rc) return rc;
  rc = enc4(&tmX, x, C, W + S - 1 - 2 * pad, H + R - 1 - 2 * pad, N, p.cw, p.Wp, p.halo_rows);
  if (rc) return rc;
  const int smem_bytes = p.stages * (int)(p.a_bytes + p.x_bytes) + 1024;
  const void* kfn = (p.ntaps == 9) ? (const void*)conv_halo_wgrad_kernel<3, 3> : (const void*)conv_halo_wgrad_kernel<4, 4>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  B200_REQUIRE(e == cudaSuccess, B200_ERR_CUDA, "conv halo wgrad: smem attribute (%d bytes): %s", smem_bytes,
               cudaGetErrorString(e));
  if (p.ntaps == 9)
    b200::launch(conv_halo_wgrad_kernel<3, 3>, p.units * p.splits, kWThreads, smem_bytes, stream, tmDy, tmX, p);
  else
    b200::launch(conv_halo_wgrad_kernel<4, 4>, p.units * p.splits, kWThreads, smem_bytes, stream, tmDy, tmX, p);
  B200_CHECK_LAUNCH("conv_halo_wgrad_kernel");
  const int Cw = window ? window : C;                 // row length of dw: [K][taps][Cw]
  const long long total = (long long)K_out * p.ntaps * (Cw / 4);
  long long blocks64 = (total + 31) / 32;
  if (blocks64 > 16LL * sm_count()) blocks64 = 16LL * sm_count();
  const int blocks = (int)blocks64;
  b200::launch(conv_halo_wgrad_reduce_kernel, blocks, 32 * wgrad_reduce_warps(p.splits), 0, stream, p.partial, dw, K_out, p.ntaps, Cw, p.cw, p.k_tiles, p.splits,
                                                            p.ncols);
  B200_CHECK_LAUNCH("conv_halo_wgrad_reduce_kernel");
  return B200_OK;
}

}  // namespace b200
