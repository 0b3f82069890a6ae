// 3x3 / stride-1 / pad-1 convolution (fprop and dgrad) as a "shift GEMM" on tcgen05 tensor cores.
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
// Same warp roles / TMEM double buffering / TMA-store epilogue / fused BN statistics as conv_igemm_kernel.
#include "common.cuh"
#include "host.h"
#include <stdlib.h>

namespace b200 {

constexpr int kHThreads = 320;
constexpr int kHTileM = 128;
constexpr int kHMaxA = 4, kHMaxB = 8;
constexpr int kHStatReplicas = 16;

struct HaloParams {
  int N, H, W, C, Kout;
  int RT, Wp;              // image rows per tile, padded row pitch W+2
  int tiles_per_img, m_tiles, n_tiles, block_n, c_chunks;
  int sa, sb;              // ring depths
  uint32_t a_bytes, a_box_bytes, b_bytes;
  int act;
  int has_res;
  double* stats;
  uint16_t a_off[9];       // pixel-row offset of each tap's view inside the halo buffer
  uint16_t b_tap[9];       // weight tap slice used with it
};

__global__ void __launch_bounds__(kHThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kHMaxA], a_empty[kHMaxA], b_full[kHMaxB], b_empty[kHMaxB];
  __shared__ __align__(8) uint64_t tmem_full[2], tmem_empty[2], res_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sA = smem;
  uint8_t* sB = sA + p.sa * p.a_bytes;
  uint8_t* epi = sB + p.sb * p.b_bytes;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.sa; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.sb; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
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
  const uint32_t tmem_base = tmem_base_s;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    if (lane == 0) {
      int ia = 0, ib = 0; uint32_t pa = 0, pb = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
        const int img = m_tile / p.tiles_per_img;
        const int h0 = (m_tile - img * p.tiles_per_img) * p.RT;
        for (int cc = 0; cc < p.c_chunks; ++cc) {
          mbar_wait(&a_empty[ia], pa ^ 1u);
          mbar_arrive_expect_tx(&a_full[ia], p.a_box_bytes);
          tma_load_4d(&tmX, &a_full[ia], sA + ia * p.a_bytes, cc * 64, -1, h0 - 1, img);
          if (++ia == p.sa) { ia = 0; pa ^= 1u; }
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&b_empty[ib], pb ^ 1u);
            mbar_arrive_expect_tx(&b_full[ib], p.b_bytes);
            tma_load_3d(&tmB, &b_full[ib], sB + ib * p.b_bytes, cc * 64, p.b_tap[t], n_tile * p.block_n);
            if (++ib == p.sb) { ib = 0; pb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int ia = 0, ib = 0; uint32_t pa = 0, pb = 0;
      const uint32_t idesc = make_idesc_bf16(kHTileM, p.block_n, 0, 0);
      int local = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
        const int acc = local & 1;
        mbar_wait(&tmem_empty[acc], ((local >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int cc = 0; cc < p.c_chunks; ++cc) {
          mbar_wait(&a_full[ia], pa);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + ia * p.a_bytes);
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&b_full[ib], pb);
            tc_fence_after();
            const uint32_t a_addr = a_base + static_cast<uint32_t>(p.a_off[t]) * 128u;
            const uint32_t b_addr = smem_u32(sB + ib * p.b_bytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = make_smem_desc(a_addr + k * 32, 16, 1024, 2);
              const uint64_t db = make_smem_desc(b_addr + k * 32, 16, 1024, 2);
              umma_bf16(d_tmem, da, db, idesc, (cc | t | k) != 0 ? 1u : 0u);
            }
            umma_commit(&b_empty[ib]);
            if (++ib == p.sb) { ib = 0; pb ^= 1u; }
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
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
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
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled(4d) failed (%d) dims=(%d,%d,%d,%d) box=(%d,%d,%d)",
               (int)r, C, W, H, N, b0, b1, b2);
  return B200_OK;
}

// Eligibility of the halo path for a 3x3/s1/p1 problem with `Cs` source channels, `Nout` produced channels on an
// H x W map.  Tiles are RT rows; require a reasonable fill of the 128 accumulator rows.
bool halo_eligible(int H, int W, int Cs, int Nout) {
  static const bool enabled = !(getenv("B200_HALO") && atoi(getenv("B200_HALO")) == 0);
  if (!enabled) return false;
  if (Cs % 64 != 0 || Nout % 64 != 0 || W + 2 > 128 || W < 8) return false;
  const int n_tiles = (Nout + 255) / 256;
  const int block_n = Nout / n_tiles;
  if (Nout % n_tiles != 0 || block_n % 64 != 0 || 256 % block_n != 0) return false;
  const int Wp = W + 2;
  int RT = 128 / Wp;
  if (RT > H) RT = H;
  const int tiles = (H + RT - 1) / RT;
  const double fill = (double)H * W / (128.0 * tiles);
  return fill >= 0.6;
}

// dir 0: fprop taps (offset r,s <-> weight tap r*3+s); dir 1: dgrad (offset 2-r, 2-s <-> weight tap r*3+s)
int launch_halo(const void* src, const void* wmat, void* out, const void* res, int N, int H, int W, int Cs, int Nout,
                int dir, int act, double* stats, cudaStream_t stream) {
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W; p.C = Cs; p.Kout = Nout;
  p.Wp = W + 2;
  p.RT = 128 / p.Wp;
  if (p.RT > H) p.RT = H;
  p.tiles_per_img = (H + p.RT - 1) / p.RT;
  p.m_tiles = N * p.tiles_per_img;
  p.n_tiles = (Nout + 255) / 256;
  p.block_n = Nout / p.n_tiles;
  p.c_chunks = Cs / 64;
  p.a_box_bytes = (uint32_t)(p.RT + 2) * p.Wp * 128u;
  uint32_t a_need = (uint32_t)(kHTileM + 2 * p.Wp + 2) * 128u;
  if (a_need < p.a_box_bytes) a_need = p.a_box_bytes;
  p.a_bytes = (a_need + 1023u) & ~1023u;
  p.b_bytes = (uint32_t)p.block_n * 128u;
  p.act = act;
  p.has_res = res != nullptr;
  p.stats = stats;
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      const int t = r * 3 + s;
      p.b_tap[t] = (uint16_t)t;
      p.a_off[t] = (uint16_t)(dir == 0 ? r * p.Wp + s : (2 - r) * p.Wp + (2 - s));
    }
  const uint32_t box_pitch = (((uint32_t)(p.RT * W) * 128u) + 1023u) & ~1023u;
  const uint32_t epi_bytes = (uint32_t)(p.block_n / 64) * box_pitch;
  const int budget = 212 * 1024 - (int)epi_bytes;
  p.sa = 2;
  p.sb = (budget - p.sa * (int)p.a_bytes) / (int)p.b_bytes;
  if (p.sb > kHMaxB) p.sb = kHMaxB;
  B200_REQUIRE(p.sb >= 2, B200_ERR_UNSUPPORTED, "conv3x3 halo: shared memory budget exceeded (W=%d, block_n=%d)", W, p.block_n);
  if (p.sb >= 6 && (int)(3 * p.a_bytes + 4 * p.b_bytes) <= budget) { p.sa = 3; p.sb = (budget - 3 * (int)p.a_bytes) / (int)p.b_bytes; if (p.sb > kHMaxB) p.sb = kHMaxB; }
  CUtensorMap tmX, tmB, tmC, tmR;
  memset(&tmR, 0, sizeof(tmR));
  int rc = enc4(&tmX, src, Cs, W, H, N, 64, p.Wp, p.RT + 2);
  if (rc) return rc;
  {
    EncodeTiledFn fn = encode_tiled_fn();
    cuuint64_t dims[3] = {(cuuint64_t)Cs, 9, (cuuint64_t)Nout};
    cuuint64_t strides[2] = {(cuuint64_t)Cs * 2, (cuuint64_t)Cs * 9 * 2};
    cuuint32_t box[3] = {64, 1, (cuuint32_t)p.block_n};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wmat), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "conv3x3 halo: weight tensor map failed (%d)", (int)r);
  }
  rc = enc4(&tmC, out, Nout, W, H, N, 64, W, p.RT);
  if (rc) return rc;
  if (res) {
    rc = enc4(&tmR, res, Nout, W, H, N, 64, W, p.RT);
    if (rc) return rc;
  }
  const int smem_bytes = p.sa * (int)p.a_bytes + p.sb * (int)p.b_bytes + (int)epi_bytes + 1024;
  cudaError_t e = cudaFuncSetAttribute((const void*)conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  B200_REQUIRE(e == cudaSuccess, B200_ERR_CUDA, "conv3x3 halo: smem attribute (%d bytes): %s", smem_bytes, cudaGetErrorString(e));
  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < sm_count() ? total : sm_count();
  conv_halo_kernel<<<grid, kHThreads, smem_bytes, stream>>>(tmX, tmB, tmC, tmR, p);
  B200_CHECK_LAUNCH("conv_halo_kernel");
  return B200_OK;
}

}  // namespace b200
