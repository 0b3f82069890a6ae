// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a): fprop, dgrad, wgrad.
//
//   fprop : y[m, k]   = sum_{tap, c} x_im2col[m, tap, c] * w[k, tap, c]        (A K-major, B K-major)
//   dgrad : dx[m, c]  = sum_{tap, k} dy_im2col[m, tap, k] * wt[c, tap, k]       (same kernel, tap table)
//   wgrad : dw[k, tap, c] += sum_{pix} dy[pix, k] * x_im2col[pix, tap, c]       (A MN-major, B MN-major)
//
// Operand tiles are staged by TMA (im2col mode for the activation side, tiled mode for weights / dy)
// into 32/64/128B-swizzled shared memory, consumed by single-thread-issued tcgen05.mma with fp32
// accumulators in TMEM, drained by 4 epilogue warps with tcgen05.ld.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue.
//
// Replaces cuDNN's convolution behind nn.Conv2d in the reference (models/resnet.py:75-78,126-132,
// 226-227) and its autograd backward (trainer.py:162).
#include "common.cuh"
#include "host.h"
#include <stdlib.h>

namespace b200 {

constexpr int kMaxStages = 8;
constexpr int kMaxTaps = 32;
constexpr int kThreads = 192;        // wgrad: TMA warp, MMA warp, 4 epilogue warps
constexpr int kIgemmThreads = 320;   // igemm: TMA warp, MMA warp, 8 epilogue warps (2 per TMEM lane quarter)
constexpr int kTileM = 128;
constexpr uint32_t kTmemCols = 512;
constexpr int kStatReplicas = 16;    // must match kReplicas of bn.cu (layout of the BN workspace accumulators)

struct TapEntry {
  uint16_t off_w, off_h;  // im2col filter offsets (added to the base pixel)
  uint16_t b_tap;         // which tap slice of the weight operand
  uint16_t pad_;
};

struct IgemmParams {
  int M_total;           // rows of the implicit GEMM = Nimg * I * J
  int I, J;              // base-pixel grid per image
  int trav;              // traversal stride of base pixels in the source tensor
  int lower_w, lower_h;  // source coordinate of base pixel (0,0)
  int N_total;           // GEMM N (channels produced)
  int block_n, n_tiles, m_tiles;
  int ck, c_chunks;      // channels per k-block, number of k-blocks per tap
  int ntaps;
  int num_stages;
  uint32_t a_bytes, b_bytes, tx_bytes;
  // output mapping: row m=(n,i,j) -> out pixel (n, i*os+oh0, j*os+ow0) of an [Nimg,OH,OW,ldo] tensor
  int OH, OW, os, oh0, ow0;
  int ldo;
  int act, out_fp32;
  int tma_store;         // 1: dense bf16 output staged in smem and written by TMA (tmC), residual via tmR
  int plain_a;           // 1: A is a dense [M_total, SC] matrix (1x1, stride 1, no padding): tiled TMA
  int b_stationary;      // 1: every (tap, k-block) weight slice stays in shared memory (small 1x1 layers): only A streams
  int own_ntile;         // 1 (with b_stationary): the CTA owns n-tile blockIdx.x % n_tiles -- its weight slices are loaded
                         // once -- and walks the m-tiles blockIdx.x / n_tiles, + gridDim.x / n_tiles, ... (the grid is a
                         // multiple of n_tiles).  Cuts the L2->SM re-streaming of the weights from once per tile to once
                         // per CTA for layers whose per-n-tile weights fit in shared memory (K <= 256 at N-tile 256)
  int epi_bufs;          // 1 or 2 output staging tiles: with 2 the residual tile of the NEXT tile is fetched while this one
                         // is converted and stored, and a store never waits for the previous one (write-heavy epilogues)
  uint32_t epi_bytes;
  int window;            // > 0: block-diagonal convolution -- n-tile b (block_n == window) reads source channels
                         // [window*b, window*b + window) only; the weight operand is [N_total][taps][window]
  double* stats;         // fused BN statistics accumulators [kStatReplicas][2][N_total] (BN workspace) or nullptr
  void* out;
  const void* res;
  const float* bias;
  TapEntry taps[kMaxTaps];
};

__device__ __forceinline__ void store_chunk16(const IgemmParams& p, const uint32_t (&v)[16], long long off,
                                              int n0, bool row_ok) {
  // v: 16 consecutive fp32 accumulators (as bits) for columns [n0, n0+16) of this thread's row
  if (!row_ok) return;
  float f[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
  const bool full = (n0 + 16 <= p.N_total);
  if (p.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (n0 + i < p.N_total) f[i] += __ldg(p.bias + n0 + i);
  }
  if (p.res != nullptr) {
    const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(p.res) + off + n0;
    if (full && ((p.ldo & 7) == 0)) {
      const uint4 r0 = *reinterpret_cast<const uint4*>(r);
      const uint4 r1 = *reinterpret_cast<const uint4*>(r + 8);
      const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float2 t = unpack_bf16x2(rr[i]);
        f[2 * i] += t.x;
        f[2 * i + 1] += t.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (n0 + i < p.N_total) f[i] += __bfloat162float(r[i]);
    }
  }
  if (p.act == B200_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
  } else if (p.act == B200_ACT_RELU6) {
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = fminf(fmaxf(f[i], 0.f), 6.f);
  }
  if (p.out_fp32) {
    float* o = reinterpret_cast<float*>(p.out) + off + n0;
    if (full && ((p.ldo & 3) == 0)) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(o + 4 * i) = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (n0 + i < p.N_total) o[i] = f[i];
    }
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off + n0;
    if (full && ((p.ldo & 7) == 0)) {
      uint4 a, b;
      a.x = pack_bf16x2(f[0], f[1]);   a.y = pack_bf16x2(f[2], f[3]);
      a.z = pack_bf16x2(f[4], f[5]);   a.w = pack_bf16x2(f[6], f[7]);
      b.x = pack_bf16x2(f[8], f[9]);   b.y = pack_bf16x2(f[10], f[11]);
      b.z = pack_bf16x2(f[12], f[13]); b.w = pack_bf16x2(f[14], f[15]);
      *reinterpret_cast<uint4*>(o) = a;
      *reinterpret_cast<uint4*>(o + 8) = b;
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (n0 + i < p.N_total) o[i] = __float2bfloat16(f[i]);
    }
  }
}

__global__ void __launch_bounds__(kIgemmThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                  const __grid_constant__ IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full[2];
  __shared__ __align__(8) uint64_t tmem_empty[2];
  __shared__ __align__(8) uint64_t res_bar[2], bstat_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t stage_bytes = p.b_stationary ? p.a_bytes : p.a_bytes + p.b_bytes;
  uint8_t* sBstat = smem + p.num_stages * stage_bytes;   // stationary weight slices [tap * c_chunks + k-block]
  uint8_t* epi_base = sBstat + (p.b_stationary ? static_cast<uint32_t>(p.ntaps * p.c_chunks) * p.b_bytes : 0u);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&bstat_bar, 1);
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], 8);
    mbar_init(&tmem_empty[1], 8);
    mbar_init(&res_bar[0], 1);
    mbar_init(&res_bar[1], 1);
    fence_mbar_init();
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.tma_store) {
      prefetch_tmap(&tmC);
      if (p.res != nullptr) prefetch_tmap(&tmR);
    }
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // prologue above (barriers, TMEM, descriptor prefetch) overlaps the predecessor grid
  const uint32_t tmem_base = tmem_base_s;

  const int total_tiles = p.m_tiles * p.n_tiles;
  const int k_iters = p.ntaps * p.c_chunks;
  // tile walk: round-robin over all (m, n) tiles, or -- owned n-tile -- over the m-tiles of one n-tile
  const int walk_first = p.own_ntile ? static_cast<int>(blockIdx.x) / p.n_tiles : static_cast<int>(blockIdx.x);
  const int walk_step = p.own_ntile ? static_cast<int>(gridDim.x) / p.n_tiles : static_cast<int>(gridDim.x);
  const int walk_end = p.own_ntile ? p.m_tiles : total_tiles;
  const int own_n = static_cast<int>(blockIdx.x) % p.n_tiles;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int IJ = p.I * p.J;
      if (p.b_stationary && walk_first < walk_end) {   // the same slices serve every tile of the CTA
        mbar_arrive_expect_tx(&bstat_bar, static_cast<uint32_t>(k_iters) * (p.tx_bytes - p.a_bytes));
        for (int t = 0; t < p.ntaps; ++t)
          for (int cc = 0; cc < p.c_chunks; ++cc)
            tma_load_3d(&tmB, &bstat_bar, sBstat + (t * p.c_chunks + cc) * p.b_bytes, cc * p.ck, p.taps[t].b_tap,
                        p.own_ntile ? own_n * p.block_n : 0);
      }
      for (int tile = walk_first; tile < walk_end; tile += walk_step) {
        const int m_tile = p.own_ntile ? tile : tile / p.n_tiles;
        const int n_tile = p.own_ntile ? own_n : tile - m_tile * p.n_tiles;
        const int m0 = m_tile * kTileM;
        const int img = m0 / IJ;
        const int rem = m0 - img * IJ;
        const int bi = rem / p.J;
        const int bj = rem - bi * p.J;
        const int base_w = bj * p.trav + p.lower_w;
        const int base_h = bi * p.trav + p.lower_h;
        for (int t = 0; t < p.ntaps; ++t) {
          const TapEntry te = p.taps[t];
          for (int cc = 0; cc < p.c_chunks; ++cc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + stage * stage_bytes;
            uint8_t* sb = sa + p.a_bytes;
            mbar_arrive_expect_tx(&full_bar[stage], p.b_stationary ? p.a_bytes : p.tx_bytes);
            const int a_c = cc * p.ck + (p.window ? n_tile * p.window : 0);
            if (p.plain_a)  // 1x1 / stride 1: the A operand is a dense [M, C] matrix -> tiled TMA (faster than im2col)
              tma_load_2d(&tmA, &full_bar[stage], sa, a_c, m0);
            else
              tma_load_im2col_4d(&tmA, &full_bar[stage], sa, a_c, base_w, base_h, img, te.off_w, te.off_h);
            if (!p.b_stationary) tma_load_3d(&tmB, &full_bar[stage], sb, cc * p.ck, te.b_tap, n_tile * p.block_n);
            if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idesc = make_idesc_bf16(kTileM, p.block_n, 0, 0);
      const uint32_t row_bytes = p.ck * 2;
      // The issuing thread is latency-bound per instruction: build the descriptor once and advance its address field
      // (bytes >> 4) instead of re-encoding it for every MMA.
      const uint64_t proto = make_smem_desc(0, 16, 8 * row_bytes, layout_type_for_row_bytes(row_bytes));
      const int ksteps = p.ck / 16;
      if (p.b_stationary && walk_first < walk_end) {
        mbar_wait(&bstat_bar, 0);
        tc_fence_after();
      }
      const uint32_t bstat_addr = smem_u32(sBstat);
      int local = 0;
      for (int tile = walk_first; tile < walk_end; tile += walk_step, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1u;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint64_t da = proto + (a_addr >> 4);
          const uint64_t db = proto + ((p.b_stationary ? bstat_addr + it * p.b_bytes : a_addr + p.a_bytes) >> 4);
          if (ksteps == 4) {
            umma_bf16(d_tmem, da, db, idesc, it != 0 ? 1u : 0u);
            umma_bf16(d_tmem, da + 2, db + 2, idesc, 1u);
            umma_bf16(d_tmem, da + 4, db + 4, idesc, 1u);
            umma_bf16(d_tmem, da + 6, db + 6, idesc, 1u);
          } else {
            for (int k = 0; k < ksteps; ++k) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (it == k_iters - 1) umma_commit(&tmem_full[acc]);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;  // the two warps of a quarter take alternate 16-column chunks
    // fused-statistics bookkeeping: this thread owns column st_col, rows [st_row0, st_row0 + st_rows) of each tile
    const int st_tid = threadIdx.x - 64;
    const int st_col = st_tid % p.block_n;
    const int st_rows = kTileM / (256 / p.block_n);
    const int st_row0 = (st_tid / p.block_n) * st_rows;
    int st_ntile = -1;
    float st_s1 = 0.f, st_s2 = 0.f;
    const int IJ = p.I * p.J;
    int local = 0;
    for (int tile = walk_first; tile < walk_end; tile += walk_step, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1u;
      const int m_tile = p.own_ntile ? tile : tile / p.n_tiles;
      const int n_tile = p.own_ntile ? own_n : tile - m_tile * p.n_tiles;
      const int m = m_tile * kTileM + q * 32 + lane;
      const bool row_ok = m < p.M_total;
      long long off = 0;
      if (row_ok) {
        const int img = m / IJ;
        const int rem = m - img * IJ;
        const int bi = rem / p.J;
        const int bj = rem - bi * p.J;
        off = ((static_cast<long long>(img) * p.OH + (bi * p.os + p.oh0)) * p.OW + (bj * p.os + p.ow0)) *
              static_cast<long long>(p.ldo);
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
      const int nbase = n_tile * p.block_n;
      if (p.tma_store) {
        // Dense bf16 output: stage the tile in 128B-swizzled shared memory and write it with TMA (coalesced,
        // clipped at the M tail); the residual tile is fetched by TMA into the same buffer and updated in place.
        const bool leader = (warp == 2 && lane == 0);
        const int row = q * 32 + lane;
        const int nbox = p.block_n >> 6;
        // staging buffer of this tile; with two buffers tile i uses buffer i & 1
        const int eb = p.epi_bufs == 2 ? (local & 1) : 0;
        uint8_t* epi = epi_base + eb * p.epi_bytes;
        if (leader) {
          // the buffer about to be (re)written must no longer be read by an earlier TMA store: one buffer -> the
          // previous store; two buffers -> the store before the previous one, unless the NEXT tile's residual is
          // prefetched below into the buffer the previous store used
          if (local > 0) {
            if (p.epi_bufs == 2 && p.res == nullptr) { if (local > 1) bulk_wait_group_read1(); }
            else bulk_wait_group_read0();
          }
          if (p.res != nullptr) {
            auto fetch = [&](int t_tile, int buf) {
              const int mt = p.own_ntile ? t_tile : t_tile / p.n_tiles;
              const int nt = p.own_ntile ? own_n : t_tile - mt * p.n_tiles;
              uint8_t* dst = epi_base + buf * p.epi_bytes;
              mbar_arrive_expect_tx(&res_bar[buf], static_cast<uint32_t>(nbox) * kTileM * 128u);
              for (int b = 0; b < nbox; ++b)
                tma_load_2d(&tmR, &res_bar[buf], dst + b * (kTileM * 128), nt * p.block_n + b * 64, mt * kTileM);
            };
            if (p.epi_bufs == 2) {
              if (local == 0) fetch(tile, 0);
              const int next = tile + walk_step;
              if (next < walk_end) fetch(next, eb ^ 1);         // lands while this tile is converted and stored
            } else {
              fetch(tile, 0);
            }
          }
        }
        named_bar_sync(1, 256);
        if (p.res != nullptr) {
          // buffer b receives tiles b, b+2, b+4, ... (two buffers) or every tile (one buffer): k-th use -> parity k & 1
          const uint32_t use = p.epi_bufs == 2 ? static_cast<uint32_t>(local >> 1) : static_cast<uint32_t>(local);
          mbar_wait(&res_bar[eb], use & 1u);
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        for (int c0 = half * 16; c0 < p.block_n; c0 += 32) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] += __ldg(p.bias + nbase + c0 + i);
          }
          uint8_t* box = epi + (c0 >> 6) * (kTileM * 128) + row * 128;
          const int j0 = (c0 & 63) >> 3;  // 16B chunk index inside the 128B row
          uint4* p0 = reinterpret_cast<uint4*>(box + (((j0) ^ (row & 7)) << 4));
          uint4* p1 = reinterpret_cast<uint4*>(box + (((j0 + 1) ^ (row & 7)) << 4));
          if (p.res != nullptr) {
            const uint4 r0 = *p0, r1 = *p1;
            const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 t = unpack_bf16x2(rr[i]);
              f[2 * i] += t.x;
              f[2 * i + 1] += t.y;
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
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);   // TMEM stage is free: the MMA warp may start tile+2
        fence_proxy_async();                             // generic-proxy smem writes -> visible to TMA
        named_bar_sync(1, 256);
        if (leader) {
          for (int b = 0; b < nbox; ++b)
            tma_store_2d(&tmC, epi + b * (kTileM * 128), nbase + b * 64, m_tile * kTileM);
          bulk_commit_group();
        }
        if (p.stats != nullptr) {
          // Fused BN statistics: per-channel sum / sum of squares of the bf16-rounded outputs of this tile,
          // read back from the staged tile (rows beyond M_total are exact zeros).  Accumulated in registers
          // across the tiles of this CTA while it stays on the same channel block, then one fp64 atomic each.
          if (st_ntile != n_tile) {
            if (st_ntile >= 0) {
              double* dst = p.stats + (blockIdx.x % kStatReplicas) * 2 * p.N_total + st_ntile * p.block_n + st_col;
              atomicAdd(dst, (double)st_s1);
              atomicAdd(dst + p.N_total, (double)st_s2);
            }
            st_ntile = n_tile; st_s1 = 0.f; st_s2 = 0.f;
          }
          const uint8_t* col = epi + (st_col >> 6) * (kTileM * 128) + (st_col & 7) * 2;
          const int j = (st_col & 63) >> 3;
          const int r_end = st_row0 + st_rows;
#pragma unroll 8
          for (int r = st_row0; r < r_end; ++r) {
            const float vv = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(col + r * 128 + ((j ^ (r & 7)) << 4)));
            st_s1 += vv;
            st_s2 = fmaf(vv, vv, st_s2);
          }
        }
        continue;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      for (int c0 = half * 16; c0 < p.block_n; c0 += 32) {
        uint32_t v0[16];
        tmem_ld16(taddr + c0, v0);
        tmem_ld_wait();
        store_chunk16(p, v0, off, nbase + c0, row_ok && (nbase + c0 < p.N_total));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (p.stats != nullptr && st_ntile >= 0) {
      double* dst = p.stats + (blockIdx.x % kStatReplicas) * 2 * p.N_total + st_ntile * p.block_n + st_col;
      atomicAdd(dst, (double)st_s1);
      atomicAdd(dst + p.N_total, (double)st_s2);
    }
    if (p.tma_store && warp == 2 && lane == 0) bulk_wait_group0();  // smem must outlive the last TMA store
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad: dw[k, tap, c] += sum_pix dy[pix, k] * x[pix @ tap, c]
// A = dy tile, MN-major (rows of smem = pixels, 128/64/32B of k-channels); B = im2col(x) tile, MN-major.
struct WgradParams {
  int M_total;          // fwd output pixels N*P*Q
  int P, Q;
  int trav;             // conv stride
  int lower_w, lower_h; // -pad
  int K_out, C, taps_total;
  int ckA, ckB;         // channels per smem box on the dy side / x side
  int bk;               // pixels per stage
  int c_chunks;         // ceil(C / ckB)
  int total_boxes;      // taps * c_chunks
  int boxes_per_cta;    // <= 512 / (kt * ckB) and <= 8
  int kt, k_groups;     // k-tiles (128 output channels each) accumulated side by side in TMEM by one CTA: the x tile is
                        // fetched once for all of them (L2 -> SM delivery, not the tensor pipe, bounds these kernels)
  int k_tiles, col_groups, splits;
  int blocks_per_split, total_blocks;  // in units of bk pixels
  int num_stages;
  uint32_t boxA_bytes, boxB_bytes, stage_bytes;
  float* dw;
  int plain_x;           // 1: x is a dense [M_total, C] matrix (1x1 stride 1): tiled TMA instead of im2col
  float* partial;        // split-K partial tiles [tile][split][128][pitch] (nullptr: splits == 1, add into dw)
  int pitch;             // kt * boxes_per_cta * ckB
  int window;            // 0 dense; 128: k-tile t pairs with source channels [128t, 128t+128) only (kt == 1)
  int S_filter;          // filter width: tap t = (r, s) = (t / S, t % S) gives the im2col offsets {s, r}.  Computed, not
                         // read from a table: ptxas 12.9 (sm_100a) mis-split a 32-bit {off_w, off_h} word loaded through
                         // the uniform datapath (LDCU + UPRMT on a stale register) -- wrong off_h for every tap row > 0
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                  const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t acc_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_bar, 1);
    fence_mbar_init();
    prefetch_tmap(&tmDy);
    prefetch_tmap(&tmX);
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // prologue above (barriers, TMEM, descriptor prefetch) overlaps the predecessor grid
  const uint32_t tmem_base = tmem_base_s;

  // work decomposition
  const int tiles = p.k_groups * p.col_groups;
  const int split = blockIdx.x / tiles;
  const int tile = blockIdx.x - split * tiles;
  const int k_group = tile % p.k_groups;
  const int cgroup = tile / p.k_groups;
  const int k0 = k_group * p.kt * kTileM;
  const int kt_valid = min(p.kt, p.k_tiles - k_group * p.kt);
  const int box0 = cgroup * p.boxes_per_cta;
  const int nboxes = min(p.boxes_per_cta, p.total_boxes - box0);
  const int blk_begin = split * p.blocks_per_split;
  const int blk_end = min(p.total_blocks, blk_begin + p.blocks_per_split);
  const int nblk = blk_end - blk_begin;
  const uint32_t a_tile = (kTileM / p.ckA) * p.boxA_bytes;   // one k-tile of dy: 128 channels x bk pixels
  const uint32_t a_region = p.kt * a_tile;

  if (nblk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const int PQ = p.P * p.Q;
        // everything that does not depend on the pixel block is computed once: the producer thread's issue rate bounds
        // the HBM-bound layers (one TMA request per 8 KB box)
        int nA_j[4];
        int nA_total = 0;
        for (int j = 0; j < 4; ++j) {
          nA_j[j] = j < kt_valid ? min(kTileM / p.ckA, (p.K_out - (k0 + j * kTileM) + p.ckA - 1) / p.ckA) : 0;
          nA_total += nA_j[j];
        }
        int box_c[8];
        uint16_t box_w[8], box_h[8];
        for (int x = 0; x < 8; ++x) {
          const int id = box0 + min(x, nboxes - 1);
          const int t = id / p.c_chunks;
          const int th = t / p.S_filter;
          box_c[x] = (id - t * p.c_chunks) * p.ckB + (p.window ? k0 : 0);
          box_w[x] = static_cast<uint16_t>(t - th * p.S_filter);
          box_h[x] = static_cast<uint16_t>(th);
        }
        const uint32_t tx = nA_total * p.boxA_bytes + nboxes * p.boxB_bytes;
        for (int b = blk_begin; b < blk_end; ++b) {
          const int pix0 = b * p.bk;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * p.stage_bytes;
          uint8_t* sb = sa + a_region;
          mbar_arrive_expect_tx(&full_bar[stage], tx);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int kj = k0 + j * kTileM;
            for (int a = 0; a < nA_j[j]; ++a)      // dy boxes actually present (K_out tail)
              tma_load_2d(&tmDy, &full_bar[stage], sa + j * a_tile + a * p.boxA_bytes, kj + a * p.ckA, pix0);
          }
          if (p.plain_x) {
#pragma unroll
            for (int x = 0; x < 8; ++x)
              if (x < nboxes) tma_load_2d(&tmX, &full_bar[stage], sb + x * p.boxB_bytes, box_c[x], pix0);
          } else {
            const int img = pix0 / PQ;
            const int rem = pix0 - img * PQ;
            const int pi = rem / p.Q;
            const int pj = rem - pi * p.Q;
            const int base_w = pj * p.trav + p.lower_w;
            const int base_h = pi * p.trav + p.lower_h;
#pragma unroll
            for (int x = 0; x < 8; ++x)
              if (x < nboxes)
                tma_load_im2col_4d(&tmX, &full_bar[stage], sb + x * p.boxB_bytes, box_c[x], base_w, base_h, img, box_w[x],
                                   box_h[x]);
          }
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint64_t protoA = make_smem_desc(0, p.boxA_bytes, 8 * p.ckA * 2, layout_type_for_row_bytes(p.ckA * 2));
        const uint64_t protoB = make_smem_desc(0, p.boxB_bytes, 8 * p.ckB * 2, layout_type_for_row_bytes(p.ckB * 2));
        const uint32_t kincA = (16u * p.ckA * 2) >> 4, kincB = (16u * p.ckB * 2) >> 4;  // 16 pixel rows per K step
        const int boxes_per_mma = min(8, 256 / p.ckB);
        const int ksteps = p.bk / 16;
        // the issuing thread is latency-bound per instruction: the (k-tile, box group) list of one stage -- descriptor
        // offsets, TMEM column, instruction descriptor -- is built once
        // slot (j, gi): k-tile j, box group gi (at most 2 groups of boxes_per_mma boxes); statically indexed -> registers
        uint32_t g_da[8], g_db[8], g_tm[8], g_id[8];
        bool g_on[8];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int gi = 0; gi < 2; ++gi) {
            const int g0 = gi * boxes_per_mma;
            const int nb = min(boxes_per_mma, nboxes - g0);
            g_on[j * 2 + gi] = j < kt_valid && g0 < nboxes;
            g_da[j * 2 + gi] = (j * a_tile) >> 4;
            g_db[j * 2 + gi] = (a_region + g0 * p.boxB_bytes) >> 4;
            g_tm[j * 2 + gi] = tmem_base + (j * p.boxes_per_cta + g0) * p.ckB;
            g_id[j * 2 + gi] = make_idesc_bf16(kTileM, max(nb, 1) * p.ckB, 1, 1);
          }
        for (int b = 0; b < nblk; ++b) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * p.stage_bytes);
          const uint64_t da_s = protoA + (a_addr >> 4), db_s = protoB + (a_addr >> 4);
          const uint32_t first = b != 0 ? 1u : 0u;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (!g_on[g]) continue;
            const uint64_t da0 = da_s + g_da[g], db0 = db_s + g_db[g];
            const uint32_t d_tmem = g_tm[g], idesc = g_id[g];
            if (ksteps == 4) {
              umma_bf16(d_tmem, da0, db0, idesc, first);
              umma_bf16(d_tmem, da0 + kincA, db0 + kincB, idesc, 1u);
              umma_bf16(d_tmem, da0 + 2 * kincA, db0 + 2 * kincB, idesc, 1u);
              umma_bf16(d_tmem, da0 + 3 * kincA, db0 + 3 * kincB, idesc, 1u);
            } else {
              for (int k = 0; k < ksteps; ++k)
                umma_bf16(d_tmem, da0 + k * kincA, db0 + k * kincB, idesc, (b | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (b == nblk - 1) umma_commit(&acc_bar);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    } else {
      const int q = warp & 3;
      mbar_wait(&acc_bar, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      for (int j = 0; j < kt_valid; ++j) {
        const int k = k0 + j * kTileM + q * 32 + lane;
        const bool row_ok = k < p.K_out;
        const int col_j = j * p.boxes_per_cta * p.ckB;
        if (p.partial != nullptr) {
          // split-K: plain stores of this CTA's fp32 tile; conv_wgrad_reduce_kernel sums the splits into dw
          float* dst = p.partial + ((static_cast<long long>(tile) * p.splits + split) * kTileM + (q * 32 + lane)) * p.pitch +
                       col_j;
          const int ncols = nboxes * p.ckB;
          for (int c0 = 0; c0 < ncols; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + col_j + c0, v);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(dst + c0 + 4 * i) =
                    make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                __uint_as_float(v[4 * i + 3]));
            }
          }
        } else {
          for (int x = 0; x < nboxes; ++x) {
            const int id = box0 + x;
            const int t = id / p.c_chunks;
            const int cc = id - t * p.c_chunks;
            const int tap = t;
            const int cbase = cc * p.ckB;
            float* dst = p.dw + (static_cast<long long>(k) * p.taps_total + tap) * p.C + cbase;
            for (int c0 = 0; c0 < p.ckB; c0 += 16) {
              uint32_t v[16];
              tmem_ld16(taddr + col_j + x * p.ckB + c0, v);
              tmem_ld_wait();
              if (row_ok) {
                if (cbase + c0 + 16 <= p.C && (p.C & 3) == 0) {
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    red_add_v4(dst + c0 + 4 * i, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                               __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                } else {
#pragma unroll
                  for (int i = 0; i < 16; ++i)
                    if (cbase + c0 + i < p.C) atomicAdd(dst + c0 + i, __uint_as_float(v[i]));
                }
              }
            }
          }
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// dw[k][tap][c] += sum over splits of the partial tiles (fixed order => deterministic).  A block owns 32 consecutive
// float4 outputs; its 8 warps each sum every 8th split (coalesced 512 B reads) and combine through shared memory --
// with one thread per output the loop over up to 148 splits was a serial chain of L2 round trips.
__global__ void __launch_bounds__(256) conv_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                                int K_out, int taps, int C, int ckB, int c_chunks,
                                                                int boxes_per_cta, int kt, int k_groups, int splits,
                                                                int pitch) {
  pdl_wait();
  __shared__ float4 red[8][32];
  const int c4n = C >> 2;
  const long long total = static_cast<long long>(K_out) * taps * c4n;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (long long base = static_cast<long long>(blockIdx.x) * 32; base < total; base += static_cast<long long>(gridDim.x) * 32) {
    const long long idx = base + lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int c = 0, tap = 0, k = 0;
    if (idx < total) {
      c = static_cast<int>(idx % c4n) * 4;
      tap = static_cast<int>((idx / c4n) % taps);
      k = static_cast<int>(idx / (static_cast<long long>(c4n) * taps));
      const int k_tile = k / kTileM, row = k - k_tile * kTileM;
      const int k_group = k_tile / kt, j = k_tile - k_group * kt;
      const int cc = c / ckB;
      const int id = tap * c_chunks + cc;
      const int cgroup = id / boxes_per_cta, x = id - cgroup * boxes_per_cta;
      const int tile = cgroup * k_groups + k_group;
      const float* src = partial + ((static_cast<long long>(tile) * splits) * kTileM + row) * pitch +
                         (j * boxes_per_cta + x) * ckB + (c - cc * ckB);
      for (int s2 = w; s2 < splits; s2 += nw) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(src + static_cast<long long>(s2) * kTileM * pitch));
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
      float4* o = reinterpret_cast<float4*>(dw + (static_cast<long long>(k) * taps + tap) * C + c);
      float4 cur = *o;
      cur.x += acc.x; cur.y += acc.y; cur.z += acc.z; cur.w += acc.w;
      *o = cur;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// host side
static int pick_ck(int channels) { return channels <= 16 ? 16 : (channels <= 32 ? 32 : 64); }

static int encode_im2col(CUtensorMap* tm, const void* base, int Nimg, int H, int W, int C, int ck, int pixels,
                         int lower_w, int lower_h, int upper_w, int upper_h, int trav, long long pix_stride = 0,
                         long long row_stride = 0, long long img_stride = 0) {
  EncodeIm2colFn fn = encode_im2col_fn();
  B200_REQUIRE(fn != nullptr, B200_ERR_CUDA, "cuTensorMapEncodeIm2col entry point unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nimg};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  if (pix_stride > 0) {  // explicit (possibly overlapping) layout, in elements
    strides[0] = (cuuint64_t)pix_stride * 2;
    strides[1] = (cuuint64_t)row_stride * 2;
    strides[2] = (cuuint64_t)img_stride * 2;
  }
  int lower[2] = {lower_w, lower_h};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)trav, (cuuint32_t)trav, 1};
  B200_REQUIRE(lower_w >= -128 && lower_w <= 127 && lower_h >= -128 && lower_h <= 127 && upper_w >= -128 &&
                   upper_w <= 127 && upper_h >= -128 && upper_h <= 127,
               B200_ERR_UNSUPPORTED, "im2col corner offsets out of the TMA range");
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower, upper,
                  (cuuint32_t)ck, (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_for_row_bytes(ck * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA,
               "cuTensorMapEncodeIm2col failed (%d) N=%d H=%d W=%d C=%d ck=%d pix=%d lower=(%d,%d) upper=(%d,%d) trav=%d",
               (int)r, Nimg, H, W, C, ck, pixels, lower_w, lower_h, upper_w, upper_h, trav);
  // Driver quirk for small tensors (as worked around by CUTLASS' im2col descriptor builder):
  // for tensors below 128 KiB, bit 21 of the second descriptor word must be cleared on drivers <= 13.1.
  int drv = 0;
  cudaDriverGetVersion(&drv);
  if (drv <= 13010 && (size_t)Nimg * H * W * C * 2 < 131072) {
    reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
  }
  return B200_OK;
}

static int encode_tiled3(CUtensorMap* tm, const void* base, int d0, int d1, int d2, int b0, int b1, int b2) {
  EncodeTiledFn fn = encode_tiled_fn();
  B200_REQUIRE(fn != nullptr, B200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)d0 * 2, (cuuint64_t)d0 * d1 * 2};
  cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(b0 * 2),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed (%d) dims=(%d,%d,%d) box=(%d,%d,%d)",
               (int)r, d0, d1, d2, b0, b1, b2);
  return B200_OK;
}

static int encode_tiled2(CUtensorMap* tm, const void* base, int d0, long long d1, int b0, int b1) {
  EncodeTiledFn fn = encode_tiled_fn();
  B200_REQUIRE(fn != nullptr, B200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
  cuuint64_t strides[1] = {(cuuint64_t)d0 * 2};
  cuuint32_t box[2] = {(cuuint32_t)b0, (cuuint32_t)b1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(b0 * 2),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled(2d) failed (%d) dims=(%d,%lld) box=(%d,%d)",
               (int)r, d0, d1, b0, b1);
  return B200_OK;
}

static const int kSmemBudget = 200 * 1024;
static const int kSmemBudgetMax = 224 * 1024;   // + 1 KB alignment slack + static barriers < the 227 KB per-CTA limit

static int set_smem_attr(const void* fn, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  B200_REQUIRE(e == cudaSuccess, B200_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", bytes, cudaGetErrorString(e));
  return B200_OK;
}

// One implicit-GEMM launch.  src: [Nimg, SH, SW, SC] bf16 (im2col source); wmat: [Nout][wtaps][SC] bf16.
struct IgemmLaunch {
  const void* src; int Nimg, SH, SW, SC;
  long long s_pix, s_row, s_img;   // optional explicit source strides (elements), 0 = dense
  const void* wmat; int Nout, wtaps;
  int I, J, trav, lower_w, lower_h;
  int ntaps; TapEntry taps[kMaxTaps];
  void* out; int OH, OW, os, oh0, ow0, ldo;
  const void* res; const float* bias; int act, out_fp32;
  double* stats;
  int window;
};

static int launch_igemm(const IgemmLaunch& L, cudaStream_t stream) {
  B200_REQUIRE(L.SC % 8 == 0, B200_ERR_UNSUPPORTED, "igemm: source channels (%d) must be a multiple of 8", L.SC);
  B200_REQUIRE(L.ntaps >= 1 && L.ntaps <= kMaxTaps, B200_ERR_UNSUPPORTED, "igemm: %d taps unsupported", L.ntaps);
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.M_total = L.Nimg * L.I * L.J;
  B200_REQUIRE(p.M_total > 0, B200_ERR_INVALID, "igemm: empty problem");
  p.I = L.I; p.J = L.J; p.trav = L.trav; p.lower_w = L.lower_w; p.lower_h = L.lower_h;
  p.N_total = L.Nout;
  p.n_tiles = (L.Nout + 255) / 256;
  p.block_n = (((L.Nout + p.n_tiles - 1) / p.n_tiles) + 15) / 16 * 16;
  p.m_tiles = (p.M_total + kTileM - 1) / kTileM;
  p.ck = pick_ck(L.SC);
  p.c_chunks = (L.SC + p.ck - 1) / p.ck;
  p.window = L.window;
  if (L.window) {   // block-diagonal: one n-tile per window, its K loop covers the window's channels only
    B200_REQUIRE(L.window % 64 == 0 && L.window <= 256 && L.SC == L.Nout && L.Nout % L.window == 0, B200_ERR_UNSUPPORTED,
                 "igemm: window %d needs C == K (C=%d K=%d), K %% window == 0", L.window, L.SC, L.Nout);
    p.n_tiles = L.Nout / L.window;
    p.block_n = L.window;
    p.ck = 64;
    p.c_chunks = L.window / 64;
  }
  p.ntaps = L.ntaps;
  p.a_bytes = kTileM * p.ck * 2;
  p.b_bytes = p.block_n * p.ck * 2;
  p.tx_bytes = p.a_bytes + p.b_bytes;
  // keep every stage 1024B aligned (128B-swizzle atoms are 1024B)
  uint32_t stage = p.a_bytes + p.b_bytes;
  if (stage % 1024) { p.b_bytes += 1024 - stage % 1024; stage = p.a_bytes + p.b_bytes; }
  // dense bf16 outputs with 64-channel granularity go out through shared memory + TMA store
  p.tma_store = (L.os == 1 && !L.out_fp32 && (p.block_n % 64) == 0 && (L.Nout % p.block_n) == 0 && L.ldo == L.Nout &&
                 L.OH == L.I && L.OW == L.J && L.oh0 == 0 && L.ow0 == 0) ? 1 : 0;
  const int epi_bytes = p.tma_store ? kTileM * p.block_n * 2 : 0;
  // small 1x1 layers (all weight slices <= 64 KB): keep the weights resident, stream only the activations
  static const bool bstat_enabled = !(getenv("B200_IGEMM_BSTAT") && atoi(getenv("B200_IGEMM_BSTAT")) == 0);
  const int b_all = L.ntaps * p.c_chunks * (int)p.b_bytes;
  int bstat_bytes = 0;
  if (bstat_enabled && p.n_tiles == 1 && !L.window && (p.a_bytes % 1024) == 0 && (p.b_bytes % 1024) == 0 && b_all <= 64 * 1024) {
    p.b_stationary = 1;
    bstat_bytes = b_all;
    stage = p.a_bytes;
  }
  // several n-tiles whose weights fit one at a time: a CTA owns an n-tile.  Worth it when the weights dominate the
  // L2->SM traffic of the round-robin walk (bytes ~ A * n_tiles + W * m_tiles vs A * n_tiles + W_tile * CTAs) and
  // every CTA still gets a few m-tiles; needs >= 3 operand stages beside the resident weights and the staging tile
  // (with two, l3 256->1024 did not gain although its L2->SM traffic fell 2.5x: pipeline depth matters as much).
  static const int own_mode = getenv("B200_IGEMM_OWN_NTILE") ? atoi(getenv("B200_IGEMM_OWN_NTILE")) : 1;
  int grid_own = 0;
  // B200_IGEMM_SMEM_KB: operand + staging budget (default 224 KB: three 48 KB stages for 256-wide tiles beside the 64 KB
  // staging tile; with 200 KB they ran a TWO-stage pipeline -- l3/l4 1x1 layers 15-20 % slower, 0.35 ms/step)
  static const int budget_kb = getenv("B200_IGEMM_SMEM_KB") ? atoi(getenv("B200_IGEMM_SMEM_KB")) : 224;
  int budget = (budget_kb >= 96 && budget_kb <= 224 ? budget_kb : 224) * 1024;
  if (bstat_enabled && own_mode && !p.b_stationary && p.n_tiles > 1 && p.n_tiles <= 8 && !L.window &&
      (p.a_bytes % 1024) == 0 && (p.b_bytes % 1024) == 0 && (L.Nout % p.block_n) == 0) {
    const int ctas = sm_count() / p.n_tiles;
    const long long a_all = (long long)p.M_total * L.SC * 2 * L.ntaps;
    const long long w_all = (long long)b_all * p.n_tiles;
    const long long rr_bytes = a_all * p.n_tiles + w_all * p.m_tiles;
    const long long own_bytes = a_all * p.n_tiles + (long long)b_all * ctas * p.n_tiles;
    const int stages_left = (kSmemBudgetMax - epi_bytes - b_all) / (int)p.a_bytes;
    if (ctas >= 1 && p.m_tiles >= 4 * ctas && stages_left >= 3 && own_bytes * 4 <= rr_bytes * 3) {
      p.b_stationary = 1;
      p.own_ntile = 1;
      bstat_bytes = b_all;
      stage = p.a_bytes;
      grid_own = ctas * p.n_tiles;
      budget = kSmemBudgetMax;
    }
  }
  // second staging tile (residual prefetched one tile ahead, store/convert overlap) when at least three operand
  // stages remain.  B200_IGEMM_EPI2: 0 = never, 2 = every epilogue, default 1 = epilogues with a residual only --
  // measured (profiles/r02_summary.md): residual epilogues gain up to 28 %, plain ones lose 0-8 % to the lost stage
  static const int epi2_mode = getenv("B200_IGEMM_EPI2") ? atoi(getenv("B200_IGEMM_EPI2")) : 1;
  p.epi_bufs = 1;
  p.epi_bytes = (uint32_t)epi_bytes;
  if ((epi2_mode >= 2 || (epi2_mode == 1 && L.res != nullptr)) && p.tma_store &&
      (budget - 2 * epi_bytes - bstat_bytes) / (int)stage >= 3)
    p.epi_bufs = 2;
  const int epi_total = epi_bytes * p.epi_bufs;
  p.num_stages = (budget - epi_total - bstat_bytes) / (int)stage;
  if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
  if (p.num_stages < 2) p.num_stages = 2;
  p.OH = L.OH; p.OW = L.OW; p.os = L.os; p.oh0 = L.oh0; p.ow0 = L.ow0; p.ldo = L.ldo;
  p.act = L.act; p.out_fp32 = L.out_fp32; p.out = L.out; p.res = L.res; p.bias = L.bias;
  p.stats = L.stats;
  if (L.stats != nullptr)
    B200_REQUIRE(p.tma_store && (256 % p.block_n) == 0 && L.res == nullptr && L.bias == nullptr && L.act == 0,
                 B200_ERR_UNSUPPORTED, "conv_fprop: fused BN statistics need a dense bf16 output with K %% 64 == 0 "
                 "(block_n=%d) and no bias/residual/activation", p.block_n);
  for (int t = 0; t < L.ntaps; ++t) p.taps[t] = L.taps[t];

  // bounding box of base pixels: [lower, lower + (I-1)*trav] in a source of extent SH x SW
  const int upper_w = L.lower_w + (L.J - 1) * L.trav + 1 - L.SW;
  const int upper_h = L.lower_h + (L.I - 1) * L.trav + 1 - L.SH;
  CUtensorMap tmA, tmB;
  int rc;
  p.plain_a = (L.ntaps == 1 && L.trav == 1 && L.lower_w == 0 && L.lower_h == 0 && L.taps[0].off_w == 0 &&
               L.taps[0].off_h == 0 && L.I == L.SH && L.J == L.SW && L.s_pix == 0) ? 1 : 0;
  if (p.plain_a)
    rc = encode_tiled2(&tmA, L.src, L.SC, (long long)p.M_total, p.ck, kTileM);
  else
    rc = encode_im2col(&tmA, L.src, L.Nimg, L.SH, L.SW, L.SC, p.ck, kTileM, L.lower_w, L.lower_h, upper_w, upper_h,
                       L.trav, L.s_pix, L.s_row, L.s_img);
  if (rc) return rc;
  rc = encode_tiled3(&tmB, L.wmat, L.window ? L.window : L.SC, L.wtaps, L.Nout, p.ck, 1, p.block_n);
  if (rc) return rc;

  CUtensorMap tmC, tmR;
  memset(&tmC, 0, sizeof(tmC));
  memset(&tmR, 0, sizeof(tmR));
  if (p.tma_store) {
    rc = encode_tiled2(&tmC, L.out, L.ldo, (long long)p.M_total, 64, kTileM);
    if (rc) return rc;
    if (L.res != nullptr) {
      rc = encode_tiled2(&tmR, L.res, L.ldo, (long long)p.M_total, 64, kTileM);
      if (rc) return rc;
    }
  }
  const int smem_bytes = p.num_stages * (int)stage + bstat_bytes + epi_total + 1024;
  rc = set_smem_attr((const void*)conv_igemm_kernel, smem_bytes);
  if (rc) return rc;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int grid = grid_own ? grid_own : (total_tiles < sm_count() ? total_tiles : sm_count());
  if (getenv("B200_IGEMM_DEBUG"))
    fprintf(stderr, "[igemm] M=%d C=%d N=%d taps=%d block_n=%d n_tiles=%d stages=%d bstat=%d own=%d epi_bufs=%d grid=%d smem=%d\n",
            p.M_total, L.SC, L.Nout, L.ntaps, p.block_n, p.n_tiles, p.num_stages, p.b_stationary, p.own_ntile, p.epi_bufs,
            grid, smem_bytes);
  b200::launch(conv_igemm_kernel, grid, kIgemmThreads, smem_bytes, stream, tmA, tmB, tmC, tmR, p);
  B200_CHECK_LAUNCH("conv_igemm_kernel");
  return B200_OK;
}

static int check_desc(const b200_conv_desc* d) {
  B200_REQUIRE(d != nullptr, B200_ERR_INVALID, "conv: null descriptor");
  B200_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->C > 0 && d->K > 0 && d->R > 0 && d->S > 0 && d->P > 0 &&
                   d->Q > 0 && d->stride >= 1 && d->stride <= 8,
               B200_ERR_INVALID, "conv: bad descriptor N=%d H=%d W=%d C=%d K=%d R=%d S=%d P=%d Q=%d stride=%d", d->N,
               d->H, d->W, d->C, d->K, d->R, d->S, d->P, d->Q, d->stride);
  B200_REQUIRE(d->R * d->S <= kMaxTaps, B200_ERR_UNSUPPORTED, "conv: %dx%d filter exceeds %d taps", d->R, d->S,
               kMaxTaps);
  B200_REQUIRE((d->x_pixel_stride == 0 && d->x_row_stride == 0 && d->x_image_stride == 0) ||
                   (d->x_pixel_stride > 0 && d->x_pixel_stride % 8 == 0 && d->x_row_stride % 8 == 0 &&
                    d->x_image_stride % 8 == 0 && d->x_row_stride > 0 && d->x_image_stride > 0),
               B200_ERR_INVALID, "conv: x strides must all be 0 (dense) or positive multiples of 8 elements");
  B200_REQUIRE(d->window == 0 || (d->window % 64 == 0 && d->C == d->K && d->C % d->window == 0 && d->x_pixel_stride == 0),
               B200_ERR_UNSUPPORTED, "conv: window %d needs C == K, C %% window == 0 (C=%d K=%d)", d->window, d->C, d->K);
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_conv_fprop(const b200_conv_desc* d, const void* x, const void* w, void* y,
                               const b200_epilogue* ep, b200_stream_t stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  B200_REQUIRE(x && w && y, B200_ERR_INVALID, "conv_fprop: null pointer");
  B200_REQUIRE(d->C % 8 == 0, B200_ERR_UNSUPPORTED, "conv_fprop: C=%d must be a multiple of 8 (pad the input)", d->C);
  if (d->stride == 1 && d->pad_h == d->pad_w && d->P == d->H + 2 * d->pad_h - d->R + 1 &&
      d->Q == d->W + 2 * d->pad_w - d->S + 1 && d->x_pixel_stride == 0 && (!ep || !ep->out_fp32) &&
      !(ep && ep->bias && ep->bn_stats_workspace) && (d->window == 0 || d->window == 64) &&
      halo_eligible(d->P, d->Q, d->C, d->K, d->R, d->S, d->pad_h)) {
    return launch_halo(x, w, y, ep ? ep->residual : nullptr, ep ? ep->bias : nullptr, d->N, d->P, d->Q, d->C, d->K,
                       d->R, d->S, d->pad_h, 0, ep ? ep->act : 0,
                       (ep && ep->bn_stats_workspace) ? reinterpret_cast<double*>(ep->bn_stats_workspace) : nullptr,
                       (cudaStream_t)stream, d->window);
  }
  if (d->R == 1 && d->S == 1 && d->stride == 1 && d->pad_h == 0 && d->pad_w == 0 && d->x_pixel_stride == 0 && !d->window &&
      (!ep || !ep->out_fp32) && !(ep && ep->bias && ep->bn_stats_workspace) &&
      pair_eligible((long long)d->N * d->H * d->W, d->C, d->K)) {   // opt-in CTA-pair kernel (conv_pair.cu)
    return launch_pair(x, w, y, ep ? ep->residual : nullptr, ep ? ep->bias : nullptr, (long long)d->N * d->H * d->W, d->C,
                       d->K, ep ? ep->act : 0,
                       (ep && ep->bn_stats_workspace) ? reinterpret_cast<double*>(ep->bn_stats_workspace) : nullptr,
                       (cudaStream_t)stream);
  }
  IgemmLaunch L;
  memset(&L, 0, sizeof(L));
  L.src = x; L.Nimg = d->N; L.SH = d->H; L.SW = d->W; L.SC = d->C;
  L.s_pix = d->x_pixel_stride; L.s_row = d->x_row_stride; L.s_img = d->x_image_stride;
  L.wmat = w; L.Nout = d->K; L.wtaps = d->R * d->S;
  L.I = d->P; L.J = d->Q; L.trav = d->stride; L.lower_w = -d->pad_w; L.lower_h = -d->pad_h;
  L.ntaps = d->R * d->S;
  for (int r = 0; r < d->R; ++r)
    for (int s = 0; s < d->S; ++s) {
      TapEntry& t = L.taps[r * d->S + s];
      t.off_w = (uint16_t)s; t.off_h = (uint16_t)r; t.b_tap = (uint16_t)(r * d->S + s); t.pad_ = 0;
    }
  L.out = y; L.OH = d->P; L.OW = d->Q; L.os = 1; L.oh0 = 0; L.ow0 = 0; L.ldo = d->K;
  L.res = ep ? ep->residual : nullptr;
  L.bias = ep ? ep->bias : nullptr;
  L.act = ep ? ep->act : 0;
  L.out_fp32 = ep ? ep->out_fp32 : 0;
  L.stats = (ep && ep->bn_stats_workspace) ? reinterpret_cast<double*>(ep->bn_stats_workspace) : nullptr;
  L.window = d->window;
  return launch_igemm(L, (cudaStream_t)stream);
}

extern "C" int b200_conv_dgrad(const b200_conv_desc* d, const void* dy, const void* wt, void* dx,
                               const void* residual, b200_stream_t stream_) {
  int rc = check_desc(d);
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_REQUIRE(dy && wt && dx, B200_ERR_INVALID, "conv_dgrad: null pointer");
  B200_REQUIRE(d->K % 8 == 0, B200_ERR_UNSUPPORTED, "conv_dgrad: K=%d must be a multiple of 8", d->K);
  const int st = d->stride;
  B200_REQUIRE(st == 1 || st == 2, B200_ERR_UNSUPPORTED, "conv_dgrad: stride %d unsupported", st);
  if (d->R == 3 && d->S == 3 && st == 1 && d->pad_h == 1 && d->pad_w == 1 && d->P == d->H && d->Q == d->W &&
      (d->window == 0 || d->window == 64) && halo_eligible(d->H, d->W, d->K, d->C, 3, 3, 1)) {
    return launch_halo(dy, wt, dx, residual, nullptr, d->N, d->H, d->W, d->K, d->C, 3, 3, 1, 1, 0, nullptr, stream,
                       d->window);
  }
  if (d->R == 1 && d->S == 1 && st == 1 && d->pad_h == 0 && d->pad_w == 0 && !d->window &&
      pair_eligible((long long)d->N * d->H * d->W, d->K, d->C)) {   // opt-in CTA-pair kernel: dx = dy * wt^T
    return launch_pair(dy, wt, dx, residual, nullptr, (long long)d->N * d->H * d->W, d->K, d->C, 0, nullptr, stream);
  }
  // dx[h,w] = sum_{r,s : (h+pad-r) % st == 0} dy[(h+pad-r)/st, (w+pad-s)/st] * w[r,s]
  // one launch per residue class (h % st, w % st); each class is a stride-1 correlation over dy.
  bool any_empty = false;
  for (int ph = 0; ph < st; ++ph)
    for (int pw = 0; pw < st; ++pw) {
      int nr = 0, ns = 0;
      for (int r = 0; r < d->R; ++r) if (((ph + d->pad_h - r) % st + st) % st == 0) ++nr;
      for (int s = 0; s < d->S; ++s) if (((pw + d->pad_w - s) % st + st) % st == 0) ++ns;
      if (nr * ns == 0 && (d->H - ph + st - 1) / st > 0 && (d->W - pw + st - 1) / st > 0) any_empty = true;
    }
  if (any_empty) {
    B200_REQUIRE(residual == nullptr, B200_ERR_UNSUPPORTED,
                 "conv_dgrad: residual add with an empty stride class is unsupported");
    cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)d->N * d->H * d->W * d->C * 2, stream);
    B200_REQUIRE(e == cudaSuccess, B200_ERR_CUDA, "conv_dgrad: memset failed: %s", cudaGetErrorString(e));
    count_launch();
  }
  for (int ph = 0; ph < st; ++ph)
    for (int pw = 0; pw < st; ++pw) {
      const int I = (d->H - ph + st - 1) / st;
      const int J = (d->W - pw + st - 1) / st;
      if (I <= 0 || J <= 0) continue;
      IgemmLaunch L;
      memset(&L, 0, sizeof(L));
      int dh[kMaxTaps], dwv[kMaxTaps], rr[kMaxTaps], ss[kMaxTaps];
      int nr = 0, ns = 0;
      for (int r = 0; r < d->R; ++r) {
        const int v = ph + d->pad_h - r;
        if (((v % st) + st) % st == 0) { dh[nr] = (v - (((v % st) + st) % st)) / st; rr[nr] = r; ++nr; }
      }
      for (int s = 0; s < d->S; ++s) {
        const int v = pw + d->pad_w - s;
        if (((v % st) + st) % st == 0) { dwv[ns] = (v - (((v % st) + st) % st)) / st; ss[ns] = s; ++ns; }
      }
      if (nr * ns == 0) continue;
      int lo_h = dh[0], lo_w = dwv[0];
      for (int i = 1; i < nr; ++i) lo_h = dh[i] < lo_h ? dh[i] : lo_h;
      for (int i = 1; i < ns; ++i) lo_w = dwv[i] < lo_w ? dwv[i] : lo_w;
      L.ntaps = nr * ns;
      for (int a = 0; a < nr; ++a)
        for (int b = 0; b < ns; ++b) {
          TapEntry& t = L.taps[a * ns + b];
          t.off_h = (uint16_t)(dh[a] - lo_h);
          t.off_w = (uint16_t)(dwv[b] - lo_w);
          t.b_tap = (uint16_t)(rr[a] * d->S + ss[b]);
          t.pad_ = 0;
        }
      L.src = dy; L.Nimg = d->N; L.SH = d->P; L.SW = d->Q; L.SC = d->K;
      L.wmat = wt; L.Nout = d->C; L.wtaps = d->R * d->S;
      L.I = I; L.J = J; L.trav = 1; L.lower_w = lo_w; L.lower_h = lo_h;
      L.out = dx; L.OH = d->H; L.OW = d->W; L.os = st; L.oh0 = ph; L.ow0 = pw; L.ldo = d->C;
      L.res = residual; L.bias = nullptr; L.act = 0; L.out_fp32 = 0;
      L.window = d->window;
      rc = launch_igemm(L, stream);
      if (rc) return rc;
    }
  return B200_OK;
}

extern "C" size_t b200_conv_wgrad_workspace_bytes(void) {
  // splits * tiles <= SM count, each partial tile is 128 x 512 fp32
  return static_cast<size_t>(sm_count() + 8) * kTileM * 512 * sizeof(float);
}

extern "C" int b200_conv_wgrad(const b200_conv_desc* d, const void* x, const void* dy, float* dw, void* workspace,
                               size_t workspace_bytes, b200_stream_t stream_) {
  int rc = check_desc(d);
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_REQUIRE(x && dy && dw, B200_ERR_INVALID, "conv_wgrad: null pointer");
  B200_REQUIRE(d->C % 8 == 0 && d->K % 8 == 0, B200_ERR_UNSUPPORTED,
               "conv_wgrad: C=%d and K=%d must be multiples of 8", d->C, d->K);
  if (d->stride == 1 && d->pad_h == d->pad_w && d->P == d->H + 2 * d->pad_h - d->R + 1 &&
      d->Q == d->W + 2 * d->pad_w - d->S + 1 && d->x_pixel_stride == 0 &&
      (d->window == 0 || d->window == 128) && halo_wgrad_eligible(d->P, d->Q, d->C, d->K, d->R, d->S, d->pad_h)) {
    // partial tiles of the halo kernel must fit the split-K workspace (units * splits <= SMs, or one split)
    const int cw = d->C == 16 ? 16 : 32;
    const int units = ((d->window ? d->window : d->C) / cw) * ((d->K + kTileM - 1) / kTileM);
    if (units <= sm_count() + 8)
      return launch_halo_wgrad(x, dy, dw, workspace, workspace_bytes, d->N, d->P, d->Q, d->C, d->K, d->R, d->S, d->pad_h,
                               stream, d->window);
  }
  B200_REQUIRE(d->window == 0 || d->window == 128, B200_ERR_UNSUPPORTED, "conv_wgrad: window must be 0 or 128");
  // window mode (block-diagonal): k-tile t pairs with input channels [128t, 128t+128) only; dw is [K][taps][128]
  const int Cw = d->window ? d->window : d->C;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.M_total = d->N * d->P * d->Q;
  p.P = d->P; p.Q = d->Q; p.trav = d->stride; p.lower_w = -d->pad_w; p.lower_h = -d->pad_h;
  p.K_out = d->K; p.C = Cw; p.taps_total = d->R * d->S;
  p.window = d->window;
  p.ckA = pick_ck(d->K);
  p.ckB = pick_ck(Cw);
  p.bk = 64;  // pixels per TMA box: fewer, larger TMA requests per byte (32-pixel boxes were request-rate bound)
  p.c_chunks = (Cw + p.ckB - 1) / p.ckB;
  p.total_boxes = p.taps_total * p.c_chunks;
  p.k_tiles = (d->K + kTileM - 1) / kTileM;
  p.boxA_bytes = p.bk * p.ckA * 2;
  p.boxB_bytes = p.bk * p.ckB * 2;
  // Tile shape: kt k-tiles x bpc channel boxes per CTA (kt * bpc * ckB <= 512 TMEM columns).  The kernels are bound by
  // operand delivery from L2, so pick the shape that fetches the fewest bytes per MMA flop:
  //   bytes per stage = kt * (dy tile) + bpc * (x box),  flops per stage ~ kt * bpc.
  static const int kt_cap = getenv("B200_WGRAD_KT") ? atoi(getenv("B200_WGRAD_KT")) : 4;
  double best = 1e30;
  p.kt = 1; p.boxes_per_cta = 1;
  for (int kt = 1; kt <= 4 && kt <= p.k_tiles && kt <= (d->window ? 1 : kt_cap); kt *= 2) {
    int bpc = 512 / (kt * p.ckB);
    if (bpc > 8) bpc = 8;
    if (bpc > p.total_boxes) bpc = p.total_boxes;
    if (bpc < 1) continue;
    const double a_bytes = (double)kt * (kTileM / p.ckA) * p.boxA_bytes, b_bytes = (double)bpc * p.boxB_bytes;
    if (a_bytes + b_bytes > 96.0 * 1024) continue;                       // at least two stages in shared memory
    const double cost = (a_bytes + b_bytes) / ((double)kt * bpc * p.ckB);
    if (cost < best * 0.999) { best = cost; p.kt = kt; p.boxes_per_cta = bpc; }
  }
  p.k_groups = (p.k_tiles + p.kt - 1) / p.kt;
  p.col_groups = (p.total_boxes + p.boxes_per_cta - 1) / p.boxes_per_cta;
  p.total_blocks = (p.M_total + p.bk - 1) / p.bk;
  const int tiles = p.k_groups * p.col_groups;
  int splits = sm_count() / tiles;  // one wave: a CTA owns the whole TMEM, two cannot share an SM
  if (splits > p.total_blocks) splits = p.total_blocks;
  if (splits < 1) splits = 1;
  p.blocks_per_split = (p.total_blocks + splits - 1) / splits;
  p.splits = (p.total_blocks + p.blocks_per_split - 1) / p.blocks_per_split;
  p.stage_bytes = p.kt * (kTileM / p.ckA) * p.boxA_bytes + p.boxes_per_cta * p.boxB_bytes;
  p.stage_bytes = (p.stage_bytes + 1023u) & ~1023u;
  // 200 KB, not the full 224: the weight gradients run on the side stream BESIDE the main stream's BatchNorm kernels,
  // whose reduce blocks need ~10 KB of shared memory on the same SM
  p.num_stages = kSmemBudget / (int)p.stage_bytes;
  if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
  if (p.num_stages < 2) p.num_stages = 2;
  p.dw = dw;
  p.pitch = p.kt * p.boxes_per_cta * p.ckB;
  p.partial = nullptr;
  if (p.splits > 1) {
    const size_t need = static_cast<size_t>(tiles) * p.splits * kTileM * p.pitch * sizeof(float);
    B200_REQUIRE(workspace != nullptr && workspace_bytes >= need, B200_ERR_INVALID,
                 "conv_wgrad: workspace too small (%zu < %zu); see b200_conv_wgrad_workspace_bytes()", workspace_bytes,
                 need);
    B200_REQUIRE((Cw & 3) == 0, B200_ERR_UNSUPPORTED, "conv_wgrad: C must be a multiple of 4");
    p.partial = static_cast<float*>(workspace);
  }
  p.S_filter = d->S;
  if (getenv("B200_WGRAD_DEBUG"))
    fprintf(stderr, "wgrad cfg: K=%d C=%d taps=%d ckA=%d ckB=%d kt=%d bpc=%d k_groups=%d col_groups=%d splits=%d bps=%d "
            "stages=%d stage=%u pitch=%d partial=%d\n", d->K, d->C, p.taps_total, p.ckA, p.ckB, p.kt, p.boxes_per_cta,
            p.k_groups, p.col_groups, p.splits, p.blocks_per_split, p.num_stages, p.stage_bytes, p.pitch,
            p.partial != nullptr);
  CUtensorMap tmDy, tmX;
  rc = encode_tiled2(&tmDy, dy, d->K, (long long)p.M_total, p.ckA, p.bk);
  if (rc) return rc;
  const int upper_w = p.lower_w + (d->Q - 1) * d->stride + 1 - d->W;
  const int upper_h = p.lower_h + (d->P - 1) * d->stride + 1 - d->H;
  p.plain_x = (d->R == 1 && d->S == 1 && d->stride == 1 && d->pad_h == 0 && d->pad_w == 0 && d->P == d->H &&
               d->Q == d->W && d->x_pixel_stride == 0) ? 1 : 0;
  if (p.plain_x)
    rc = encode_tiled2(&tmX, x, d->C, (long long)p.M_total, p.ckB, p.bk);
  else
    rc = encode_im2col(&tmX, x, d->N, d->H, d->W, d->C, p.ckB, p.bk, p.lower_w, p.lower_h, upper_w, upper_h,
                       d->stride, d->x_pixel_stride, d->x_row_stride, d->x_image_stride);
  if (rc) return rc;
  const int smem_bytes = p.num_stages * (int)p.stage_bytes + 1024;
  rc = set_smem_attr((const void*)conv_wgrad_kernel, smem_bytes);
  if (rc) return rc;
  const int grid = tiles * p.splits;
  b200::launch(conv_wgrad_kernel, grid, kThreads, smem_bytes, stream, tmDy, tmX, p);
  B200_CHECK_LAUNCH("conv_wgrad_kernel");
  if (p.partial != nullptr) {
    const long long total = static_cast<long long>(d->K) * p.taps_total * (Cw / 4);
    long long blocks = (total + 31) / 32;
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    b200::launch(conv_wgrad_reduce_kernel, static_cast<int>(blocks), 32 * wgrad_reduce_warps(p.splits), 0, stream, p.partial, dw, d->K, p.taps_total, Cw, p.ckB,
                                                                          p.c_chunks, p.boxes_per_cta, p.kt, p.k_groups,
                                                                          p.splits, p.pitch);
    B200_CHECK_LAUNCH("conv_wgrad_reduce_kernel");
  }
  return B200_OK;
}
