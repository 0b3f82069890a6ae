// EXPERIMENTAL (off unless B200_IGEMM_PAIR=1; written at the end of round 1 after the GPU budget was spent, so it has
// been compiled and reviewed but NOT yet run -- validate with `tools/run_gpu_diag.sh "p1_*"` before enabling).
//
// 1x1 / stride-1 convolution (fprop, and dgrad through the transposed weights) as a GEMM on CTA PAIRS:
//   D[M, N] = A[M, C] * B[N, C]^T,  A = activations (dense NHWC rows), B = weights [N][C].
// The single-CTA igemm kernel is bound by L2->SM delivery on the wide layers (layer3/4: every 128-row tile re-streams
// a 256-column weight block: 16 KB of A + 32 KB of B per 64-channel step).  Here a cluster of two CTAs owns a 256 x N
// tile: CTA r loads rows [128r, 128r+128) of A and rows [r*N/2, (r+1)*N/2) of B, the leader issues
// tcgen05.mma.cta_group::2 (M = 256) which reads both halves of B from the two shared memories, and each CTA drains
// its own 128 TMEM lanes.  Per-SM traffic per step: 16 KB + 16 KB instead of 16 KB + 32 KB.
// Protocol (validated piecewise by tools/pair_probe.cu):
//   full[s]   lives in the LEADER: leader arrive.expect_tx(bytes of both CTAs), peer remote-arrives; the TMA loads of
//             both CTAs (cp.async.bulk.tensor ... cta_group::2) complete_tx on the leader's barrier.
//   empty[s], tmem_full[a]  local to each CTA, signalled by the leader's multicast tcgen05.commit (mask 0b11).
//   tmem_empty[a]  in the leader, 16 arrivals (8 epilogue warps of each CTA; the peer's are remote arrives).
#include "common.cuh"
#include "host.h"
#include <stdlib.h>

namespace b200 {

constexpr int kPThreads = 320;      // TMA warp, MMA warp, 8 epilogue warps (per CTA)
constexpr int kPMaxStages = 6;
constexpr int kPStatReplicas = 16;

struct PairParams {
  int M_total, N_total, C;
  int block_n, n_tiles, m_pairs, c_chunks, num_stages;
  uint32_t a_bytes, bh_bytes;       // 128 x 64 ch of A, block_n/2 x 64 ch of B
  int act, has_res;
  const float* bias;
  double* stats;
};

__device__ __forceinline__ uint32_t pair_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t pair_cluster_id() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t pair_nclusters() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void pair_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared-memory object of THIS CTA) in the leader CTA (rank 0)
__device__ __forceinline__ uint32_t leader_addr(const void* p) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(p)), "r"(0));
  return ra;
}
__device__ __forceinline__ void remote_arrive(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const void* tmap, uint32_t cluster_bar_addr, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ PairParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kPMaxStages], empty_bar[kPMaxStages];
  __shared__ __align__(8) uint64_t tmem_full[2], tmem_empty[2], res_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = pair_ctarank();
  const bool is_leader = rank == 0;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t stage_bytes = p.a_bytes + p.bh_bytes;
  uint8_t* epi = smem + p.num_stages * stage_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 2);          // used in the leader only: one arrival per CTA of the pair
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full[0], 1); mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], 16); mbar_init(&tmem_empty[1], 16);   // leader only: 8 epilogue warps x 2 CTAs
    mbar_init(&res_bar, 1);
    fence_mbar_init();
    prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmC);
    if (p.has_res) prefetch_tmap(&tmR);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  pair_cluster_sync();                     // barriers of BOTH CTAs initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int total_tiles = p.m_pairs * p.n_tiles;
  const int cid = static_cast<int>(pair_cluster_id()), ncl = static_cast<int>(pair_nclusters());

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = cid; tile < total_tiles; tile += ncl) {
        const int m_pair = tile / p.n_tiles, n_tile = tile - m_pair * p.n_tiles;
        const int row0 = m_pair * 256 + static_cast<int>(rank) * 128;
        const int col0 = n_tile * p.block_n + static_cast<int>(rank) * (p.block_n >> 1);
        for (int cc = 0; cc < p.c_chunks; ++cc) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);            // local: the leader's multicast commit frees both
          const uint32_t lbar = leader_addr(&full_bar[stage]);
          if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2u * stage_bytes);
          else remote_arrive(lbar);
          uint8_t* sa = smem + stage * stage_bytes;
          tma_load_2d_pair(&tmA, lbar, sa, cc * 64, row0);
          tma_load_2d_pair(&tmB, lbar, sa + p.a_bytes, cc * 64, col0);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (is_leader && lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t idesc = make_idesc_bf16(256, p.block_n, 0, 0);
      const uint64_t proto = make_smem_desc(0, 16, 1024, 2);
      int local = 0;
      for (int tile = cid; tile < total_tiles; tile += ncl, ++local) {
        const int acc = local & 1;
        mbar_wait(&tmem_empty[acc], ((local >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int cc = 0; cc < p.c_chunks; ++cc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint64_t da = proto + (a_addr >> 4);
          const uint64_t db = proto + ((a_addr + p.a_bytes) >> 4);
          umma_pair(d_tmem, da, db, idesc, cc != 0 ? 1u : 0u);
          umma_pair(d_tmem, da + 2, db + 2, idesc, 1u);
          umma_pair(d_tmem, da + 4, db + 4, idesc, 1u);
          umma_pair(d_tmem, da + 6, db + 6, idesc, 1u);
          umma_commit_pair(&empty_bar[stage]);
          if (cc == p.c_chunks - 1) umma_commit_pair(&tmem_full[acc]);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ---- epilogue (both CTAs): TMEM -> (+bias, +residual, act) -> bf16 -> swizzled staging -> TMA store ----
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const bool store_leader = (warp == 2 && lane == 0);
    const int row = q * 32 + lane;
    const int nbox = p.block_n >> 6;
    const int st_tid = threadIdx.x - 64;
    const int st_col = st_tid % p.block_n;
    const int st_rows = 128 / (256 / p.block_n);
    const int st_row0 = (st_tid / p.block_n) * st_rows;
    int st_ntile = -1;
    float st_s1 = 0.f, st_s2 = 0.f;
    int local = 0;
    for (int tile = cid; tile < total_tiles; tile += ncl, ++local) {
      const int acc = local & 1;
      const int m_pair = tile / p.n_tiles, n_tile = tile - m_pair * p.n_tiles;
      const int row0 = m_pair * 256 + static_cast<int>(rank) * 128;
      const int nbase = n_tile * p.block_n;
      if (store_leader && local > 0) bulk_wait_group_read0();
      named_bar_sync(1, 256);
      if (p.has_res) {
        if (store_leader) {
          mbar_arrive_expect_tx(&res_bar, static_cast<uint32_t>(nbox) * 128u * 128u);
          for (int b = 0; b < nbox; ++b) tma_load_2d(&tmR, &res_bar, epi + b * (128 * 128), nbase + b * 64, row0);
        }
        mbar_wait(&res_bar, static_cast<uint32_t>(local & 1));
      }
      mbar_wait(&tmem_full[acc], (local >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
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
        uint8_t* box = epi + (c0 >> 6) * (128 * 128) + row * 128;
        const int j0 = (c0 & 63) >> 3;
        uint4* p0 = reinterpret_cast<uint4*>(box + (((j0) ^ (row & 7)) << 4));
        uint4* p1 = reinterpret_cast<uint4*>(box + (((j0 + 1) ^ (row & 7)) << 4));
        if (p.has_res) {
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
      if (lane == 0) remote_arrive(leader_addr(&tmem_empty[acc]));   // for rank 0 this maps to its own barrier
      fence_proxy_async();
      named_bar_sync(1, 256);
      if (store_leader) {
        for (int b = 0; b < nbox; ++b) tma_store_2d(&tmC, epi + b * (128 * 128), nbase + b * 64, row0);
        bulk_commit_group();
      }
      if (p.stats != nullptr) {
        if (st_ntile != n_tile) {
          if (st_ntile >= 0) {
            double* dst = p.stats + (blockIdx.x % kPStatReplicas) * 2 * p.N_total + st_ntile * p.block_n + st_col;
            atomicAdd(dst, (double)st_s1);
            atomicAdd(dst + p.N_total, (double)st_s2);
          }
          st_ntile = n_tile; st_s1 = 0.f; st_s2 = 0.f;
        }
        const uint8_t* col = epi + (st_col >> 6) * (128 * 128) + (st_col & 7) * 2;
        const int j = (st_col & 63) >> 3;
        const int r_end = st_row0 + st_rows;
#pragma unroll 8
        for (int r = st_row0; r < r_end; ++r) {
          const float vv = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(col + r * 128 + ((j ^ (r & 7)) << 4)));
          st_s1 += vv;
          st_s2 = fmaf(vv, vv, st_s2);
        }
      }
    }
    if (p.stats != nullptr && st_ntile >= 0) {
      double* dst = p.stats + (blockIdx.x % kPStatReplicas) * 2 * p.N_total + st_ntile * p.block_n + st_col;
      atomicAdd(dst, (double)st_s1);
      atomicAdd(dst + p.N_total, (double)st_s2);
    }
    if (store_leader) bulk_wait_group0();
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  pair_cluster_sync();      // the peer may still be reading shared memory of this CTA through the leader's MMAs
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

static int enc2p(CUtensorMap* tm, const void* base, long long d0, long long d1, int b0, int b1) {
  EncodeTiledFn fn = encode_tiled_fn();
  B200_REQUIRE(fn != nullptr, B200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
  cuuint64_t strides[1] = {(cuuint64_t)d0 * 2};
  cuuint32_t box[2] = {(cuuint32_t)b0, (cuuint32_t)b1};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "conv pair: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return B200_OK;
}

// D[M, Nout] = A[M, C] * W[Nout, C]^T with dense bf16 rows; only wide layers profit (the kernel is opt-in)
bool pair_eligible(long long M, int C, int Nout) {
  static const bool enabled = getenv("B200_IGEMM_PAIR") && atoi(getenv("B200_IGEMM_PAIR")) != 0;
  if (!enabled) return false;
  if (C % 64 != 0 || C < 256 || Nout % 128 != 0 || M < 256) return false;
  const int n_tiles = (Nout + 255) / 256;
  return Nout % n_tiles == 0 && (Nout / n_tiles) % 128 == 0;
}

int launch_pair(const void* a, const void* w, void* out, const void* res, const float* bias, long long M, int C, int Nout,
                int act, double* stats, cudaStream_t stream) {
  PairParams p;
  memset(&p, 0, sizeof(p));
  p.M_total = (int)M; p.N_total = Nout; p.C = C;
  p.n_tiles = (Nout + 255) / 256;
  p.block_n = Nout / p.n_tiles;
  p.m_pairs = (int)((M + 255) / 256);
  p.c_chunks = C / 64;
  p.a_bytes = 128 * 128;
  p.bh_bytes = (uint32_t)(p.block_n / 2) * 128u;
  p.act = act; p.has_res = res != nullptr; p.bias = bias; p.stats = stats;
  const int epi_bytes = 128 * p.block_n * 2;
  p.num_stages = (212 * 1024 - epi_bytes) / (int)(p.a_bytes + p.bh_bytes);
  if (p.num_stages > kPMaxStages) p.num_stages = kPMaxStages;
  B200_REQUIRE(p.num_stages >= 2, B200_ERR_UNSUPPORTED, "conv pair: shared memory budget exceeded");
  CUtensorMap tmA, tmB, tmC, tmR;
  memset(&tmR, 0, sizeof(tmR));
  int rc = enc2p(&tmA, a, C, M, 64, 128);
  if (rc) return rc;
  rc = enc2p(&tmB, w, C, Nout, 64, p.block_n / 2);
  if (rc) return rc;
  rc = enc2p(&tmC, out, Nout, M, 64, 128);
  if (rc) return rc;
  if (res) {
    rc = enc2p(&tmR, res, Nout, M, 64, 128);
    if (rc) return rc;
  }
  const int smem_bytes = p.num_stages * (int)(p.a_bytes + p.bh_bytes) + epi_bytes + 1024;
  cudaError_t e = cudaFuncSetAttribute((const void*)conv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  B200_REQUIRE(e == cudaSuccess, B200_ERR_CUDA, "conv pair: smem attribute (%d bytes): %s", smem_bytes, cudaGetErrorString(e));
  const int total = p.m_pairs * p.n_tiles;
  int clusters = sm_count() / 2;
  if (clusters > total) clusters = total;
  conv_pair_kernel<<<2 * clusters, kPThreads, smem_bytes, stream>>>(tmA, tmB, tmC, tmR, p);
  B200_CHECK_LAUNCH("conv_pair_kernel");
  return B200_OK;
}

}  // namespace b200
