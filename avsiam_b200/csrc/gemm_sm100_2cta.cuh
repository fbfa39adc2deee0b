// Two-CTA (cta_group::2) variant of the persistent tcgen05 GEMM of gemm_sm100.cuh for the forward / dgrad products with
// bf16 outputs:  a CLUSTER of two CTAs (one SM pair) owns a 256 x 256 output tile.  Each CTA stages its own 128 rows of A
// and HALF of the B tile (128 of the 256 columns) — 32 KB per k-block instead of 48 KB — and the leader CTA's single
// issuing thread runs tcgen05.mma.cta_group::2 (M = 256), which reads both CTAs' shared memory and accumulates each CTA's
// 128 rows into that CTA's own tensor memory.  Why: the one-CTA kernel consumes 48 KB of operands per 512 tensor cycles,
// so its 192 KB ring holds ~2 000 cycles of MMAs — about one loaded-memory round trip — and the tensor pipe idles a third
// of the time on the K = 512 / 768 shapes (profiles/r02_gemm_bound_probe.log).  With 32 KB per
// k-block the same shared memory holds 6 stages (3 000 cycles) and the L2 -> SM operand traffic drops by a third.
//
// Protocol (the one of CUTLASS's 2-SM pipelines, restated):
//   full[s]    lives in the LEADER: it arms 2 x stage bytes; both CTAs' TMA loads (cp.async.bulk.tensor ... .cta_group::2)
//              complete their bytes on the leader's barrier.
//   empty[s]   one per CTA: the leader's tcgen05.commit.cta_group::2 ... multicast::cluster arrives on both when the MMAs
//              that read stage s have retired; each CTA's producer waits on its own.
//   tfull[a]   one per CTA (multicast commit): accumulator a of the tile is complete in both CTAs' tensor memory.
//   tempty[a]  lives in the leader (the issuer waits on it): 2 x 8 epilogue warps arrive, the peer's through
//              mbarrier.arrive ... shared::cluster.
// The epilogue is the one-CTA kernel's, per CTA, on its own 128 rows.
#pragma once
#include "gemm_sm100.cuh"

namespace avs {

struct Gemm2Cfg {
  static constexpr int BLOCK_N = 256;                                   // per pair; each CTA stages BLOCK_N / 2 columns of B
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;       // this CTA's 128 rows
  static constexpr int B_BYTES = (BLOCK_N / 2) * GEMM_BLOCK_K * 2;      // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;                 // 32 KB
  static constexpr int TMEM_COLS = 512;
  static constexpr int BAR_BYTES = 1024;   // [8 full][8 empty][2 tfull][2 tempty][16 x 4 in-stream] barriers + the TMEM slot
  static __host__ __device__ int epi_bytes_per_warp(int tma_epi, int has_in, int has_aux_out) {
    return GemmCfg<BLOCK_N>::epi_bytes_per_warp(tma_epi, has_in, has_aux_out);
  }
  // staging bytes of the whole CTA: 8 slots of [in ring][out][aux] with 8 epilogue warps; with 16 warps the 8 PAIRS share
  // the out (/ aux) tiles and every warp keeps its own in-stream ring
  static __host__ __device__ int staging_bytes(int ew, int has_in, int has_aux_out) {
    if (ew == 16)
      return 16 * (has_in ? GEMM_IN_DEPTH * GEMM_EPI_BUF : 0) + 8 * GEMM_OUT_BUF * (1 + (has_aux_out ? 1 : 0));
    return GEMM_EPI_WARPS * epi_bytes_per_warp(1, has_in, has_aux_out);
  }
  static __host__ int pick_stages(int staging, int extra = 0) {
    int s = (GEMM_SMEM_LIMIT - 1024 - BAR_BYTES - extra - staging) / STAGE_BYTES;
    return s > GEMM_MAX_STAGES ? GEMM_MAX_STAGES : s;
  }
  static __host__ int smem_bytes(int stages, int staging, int extra = 0) {
    return stages * STAGE_BYTES + staging + 1024 + BAR_BYTES + extra;
  }
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");   // (.release.cluster costs a MEMBAR.GPU)
}
// TMA load into THIS CTA's shared memory whose bytes complete on a barrier given by shared::cluster address (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared::cta offset in BOTH CTAs of the pair when the issued MMAs retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

// EW = epilogue warps per CTA: 8 (two per TMEM lane quarter, each walking half of the tile's columns through its own
// staging tiles) or, for the GELU class, 16: four per quarter, where the two warps of a PAIR fill the left / right 32 columns
// of the same 64-column staging tile and meet on a 64-thread named barrier — the staging memory, and with it the depth of
// the operand ring, stays what it is with 8 warps.  The GELU (+ derivative) epilogue is ~15 FP32 / MUFU instructions per
// element with two warps per scheduler eligible 44 % of the time; four warps per scheduler hide the
// TMEM-load / MUFU / barrier waits.
template <int B_MAJOR, int EPI, int EW = GEMM_EPI_WARPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_in,
                  const __grid_constant__ CUtensorMap tma_aux, const GemmArgs args) {
  using Cfg = Gemm2Cfg;
  constexpr int A_MAJOR = MAJOR_K;
  constexpr int BLOCK_N = Cfg::BLOCK_N;
  static_assert(EPI != GEMM_E_F32, "the pair kernel writes bf16 outputs (forward / dgrad products)");
  constexpr bool TMA_EPI = true;
  constexpr bool HAS_IN = EPI == GEMM_E_RESID || EPI == GEMM_E_MUL;
  const int STAGES = args.stages;
  extern __shared__ uint8_t smem_raw[];
  // identical shared-memory layout in both CTAs (the MMA addresses the peer's tiles by the same offsets): the dynamic
  // shared window starts at the same shared::cta address in every CTA of a launch
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  const int epi_per_warp = Cfg::epi_bytes_per_warp(args.tma_epi, args.has_in, args.has_aux_out);
  uint8_t* smem_epi = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + Cfg::staging_bytes(EW, args.has_in, args.has_aux_out));
  uint64_t* full_bar = bars;                               // [MAX_STAGES]   (the leader's are used)
  uint64_t* empty_bar = bars + GEMM_MAX_STAGES;            // [MAX_STAGES]
  uint64_t* tfull_bar = bars + 2 * GEMM_MAX_STAGES;        // [2]
  uint64_t* tempty_bar = bars + 2 * GEMM_MAX_STAGES + 2;   // [2]            (the leader's are used)
  uint64_t* in_bar = bars + 2 * GEMM_MAX_STAGES + 4;       // [EPI_WARPS][4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GEMM_MAX_STAGES + 4 + 4 * 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const int cluster_id = (int)(blockIdx.x >> 1);
  const int num_clusters = (int)(gridDim.x >> 1);

  const int m_tiles = (args.M + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M);   // 256-row tiles
  const int n_tiles = (args.N + BLOCK_N - 1) / BLOCK_N;
  const int total_kb = (args.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int mn_tiles = m_tiles * n_tiles;
  const int num_tiles = mn_tiles;                          // split_k == 1

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_c);
    if (HAS_IN) tma_prefetch_desc(&tma_in);
    if (EPI == GEMM_E_GELU && args.has_aux_out) tma_prefetch_desc(&tma_aux);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 2 * EW);
    }
    for (int i = 0; i < 4 * 16; ++i) mbar_init(&in_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();          // barriers initialised and tensor memory allocated in BOTH CTAs before anyone signals the peer
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tempty_leader = mapa_shared(smem_u32(&tempty_bar[0]), 0);

  if (warp == 0) {
    // ============================ TMA producer (both CTAs) ============================
    const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[0]), 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int m0 = (t / n_tiles) * (2 * GEMM_BLOCK_M) + (int)cta_rank * GEMM_BLOCK_M;
      const int nb0 = (t % n_tiles) * BLOCK_N + (int)cta_rank * (BLOCK_N / 2);   // this CTA's half of the B tile
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          const uint32_t fb = full_leader + (uint32_t)stage * 8u;
          uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
          uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
          const int k0 = kb * GEMM_BLOCK_K;
          tma_load_2d_pair(sa, &tma_a, fb, k0, m0);
          if (B_MAJOR == MAJOR_K) {
            tma_load_2d_pair(sb, &tma_b, fb, k0, nb0);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 2 / 64; ++j)
              tma_load_2d_pair(sb + j * (GEMM_BLOCK_K * 128), &tma_b, fb, nb0 + j * 64, k0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA only) ============================
    if (cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BLOCK_M, BLOCK_N, A_MAJOR, B_MAJOR);
      constexpr uint32_t mn_lbo = GEMM_BLOCK_K * 128, mn_sbo = 1024;
      const uint64_t da0 = make_smem_desc(smem_u32(smem_a), 0, 1024);
      const uint64_t db0 = (B_MAJOR == MAJOR_K) ? make_smem_desc(smem_u32(smem_b), 0, 1024)
                                                : make_smem_desc(smem_u32(smem_b), mn_lbo, mn_sbo);
      constexpr uint32_t A_K16 = 32 >> 4, B_K16 = ((B_MAJOR == MAJOR_K) ? 32 : 2048) >> 4;
      constexpr uint32_t A_STAGE16 = Cfg::A_BYTES >> 4, B_STAGE16 = Cfg::B_BYTES >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint64_t da = da0 + (uint64_t)((uint32_t)stage * A_STAGE16);
            const uint64_t db = db0 + (uint64_t)((uint32_t)stage * B_STAGE16);
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k)
              umma_bf16_ss_pair(tmem_d, da + (uint64_t)(k * A_K16), db + (uint64_t)(k * B_K16), idesc,
                                (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage]);                      // frees the stage in both CTAs
            if (kb == total_kb - 1) umma_commit_pair(&tfull_bar[acc]);  // accumulator complete in both CTAs
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if constexpr (EW == 16 && EPI == GEMM_E_MUL) {
    // ============================ epilogue, 16 warps in pairs (product class) ============================
    // out = acc * in-stream (the stored gelu', or gelu'(pre) evaluated here) with the fused column sums; every warp keeps
    // its own ring of [32 x 32] in-stream tiles (its two chunks of a tile, requested one tile ahead), pairs share the
    // [32 x 64] output staging tiles.
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int kq = ew >> 2;
    const int pairi = kq >> 1;
    const int side = kq & 1;
    const int slot = quarter * 2 + pairi;
    const int bar_id = 1 + slot;           // pair barriers 1..8; 9 = all 16 epilogue warps (column sums)
    const GemmEpilogue& ep = args.epi;
    uint8_t* in_buf = smem_epi + ew * (GEMM_IN_DEPTH * GEMM_EPI_BUF);
    uint8_t* out_buf = smem_epi + 16 * (GEMM_IN_DEPTH * GEMM_EPI_BUF) + slot * GEMM_OUT_BUF;
    uint64_t* my_in_bar = in_bar + 4 * ew;
    float* s_colsum = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + Cfg::BAR_BYTES);   // [2][4][256]
    auto issue_in = [&](int q) {
      const int t = cluster_id + (q >> 1) * num_clusters;
      if (t >= num_tiles) return;
      const int m0 = (t / n_tiles) * (2 * GEMM_BLOCK_M) + (int)cta_rank * GEMM_BLOCK_M;
      const int col = (t % n_tiles) * BLOCK_N + pairi * 128 + (q & 1) * 64 + side * 32;
      const int sl = q % GEMM_IN_DEPTH;
      mbar_arrive_expect_tx(&my_in_bar[sl], GEMM_EPI_BUF);
      tma_load_2d(in_buf + sl * GEMM_EPI_BUF, &tma_in, &my_in_bar[sl], col, m0 + quarter * 32);
    };
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < GEMM_IN_DEPTH; ++i) issue_in(i);
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    int q = 0;
    int tile_par = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int m0 = (t / n_tiles) * (2 * GEMM_BLOCK_M) + (int)cta_rank * GEMM_BLOCK_M;
      const int n0 = (t % n_tiles) * BLOCK_N;
      const bool row_ok = (m0 + quarter * 32 + lane) < args.M;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int b = 0; b < 2; ++b, ++q) {
        const int ccol = pairi * 128 + b * 64 + side * 32;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N + ccol), r);
        tmem_ld_wait();
        if (b == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader + (uint32_t)acc * 8u);
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        uint4 in4[4];
        {
          const int in_slot = q % GEMM_IN_DEPTH;
          mbar_wait(&my_in_bar[in_slot], (uint32_t)((q / GEMM_IN_DEPTH) & 1));
#pragma unroll
          for (int j = 0; j < 4; ++j)
            in4[j] = *reinterpret_cast<const uint4*>(in_buf + in_slot * GEMM_EPI_BUF + epi_tile_off(lane, j));
          __syncwarp();                     // every lane has read its row: the tile may be refilled
          if (lane == 0) issue_in(q + GEMM_IN_DEPTH);
        }
        if (ep.flags & EPI_MUL_AUX) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f;
            f = unpack_bf16x2(in4[j].x); v[8 * j] *= f.x; v[8 * j + 1] *= f.y;
            f = unpack_bf16x2(in4[j].y); v[8 * j + 2] *= f.x; v[8 * j + 3] *= f.y;
            f = unpack_bf16x2(in4[j].z); v[8 * j + 4] *= f.x; v[8 * j + 5] *= f.y;
            f = unpack_bf16x2(in4[j].w); v[8 * j + 6] *= f.x; v[8 * j + 7] *= f.y;
          }
        } else {   // EPI_DGELU
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f;
            f = unpack_bf16x2(in4[j].x); v[8 * j] *= dgelu_erf(f.x); v[8 * j + 1] *= dgelu_erf(f.y);
            f = unpack_bf16x2(in4[j].y); v[8 * j + 2] *= dgelu_erf(f.x); v[8 * j + 3] *= dgelu_erf(f.y);
            f = unpack_bf16x2(in4[j].z); v[8 * j + 4] *= dgelu_erf(f.x); v[8 * j + 5] *= dgelu_erf(f.y);
            f = unpack_bf16x2(in4[j].w); v[8 * j + 6] *= dgelu_erf(f.x); v[8 * j + 7] *= dgelu_erf(f.y);
          }
        }
        if (ep.alpha != 1.0f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= ep.alpha;
        }
        uint4 o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[j].x = pack_bf16x2(v[8 * j], v[8 * j + 1]); o[j].y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
          o[j].z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); o[j].w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        }
        if (ep.colsum != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = row_ok ? v[j] : 0.f;
          warp_colsum<32>(v, lane);                   // lane L now holds the sum of column L over the warp's 32 rows
          s_colsum[((tile_par * 4 + quarter) << 8) + ccol + lane] = v[0];
        }
        if (side == 0 && lane == 0) bulk_wait_read0();
        asm volatile("bar.sync %0, 64;\n" ::"r"(bar_id) : "memory");
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(out_buf + out_tile_off(lane, side * 4 + j)) = o[j];
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 64;\n" ::"r"(bar_id) : "memory");
        if (side == 0 && lane == 0) {
          const int nc0 = n0 + pairi * 128 + b * 64;
          if (nc0 < args.N && m0 + quarter * 32 < args.M) tma_store_2d(&tma_c, out_buf, nc0, m0 + quarter * 32);
          bulk_commit();
        }
      }
      if (ep.colsum != nullptr) {
        asm volatile("bar.sync 9, 512;\n" ::: "memory");      // every epilogue warp has added its rows of this tile
        const int et = ew * 32 + lane;
        if (et < BLOCK_N && n0 + et < args.N) {
          const float* sc = s_colsum + (tile_par << 10) + et;
          atomicAdd(ep.colsum + n0 + et, (sc[0] + sc[256]) + (sc[512] + sc[768]));
        }
        tile_par ^= 1;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (side == 0 && lane == 0) bulk_wait0();
  } else if constexpr (EW == 16) {
    // ============================ epilogue, 16 warps in pairs (GELU class) ============================
    static_assert(EW != 16 || EPI == GEMM_E_GELU || EPI == GEMM_E_MUL, "16 epilogue warps: GELU and product classes");
    const int ew = warp - 2;
    const int quarter = warp & 3;          // TMEM lane quarter (= scheduler) of this warp
    const int kq = ew >> 2;                // 0..3: which of the quarter's four warps
    const int pairi = kq >> 1;             // the pair's half of the tile's columns (128 each)
    const int side = kq & 1;               // left / right 32 columns of each 64-column block
    const int slot = quarter * 2 + pairi;  // staging slot shared by the pair
    const int bar_id = 1 + slot;           // named barrier of the pair (64 threads)
    const GemmEpilogue& ep = args.epi;
    uint8_t* out_buf = smem_epi + slot * epi_per_warp;
    uint8_t* aux_buf = out_buf + GEMM_OUT_BUF;
    const bool has_aux = args.has_aux_out != 0;
    const bool aux_grad = has_aux && (ep.flags & EPI_AUX_GRAD);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int m0 = (t / n_tiles) * (2 * GEMM_BLOCK_M) + (int)cta_rank * GEMM_BLOCK_M;
      const int n0 = (t % n_tiles) * BLOCK_N;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int b = 0; b < 2; ++b) {
        const int ccol = pairi * 128 + b * 64 + side * 32;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N + ccol), r);
        tmem_ld_wait();
        if (b == 1) {  // accumulator fully read by this warp: hand it back to the pair's issuer (CTA 0)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader + (uint32_t)acc * 8u);
        }
        const int nc = n0 + ccol;
        const bool col_ok = nc < args.N;
        const bool full = (nc + 32 <= args.N);
        uint4 o[4], ax[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {      // 16 columns at a time: bounded live registers (112 per thread at 576 threads)
          float v[16], dg[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[16 * h + j]);
          if (ep.bias != nullptr && col_ok) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              if (full || nc + 16 * h + j < args.N) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(ep.bias + nc + 16 * h + j));
                v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
              }
            }
          }
          if (aux_grad) {
#pragma unroll
            for (int j = 0; j < 16; ++j) gelu_and_grad(v[j], v[j], dg[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) { dg[j] = v[j]; v[j] = gelu_erf(v[j]); }   // aux (if any) = pre-activation
          }
          if (ep.alpha != 1.0f) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] *= ep.alpha;
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            o[2 * h + j].x = pack_bf16x2(v[8 * j], v[8 * j + 1]); o[2 * h + j].y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
            o[2 * h + j].z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); o[2 * h + j].w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
            ax[2 * h + j].x = pack_bf16x2(dg[8 * j], dg[8 * j + 1]); ax[2 * h + j].y = pack_bf16x2(dg[8 * j + 2], dg[8 * j + 3]);
            ax[2 * h + j].z = pack_bf16x2(dg[8 * j + 4], dg[8 * j + 5]); ax[2 * h + j].w = pack_bf16x2(dg[8 * j + 6], dg[8 * j + 7]);
          }
        }
        // the pair's previous stores have finished reading the staging tiles (only the issuing thread holds bulk groups)
        if (side == 0 && lane == 0) bulk_wait_read0();
        asm volatile("bar.sync %0, 64;\n" ::"r"(bar_id) : "memory");
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(out_buf + out_tile_off(lane, side * 4 + j)) = o[j];
        if (has_aux) {
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(aux_buf + out_tile_off(lane, side * 4 + j)) = ax[j];
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 64;\n" ::"r"(bar_id) : "memory");
        if (side == 0 && lane == 0) {
          const int nc0 = n0 + pairi * 128 + b * 64;   // first column of the 64-column block
          if (nc0 < args.N && m0 + quarter * 32 < args.M) {   // TMA clips the M / N tails
            tma_store_2d(&tma_c, out_buf, nc0, m0 + quarter * 32);
            if (has_aux) tma_store_2d(&tma_aux, aux_buf, nc0, m0 + quarter * 32);
          }
          bulk_commit();
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (side == 0 && lane == 0) bulk_wait0();  // all stores retired before the CTA exits
  } else {
    // ============================ epilogue (8 warps) ============================
    // warp w may only touch TMEM lanes 32*(w%4)..+32; the two warps sharing a quarter split the tile's columns.
    const int ew = warp - 2;
    // column sums of the output (the bias gradient of the upstream Linear): every epilogue warp writes the sums over
    // its 32 rows into its lane quarter's slot [tile parity][quarter][column] (plain stores — shared-memory float
    // atomics are compare-and-swap loops), one barrier per tile, then one global atomic per column and tile
    float* s_colsum = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + Cfg::BAR_BYTES);   // [2][4][256]
    const int quarter = warp & 3;
    const int half = ew >> 2;
    constexpr int CH = BLOCK_N / (2 * GEMM_EPI_CHUNK);  // chunks per warp per tile
    const GemmEpilogue& ep = args.epi;
    uint8_t* my_epi = smem_epi + ew * epi_per_warp;
    uint8_t* in_buf = my_epi;                                        // [2][2 KB] when has_in
    uint8_t* out_buf = my_epi + (HAS_IN ? GEMM_IN_DEPTH * GEMM_EPI_BUF : 0);   // 1024-byte aligned (128 B swizzle pattern)
    uint8_t* aux_buf = out_buf + GEMM_OUT_BUF;
    uint64_t* my_in_bar = in_bar + 4 * ew;
    constexpr bool tma_epi = TMA_EPI;
    constexpr bool has_in = HAS_IN;

    // flat per-warp chunk sequence q = tile_iteration * CH + chunk; `in` tiles are prefetched two chunks ahead
#ifdef AVS_GEMM_DEBUG
    const bool dbg_no_in = (args.desc_variant & 4) != 0;   // timing experiment: no in-stream TMA loads (wrong results)
    const bool dbg_no_store = (args.desc_variant & 2) != 0;
#else
    constexpr bool dbg_no_in = false, dbg_no_store = false;
#endif
    auto issue_in = [&](int q) {
      if (dbg_no_in) return;
      const int t = cluster_id + (q / CH) * num_clusters;
      if (t >= num_tiles) return;
      const int mn = t % mn_tiles;
      const int m0 = (mn / n_tiles) * (2 * GEMM_BLOCK_M) + (int)cta_rank * GEMM_BLOCK_M;
      const int n0 = (mn % n_tiles) * BLOCK_N;
      const int col = n0 + half * (BLOCK_N / 2) + (q % CH) * GEMM_EPI_CHUNK;
      const int slot = q % GEMM_IN_DEPTH;
      mbar_arrive_expect_tx(&my_in_bar[slot], GEMM_EPI_BUF);
      tma_load_2d(in_buf + slot * GEMM_EPI_BUF, &tma_in, &my_in_bar[slot], col, m0 + quarter * 32);
    };
    if (has_in && lane == 0) {
#pragma unroll
      for (int i = 0; i < GEMM_IN_DEPTH; ++i) issue_in(i);
    }

    int acc = 0;
    uint32_t acc_phase = 0;
    int q = 0;
    int tile_par = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int ks = t / mn_tiles;   // output tile fastest: the CTAs that run together share a k-range (see mn_tiles)
      const int mn = t - ks * mn_tiles;
      const int m0 = (mn / n_tiles) * (2 * GEMM_BLOCK_M) + (int)cta_rank * GEMM_BLOCK_M;
      const int n0 = (mn % n_tiles) * BLOCK_N;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < args.M;
      const float* rowadd_ptr = nullptr;
      if ((EPI == GEMM_E_PLAIN || EPI == GEMM_E_F32) && ep.rowadd != nullptr && row_ok) {
        const int ri = ep.rowidx ? ep.rowidx[row] : (row % ep.rowadd_rows);
        rowadd_ptr = ep.rowadd + (long long)ri * args.N;
      }
      const bool lead_split = (ks == 0);
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N + half * (BLOCK_N / 2));
      uint32_t rn[GEMM_TMEM_PF ? 32 : 1];
      if constexpr (GEMM_TMEM_PF) tmem_ld_32x32b_x32(taddr0, reinterpret_cast<uint32_t(&)[32]>(rn));
#pragma unroll 1
      for (int c = 0; c < CH; ++c, ++q) {
        const int ccol = half * (BLOCK_N / 2) + c * GEMM_EPI_CHUNK;  // column offset inside the tile
        uint32_t r[32];
        if constexpr (GEMM_TMEM_PF) {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = rn[j];
          if (c + 1 < CH) tmem_ld_32x32b_x32(taddr0 + (uint32_t)((c + 1) * GEMM_EPI_CHUNK), reinterpret_cast<uint32_t(&)[32]>(rn));
        } else {
          tmem_ld_32x32b_x32(taddr0 + (uint32_t)(c * GEMM_EPI_CHUNK), r);
          tmem_ld_wait();
        }
        if (c == CH - 1) {  // accumulator fully read by this warp: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader + (uint32_t)acc * 8u);   // the pair's issuer lives in CTA 0
        }
        const int nc = n0 + ccol;
        const bool col_ok = nc < args.N;
        const bool full = (nc + 32 <= args.N);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (ep.bias != nullptr && lead_split && col_ok) {
          if (full) {   // interior chunk: no per-group predicates (the epilogue is issue-bound)
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + nc + j));
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (nc + j < args.N) {
                const float4 b = *reinterpret_cast<const float4*>(ep.bias + nc + j);
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
            }
          }
        }
        if ((EPI == GEMM_E_PLAIN || EPI == GEMM_E_F32) && rowadd_ptr != nullptr && lead_split && col_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (full || nc + j < args.N) {
              const float4 b = *reinterpret_cast<const float4*>(rowadd_ptr + nc + j);
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
        }
        if constexpr (tma_epi) {
          // ---------------- bf16 outputs: staged in smem, moved by TMA ----------------
          // Everything is computed in registers first; the wait for the previous chunk's TMA stores (they read the
          // staging tiles) comes as late as possible so that it overlaps this chunk's math.
          uint4 ax[4];
          if constexpr (EPI == GEMM_E_GELU) {
            if (args.has_aux_out && (ep.flags & EPI_AUX_GRAD)) {
              // the derivative is evaluated here, where the tanh is already paid for, and stored instead of the
              // pre-activation: the dgrad GEMM of fc2 then only multiplies
              float dg[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) gelu_and_grad(v[j], v[j], dg[j]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                ax[j].x = pack_bf16x2(dg[8 * j], dg[8 * j + 1]); ax[j].y = pack_bf16x2(dg[8 * j + 2], dg[8 * j + 3]);
                ax[j].z = pack_bf16x2(dg[8 * j + 4], dg[8 * j + 5]); ax[j].w = pack_bf16x2(dg[8 * j + 6], dg[8 * j + 7]);
              }
            } else {
              if (args.has_aux_out) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  ax[j].x = pack_bf16x2(v[8 * j], v[8 * j + 1]); ax[j].y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                  ax[j].z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); ax[j].w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                }
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
            }
          }
          uint4 in4[4];
          if constexpr (has_in) {
            const int in_slot = q % GEMM_IN_DEPTH;
            if (!dbg_no_in) mbar_wait(&my_in_bar[in_slot], (uint32_t)((q / GEMM_IN_DEPTH) & 1));
#pragma unroll
            for (int j = 0; j < 4; ++j)
              in4[j] = *reinterpret_cast<const uint4*>(in_buf + in_slot * GEMM_EPI_BUF + epi_tile_off(lane, j));
            __syncwarp();                     // every lane has read its row: the tile may be refilled
            if (lane == 0) issue_in(q + GEMM_IN_DEPTH);
          }
          if constexpr (EPI == GEMM_E_MUL) {
           if (ep.flags & EPI_MUL_AUX) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 f;
              f = unpack_bf16x2(in4[j].x); v[8 * j] *= f.x; v[8 * j + 1] *= f.y;
              f = unpack_bf16x2(in4[j].y); v[8 * j + 2] *= f.x; v[8 * j + 3] *= f.y;
              f = unpack_bf16x2(in4[j].z); v[8 * j + 4] *= f.x; v[8 * j + 5] *= f.y;
              f = unpack_bf16x2(in4[j].w); v[8 * j + 6] *= f.x; v[8 * j + 7] *= f.y;
            }
           } else {   // EPI_DGELU
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 f;
              f = unpack_bf16x2(in4[j].x); v[8 * j] *= dgelu_erf(f.x); v[8 * j + 1] *= dgelu_erf(f.y);
              f = unpack_bf16x2(in4[j].y); v[8 * j + 2] *= dgelu_erf(f.x); v[8 * j + 3] *= dgelu_erf(f.y);
              f = unpack_bf16x2(in4[j].z); v[8 * j + 4] *= dgelu_erf(f.x); v[8 * j + 5] *= dgelu_erf(f.y);
              f = unpack_bf16x2(in4[j].w); v[8 * j + 6] *= dgelu_erf(f.x); v[8 * j + 7] *= dgelu_erf(f.y);
            }
           }
          }
          if (ep.alpha != 1.0f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= ep.alpha;
          }
          if constexpr (EPI == GEMM_E_RESID) {  // residual add
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 f;
              f = unpack_bf16x2(in4[j].x); v[8 * j] += f.x; v[8 * j + 1] += f.y;
              f = unpack_bf16x2(in4[j].y); v[8 * j + 2] += f.x; v[8 * j + 3] += f.y;
              f = unpack_bf16x2(in4[j].z); v[8 * j + 4] += f.x; v[8 * j + 5] += f.y;
              f = unpack_bf16x2(in4[j].w); v[8 * j + 6] += f.x; v[8 * j + 7] += f.y;
            }
          }
          if (EPI == GEMM_E_MUL && ep.colsum != nullptr) {
            float cs[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) cs[j] = row_ok ? v[j] : 0.f;
            warp_colsum<32>(cs, lane);                  // lane L now holds the sum of column L over the warp's 32 rows
            s_colsum[((tile_par * 4 + quarter) << 8) + ccol + lane] = cs[0];
          }
          uint4 o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            o[j].x = pack_bf16x2(v[8 * j], v[8 * j + 1]); o[j].y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
            o[j].z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); o[j].w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
          }
          const int hsel = c & 1;   // which 64-byte half of the 128-byte staging rows this chunk fills
          if (hsel == 0) {
            if (lane == 0) bulk_wait_read0();  // the previous pair's stores have finished reading the staging tiles
            __syncwarp();
          }
          if (EPI == GEMM_E_GELU && args.has_aux_out) {
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(aux_buf + out_tile_off(lane, hsel * 4 + j)) = ax[j];
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(out_buf + out_tile_off(lane, hsel * 4 + j)) = o[j];
          if (hsel == 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              const int nc0 = nc - GEMM_EPI_CHUNK;   // first column of the pair
              if (nc0 < args.N && m0 + quarter * 32 < args.M && !dbg_no_store) {  // TMA clips the M / N tails
                tma_store_2d(&tma_c, out_buf, nc0, m0 + quarter * 32);
                if (EPI == GEMM_E_GELU && args.has_aux_out) tma_store_2d(&tma_aux, aux_buf, nc0, m0 + quarter * 32);
              }
              bulk_commit();
            }
          }
        } else if (row_ok && col_ok) {
          // ---------------- fp32 outputs (wgrad / split-K accumulation): direct stores ----------------
          if (ep.alpha != 1.0f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= ep.alpha;
          }
          float* cp = reinterpret_cast<float*>(args.C) + (long long)row * args.ldc + nc;
          if (ep.flags & EPI_OUT_ATOMIC) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (full || nc + j < args.N) red_add_f32x4(cp + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (full || nc + j < args.N)
                *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
      }
      if (EPI == GEMM_E_MUL && ep.colsum != nullptr) {
        asm volatile("bar.sync 1, 256;\n" ::: "memory");      // every epilogue warp has added its rows of this tile
        const int et = ew * 32 + lane;
        if (et < BLOCK_N && n0 + et < args.N) {
          const float* sc = s_colsum + (tile_par << 10) + et;
          atomicAdd(ep.colsum + n0 + et, (sc[0] + sc[256]) + (sc[512] + sc[768]));
        }
        tile_par ^= 1;   // the next tile writes the other buffer: its barrier orders the one after against these reads
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (tma_epi && lane == 0) bulk_wait0();  // all stores retired before the CTA exits
  }

  tc_fence_before();
  cluster_sync_all();     // neither CTA leaves (or frees tensor memory) while the peer may still read its tiles / signal it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace avs
