// tcgen05 / TMEM attention for the SHORT sequences of the shared encoder (head_dim 64, S <= 128): audio 128 kept
// tokens, video 49 kept tokens, and the mixed-ratio chunks of forward_encoder_mmixed.
//
// Replaces F.scaled_dot_product_attention forward + backward in Attention.forward (cav_mae_base.py:58-77) for these
// shapes. A whole head is ONE 128 x 128 score tile, so there is no online softmax and no streaming: what bounded the
// previous kernels (one short-lived CTA per head: TMEM allocation, barrier set-up and a cold pipeline for ~3 k cycles
// of work; mma.sync tiles for S < 96) is removed by making the CTAs PERSISTENT — they walk over (sequence, head)
// tiles with the operands of the next tile already in flight (TMA, two shared-memory stages) — and by packing TWO
// sequences of S <= 64 tokens into one tile (block-diagonal mask), so the video sequences fill the 128 TMEM lanes.
//
// Tile = two 64-row slots. S <= 64: slot k holds sequence 2p+k; 64 < S <= 128: the slots hold rows [0,64) / [64,128) of
// sequence p. Rows a TMA box reads past the end of its sequence belong to the next sequence (finite values, masked:
// their probabilities are exactly 0) or lie outside the tensor (zero fill).
//
// Forward  (6 warps, 2 CTAs / SM, 256 TMEM columns): S = Q K^T -> thread-per-row softmax (two passes over tensor
//           memory: maximum, then exponentials) -> P (bf16) written back over S in tensor memory -> O = P V with the A
//           operand read from tensor memory -> O / rowsum stored straight from registers (128 contiguous bytes per row).
// Backward (10 warps, 1 CTA / SM, 448 TMEM columns): S = Q K^T and dP = dO V^T -> P = exp2(S c - lse),
//           delta = rowsum(P o dP) taken in the same pass (no O read, no delta pre-pass), dS = P o (dP - delta) / sqrt(hd)
//           -> P and dS (bf16) into swizzled shared-memory tiles laid out so that the SAME bytes serve as the MN-major A
//           operand of dV = P^T dO / dK = dS^T Q and as the K-major A operand of dQ = dS K -> the three gradients are
//           read from tensor memory and stored; the qkv-bias gradient (column sums) is accumulated in shared memory
//           over all tiles of the CTA and flushed once.
// Both are HBM-bound by design (forward 64 KB, backward 112 KB per head in / out); tensor and MUFU work hide under it.
#include "../../include/avsiam_b200.h"
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int HD = 64;
constexpr int SLOT_ROWS = 64;
constexpr int MAT_BYTES = 128 * HD * 2;        // one [128 x 64] bf16 operand tile (128-byte rows, SWIZZLE_128B) = 16 KB
constexpr int SLOT_BYTES = SLOT_ROWS * HD * 2;  // 8 KB

struct SmallArgs {
  int n_seq, S, H, D;      // D = H * 64
  int pair;                // 1: two sequences per tile (S <= 64)
  int n_tiles;
  long long ld_o, ld_qkv;
  bf16* out;               // fwd: [rows, ld_o]
  float* lse2;             // [n_seq, H, S] (log2 domain: max * scale_log2 + log2(rowsum))
  bf16* dqkv;              // bwd: [rows, ld_qkv]
  float* dbias;            // bwd: [3 D] or null
  float scale, scale_log2;
};

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// tile -> rows. row0_k = first global token row loaded into slot k (may lie past the tensor: TMA zero-fills), seq_k =
// the sequence of slot k (-1: none). Scalars, not arrays: a slot index that is only known at run time would put an
// array into local memory.
struct TileGeom {
  int head;
  int row0_0, row0_1;
  int seq0, seq1;
};
__device__ __forceinline__ TileGeom tile_geom(const SmallArgs& a, int head, int p) {
  TileGeom g;
  g.head = head;
  if (a.pair) {
    g.seq0 = 2 * p;
    g.seq1 = (2 * p + 1 < a.n_seq) ? 2 * p + 1 : -1;
    g.row0_0 = 2 * p * a.S;
    g.row0_1 = (2 * p + 1) * a.S;
  } else {
    g.seq0 = g.seq1 = p;
    g.row0_0 = p * a.S;
    g.row0_1 = p * a.S + SLOT_ROWS;
  }
  return g;
}

// Per-thread view of the tile: query row i (= TMEM lane), its global row and validity, and how many key columns of
// each 32-column chunk it may attend to.  Everything that decides whether a tcgen05.ld is executed is warp-uniform (a
// warp's 32 rows share a slot).
struct RowView {
  bool row_valid;
  long long grow;    // global token row of query i
  int seq, r_local;  // sequence and position inside it
  int c_lo, c_hi;    // valid key columns [c_lo, c_hi) of the 128-column tile (warp-uniform)
};
__device__ __forceinline__ RowView row_view(const SmallArgs& a, const TileGeom& g, int i) {
  RowView v;
  if (a.pair) {
    const int slot = i >> 6, r = i & 63;
    v.seq = slot ? g.seq1 : g.seq0;
    v.r_local = r;
    v.row_valid = (v.seq >= 0) && r < a.S;
    v.grow = (long long)(slot ? g.row0_1 : g.row0_0) + r;
    v.c_lo = slot * 64;
    v.c_hi = (v.seq >= 0) ? slot * 64 + a.S : slot * 64;
  } else {
    v.seq = g.seq0;
    v.r_local = i;
    v.row_valid = i < a.S;
    v.grow = (long long)g.row0_0 + i;
    v.c_lo = 0;
    v.c_hi = a.S;
  }
  return v;
}
// valid columns of 32-column chunk c for a row with valid range [c_lo, c_hi): 0..32 (c_lo is a multiple of 64)
__device__ __forceinline__ int chunk_valid(int c, int c_lo, int c_hi) {
  if (32 * c < c_lo) return 0;
  const int n = c_hi - 32 * c;
  return n < 0 ? 0 : (n > 32 ? 32 : n);
}

// =====================================================================================================
// Forward
// =====================================================================================================
constexpr int SF_SOFTMAX_WARPS = 8;
constexpr int SF_THREADS = 32 * (2 + SF_SOFTMAX_WARPS);   // warp 0 loader, warp 1 MMA issuer, warps 2-9 softmax / epilogue
constexpr int SF_STAGE_BYTES = 3 * MAT_BYTES;             // Q | K | V
constexpr int SF_SMEM = 2 * SF_STAGE_BYTES + 1024 /* alignment */ + 2 * 2 * 128 * 4 /* max / sum exchange */ + 256;
// P gets its own columns: with two warps per lane quarter a P chunk written over S could land on columns the partner
// warp has not read yet
constexpr int SF_COL_S = 0, SF_COL_O = 128, SF_COL_P = 192, SF_TMEM_COLS = 256;

// tile t of this CTA's contiguous range: head-major, so a CTA stays on one head (two at most)
__device__ __forceinline__ void tile_range(const SmallArgs& a, int& t0, int& t1) {
  t0 = (int)((long long)blockIdx.x * a.n_tiles / gridDim.x);
  t1 = (int)((long long)(blockIdx.x + 1) * a.n_tiles / gridDim.x);
}

__global__ void __launch_bounds__(SF_THREADS, 2)
attn_small_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const SmallArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* xch = reinterpret_cast<float*>(smem + 2 * SF_STAGE_BYTES);   // [max | sum][2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 2 * 2 * 128);
  uint64_t* full = bars;          // [2]
  uint64_t* empty = bars + 2;     // [2]
  uint64_t* s_ready = bars + 4;
  uint64_t* p_ready = bars + 5;
  uint64_t* o_ready = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int groups = a.n_tiles / a.H;
  int t0, t1;
  tile_range(a, t0, t1);

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_qkv);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(s_ready, 1);
    mbar_init(p_ready, SF_SOFTMAX_WARPS);
    mbar_init(o_ready, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, SF_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- loader
    int it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
      const int st = it & 1;
      mbar_wait_backoff(&empty[st], ((it >> 1) & 1) ^ 1, 64);
      if (elect_one_sync()) {
        const TileGeom g = tile_geom(a, t / groups, t % groups);
        uint8_t* base = smem + st * SF_STAGE_BYTES;
        mbar_arrive_expect_tx(&full[st], SF_STAGE_BYTES);
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          tma_load_2d(base + m * MAT_BYTES, &tm_qkv, &full[st], m * a.D + g.head * HD, g.row0_0);
          tma_load_2d(base + m * MAT_BYTES + SLOT_BYTES, &tm_qkv, &full[st], m * a.D + g.head * HD, g.row0_1);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_s = idesc_bf16(128, 128, 0, 0);   // S = Q K^T : both operands K-major (head_dim contiguous)
    constexpr uint32_t idesc_o = idesc_bf16(128, HD, 0, 1);    // O = P V   : A in tensor memory, V [keys x hd] MN-major
    int it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
      const int st = it & 1;
      const uint32_t sb = smem_u32(smem + st * SF_STAGE_BYTES);
      mbar_wait(&full[st], (it >> 1) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t dq = desc_sw128(sb, 0, 1024), dk = desc_sw128(sb + MAT_BYTES, 0, 1024);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16_ss(tmem + SF_COL_S, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
        umma_commit(s_ready);
      }
      __syncwarp();
      mbar_wait(p_ready, it & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t dv = desc_sw128(sb + 2 * MAT_BYTES, 128 * 128, 1024);
#pragma unroll
        for (int k = 0; k < 128 / 16; ++k)
          umma_bf16_ts(tmem + SF_COL_O, tmem + SF_COL_P + 8 * k, dv + (uint64_t)(k * (2048 >> 4)), idesc_o, k > 0);
        umma_commit(o_ready);
        umma_commit(&empty[st]);
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue
    // thread = (query row i = TMEM lane, half): the two warps of a lane quarter take the 32-column chunks {half, half+2}
    // of the 128 key columns (interleaved, so that a packed pair of 64-token sequences also splits evenly) and agree on
    // the row maximum and the row sum through shared memory and a 64-thread named barrier.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int i = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    float* xmax = xch;
    float* xsum = xch + 256;
    int it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
      const TileGeom g = tile_geom(a, t / groups, t % groups);
      const RowView rv = row_view(a, g, i);
      const int c_lo = __shfl_sync(0xffffffffu, rv.c_lo, 0), c_hi = __shfl_sync(0xffffffffu, rv.c_hi, 0);   // uniform
      mbar_wait(s_ready, it & 1);
      tc_fence_after();
      // pass 1: maximum over this thread's chunks
      float m = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half + 2 * cc;
        const int nv = chunk_valid(c, c_lo, c_hi);
        if (nv > 0) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(lane_addr + SF_COL_S + 32 * c, r);
          tmem_ld_wait();
          if (nv == 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(r[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nv) m = fmaxf(m, __uint_as_float(r[j]));
          }
        }
      }
      xmax[half * 128 + i] = m;
      asm volatile("bar.sync %0, 64;\n" ::"r"(1 + quarter) : "memory");
      m = fmaxf(m, xmax[(half ^ 1) * 128 + i]);
      if (m == -INFINITY) m = 0.f;   // slot without a sequence
      const float ms = m * a.scale_log2;
      // pass 2: exponentials, P (bf16) into tensor memory
      float sum = 0.f;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half + 2 * cc;
        const int nv = chunk_valid(c, c_lo, c_hi);
        uint32_t pk[16];
        if (nv > 0) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(lane_addr + SF_COL_S + 32 * c, r);
          tmem_ld_wait();
          if (nv == 32) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float p0 = exp2f(fmaf(__uint_as_float(r[j]), a.scale_log2, -ms));
              const float p1 = exp2f(fmaf(__uint_as_float(r[j + 1]), a.scale_log2, -ms));
              sum += p0 + p1;
              pk[j >> 1] = pack_bf16x2(p0, p1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float p0 = exp2f(fmaf(__uint_as_float(r[j]), a.scale_log2, -ms));
              float p1 = exp2f(fmaf(__uint_as_float(r[j + 1]), a.scale_log2, -ms));
              if (j >= nv) p0 = 0.f;
              if (j + 1 >= nv) p1 = 0.f;
              sum += p0 + p1;
              pk[j >> 1] = pack_bf16x2(p0, p1);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = 0u;
        }
        tmem_st_32x32b_x16(lane_addr + SF_COL_P + 16 * c, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      xsum[half * 128 + i] = sum;
      asm volatile("bar.sync %0, 64;\n" ::"r"(1 + quarter) : "memory");
      sum += xsum[(half ^ 1) * 128 + i];
      if (half == 0 && rv.row_valid) a.lse2[((long long)rv.seq * a.H + g.head) * a.S + rv.r_local] = ms + log2f(sum);
      const float inv = rv.row_valid ? 1.0f / sum : 0.f;
      mbar_wait(o_ready, it & 1);
      tc_fence_after();
      {
        uint32_t r[32];
        tmem_ld_32x32b_x32(lane_addr + SF_COL_O + 32 * half, r);
        tmem_ld_wait();
        if (rv.row_valid) {
          bf16* orow = a.out + rv.grow * a.ld_o + g.head * HD + 32 * half;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(r[j + 2]) * inv, __uint_as_float(r[j + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(r[j + 4]) * inv, __uint_as_float(r[j + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(r[j + 6]) * inv, __uint_as_float(r[j + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + j) = o;
          }
        }
      }
      tc_fence_before();   // the O reads precede (through p_ready of the next tile) the next P V product
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, SF_TMEM_COLS);
  }
}

// =====================================================================================================
// Backward
// =====================================================================================================
constexpr int SB_SOFTMAX_WARPS = 8;
constexpr int SB_THREADS = 32 * (2 + SB_SOFTMAX_WARPS);   // warp 0 loader, warp 1 MMA issuer, warps 2-9 softmax / epilogue
constexpr int SB_STAGE_BYTES = 4 * MAT_BYTES;             // Q | K | V | dO
constexpr int SB_TILE_BYTES = 128 * 128 * 2;              // P / dS: [128 queries][128 keys] bf16, two 64-key atoms
constexpr int SB_COL_S = 0, SB_COL_DP = 128, SB_COL_DQ = 256, SB_COL_DK = 320, SB_COL_DV = 384, SB_TMEM_COLS = 512;
constexpr int SB_SMEM = 2 * SB_STAGE_BYTES + 2 * SB_TILE_BYTES + 2 * 2 * 128 * 4 /* delta exchange */ + 256 + 1024;

// byte offset of the 16-byte chunk holding keys [8 kc, 8 kc + 8) of query row i in a P / dS tile
__device__ __forceinline__ uint32_t pds_off(int i, int kc) {
  return (uint32_t)((kc >> 3) * (128 * 128) + i * 128 + (((kc & 7) ^ (i & 7)) << 4));
}

__global__ void __launch_bounds__(SB_THREADS, 1)
attn_small_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                      const SmallArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* p_tile = smem + 2 * SB_STAGE_BYTES;
  uint8_t* ds_tile = p_tile + SB_TILE_BYTES;
  float* xch = reinterpret_cast<float*>(ds_tile + SB_TILE_BYTES);   // [2 buffers][2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 2 * 2 * 128);
  uint64_t* full = bars;           // [2]
  uint64_t* empty = bars + 2;      // [2]
  uint64_t* sdp_ready = bars + 4;
  uint64_t* pds_ready = bars + 5;
  uint64_t* grad_ready = bars + 6; // [3]: dQ, dK, dV
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int groups = a.n_tiles / a.H;
  int t0, t1;
  tile_range(a, t0, t1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(sdp_ready, 1);
    mbar_init(pds_ready, SB_SOFTMAX_WARPS);
    for (int m = 0; m < 3; ++m) mbar_init(&grad_ready[m], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, SB_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- loader
    int it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
      const int st = it & 1;
      mbar_wait_backoff(&empty[st], ((it >> 1) & 1) ^ 1, 64);
      if (elect_one_sync()) {
        const TileGeom g = tile_geom(a, t / groups, t % groups);
        uint8_t* base = smem + st * SB_STAGE_BYTES;
        mbar_arrive_expect_tx(&full[st], SB_STAGE_BYTES);
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          tma_load_2d(base + m * MAT_BYTES, &tm_qkv, &full[st], m * a.D + g.head * HD, g.row0_0);
          tma_load_2d(base + m * MAT_BYTES + SLOT_BYTES, &tm_qkv, &full[st], m * a.D + g.head * HD, g.row0_1);
        }
        tma_load_2d(base + 3 * MAT_BYTES, &tm_do, &full[st], g.head * HD, g.row0_0);
        tma_load_2d(base + 3 * MAT_BYTES + SLOT_BYTES, &tm_do, &full[st], g.head * HD, g.row0_1);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_kk = idesc_bf16(128, 128, 0, 0);   // S = Q K^T, dP = dO V^T
    constexpr uint32_t idesc_mm = idesc_bf16(128, HD, 1, 1);    // dV = P^T dO, dK = dS^T Q  (A and B MN-major)
    constexpr uint32_t idesc_km = idesc_bf16(128, HD, 0, 1);    // dQ = dS K                 (A K-major, B MN-major)
    const uint32_t sp = smem_u32(p_tile), sds = smem_u32(ds_tile);
    int it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
      const int st = it & 1;
      const uint32_t sb = smem_u32(smem + st * SB_STAGE_BYTES);
      const uint32_t sq = sb, sk = sb + MAT_BYTES, sv = sb + 2 * MAT_BYTES, sdo = sb + 3 * MAT_BYTES;
      mbar_wait(&full[st], (it >> 1) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t dq = desc_sw128(sq, 0, 1024), dk = desc_sw128(sk, 0, 1024);
        const uint64_t dv = desc_sw128(sv, 0, 1024), ddo = desc_sw128(sdo, 0, 1024);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16_ss(tmem + SB_COL_S, dq + 2 * k, dk + 2 * k, idesc_kk, k > 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16_ss(tmem + SB_COL_DP, ddo + 2 * k, dv + 2 * k, idesc_kk, k > 0);
        umma_commit(sdp_ready);
      }
      __syncwarp();
      mbar_wait(pds_ready, it & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        // MN-major tiles: 64-element atoms along M/N are 128 x 128 B apart (LBO), 8-row reduction groups 1024 B (SBO);
        // one k-step = 16 reduction rows = 2048 B.  Each gradient signals its own barrier, so the stores of dQ run
        // under the dK / dV products.
        const uint64_t ap = desc_sw128(sp, 128 * 128, 1024), ads = desc_sw128(sds, 128 * 128, 1024);
        const uint64_t bdo = desc_sw128(sdo, 128 * 128, 1024), bq = desc_sw128(sq, 128 * 128, 1024);
        const uint64_t bk = desc_sw128(sk, 128 * 128, 1024);
        // dQ = dS K: the dS tile read K-major (rows = queries, 16 keys = 32 B per k-step, second 64-key atom at +16 KB)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t adq = desc_sw128(sds + (uint32_t)((k >> 2) * (128 * 128) + (k & 3) * 32), 0, 1024);
          umma_bf16_ss(tmem + SB_COL_DQ, adq, bk + (uint64_t)(k * 128), idesc_km, k > 0);
        }
        umma_commit(&grad_ready[0]);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16_ss(tmem + SB_COL_DK, ads + (uint64_t)(k * 128), bq + (uint64_t)(k * 128), idesc_mm, k > 0);
        umma_commit(&grad_ready[1]);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16_ss(tmem + SB_COL_DV, ap + (uint64_t)(k * 128), bdo + (uint64_t)(k * 128), idesc_mm, k > 0);
        umma_commit(&grad_ready[2]);
        umma_commit(&empty[st]);
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue
    // thread = (row i of the tile, half): the two warps of a TMEM lane quarter take the 32-column chunks {half, half+2}
    // of the key columns and, in the epilogue, 32 of the 64 gradient columns
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int i = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    // qkv-bias gradient: this thread's column (32 half + lane) of dQ / dK / dV summed over the warp's rows, kept in
    // registers while the CTA stays on a head and added to global memory when the head changes
    float bsum[3] = {0.f, 0.f, 0.f};
    int bhead = -1;
    auto flush_bias = [&]() {
      if (a.dbias != nullptr && bhead >= 0) {
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          atomicAdd(a.dbias + m * a.D + bhead * HD + 32 * half + lane, bsum[m]);
          bsum[m] = 0.f;
        }
      }
    };
    int it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
      const TileGeom g = tile_geom(a, t / groups, t % groups);
      if (g.head != bhead) {
        flush_bias();
        bhead = g.head;
      }
      const RowView rv = row_view(a, g, i);
      const int c_lo = __shfl_sync(0xffffffffu, rv.c_lo, 0), c_hi = __shfl_sync(0xffffffffu, rv.c_hi, 0);   // uniform
      // a query row outside its sequence attends to nothing: lse = +inf makes every probability exactly 0
      const float lse = rv.row_valid ? a.lse2[((long long)rv.seq * a.H + g.head) * a.S + rv.r_local] : INFINITY;
      mbar_wait(sdp_ready, it & 1);
      tc_fence_after();
      // pass 1: P (kept packed in registers, written to the P tile), dP kept in registers, partial delta
      uint32_t pk[2][16], rd[2][32];
      float dsum = 0.f;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half + 2 * cc;   // 32-column chunk of the tile
        const int nv = chunk_valid(c, c_lo, c_hi);
        if (nv > 0) {
          uint32_t rs[32];
          tmem_ld_32x32b_x32(lane_addr + SB_COL_S + 32 * c, rs);
          tmem_ld_32x32b_x32(lane_addr + SB_COL_DP + 32 * c, rd[cc]);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float p0 = exp2f(fmaf(__uint_as_float(rs[j]), a.scale_log2, -lse));
            float p1 = exp2f(fmaf(__uint_as_float(rs[j + 1]), a.scale_log2, -lse));
            if (nv < 32) {   // warp-uniform: only the chunk that holds the end of the sequence masks per column
              if (j >= nv) p0 = 0.f;
              if (j + 1 >= nv) p1 = 0.f;
            }
            dsum = fmaf(p0, __uint_as_float(rd[cc][j]), dsum);
            dsum = fmaf(p1, __uint_as_float(rd[cc][j + 1]), dsum);
            pk[cc][j >> 1] = pack_bf16x2(p0, p1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[cc][j] = 0u;
#pragma unroll
          for (int j = 0; j < 32; ++j) rd[cc][j] = 0u;
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          *reinterpret_cast<uint4*>(p_tile + pds_off(i, 4 * c + q4)) =
              make_uint4(pk[cc][4 * q4], pk[cc][4 * q4 + 1], pk[cc][4 * q4 + 2], pk[cc][4 * q4 + 3]);
      }
      // delta = rowsum(P o dP) over both halves
      float* xb = xch + (it & 1) * 256;
      xb[half * 128 + i] = dsum;
      asm volatile("bar.sync %0, 64;\n" ::"r"(1 + quarter) : "memory");
      const float delta = dsum + xb[(half ^ 1) * 128 + i];
      // pass 2: dS = P o (dP - delta) * scale
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half + 2 * cc;
        uint32_t dk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 p = unpack_bf16x2(pk[cc][j >> 1]);
          const float d0 = p.x * (__uint_as_float(rd[cc][j]) - delta) * a.scale;
          const float d1 = p.y * (__uint_as_float(rd[cc][j + 1]) - delta) * a.scale;
          dk[j >> 1] = pack_bf16x2(d0, d1);
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          *reinterpret_cast<uint4*>(ds_tile + pds_off(i, 4 * c + q4)) =
              make_uint4(dk[4 * q4], dk[4 * q4 + 1], dk[4 * q4 + 2], dk[4 * q4 + 3]);
      }
      fence_async_smem();      // the tiles are read by the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_ready);

      // epilogue: dQ (lanes = queries), dK / dV (lanes = keys); this thread takes 32 of the 64 columns
      bf16* drow = a.dqkv + rv.grow * a.ld_qkv + g.head * HD + 32 * half;
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const int col = (m == 0) ? SB_COL_DQ : (m == 1 ? SB_COL_DK : SB_COL_DV);
        mbar_wait(&grad_ready[m], it & 1);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld_32x32b_x32(lane_addr + col + 32 * half, r);
        tmem_ld_wait();
        if (rv.row_valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(r[j]), __uint_as_float(r[j + 1]));
            o.y = pack_bf16x2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            o.z = pack_bf16x2(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5]));
            o.w = pack_bf16x2(__uint_as_float(r[j + 6]), __uint_as_float(r[j + 7]));
            *reinterpret_cast<uint4*>(drow + m * a.D + j) = o;
          }
        }
        if (a.dbias != nullptr) {   // column sums over the warp's 32 rows (rows outside a sequence hold exact zeros)
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          warp_colsum<32>(v, lane);
          bsum[m] += v[0];
        }
      }
      tc_fence_before();
    }
    flush_bias();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, SB_TMEM_COLS);
  }
}

bool small_enabled() {
  static const bool on = [] {
    const char* e = getenv("AVS_ATTN_SMALL");
    return !(e && e[0] == '0');
  }();
  return on;
}

int fill_args(SmallArgs& a, long long ld_qkv, long long ld_o, int n_seq, int S, int H) {
  a.n_seq = n_seq; a.S = S; a.H = H; a.D = H * HD;
  a.pair = (S <= SLOT_ROWS) ? 1 : 0;
  const int groups = a.pair ? (n_seq + 1) / 2 : n_seq;
  a.n_tiles = groups * H;
  a.ld_o = ld_o; a.ld_qkv = ld_qkv;
  a.scale = 0.125f;   // 64^-0.5
  a.scale_log2 = a.scale * 1.4426950408889634f;
  return 0;
}

}  // namespace

// Both return -2 when the shape is not covered (head_dim != 64 or S > 128): the caller falls through to the other CUDA
// kernels of this library.
int avs_attention_small_fwd(const void* qkv, long long ld_qkv, void* out, long long ld_o, float* lse2, int n_seq, int S,
                            int H, int head_dim, void* stream) {
  if (head_dim != HD || S > 128 || !small_enabled()) return -2;
  SmallArgs a = {};
  fill_args(a, ld_qkv, ld_o, n_seq, S, H);
  a.out = (bf16*)out; a.lse2 = lse2;
  CUtensorMap tm;
  if (int rc = avs_make_tmap_2d_bf16(&tm, qkv, (long long)n_seq * S, 3LL * a.D, ld_qkv, HD, SLOT_ROWS, 128)) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SF_SMEM);
    if (e != cudaSuccess) { avs_set_error("attn_small_fwd: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr = true;
  }
  const int grid = a.n_tiles < 2 * avs_num_sms() ? a.n_tiles : 2 * avs_num_sms();
  attn_small_fwd_kernel<<<grid, SF_THREADS, SF_SMEM, (cudaStream_t)stream>>>(tm, a);
  return avs_check_launch("attn_small_fwd_kernel");
}

int avs_attention_small_bwd(const void* qkv, long long ld_qkv, const void* dout, long long ld_o, const float* lse2,
                            void* dqkv, float* dbias, int n_seq, int S, int H, int head_dim, void* stream) {
  if (head_dim != HD || S > 128 || !small_enabled()) return -2;
  SmallArgs a = {};
  fill_args(a, ld_qkv, ld_o, n_seq, S, H);
  a.lse2 = const_cast<float*>(lse2); a.dqkv = (bf16*)dqkv; a.dbias = dbias;
  CUtensorMap tm, tm_do;
  if (int rc = avs_make_tmap_2d_bf16(&tm, qkv, (long long)n_seq * S, 3LL * a.D, ld_qkv, HD, SLOT_ROWS, 128)) return rc;
  if (int rc = avs_make_tmap_2d_bf16(&tm_do, dout, (long long)n_seq * S, a.D, ld_o, HD, SLOT_ROWS, 128)) return rc;
  const int smem = SB_SMEM;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { avs_set_error("attn_small_bwd: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr = true;
  }
  const int grid = a.n_tiles < avs_num_sms() ? a.n_tiles : avs_num_sms();
  attn_small_bwd_kernel<<<grid, SB_THREADS, smem, (cudaStream_t)stream>>>(tm, tm_do, a);
  return avs_check_launch("attn_small_bwd_kernel");
}
