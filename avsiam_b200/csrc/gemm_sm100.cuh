// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA -> smem ring -> tcgen05.mma (TMEM accumulators,
// double-buffered) -> tcgen05.ld epilogue with fused bias / GELU / dGELU / row-table add / residual.
//
//   D[M,N] = epilogue( sum_k A(m,k) * B(n,k) )
//
// Operand "major" selects how the operand sits in global memory (row-major tensors in all cases):
//   A  MAJOR_K : A is [M, K]  (reduction contiguous)       MAJOR_MN : A is [K, M]  (M contiguous)
//   B  MAJOR_K : B is [N, K]                                MAJOR_MN : B is [K, N]
// Forward Linear = (K,K) with B = weight[N,K]; dgrad = (K,MN) with B = weight[N_out,K_in] read as [K_red, N];
// wgrad = (MN,MN) with A = dY[tokens, N_out], B = X[tokens, K_in] -> no transposes anywhere in backward.
//
// Replaces the cuBLAS calls behind nn.Linear / Conv2d(k=s=16) in the reference
// (src/models/cav_mae_base.py:51,55,96,311,334-335 and timm Mlp fc1/fc2).
#pragma once
#include "common.cuh"

namespace avs {

enum { MAJOR_K = 0, MAJOR_MN = 1 };

// epilogue flags
enum {
  EPI_GELU = 1,        // v = gelu(v)            (aux_out, if set, receives the pre-activation)
  EPI_DGELU = 2,       // v = v * gelu'(aux_in)
  EPI_OUT_F32 = 4,     // C is fp32 (else bf16)
  EPI_OUT_ATOMIC = 8,  // C is fp32 and accumulated with red.global.add (split-K / grad accumulation)
  EPI_AUX_GRAD = 16,   // with EPI_GELU: aux_out receives gelu'(pre) instead of pre
  EPI_MUL_AUX = 32,    // v = v * aux_in  (aux_in holds the stored gelu'; the dGELU epilogue without the derivative math)
};

// Epilogue class = template parameter of the kernel: every instantiation carries only the code of its own epilogue.
// With all variants behind run-time flags in one kernel (6.7 k SASS instructions, 107 KB) the shapes whose epilogue is the
// bound ran 8 % slower than with the 5.7 k-instruction kernel that lacked two of the variants — the epilogue loop no
// longer fitted the instruction cache next to the producer / issuer code.
enum {
  GEMM_E_PLAIN = 0,   // bf16 out: bias, row table, alpha
  GEMM_E_GELU = 1,    // bf16 out: bias + GELU, optional second output (pre-activation or gelu')
  GEMM_E_RESID = 2,   // bf16 out: bias + residual input stream
  GEMM_E_MUL = 3,     // bf16 out: product with an input stream (stored gelu', or gelu'(pre) evaluated here), column sums
  GEMM_E_F32 = 4,     // fp32 out (overwrite or red.global.add): wgrad / split-K
};

struct GemmEpilogue {
  int flags;
  float alpha;          // final scale (before the residual add)
  const float* bias;    // [N] fp32 or null
  const bf16* resid;    // [M, ld_resid] or null
  long long ld_resid;
  const bf16* aux_in;   // [M, ld_aux]
  bf16* aux_out;        // [M, ld_aux]
  long long ld_aux;
  const float* rowadd;  // [rowadd_rows, N] fp32 table or null  (positional embedding)
  const int* rowidx;    // [M] int32 row index into rowadd, or null => m % rowadd_rows
  int rowadd_rows;
  float* colsum;        // [N] fp32 or null: += column sums of the output (bf16 / TMA epilogue only)
};

struct GemmArgs {
  int M, N, K;          // K = reduction length
  void* C;
  long long ldc;
  int split_k;          // >=1; >1 requires EPI_OUT_ATOMIC
  int kb_per_split;     // k-blocks (of 64) per split
  int desc_variant;     // -DAVS_GEMM_DEBUG builds only: 1 swaps LBO/SBO of MN-major descriptors (probe), 2 skips the TMA
                        // stores, 4 skips the in-stream loads (timing experiments, wrong results); ignored otherwise
  int stages;           // smem ring depth (runtime: whatever fits beside the epilogue staging buffers)
  int tma_epi;          // 1: bf16 C (and aux_out) leave through TMA stores, resid/aux_in arrive through TMA loads
  int has_in;           // tma_epi: a [M,N] bf16 input tile stream exists (resid or aux_in — never both)
  int has_aux_out;      // tma_epi: the GELU pre-activation is stored as a second output stream
  GemmEpilogue epi;
};

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_EPI_WARPS = 8;  // two warps per TMEM lane quarter, each taking half of the tile's columns
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;  // warp0 TMA, warp1 MMA, warps2-9 epilogue
constexpr int GEMM_MAX_STAGES = 8;
constexpr int GEMM_EPI_CHUNK = 32;                 // columns per epilogue chunk (one tcgen05.ld 32x32b.x32)
constexpr int GEMM_EPI_BUF = 32 * GEMM_EPI_CHUNK * 2;  // one [32 rows x 32 cols] bf16 input staging tile = 2 KB
#ifndef AVS_GEMM_IN_DEPTH
#define AVS_GEMM_IN_DEPTH 2
#endif
#ifndef AVS_GEMM_TMEM_PF
#define AVS_GEMM_TMEM_PF 0
#endif
#ifndef AVS_GEMM_L2_PF
#define AVS_GEMM_L2_PF 0
#endif
constexpr int GEMM_L2_PF = AVS_GEMM_L2_PF;         // k-blocks of L2 prefetch ahead of the A-operand loads (0 = off)
constexpr int GEMM_IN_DEPTH = AVS_GEMM_IN_DEPTH;   // input tiles in flight per epilogue warp (ring; chunk q uses slot q % depth)
constexpr bool GEMM_TMEM_PF = AVS_GEMM_TMEM_PF != 0;   // tcgen05.ld of chunk c+1 issued before chunk c's math
static_assert(GEMM_IN_DEPTH >= 2 && GEMM_IN_DEPTH <= 4, "input-tile ring depth");
// Output staging tiles are [32 rows x 64 cols] (128-byte rows, SWIZZLE_128B) and leave every SECOND chunk: the TMA
// unit turns each box row into one L2 write request, so 64-byte rows (the 32-column tiles of v2) made the stores —
// 2048 row requests per 128x256 tile — the bound of every bf16-output GEMM (ncu: MMA warp polling tmem_empty).
constexpr int GEMM_OUT_BUF = 32 * 64 * 2;
constexpr int GEMM_SMEM_LIMIT = 226 * 1024;   // dynamic shared memory per CTA
constexpr int GEMM_COLSUM_BYTES = 2 * 4 * 256 * 4;   // [tile parity][lane quarter][column] fp32, behind the barriers

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int BAR_BYTES = 512;
  // per epilogue warp: [in x2][out][aux_out] staging tiles (only the ones the launch uses).  (Double-buffering the
  // output tiles was measured and bought nothing: the mainloop, not the store latency, bounds these kernels.)
  static __host__ __device__ int epi_bytes_per_warp(int tma_epi, int has_in, int has_aux_out) {
    return tma_epi ? GEMM_EPI_BUF * (has_in ? GEMM_IN_DEPTH : 0) + GEMM_OUT_BUF * (1 + (has_aux_out ? 1 : 0)) : 0;
  }
  static __host__ int pick_stages(int epi_per_warp, int extra = 0) {
    int s = (GEMM_SMEM_LIMIT - 1024 - BAR_BYTES - extra - GEMM_EPI_WARPS * epi_per_warp) / STAGE_BYTES;
    return s > GEMM_MAX_STAGES ? GEMM_MAX_STAGES : s;
  }
  static __host__ int smem_bytes(int stages, int epi_per_warp, int extra = 0) {
    return stages * STAGE_BYTES + GEMM_EPI_WARPS * epi_per_warp + 1024 + BAR_BYTES + extra;
  }
};

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, int a_major, int b_major) {
  return (1u << 4)                      // C format F32
         | (1u << 7)                    // A format BF16
         | (1u << 10)                   // B format BF16
         | ((uint32_t)a_major << 15)    // A major (0 = K, 1 = MN)
         | ((uint32_t)b_major << 16)    // B major
         | ((uint32_t)(n >> 3) << 17)   // N / 8
         | ((uint32_t)(m >> 4) << 24);  // M / 16
}

__device__ __forceinline__ void red_add_f32x4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// byte offset of 16-byte chunk `j` (0..3) of row `r` inside a [32 x 32] bf16 staging tile laid out the way TMA's
// SWIZZLE_64B expects it (chunk index XOR address bits 7..8) — also conflict-free for one-row-per-lane accesses.
__device__ __forceinline__ uint32_t epi_tile_off(int r, int j) { return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }

// byte offset of 16-byte chunk `j` (0..7) of row `r` inside a [32 x 64] bf16 output staging tile, SWIZZLE_128B
__device__ __forceinline__ uint32_t out_tile_off(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

template <int A_MAJOR, int B_MAJOR, int BLOCK_N, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_in,
                 const __grid_constant__ CUtensorMap tma_aux, const GemmArgs args) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr bool TMA_EPI = EPI != GEMM_E_F32;                        // bf16 outputs leave through TMA stores
  constexpr bool HAS_IN = EPI == GEMM_E_RESID || EPI == GEMM_E_MUL;  // a bf16 [M, N] input tile stream exists
  const int STAGES = args.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  const int epi_per_warp = Cfg::epi_bytes_per_warp(args.tma_epi, args.has_in, args.has_aux_out);
  uint8_t* smem_epi = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + GEMM_EPI_WARPS * epi_per_warp);
  uint64_t* full_bar = bars;                               // [MAX_STAGES]
  uint64_t* empty_bar = bars + GEMM_MAX_STAGES;            // [MAX_STAGES]
  uint64_t* tfull_bar = bars + 2 * GEMM_MAX_STAGES;        // [2]
  uint64_t* tempty_bar = bars + 2 * GEMM_MAX_STAGES + 2;   // [2]
  uint64_t* in_bar = bars + 2 * GEMM_MAX_STAGES + 4;       // [EPI_WARPS][4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GEMM_MAX_STAGES + 4 + 4 * GEMM_EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (args.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int n_tiles = (args.N + BLOCK_N - 1) / BLOCK_N;
  const int total_kb = (args.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  // Work item t = (k-split ks, output tile mn) with mn FASTEST: the ~148 CTAs in flight then cover all output tiles of a
  // few k-ranges, so every A / B slice is fetched from HBM once and shared through L2 (with ks fastest each CTA streamed
  // private k-ranges and the long-reduction wgrad GEMMs re-read their operands once per output tile), and concurrent
  // red.global.add traffic goes to different output tiles instead of piling onto one.
  const int mn_tiles = m_tiles * n_tiles;
  const int num_tiles = mn_tiles * args.split_k;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (TMA_EPI) {
      tma_prefetch_desc(&tma_c);
      if (HAS_IN) tma_prefetch_desc(&tma_in);
      if (EPI == GEMM_E_GELU && args.has_aux_out) tma_prefetch_desc(&tma_aux);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], GEMM_EPI_WARPS);
    }
    for (int i = 0; i < 4 * GEMM_EPI_WARPS; ++i) mbar_init(&in_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================ TMA producer ============================
    // The whole warp runs the loop convergently (loop state stays in uniform registers); one elected lane issues.
    int stage = 0;
    uint32_t phase = 0;
    // L2 prefetch cursor (K-major A = the activation stream of forward / dgrad products): runs GEMM_L2_PF k-blocks ahead
    // of the load cursor through this CTA's own (tile, k-block) sequence, so that the ring's TMA loads find their A boxes
    // in L2 instead of waiting for HBM behind the epilogue's stores (the ring holds only 3-4 k-blocks = ~1 us of MMAs)
    constexpr bool L2_PF = GEMM_L2_PF > 0 && A_MAJOR == MAJOR_K;
    int tp = blockIdx.x, kbp = 0, kb1p = 0, m0p = 0;
    auto pf_tile = [&]() {
      if (tp < num_tiles) {
        const int ksp = tp / mn_tiles;
        m0p = ((tp - ksp * mn_tiles) / n_tiles) * GEMM_BLOCK_M;
        kbp = ksp * args.kb_per_split;
        kb1p = min(total_kb, kbp + args.kb_per_split);
      }
    };
    auto pf_step = [&]() {
      if (tp >= num_tiles) return;
      if (elect_one_sync()) tma_prefetch_l2_2d(&tma_a, kbp * GEMM_BLOCK_K, m0p);
      if (++kbp == kb1p) {
        tp += gridDim.x;
        pf_tile();
      }
    };
    if constexpr (L2_PF) {
      pf_tile();
#pragma unroll 1
      for (int i = 0; i < GEMM_L2_PF; ++i) pf_step();
    }
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int ks = t / mn_tiles;   // output tile fastest: the CTAs that run together share a k-range (see mn_tiles)
      const int mn = t - ks * mn_tiles;
      const int m0 = (mn / n_tiles) * GEMM_BLOCK_M;
      const int n0 = (mn % n_tiles) * BLOCK_N;
      const int kb0 = ks * args.kb_per_split;
      const int kb1 = min(total_kb, kb0 + args.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        if constexpr (L2_PF) pf_step();
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
          uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
          const int k0 = kb * GEMM_BLOCK_K;
          if (A_MAJOR == MAJOR_K) {
            tma_load_2d(sa, &tma_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < GEMM_BLOCK_M / 64; ++j)
              tma_load_2d(sa + j * (GEMM_BLOCK_K * 128), &tma_a, &full_bar[stage], m0 + j * 64, k0);
          }
          if (B_MAJOR == MAJOR_K) {
            tma_load_2d(sb, &tma_b, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(sb + j * (GEMM_BLOCK_K * 128), &tma_b, &full_bar[stage], n0 + j * 64, k0);
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
    // ============================ MMA issuer ============================
    // One k-block is only 4 MMAs (~512 tensor cycles at BLOCK_N = 256), so the issuing warp's scalar work must stay
    // well below that: the warp runs convergently, every descriptor is a precomputed 64-bit base plus a 16-byte-unit
    // offset (stage and k-step), and one elected lane issues the 4 MMAs and the commit back to back.
    constexpr uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, BLOCK_N, A_MAJOR, B_MAJOR);
    // K-major: step 16 elements (32 B) inside the 128 B swizzle row; 8-row groups 1024 B apart.
    // MN-major: step 16 reduction rows (2 swizzle atoms = 2048 B); atoms along MN are BLOCK_K*128 B apart (LBO),
    //           8-row K groups 1024 B apart (SBO).
#ifdef AVS_GEMM_DEBUG
    const uint32_t mn_lbo = (args.desc_variant & 1) ? 1024 : GEMM_BLOCK_K * 128;
    const uint32_t mn_sbo = (args.desc_variant & 1) ? GEMM_BLOCK_K * 128 : 1024;
#else
    constexpr uint32_t mn_lbo = GEMM_BLOCK_K * 128, mn_sbo = 1024;
#endif
    const uint64_t da0 = (A_MAJOR == MAJOR_K) ? make_smem_desc(smem_u32(smem_a), 0, 1024)
                                              : make_smem_desc(smem_u32(smem_a), mn_lbo, mn_sbo);
    const uint64_t db0 = (B_MAJOR == MAJOR_K) ? make_smem_desc(smem_u32(smem_b), 0, 1024)
                                              : make_smem_desc(smem_u32(smem_b), mn_lbo, mn_sbo);
    constexpr uint32_t A_K16 = ((A_MAJOR == MAJOR_K) ? 32 : 2048) >> 4, B_K16 = ((B_MAJOR == MAJOR_K) ? 32 : 2048) >> 4;
    constexpr uint32_t A_STAGE16 = Cfg::A_BYTES >> 4, B_STAGE16 = Cfg::B_BYTES >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int ks = t / mn_tiles;
      const int kb0 = ks * args.kb_per_split;
      const int kb1 = min(total_kb, kb0 + args.kb_per_split);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t da = da0 + (uint64_t)((uint32_t)stage * A_STAGE16);
          const uint64_t db = db0 + (uint64_t)((uint32_t)stage * B_STAGE16);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k)
            umma_bf16_ss(tmem_d, da + (uint64_t)(k * A_K16), db + (uint64_t)(k * B_K16), idesc,
                         (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
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
      const int t = blockIdx.x + (q / CH) * (int)gridDim.x;
      if (t >= num_tiles) return;
      const int mn = t % mn_tiles;
      const int m0 = (mn / n_tiles) * GEMM_BLOCK_M;
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
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int ks = t / mn_tiles;   // output tile fastest: the CTAs that run together share a k-range (see mn_tiles)
      const int mn = t - ks * mn_tiles;
      const int m0 = (mn / n_tiles) * GEMM_BLOCK_M;
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
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
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
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace avs
