// Host side of the tcgen05 GEMM: TMA descriptor encoding, tile/split-K selection, C-ABI entry.
#include <cudaTypedefs.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "../../include/avsiam_b200.h"
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "gemm_sm100_2cta.cuh"

using namespace avs;

#ifndef AVS_GEMM_MUL16_DEFAULT
#define AVS_GEMM_MUL16_DEFAULT 1
#endif
#ifndef AVS_GEMM_EW16_DEFAULT
#define AVS_GEMM_EW16_DEFAULT 1
#endif
#ifndef AVS_GEMM_2CTA_DEFAULT
#define AVS_GEMM_2CTA_DEFAULT 1   // AVS_GEMM_2CTA=0 in the environment keeps the one-CTA kernel everywhere (A/B)
#endif

#ifdef AVS_GEMM_DEBUG   // probe / timing builds only: the release library carries no debug switches in its epilogue
static int g_desc_variant = getenv("AVS_GEMM_DEBUG") ? atoi(getenv("AVS_GEMM_DEBUG")) : 0;
extern "C" void avs_debug_set_desc_variant(int v) { g_desc_variant = v; }
#else
static const int g_desc_variant = 0;
#endif

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// TMA descriptor cache: a training step issues the same ~280 GEMMs on the same buffers every iteration (the caching
// allocator hands back the same addresses), so each (pointer, shape, pitch, box, swizzle) is encoded once and copied
// afterwards — 128 bytes instead of a driver call, five times per GEMM. avs_reset() drops the cache (call it after
// freeing / re-laying-out buffers if the process keeps running with different shapes; entries are keyed by value, so
// a stale entry can never describe a different tensor than the one asked for).
struct TmapKey {
  const void* ptr;
  long long rows, cols, ld;
  int box_cols, box_rows, swizzle;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_cols == o.box_cols &&
           box_rows == o.box_rows && swizzle == o.swizzle;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    h ^= (size_t)k.rows * 0xC2B2AE3D27D4EB4Full + (size_t)k.cols * 0x165667B19E3779F9ull + (size_t)k.ld;
    h ^= ((size_t)k.box_cols << 40) ^ ((size_t)k.box_rows << 20) ^ (size_t)k.swizzle;
    return h;
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static std::mutex g_tmap_mutex;
static const size_t TMAP_CACHE_MAX = 1 << 16;

extern "C" void avs_reset(void) {
  std::lock_guard<std::mutex> lock(g_tmap_mutex);
  g_tmap_cache.clear();
}

static int encode_tmap_2d(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld,
                          int box_cols, int box_rows, CUtensorMapSwizzle swizzle);

// 2-D bf16 row-major tensor [rows, cols] with pitch ld (elements); box = {box_cols (inner), box_rows}, 128B swizzle.
static int make_tmap_2d(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld,
                        int box_cols, int box_rows, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  const TmapKey key{ptr, rows, cols, ld, box_cols, box_rows, (int)swizzle};
  {
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *map = it->second;
      return 0;
    }
  }
  if (int rc = encode_tmap_2d(map, ptr, rows, cols, ld, box_cols, box_rows, swizzle)) return rc;
  std::lock_guard<std::mutex> lock(g_tmap_mutex);
  if (g_tmap_cache.size() >= TMAP_CACHE_MAX) g_tmap_cache.clear();
  g_tmap_cache.emplace(key, *map);
  return 0;
}

static int encode_tmap_2d(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld,
                          int box_cols, int box_rows, CUtensorMapSwizzle swizzle) {
  auto fn = get_encode_fn();
  AVS_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  AVS_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA: base pointer must be 16-byte aligned");
  AVS_REQUIRE((ld * 2) % 16 == 0, "TMA: row pitch must be a multiple of 8 bf16 elements (got %lld)", ld);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AVS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld cols=%lld ld=%lld", (int)r,
              rows, cols, ld);
  return 0;
}

struct GemmMaps {
  CUtensorMap a, b, c, in, aux;
};

int avs_make_tmap_2d_bf16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_cols,
                          int box_rows, int swizzle_bytes) {
  return make_tmap_2d(map, ptr, rows, cols, ld, box_cols, box_rows,
                      swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int AM, int BM, int BN, int EPI>
static int launch_gemm(const GemmMaps& tm, GemmArgs& args, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;  // idempotent; benign race
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<AM, BM, BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GEMM_SMEM_LIMIT);
    if (e != cudaSuccess) {
      avs_set_error("cudaFuncSetAttribute(gemm smem=%d): %s", GEMM_SMEM_LIMIT, cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  const int epw = Cfg::epi_bytes_per_warp(args.tma_epi, args.has_in, args.has_aux_out);
  const int extra = (EPI == GEMM_E_MUL) ? GEMM_COLSUM_BYTES : 0;
  args.stages = Cfg::pick_stages(epw, extra);
  const int smem = Cfg::smem_bytes(args.stages, epw, extra);
  gemm_bf16_kernel<AM, BM, BN, EPI><<<grid, GEMM_THREADS, smem, stream>>>(tm.a, tm.b, tm.c, tm.in, tm.aux, args);
  return avs_check_launch("gemm_bf16_kernel");
}

// ---- two-CTA (cta_group::2) kernel: forward / dgrad products with bf16 outputs at N > 128 (gemm_sm100_2cta.cuh) ----
template <int BM, int EPI, int EW = GEMM_EPI_WARPS>
static int launch_gemm_pair(const GemmMaps& tm, GemmArgs& args, int grid, cudaStream_t stream) {
  static bool attr_set = false;  // idempotent; benign race
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_bf16_kernel<BM, EPI, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GEMM_SMEM_LIMIT);
    if (e != cudaSuccess) {
      avs_set_error("cudaFuncSetAttribute(gemm pair smem=%d): %s", GEMM_SMEM_LIMIT, cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  const int staging = Gemm2Cfg::staging_bytes(EW, args.has_in, args.has_aux_out);
  const int extra = (EPI == GEMM_E_MUL) ? GEMM_COLSUM_BYTES : 0;
  args.stages = Gemm2Cfg::pick_stages(staging, extra);
  const int smem = Gemm2Cfg::smem_bytes(args.stages, staging, extra);
  gemm2_bf16_kernel<BM, EPI, EW><<<grid, 64 + 32 * EW, smem, stream>>>(tm.a, tm.b, tm.c, tm.in, tm.aux, args);   // __cluster_dims__(2)
  return avs_check_launch("gemm2_bf16_kernel");
}

template <int BM>
static int launch_gemm_pair_class(int cls, const GemmMaps& tm, GemmArgs& args, int grid, cudaStream_t stream) {
  switch (cls) {
    case GEMM_E_PLAIN: return launch_gemm_pair<BM, GEMM_E_PLAIN>(tm, args, grid, stream);
    case GEMM_E_GELU: {   // fc1: 16 epilogue warps in pairs (AVS_GEMM_EW16=0 keeps 8)
      static const bool ew16 = getenv("AVS_GEMM_EW16") ? atoi(getenv("AVS_GEMM_EW16")) != 0 : AVS_GEMM_EW16_DEFAULT != 0;
      if (ew16) return launch_gemm_pair<BM, GEMM_E_GELU, 16>(tm, args, grid, stream);
      return launch_gemm_pair<BM, GEMM_E_GELU>(tm, args, grid, stream);
    }
    case GEMM_E_RESID: return launch_gemm_pair<BM, GEMM_E_RESID>(tm, args, grid, stream);
    default: {            // fc2 dgrad x derivative: 16 epilogue warps, 3 operand stages (AVS_GEMM_MUL16=0: 8 warps, 4 stages)
      static const bool mul16 = getenv("AVS_GEMM_MUL16") ? atoi(getenv("AVS_GEMM_MUL16")) != 0 : AVS_GEMM_MUL16_DEFAULT != 0;
      if (mul16) return launch_gemm_pair<BM, GEMM_E_MUL, 16>(tm, args, grid, stream);
      return launch_gemm_pair<BM, GEMM_E_MUL>(tm, args, grid, stream);
    }
  }
}

static bool gemm_pair_enabled() {
  static const bool on = [] {
    const char* e = getenv("AVS_GEMM_2CTA");
    return e ? atoi(e) != 0 : AVS_GEMM_2CTA_DEFAULT != 0;
  }();
  return on;
}

template <int AM, int BM, int BN>
static int launch_gemm_class(int cls, const GemmMaps& tm, GemmArgs& args, int grid, cudaStream_t stream) {
  switch (cls) {
    case GEMM_E_PLAIN: return launch_gemm<AM, BM, BN, GEMM_E_PLAIN>(tm, args, grid, stream);
    case GEMM_E_GELU: return launch_gemm<AM, BM, BN, GEMM_E_GELU>(tm, args, grid, stream);
    case GEMM_E_RESID: return launch_gemm<AM, BM, BN, GEMM_E_RESID>(tm, args, grid, stream);
    case GEMM_E_MUL: return launch_gemm<AM, BM, BN, GEMM_E_MUL>(tm, args, grid, stream);
    default: return launch_gemm<AM, BM, BN, GEMM_E_F32>(tm, args, grid, stream);
  }
}

extern "C" int avs_gemm_bf16(const void* A, long long lda, int a_major, const void* B, long long ldb, int b_major,
                             void* C, long long ldc, int M, int N, int K, const avs_gemm_epilogue_t* epi,
                             int split_k, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  AVS_REQUIRE(A && B && C && epi, "avs_gemm_bf16: null pointer");
  AVS_REQUIRE(M > 0 && N > 0 && K > 0, "avs_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  AVS_REQUIRE(N % 8 == 0, "avs_gemm_bf16: N must be a multiple of 8 (got %d)", N);
  AVS_REQUIRE((a_major == 0 || a_major == 1) && (b_major == 0 || b_major == 1), "avs_gemm_bf16: bad major");
  AVS_REQUIRE(!(a_major == 1 && b_major == 0), "avs_gemm_bf16: (A MN-major, B K-major) is not instantiated");
  const bool out_f32 = (epi->flags & (AVS_EPI_OUT_F32 | AVS_EPI_OUT_ATOMIC)) != 0;
  AVS_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0 && (ldc * (out_f32 ? 4 : 2)) % 16 == 0,
              "avs_gemm_bf16: C must be 16-byte aligned with a 16-byte-multiple pitch");
  if (epi->flags & (AVS_EPI_DGELU | AVS_EPI_MUL_AUX))
    AVS_REQUIRE(epi->aux_in != nullptr, "avs_gemm_bf16: DGELU / MUL_AUX need aux_in");
  AVS_REQUIRE(!((epi->flags & AVS_EPI_DGELU) && (epi->flags & AVS_EPI_MUL_AUX)), "avs_gemm_bf16: DGELU and MUL_AUX exclude each other");
  if (epi->flags & AVS_EPI_AUX_GRAD)
    AVS_REQUIRE((epi->flags & AVS_EPI_GELU) && epi->aux_out, "avs_gemm_bf16: AUX_GRAD needs GELU and aux_out");
  if (epi->bias) AVS_REQUIRE((reinterpret_cast<uintptr_t>(epi->bias) & 15) == 0, "avs_gemm_bf16: bias alignment");
  if (epi->rowadd) AVS_REQUIRE(epi->rowadd_rows > 0, "avs_gemm_bf16: rowadd_rows must be > 0");
  // epilogue class (one kernel instantiation each; gemm_sm100.cuh)
  const int cls = out_f32 ? GEMM_E_F32
                  : (epi->flags & AVS_EPI_GELU) ? GEMM_E_GELU
                  : (epi->flags & (AVS_EPI_DGELU | AVS_EPI_MUL_AUX)) ? GEMM_E_MUL
                  : epi->resid ? GEMM_E_RESID : GEMM_E_PLAIN;
  AVS_REQUIRE(!epi->colsum || cls == GEMM_E_MUL, "avs_gemm_bf16: colsum is fused into the DGELU / MUL_AUX epilogues only");
  AVS_REQUIRE(!epi->rowadd || cls == GEMM_E_PLAIN || cls == GEMM_E_F32,
              "avs_gemm_bf16: rowadd combines with bias / alpha only (patch-embed epilogue)");
  AVS_REQUIRE(!(epi->resid && cls == GEMM_E_GELU), "avs_gemm_bf16: GELU and resid cannot be combined");

  const int BN = (N > 128) ? 256 : 128;
  // forward / dgrad products with bf16 outputs and N > 128: a pair of CTAs per 256 x 256 tile (gemm_sm100_2cta.cuh)
  const bool pair = gemm_pair_enabled() && !out_f32 && a_major == 0 && BN == 256 && split_k <= 1 && M > GEMM_BLOCK_M;
  const int m_tiles = ceil_div(M, pair ? 2 * GEMM_BLOCK_M : GEMM_BLOCK_M), n_tiles = ceil_div(N, BN);
  const int total_kb = ceil_div(K, GEMM_BLOCK_K);
  const int sms = avs_num_sms();
  if (split_k <= 0) {  // auto: only meaningful with atomic accumulation
    split_k = 1;
    if (epi->flags & AVS_EPI_OUT_ATOMIC) {
      // wgrad: few output tiles, very long reduction.  Pick the split that wastes the least of the last wave
      // (work items = tiles x splits over `sms` persistent CTAs), keeping >= 8 k-blocks per split.
      const int mn = m_tiles * n_tiles;
      const int max_split = total_kb / 8 > 0 ? (total_kb / 8 < 64 ? total_kb / 8 : 64) : 1;
      double best = -1.0;
      for (int s = 1; s <= max_split; ++s) {
        const int kbs = ceil_div(total_kb, s);
        const int s_eff = ceil_div(total_kb, kbs);
        const long long items = (long long)mn * s_eff;
        const long long rounds = (items + sms - 1) / sms;
        // time ~ rounds x (k-blocks per item + fixed per-item epilogue cost of ~6 k-block equivalents)
        const double t = (double)rounds * (kbs + 6);
        if (best < 0 || t < best * 0.995) { best = t; split_k = s_eff; }
      }
    }
  }
  AVS_REQUIRE(split_k == 1 || (epi->flags & AVS_EPI_OUT_ATOMIC), "avs_gemm_bf16: split_k>1 needs AVS_EPI_OUT_ATOMIC");
  if (split_k > total_kb) split_k = total_kb;
  const int kb_per_split = ceil_div(total_kb, split_k);
  split_k = ceil_div(total_kb, kb_per_split);

  GemmMaps tm;
  memset(&tm, 0, sizeof(tm));
  int rc;
  if (a_major == 0) rc = make_tmap_2d(&tm.a, A, M, K, lda, GEMM_BLOCK_K, GEMM_BLOCK_M);
  else rc = make_tmap_2d(&tm.a, A, K, M, lda, 64, GEMM_BLOCK_K);
  if (rc) return rc;
  if (b_major == 0) rc = make_tmap_2d(&tm.b, B, N, K, ldb, GEMM_BLOCK_K, pair ? BN / 2 : BN);   // pair: each CTA loads half
  else rc = make_tmap_2d(&tm.b, B, K, N, ldb, 64, GEMM_BLOCK_K);
  if (rc) return rc;
  // bf16 outputs leave as [32 x 64] tiles (128-byte swizzle), the residual / dGELU operand arrives as [32 x 32]
  // tiles (64-byte swizzle), all moved by TMA
  const bool tma_epi = !out_f32;
  const void* in_ptr = nullptr;
  long long in_ld = 0;
  if (tma_epi) {
    AVS_REQUIRE(!(epi->resid && (epi->flags & (AVS_EPI_DGELU | AVS_EPI_MUL_AUX))),
                "avs_gemm_bf16: resid and DGELU / MUL_AUX cannot be combined");
    if ((rc = make_tmap_2d(&tm.c, C, M, N, ldc, 2 * GEMM_EPI_CHUNK, 32, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if (epi->flags & (AVS_EPI_DGELU | AVS_EPI_MUL_AUX)) { in_ptr = epi->aux_in; in_ld = epi->ld_aux; }
    else if (epi->resid) { in_ptr = epi->resid; in_ld = epi->ld_resid; }
    if (in_ptr && (rc = make_tmap_2d(&tm.in, in_ptr, M, N, in_ld, GEMM_EPI_CHUNK, 32, CU_TENSOR_MAP_SWIZZLE_64B)))
      return rc;
    if ((epi->flags & AVS_EPI_GELU) && epi->aux_out &&
        (rc = make_tmap_2d(&tm.aux, epi->aux_out, M, N, epi->ld_aux, 2 * GEMM_EPI_CHUNK, 32, CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  } else {
    AVS_REQUIRE(!epi->resid && !(epi->flags & (AVS_EPI_GELU | AVS_EPI_DGELU | AVS_EPI_MUL_AUX)),
                "avs_gemm_bf16: fp32 outputs support bias / rowadd / alpha only");
  }

  GemmArgs args;
  args.M = M; args.N = N; args.K = K;
  args.C = C; args.ldc = ldc;
  args.split_k = split_k; args.kb_per_split = kb_per_split;
  args.desc_variant = g_desc_variant;
  args.stages = 0;
  args.tma_epi = tma_epi ? 1 : 0;
  args.has_in = in_ptr ? 1 : 0;
  args.has_aux_out = (tma_epi && (epi->flags & AVS_EPI_GELU) && epi->aux_out) ? 1 : 0;
  args.epi.flags = epi->flags;
  args.epi.alpha = epi->alpha;
  args.epi.bias = epi->bias;
  args.epi.resid = reinterpret_cast<const bf16*>(epi->resid);
  args.epi.ld_resid = epi->ld_resid;
  args.epi.aux_in = reinterpret_cast<const bf16*>(epi->aux_in);
  args.epi.aux_out = reinterpret_cast<bf16*>(epi->aux_out);
  args.epi.ld_aux = epi->ld_aux;
  args.epi.rowadd = epi->rowadd;
  args.epi.rowidx = epi->rowidx;
  args.epi.rowadd_rows = epi->rowadd_rows;
  args.epi.colsum = epi->colsum;

  const int num_tiles = m_tiles * n_tiles * split_k;
  if (pair) {
    const int clusters = num_tiles < sms / 2 ? num_tiles : sms / 2;
    if (b_major == 0) return launch_gemm_pair_class<MAJOR_K>(cls, tm, args, 2 * clusters, stream);
    return launch_gemm_pair_class<MAJOR_MN>(cls, tm, args, 2 * clusters, stream);
  }
  const int grid = num_tiles < sms ? num_tiles : sms;
  const int key = a_major * 2 + b_major;
  if (BN == 256) {
    if (key == 0) return launch_gemm_class<MAJOR_K, MAJOR_K, 256>(cls, tm, args, grid, stream);
    if (key == 1) return launch_gemm_class<MAJOR_K, MAJOR_MN, 256>(cls, tm, args, grid, stream);
    return launch_gemm_class<MAJOR_MN, MAJOR_MN, 256>(cls, tm, args, grid, stream);
  } else {
    if (key == 0) return launch_gemm_class<MAJOR_K, MAJOR_K, 128>(cls, tm, args, grid, stream);
    if (key == 1) return launch_gemm_class<MAJOR_K, MAJOR_MN, 128>(cls, tm, args, grid, stream);
    return launch_gemm_class<MAJOR_MN, MAJOR_MN, 128>(cls, tm, args, grid, stream);
  }
}
