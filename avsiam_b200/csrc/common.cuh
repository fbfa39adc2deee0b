// Shared device/host helpers for the avsiam_b200 kernels (sm_100a only).
// PTX wrappers for mbarrier / TMA / tcgen05 / TMEM plus the C-ABI error plumbing.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// C-ABI error state (thread-local; see include/avsiam_b200.h "Errors")
// ---------------------------------------------------------------------------------------------
void avs_set_error(const char* fmt, ...);
int avs_check_launch(const char* what);  // returns 0 or the cudaError_t (>0), recording the message

#define AVS_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      avs_set_error(__VA_ARGS__);       \
      return -1;                        \
    }                                   \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
int avs_num_sms();
// 2-D bf16 row-major tensor map [rows, cols] (pitch ld elements), box {box_cols, box_rows}; swizzle_bytes in {64,128}
int avs_make_tmap_2d_bf16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_cols,
                          int box_rows, int swizzle_bytes);
// tcgen05 attention (attention_tc.cu); each returns -2 when the shape is not covered (the caller then uses the
// mma.sync kernels in attention.cu — both are CUDA paths of this library)
int avs_attention_bwd_tc(const void* qkv, long long ld_qkv, const void* dout, long long ld_o, const float* lse2,
                         const float* delta, void* dqkv, float* dbias, int n_seq, int S, int H, int head_dim,
                         void* stream);
int avs_attention_fwd_tc(const void* qkv, long long ld_qkv, void* out, long long ld_o, float* lse2, int n_seq, int S,
                         int H, int head_dim, void* stream);
// short sequences of the shared encoder (head_dim 64, S <= 128): persistent tcgen05 kernels, attention_small.cu
int avs_attention_small_fwd(const void* qkv, long long ld_qkv, void* out, long long ld_o, float* lse2, int n_seq, int S,
                            int H, int head_dim, void* stream);
int avs_attention_small_bwd(const void* qkv, long long ld_qkv, const void* dout, long long ld_o, const float* lse2,
                            void* dqkv, float* dbias, int n_seq, int S, int H, int head_dim, void* stream);
bool avs_attention_tc_enabled();  // AVS_ATTN_TC=0 in the environment selects the mma.sync kernels (A/B timing)

// ---------------------------------------------------------------------------------------------
// Device PTX helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// For producers that wait a long time for a buffer to drain (persistent kernels): the plain loop above retries every
// ~12 cycles and takes issue slots and mbarrier bandwidth from the compute warps of the same SM sub-partition.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t ns = 256) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}

// ---- TMA ----
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// L2 prefetch of one box of a tiled tensor (a hint: no barrier, no shared-memory destination)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1)
               : "memory");
}
// 2-D tiled load: coordinates are (c0 = innermost/contiguous, c1 = row)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs / fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (M=128 lanes x 16 bf16 per K step = 8 packed 32-bit columns)
// is read from tensor memory — used by attention to feed P / dS straight from the softmax warps.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tcgen05.commit: arrive on an mbarrier when all previously issued MMAs of this thread retire.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base_lane+i).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
// registers -> TMEM: thread i writes lane (base_lane+i), 16 / 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// ---- misc math / vector helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
// exact-erf GELU (timm Mlp uses nn.GELU default, approximate='none'):  gelu(x) = x * Phi(x).
// The GEMM epilogues evaluate it ~10^8 times per step on warps that must keep pace with the tensor pipe.  A pure
// FMA-pipe polynomial costs ~13 issue slots per element and made the fc1 GEMMs epilogue-bound (ncu: 621
// instructions per 32-column chunk), so the normal CDF is evaluated as a sigmoid of a fitted odd polynomial,
//   Phi(x) ~= 1 / (1 + exp2(x * Q(t))),  t = min(x^2, 5.5^2),
// which moves the transcendental part to the MUFU pipe and leaves 6 FMA-pipe instructions; the fit itself has
// |gelu err| <= 2.6e-5, |gelu' err| <= 1.1e-4  (bf16 ulp at 1 is 3.9e-3; the clamp keeps the fit in range, beyond it
// the sigmoid is saturated).  gelu' is the exact derivative of the approximation,
//   gelu'(x) = s + x s (1 - s) R(t),  R = P + 2 t P'.   Coefficients: tools/fit_gelu_poly.py.
// The sigmoid is evaluated with the hardware tanh (sigmoid(2y) = (1 + tanh y) / 2, y = x (c0 + c1 t + c2 t^2),
// c0 = 0.79751 ~ sqrt(2/pi)): ONE MUFU per element instead of ex2 + rcp.  That matters because the GELU / dGELU
// epilogues of the K = 512 / 768 GEMMs are MUFU-bound, not FMA-bound: MUFU issues 16 lanes per clock per SM
// (tools/mufu_probe.cu), so 2 MUFU x 128 x 256 elements = 4096 cycles per tile against a 3072-cycle mainloop.
// tanh.approx.f32 has a relative error of 2^-11 (absolute 4.9e-4 near saturation), i.e. |gelu err| <= 2.5e-4 |x| —
// below the bf16 rounding of the stored activation for x > 0, above it (but < 1.3e-3 absolute) in the negative tail.
__device__ __forceinline__ float tanh_approx(float y) {
  float r;
  asm("tanh.approx.f32 %0, %1;\n" : "=f"(r) : "f"(y));
  return r;
}
__device__ __forceinline__ float gelu_tanh_arg(float x, float t) {
  float q = -3.515168559e-04f;
  q = fmaf(q, t, 3.700564594e-02f);
  q = fmaf(q, t, 7.975078770e-01f);
  return x * q;
}
__device__ __forceinline__ float gelu_sigmoid(float x) {   // Phi(x)
  const float t = fminf(x * x, 30.25f);
  return fmaf(0.5f, tanh_approx(gelu_tanh_arg(x, t)), 0.5f);
}
__device__ __forceinline__ float gelu_erf(float x) {
  const float t = fminf(x * x, 30.25f);
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_approx(gelu_tanh_arg(x, t)), hx);
}
__device__ __forceinline__ float dgelu_erf(float x) {
  const float t = fminf(x * x, 30.25f);
  const float s = fmaf(0.5f, tanh_approx(gelu_tanh_arg(x, t)), 0.5f);
  float r = -3.515168559e-03f;
  r = fmaf(r, t, 2.220338732e-01f);
  r = fmaf(r, t, 1.595015764e+00f);
  return fmaf(x * fmaf(-s, s, s), r, s);
}

// gelu(x) and gelu'(x) together (the forward epilogue that stores the derivative instead of the pre-activation): the
// tanh, the clamp and x^2 are shared, the derivative costs 5 more FMA-pipe instructions.
__device__ __forceinline__ void gelu_and_grad(float x, float& g, float& dg) {
  const float t = fminf(x * x, 30.25f);
  const float s = fmaf(0.5f, tanh_approx(gelu_tanh_arg(x, t)), 0.5f);
  float r = -3.515168559e-03f;
  r = fmaf(r, t, 2.220338732e-01f);
  r = fmaf(r, t, 1.595015764e+00f);
  g = x * s;
  dg = fmaf(x * fmaf(-s, s, s), r, s);
}

// Column sums over the 32 rows a warp holds (one row per lane, N columns per lane): recursive halving, so N = 16
// costs 16 shuffles instead of 80.  On return lane L holds the sum of column  (N == 32 ? L : (L >> 1) & 15)  in v[0].
template <int N>
__device__ __forceinline__ void warp_colsum(float (&v)[N], int lane) {
#pragma unroll
  for (int n = N / 2, o = 16; n >= 1; n >>= 1, o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = up ? v[i] : v[i + n];
      const float keep = up ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  if (N == 16) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// ---- TMA store / bulk-group plumbing (epilogues) ----
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
#endif  // __CUDACC__
