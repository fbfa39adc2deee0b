// Standalone bring-up probe for avs_gemm_bf16 (no torch): checks each operand-major combination and epilogue
// against a CPU double-precision reference, then times the ViT-B shapes.  Build: tools/build_probe.sh
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../include/avsiam_b200.h"
extern "C" void avs_debug_set_desc_variant(int v);

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }
static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

// logical A(m,k), B(n,k) stored per major
static int run_case(const char* name, int M, int N, int K, int am, int bm, int flags, bool bias, bool resid,
                    int split_k) {
  std::vector<float> A((size_t)M * K), B((size_t)N * K), bias_h(N), R((size_t)M * N), AUX((size_t)M * N);
  for (auto& x : A) x = bf(frand());
  for (auto& x : B) x = bf(frand());
  for (auto& x : bias_h) x = frand();
  for (auto& x : R) x = bf(frand());
  for (auto& x : AUX) x = bf(frand() * 2);
  std::vector<__nv_bfloat16> Ad((size_t)M * K), Bd((size_t)N * K), Rd((size_t)M * N), AUXd((size_t)M * N);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      size_t idx = am == 0 ? (size_t)m * K + k : (size_t)k * M + m;
      Ad[idx] = __float2bfloat16(A[(size_t)m * K + k]);
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      size_t idx = bm == 0 ? (size_t)n * K + k : (size_t)k * N + n;
      Bd[idx] = __float2bfloat16(B[(size_t)n * K + k]);
    }
  for (size_t i = 0; i < R.size(); ++i) Rd[i] = __float2bfloat16(R[i]), AUXd[i] = __float2bfloat16(AUX[i]);
  const bool f32 = flags & (AVS_EPI_OUT_F32 | AVS_EPI_OUT_ATOMIC);
  void *dA, *dB, *dC, *dR, *dAUX, *dPre;
  float* dbias;
  CK(cudaMalloc(&dA, Ad.size() * 2));
  CK(cudaMalloc(&dB, Bd.size() * 2));
  CK(cudaMalloc(&dC, (size_t)M * N * 4));
  CK(cudaMalloc(&dR, Rd.size() * 2));
  CK(cudaMalloc(&dAUX, Rd.size() * 2));
  CK(cudaMalloc(&dPre, Rd.size() * 2));
  CK(cudaMalloc(&dbias, N * 4));
  CK(cudaMemcpy(dA, Ad.data(), Ad.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bd.data(), Bd.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dR, Rd.data(), Rd.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAUX, AUXd.data(), Rd.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, bias_h.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0, (size_t)M * N * 4));
  avs_gemm_epilogue_t ep = {};
  ep.flags = flags;
  ep.alpha = 1.0f;
  ep.bias = bias ? dbias : nullptr;
  ep.resid = resid ? dR : nullptr;
  ep.ld_resid = N;
  ep.aux_in = dAUX;
  ep.aux_out = (flags & AVS_EPI_GELU) ? dPre : nullptr;
  ep.ld_aux = N;
  int rc = avs_gemm_bf16(dA, am == 0 ? K : M, am, dB, bm == 0 ? K : N, bm, dC, N, M, N, K, &ep, split_k, nullptr);
  if (rc != 0) {
    printf("[%s] launch rc=%d err=%s\n", name, rc, avs_last_error());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[%s] sync error %s\n", name, cudaGetErrorString(e));
    exit(3);
  }
  std::vector<float> C((size_t)M * N);
  if (f32) {
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  } else {
    std::vector<__nv_bfloat16> Cb((size_t)M * N);
    CK(cudaMemcpy(Cb.data(), dC, Cb.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < C.size(); ++i) C[i] = __bfloat162float(Cb[i]);
  }
  double max_err = 0, max_ref = 0;
  int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
      if (bias) acc += bias_h[n];
      if (flags & AVS_EPI_GELU) acc = 0.5 * acc * (1.0 + erf(acc * 0.7071067811865476));
      if (flags & AVS_EPI_DGELU) {
        double x = AUX[(size_t)m * N + n];
        acc *= 0.5 * (1.0 + erf(x * 0.7071067811865476)) + x * 0.3989422804014327 * exp(-0.5 * x * x);
      }
      if (resid) acc += R[(size_t)m * N + n];
      double err = fabs(acc - C[(size_t)m * N + n]);
      double tol = 0.02 * fabs(acc) + 0.05;
      if (err > tol && bad++ < 5) printf("   mismatch m=%d n=%d ref=%f got=%f\n", m, n, acc, C[(size_t)m * N + n]);
      if (err > max_err) max_err = err;
      if (fabs(acc) > max_ref) max_ref = fabs(acc);
    }
  printf("[%s] M=%d N=%d K=%d am=%d bm=%d flags=%d split=%d : max_err=%.4f (max_ref %.2f) bad=%d %s\n", name, M, N, K,
         am, bm, flags, split_k, max_err, max_ref, bad, bad ? "FAIL" : "ok");
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dR); cudaFree(dAUX); cudaFree(dPre); cudaFree(dbias);
  return bad ? 1 : 0;
}

static void time_case(const char* name, int M, int N, int K, int am, int bm, int flags, int split_k) {
  void *dA, *dB, *dC;
  CK(cudaMalloc(&dA, (size_t)M * K * 2));
  CK(cudaMalloc(&dB, (size_t)N * K * 2));
  CK(cudaMalloc(&dC, (size_t)M * N * 4));
  CK(cudaMemset(dA, 0x3c, (size_t)M * K * 2));
  CK(cudaMemset(dB, 0x3c, (size_t)N * K * 2));
  CK(cudaMemset(dC, 0, (size_t)M * N * 4));
  avs_gemm_epilogue_t ep = {};
  ep.flags = flags;
  ep.alpha = 1.0f;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i)
    avs_gemm_bf16(dA, am == 0 ? K : M, am, dB, bm == 0 ? K : N, bm, dC, N, M, N, K, &ep, split_k, nullptr);
  cudaEventRecord(e0);
  const int iters = 20;
  for (int i = 0; i < iters; ++i)
    avs_gemm_bf16(dA, am == 0 ? K : M, am, dB, bm == 0 ? K : N, bm, dC, N, M, N, K, &ep, split_k, nullptr);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[%s] sync error %s\n", name, cudaGetErrorString(e));
    exit(3);
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= iters;
  printf("[time %s] M=%d N=%d K=%d am=%d bm=%d split=%d: %.3f ms  %.1f TFLOP/s\n", name, M, N, K, am, bm, split_k, ms,
         2.0 * M * N * K / ms / 1e9);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
}

int main(int argc, char** argv) {
  int fails = 0;
  printf("avs_version %d\n", avs_version());
  // forward (K,K)
  fails += run_case("fwd-small", 128, 256, 64, 0, 0, 0, false, false, 1);
  fails += run_case("fwd-k256", 128, 256, 256, 0, 0, 0, false, false, 1);
  fails += run_case("fwd-ragged", 300, 512, 192, 0, 0, 0, true, false, 1);
  fails += run_case("fwd-multi", 1000, 768, 768, 0, 0, 0, true, true, 1);
  fails += run_case("fwd-gelu", 300, 512, 192, 0, 0, AVS_EPI_GELU, true, false, 1);
  fails += run_case("fwd-n128", 300, 128, 200, 0, 0, 0, true, false, 1);
  fails += run_case("fwd-n64-f32", 300, 64, 128, 0, 0, AVS_EPI_OUT_F32, true, false, 1);
  // dgrad (K,MN)
  for (int v = 0; v < 2; ++v) {
    avs_debug_set_desc_variant(v);
    printf("--- MN-major descriptor variant %d ---\n", v);
    int f = 0;
    f += run_case("dgrad-small", 128, 256, 64, 0, 1, 0, false, false, 1);
    f += run_case("dgrad-ragged", 300, 512, 192, 0, 1, AVS_EPI_DGELU, false, false, 1);
    f += run_case("wgrad-small", 128, 256, 64, 1, 1, AVS_EPI_OUT_F32, false, false, 1);
    f += run_case("wgrad-ragged", 256, 768, 354, 1, 1, AVS_EPI_OUT_ATOMIC, false, false, 3);
    f += run_case("wgrad-n128", 384, 128, 500, 1, 1, AVS_EPI_OUT_ATOMIC, false, false, 0);
    printf("--- variant %d: %s ---\n", v, f ? "FAIL" : "ok");
    if (v == 0) fails += f;
    if (f == 0) break;
  }
  // timings at ViT-B B=256 shapes
  time_case("qkv", 45312, 2304, 768, 0, 0, 0, 1);
  time_case("fc1", 45312, 3072, 768, 0, 0, 0, 1);
  time_case("fc2", 45312, 768, 3072, 0, 0, 0, 1);
  time_case("dec-fc1", 181248, 2048, 512, 0, 0, 0, 1);
  time_case("dgrad-fc1", 45312, 768, 3072, 0, 1, 0, 1);
  time_case("wgrad-fc1", 3072, 768, 45312, 1, 1, AVS_EPI_OUT_ATOMIC, 0);
  time_case("wgrad-qkv", 2304, 768, 45312, 1, 1, AVS_EPI_OUT_ATOMIC, 0);
  time_case("square8k", 8192, 8192, 8192, 0, 0, 0, 1);
  printf("PROBE %s (fails=%d)\n", fails ? "FAILED" : "PASSED", fails);
  return fails ? 1 : 0;
}
