// Library-wide plumbing: thread-local error string, launch counter, device query.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/avsiam_b200.h"
#include "common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void avs_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int avs_check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    avs_set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int avs_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

extern "C" const char* avs_last_error(void) { return g_err; }
extern "C" int avs_version(void) { return 100; }
extern "C" long long avs_launch_count(void) { return g_launches.load(); }
extern "C" void avs_reset_launch_count(void) { g_launches.store(0); }
