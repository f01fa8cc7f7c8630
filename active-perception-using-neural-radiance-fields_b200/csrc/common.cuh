// Shared helpers for the apnerf sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#define APNERF_API extern "C" __attribute__((visibility("default")))

// Error plumbing of the C-ABI: every entry point returns 0 or a cudaError_t value and leaves a
// message for apnerf_last_error().
void apnerf_set_error(const char* where, cudaError_t e);
void apnerf_set_error_msg(const char* msg);

#define APNERF_CHECK_LAUNCH(name)                         \
  do {                                                    \
    cudaError_t _e = cudaGetLastError();                  \
    if (_e != cudaSuccess) {                              \
      apnerf_set_error(name, _e);                         \
      return (int)_e;                                     \
    }                                                     \
  } while (0)

#define APNERF_CUDA(call)                                 \
  do {                                                    \
    cudaError_t _e = (call);                              \
    if (_e != cudaSuccess) {                              \
      apnerf_set_error(#call, _e);                        \
      return (int)_e;                                     \
    }                                                     \
  } while (0)

#define APNERF_REQUIRE(cond, msg)                         \
  do {                                                    \
    if (!(cond)) {                                        \
      apnerf_set_error_msg(msg);                          \
      return (int)cudaErrorInvalidValue;                  \
    }                                                     \
  } while (0)

// SM count of the CURRENT device (cached per device: a process may drive several GPUs).
static inline int apnerf_num_sms() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cache[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

static inline int ceil_div_i(long long a, long long b) { return (int)((a + b - 1) / b); }

// Grid for a grid-stride kernel: enough CTAs to cover n, capped at a multiple of the SM count.
static inline int grid_for(long long n, int threads, int ctas_per_sm) {
  long long need = (n + threads - 1) / threads;
  long long cap = (long long)apnerf_num_sms() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// Tuning knobs (environment overrides are for experiments; defaults are the measured best).
#include <stdlib.h>
static inline float apnerf_skip_min_steps() {
  static float v = -1.f;
  if (v < 0.f) {
    const char* e = getenv("APNERF_SKIP_MIN");
    v = e ? (float)atof(e) : 256.f;
  }
  return v;
}
static inline void apnerf_march_cfg(int& threads, int& ctas_per_sm) {
  static int t = 0, c = 0;
  if (t == 0) {
    const char* e = getenv("APNERF_MARCH_CFG");
    if (!e || sscanf(e, "%d,%d", &t, &c) != 2) t = 256, c = 4;
  }
  threads = t, ctas_per_sm = c;
}
