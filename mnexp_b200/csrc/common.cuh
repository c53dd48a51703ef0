// Shared device helpers for the LSTUR sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lstur_b200.h"

namespace lstur {

void set_error(const char* fmt, ...);
extern unsigned long long g_launch_count;   // kernels launched through this library (bench.py: gpu_launches)

#define LSTUR_CHECK_LAUNCH(name)                                              \
  do {                                                                        \
    ++lstur::g_launch_count;                                                  \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) {                                                 \
      lstur::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return LSTUR_ERR_CUDA;                                                  \
    }                                                                         \
  } while (0)

#define LSTUR_REQUIRE(cond, name)                                             \
  do {                                                                        \
    if (!(cond)) {                                                            \
      lstur::set_error("%s: invalid argument: %s", name, #cond);              \
      return LSTUR_ERR_ARG;                                                   \
    }                                                                         \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int warp_or(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Counter-based RNG for dropout: lowbias32 of (index, seed).  The same
// function is replicated in numpy by mnexp_b200/rng.py so that tests can hand
// the exact mask to the oracle.
__device__ __host__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

__device__ __host__ __forceinline__ uint32_t rng_u32(uint32_t seed, uint64_t idx) {
  uint32_t lo = (uint32_t)idx, hi = (uint32_t)(idx >> 32);
  return lowbias32(lo ^ lowbias32(hi + 0x9e3779b9u * (seed + 1u)));
}

// keep-probability threshold on the top 24 bits: keep iff (u >> 8) >= thr.
__device__ __host__ __forceinline__ uint32_t dropout_threshold(float p) {
  return (uint32_t)(p * 16777216.0f);
}

__device__ __forceinline__ float hard_sigmoid_f(float x) { return fminf(fmaxf(0.2f * x + 0.5f, 0.f), 1.f); }

}  // namespace lstur
