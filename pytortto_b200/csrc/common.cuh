// Shared helpers for libtortto_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tortto_b200.h"

namespace ttb {

void set_error(const char* fmt, ...);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Checks the launch; returns non-zero (and records the text) on failure.  Never synchronises.
inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int sm_count();

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid size for a grid-stride elementwise kernel: a few waves of 148 SMs worth of CTAs
inline int elementwise_grid(int64_t work_items, int threads, int ctas_per_sm = 8) {
  int64_t need = ceil_div(work_items, threads);
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

#define TTB_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ttb::set_error(__VA_ARGS__);      \
      return 2;                         \
    }                                   \
  } while (0)

__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st_f4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
// streaming variants: read-once / write-once data should not displace reusable lines in L1
__device__ __forceinline__ float4 ld_f4_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

}  // namespace ttb
