// Shared helpers for libtortto_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
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

// Experiment / tuning knobs (TTB_* environment variables) exist only in the tuning build
// (`python -m pytortto_b200.build --tuning` -> libtortto_b200_tuning.so, compiled with -DTTB_TUNING and selected with
// TORTTO_B200_LIB=tuning); the release library never reads the environment and always takes the default.
#ifdef TTB_TUNING
int tuning_knob(const char* name, int dflt);
#else
constexpr int tuning_knob(const char*, int dflt) { return dflt; }
#endif

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid size for a grid-stride elementwise kernel: a few waves of 148 SMs worth of CTAs
inline int elementwise_grid(int64_t work_items, int threads, int ctas_per_sm = 8) {
  int64_t need = ceil_div(work_items, threads);
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------------
// A training step is ~170 back-to-back kernels, many of them a few microseconds long; with plain stream order each
// one pays the launch latency + ramp of an empty GPU after its predecessor has fully drained.  Every kernel of this
// library is therefore launched with the programmatic-stream-serialization attribute and starts with
//   pdl_launch_dependents()  - the NEXT kernel of the stream may be scheduled now: its CTAs become resident as this grid's
//                              CTAs retire, and whatever it does before its own wait (barrier init, TMEM allocation,
//                              tensor-map prefetch) overlaps this grid's tail;
//   pdl_wait()               - blocks until the PREVIOUS grid has completed and its writes are visible.  No kernel touches
//                              global memory before this point, so the usual stream-order guarantees hold unchanged
//                              (completion is transitive: the previous grid only completes after its own wait returned).
// Both instructions are no-ops when the launch carries no programmatic dependency; the relation survives CUDA-graph
// capture as a programmatic edge.  TTB_PDL=0 (tuning build) launches without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() {
  pdl_launch_dependents();
  pdl_wait();
}

template <class... KArgs, class... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static const int pdl = tuning_knob("TTB_PDL", 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface through check_launch (cudaGetLastError)
}

// What a convolution epilogue does to an accumulator element of output channel k before it is stored (the device-side
// form of ttb_conv_epilogue):
//   v = acc * scale[k] + bias[k]   (either may be null)     - conv bias, or an eval-mode BatchNorm folded to scale / shift
//   v += accum[same element]       (accum may be null)      - a residual / a gradient already pending for the same tensor
//   v = max(v, 0)                  (relu != 0)
//   stats (may be null): per-chunk partial sums [chunks][2][K] (double) of v and v*v for the BatchNorm that follows
//   bn_x != null: `stats` receives instead the two sums a BatchNorm BACKWARD needs over the stored values v (a dgrad whose
//   output is the gradient that BatchNorm(+ReLU) node reads next): sum(g) and sum(g * (x - mean[k])) with g = v, or
//   g = v where fmaf(x - mean, rscale, rshift) > 0 else 0 (the ReLU mask recomputed exactly as forward computed it);
//   bn_x = the BatchNorm's input, same shape / layout as the output
struct Epilogue {
  const float* scale;
  const float* bias;
  const float* accum;
  int relu;
  double* stats;
  const float* bn_x;
  const float* bn_mean;
  const float* bn_rscale;  // may be null (no ReLU behind the BatchNorm): then bn_rshift is unused
  const float* bn_rshift;
};
inline Epilogue bias_epilogue(const float* bias) { return Epilogue{nullptr, bias, nullptr, 0, nullptr}; }

#define TTB_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ttb::set_error(__VA_ARGS__);      \
      return 2;                         \
    }                                   \
  } while (0)

__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st_f4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
// 4 floats -> 4 bf16 (round to nearest even), one 8-byte store
__device__ __forceinline__ void st_bf16x4(__nv_bfloat16* p, const float4& v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 packed;
  packed.x = *reinterpret_cast<uint32_t*>(&lo);
  packed.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = packed;
}
// streaming variants: read-once / write-once data should not displace reusable lines in L1
__device__ __forceinline__ float4 ld_f4_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

}  // namespace ttb
