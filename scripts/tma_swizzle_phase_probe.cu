// Hardware probe for the "flat-shift halo tile" (DESIGN.md section 8), part 3: when a SWIZZLE_128B TMA box is written to a
// shared-memory address that is 128-byte but NOT 1024-byte aligned, is the 16-byte-chunk XOR pattern taken from the
// ABSOLUTE shared-memory address bits [7:9] (then boxes of one padded image row each can be stacked at any row offset and
// tcgen05.mma, which also goes by address bits - umma_shift_probe.cu - reads them consistently), or from the row index
// INSIDE the box (then every box must start on a 1024-byte boundary)?
//
// One CTA loads a box of 16 rows x 32 fp32 (value = 100*row + col) to smem base + off*128 for off = 0..8 and dumps the raw
// shared memory; the host reports which hypothesis explains where every 16-byte chunk landed.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I pytortto_b200/csrc \
//              -o scripts/_build/tma_swizzle_phase_probe scripts/tma_swizzle_phase_probe.cu
// NOT part of the product path; nothing imports it.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "sm100_ptx.cuh"

using namespace ttb::ptx;

constexpr int kBoxRows = 16, kDumpRows = 32;

__global__ void __launch_bounds__(128)
probe_kernel(const __grid_constant__ CUtensorMap tm, float* __restrict__ dump, int off_rows) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar;
  float* sf = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < kDumpRows * 32; i += blockDim.x) sf[i] = -1.f;
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();  // the generic-proxy fill above must be ordered before the async-proxy (TMA) write
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&full_bar, kBoxRows * 128);
    tma_load_2d(smem + off_rows * 128, &tm, &full_bar, 0, 0);
  }
  mbar_wait(&full_bar, 0);
  for (int i = threadIdx.x; i < kDumpRows * 32; i += blockDim.x) dump[i] = sf[i];
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(fn);
  static float h[kBoxRows * 32], d[kDumpRows * 32];
  for (int r = 0; r < kBoxRows; ++r)
    for (int c = 0; c < 32; ++c) h[r * 32 + c] = (float)(100 * r + c);
  float *dsrc, *ddump;
  CK(cudaMalloc(&dsrc, sizeof(h)));
  CK(cudaMalloc(&ddump, sizeof(d)));
  CK(cudaMemcpy(dsrc, h, sizeof(h), cudaMemcpyHostToDevice));
  CUtensorMap tm;
  cuuint64_t dims[2] = {32, (cuuint64_t)kBoxRows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {32, (cuuint32_t)kBoxRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dsrc, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return 1;
  }
  const size_t smem = kDumpRows * 128 + 1024;
  printf("off = destination row offset (x128 B) of a 16-row SWIZZLE_128B box; chunks explained by each hypothesis (of 128)\n");
  printf("  off | absolute-address pattern | box-relative pattern | rows found where expected\n");
  for (int off = 0; off <= 8; ++off) {
    probe_kernel<<<1, 128, smem>>>(tm, ddump, off);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%5d | launch failed: %s\n", off, cudaGetErrorString(e));
      return 1;
    }
    CK(cudaMemcpy(d, ddump, sizeof(d), cudaMemcpyDeviceToHost));
    int abs_ok = 0, rel_ok = 0, row_ok = 0;
    for (int i = 0; i < kBoxRows; ++i) {
      const int srow = off + i;  // smem row the box row should occupy
      bool all_in_row = true;
      for (int c = 0; c < 8; ++c) {  // 16-byte chunk c of box row i holds values 100*i + 4c .. 4c+3
        const float want = (float)(100 * i + 4 * c);
        const int ca = c ^ (srow & 7), cr = c ^ (i & 7);
        if (d[srow * 32 + ca * 4] == want) ++abs_ok;
        if (d[srow * 32 + cr * 4] == want) ++rel_ok;
        bool found = false;
        for (int cc = 0; cc < 8; ++cc) found |= d[srow * 32 + cc * 4] == want;
        all_in_row &= found;
      }
      row_ok += all_in_row;
    }
    printf("%5d | %24d | %20d | %d of %d\n", off, abs_ok, rel_ok, row_ok, kBoxRows);
  }
  return 0;
}
