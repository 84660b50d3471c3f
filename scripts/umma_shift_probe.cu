// Hardware probe for the "flat-shift halo tile" (DESIGN.md section 8): can a tcgen05.mma A operand start at an
// arbitrary 128-byte ROW of a 128B-swizzled shared-memory tile that one TMA box filled?
//
// One CTA loads A_full[192 rows][32 fp32] and B[64 rows][32 fp32] (K-major, CU_TENSOR_MAP_SWIZZLE_128B), then computes
// D[128][64] = A_full[shift .. shift+128) * B^T with the A descriptor's start address moved by `shift` rows, for every
// shift in 0..9 and every value 0..7 of the descriptor's base-offset field (bits 49-51), and compares with the host.
// Expected outcome if the idea works: for each shift exactly one base-offset value (probably shift % 8, or 0 if the
// hardware derives the swizzle phase from the address bits alone) reproduces the host result.
//
// Build (no GPU needed):  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I pytortto_b200/csrc \
//                              -o scripts/_build/umma_shift_probe scripts/umma_shift_probe.cu
// Run on a B200:          scripts/_build/umma_shift_probe
// NOT part of the product path; nothing imports it.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sm100_ptx.cuh"

using namespace ttb::ptx;

constexpr int kRows = 192, kN = 64, kK = 32, kM = 128;

__device__ __forceinline__ uint64_t desc_sw128_bo(uint32_t smem_addr, uint32_t base_offset) {
  uint64_t d = umma_desc_sw128(smem_addr, 16, 1024);
  d |= (uint64_t)(base_offset & 7) << 49;
  return d;
}

__global__ void __launch_bounds__(128)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ out,
             int shift_rows, int base_offset) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                      // 192 x 128 B
  uint8_t* sB = smem + kRows * 128;        // 64 x 128 B (offset 24576 = 24 * 1024: atom aligned)
  __shared__ uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(&full_bar, 1);
      mbar_init(&done_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<64>(&tmem_base_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&full_bar, (kRows + kN) * 128);
    tma_load_2d(sA, &tmA, &full_bar, 0, 0);
    tma_load_2d(sB, &tmB, &full_bar, 0, 0);
  }
  mbar_wait(&full_bar, 0);
  tc_fence_after();
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(2 /*tf32*/, 0, 0, kM, kN);
      const uint32_t a0 = smem_u32(sA) + (uint32_t)shift_rows * 128u;
      const uint32_t b0 = smem_u32(sB);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma_tf32(tmem_base, desc_sw128_bo(a0 + k * 32, (uint32_t)base_offset), umma_desc_sw128(b0 + k * 32, 16, 1024), idesc,
                 k != 0);
      mma_commit(&done_bar);
    }
    __syncwarp();
  }
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  for (int cb = 0; cb < kN / 32; ++cb) {
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cb * 32), r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * kN + cb * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem_base);
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

static int make_map(PFN_encodeTiled enc, CUtensorMap* tm, void* base, int rows) {
  cuuint64_t dims[2] = {(cuuint64_t)kK, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kK * 4};
  cuuint32_t box[2] = {(cuuint32_t)kK, (cuuint32_t)rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return 1;
  }
  return 0;
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(fn);
  static float hA[kRows * kK], hB[kN * kK], hD[kM * kN];
  for (int r = 0; r < kRows; ++r)
    for (int k = 0; k < kK; ++k) hA[r * kK + k] = (float)((r * 7 + k * 3) % 13 - 6);  // small integers: exact in tf32
  for (int n = 0; n < kN; ++n)
    for (int k = 0; k < kK; ++k) hB[n * kK + k] = (float)((n * 5 + k * 11) % 9 - 4);
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, sizeof(hA)));
  CK(cudaMalloc(&dB, sizeof(hB)));
  CK(cudaMalloc(&dD, sizeof(hD)));
  CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
  CUtensorMap tmA, tmB;
  if (make_map(enc, &tmA, dA, kRows) || make_map(enc, &tmB, dB, kN)) return 1;
  const size_t smem = (kRows + kN) * 128 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  printf("rows of A_full shifted by `shift`; columns = descriptor base-offset field; entry = max |D - host| (0 = exact)\n");
  printf("shift |");
  for (int bo = 0; bo < 8; ++bo) printf("   bo=%d  ", bo);
  printf("\n");
  for (int shift = 0; shift <= 9; ++shift) {
    printf("%5d |", shift);
    for (int bo = 0; bo < 8; ++bo) {
      CK(cudaMemset(dD, 0xff, sizeof(hD)));
      probe_kernel<<<1, 128, smem>>>(tmA, tmB, dD, shift, bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("  %s\n", cudaGetErrorString(e));
        return 1;
      }
      CK(cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost));
      double worst = 0.0;
      for (int m = 0; m < kM; ++m)
        for (int n = 0; n < kN; ++n) {
          double ref = 0.0;
          for (int k = 0; k < kK; ++k) ref += (double)hA[(m + shift) * kK + k] * hB[n * kK + k];
          double d = fabs((double)hD[m * kN + n] - ref);
          if (!(d <= worst)) worst = d;  // NaN-safe
        }
      printf(" %8.3g", worst);
    }
    printf("\n");
  }
  return 0;
}
