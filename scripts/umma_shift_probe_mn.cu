// Hardware probe, MN-major twin of umma_shift_probe.cu (wgrad operands: the reduction index = pixel is the ROW of the
// shared-memory tiles, fp32 MN-major needs the 128B-span / 32B-atom swizzle): can the A operand start at an arbitrary
// pixel row, i.e. could a wgrad reuse ONE haloed x tile for all filter taps?
//
// One CTA loads A[64 pixel rows][128 fp32] as 4 slabs of 32 columns and B[64 rows][64 fp32] as 2 slabs
// (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), then computes D[m][n] = sum_{k<32} A[k + shift][m] * B[k][n] with the A
// descriptor's start address moved by `shift` rows, for shift 0..9 and base-offset field 0..7, and compares with the host.
//
// Build (no GPU needed):  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I pytortto_b200/csrc \
//                              -o scripts/_build/umma_shift_probe_mn scripts/umma_shift_probe_mn.cu
// Run on a B200:          scripts/_build/umma_shift_probe_mn
// NOT part of the product path; nothing imports it.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sm100_ptx.cuh"

using namespace ttb::ptx;

constexpr int kRows = 64, kN = 64, kK = 32, kM = 128;  // kRows pixel rows loaded, kK of them reduced

// MN-major fp32: layout type 1 (128B span / 32B atom), 128-byte-wide slabs kRows*128 bytes apart, 4-row groups 512 B apart
__device__ __forceinline__ uint64_t desc_mn_bo(uint32_t smem_addr, uint32_t base_offset) {
  uint64_t d = umma_desc(smem_addr, kRows * 128, 512, 1);
  d |= (uint64_t)(base_offset & 7) << 49;
  return d;
}

__global__ void __launch_bounds__(128)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ out,
             int shift_rows, int base_offset) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                          // 4 slabs x 64 rows x 128 B
  uint8_t* sB = smem + 4 * kRows * 128;        // 2 slabs x 64 rows x 128 B
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
    mbar_expect_tx(&full_bar, 6 * kRows * 128);
    for (int sl = 0; sl < 4; ++sl) tma_load_2d(sA + sl * kRows * 128, &tmA, &full_bar, sl * 32, 0);
    for (int sl = 0; sl < 2; ++sl) tma_load_2d(sB + sl * kRows * 128, &tmB, &full_bar, sl * 32, 0);
  }
  mbar_wait(&full_bar, 0);
  tc_fence_after();
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(2 /*tf32*/, 1, 1, kM, kN);
      const uint32_t a0 = smem_u32(sA) + (uint32_t)shift_rows * 128u;
      const uint32_t b0 = smem_u32(sB);
#pragma unroll
      for (int k = 0; k < kK / 8; ++k)  // 8 pixel rows per MMA
        mma_tf32(tmem_base, desc_mn_bo(a0 + k * 1024, (uint32_t)base_offset), desc_mn_bo(b0 + k * 1024, 0), idesc, k != 0);
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

static int make_map(PFN_encodeTiled enc, CUtensorMap* tm, void* base, int cols) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)kRows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)kRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
  static float hA[kRows * kM], hB[kRows * kN], hD[kM * kN];
  for (int r = 0; r < kRows; ++r)
    for (int m = 0; m < kM; ++m) hA[r * kM + m] = (float)((r * 7 + m * 3) % 13 - 6);  // small integers: exact in tf32
  for (int r = 0; r < kRows; ++r)
    for (int n = 0; n < kN; ++n) hB[r * kN + n] = (float)((n * 5 + r * 11) % 9 - 4);
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, sizeof(hA)));
  CK(cudaMalloc(&dB, sizeof(hB)));
  CK(cudaMalloc(&dD, sizeof(hD)));
  CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
  CUtensorMap tmA, tmB;
  if (make_map(enc, &tmA, dA, kM) || make_map(enc, &tmB, dB, kN)) return 1;
  const size_t smem = 6 * kRows * 128 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  printf("MN-major: pixel rows of A shifted by `shift`; columns = descriptor base-offset field; entry = max |D - host| (0 = exact)\n");
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
          for (int k = 0; k < kK; ++k) ref += (double)hA[(k + shift) * kM + m] * hB[k * kN + n];
          double d = fabs((double)hD[m * kN + n] - ref);
          if (!(d <= worst)) worst = d;  // NaN-safe
        }
      printf(" %8.3g", worst);
    }
    printf("\n");
  }
  return 0;
}
