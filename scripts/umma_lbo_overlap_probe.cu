// Hardware probe for the haloed wgrad tile: may the 32-channel (fp32) / 64-channel (bf16) MN-major slabs that make up the
// M = 128 rows of one tcgen05.mma OVERLAP in shared memory, i.e. can the descriptor's leading-dimension byte offset (the
// distance between consecutive 128-byte-wide slabs) be a few pixel rows (n * 128 B) instead of a whole slab?
// If so, ONE haloed x tile [pixels][channels] serves several filter taps in a single MMA: slab j of the A operand is the same
// tile read j * lbo_rows pixel rows later - the taps (r, s = 0..3) of a filter row for lbo_rows = 1.
//
//   D[j * SL + c][n] = sum_{k < 32} A[k + shift + j * lbo_rows][c] * B[k][n]      SL = 32 (fp32) / 64 (bf16), j < 128 / SL
//
// A is loaded by ONE TMA box [kRows pixel rows][128 bytes of channels] (SWIZZLE_128B_ATOM_32B for fp32, SWIZZLE_128B for
// bf16), B[64 rows][64] as usual.  For every (shift 0..9) x (lbo_rows 1, 2, 3, 34) the result is compared with the host.
//
// Build (no GPU needed):  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I pytortto_b200/csrc \
//                              -o scripts/_build/umma_lbo_overlap_probe scripts/umma_lbo_overlap_probe.cu
// NOT part of the product path; nothing imports it.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sm100_ptx.cuh"

using namespace ttb::ptx;

constexpr int kRows = 160, kBRows = 64, kN = 64, kK = 32, kM = 128;

template <bool BF16>
__global__ void __launch_bounds__(128)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ out,
             int shift_rows, int lbo_rows) {
  constexpr int kBSlabs = BF16 ? 1 : 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                  // ONE slab: kRows x 128 B
  uint8_t* sB = smem + kRows * 128;    // kBSlabs x 64 rows x 128 B
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
    mbar_expect_tx(&full_bar, (kRows + kBSlabs * kBRows) * 128);
    tma_load_2d(sA, &tmA, &full_bar, 0, 0);
    for (int sl = 0; sl < kBSlabs; ++sl) tma_load_2d(sB + sl * kBRows * 128, &tmB, &full_bar, sl * 32, 0);
  }
  mbar_wait(&full_bar, 0);
  tc_fence_after();
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(BF16 ? 1 : 2, 1, 1, kM, kN);
      constexpr uint32_t kRowsPerMma = BF16 ? 16 : 8;
      constexpr uint32_t sbo = BF16 ? 1024 : 512, layout = BF16 ? 2 : 1;
      const uint32_t a0 = smem_u32(sA) + (uint32_t)shift_rows * 128u;
      const uint32_t b0 = smem_u32(sB);
#pragma unroll
      for (uint32_t k = 0; k < kK / kRowsPerMma; ++k) {
        const uint64_t da = umma_desc(a0 + k * kRowsPerMma * 128, (uint32_t)lbo_rows * 128u, sbo, layout);
        const uint64_t db = umma_desc(b0 + k * kRowsPerMma * 128, kBRows * 128, sbo, layout);
        if (BF16) mma_bf16(tmem_base, da, db, idesc, k != 0);
        else mma_tf32(tmem_base, da, db, idesc, k != 0);
      }
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

static int make_map(PFN_encodeTiled enc, CUtensorMap* tm, void* base, int cols, int rows, bool bf16) {
  const int es = bf16 ? 2 : 4, per_row = 128 / es;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * es};
  cuuint32_t box[2] = {(cuuint32_t)per_row, (cuuint32_t)rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, bf16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return 1;
  }
  return 0;
}

template <bool BF16>
static int run(PFN_encodeTiled enc) {
  constexpr int SL = BF16 ? 64 : 32;  // channels per slab
  static float hA[kRows * SL], hB[kBRows * kN], hD[kM * kN];
  for (int r = 0; r < kRows; ++r)
    for (int c = 0; c < SL; ++c) hA[r * SL + c] = (float)((r * 7 + c * 3 + (r * c) % 5) % 13 - 6);  // small integers: exact
  for (int r = 0; r < kBRows; ++r)
    for (int n = 0; n < kN; ++n) hB[r * kN + n] = (float)((n * 5 + r * 11) % 9 - 4);
  void *dA, *dB;
  float* dD;
  const size_t es = BF16 ? 2 : 4;
  CK(cudaMalloc(&dA, sizeof(hA) / 4 * es));
  CK(cudaMalloc(&dB, sizeof(hB) / 4 * es));
  CK(cudaMalloc(&dD, sizeof(hD)));
  if (BF16) {
    static __nv_bfloat16 tA[kRows * SL], tB[kBRows * kN];
    for (int i = 0; i < kRows * SL; ++i) tA[i] = __float2bfloat16(hA[i]);
    for (int i = 0; i < kBRows * kN; ++i) tB[i] = __float2bfloat16(hB[i]);
    CK(cudaMemcpy(dA, tA, sizeof(tA), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, tB, sizeof(tB), cudaMemcpyHostToDevice));
  } else {
    CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
  }
  CUtensorMap tmA, tmB;
  if (make_map(enc, &tmA, dA, SL, kRows, BF16) || make_map(enc, &tmB, dB, kN, kBRows, BF16)) return 1;
  const size_t smem = (kRows + 2 * kBRows) * 128 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  printf("%s MN-major, M = 128 as %d overlapping slabs; rows = start shift, columns = slab distance in pixel rows; entry = max "
         "|D - host| (0 = exact)\n", BF16 ? "bf16 (SWIZZLE_128B, layout 2)" : "fp32/tf32 (SWIZZLE_128B_ATOM_32B, layout 1)", kM / SL);
  const int lbos[4] = {1, 2, 3, 34};
  printf("shift |   lbo=1    lbo=2    lbo=3   lbo=34\n");
  for (int shift = 0; shift <= 9; ++shift) {
    printf("%5d |", shift);
    for (int li = 0; li < 4; ++li) {
      const int lbo = lbos[li];
      CK(cudaMemset(dD, 0xff, sizeof(hD)));
      probe_kernel<BF16><<<1, 128, smem>>>(tmA, tmB, dD, shift, lbo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("  %s\n", cudaGetErrorString(e));
        return 1;
      }
      CK(cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost));
      double worst = 0.0;
      for (int m = 0; m < kM; ++m)
        for (int n = 0; n < kN; ++n) {
          const int j = m / SL, c = m % SL;
          double ref = 0.0;
          for (int k = 0; k < kK; ++k) ref += (double)hA[(k + shift + j * lbo) * SL + c] * hB[k * kN + n];
          double d = fabs((double)hD[m * kN + n] - ref);
          if (!(d <= worst)) worst = d;  // NaN-safe
        }
      printf(" %8.3g", worst);
    }
    printf("\n");
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
  if (run<false>(enc)) return 1;
  if (run<true>(enc)) return 1;
  return 0;
}
