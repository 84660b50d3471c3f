// tcgen05 implicit-GEMM convolution for sm_100a: fprop, dgrad and wgrad with NO materialised im2col.
//
// Replaces the reference's "as_strided window view + einsum" (autograd/grad_nn.py:595-682), whose tensordot
// reshape copies the whole im2col matrix (kh*kw x the activation) on every call.
//
// GEMM views (activations NHWC fp32, weights KRSC fp32, TF32 tensor-core math, fp32 accumulation in TMEM):
//   fprop : Y[m, k]      = sum_{tap,c} A[m, (tap,c)] * W[k, (tap,c)]       m = (n,p,q) output pixel
//   dgrad : dX[m, c]     = sum_{tap,k} dY[m', (tap,k)] * Wt[c, (tap,k)]     per stride-parity class of input pixels
//   wgrad : dW[k,(tap,c)] = sum_m dY[m, k] * A[m, (tap,c)]                  split over pixel ranges
// A-tiles (128 pixels x 32 channels for one filter tap) are fetched by ONE TMA instruction each, using an
// im2col-mode tensor map: the hardware walks output pixels across rows and images, applies stride, adds the
// filter-tap offset and zero-fills the padding halo.  Weight / dY tiles use tiled-mode TMA.  Every tile lands in
// shared memory in the 128-byte-swizzled layout tcgen05.mma consumes directly.
//
// Kernel structure: warp-specialised - TMA producer warp(s), one MMA-issuing warp (+TMEM owner), four epilogue
// warps (TMEM -> registers -> padded smem transpose -> coalesced 128 B row-segment stores).  A ring of NSTAGES {A,B}
// stages is handed over with full/empty mbarriers; tcgen05.commit releases stages.  fprop / dgrad: persistent CTAs
// over a static tile list, several producer warps, two TMEM accumulators (igemm_fwd_persist_kernel); wgrad: one tile
// and one pixel range per CTA (igemm_wgrad_kernel).  Producer and MMA loops run warp-uniform with the asynchronous
// instruction under elect.sync - from an `if (lane == 0)` branch ptxas wraps each one in a divergence waterfall.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "conv_epilogue.cuh"
#include "sm100_ptx.cuh"

namespace ttb {

// conv_wgrad_halo.cu: haloed-tile wgrad of the narrow stride-1 layers
bool halo_wgrad_selected(const ttb_conv_desc* d);
size_t halo_wgrad_workspace(const ttb_conv_desc* d);
int halo_wgrad(const ttb_conv_desc* d, const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes, cudaStream_t st,
               int* splits_out);

constexpr int kMaxTaps = 64;
constexpr int kTileM = 128;          // UMMA M (rows of the accumulator = TMEM lanes)

// operand element type of one launch: fp32 read as TF32 (32 per 128-byte swizzle row) or bf16 (64 per row)
struct Elem {
  bool bf16;
  int size;    // bytes
  int per_row; // elements per 128-byte row = channels per K-block / per MN slab
  CUtensorMapDataType dt;
};
static inline Elem elem_of(bool bf16) {
  return bf16 ? Elem{true, 2, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16} : Elem{false, 4, 32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32};
}
constexpr int kThreadsIgemm = 192;

// ------------------------------------------------------------------------------------------------------------
// driver entry points (resolved through the runtime so the library has no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encodeTiled = nullptr;
static PFN_encodeIm2col g_encodeIm2col = nullptr;

static int load_driver_fns() {
  if (g_encodeTiled && g_encodeIm2col) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return 1;
  }
  g_encodeTiled = (PFN_encodeTiled)fn;
  fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) {
    set_error("cuTensorMapEncodeIm2col not available from the driver");
    return 1;
  }
  g_encodeIm2col = (PFN_encodeIm2col)fn;
  return 0;
}

// 2-D row-major matrix [rows][cols] -> tiled map with box [box_rows][128 bytes], 128B swizzle, zero OOB fill
// (`row_stride`: elements between rows, 0 = dense; a channel group of a wider tensor has cols < row_stride)
static int make_tiled_2d(CUtensorMap* tm, const void* base, Elem el, uint64_t rows, uint64_t cols, uint32_t box_rows,
                         CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, uint64_t row_stride = 0) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {(row_stride ? row_stride : cols) * (uint64_t)el.size};
  cuuint32_t box[2] = {(cuuint32_t)el.per_row, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encodeTiled(tm, el.dt, 2, (void*)base, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box_rows=%u", (int)r, (unsigned long long)rows,
              (unsigned long long)cols, box_rows);
    return 1;
  }
  return 0;
}

// generic tiled map (rank <= 5): dims / byte strides (strides[i] = stride of dim i+1) / box / traversal strides
static int make_tiled_nd(CUtensorMap* tm, const void* base, Elem el, int rank, const cuuint64_t* dims,
                         const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapSwizzle swz) {
  CUresult r = g_encodeTiled(tm, el.dt, (cuuint32_t)rank, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(rank %d) failed (%d)", rank, (int)r);
    return 1;
  }
  return 0;
}

// for the other tensor-core translation units (conv_wgrad_halo.cu): a tiled map of rank <= 5 with the swizzle of an MN-major
// (fp32: 128B span / 32B atom, bf16: 128B) or K-major (128B) operand
int tma_make_tiled(CUtensorMap* tm, const void* base, bool bf16, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides, bool mn_major) {
  if (load_driver_fns()) return 1;
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = elem_strides[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  const CUtensorMapSwizzle swz = (mn_major && !bf16) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  return make_tiled_nd(tm, base, elem_of(bf16), rank, d, s, b, e, swz);
}

// NHWC activation [n][h][w][c] -> im2col map: `pixels` base pixels x 128 bytes of channels per load, traversal strides
// (tw, th), bounding-box corners in W/H order, 128B swizzle, zero fill outside the image.
// (`cs`: elements between pixels, 0 = c; a channel group of a wider tensor has c < cs)
static int make_im2col_4d(CUtensorMap* tm, const void* base, Elem el, int n, int h, int w, int c, int lower_w, int lower_h,
                          int upper_w, int upper_h, int tw, int th, uint32_t pixels,
                          CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, int cs = 0) {
  const cuuint64_t es = (cuuint64_t)el.size;
  if (cs == 0) cs = c;
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)cs * es, (cuuint64_t)w * cs * es, (cuuint64_t)h * w * cs * es};
  int lower[2] = {lower_w, lower_h};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)tw, (cuuint32_t)th, 1};
  CUresult r = g_encodeIm2col(tm, el.dt, 4, (void*)base, dims, strides, lower, upper,
                              (cuuint32_t)el.per_row, pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col failed (%d) nhwc=%d,%d,%d,%d lower=(%d,%d) upper=(%d,%d) stride=(%d,%d)", (int)r, n,
              h, w, c, lower_w, lower_h, upper_w, upper_h, tw, th);
    return 1;
  }
  // Known driver issue (<= 13.1) with im2col maps over tensors smaller than 128 KiB: bit 21 of the second
  // descriptor word must be cleared (same workaround CUTLASS applies, cute/atom/copy_traits_sm90_im2col.hpp).
  int drv = 0;
  cudaDriverGetVersion(&drv);
  if (drv <= 13010 && (uint64_t)n * h * w * cs * es < 131072) reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// kernel parameters
// ------------------------------------------------------------------------------------------------------------
struct OutMap {            // accumulator row m -> element offset of the output row
  float* out;
  int64_t n_stride, h_stride, w_stride, base;  // in elements
  int p_dim, q_dim;        // m = (n*p_dim + i)*q_dim + j
  int q_valid;             // rows with j >= q_valid are dropped (row-shared A strips index a W-padded grid); 0 = all kept
  int m_total;             // rows that exist
  int n_total;             // valid columns (row length to write)
};

struct FwdParams {         // fprop / dgrad (K-major A via im2col TMA, K-major B via tiled TMA)
  CUtensorMap tmA, tmB;
  OutMap o;
  Epilogue ep;             // per-output-channel scale / bias, accumulate-into, ReLU (common.cuh)
  int c_blocks;            // reduction channels / 32
  int num_taps;
  int base_w, base_h;      // coordinates of base pixel (i=0, j=0)
  int trav_w, trav_h;      // traversal strides
  int b_koff[kMaxTaps];    // column offset of the tap's K-slice in the weight matrix
  uint16_t off_w[kMaxTaps], off_h[kMaxTaps];
  int dbg;                 // TTB_IGEMM_DBG bits (timing experiments, wrong results): 1 no MMA, 2 no A loads, 4 no B loads
  long long* trace;        // optional clock64() timeline of CTA (0,0): see scripts/igemm_trace.py
};

struct WgradParams {       // wgrad (MN-major A = dY via tiled TMA, MN-major B = im2col(x) via im2col TMA)
  CUtensorMap tmDy, tmX;
  OutMap o;                // rows = output channel k, columns = (tap, c) flattened; out = partial buffer of this split
  int64_t split_stride;    // elements between the partial buffers of consecutive splits
  int c;                   // input channels
  int pixel_steps_total;   // ceil(M / KP)
  int steps_per_split;
  int p_dim, q_dim;        // output pixel grid (for decomposing the pixel index of a step)
  int base_w, base_h, trav_w, trav_h;
  uint32_t desc_lbo, desc_sbo, desc_layout;  // UMMA smem-descriptor fields for the MN-major operands
  // "rectangular" fast path: the KP pixels of a step form a box (rect_w x rect_h x rect_n) of the output grid, so all
  // channel slabs of one tap come from ONE tiled 5-D load (tmX = [32|64 ch][W][H][N][C/32|64]) and all slabs of dY
  // from ONE tiled 3-D load (tmDy = [ch][pixels][K/32|64]) - 2..5 TMA instructions per step instead of 12.
  int rect;                // 0: im2col / per-slab loads
  int pad_w, pad_h, dil_w, dil_h;
  uint16_t off_w[kMaxTaps], off_h[kMaxTaps];
  int dbg;                 // TTB_IGEMM_DBG bits (timing experiments, wrong results): 1 no MMA, 2 no dY loads, 4 no x loads
};

// ------------------------------------------------------------------------------------------------------------
// accumulator row m of a tile -> element offset of its output row (conv_epilogue.cuh does the rest)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t out_row_offset(const OutMap& o, int m) {
  if (m >= o.m_total) return -1;
  const int j = m % o.q_dim;
  if (o.q_valid && j >= o.q_valid) return -1;
  const int t = m / o.q_dim;
  const int i = t % o.p_dim;
  const int n = t / o.p_dim;
  return o.base + (int64_t)n * o.n_stride + (int64_t)i * o.h_stride + (int64_t)j * o.w_stride;
}

// Up to kMaxMulti independent problems of the same shape class in one launch: the stride-parity classes of a strided
// dgrad run as one persistent grid instead of 4 small back-to-back grids.
constexpr int kMaxMulti = 4;
struct FwdParamsMulti {
  FwdParams p[kMaxMulti];
};

// ------------------------------------------------------------------------------------------------------------
// Persistent fprop / dgrad kernel: one CTA per SM walks a static list of output tiles (tile = blockIdx.x + i*gridDim.x)
// of up to kMaxMulti problems (the stride-parity classes of a strided dgrad).
//   * NPROD producer threads (one per warp) issue the TMA loads, K-block g belongs to producer g % NPROD: a K-block costs
//     one thread ~450 cycles of dependent mbarrier.try_wait / arrive.expect_tx / 2 x cp.async.bulk.tensor latency
//     (measured), more than the 128-256 cycles the tensor core needs for it, so a single producer starves the MMAs.
//     NSTAGES % NPROD == 0 keeps every stage with one producer (no parity aliasing between threads).
//   * the smem ring (~192 KB) keeps running across tile boundaries: no pipeline drain / refill per tile.
//   * two accumulator buffers in TMEM: the epilogue of tile i overlaps the main loop of tile i+1.
// Warp roles: [0, NPROD) producers, NPROD = MMA issuer + TMEM owner, NPROD+1 .. NPROD+4 epilogue.
// ------------------------------------------------------------------------------------------------------------
// STATS: the epilogue also accumulates the per-channel sum / sum of squares of everything this CTA stores and writes
// them as one segment of a [gridDim.x / nt][2][n_total] double partial buffer (P.ep.stats); the host then sizes the grid
// as a multiple of nt, so that every CTA keeps the same N tile for all of its tiles (conv_epilogue.cuh).
// WT > 1 ("row-shared A strip", the 3x3 / stride-1 layers): the WT filter taps of one filter ROW differ only by a shift of
// one pixel, so output pixels are indexed over a grid that is padded in W (q_dim = Q + WT - 1 positions per row, the last
// WT - 1 computed and dropped) and ONE im2col load of kTileM + WT - 1 consecutive positions serves all WT taps: tap j is the
// same strip read through a UMMA descriptor that starts j rows (j * 128 bytes) later (a 128B-swizzled K-major operand may
// start at any 128-byte row: profiles/r1_umma_row_shift_probe.txt).  A stage then holds the strip + WT weight tiles and
// costs one hand-shake per 4*WT MMAs; A traffic from L2 drops from R*S to R loads per tile (the 64 / 128-channel layers are
// bound by exactly that feed: 97-128 B/clk/SM wanted at full tensor rate against ~100 available).
template <int BN, int NSTAGES, int NPROD, int WT, bool BF16, int STATS>
__global__ void __launch_bounds__((NPROD + 5) * 32, BN <= 64 ? 2 : 1)  // narrow tiles: two CTAs per SM (<= 128 registers)
igemm_fwd_persist_kernel(const __grid_constant__ FwdParamsMulti PM, const int count) {
  pdl_launch_dependents();  // (the matching pdl_wait() follows the prologue below)
  static_assert(NSTAGES % NPROD == 0, "a stage must always be filled by the same producer thread");
  constexpr int kElems = BF16 ? 64 : 32;
  constexpr uint32_t kARows = kTileM + WT - 1;
  constexpr uint32_t kALoadBytes = kARows * 128;                       // what the TMA writes
  constexpr uint32_t kABytes = (kALoadBytes + 1023u) & ~1023u;         // (weight tiles stay 1024-byte aligned)
  constexpr uint32_t kBBytes = BN * 128;
  constexpr uint32_t kStageBytes = kABytes + WT * kBBytes;  // a pipeline stage (one mbarrier hand-shake): strip + WT weight tiles
  constexpr int kAccCols = BN < 32 ? 32 : BN;
  constexpr int kTmemCols = 2 * kAccCols;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* staging = reinterpret_cast<float*>(smem + (size_t)NSTAGES * kStageBytes);
  __shared__ uint64_t full_bar[NSTAGES], empty_bar[NSTAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_smem;

  // The producer / MMA loops below are executed by WHOLE warps with warp-uniform control flow and operands, and only
  // the asynchronous instruction itself sits under elect.sync.  Issued from a `lane == 0` branch instead, every
  // UTMALDG / UTCHMMA / UTCBAR gets a divergence "waterfall" (ELECT + R2UR.BROADCAST + BRA.U.ANY loop) around it and a
  // K-block costs ~450 cycles of issue latency - more than its 128..512 cycles of tensor-core time (measured).
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int nt = (PM.p[0].o.n_total + BN - 1) / BN;  // N tiles (same for every problem of the launch)
  int tiles_total = 0;
  for (int c = 0; c < count; ++c) tiles_total += ((PM.p[c].o.m_total + kTileM - 1) / kTileM) * nt;
  long long* const tr = (PM.p[0].trace && blockIdx.x == 0) ? PM.p[0].trace : nullptr;

  if (warp == 0 && ptx::elect_one())
    for (int c = 0; c < count; ++c) {
      ptx::prefetch_tmap(&PM.p[c].tmA);
      ptx::prefetch_tmap(&PM.p[c].tmB);
    }
  if (warp == NPROD) {
    if (lane == 0) {
      for (int s = 0; s < NSTAGES; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(&acc_full[b], 1);
        ptx::mbar_init(&acc_empty[b], 4);  // one arrival per epilogue warp
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<kTmemCols>(&tmem_base_smem);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();  // barriers, TMEM and tensor-map prefetch above overlapped the previous kernel's tail; global memory from here on
  if (tr && threadIdx.x == 0) tr[1] = clock64();

  // tile index -> problem, first row, first column
  auto decode = [&](int t, int& cls, int& m0, int& n0) {
    cls = 0;
    while (cls + 1 < count) {
      const int tc = ((PM.p[cls].o.m_total + kTileM - 1) / kTileM) * nt;
      if (t < tc) break;
      t -= tc;
      ++cls;
    }
    m0 = (t / nt) * kTileM;
    n0 = (t % nt) * BN;
  };

  if (warp < NPROD) {
    // ===================== TMA producers (whole warp, one elected lane issues) =====================
    const int dbg = PM.p[0].dbg;
    const uint32_t tx_bytes = ((dbg & 2) ? 0 : kALoadBytes) + ((dbg & 4) ? 0 : WT * kBBytes);
    uint32_t g = 0;  // stages filled by this CTA so far, counted by every producer; stage-fill g belongs to producer g % NPROD
    for (int t = blockIdx.x; t < tiles_total; t += gridDim.x) {
      int cls, m0, n0;
      decode(t, cls, m0, n0);
      const FwdParams& P = PM.p[cls];
      const int j = m0 % P.o.q_dim;
      const int tt = m0 / P.o.q_dim;
      const int i = tt % P.o.p_dim;
      const int n = tt / P.o.p_dim;
      const int cw = P.base_w + j * P.trav_w, ch = P.base_h + i * P.trav_h;
      const int num_kb = (P.num_taps / WT) * P.c_blocks;   // stage fills per tile: (tap or tap row) x channel block
      for (int kb = 0; kb < num_kb; ++kb, ++g) {
        if ((int)(g % NPROD) != warp) continue;
        const uint32_t stage = g % NSTAGES, phase = (g / NSTAGES) & 1u;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (tr && lane == 0 && g < 1000) tr[2048 + g] = clock64();
        uint8_t* sa = smem + stage * kStageBytes;
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&full_bar[stage], tx_bytes);
          const int tap0 = (kb / P.c_blocks) * WT;
          const int cb = kb - (kb / P.c_blocks) * P.c_blocks;
          // (WT > 1: the strip is loaded at the row's smallest W offset, 0; tap j reads it shifted by off_w[tap0 + j] rows)
          if (!(dbg & 2))
            ptx::tma_load_im2col_4d(sa, &P.tmA, &full_bar[stage], cb * kElems, cw, ch, n, WT > 1 ? (uint16_t)0 : P.off_w[tap0],
                                    P.off_h[tap0]);
          if (!(dbg & 4)) {
#pragma unroll
            for (int u = 0; u < WT; ++u)
              ptx::tma_load_2d(sa + kABytes + u * kBBytes, &P.tmB, &full_bar[stage], P.b_koff[tap0 + u] + cb * kElems, n0);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == NPROD) {
    // ===================== MMA issuer (whole warp, one elected lane issues) =====================
    constexpr uint32_t idesc = ptx::umma_idesc(BF16 ? 1 /*bf16*/ : 2 /*tf32*/, 0, 0, kTileM, BN);
    const int dbg = PM.p[0].dbg;
    uint32_t g = 0;
    int it = 0;
    for (int t = blockIdx.x; t < tiles_total; t += gridDim.x, ++it) {
      int cls, m0, n0;
      decode(t, cls, m0, n0);
      const FwdParams& P = PM.p[cls];
      const int num_kb = (P.num_taps / WT) * P.c_blocks;
      const int buf = it & 1;
      ptx::mbar_wait(&acc_empty[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator
      ptx::tc_fence_after();
      if (tr && lane == 0 && it < 250) tr[16 + 2 * it] = clock64();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * kAccCols);
      for (int kb = 0; kb < num_kb; ++kb, ++g) {
        const uint32_t stage = g % NSTAGES, phase = (g / NSTAGES) & 1u;
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (tr && lane == 0 && g < 1000) tr[3072 + g] = clock64();
        const uint32_t s0 = ptx::smem_u32(smem + stage * kStageBytes);
        const int tap0 = (kb / P.c_blocks) * WT;
        uint32_t shift[WT];  // A-strip row shift of each tap of this stage (warp-uniform)
#pragma unroll
        for (int u = 0; u < WT; ++u) shift[u] = WT > 1 ? (uint32_t)P.off_w[tap0 + u] * 128u : 0u;
        if (ptx::elect_one()) {
          if (!(dbg & 1)) {
            const int nmma = (dbg & 8) ? 1 : (dbg & 16) ? 2 : 4;  // experiment: fewer MMAs per K-block (wrong results)
#pragma unroll
            for (int u = 0; u < WT; ++u) {
              const uint32_t sa = s0 + shift[u], sb = s0 + kABytes + u * kBBytes;
#pragma unroll
              for (int k = 0; k < 4; ++k) {  // one MMA consumes 32 bytes of K per row: 8 tf32 or 16 bf16 elements
                if (k >= nmma) break;
                uint64_t da = ptx::umma_desc_sw128(sa + k * 32, 16, 1024);
                uint64_t db = ptx::umma_desc_sw128(sb + k * 32, 16, 1024);
                if (BF16) ptx::mma_bf16(d_tmem, da, db, idesc, (kb | u | k) != 0);
                else ptx::mma_tf32(d_tmem, da, db, idesc, (kb | u | k) != 0);
              }
            }
          }
          ptx::mma_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs above have read it
        }
        __syncwarp();
      }
      if (ptx::elect_one()) ptx::mma_commit(&acc_full[buf]);  // accumulator of this tile complete
      __syncwarp();
      if (tr && lane == 0 && it < 250) tr[17 + 2 * it] = clock64();
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - (NPROD + 1);
    float* const my_stage = staging + ew * (32 * kStagePitch);
    EpiStats<BN> es;
    if (STATS) es.reset();
    int it = 0, n0_last = 0;
    for (int t = blockIdx.x; t < tiles_total; t += gridDim.x, ++it) {
      int cls, m0, n0;
      decode(t, cls, m0, n0);
      const FwdParams& P = PM.p[cls];
      const int buf = it & 1;
      const int64_t my_off = out_row_offset(P.o, m0 + (warp & 3) * 32 + lane);
      // While this warp waits for the tile's MMAs, the rows of the tensors its epilogue will READ at the output positions
      // (the pending gradient / residual, the BatchNorm input of the backward sums) are pulled into L2: their loads inside
      // the epilogue are dependent round trips per 32-column block, ~4x shorter from L2 than from HBM.
      if ((STATS == 2 || P.ep.accum != nullptr) && my_off >= 0) {
#pragma unroll
        for (int cb = 0; cb < BN / 32; ++cb) {
          if (n0 + cb * 32 >= P.o.n_total) break;
          if (P.ep.accum != nullptr) ptx::prefetch_l2(P.ep.accum + my_off + n0 + cb * 32);
          if (STATS == 2) ptx::prefetch_l2(P.ep.bn_x + my_off + n0 + cb * 32);
        }
      }
      ptx::mbar_wait(&acc_full[buf], ((uint32_t)it >> 1) & 1u);
      ptx::tc_fence_after();
      if (tr && ew == 0 && lane == 0 && it < 250) tr[528 + 2 * it] = clock64();
      epilogue_tile<BN, STATS>(tmem_base + (uint32_t)(buf * kAccCols), my_stage, P.o.out, my_off, n0, P.o.n_total, P.ep,
                               warp & 3, es);
      n0_last = n0;
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[buf]);
      if (tr && ew == 0 && lane == 0 && it < 250) tr[529 + 2 * it] = clock64();
    }
    if (STATS && it > 0) {
      const FwdParams& P = PM.p[0];
      epilogue_stats_flush<BN>(es, staging, ew, 1, P.ep.stats + (int64_t)(blockIdx.x / nt) * 2 * P.o.n_total, n0_last,
                               P.o.n_total);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == NPROD) ptx::tmem_dealloc<kTmemCols>(tmem_base);
  if (tr && threadIdx.x == 0) tr[2] = clock64();
}

// ------------------------------------------------------------------------------------------------------------
// wgrad kernel: D[k (128 lanes), (tap,c) (BN columns)] += dY[pixels, k]^T * A[pixels, (tap,c)]
// ------------------------------------------------------------------------------------------------------------
template <int BN, int KP, int NSTAGES, bool BF16>
__global__ void __launch_bounds__(kThreadsIgemm, 1)
igemm_wgrad_kernel(const __grid_constant__ WgradParams P) {
  pdl_launch_dependents();  // (the matching pdl_wait() follows the prologue below)
  constexpr int kSlabCh = BF16 ? 64 : 32;            // channels per 128-byte-wide MN slab
  constexpr int kASlabs = kTileM / kSlabCh;          // 128 output channels
  constexpr int kMmaRows = BF16 ? 16 : 8;            // pixels (K) consumed per MMA
  static_assert(BN % kSlabCh == 0, "N tile must be whole slabs");
  constexpr uint32_t kSlabBytes = KP * 128;          // KP pixel rows x 128 bytes of channels
  constexpr uint32_t kABytes = kASlabs * kSlabBytes;
  constexpr uint32_t kBBytes = (BN / kSlabCh) * kSlabBytes;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static_assert(NSTAGES * kStageBytes >= 4 * 32 * kStagePitch * 4, "epilogue staging must fit in the pipeline smem");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[NSTAGES], empty_bar[NSTAGES], accum_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform (see fwd kernel)
  const int lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * kTileM;   // first output channel of the tile
  const int n0 = blockIdx.y * BN;       // first (tap,c) column of the tile
  const int split = blockIdx.z;
  const int step0 = split * P.steps_per_split;
  int steps = P.pixel_steps_total - step0;
  if (steps > P.steps_per_split) steps = P.steps_per_split;
  const int ncol = P.o.n_total - n0 < BN ? P.o.n_total - n0 : BN;  // valid columns (whole slabs)
  const int nslab = ncol / kSlabCh;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&P.tmDy);
    ptx::prefetch_tmap(&P.tmX);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NSTAGES; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      ptx::mbar_init(&accum_bar, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<kTmemCols>(&tmem_base_smem);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();  // (see igemm_fwd_persist_kernel)

  if (warp == 0) {
    {  // whole warp runs the loop; one elected lane issues (no divergence waterfall around the TMA instructions)
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = ((P.dbg & 2) ? 0 : kABytes) + ((P.dbg & 4) ? 0 : nslab * kSlabBytes);
      for (int s = 0; s < steps; ++s) {
        const int pix0 = (step0 + s) * KP;
        int j = pix0 % P.q_dim;
        int t = pix0 / P.q_dim;
        int i = t % P.p_dim;
        int n = t / P.p_dim;
        const int cw = P.base_w + j * P.trav_w, ch = P.base_h + i * P.trav_h;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * kStageBytes;
        if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&full_bar[stage], tx_bytes);
        if (P.rect) {
          // dY: all kASlabs channel slabs in one 3-D box (channel-in-slab, pixel, slab)
          if (!(P.dbg & 2)) ptx::tma_load_3d(sa, &P.tmDy, &full_bar[stage], 0, pix0, k0 / kSlabCh);
          // x: one 5-D box per filter tap covering every channel slab of that tap inside this N tile
          int sl = 0;
          while (sl < nslab && !(P.dbg & 4)) {
            const int col = n0 + sl * kSlabCh;
            const int tap = col / P.c;
            const int c0 = col - tap * P.c;
            int run = (P.c - c0) / kSlabCh;  // slabs left in this tap
            if (run > nslab - sl) run = nslab - sl;
            // run is baked into the tensor map's box (slabs per tap inside a tile is constant: see host code)
            ptx::tma_load_5d(sa + kABytes + sl * kSlabBytes, &P.tmX, &full_bar[stage], 0,
                             j * P.trav_w - P.pad_w + P.off_w[tap], i * P.trav_h - P.pad_h + P.off_h[tap], n,
                             c0 / kSlabCh);
            sl += run;
          }
        } else {
#pragma unroll
          for (int sl = 0; sl < kASlabs; ++sl)  // dY[pix0 .. pix0+KP, k0 + slab]  (rows past the tensor are zero-filled)
            ptx::tma_load_2d(sa + sl * kSlabBytes, &P.tmDy, &full_bar[stage], k0 + sl * kSlabCh, pix0);
          for (int sl = 0; sl < nslab; ++sl) {
            const int col = n0 + sl * kSlabCh;
            const int tap = col / P.c;
            const int c0 = col - tap * P.c;
            ptx::tma_load_im2col_4d(sa + kABytes + sl * kSlabBytes, &P.tmX, &full_bar[stage], c0, cw, ch, n,
                                    P.off_w[tap], P.off_h[tap]);
          }
        }
        }
        __syncwarp();
        if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = ptx::umma_idesc(BF16 ? 1 /*bf16*/ : 2 /*tf32*/, 1, 1, kTileM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < steps; ++s) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + stage * kStageBytes);
        const uint32_t sb = sa + kABytes;
        if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < KP / kMmaRows; ++k) {
          if (P.dbg & 1) break;  // 8 (tf32) / 16 (bf16) pixel rows of every slab per MMA
          // MN-major operands: fp32 must use the 128B-span / 32B-atom swizzle (4-row K groups 512 B apart), bf16 the
          // plain 128B swizzle (8-row K groups 1024 B apart); 128-byte-wide slabs are kSlabBytes apart
          uint64_t da = ptx::umma_desc(sa + k * (kMmaRows * 128), P.desc_lbo, P.desc_sbo, P.desc_layout);
          uint64_t db = ptx::umma_desc(sb + k * (kMmaRows * 128), P.desc_lbo, P.desc_sbo, P.desc_layout);
          if (BF16) ptx::mma_bf16(tmem_base, da, db, idesc, (s | k) != 0);
          else ptx::mma_tf32(tmem_base, da, db, idesc, (s | k) != 0);
        }
        ptx::mma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
      }
      if (ptx::elect_one()) ptx::mma_commit(&accum_bar);
      __syncwarp();
    }
  } else {
    if (lane == 0) ptx::mbar_wait(&accum_bar, 0);  // one poller per warp: 128 spinning threads slow every other mbarrier op of the SM
    __syncwarp();
    ptx::tc_fence_after();
    OutMap o = P.o;
    o.out = P.o.out + (int64_t)split * P.split_stride;
    EpiStats<BN> none;
    epilogue_tile<BN, false>(tmem_base, reinterpret_cast<float*>(smem) + (warp - 2) * (32 * kStagePitch), o.out,
                             out_row_offset(o, k0 + (warp & 3) * 32 + lane), n0, o.n_total, Epilogue{}, warp & 3, none);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------
// small helper kernels
// ------------------------------------------------------------------------------------------------------------
// w[K][T][C] -> wt[C][T][K]   (T = R*S taps)
template <class T_>
__global__ void repack_krsc_to_crsk_kernel(const T_* __restrict__ w, T_* __restrict__ wt, int K, int T, int C) {
  pdl_entry();
  int64_t total = (int64_t)K * T * C;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    int k = (int)(e % K);
    int64_t r = e / K;
    int t = (int)(r % T);
    int c = (int)(r / T);
    wt[e] = w[((int64_t)k * T + t) * C + c];
  }
}

__global__ void sum_splits_kernel(const float* __restrict__ partial, int splits, int64_t n, float* __restrict__ out) {
  pdl_entry();
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[(int64_t)s * n + e];
    out[e] = acc;
  }
}

void launch_sum_splits(const float* partial, int splits, int64_t n, float* out, cudaStream_t st) {
  launch_k(sum_splits_kernel, elementwise_grid(n, 256), 256, 0, st, partial, splits, n, out);
}

// Multi-tensor forms: one launch repacks the dgrad weights / sums the wgrad splits of up to kMaxBatch layers (a training
// step has ~20 of each, every one a few-microsecond latency-bound launch on its own).  Work is cut into units (tiles
// for the re-ordering, 1024-element chunks for the sums); start[i] = first unit of tensor i.
constexpr int kMaxBatch = 32;
constexpr int kBatchChunk = 1024;
struct RepackBatch {
  const float* w[kMaxBatch];
  float* wt[kMaxBatch];
  int K[kMaxBatch], T[kMaxBatch], C[kMaxBatch];
  int start[kMaxBatch + 1];
  int count;
};
struct SumBatch {
  const float* partial[kMaxBatch];
  float* out[kMaxBatch];
  int splits[kMaxBatch];
  int64_t n[kMaxBatch];
  int start[kMaxBatch + 1];
  int count;
};

// work unit = one 32 x 32 (k, c) tile of one filter tap, transposed through shared memory so that both the read
// (c contiguous in [K][T][C]) and the write (k contiguous in [C][T][K]) are coalesced; start[i] = first tile of tensor i
__global__ void __launch_bounds__(256) repack_multi_kernel(const __grid_constant__ RepackBatch B) {
  pdl_entry();
  __shared__ float tile[32][33];
  const int tiles = B.start[B.count];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int u = blockIdx.x; u < tiles; u += gridDim.x) {
    int i = 0;
    while (i + 1 < B.count && u >= B.start[i + 1]) ++i;
    const int K = B.K[i], T = B.T[i], C = B.C[i];
    const int kt = (K + 31) / 32, ct = (C + 31) / 32;
    int v = u - B.start[i];
    const int c0 = (v % ct) * 32;
    v /= ct;
    const int k0 = (v % kt) * 32;
    const int t = v / kt;
    const float* __restrict__ w = B.w[i];
    float* __restrict__ wt = B.wt[i];
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int k = k0 + r, c = c0 + tx;
      tile[r][tx] = (k < K && c < C) ? w[((int64_t)k * T + t) * C + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int c = c0 + r, k = k0 + tx;
      if (c < C && k < K) wt[((int64_t)c * T + t) * K + k] = tile[tx][r];
    }
    __syncthreads();
  }
}

// the same tiling, bf16 outputs: the straight copy (fprop's B operand, [K][T][C]) and the transposed copy (dgrad's,
// [C][T][K]) of one fp32 tile; either destination may be null
struct PackBf16Batch {
  const float* w[kMaxBatch];
  __nv_bfloat16* wh[kMaxBatch];
  __nv_bfloat16* wth[kMaxBatch];
  int K[kMaxBatch], T[kMaxBatch], C[kMaxBatch];
  int start[kMaxBatch + 1];
  int count;
};

__global__ void __launch_bounds__(256) pack_bf16_multi_kernel(const __grid_constant__ PackBf16Batch B) {
  pdl_entry();
  __shared__ float tile[32][33];
  const int tiles = B.start[B.count];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int u = blockIdx.x; u < tiles; u += gridDim.x) {
    int i = 0;
    while (i + 1 < B.count && u >= B.start[i + 1]) ++i;
    const int K = B.K[i], T = B.T[i], C = B.C[i];
    const int kt = (K + 31) / 32, ct = (C + 31) / 32;
    int v = u - B.start[i];
    const int c0 = (v % ct) * 32;
    v /= ct;
    const int k0 = (v % kt) * 32;
    const int t = v / kt;
    const float* __restrict__ w = B.w[i];
    __nv_bfloat16* __restrict__ wh = B.wh[i];
    __nv_bfloat16* __restrict__ wth = B.wth[i];
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int k = k0 + r, c = c0 + tx;
      float val = 0.f;
      if (k < K && c < C) {
        const int64_t e = ((int64_t)k * T + t) * C + c;
        val = w[e];
        if (wh) wh[e] = __float2bfloat16_rn(val);
      }
      tile[r][tx] = val;
    }
    __syncthreads();
    if (wth) {
#pragma unroll
      for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, k = k0 + tx;
        if (c < C && k < K) wth[((int64_t)c * T + t) * K + k] = __float2bfloat16_rn(tile[tx][r]);
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) sum_splits_multi_kernel(const __grid_constant__ SumBatch B) {
  pdl_entry();
  const int chunks = B.start[B.count];
  for (int ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
    int i = 0;
    while (i + 1 < B.count && ch >= B.start[i + 1]) ++i;
    const int64_t n = B.n[i];
    const int splits = B.splits[i];
    const float* __restrict__ partial = B.partial[i];
    float* __restrict__ out = B.out[i];
    const int64_t e0 = (int64_t)(ch - B.start[i]) * kBatchChunk;
    for (int64_t e = e0 + threadIdx.x; e < e0 + kBatchChunk && e < n; e += blockDim.x) {
      float acc = 0.f;
      for (int sp = 0; sp < splits; ++sp) acc += partial[(int64_t)sp * n + e];  // fixed order: deterministic
      out[e] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// host: support checks, tile selection, launches
// ------------------------------------------------------------------------------------------------------------
static bool in_corner_range(int v) { return v >= -128 && v <= 127; }

static bool common_ok(const ttb_conv_desc* d) {
  if (d->math_mode != TTB_MATH_TF32 && d->math_mode != TTB_MATH_BF16) return false;
  if (d->groups != 1) return false;
  if (d->r * d->s > kMaxTaps) return false;
  if (d->n <= 0 || d->p <= 0 || d->q <= 0) return false;
  if ((int64_t)d->n * d->p * d->q >= (1ll << 31) || (int64_t)d->n * d->h * d->w >= (1ll << 31)) return false;
  // (the epilogue addresses output rows by 32-bit offsets in float4 units)
  if ((int64_t)d->n * d->p * d->q * d->k >= (1ll << 33) || (int64_t)d->n * d->h * d->w * d->c >= (1ll << 33)) return false;
  if (d->stride_h > 8 || d->stride_w > 8) return false;
  if ((d->r - 1) * d->dil_h > 255 || (d->s - 1) * d->dil_w > 255) return false;
  return true;
}

// channels per K-block / MN slab in the math mode of the descriptor (32 for TF32, 64 for BF16)
int igemm_channel_block(const ttb_conv_desc* d) { return d->math_mode == TTB_MATH_BF16 ? 64 : 32; }

bool igemm_supported(const ttb_conv_desc* d, int pass) {
  if (!common_ok(d)) return false;
  const int blk = igemm_channel_block(d);
  const int up_h = d->pad_h - (d->r - 1) * d->dil_h, up_w = d->pad_w - (d->s - 1) * d->dil_w;
  if (pass == 0 || pass == 2) {
    // the traversal box must reproduce exactly (P, Q): true for the reference's floor formula when the
    // bottom/right remainder is smaller than the stride
    if (d->c % blk != 0 || d->k % 8 != 0) return false;
    if (!in_corner_range(-d->pad_h) || !in_corner_range(-d->pad_w) || !in_corner_range(up_h) || !in_corner_range(up_w))
      return false;
    if ((d->h + up_h + d->pad_h - 1) / d->stride_h + 1 != d->p) return false;
    if ((d->w + up_w + d->pad_w - 1) / d->stride_w + 1 != d->q) return false;
    if (d->h + up_h + d->pad_h < 1 || d->w + up_w + d->pad_w < 1) return false;
    return true;
  }
  // dgrad: reduction over output channels k, output columns = input channels c
  if (d->k % blk != 0 || d->c % 8 != 0) return false;
  if (d->pad_h + (d->r - 1) * d->dil_h > 100 || d->pad_w + (d->s - 1) * d->dil_w > 100) return false;
  return true;
}

// Widest N tile that still gives every SM a CTA; narrow channel counts get a matching narrow tile.
static int pick_bn_wide(int64_t m_total, int n_total) {
  const int64_t mtiles = ceil_div(m_total, kTileM);
  const int sms = sm_count();
  // 256-wide tiles are the only tensor-bound TF32 configuration (see launch_persist_bn): take them as soon as they
  // fill most of one wave of SMs
  if (n_total > 128 && mtiles * ceil_div(n_total, 256) * 5 >= sms * 4) return 256;
  if (n_total > 64 && (mtiles * ceil_div(n_total, 128) >= sms || n_total > 128)) return 128;
  if (n_total > 32) return 64;
  return 32;
}

// `nkb`: K-blocks (stage fills of one tap) a tile's main loop runs = taps x reduction channels / 32 (64 in bf16).  A SHALLOW
// reduction - the 1 x 1 convolutions of a bottleneck network, the 1 x 1 stride-2 shortcuts, the few-tap classes of a
// strided dgrad - spends 4 nkb MMAs (<= 2-4 k cycles) per tile against 2.4 / 4.5 / 9 k cycles of epilogue for 64 / 128 / 256
// columns on the CTA's four epilogue warps: the kernel is bound by the epilogue, not by the tensor pipe.  Narrow tiles
// run TWO CTAs per SM, i.e. eight epilogue warps storing, and halve the epilogue of a tile.  Measured on B200
// (profiles/r2_force_bn_*.txt, per-layer device time of a step): ResNet-50 1x1 layers with <= 8 K-blocks 1.3-1.6x faster on
// 64-wide tiles (64->256 at 56x56: 410 -> 289 us, the 256->64 dgrad with its pending gradient: 601 -> 403 us), 16-32
// K-blocks 1.1-1.4x on 128-wide instead of 256-wide tiles; the 3 x 3 layers (>= 18 K-blocks) keep the wide tiles.
static int pick_bn(int64_t m_total, int n_total, int nkb) {
  if (const int v = tuning_knob("TTB_FORCE_BN", 0)) {  // experiment switch (tuning build only)
    if (v == 256 && n_total > 128) return 256;
    if (v >= 128 && n_total > 64) return 128;
    if (v >= 64 && n_total > 32) return 64;
    if (v == 32) return 32;
  }
  int bn = pick_bn_wide(m_total, n_total);
  if (tuning_knob("TTB_SHALLOW_BN", 1)) {
    if (nkb <= 8 && bn > 64) bn = 64;
    else if (nkb <= 32 && bn > 128) bn = 128;
  }
  return bn;
}

constexpr size_t persist_smem_bytes(int bn, int nstages, int wt) {
  return (size_t)nstages * ((((kTileM + wt - 1) * 128 + 1023) & ~1023) + wt * bn * 128) + kEpilogueStagingBytes + 1024;
}

// Taps per pipeline stage along W (the kernel's WT): 3 = the three taps of a filter row share one A strip (3-wide filters,
// unit stride and dilation in W, 32 / 64 / 128-wide tiles - wider tiles are tensor-bound already and have no room for three
// weight tiles per stage; rows shorter than 8 outputs would waste > 25 % of the MMAs on the dropped positions), else 1.
static int pick_wt(int taps_w, int stride_w, int dil_w, int bn, int q_out) {
  if (taps_w != 3 || stride_w != 1 || dil_w != 1 || (bn != 32 && bn != 64 && bn != 128) || q_out < 8) return 1;
  return tuning_knob("TTB_WT", 3) == 3 ? 3 : 1;
}

// CTAs of a persistent launch: one per SM (two for the narrow tiles whose ring is sized for it), never more than tiles;
// with epilogue statistics a multiple of the N-tile count, so that a CTA keeps one N tile (see the kernel)
static unsigned persist_grid(int64_t tiles, int nt, size_t smem, bool stats) {
  int sms = sm_count() * (smem <= 113 * 1024 ? 2 : 1);
  sms = tuning_knob("TTB_PERSIST_GRID", sms);  // experiment switch (tuning build only): cap the number of CTAs
  int64_t grid = tiles < sms ? tiles : sms;
  if (stats && nt > 1) grid = grid >= nt ? grid / nt * nt : nt;  // (tiles is a multiple of nt)
  return (unsigned)grid;
}

template <int BN, int NSTAGES, int NPROD, int WT, bool BF16, int STATS>
static int launch_persist(const FwdParamsMulti& PM, int count, cudaStream_t st) {
  constexpr size_t smem = persist_smem_bytes(BN, NSTAGES, WT);
  static_assert(smem <= 232448, "shared memory budget of one SM");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(igemm_fwd_persist_kernel<BN, NSTAGES, NPROD, WT, BF16, STATS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("igemm: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return 1;
    }
    attr_set = true;
  }
  const int nt = (int)ceil_div(PM.p[0].o.n_total, BN);
  int64_t tiles = 0;
  for (int i = 0; i < count; ++i) tiles += (int64_t)ceil_div(PM.p[i].o.m_total, kTileM) * nt;
  // narrow tiles leave room for two CTAs per SM (two independent MMA-issue streams: a 64-column K-block is 128
  // tensor-core cycles but ~300 cycles of issue-side latency per CTA)
  const unsigned grid = persist_grid(tiles, nt, smem, STATS != 0);
  launch_k(igemm_fwd_persist_kernel<BN, NSTAGES, NPROD, WT, BF16, STATS>, grid, (NPROD + 5) * 32, smem, st, PM, count);
  return check_launch("igemm_fwd_persist_kernel");
}

// <BN, stages, producer warps, taps per stage>.  Measured on B200 (TF32): the MMA warp pays ~170 cycles per
// stage hand-shake + ~30 cycles per MMA issue against 32 / 64 / 128 tensor-core cycles per MMA at N = 64 / 128 / 256,
// so N = 256 is tensor-bound (512 cycles per K-block), N = 128 nearly (300 vs 256) and N <= 64 issue-bound: narrow
// tiles run two CTAs per SM (two independent issue streams; their ring is sized to let two fit).
struct PersistCfg { int nstages, nprod; };
static PersistCfg persist_cfg(int bn, int wt) {
  if (wt == 3) return bn == 64 ? PersistCfg{2, 2} : PersistCfg{3, 3};
  switch (bn) {
    case 256: return {4, 2};
    case 128: return {6, 3};
    default: return {3, 3};
  }
}

template <bool BF16, int STATS>
static int launch_persist_sel(const FwdParamsMulti& PM, int count, int bn, int wt, cudaStream_t st) {
  if (wt == 3) {
    if (bn == 128) return launch_persist<128, 3, 3, 3, BF16, STATS>(PM, count, st);
    if (bn == 32) return launch_persist<32, 3, 3, 3, BF16, STATS>(PM, count, st);  // (32-filter layers: UNet level 1)
    return launch_persist<64, 2, 2, 3, BF16, STATS>(PM, count, st);
  }
  switch (bn) {
    case 256: return launch_persist<256, 4, 2, 1, BF16, STATS>(PM, count, st);
    case 128: return launch_persist<128, 6, 3, 1, BF16, STATS>(PM, count, st);
    case 64: return launch_persist<64, 3, 3, 1, BF16, STATS>(PM, count, st);
    default: return launch_persist<32, 3, 3, 1, BF16, STATS>(PM, count, st);
  }
}

static int launch_persist_bn(const FwdParamsMulti& PM, int count, int bn, int wt, bool bf16, cudaStream_t st) {
  // 0 none, 1 forward statistics of the stored values, 2 BatchNorm-backward sums against Epilogue::bn_x (conv_epilogue.cuh)
  const int stats = (count == 1 && PM.p[0].ep.stats != nullptr) ? (PM.p[0].ep.bn_x != nullptr ? 2 : 1) : 0;
  if (bf16) {
    if (stats == 2) return launch_persist_sel<true, 2>(PM, count, bn, wt, st);
    return stats ? launch_persist_sel<true, 1>(PM, count, bn, wt, st) : launch_persist_sel<true, 0>(PM, count, bn, wt, st);
  }
  if (stats == 2) return launch_persist_sel<false, 2>(PM, count, bn, wt, st);
  return stats ? launch_persist_sel<false, 1>(PM, count, bn, wt, st) : launch_persist_sel<false, 0>(PM, count, bn, wt, st);
}

// rows of the [chunks][2][K] statistics partial buffer an fprop launch of this problem writes (Epilogue::stats)
int igemm_fprop_stats_chunks(const ttb_conv_desc* d) {
  const int bn = pick_bn((int64_t)d->n * d->p * d->q, d->k, d->r * d->s * (d->c / igemm_channel_block(d)));
  // 256-wide tiles (the deep layers: about one tile per CTA, so the epilogue is not hidden behind a next tile's main loop)
  // measured +5.8 us with statistics against a ~5 us statistics pass over their small outputs: no statistics there -
  // unless the output is large (the 1 x 1 expansions of ResNet-50: 100-800 MB, several tiles per CTA), where the separate
  // pass is a full HBM read of the tensor (38 of that network's 53 BatchNorms ran one: 3.8 ms of a 56 ms step)
  if (bn == 256 && (int64_t)d->n * d->p * d->q * d->k < tuning_knob("TTB_STATS256_MIN_MELEMS", 8) * (int64_t)(1 << 20)) return 0;
  const int wt = pick_wt(d->s, d->stride_w, d->dil_w, bn, d->q);
  const int64_t m_total = (int64_t)d->n * d->p * (d->q + wt - 1);
  const PersistCfg c = persist_cfg(bn, wt);
  const int nt = (int)ceil_div(d->k, bn);
  const int64_t tiles = ceil_div(m_total, kTileM) * nt;
  return (int)(persist_grid(tiles, nt, persist_smem_bytes(bn, c.nstages, wt), true) / nt);
}

// rows of the [chunks][2][C] partial buffer a stride-1 dgrad of this problem writes when its epilogue emits the sums of a
// BatchNorm backward over dx (Epilogue::bn_x); 0 = not available.  Mirrors igemm_dgrad's single-class tile choice.
int igemm_dgrad_stats_chunks(const ttb_conv_desc* d) {
  if (d->stride_h != 1 || d->stride_w != 1 || d->groups != 1 || !igemm_supported(d, 1)) return 0;
  const int nkb = d->r * d->s * (d->k / igemm_channel_block(d));
  const int bn = pick_bn((int64_t)d->n * d->h * d->w, d->c, nkb);
  // (256-wide tiles on small outputs: about one tile per CTA, the longer epilogue would not be hidden - as in fprop)
  if (bn == 256 && (int64_t)d->n * d->h * d->w * d->c < tuning_knob("TTB_STATS256_MIN_MELEMS", 8) * (int64_t)(1 << 20)) return 0;
  const int wt = pick_wt(d->s, 1, d->dil_w, bn, d->w);
  const int64_t m_total = (int64_t)d->n * d->h * (d->w + wt - 1);
  const PersistCfg c = persist_cfg(bn, wt);
  const int nt = (int)ceil_div(d->c, bn);
  const int64_t tiles = ceil_div(m_total, kTileM) * nt;
  return (int)(persist_grid(tiles, nt, persist_smem_bytes(bn, c.nstages, wt), true) / nt);
}

static int wgrad_variant() {  // bring-up / tuning knob (tuning build only): 0 = default
  static const int v = tuning_knob("TTB_WGRAD_VARIANT", 0);
  return v;
}

static int wgrad_plan(const ttb_conv_desc* d, int* bn, int* splits, int* steps_per_split, int* steps_total) {
  // 64 pixels per pipeline step: the TMA unit pays a fixed cost per box (measured: ~115 cycles + ~3 cycles per
  // 128-byte row), so few large boxes beat many small ones (KP = 32 ran the layer-1 wgrad at 89 TFLOP/s, KP = 64 at 162)
  const int variant = wgrad_variant();
  const bool bf16 = d->math_mode == TTB_MATH_BF16;
  const int KP = (variant == 3 && !bf16) ? 32 : 64;
  const int ncols = d->r * d->s * d->c;
  *bn = ncols >= 256 ? 256 : (ncols >= 128 ? 128 : (ncols >= 64 ? 64 : 32));
  if (variant == 2 && *bn == 256) *bn = 128;
  const int64_t m = (int64_t)d->n * d->p * d->q;
  const int total = (int)ceil_div(m, KP);
  const int64_t tiles = ceil_div(d->k, kTileM) * ceil_div(ncols, *bn);
  // one CTA per SM is resident (the stage ring takes most of the shared memory), so size the grid to whole waves:
  // the largest split count with tiles*splits <= waves*SMs, for the smallest wave count that gives >= 8 steps per CTA
  const int64_t sms = sm_count();
  int64_t sp = 1;
  for (int waves = 1; waves <= 2; ++waves) {
    int64_t cand = (waves * sms) / tiles;
    if (cand < 1) cand = 1;
    if (cand > total / 8) cand = total / 8;
    if (cand < 1) cand = 1;
    sp = cand;
    if (tiles * cand >= (waves * sms * 3) / 4 || tiles > sms) break;  // this wave count is filled well enough
  }
  if (sp > 512) sp = 512;
  int sps = (int)ceil_div(total, sp);
  sp = ceil_div(total, sps);
  *splits = (int)sp;
  *steps_per_split = sps;
  *steps_total = total;
  return KP;
}

// workspace of the igemm passes themselves (operands already in the element type of the math mode)
size_t igemm_workspace_size(const ttb_conv_desc* d, int pass) {
  const size_t welems = (size_t)d->k * d->r * d->s * d->c;
  if (pass == 0) return 0;
  if (pass == 1) return welems * (d->math_mode == TTB_MATH_BF16 ? 2 : 4);  // repacked weights [C][R][S][K]
  int bn, splits, sps, total;
  wgrad_plan(d, &bn, &splits, &sps, &total);
  const size_t im2col = splits > 1 ? (size_t)splits * welems * sizeof(float) : 0;
  // (a group of a grouped convolution always takes the im2col kernel: the larger of the two plans)
  const size_t halo = halo_wgrad_selected(d) ? halo_wgrad_workspace(d) : 0;
  return halo > im2col ? halo : im2col;
}

static long long* g_trace = nullptr;
void igemm_set_trace(long long* p) { g_trace = p; }

static int igemm_dbg() {  // timing experiments that produce WRONG results: tuning build only, 0 in the release library
  static const int v = tuning_knob("TTB_IGEMM_DBG", 0);
  return v;
}

// One fprop problem -> kernel parameters.  x, w: operands in the element type of d->math_mode (fp32 for TF32, bf16 for BF16);
// y, bias: fp32.  x_ctot / y_ctot: channel count of the tensors x / y live in when `d` describes ONE GROUP of a grouped
// convolution (x, y, ep.* then point at the group's first channel); 0 = dense.
// wt > 1 (row-shared A strips, see the kernel): the traversal box is widened by wt - 1 positions in W - output rows are
// indexed over Q + wt - 1 positions, the last wt - 1 dropped - and one load brings kTileM + wt - 1 positions.
static int fprop_params(FwdParams& P, const ttb_conv_desc* d, const void* x, const void* w, const Epilogue& ep, float* y, int bn,
                        int wt, int x_ctot, int y_ctot) {
  const Elem el = elem_of(d->math_mode == TTB_MATH_BF16);
  memset(&P, 0, sizeof(P));
  const int up_h = d->pad_h - (d->r - 1) * d->dil_h, up_w = d->pad_w - (d->s - 1) * d->dil_w + (wt - 1);
  if (make_im2col_4d(&P.tmA, x, el, d->n, d->h, d->w, d->c, -d->pad_w, -d->pad_h, up_w, up_h, d->stride_w, d->stride_h,
                     kTileM + wt - 1, CU_TENSOR_MAP_SWIZZLE_128B, x_ctot))
    return 1;
  const int64_t yk = y_ctot ? y_ctot : d->k;
  P.o.out = y;
  P.o.n_stride = (int64_t)d->p * d->q * yk;
  P.o.h_stride = (int64_t)d->q * yk;
  P.o.w_stride = yk;
  P.o.base = 0;
  P.o.p_dim = d->p;
  P.o.q_dim = d->q + wt - 1;
  P.o.q_valid = wt > 1 ? d->q : 0;
  P.o.m_total = d->n * d->p * (d->q + wt - 1);
  P.o.n_total = d->k;
  P.ep = ep;
  P.ep.relu = (ep.relu ? 1 : 0) | (tuning_knob("TTB_EPI_DBG", 0) << 8);  // (experiment bits: tuning build only)
  P.c_blocks = d->c / el.per_row;
  P.num_taps = d->r * d->s;
  P.base_w = -d->pad_w;
  P.base_h = -d->pad_h;
  P.dbg = igemm_dbg();
  P.trace = g_trace;
  P.trav_w = d->stride_w;
  P.trav_h = d->stride_h;
  for (int r = 0; r < d->r; ++r)
    for (int s = 0; s < d->s; ++s) {
      int t = r * d->s + s;
      P.b_koff[t] = t * d->c;
      P.off_w[t] = (uint16_t)(s * d->dil_w);
      P.off_h[t] = (uint16_t)(r * d->dil_h);
    }
  // weight matrix [K rows][R*S*C cols]; the TMA box height is the kernel's N tile
  return make_tiled_2d(&P.tmB, w, el, (uint64_t)d->k, (uint64_t)d->r * d->s * d->c, (uint32_t)bn);
}

int igemm_fprop(const ttb_conv_desc* d, const void* x, const void* w, const Epilogue& ep, float* y, void* /*ws*/,
                size_t /*ws_bytes*/, cudaStream_t st) {
  if (load_driver_fns()) return 1;
  static thread_local FwdParamsMulti PM1;
  const int bn = pick_bn((int64_t)d->n * d->p * d->q, d->k, d->r * d->s * (d->c / igemm_channel_block(d)));
  const int wt = pick_wt(d->s, d->stride_w, d->dil_w, bn, d->q);
  if (fprop_params(PM1.p[0], d, x, w, ep, y, bn, wt, 0, 0)) return 1;
  return launch_persist_bn(PM1, 1, bn, wt, d->math_mode == TTB_MATH_BF16, st);
}

// Grouped convolution: `dg` describes one group (c = C/groups, k = K/groups, groups = 1); group g reads x + g*x_goff (a
// channel slice of a tensor with x_ctot channels per pixel, or its own dense buffer when x_ctot == 0) and w + g*w_goff and
// writes channels [g*k, (g+1)*k) of y (y_ctot channels per pixel).  Up to kMaxMulti groups share one persistent launch.
int igemm_fprop_grouped(const ttb_conv_desc* dg, int groups, const void* x, size_t x_goff_bytes, int x_ctot, const void* w,
                        size_t w_goff_bytes, const Epilogue& ep, float* y, int y_ctot, cudaStream_t st) {
  if (load_driver_fns()) return 1;
  static thread_local FwdParamsMulti PM;
  const int bn = pick_bn((int64_t)dg->n * dg->p * dg->q * (groups < kMaxMulti ? groups : kMaxMulti), dg->k,
                         dg->r * dg->s * (dg->c / igemm_channel_block(dg)));
  const int wt = pick_wt(dg->s, dg->stride_w, dg->dil_w, bn, dg->q);
  for (int g0 = 0; g0 < groups; g0 += kMaxMulti) {
    const int cnt = groups - g0 < kMaxMulti ? groups - g0 : kMaxMulti;
    for (int i = 0; i < cnt; ++i) {
      const int g = g0 + i;
      Epilogue e = ep;
      e.stats = nullptr;
      if (e.scale) e.scale += (size_t)g * dg->k;
      if (e.bias) e.bias += (size_t)g * dg->k;
      if (e.accum) e.accum += (size_t)g * dg->k;
      if (fprop_params(PM.p[i], dg, reinterpret_cast<const char*>(x) + g * x_goff_bytes,
                       reinterpret_cast<const char*>(w) + g * w_goff_bytes, e, y + (size_t)g * dg->k, bn, wt, x_ctot, y_ctot))
        return 1;
    }
    if (launch_persist_bn(PM, cnt, bn, wt, dg->math_mode == TTB_MATH_BF16, st)) return 1;
  }
  return 0;
}

// `prepacked` != null: the weights are already in the [C][R][S][K] order (igemm_pack_dgrad_weights), w / ws unused
// dy_ctot / dx_ctot: channel counts of the tensors dy / dx live in when `d` is ONE GROUP of a grouped convolution (dy, dx,
// accum point at the group's first channel; the caller zeroes dx where no tap reaches: zero_done); 0 = dense.
// `bn` (may be null; stride-1 problems only, see igemm_dgrad_stats_chunks): stats + bn_* fields of an Epilogue - the epilogue
// also emits the BatchNorm-backward sums of dx against the BatchNorm input bn->bn_x.
int igemm_dgrad(const ttb_conv_desc* d, const void* dy, const void* w, float* dx, void* ws, size_t ws_bytes,
                cudaStream_t st, const void* prepacked, const float* accum, int dy_ctot, int dx_ctot, bool zero_done,
                const Epilogue* bn) {
  // (`accum`, may be null: added to dx in the epilogue - the gradient already pending for the same tensor)
  const int64_t xc = dx_ctot ? dx_ctot : d->c;
  if (load_driver_fns()) return 1;
  const Elem el = elem_of(d->math_mode == TTB_MATH_BF16);
  const size_t wbytes = (size_t)d->k * d->r * d->s * d->c * el.size;
  if (!prepacked)
    TTB_REQUIRE(ws != nullptr && ws_bytes >= wbytes, "conv2d_dgrad: workspace of %zu bytes needed, %zu given", wbytes, ws_bytes);
  const int T = d->r * d->s;
  if (!prepacked) {
    int64_t total = (int64_t)d->k * T * d->c;
    if (el.bf16)
      launch_k(repack_krsc_to_crsk_kernel<uint16_t>, elementwise_grid(total, 256), 256, 0, st, 
          reinterpret_cast<const uint16_t*>(w), reinterpret_cast<uint16_t*>(ws), d->k, T, d->c);
    else
      launch_k(repack_krsc_to_crsk_kernel<float>, elementwise_grid(total, 256), 256, 0, st, 
          reinterpret_cast<const float*>(w), reinterpret_cast<float*>(ws), d->k, T, d->c);
    if (check_launch("repack_krsc_to_crsk")) return 1;
  }
  const void* wt = prepacked ? prepacked : ws;
  // Stride-parity classes: input rows h = a + sh*i only receive taps r with (a + ph - r*dh) % sh == 0, from output
  // row p = i + (a + ph - r*dh)/sh.  Each class is a stride-1 gather over dY - no zero insertion, no wasted MACs.
  bool need_zero = false;
  struct Cls { int a, b, nr, ns; int rr[kMaxTaps], ro[kMaxTaps], ss[kMaxTaps], so[kMaxTaps]; };
  static thread_local Cls cls;
  static thread_local FwdParamsMulti PM;
  int n_multi = 0, nkb_max = 0;
  int64_t m_all = 0;
  const bool multi = d->stride_h * d->stride_w > 1 && d->stride_h * d->stride_w <= kMaxMulti;
  TTB_REQUIRE(!bn || (d->stride_h == 1 && d->stride_w == 1 && !dy_ctot && !dx_ctot),
              "conv2d_dgrad: BatchNorm-backward statistics need a dense stride-1 problem");
  for (int pass = 0; pass < 2; ++pass) {
    for (int a = 0; a < d->stride_h; ++a)
      for (int b = 0; b < d->stride_w; ++b) {
        const int ha = (d->h - a + d->stride_h - 1) / d->stride_h, wb = (d->w - b + d->stride_w - 1) / d->stride_w;
        if (ha <= 0 || wb <= 0) continue;
        cls.nr = cls.ns = 0;
        for (int r = 0; r < d->r; ++r) {
          int t = a + d->pad_h - r * d->dil_h;
          if (((t % d->stride_h) + d->stride_h) % d->stride_h) continue;
          cls.rr[cls.nr] = r;
          cls.ro[cls.nr++] = (t - (((t % d->stride_h) + d->stride_h) % d->stride_h)) / d->stride_h;
        }
        for (int s = 0; s < d->s; ++s) {
          int t = b + d->pad_w - s * d->dil_w;
          if (((t % d->stride_w) + d->stride_w) % d->stride_w) continue;
          cls.ss[cls.ns] = s;
          cls.so[cls.ns++] = (t - (((t % d->stride_w) + d->stride_w) % d->stride_w)) / d->stride_w;
        }
        if (cls.nr == 0 || cls.ns == 0) {
          need_zero = true;
          continue;
        }
        if (pass == 0) continue;
        int lo_h = cls.ro[0], lo_w = cls.so[0];
        for (int i = 0; i < cls.nr; ++i) lo_h = cls.ro[i] < lo_h ? cls.ro[i] : lo_h;
        for (int i = 0; i < cls.ns; ++i) lo_w = cls.so[i] < lo_w ? cls.so[i] : lo_w;
        // (a stride-1 dgrad is one class: its N tile is known here, and with it whether the three taps of a filter row
        // share one strip of dY - see pick_wt / the kernel's WT)
        const int nkb = cls.nr * cls.ns * (d->k / el.per_row);
        if (nkb > nkb_max) nkb_max = nkb;
        const int bn1 = multi ? 0 : pick_bn((int64_t)d->n * ha * wb, d->c, nkb);
        const int wtaps = multi ? 1 : pick_wt(cls.ns, d->stride_w, d->dil_w, bn1, wb);
        const int up_h = ha - d->p + lo_h, up_w = wb - d->q + lo_w + (wtaps - 1);
        TTB_REQUIRE(in_corner_range(lo_h) && in_corner_range(lo_w) && in_corner_range(up_h) && in_corner_range(up_w),
                    "conv2d_dgrad: traversal box out of TMA range");
        FwdParams Pone;
        FwdParams& P = multi ? PM.p[n_multi] : Pone;
        memset(&P, 0, sizeof(P));
        if (make_im2col_4d(&P.tmA, dy, el, d->n, d->p, d->q, d->k, lo_w, lo_h, up_w, up_h, 1, 1, kTileM + wtaps - 1,
                           CU_TENSOR_MAP_SWIZZLE_128B, dy_ctot))
          return 1;
        P.o.out = dx;
        P.o.n_stride = (int64_t)d->h * d->w * xc;
        P.o.h_stride = (int64_t)d->stride_h * d->w * xc;
        P.o.w_stride = (int64_t)d->stride_w * xc;
        P.o.base = ((int64_t)a * d->w + b) * xc;
        P.o.p_dim = ha;
        P.o.q_dim = wb + wtaps - 1;
        P.o.q_valid = wtaps > 1 ? wb : 0;
        P.o.m_total = d->n * ha * (wb + wtaps - 1);
        P.o.n_total = d->c;
        P.ep = Epilogue{nullptr, nullptr, accum, tuning_knob("TTB_EPI_DBG", 0) << 8, nullptr};
        if (bn) {
          P.ep.stats = bn->stats;
          P.ep.bn_x = bn->bn_x; P.ep.bn_mean = bn->bn_mean; P.ep.bn_rscale = bn->bn_rscale; P.ep.bn_rshift = bn->bn_rshift;
        }
        P.c_blocks = d->k / el.per_row;
        P.num_taps = cls.nr * cls.ns;
        P.base_w = lo_w;
        P.base_h = lo_h;
        P.trav_w = P.trav_h = 1;
        for (int i = 0; i < cls.nr; ++i)
          for (int j = 0; j < cls.ns; ++j) {
            int t = i * cls.ns + j;
            P.b_koff[t] = (cls.rr[i] * d->s + cls.ss[j]) * d->k;
            P.off_h[t] = (uint16_t)(cls.ro[i] - lo_h);
            P.off_w[t] = (uint16_t)(cls.so[j] - lo_w);
          }
        if (multi) {  // launched together below; the N tile is chosen for the combined grid
          m_all += P.o.m_total;
          ++n_multi;
          continue;
        }
        if (make_tiled_2d(&P.tmB, wt, el, (uint64_t)d->c, (uint64_t)T * d->k, (uint32_t)bn1)) return 1;
        static thread_local FwdParamsMulti PM1;
        PM1.p[0] = P;
        if (launch_persist_bn(PM1, 1, bn1, wtaps, el.bf16, st)) return 1;
      }
    if (pass == 1 && n_multi > 0) {
      const int bn = pick_bn(m_all, d->c, nkb_max);  // (the deepest class decides: the others only get shorter)
      for (int i = 0; i < n_multi; ++i)
        if (make_tiled_2d(&PM.p[i].tmB, wt, el, (uint64_t)d->c, (uint64_t)T * d->k, (uint32_t)bn)) return 1;
      if (launch_persist_bn(PM, n_multi, bn, 1, el.bf16, st)) return 1;
    }
    if (pass == 0 && need_zero && !zero_done) {  // input pixels no filter tap reaches: zero (or just the pending gradient)
      const size_t bytes = (size_t)d->n * d->h * d->w * d->c * sizeof(float);
      cudaError_t e = accum ? cudaMemcpyAsync(dx, accum, bytes, cudaMemcpyDeviceToDevice, st) : cudaMemsetAsync(dx, 0, bytes, st);
      if (e != cudaSuccess) {
        set_error("conv2d_dgrad: memset failed: %s", cudaGetErrorString(e));
        return 1;
      }
    }
  }
  return 0;
}

template <int BN, int KP, int NSTAGES, bool BF16>
static int launch_wgrad(const WgradParams& P, int ktiles, int ntiles, int splits, cudaStream_t st) {
  constexpr int kSlabCh = BF16 ? 64 : 32;
  constexpr size_t smem = (size_t)NSTAGES * ((kTileM / kSlabCh + BN / kSlabCh) * KP * 128) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(igemm_wgrad_kernel<BN, KP, NSTAGES, BF16>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("igemm wgrad: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return 1;
    }
    attr_set = true;
  }
  dim3 grid(ktiles, ntiles, splits);
  launch_k(igemm_wgrad_kernel<BN, KP, NSTAGES, BF16>, grid, kThreadsIgemm, smem, st, P);
  return check_launch("igemm_wgrad_kernel");
}

// `splits_out` != null: the caller sums the splits (igemm_sum_splits_multi); *splits_out = number of partial buffers
// [K*R*S*C] at the start of ws (<= 1: dw is already final)
// x_ctot / dy_ctot: channel counts of the tensors x / dy live in when `d` is ONE GROUP of a grouped convolution (x, dy point
// at the group's first channel, dw at the group's filters); 0 = dense.
int igemm_wgrad(const ttb_conv_desc* d, const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes,
                cudaStream_t st, int* splits_out, int x_ctot, int dy_ctot) {
  if (load_driver_fns()) return 1;
  if (x_ctot == 0 && dy_ctot == 0 && halo_wgrad_selected(d)) return halo_wgrad(d, x, dy, dw, ws, ws_bytes, st, splits_out);
  const uint64_t xc = x_ctot ? x_ctot : d->c, yk = dy_ctot ? dy_ctot : d->k;
  const Elem el = elem_of(d->math_mode == TTB_MATH_BF16);
  int bn, splits, sps, total;
  const int KP = wgrad_plan(d, &bn, &splits, &sps, &total);
  const int64_t wsize = (int64_t)d->k * d->r * d->s * d->c;
  if (splits > 1)
    TTB_REQUIRE(ws != nullptr && ws_bytes >= (size_t)splits * wsize * sizeof(float),
                "conv2d_wgrad: workspace of %zu bytes needed, %zu given", (size_t)splits * wsize * sizeof(float), ws_bytes);
  WgradParams P;
  memset(&P, 0, sizeof(P));
  const int64_t m = (int64_t)d->n * d->p * d->q;
  // MN-major operands.  fp32: 128B-span / 32B-atom swizzle (TMA) <-> UMMA layout type 1, 4-row K groups 512 B apart.
  // bf16: plain 128B swizzle <-> layout type 2, 8-row K groups 1024 B apart.  128-byte-wide slabs KP*128 B apart.
  // (TTB_WGRAD_* overrides exist in the tuning build only.)
  CUtensorMapSwizzle swz = el.bf16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  P.desc_lbo = (uint32_t)KP * 128;
  P.desc_sbo = el.bf16 ? 1024 : 512;
  P.desc_layout = el.bf16 ? 2 : 1;
  swz = (CUtensorMapSwizzle)tuning_knob("TTB_WGRAD_SWIZZLE", (int)swz);
  P.desc_lbo = (uint32_t)tuning_knob("TTB_WGRAD_LBO", (int)P.desc_lbo);
  P.desc_sbo = (uint32_t)tuning_knob("TTB_WGRAD_SBO", (int)P.desc_sbo);
  P.desc_layout = (uint32_t)tuning_knob("TTB_WGRAD_LAYOUT", (int)P.desc_layout);
  const int ncols = d->r * d->s * d->c;
  // Rectangular fast path: when the KP output pixels of a step are a box of the (n, p, q) grid and taps / N tiles
  // nest, every tap's channel slabs arrive in ONE 5-D tiled load and dY's slabs in ONE 3-D load.
  const int slab = el.per_row;
  int bw = 0, bh = 0, bnimg = 0;
  {
    bool ok = tuning_knob("TTB_WGRAD_RECT", 1) != 0 && d->k % slab == 0 && (d->c % bn == 0 || bn % d->c == 0) && m % KP == 0;
    if (ok) {
      if (d->q >= KP) {
        ok = d->q % KP == 0;
        bw = KP; bh = 1; bnimg = 1;
      } else if (KP % d->q == 0) {
        const int rows = KP / d->q;
        if (d->p % rows == 0) { bw = d->q; bh = rows; bnimg = 1; }
        else if (rows % d->p == 0) { bw = d->q; bh = d->p; bnimg = rows / d->p; }
        else ok = false;
      } else {
        ok = false;
      }
    }
    if (ok && (bw * d->stride_w > 256 || bh * d->stride_h > 256 || bnimg > 256)) ok = false;
    P.rect = ok ? 1 : 0;
  }
  P.dbg = igemm_dbg();
  if (P.rect) {
    const cuuint64_t es = (cuuint64_t)el.size;
    {  // dY viewed as [K/slab][pixels][slab]
      cuuint64_t dims[3] = {(cuuint64_t)slab, (cuuint64_t)m, (cuuint64_t)(d->k / slab)};
      cuuint64_t strides[2] = {(cuuint64_t)yk * es, (cuuint64_t)slab * es};
      cuuint32_t box[3] = {(cuuint32_t)slab, (cuuint32_t)KP, (cuuint32_t)(kTileM / slab)};
      cuuint32_t estr[3] = {1, 1, 1};
      if (make_tiled_nd(&P.tmDy, dy, el, 3, dims, strides, box, estr, swz)) return 1;
    }
    {  // x viewed as [C/slab][N][H][W][slab]; box = one tap's slabs for a (bnimg x bh x bw) block of output pixels
      const int per_box = (d->c < bn ? d->c : bn) / slab;
      cuuint64_t dims[5] = {(cuuint64_t)slab, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n, (cuuint64_t)(d->c / slab)};
      cuuint64_t strides[4] = {(cuuint64_t)xc * es, (cuuint64_t)d->w * xc * es, (cuuint64_t)d->h * d->w * xc * es,
                               (cuuint64_t)slab * es};
      cuuint32_t box[5] = {(cuuint32_t)slab, (cuuint32_t)(bw * d->stride_w), (cuuint32_t)(bh * d->stride_h), (cuuint32_t)bnimg,
                           (cuuint32_t)per_box};
      cuuint32_t estr[5] = {1, (cuuint32_t)d->stride_w, (cuuint32_t)d->stride_h, 1, 1};
      if (make_tiled_nd(&P.tmX, x, el, 5, dims, strides, box, estr, swz)) return 1;
    }
  } else {
    // dY as a [pixels][K] matrix; box = KP pixel rows x 128 bytes of channels
    if (make_tiled_2d(&P.tmDy, dy, el, (uint64_t)m, (uint64_t)d->k, (uint32_t)KP, swz, dy_ctot)) return 1;
    const int up_h = d->pad_h - (d->r - 1) * d->dil_h, up_w = d->pad_w - (d->s - 1) * d->dil_w;
    if (make_im2col_4d(&P.tmX, x, el, d->n, d->h, d->w, d->c, -d->pad_w, -d->pad_h, up_w, up_h, d->stride_w, d->stride_h, KP,
                       swz, x_ctot))
      return 1;
  }
  P.pad_w = d->pad_w;
  P.pad_h = d->pad_h;
  P.dil_w = d->dil_w;
  P.dil_h = d->dil_h;
  P.o.out = splits > 1 ? reinterpret_cast<float*>(ws) : dw;
  P.o.n_stride = ncols;
  P.o.h_stride = 0;
  P.o.w_stride = 0;
  P.o.base = 0;
  P.o.p_dim = 1;
  P.o.q_dim = 1;
  P.o.m_total = d->k;
  P.o.n_total = ncols;
  P.split_stride = wsize;
  P.c = d->c;
  P.pixel_steps_total = total;
  P.steps_per_split = sps;
  P.p_dim = d->p;
  P.q_dim = d->q;
  P.base_w = -d->pad_w;
  P.base_h = -d->pad_h;
  P.trav_w = d->stride_w;
  P.trav_h = d->stride_h;
  for (int r = 0; r < d->r; ++r)
    for (int s = 0; s < d->s; ++s) {
      int t = r * d->s + s;
      P.off_w[t] = (uint16_t)(s * d->dil_w);
      P.off_h[t] = (uint16_t)(r * d->dil_h);
    }
  const int ktiles = (int)ceil_div(d->k, kTileM), ntiles = (int)ceil_div(ncols, bn);
  int rc;
  if (el.bf16) {
    switch (bn) {
      case 256: rc = launch_wgrad<256, 64, 4, true>(P, ktiles, ntiles, splits, st); break;
      case 128: rc = launch_wgrad<128, 64, 6, true>(P, ktiles, ntiles, splits, st); break;
      default: rc = launch_wgrad<64, 64, 6, true>(P, ktiles, ntiles, splits, st); break;
    }
  } else if (KP == 32) {  // legacy small-box configuration, kept for A/B measurements (TTB_WGRAD_VARIANT=3)
    switch (bn) {
      case 256: rc = launch_wgrad<256, 32, 4, false>(P, ktiles, ntiles, splits, st); break;
      case 128: rc = launch_wgrad<128, 32, 4, false>(P, ktiles, ntiles, splits, st); break;
      case 64: rc = launch_wgrad<64, 32, 6, false>(P, ktiles, ntiles, splits, st); break;
      default: rc = launch_wgrad<32, 32, 6, false>(P, ktiles, ntiles, splits, st); break;
    }
  } else {
    switch (bn) {
      case 256: rc = launch_wgrad<256, 64, 2, false>(P, ktiles, ntiles, splits, st); break;
      case 128: rc = launch_wgrad<128, 64, 3, false>(P, ktiles, ntiles, splits, st); break;
      case 64: rc = launch_wgrad<64, 64, 4, false>(P, ktiles, ntiles, splits, st); break;
      default: rc = launch_wgrad<32, 64, 4, false>(P, ktiles, ntiles, splits, st); break;
    }
  }
  if (rc) return rc;
  if (splits_out) {
    *splits_out = splits;
    return 0;
  }
  if (splits > 1) {
    launch_k(sum_splits_kernel, elementwise_grid(wsize, 256), 256, 0, st, reinterpret_cast<const float*>(ws), splits, wsize, dw);
    return check_launch("wgrad sum_splits");
  }
  return 0;
}

int igemm_pack_dgrad_weights(int count, const ttb_conv_desc* const* descs, const float* const* w, float* const* wt,
                             cudaStream_t st) {
  for (int base = 0; base < count; base += kMaxBatch) {
    static thread_local RepackBatch B;
    B.count = count - base < kMaxBatch ? count - base : kMaxBatch;
    int chunks = 0;
    for (int i = 0; i < B.count; ++i) {
      const ttb_conv_desc* d = descs[base + i];
      B.w[i] = w[base + i];
      B.wt[i] = wt[base + i];
      B.K[i] = d->k;
      B.T[i] = d->r * d->s;
      B.C[i] = d->c;
      B.start[i] = chunks;
      chunks += d->r * d->s * (int)ceil_div(d->k, 32) * (int)ceil_div(d->c, 32);  // 32 x 32 tiles, see the kernel
    }
    B.start[B.count] = chunks;
    if (chunks == 0) continue;
    const int grid = chunks < sm_count() * 8 ? chunks : sm_count() * 8;
    launch_k(repack_multi_kernel, grid, 256, 0, st, B);
    if (check_launch("repack_multi")) return 1;
  }
  return 0;
}

int igemm_pack_weights_bf16(int count, const ttb_conv_desc* const* descs, const float* const* w, void* const* w_bf16,
                            void* const* wt_bf16, cudaStream_t st) {
  for (int base = 0; base < count; base += kMaxBatch) {
    static thread_local PackBf16Batch B;
    B.count = count - base < kMaxBatch ? count - base : kMaxBatch;
    int chunks = 0;
    for (int i = 0; i < B.count; ++i) {
      const ttb_conv_desc* d = descs[base + i];
      B.w[i] = w[base + i];
      B.wh[i] = reinterpret_cast<__nv_bfloat16*>(w_bf16[base + i]);
      B.wth[i] = reinterpret_cast<__nv_bfloat16*>(wt_bf16[base + i]);
      B.K[i] = d->k;
      B.T[i] = d->r * d->s;
      B.C[i] = d->c / d->groups;
      B.start[i] = chunks;
      chunks += d->r * d->s * (int)ceil_div(d->k, 32) * (int)ceil_div(B.C[i], 32);
    }
    B.start[B.count] = chunks;
    if (chunks == 0) continue;
    const int grid = chunks < sm_count() * 8 ? chunks : sm_count() * 8;
    launch_k(pack_bf16_multi_kernel, grid, 256, 0, st, B);
    if (check_launch("pack_bf16_multi")) return 1;
  }
  return 0;
}

int igemm_sum_splits_multi(int count, const float* const* partials, const int* splits, const int64_t* sizes,
                           float* const* outs, cudaStream_t st) {
  for (int base = 0; base < count; base += kMaxBatch) {
    static thread_local SumBatch B;
    B.count = count - base < kMaxBatch ? count - base : kMaxBatch;
    int chunks = 0;
    for (int i = 0; i < B.count; ++i) {
      B.partial[i] = partials[base + i];
      B.out[i] = outs[base + i];
      B.splits[i] = splits[base + i];
      B.n[i] = sizes[base + i];
      B.start[i] = chunks;
      chunks += (int)ceil_div(sizes[base + i], (int64_t)kBatchChunk);
    }
    B.start[B.count] = chunks;
    if (chunks == 0) continue;
    const int grid = chunks < sm_count() * 8 ? chunks : sm_count() * 8;
    launch_k(sum_splits_multi_kernel, grid, 256, 0, st, B);
    if (check_launch("sum_splits_multi")) return 1;
  }
  return 0;
}

}  // namespace ttb
