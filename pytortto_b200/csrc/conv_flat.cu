// "Flat-shift halo tile" fprop / dgrad: an EXPERIMENT kept in the tuning build only (-DTTB_TUNING; the release library
// does not contain it).  Validated on B200 (parity vs the exact fp32 kernels on every case of scripts/flat_check.py and
// tests/test_gpu_flat.py), but not faster than the production im2col kernel where it matters: on the 64 -> 64 channel
// 3x3 layer at batch 256 the resident variant measured 54.8 - 61.2 us against 56.5 - 56.8 us (graph replay, same box), and
// with epilogue statistics 65.5 against 58.2 us (profiles/r2_conv_layer_probe.txt).  The layer is bound by the N = 64
// tcgen05.mma itself (~62 tensor-pipe cycles per MMA, ncu: the 4 KB A operand is re-read from shared memory by every
// MMA) and by the strip loads, not by the A re-reads from L2 this design removes.  Two variants:
//   * RESIDENT (TTB_FLAT=-1: the 64 -> 64 channel 3x3 layers, TF32): the WHOLE weight matrix (K x R*S*C, 144 KB for
//     64 x 576 fp32) is loaded into shared memory once per CTA and stays there for all of the CTA's tiles; the main
//     loop then has no weight hand-shakes at all - per 32-channel slab one strip wait and 9 taps x 4 tcgen05.mma issued
//     back to back.  The production im2col kernel is issue-bound on these layers (one mbarrier hand-shake per 4 small
//     N = 64 MMAs: ~340 cycles per 128 tensor-core cycles, measured) and re-reads every input pixel 9 times from L2.
//   * STREAMED (weight tiles through a ring, any channel count): measured SLOWER than the production kernel (one CTA per
//     SM, same hand-shake count, 13-27 % dropped outputs) - kept for experiments in the tuning build (TTB_FLAT=1).
//
// "Flat-shift halo tile" fprop for stride-1 / dilation-1 convolutions (the 3x3 layers that dominate every ResNet): the
// production kernel (conv_igemm.cu) loads one im2col A tile per (filter tap, 32-channel slab) - 9 loads of the same
// pixels for a 3x3 filter - and is bound by operand feed / instruction issue on the 64- and 128-channel layers.  Here
// output pixels are addressed by a flat index over the ZERO-PADDED image,
//     f = (n*Hp + p)*Wp + q,   Hp = H + 2*pad_h, Wp = W + 2*pad_w,
// so that every filter tap is a pure shift of the input: input row = f + r*Wp + s.  A tile of 128 consecutive f needs the
// padded-flat input rows [f0, f0 + 128 + (R-1)*Wp + (S-1)); they are brought into shared memory ONCE per 32-channel slab,
// as one tiled-TMA box per padded image row (box = {32 channels, Wp pixels from w = -pad_w}: out-of-bounds zero fill IS
// the padding), and the R*S taps are R*S UMMA descriptors whose start address moves by whole 128-byte rows.
// Measured facts this relies on (profiles/r1_umma_row_shift_probe.txt):
//   * a 128B-swizzled K-major tcgen05.mma operand may start at ANY 128-byte row (base-offset field 0),
//   * a SWIZZLE_128B TMA box written to a 128-byte-aligned (not 1024-byte-aligned) address takes its XOR phase from the
//     absolute shared-memory address, so rows of Wp pixels (Wp not a multiple of 8) can be stacked back to back.
// Outputs at the padding positions (q >= Q or p >= P) are computed and dropped: (Hp*Wp)/(P*Q) - 1 = 13 % extra MMA work at
// 32x32, 27 % at 16x16 - the kernel is meant for the large-image layers.
//
// Structure (persistent, like igemm_fwd_persist_kernel): warp 0 = strip producer (A rows), warps 1-2 = weight-tile
// producers (alternating K-blocks), warp 3 = MMA issuer + TMEM owner, warps 4-7 = epilogue; two TMEM accumulators.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "conv_epilogue.cuh"
#include "sm100_ptx.cuh"

namespace ttb {

#ifdef TTB_TUNING
namespace {

constexpr int kFlatTileM = 128;
constexpr int kFlatMaxTaps = 64;
constexpr int kFlatStagePitch = kStagePitch;  // floats per staged epilogue row (conv_epilogue.cuh)
constexpr int kFlatSlabsPerStrip = 2;  // 32-channel slabs brought in per strip stage (64 channels)
constexpr int kFlatStripStages = 2;
constexpr int kFlatBProducers = 2;
constexpr int kFlatThreads = (1 + kFlatBProducers + 1 + 4) * 32;

struct FlatParams {
  CUtensorMap tmX;   // 4-D tiled over NHWC x: dims (C, W, H, N), box (32, Wp, 1, 1)
  CUtensorMap tmB;   // 2-D tiled over the weight matrix [rows = output channels][cols = (tap, c)], box (32, BN)
  float* out;        // dense NHWC output [N][P][Q][K]
  Epilogue ep;       // per-channel scale / bias, residual, ReLU, statistics (common.cuh)
  int n_img, hp, wp, pad_h, pad_w;   // padded grid
  int p_out, q_out, k_out;           // valid outputs per image, output channels
  int taps_r, taps_s;                // filter size
  int c_blocks;                      // input channels / 32
  int rows_max;                      // padded rows a strip stage holds per slab
  int64_t m_flat;                    // n_img * hp * wp
  int b_koff[kFlatMaxTaps];          // column of tap t's K-slice in the weight matrix (dgrad: flipped taps)
};

// accumulator row f (padded-flat index) -> element offset of its output row; rows at padding positions are dropped (-1)
__device__ __forceinline__ int64_t flat_row_offset(const FlatParams& P, int64_t f) {
  if (f >= P.m_flat) return -1;
  const int q = (int)(f % P.wp);
  const int64_t t = f / P.wp;
  const int p = (int)(t % P.hp);
  const int64_t n = t / P.hp;
  if (q >= P.q_out || p >= P.p_out) return -1;
  return ((n * P.p_out + p) * P.q_out + q) * (int64_t)P.k_out;
}

// RES: resident weights - NB is ignored, `P.c_blocks * taps` weight tiles live in shared memory for the whole kernel and a
// strip stage holds ONE 32-channel slab (kSlabsPerStrip = 1)
// STATS: per-channel sum / sum of squares of the stored outputs, one [2][K] double row per CTA (needs one N tile: nt == 1)
template <int BN, int NB, bool RES, bool STATS>
__global__ void __launch_bounds__(kFlatThreads, 1)
igemm_flat_kernel(const __grid_constant__ FlatParams P) {
  pdl_launch_dependents();  // (the matching pdl_wait() follows the prologue below)
  static_assert(NB % kFlatBProducers == 0, "a weight stage must always be filled by the same producer");
  constexpr uint32_t kBBytes = BN * 128;
  constexpr int kAccCols = BN < 32 ? 32 : BN;
  constexpr int kTmemCols = 2 * kAccCols;
  constexpr int kSlabsPerStrip = RES ? 1 : kFlatSlabsPerStrip;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t slab_bytes = (uint32_t)P.rows_max * (uint32_t)P.wp * 128u;       // one 32-channel slab of a strip
  const uint32_t strip_bytes = ((kSlabsPerStrip * slab_bytes) + 1023u) & ~1023u;
  const int n_btiles = RES ? P.c_blocks * P.taps_r * P.taps_s : NB;
  uint8_t* strip0 = smem;                                          // kFlatStripStages strips
  uint8_t* bring = smem + kFlatStripStages * strip_bytes;           // NB weight tiles (RES: every weight tile)
  float* staging = reinterpret_cast<float*>(bring + (size_t)n_btiles * kBBytes);  // epilogue staging
  __shared__ uint64_t strip_full[kFlatStripStages], strip_empty[kFlatStripStages], b_full[NB], b_empty[NB], acc_full[2],
      acc_empty[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int nt = (P.k_out + BN - 1) / BN;
  const int64_t mt = (P.m_flat + kFlatTileM - 1) / kFlatTileM;
  const int64_t tiles = mt * nt;
  const int taps = P.taps_r * P.taps_s;
  const int num_groups = (P.c_blocks + kSlabsPerStrip - 1) / kSlabsPerStrip;  // strip loads per tile
  constexpr int kMmaWarp = 1 + kFlatBProducers;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&P.tmX);
    ptx::prefetch_tmap(&P.tmB);
  }
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kFlatStripStages; ++s) {
        ptx::mbar_init(&strip_full[s], 1);
        ptx::mbar_init(&strip_empty[s], 1);
      }
      for (int s = 0; s < NB; ++s) {
        ptx::mbar_init(&b_full[s], 1);
        ptx::mbar_init(&b_empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(&acc_full[b], 1);
        ptx::mbar_init(&acc_empty[b], 4);
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
  pdl_wait();

  if (warp == 0) {
    // ===================== strip producer: padded input rows, one TMA box per (slab, padded row) =====================
    uint32_t g = 0;  // strip stages filled so far
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int64_t f0 = (t / nt) * kFlatTileM;
      const int64_t row0 = f0 / P.wp;  // first padded row (over n*Hp + hp) the tile touches
      int64_t last = f0 + kFlatTileM - 1 + (int64_t)(P.taps_r - 1) * P.wp + (P.taps_s - 1);
      int64_t row1 = last / P.wp;
      const int64_t rows_total = (int64_t)P.n_img * P.hp;
      if (row1 >= rows_total) row1 = rows_total - 1;  // rows past the last image only feed dropped outputs
      const int nrows = (int)(row1 - row0 + 1);
      for (int grp = 0; grp < num_groups; ++grp, ++g) {
        const uint32_t stage = g % kFlatStripStages, phase = (g / kFlatStripStages) & 1u;
        const int slabs = P.c_blocks - grp * kSlabsPerStrip < kSlabsPerStrip ? P.c_blocks - grp * kSlabsPerStrip
                                                                                     : kSlabsPerStrip;
        ptx::mbar_wait(&strip_empty[stage], phase ^ 1u);
        if (ptx::elect_one()) {
          uint8_t* base = strip0 + stage * strip_bytes;
          ptx::mbar_expect_tx(&strip_full[stage], (uint32_t)slabs * (uint32_t)nrows * (uint32_t)P.wp * 128u);
          for (int sl = 0; sl < slabs; ++sl)
            for (int j = 0; j < nrows; ++j) {
              const int64_t row = row0 + j;
              const int n = (int)(row / P.hp);
              const int hp = (int)(row - (int64_t)n * P.hp);
              ptx::tma_load_4d(base + sl * slab_bytes + (uint32_t)j * (uint32_t)P.wp * 128u, &P.tmX, &strip_full[stage],
                               (grp * kSlabsPerStrip + sl) * 32, -P.pad_w, hp - P.pad_h, n);
            }
        }
        __syncwarp();
      }
    }
  } else if (warp < kMmaWarp) {
    // ===================== weight-tile producers (K-block g belongs to producer g % kFlatBProducers) =====================
    const int me = warp - 1;
    uint32_t g = 0;
    if (RES) {
      // resident weights: every (tap, slab) tile once, all on one barrier; tile index = tap * c_blocks + slab
      if (me == 0 && blockIdx.x < tiles) {
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&b_full[0], (uint32_t)n_btiles * kBBytes);
          for (int tap = 0; tap < taps; ++tap)
            for (int sl = 0; sl < P.c_blocks; ++sl)
              ptx::tma_load_2d(bring + (size_t)(tap * P.c_blocks + sl) * kBBytes, &P.tmB, &b_full[0], P.b_koff[tap] + sl * 32, 0);
        }
        __syncwarp();
      }
    } else
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int n0 = (int)(t % nt) * BN;
      for (int grp = 0; grp < num_groups; ++grp) {
        const int slabs = P.c_blocks - grp * kSlabsPerStrip < kSlabsPerStrip ? P.c_blocks - grp * kSlabsPerStrip
                                                                                     : kSlabsPerStrip;
        for (int tap = 0; tap < taps; ++tap) {
          const int kb0 = P.b_koff[tap];
          for (int sl = 0; sl < slabs; ++sl, ++g) {
            if ((int)(g % kFlatBProducers) != me) continue;
            const uint32_t stage = g % NB, phase = (g / NB) & 1u;
            ptx::mbar_wait(&b_empty[stage], phase ^ 1u);
            if (ptx::elect_one()) {
              ptx::mbar_expect_tx(&b_full[stage], kBBytes);
              ptx::tma_load_2d(bring + stage * kBBytes, &P.tmB, &b_full[stage], kb0 + (grp * kSlabsPerStrip + sl) * 32, n0);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::umma_idesc(2 /*tf32*/, 0, 0, kFlatTileM, BN);
    uint32_t gs = 0, gb = 0;
    int it = 0;
    if (RES && blockIdx.x < tiles) {
      ptx::mbar_wait(&b_full[0], 0);  // the whole weight matrix is in shared memory from here on
      ptx::tc_fence_after();
    }
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int64_t f0 = (t / nt) * kFlatTileM;
      const int lead = (int)(f0 - (f0 / P.wp) * P.wp);  // tile start inside its first padded row
      const int buf = it & 1;
      ptx::mbar_wait(&acc_empty[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * kAccCols);
      bool first = true;
      for (int grp = 0; grp < num_groups; ++grp, ++gs) {
        const uint32_t sstage = gs % kFlatStripStages, sphase = (gs / kFlatStripStages) & 1u;
        const int slabs = P.c_blocks - grp * kSlabsPerStrip < kSlabsPerStrip ? P.c_blocks - grp * kSlabsPerStrip
                                                                                     : kSlabsPerStrip;
        ptx::mbar_wait(&strip_full[sstage], sphase);
        ptx::tc_fence_after();
        const uint32_t strip = ptx::smem_u32(strip0 + sstage * strip_bytes);
        if (RES) {
          // Resident weights, 3x3 taps: no hand-shake per K-block and NO per-tap scalar work - the 9 x 4 MMAs of a slab are
          // one straight-line block under a single elect (every descriptor is a warp-uniform base plus a compile-time
          // multiple of the row pitch).  With the tap loop rolled (index arithmetic, elect and reconvergence per tap) the
          // MMA warp needed ~410 cycles per tap against 128 tensor-core cycles (measured, profiles/r2_flat_resident.md).
          const uint32_t a0 = strip + (uint32_t)lead * 128u;
          const uint32_t wp128 = (uint32_t)P.wp * 128u;
          const uint32_t b0 = ptx::smem_u32(bring) + (uint32_t)grp * kBBytes;
          const uint32_t bstep = (uint32_t)P.c_blocks * kBBytes;  // weight tiles of consecutive taps
          const uint32_t acc0 = (uint32_t)grp;                     // 0 only for the first slab of the tile
          if (ptx::elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t sa = a0 + (uint32_t)(tap / 3) * wp128 + (uint32_t)(tap % 3) * 128u;
              const uint32_t sb = b0 + (uint32_t)tap * bstep;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = ptx::umma_desc_sw128(sa + k * 32, 16, 1024);
                const uint64_t db = ptx::umma_desc_sw128(sb + k * 32, 16, 1024);
                ptx::mma_tf32(d_tmem, da, db, idesc, acc0 | (uint32_t)(tap | k));
              }
            }
            ptx::mma_commit(&strip_empty[sstage]);
          }
          __syncwarp();
          continue;
        }
        for (int tap = 0; tap < taps; ++tap) {
          const int r = tap / P.taps_s, s = tap - r * P.taps_s;
          const uint32_t row_off = (uint32_t)(lead + r * P.wp + s) * 128u;  // the tap is a shift by whole rows
          for (int sl = 0; sl < slabs; ++sl, ++gb) {
            const uint32_t bstage = gb % NB, bphase = (gb / NB) & 1u;
            ptx::mbar_wait(&b_full[bstage], bphase);
            ptx::tc_fence_after();
            const uint32_t sa = strip + (uint32_t)sl * slab_bytes + row_off;
            const uint32_t sb = ptx::smem_u32(bring + bstage * kBBytes);
            if (ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = ptx::umma_desc_sw128(sa + k * 32, 16, 1024);
                const uint64_t db = ptx::umma_desc_sw128(sb + k * 32, 16, 1024);
                ptx::mma_tf32(d_tmem, da, db, idesc, !(first && k == 0));
              }
              ptx::mma_commit(&b_empty[bstage]);
            }
            __syncwarp();
            first = false;
          }
        }
        if (ptx::elect_one()) ptx::mma_commit(&strip_empty[sstage]);  // every MMA that read this strip has been issued
        __syncwarp();
      }
      if (ptx::elect_one()) ptx::mma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - (kMmaWarp + 1);
    float* const my_stage = staging + ew * (32 * kFlatStagePitch);
    EpiStats<BN> es;
    if (STATS) es.reset();
    int it = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int64_t f0 = (t / nt) * kFlatTileM;
      const int n0 = (int)(t % nt) * BN;
      const int buf = it & 1;
      ptx::mbar_wait(&acc_full[buf], ((uint32_t)it >> 1) & 1u);
      ptx::tc_fence_after();
      epilogue_tile<BN, STATS>(tmem_base + (uint32_t)(buf * kAccCols), my_stage, P.out,
                               flat_row_offset(P, f0 + (warp & 3) * 32 + lane), n0, P.k_out, P.ep, warp & 3, es);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[buf]);
    }
    if (STATS && it > 0)
      epilogue_stats_flush<BN>(es, staging, ew, 1, P.ep.stats + (int64_t)blockIdx.x * 2 * P.k_out, 0, P.k_out);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) ptx::tmem_dealloc<kTmemCols>(tmem_base);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled g_flat_encode = nullptr;

int flat_load_driver() {
  if (g_flat_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return 1;
  }
  g_flat_encode = (PFN_encodeTiled)fn;
  return 0;
}

constexpr size_t kFlatSmemLimit = 232448 - 512;  // 227 KB per CTA minus the kernel's static barriers

static unsigned flat_grid(int64_t m_flat, int k_out, int bn) {
  const int64_t tiles = ceil_div(m_flat, kFlatTileM) * ceil_div(k_out, bn);
  const int sms = sm_count();
  return (unsigned)(tiles < sms ? tiles : sms);
}

template <int BN, int NB, bool RES, bool STATS = false>
int flat_launch(const FlatParams& P, size_t strip_bytes_total, cudaStream_t st) {
  const size_t btiles = RES ? (size_t)P.c_blocks * P.taps_r * P.taps_s : (size_t)NB;
  const size_t smem = strip_bytes_total + btiles * BN * 128 + 4 * 32 * kFlatStagePitch * 4 + 1024;
  if (smem > kFlatSmemLimit) {
    set_error("conv flat path: %zu bytes of shared memory needed", smem);
    return 1;
  }
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(igemm_flat_kernel<BN, NB, RES, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("conv flat path: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return 1;
    }
    attr = smem;
  }
  const unsigned grid = flat_grid(P.m_flat, P.k_out, BN);
  launch_k(igemm_flat_kernel<BN, NB, RES, STATS>, grid, kFlatThreads, smem, st, P);
  return check_launch("igemm_flat_kernel");
}

}  // namespace

// TTB_FLAT: 0 = never (default), -1 = the resident-weight variant where it applies, 1 = every eligible problem (also the
// streamed variant)
static int flat_mode() {
  static const int mode = tuning_knob("TTB_FLAT", 0);
  return mode;
}

// geometry of one flat-shift problem: correlation of `in` (NHWC, c_in channels, zero-padded by pad) with r x s taps
struct FlatProblem {
  int n, c_in, h_in, w_in, k_out, r, s, pad_h, pad_w, p_out, q_out;
};

static int flat_rows_max_of(const FlatProblem& g) {
  const int wp = g.w_in + 2 * g.pad_w;
  // 128 + S - 1 consecutive padded-flat elements starting anywhere inside a row, plus the R - 1 rows below
  return (wp - 1 + kFlatTileM + g.s - 1 + wp - 1) / wp + (g.r - 1);
}

static size_t flat_strip_bytes(const FlatProblem& g, int slabs_per_strip) {
  const int wp = g.w_in + 2 * g.pad_w;
  return (size_t)kFlatStripStages * (((size_t)slabs_per_strip * flat_rows_max_of(g) * wp * 128 + 1023) & ~(size_t)1023);
}

static bool flat_geometry_ok(const FlatProblem& g) {
  if (g.c_in % 32 != 0 || g.k_out % 8 != 0 || g.r * g.s > kFlatMaxTaps || g.r * g.s < 2) return false;
  if (g.pad_h < 0 || g.pad_w < 0) return false;
  const int hp = g.h_in + 2 * g.pad_h, wp = g.w_in + 2 * g.pad_w;
  if (wp > 256 || hp - g.r + 1 != g.p_out || wp - g.s + 1 != g.q_out || g.p_out < 1 || g.q_out < 1) return false;
  return flat_strip_bytes(g, kFlatSlabsPerStrip) + 4 * 128 * 128 + 4 * 32 * kFlatStagePitch * 4 + 1024 <= kFlatSmemLimit;  // widest weight ring: 4 x (128 x 128 B)
}

// resident-weight variant: one N tile (K <= 64) whose whole weight matrix fits next to two one-slab strip stages, and
// enough tiles that loading the weights once per CTA (~150 KB) is amortised
static bool flat_resident_ok(const FlatProblem& g) {
  if (!flat_geometry_ok(g) || g.k_out > 64 || g.k_out < 33 || g.r != 3 || g.s != 3) return false;
  const size_t wbytes = (size_t)(g.c_in / 32) * g.r * g.s * 64 * 128;
  if (flat_strip_bytes(g, 1) + wbytes + 4 * 32 * kFlatStagePitch * 4 + 1024 > kFlatSmemLimit) return false;
  const int hp = g.h_in + 2 * g.pad_h, wp = g.w_in + 2 * g.pad_w;
  const int64_t tiles = ceil_div((int64_t)g.n * hp * wp, kFlatTileM);
  return tiles >= 4 * (int64_t)sm_count();
}

// in: NHWC fp32 [n][h_in][w_in][c_in]; wmat: [k_out rows][r*s*c_in cols] fp32 with tap t's slice at column koff[t];
// out: dense NHWC fp32 [n][p_out][q_out][k_out]
static int flat_run(const FlatProblem& g, const float* in, const float* wmat, const int* koff, const Epilogue& ep, float* out,
                    cudaStream_t st) {
  if (flat_load_driver()) return 1;
  static thread_local FlatParams P;
  memset(&P, 0, sizeof(P));
  const int hp = g.h_in + 2 * g.pad_h, wp = g.w_in + 2 * g.pad_w;
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.c_in, (cuuint64_t)g.w_in, (cuuint64_t)g.h_in, (cuuint64_t)g.n};
    cuuint64_t strides[3] = {(cuuint64_t)g.c_in * 4, (cuuint64_t)g.w_in * g.c_in * 4, (cuuint64_t)g.h_in * g.w_in * g.c_in * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)wp, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_flat_encode(&P.tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)in, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv flat path: cuTensorMapEncodeTiled(activation) failed (%d)", (int)r);
      return 1;
    }
  }
  // widest N tile that keeps most SMs busy; the weight ring shrinks as the tile grows
  const int64_t mt = ceil_div((int64_t)g.n * hp * wp, kFlatTileM);
  int bn = 64;
  if (g.k_out > 64 && mt * ceil_div(g.k_out, 128) * 5 >= (int64_t)sm_count() * 4) bn = 128;
  if (g.k_out <= 32) bn = 32;
  {
    const cuuint64_t cols = (cuuint64_t)g.r * g.s * g.c_in;
    cuuint64_t dims[2] = {cols, (cuuint64_t)g.k_out};
    cuuint64_t strides[1] = {cols * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_flat_encode(&P.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wmat, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv flat path: cuTensorMapEncodeTiled(weights) failed (%d)", (int)r);
      return 1;
    }
  }
  P.out = out;
  P.ep = ep;
  P.ep.relu = (ep.relu ? 1 : 0) | (tuning_knob("TTB_EPI_DBG", 0) << 8);  // (experiment bits: tuning build only)
  P.n_img = g.n;
  P.hp = hp;
  P.wp = wp;
  P.pad_h = g.pad_h;
  P.pad_w = g.pad_w;
  P.p_out = g.p_out;
  P.q_out = g.q_out;
  P.k_out = g.k_out;
  P.taps_r = g.r;
  P.taps_s = g.s;
  P.c_blocks = g.c_in / 32;
  P.rows_max = flat_rows_max_of(g);
  P.m_flat = (int64_t)g.n * hp * wp;
  for (int t = 0; t < g.r * g.s; ++t) P.b_koff[t] = koff[t];
  if (flat_resident_ok(g))
    return ep.stats ? flat_launch<64, 6, true, true>(P, flat_strip_bytes(g, 1), st)
                    : flat_launch<64, 6, true, false>(P, flat_strip_bytes(g, 1), st);
  if (ep.stats) {
    set_error("conv flat path: epilogue statistics need the resident-weight variant");
    return 1;
  }
  const size_t strips = flat_strip_bytes(g, kFlatSlabsPerStrip);
  switch (bn) {
    case 128: return flat_launch<128, 4, false>(P, strips, st);
    case 64: return flat_launch<64, 6, false>(P, strips, st);
    default: return flat_launch<32, 6, false>(P, strips, st);
  }
}

static bool flat_common_ok(const ttb_conv_desc* d) {
  return flat_mode() != 0 && d->math_mode == TTB_MATH_TF32 && d->groups == 1 && d->stride_h == 1 && d->stride_w == 1 &&
         d->dil_h == 1 && d->dil_w == 1;
}

static bool flat_take(const FlatProblem& g) { return flat_mode() == 1 ? flat_geometry_ok(g) : flat_resident_ok(g); }

static FlatProblem flat_fprop_problem(const ttb_conv_desc* d) {
  return FlatProblem{d->n, d->c, d->h, d->w, d->k, d->r, d->s, d->pad_h, d->pad_w, d->p, d->q};
}

// dgrad of a stride-1 convolution = the same correlation over dY with the taps flipped and padding R-1-pad
static FlatProblem flat_dgrad_problem(const ttb_conv_desc* d) {
  return FlatProblem{d->n, d->k, d->p, d->q, d->c, d->r, d->s, d->r - 1 - d->pad_h, d->s - 1 - d->pad_w, d->h, d->w};
}

// 1 when the flat-shift kernel takes the problem: TF32, stride 1, dilation 1, groups 1, C % 32 == 0 and the
// resident-weight variant applies (33..64 output channels, weights + strips fit in shared memory, >= 4 tiles per SM)
bool flat_fprop_supported(const ttb_conv_desc* d) { return flat_common_ok(d) && flat_take(flat_fprop_problem(d)); }
bool flat_dgrad_supported(const ttb_conv_desc* d) { return flat_common_ok(d) && flat_take(flat_dgrad_problem(d)); }

// x: NHWC fp32, w: [K][R][S][C] fp32, y: dense NHWC fp32
int flat_fprop(const ttb_conv_desc* d, const float* x, const float* w, const Epilogue& ep, float* y, cudaStream_t st) {
  int koff[kFlatMaxTaps];
  for (int t = 0; t < d->r * d->s; ++t) koff[t] = t * d->c;
  return flat_run(flat_fprop_problem(d), x, w, koff, ep, y, st);
}

// rows of the [chunks][2][K] statistics partial buffer flat_fprop writes for this problem (Epilogue::stats); 0: the
// resident-weight variant (the only one that emits statistics) does not take the problem
int flat_fprop_stats_chunks(const ttb_conv_desc* d) {
  const FlatProblem g = flat_fprop_problem(d);
  if (!flat_common_ok(d) || !flat_resident_ok(g)) return 0;
  return (int)flat_grid((int64_t)g.n * (g.h_in + 2 * g.pad_h) * (g.w_in + 2 * g.pad_w), g.k_out, 64);
}

// dy: NHWC fp32 [N][P][Q][K], w_packed: [C][R][S][K] fp32 (the dgrad re-ordering), dx: dense NHWC fp32 [N][H][W][C]
int flat_dgrad(const ttb_conv_desc* d, const float* dy, const float* w_packed, float* dx, cudaStream_t st, const float* accum) {
  int koff[kFlatMaxTaps];
  for (int r = 0; r < d->r; ++r)
    for (int s = 0; s < d->s; ++s) koff[r * d->s + s] = ((d->r - 1 - r) * d->s + (d->s - 1 - s)) * d->k;  // flipped taps
  return flat_run(flat_dgrad_problem(d), dy, w_packed, koff, Epilogue{nullptr, nullptr, accum, 0, nullptr}, dx, st);
}

#else  // release build: the production im2col kernels take every problem
bool flat_fprop_supported(const ttb_conv_desc*) { return false; }
bool flat_dgrad_supported(const ttb_conv_desc*) { return false; }
int flat_fprop(const ttb_conv_desc*, const float*, const float*, const Epilogue&, float*, cudaStream_t) { return 1; }
int flat_fprop_stats_chunks(const ttb_conv_desc*) { return 0; }
int flat_dgrad(const ttb_conv_desc*, const float*, const float*, float*, cudaStream_t, const float*) { return 1; }
#endif

}  // namespace ttb
