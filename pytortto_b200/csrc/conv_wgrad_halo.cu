// Haloed-tile wgrad for stride-1 filters (tcgen05, sm_100a): ONE x tile serves every filter tap.
//
//   dW[k][r][s][c] = sum_{n,p,q} dY[n][p][q][k] * x[n][p + r - ph][q + s - pw][c]          (reference: _conv2d_backward_w,
//                                                                                            autograd/grad_nn.py:646-656)
// The im2col wgrad kernel (conv_igemm.cu) loads, for every step of 64 output pixels, one x box PER FILTER TAP - nine
// re-reads of the same pixels for a 3x3 filter - and puts the output channels on the 128 accumulator rows, so a 64-filter
// layer wastes half of every MMA.  For the narrow layers (<= 128 filters: the 64-channel stage of the ResNets, every UNet
// level) both costs dominate: the kernel is bound by TMA rows, not by the tensor pipe (84 us for 19.3 GFLOP).
//
// Here a step is a BH x BW box of output pixels of one image.  Shared memory receives
//   * the dY box  [K/slab][BH*BW pixel rows][128 B]                       (one tiled 5-D TMA load), and
//   * the x HALO  [c-slabs][(BH + R - 1) x (BW + S - 1) pixel rows][128 B] (one tiled 5-D TMA load; out-of-image rows /
//     columns are zero-filled by the TMA unit = the convolution's padding),
// both MN-major (the reduction index - the pixel - is the row of the tile).  A filter tap (r, s) is then nothing but the
// halo tile read (r * (BW+S-1) + s) pixel rows later, and because the 128-byte-wide slabs that make up the M = 128 rows of
// one tcgen05.mma may OVERLAP in shared memory (leading-dimension byte offset = ONE pixel row; measured exact for tf32 and
// bf16: profiles/r2_umma_lbo_overlap_probe.txt) a single MMA covers the taps s, s+1, .. of a filter row at once:
//   accumulator (c-slab, r, tap group)[lane = (tap in group, channel in slab)][column = k]
//       += halo[pixel rows of K-group g, shifted by tap (r, s)]^T  *  dY[pixel rows of K-group g]
// The accumulator rows are (tap, input channel) and its COLUMNS the output channels, so a 64-filter layer issues N = 64 MMAs
// over full 128-row tiles (tf32: 4 taps per MMA, 3 of them real for a 3-wide filter), x is read from L2 about (1 + 2/BH) x
// instead of 9 x, and every accumulator of the CTA stays in TMEM (<= 512 columns) for its whole pixel range.
// The pixel range is split over the grid; partial dW buffers are summed in fixed order by the caller's split reduction.
#include <cuda.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace ttb {

// conv_igemm.cu
int tma_make_tiled(CUtensorMap* tm, const void* base, bool bf16, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides, bool mn_major);
void launch_sum_splits(const float* partial, int splits, int64_t n, float* out, cudaStream_t st);

constexpr int kHaloThreads = 224;
constexpr int kHaloMaxStages = 8;
constexpr uint32_t kHaloSmemMax = 226u * 1024u;  // dynamic shared memory the kernel may be given (232448 - static)

struct WgradHaloParams {
  CUtensorMap tmX, tmDy;
  float* out;              // dW, or the first of `splits` partial buffers
  int64_t split_stride;    // elements between partial buffers
  int C, K, R, S;
  int BH, BW, HW;          // box of output pixels (rows, columns); halo tile width BW + S - 1
  int boxes_w, boxes_h, boxes_total, boxes_per_split;
  int pad_h, pad_w;
  int ncs, nr, rgroups;    // input-channel slabs / filter rows of one CTA; R / nr
  uint32_t x_slab_bytes;   // (BH + nr - 1) * HW * 128
  uint32_t x_bytes;        // ncs * x_slab_bytes (what the TMA writes)
  uint32_t dy_off;         // x_bytes rounded up to 1024
  uint32_t dy_bytes;
  uint32_t stage_bytes;
  int nstages;
  uint32_t tmem_cols;
  int dbg;                 // timing experiments (tuning build, WRONG results): 1 no MMAs, 2 no x loads, 4 no dY loads
};

__device__ __forceinline__ void tmem_alloc_dyn(uint32_t* dst_smem, uint32_t cols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(dst_smem)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc_dyn(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}

// Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..5 = epilogue (TMEM lane quarter = warp % 4), 6 = second MMA
// issuer: an N = 64 MMA is 32 tensor-pipe cycles but ~50 cycles of issue for one thread (measured), so the accumulators are
// split between two issuing warps (each commits its own MMAs; a stage is free / the result complete after both commits).
// NCS / NR / SG (input-channel slabs, filter rows, tap groups per filter row of one CTA) are compile-time: the MMA-issuing
// thread runs alone, so everything it executes between two MMAs is issue latency - with run-time loops, a division and a
// 64-bit descriptor build per MMA the first version of this kernel spent ~240 cycles per 32-cycle MMA (measured).  Here the
// accumulator loops are unrolled over descriptor offsets held in registers and one MMA costs two 32-bit adds.
template <bool BF16, int NCS, int NR, int SG>
__global__ void __launch_bounds__(kHaloThreads, 1) wgrad_halo_kernel(const __grid_constant__ WgradHaloParams P) {
  pdl_launch_dependents();  // (the matching pdl_wait() follows the prologue)
  constexpr int kSlabCh = BF16 ? 64 : 32;           // channels per 128-byte row
  constexpr int TW = 128 / kSlabCh;                 // overlapping slabs = filter taps along W covered by one MMA
  constexpr int kMmaRows = BF16 ? 16 : 8;           // pixels (K) per MMA
  constexpr uint32_t kSbo = BF16 ? 1024 : 512, kLayout = BF16 ? 2 : 1;
  constexpr int NACC = NCS * NR * SG;               // accumulators of this CTA, K columns each
  constexpr int kIssuers = NACC > 1 ? 2 : 1;        // MMA-issuing warps
  constexpr int kAccSplit = (NACC + kIssuers - 1) / kIssuers;  // accumulators [0, kAccSplit) belong to warp 1, the rest to warp 6

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[kHaloMaxStages], empty_bar[kHaloMaxStages], accum_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform (see conv_igemm.cu)
  const int lane = threadIdx.x & 31;
  const int cs0 = (blockIdx.y / P.rgroups) * NCS; // first input-channel slab of this CTA
  const int r0 = (blockIdx.y % P.rgroups) * NR;     // first filter row
  const int split = blockIdx.x;
  const int box0 = split * P.boxes_per_split;
  int nbox = P.boxes_total - box0;
  if (nbox > P.boxes_per_split) nbox = P.boxes_per_split;
  const int KP = P.BH * P.BW;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&P.tmX);
    ptx::prefetch_tmap(&P.tmDy);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < P.nstages; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], kIssuers);
      }
      ptx::mbar_init(&accum_bar, kIssuers);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_dyn(&tmem_base_smem, P.tmem_cols);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (whole warp, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int b = 0; b < nbox; ++b) {
      const int box = box0 + b;
      const int bj = box % P.boxes_w;
      const int t = box / P.boxes_w;
      const int bi = t % P.boxes_h;
      const int n = t / P.boxes_h;
      ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
      uint8_t* sx = smem + (size_t)stage * P.stage_bytes;
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&full_bar[stage], ((P.dbg & 2) ? 0u : P.x_bytes) + ((P.dbg & 4) ? 0u : P.dy_bytes));
        // x halo: [slab channels][W][H][N][C/slab] box (slab, HW, BH + nr - 1, 1, ncs); rows / columns outside the image
        // (the padding) are zero-filled
        if (!(P.dbg & 2)) ptx::tma_load_5d(sx, &P.tmX, &full_bar[stage], 0, bj * P.BW - P.pad_w, bi * P.BH - P.pad_h + r0, n, cs0);
        // dY: [slab channels][Q][P][N][K/slab] box (slab, BW, BH, 1, K/slab); pixels past the grid are zero-filled
        if (!(P.dbg & 4)) ptx::tma_load_5d(sx + P.dy_off, &P.tmDy, &full_bar[stage], 0, bj * P.BW, bi * P.BH, n, 0);
      }
      __syncwarp();
      if (++stage == P.nstages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1 || warp == 6) {
    // ===================== MMA issuers (whole warp, one elected lane issues) =====================
    const bool second = warp == 6;
    const uint32_t idesc = ptx::umma_idesc(BF16 ? 1 /*bf16*/ : 2 /*tf32*/, 1, 1, 128, (uint32_t)P.K);
    const int ngroups = (P.dbg & 1) ? 0 : KP / kMmaRows;
    // descriptor = {high word: LBO-independent fields, low word: start address >> 4 | LBO >> 4 << 16}; every operand of the
    // kernel differs from its stage's first one only in the start address, i.e. by an addend to the low word
    const uint32_t a_hi = (uint32_t)(ptx::umma_desc(0, 128u, kSbo, kLayout) >> 32);
    const uint32_t b_hi = (uint32_t)(ptx::umma_desc(0, (uint32_t)KP * 128u, kSbo, kLayout) >> 32);
    const uint32_t a_lo0 = (uint32_t)ptx::umma_desc(0, 128u, kSbo, kLayout);
    const uint32_t b_lo0 = (uint32_t)ptx::umma_desc(0, (uint32_t)KP * 128u, kSbo, kLayout);
    uint32_t acc_off[NACC];  // (c-slab a, filter row r, tap group sg) -> start-address addend (16-byte units)
#pragma unroll
    for (int a = 0; a < NCS; ++a)
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int sg = 0; sg < SG; ++sg)
          acc_off[(a * NR + r) * SG + sg] = ((uint32_t)a * P.x_slab_bytes + (uint32_t)(r * P.HW + sg * TW) * 128u) >> 4;
    const uint32_t row_skip = (uint32_t)(P.HW - P.BW) * 8u;  // halo columns between two box rows (16-byte units)
    int stage = 0;
    uint32_t phase = 0;
    for (int b = 0; b < ((second && kIssuers == 1) ? 0 : nbox); ++b) {
      ptx::mbar_wait(&full_bar[stage], phase);
      ptx::tc_fence_after();
      const uint32_t sx = ptx::smem_u32(smem + (size_t)stage * P.stage_bytes);
      if (ptx::elect_one()) {
        uint32_t a_lo = a_lo0 + (sx >> 4), b_lo = b_lo0 + ((sx + P.dy_off) >> 4);
        int gj = 0;
        for (int g = 0; g < ngroups; ++g) {  // kMmaRows consecutive pixels of one box row
          const uint64_t db = ((uint64_t)b_hi << 32) | b_lo;
          const uint32_t accumulate = (uint32_t)(b | g);
          // slabs 128 bytes = ONE pixel row apart: slab j of the A operand is filter tap s = sg * TW + j
          auto issue = [&](auto lo, auto hi) {
#pragma unroll
            for (int i = decltype(lo)::value; i < decltype(hi)::value; ++i) {
              const uint64_t da = ((uint64_t)a_hi << 32) | (a_lo + acc_off[i]);
              if (BF16) ptx::mma_bf16(tmem_base + (uint32_t)i * (uint32_t)P.K, da, db, idesc, accumulate);
              else ptx::mma_tf32(tmem_base + (uint32_t)i * (uint32_t)P.K, da, db, idesc, accumulate);
            }
          };
          if (second) issue(std::integral_constant<int, kAccSplit>{}, std::integral_constant<int, NACC>{});
          else issue(std::integral_constant<int, 0>{}, std::integral_constant<int, kAccSplit>{});
          a_lo += kMmaRows * 8u;
          b_lo += kMmaRows * 8u;
          gj += kMmaRows;
          if (gj == P.BW) { gj = 0; a_lo += row_skip; }
        }
        ptx::mma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == P.nstages) { stage = 0; phase ^= 1u; }
    }
    if (!(second && kIssuers == 1) && ptx::elect_one()) ptx::mma_commit(&accum_bar);
    __syncwarp();
  } else {
    // ===================== epilogue: TMEM -> dW partial =====================
    // lane (tap in group, channel in slab) holds column k of its accumulator row: for a fixed k the 32 lanes of a warp are
    // 32 consecutive input channels of one tap = one 128-byte segment of dW[k][r][s][:]
    if (lane == 0) ptx::mbar_wait(&accum_bar, 0);
    __syncwarp();
    ptx::tc_fence_after();
    const int lane_block = warp & 3;
    const int row = lane_block * 32 + lane;
    const int tap_in_group = row / kSlabCh, c_in = row % kSlabCh;
    const int64_t rsc = (int64_t)P.R * P.S * P.C;
    float* const out = P.out + (int64_t)split * P.split_stride;
    uint32_t acc = 0;
    for (int a = 0; a < NCS; ++a)
      for (int r = 0; r < NR; ++r)
        for (int sg = 0; sg < SG; ++sg, ++acc) {
          const int s = sg * TW + tap_in_group;  // warp-uniform (32 divides the slab width)
          float* const dst = out + ((int64_t)(r0 + r) * P.S + s) * P.C + (cs0 + a) * kSlabCh + c_in;
          for (int kc = 0; kc < P.K; kc += 32) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(lane_block * 32) << 16) + acc * (uint32_t)P.K + (uint32_t)kc, v);
            ptx::tmem_ld_wait();
            if (s < P.S && nbox > 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j) dst[(int64_t)(kc + j) * rsc] = __uint_as_float(v[j]);
            } else if (s < P.S) {
#pragma unroll
              for (int j = 0; j < 32; ++j) dst[(int64_t)(kc + j) * rsc] = 0.f;
            }
          }
        }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc_dyn(tmem_base, P.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------------
struct HaloPlan {
  int BH, BW, HW, ncs, nr, rgroups, cgroups, nstages, splits, boxes_w, boxes_h, boxes_total, boxes_per_split;
  uint32_t x_slab_bytes, x_bytes, dy_off, dy_bytes, stage_bytes, tmem_cols;
  size_t smem;
};

static bool halo_plan(const ttb_conv_desc* d, HaloPlan* h) {
  const bool bf16 = d->math_mode == TTB_MATH_BF16;
  if (d->math_mode != TTB_MATH_TF32 && !bf16) return false;
  const int slab = bf16 ? 64 : 32, tw = 128 / slab, mma_rows = bf16 ? 16 : 8;
  if (d->groups != 1 || d->stride_h != 1 || d->stride_w != 1 || d->dil_h != 1 || d->dil_w != 1) return false;
  if (d->s < 2 || d->s > 4 || d->r < 1 || d->r > 8) return false;  // (1-wide filters would waste the overlapped taps)
  if (d->c % slab != 0 || d->k % 32 != 0 || d->k > 256 || d->k % slab != 0) return false;
  if (d->n <= 0 || d->p <= 0 || d->q < mma_rows) return false;
  if (d->pad_h > 64 || d->pad_w > 64) return false;
  if ((int64_t)d->n * d->h * d->w * d->c >= (1ll << 40)) return false;
  const int sg = (d->s + tw - 1) / tw;
  const int cap = 512 / (sg * d->k);  // accumulators (c-slab, filter row) a CTA can keep in TMEM
  if (cap < 1) return false;
  const int cslabs = d->c / slab;
  // (the kernel is instantiated for 1 or 3 filter rows and 1 or 2 channel slabs per CTA; other filter heights run one row
  // per CTA)
  const int nr = (d->r == 3 && cap >= 3) ? 3 : 1;
  int ncs = (cap / nr >= 2 && cslabs % 2 == 0) ? 2 : 1;
  const int kp_target = tuning_knob("TTB_HALO_KP", 64);
  int bw = (d->q + mma_rows - 1) / mma_rows * mma_rows;
  if (bw > 64) bw = 64;
  int bh = kp_target / bw;
  if (bh < 1) bh = 1;
  if (bh > d->p) bh = d->p;
  const int hw = bw + d->s - 1;
  for (;;) {
    const uint32_t x_slab = (uint32_t)(bh + nr - 1) * hw * 128u;
    h->x_slab_bytes = x_slab;
    h->x_bytes = x_slab * ncs;
    h->dy_off = (h->x_bytes + 1023u) & ~1023u;
    h->dy_bytes = (uint32_t)(d->k / slab) * bh * bw * 128u;
    h->stage_bytes = (h->dy_off + h->dy_bytes + 1023u) & ~1023u;
    int st = (int)((kHaloSmemMax - 2048u) / h->stage_bytes);
    if (st >= 2 || ncs == 1) {
      if (st < 2) return false;
      h->nstages = st > kHaloMaxStages ? kHaloMaxStages : st;
      break;
    }
    ncs = 1;  // (stage too large: one channel slab per CTA)
  }
  if (bh + nr - 1 > 256 || hw > 256) return false;
  h->BH = bh; h->BW = bw; h->HW = hw; h->ncs = ncs; h->nr = nr;
  h->rgroups = d->r / nr;
  h->cgroups = cslabs / ncs;
  h->boxes_w = (d->q + bw - 1) / bw;
  h->boxes_h = (d->p + bh - 1) / bh;
  h->boxes_total = d->n * h->boxes_w * h->boxes_h;
  const int gy = h->rgroups * h->cgroups;
  int sp = sm_count() / gy;
  if (sp < 1) sp = 1;
  if (sp > h->boxes_total) sp = h->boxes_total;
  h->boxes_per_split = (h->boxes_total + sp - 1) / sp;
  h->splits = (h->boxes_total + h->boxes_per_split - 1) / h->boxes_per_split;
  uint32_t cols = (uint32_t)(ncs * nr * sg * d->k), pow2 = 32;
  while (pow2 < cols) pow2 <<= 1;
  h->tmem_cols = pow2;
  h->smem = (size_t)h->nstages * h->stage_bytes + 1024;
  return true;
}

// Which wgrads take the haloed kernel: the narrow layers, where the im2col kernel wastes accumulator rows (K <= 64) or
// is bound by its per-tap reloads.  TTB_WGRAD_HALO (tuning build): 0 never, 2 whenever the geometry allows.
bool halo_wgrad_selected(const ttb_conv_desc* d) {
  const int mode = tuning_knob("TTB_WGRAD_HALO", 1);
  if (mode == 0) return false;
  HaloPlan h;
  if (!halo_plan(d, &h)) return false;
  if (mode == 2) return true;
  return d->k <= tuning_knob("TTB_HALO_MAXK", 128);
}

size_t halo_wgrad_workspace(const ttb_conv_desc* d) {
  HaloPlan h;
  if (!halo_plan(d, &h)) return 0;
  return h.splits > 1 ? (size_t)h.splits * d->k * d->r * d->s * d->c * sizeof(float) : 0;
}

int halo_wgrad(const ttb_conv_desc* d, const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes, cudaStream_t st,
               int* splits_out) {
  HaloPlan h;
  TTB_REQUIRE(halo_plan(d, &h), "conv2d_wgrad (halo): unsupported problem");
  const bool bf16 = d->math_mode == TTB_MATH_BF16;
  const uint64_t es = bf16 ? 2 : 4, slab = bf16 ? 64 : 32;
  const int64_t wsize = (int64_t)d->k * d->r * d->s * d->c;
  if (h.splits > 1)
    TTB_REQUIRE(ws != nullptr && ws_bytes >= (size_t)h.splits * wsize * sizeof(float),
                "conv2d_wgrad: workspace of %zu bytes needed, %zu given", (size_t)h.splits * wsize * sizeof(float), ws_bytes);
  static thread_local WgradHaloParams P;
  memset(&P, 0, sizeof(P));
  {  // x viewed as [C/slab][N][H][W][slab]
    const uint64_t dims[5] = {slab, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->n, (uint64_t)d->c / slab};
    const uint64_t strides[4] = {(uint64_t)d->c * es, (uint64_t)d->w * d->c * es, (uint64_t)d->h * d->w * d->c * es, slab * es};
    const uint32_t box[5] = {(uint32_t)slab, (uint32_t)h.HW, (uint32_t)(h.BH + h.nr - 1), 1u, (uint32_t)h.ncs};
    const uint32_t estr[5] = {1, 1, 1, 1, 1};
    if (tma_make_tiled(&P.tmX, x, bf16, 5, dims, strides, box, estr, true)) return 1;
  }
  {  // dY viewed as [K/slab][N][P][Q][slab]
    const uint64_t dims[5] = {slab, (uint64_t)d->q, (uint64_t)d->p, (uint64_t)d->n, (uint64_t)d->k / slab};
    const uint64_t strides[4] = {(uint64_t)d->k * es, (uint64_t)d->q * d->k * es, (uint64_t)d->p * d->q * d->k * es, slab * es};
    const uint32_t box[5] = {(uint32_t)slab, (uint32_t)h.BW, (uint32_t)h.BH, 1u, (uint32_t)(d->k / slab)};
    const uint32_t estr[5] = {1, 1, 1, 1, 1};
    if (tma_make_tiled(&P.tmDy, dy, bf16, 5, dims, strides, box, estr, true)) return 1;
  }
  P.out = h.splits > 1 ? reinterpret_cast<float*>(ws) : dw;
  P.split_stride = wsize;
  P.C = d->c; P.K = d->k; P.R = d->r; P.S = d->s;
  P.BH = h.BH; P.BW = h.BW; P.HW = h.HW;
  P.boxes_w = h.boxes_w; P.boxes_h = h.boxes_h; P.boxes_total = h.boxes_total; P.boxes_per_split = h.boxes_per_split;
  P.pad_h = d->pad_h; P.pad_w = d->pad_w;
  P.ncs = h.ncs; P.nr = h.nr; P.rgroups = h.rgroups;
  P.x_slab_bytes = h.x_slab_bytes; P.x_bytes = h.x_bytes; P.dy_off = h.dy_off; P.dy_bytes = h.dy_bytes;
  P.stage_bytes = h.stage_bytes; P.nstages = h.nstages; P.tmem_cols = h.tmem_cols;
  P.dbg = tuning_knob("TTB_HALO_DBG", 0);
  const int sg = (d->s + (bf16 ? 2 : 4) - 1) / (bf16 ? 2 : 4);
  const dim3 grid((unsigned)h.splits, (unsigned)(h.rgroups * h.cgroups), 1);
  int rc = -1;
#define TTB_HALO_CASE(BF, NCS_, NR_, SG_)                                                                                 \
  if (bf16 == BF && h.ncs == NCS_ && h.nr == NR_ && sg == SG_) {                                                          \
    static bool attr_set = false;                                                                                         \
    if (!attr_set) {                                                                                                      \
      cudaError_t e = cudaFuncSetAttribute(wgrad_halo_kernel<BF, NCS_, NR_, SG_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)kHaloSmemMax);                                                            \
      if (e != cudaSuccess) {                                                                                             \
        set_error("wgrad (halo): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));                                \
        return 1;                                                                                                         \
      }                                                                                                                   \
      attr_set = true;                                                                                                    \
    }                                                                                                                     \
    launch_k(wgrad_halo_kernel<BF, NCS_, NR_, SG_>, grid, kHaloThreads, h.smem, st, P);                                   \
    rc = 0;                                                                                                               \
  }
  TTB_HALO_CASE(false, 1, 1, 1) TTB_HALO_CASE(false, 1, 3, 1) TTB_HALO_CASE(false, 2, 1, 1) TTB_HALO_CASE(false, 2, 3, 1)
  TTB_HALO_CASE(true, 1, 1, 1) TTB_HALO_CASE(true, 1, 3, 1) TTB_HALO_CASE(true, 2, 1, 1) TTB_HALO_CASE(true, 2, 3, 1)
  TTB_HALO_CASE(true, 1, 1, 2) TTB_HALO_CASE(true, 1, 3, 2) TTB_HALO_CASE(true, 2, 1, 2) TTB_HALO_CASE(true, 2, 3, 2)
#undef TTB_HALO_CASE
  TTB_REQUIRE(rc == 0, "wgrad (halo): no kernel instance for ncs=%d nr=%d sg=%d", h.ncs, h.nr, sg);
  if (check_launch("wgrad_halo_kernel")) return 1;
  if (splits_out) {
    *splits_out = h.splits;
    return 0;
  }
  if (h.splits > 1) launch_sum_splits(reinterpret_cast<const float*>(ws), h.splits, wsize, dw, st);
  return check_launch("wgrad sum_splits");
}

}  // namespace ttb
