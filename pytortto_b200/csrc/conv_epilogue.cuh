// Shared epilogue of the tcgen05 convolution kernels (conv_igemm.cu, conv_flat.cu): accumulator tile (128 TMEM lanes x BN
// fp32 columns) -> per-channel scale / bias -> residual add -> ReLU -> global rows, and - optionally - the per-channel
// sum / sum of squares of what was stored, so that the BatchNorm that follows the convolution does not have to read the
// tensor again for its statistics (reference: BatchNorm.forward's xp.mean / xp.var passes, autograd/grad_nn.py:923-924).
#pragma once
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace ttb {

constexpr int kStagePitch = 36;      // floats per staged epilogue row (32 + 4 pad: conflict-free float4 access)
constexpr int kEpilogueStagingBytes = 4 * 32 * kStagePitch * 4;  // four epilogue warps x 32 rows

// Running per-lane column statistics of one epilogue warp, kept in registers across all tiles of a persistent CTA.
// Lane (sub = lane / 8, c4 = lane % 8) stores - and therefore accumulates - columns 4*c4 .. 4*c4+3 of every 32-column block
// for rows sub, sub + 4, ... of the warp's 32 rows.  Sums are SHIFTED by a lane-private K (the first value the lane saw
// in that column): sum(v - K), sum((v - K)^2) in fp32 do not cancel when |mean| >> sd, and each lane converts its own
// sums to plain double sums at the end (epilogue_stats_flush), so K never has to agree between lanes.
template <int BN>
struct EpiStats {
  float k[BN / 32][4], s0[BN / 32][4], s1[BN / 32][4];
  int cnt;  // rows accumulated so far (identical for every column block)
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int b = 0; b < BN / 32; ++b)
#pragma unroll
      for (int j = 0; j < 4; ++j) k[b][j] = s0[b][j] = s1[b][j] = 0.f;
    cnt = 0;
  }
};

// One tile.  `my_off`: element offset of the output row this thread's accumulator row maps to (< 0: the row does not
// exist / is dropped); `st`: this warp's 32 x kStagePitch staging floats; `lane_block` = TMEM lane block (warp_id % 4).
// STATS: 0 none, 1 sum / sum of squares of the stored values (forward statistics), 2 the BatchNorm-backward sums (Epilogue::bn_x)
template <int BN, int STATS>
__device__ __forceinline__ void epilogue_tile(uint32_t tmem_acc, float* st, float* __restrict__ out, int64_t my_off, int n0,
                                              int n_total, const Epilogue& ep, int lane_block, EpiStats<BN>& es) {
  const int lane = threadIdx.x & 31;
  const int sub = lane >> 3, c4 = lane & 7;
  const bool fresh = STATS == 1 && es.cnt == 0;
  int tile_rows = 0;
  // output offsets (in float4 units; every row offset is a multiple of 4 elements) of the 8 rows this lane stores, fetched
  // from their owner lanes once per tile instead of once per column block
  const int my_off4 = my_off < 0 ? -1 : (int)(my_off >> 2);
  int off4[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    off4[it] = __shfl_sync(0xffffffffu, my_off4, it * 4 + sub);
    if (STATS && off4[it] >= 0) ++tile_rows;
  }
#ifdef TTB_TUNING
  const int dbg = ep.relu >> 8;  // timing experiments (wrong results): 1 no global stores, 2 no staging, 4 no TMEM load
#else
  constexpr int dbg = 0;
#endif
  float4* const out4 = reinterpret_cast<float4*>(out);
  const float4* const accum4 = reinterpret_cast<const float4*>(ep.accum);
  const bool relu = (ep.relu & 1) != 0;
  // The rows of the tensor that is added (residual / pending gradient) are fetched a whole column block at a time, eight
  // independent loads per lane, and - for tiles of >= 128 columns (one CTA per SM: registers to spare) - one block AHEAD of
  // the block being stored, so their HBM latency overlaps the TMEM load, the staging and the stores of the previous block.
  // Loaded one by one next to the store that consumes them (8 dependent round trips per block) a ResNet-50 dgrad with a
  // pending gradient ran at 16 TFLOP/s.
  constexpr bool kAhead = BN >= 128;
  auto load_accum = [&](float4 (&a)[8], int cb) {
    const int col0 = n0 + cb * 32;
    const bool ok = col0 + c4 * 4 < n_total;
#pragma unroll
    for (int it = 0; it < 8; ++it)
      a[it] = (off4[it] >= 0 && ok) ? ld_f4_stream(reinterpret_cast<const float*>(accum4 + ((int64_t)off4[it] + (col0 >> 2) + c4)))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  float4 acc[8], acc_next[8];
  if (kAhead && accum4) load_accum(acc_next, 0);
  // BatchNorm-backward sums (Epilogue::bn_x): the BatchNorm's input at the positions this lane stores, eight independent
  // loads per column block issued before the TMEM load, and the lane's four per-channel constants
  constexpr bool bwd = STATS == 2;
  const float4* const bnx4 = reinterpret_cast<const float4*>(ep.bn_x);
  // (fully unrolled: the statistics live in registers indexed by cb)
#pragma unroll
  for (int cb = 0; cb < BN / 32; ++cb) {
    const int col0 = n0 + cb * 32;
    if (col0 >= n_total) break;  // warp-uniform
    if (accum4) {
      if (kAhead) {
#pragma unroll
        for (int it = 0; it < 8; ++it) acc[it] = acc_next[it];
        if (cb + 1 < BN / 32) load_accum(acc_next, cb + 1);  // (columns past n_total are not loaded: see load_accum)
      } else {
        load_accum(acc, cb);
      }
    }
    float4 xs[8];
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rsc = mu, rsh = mu;
    if (STATS && bwd) {
      const bool okc = col0 + c4 * 4 < n_total;
#pragma unroll
      for (int it = 0; it < 8; ++it)
        xs[it] = (off4[it] >= 0 && okc) ? ld_f4_stream(reinterpret_cast<const float*>(bnx4 + ((int64_t)off4[it] + (col0 >> 2) + c4)))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
      if (okc) {
        mu = ld_f4(ep.bn_mean + col0 + c4 * 4);
        if (ep.bn_rscale) {
          rsc = ld_f4(ep.bn_rscale + col0 + c4 * 4);
          rsh = ld_f4(ep.bn_rshift + col0 + c4 * 4);
        }
      }
    }
    uint32_t r[32];
    if (!(dbg & 4)) {
      ptx::tmem_ld_32x32(tmem_acc + ((uint32_t)(lane_block * 32) << 16) + (uint32_t)(cb * 32), r);
      ptx::tmem_ld_wait();
    }
    if (ep.scale) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < n_total) r[j] = __float_as_uint(__uint_as_float(r[j]) * __ldg(ep.scale + col0 + j));
    }
    if (ep.bias) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < n_total) r[j] = __float_as_uint(__uint_as_float(r[j]) + __ldg(ep.bias + col0 + j));
    }
    // own row -> smem (8 x float4), then 4 rows x 128 B per store instruction
    float4* srow = reinterpret_cast<float4*>(st + lane * kStagePitch);
    if (!(dbg & 2)) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        srow[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                              __uint_as_float(r[4 * j + 3]));
    }
    __syncwarp();
    const bool col_ok = col0 + c4 * 4 < n_total;
    const int colq = (col0 >> 2) + c4;
    bool seen = false;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + sub;
      float4 v = (dbg & 2) ? make_float4(__uint_as_float(r[it]), 0.f, 0.f, 0.f)
                           : *reinterpret_cast<const float4*>(st + rr * kStagePitch + c4 * 4);
      if (off4[it] >= 0 && col_ok) {
        const int64_t o4 = (int64_t)off4[it] + colq;
        if (accum4) {
          const float4 a = acc[it];
          v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        }
        if (relu) {
          v.x = v.x < 0.f ? 0.f : v.x; v.y = v.y < 0.f ? 0.f : v.y;
          v.z = v.z < 0.f ? 0.f : v.z; v.w = v.w < 0.f ? 0.f : v.w;
        }
        if (!(dbg & 1)) out4[o4] = v;
        if (STATS && bwd) {  // (es.k stays 0: the flush then passes s0 / s1 through unchanged)
          const float4 xv = xs[it];
          const float cx = xv.x - mu.x, cy = xv.y - mu.y, cz = xv.z - mu.z, cw = xv.w - mu.w;
          float4 g = v;
          if (ep.bn_rscale) {  // exactly forward's expression (bn_apply4): bit-identical mask decisions
            g.x = fmaf(cx, rsc.x, rsh.x) > 0.f ? g.x : 0.f; g.y = fmaf(cy, rsc.y, rsh.y) > 0.f ? g.y : 0.f;
            g.z = fmaf(cz, rsc.z, rsh.z) > 0.f ? g.z : 0.f; g.w = fmaf(cw, rsc.w, rsh.w) > 0.f ? g.w : 0.f;
          }
          es.s0[cb][0] += g.x; es.s0[cb][1] += g.y; es.s0[cb][2] += g.z; es.s0[cb][3] += g.w;
          es.s1[cb][0] = fmaf(g.x, cx, es.s1[cb][0]); es.s1[cb][1] = fmaf(g.y, cy, es.s1[cb][1]);
          es.s1[cb][2] = fmaf(g.z, cz, es.s1[cb][2]); es.s1[cb][3] = fmaf(g.w, cw, es.s1[cb][3]);
        } else if (STATS) {
          if (fresh && !seen) { es.k[cb][0] = v.x; es.k[cb][1] = v.y; es.k[cb][2] = v.z; es.k[cb][3] = v.w; }
          seen = true;
          const float dx = v.x - es.k[cb][0], dy = v.y - es.k[cb][1], dz = v.z - es.k[cb][2], dw = v.w - es.k[cb][3];
          es.s0[cb][0] += dx; es.s0[cb][1] += dy; es.s0[cb][2] += dz; es.s0[cb][3] += dw;
          es.s1[cb][0] = fmaf(dx, dx, es.s1[cb][0]); es.s1[cb][1] = fmaf(dy, dy, es.s1[cb][1]);
          es.s1[cb][2] = fmaf(dz, dz, es.s1[cb][2]); es.s1[cb][3] = fmaf(dw, dw, es.s1[cb][3]);
        }
      }
    }
    __syncwarp();
  }
  if (STATS) es.cnt += tile_rows;
}

// End of a CTA's tile loop (all four epilogue warps call it): lane sums -> plain double sums -> warp -> CTA, then one row
// segment of the partial buffer: row[c_total*0 + n0 + col] = sum(v), row[c_total + n0 + col] = sum(v*v) over every output
// row this CTA stored.  `stage_smem`: the CTA's epilogue staging area (each warp reuses its own slice: BN*2 doubles
// <= 32*kStagePitch floats); `bar_id`: a named barrier reserved for the 128 epilogue threads.
template <int BN>
__device__ __forceinline__ void epilogue_stats_flush(const EpiStats<BN>& es, float* stage_smem, int ep_warp, int bar_id,
                                                     double* __restrict__ row, int n0, int n_total) {
  static_assert(BN * 2 * sizeof(double) <= 32 * kStagePitch * sizeof(float), "per-warp statistics must fit its staging slice");
  const int lane = threadIdx.x & 31;
  const int sub = lane >> 3, c4 = lane & 7;
  double* mine = reinterpret_cast<double*>(stage_smem + ep_warp * (32 * kStagePitch));
  const double n = (double)es.cnt;
#pragma unroll
  for (int cb = 0; cb < BN / 32; ++cb) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double k = (double)es.k[cb][j], t0 = (double)es.s0[cb][j], t1 = (double)es.s1[cb][j];
      double d0 = t0 + n * k;
      double d1 = t1 + 2.0 * k * t0 + n * k * k;
      d0 += __shfl_xor_sync(0xffffffffu, d0, 8);
      d1 += __shfl_xor_sync(0xffffffffu, d1, 8);
      d0 += __shfl_xor_sync(0xffffffffu, d0, 16);
      d1 += __shfl_xor_sync(0xffffffffu, d1, 16);
      if (sub == 0) {
        mine[(cb * 32 + c4 * 4 + j) * 2 + 0] = d0;
        mine[(cb * 32 + c4 * 4 + j) * 2 + 1] = d1;
      }
    }
  }
  asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
  const int e = ep_warp * 32 + lane;
  for (int col = e; col < BN; col += 128) {
    if (n0 + col >= n_total) break;
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {  // fixed order: deterministic
      const double* p = reinterpret_cast<const double*>(stage_smem + w * (32 * kStagePitch));
      t0 += p[col * 2 + 0];
      t1 += p[col * 2 + 1];
    }
    row[n0 + col] = t0;
    row[n_total + n0 + col] = t1;
  }
}

}  // namespace ttb
