// BatchNorm2d forward/backward on NHWC activations viewed as [M rows][C channels] (HBM-bound).
//
// Reference: BatchNorm.forward/backward, /root/reference/src/tortto/autograd/grad_nn.py:909-989.
// The reference makes ~8 full-tensor passes forward and ~10 backward; here forward is 2 reads + 1 write
// (column reduce, then one fused normalise[+ReLU] pass) and backward is 4 reads + 1 write (column reduce of
// sum(dy), sum(dy*(x-mean)); then one fused dx pass).
//
// Column reductions: thread -> fixed channel quad (float4 = 4 channels), rows strided across the block and
// the grid; per-thread fp32 partials over its share of one wave's rows (tens to a few hundred), block tree in shared memory, one double
// partial per (chunk, channel) written to HBM, and a tiny second kernel sums chunks in double in a FIXED
// order (deterministic; no atomics).  The [2][C] double sums are what a data-parallel run all-reduces.
#include <cuda_bf16.h>

#include "common.cuh"
#include "bn_finalize.cuh"

namespace ttb {

constexpr int kBnThreads = 256;
constexpr int kBnWaveCtas = 8;  // CTAs per SM the column reductions are sized for (measured choice)

struct ColGeom {
  int tx;        // threads along channel quads (power of two <= 256)
  int ty;        // threads along rows = 256 / tx
  int qblocks;   // gridDim.x = ceil(cq / tx)
  int chunks;    // gridDim.y
  int64_t rows_per_chunk;
};

// resident CTAs per SM of the column-reduce kernels (smallest over the variants), queried once: the grid is sized to
// ONE full wave so no SM idles through a partial second wave
static int col_reduce_ctas_per_sm();

static ColGeom col_geom(int64_t m, int c) {
  ColGeom g;
  int cq = (c + 3) / 4;
  int tx = 1;
  while (tx < cq && tx < kBnThreads) tx <<= 1;
  g.tx = tx;
  g.ty = kBnThreads / tx;
  g.qblocks = (cq + tx - 1) / tx;
  // Spread the rows over ONE full wave of CTAs (SMs x resident CTAs per SM) whenever there are enough rows: the
  // kernel is a chain of dependent load batches per thread, so its latency is (rows per thread / unroll) x DRAM
  // latency - small tensors want many short threads, not few long ones.  At least kMinRowsPerThread rows per thread.
  constexpr int kMinRowsPerThread = 4;
  // kBnWaveCtas CTAs per SM are enough to saturate HBM (8 x float4 in flight per thread) and keep the partial buffer the
  // finalize kernel has to sum small: its latency is (chunks / 128) dependent L2 round trips
  int per_sm = col_reduce_ctas_per_sm();
  const int want = tuning_knob("TTB_BN_CTAS_PER_SM", kBnWaveCtas);
  if (want > 0 && want < per_sm) per_sm = want;
  int64_t cap = (int64_t)sm_count() * per_sm / g.qblocks;
  // ... but keep the partial buffer [chunks][2][C] doubles under ~1 MB so the finalize kernel stays a few microseconds
  const int64_t cap_bytes = (int64_t)(1 << 20) / ((int64_t)2 * c * 8);
  if (cap > cap_bytes) cap = cap_bytes;
  if (cap < 1) cap = 1;
  int64_t rows_per_chunk = ceil_div(m, cap);
  rows_per_chunk = ceil_div(rows_per_chunk, g.ty) * g.ty;
  if (rows_per_chunk < (int64_t)g.ty * kMinRowsPerThread) rows_per_chunk = (int64_t)g.ty * kMinRowsPerThread;
  int64_t chunks = ceil_div(m, rows_per_chunk);
  if (chunks < 1) chunks = 1;
  g.chunks = (int)chunks;
  g.rows_per_chunk = rows_per_chunk;
  return g;
}

// MODE 0: s0 = sum(a), s1 = sum(a*a)                          (forward statistics; a = x)
//         accumulated per thread as sum(a-K), sum((a-K)^2) with K = the channel's value in row 0 of the tensor (the same
//         K in every CTA) and converted to the unshifted double sums when the block writes its partial: fp32
//         accumulation of x^2 loses the variance when |mean| >> sd (sum x^2 - n*mean^2 cancels); shifted sums do not,
//         and the reference is two-pass (xp.mean, xp.var: grad_nn.py:923-924).
// MODE 1: s0 = sum(g), s1 = sum(g*(b-mean)), g = a or masked  (backward; a = dy, b = x, mask = relu_out > 0)
// MODE 2: MODE 0 over the sum a + b, which is also written to `sum_out` (must NOT alias a or b - the shift K is row 0 of
//         the inputs, read by every CTA): the residual `Add` and the
//         statistics pass of the BatchNorm that reads the sum, in one pass (2 reads + 1 write instead of 3 reads + 1 write)
template <int MODE, bool VEC>
__global__ void __launch_bounds__(kBnThreads)
col_reduce_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ mean,
                  const float* __restrict__ mask, const float* __restrict__ rscale, const float* __restrict__ rshift,
                  int64_t m, int c, int tx_n, int64_t rows_per_chunk, double* __restrict__ partials, float* __restrict__ sum_out) {
  pdl_entry();
  // ReLU mask of the fused BatchNorm+ReLU node: either read (mask = the ReLU output) or RECOMPUTED from x with the
  // forward pass's own mean / scale / beta (rscale != null): fmaf(x - mean, scale, beta) > 0 is bit-for-bit what forward tested,
  // and it saves one of the three reads of this pass
  extern __shared__ float4 sm[];  // [2][ty][tx]
  const int tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n, ty_n = kBnThreads / tx_n;
  const int quad = blockIdx.x * tx_n + tx;
  const int ch = quad * 4;
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  if (ch < c) {
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rsc = mu, rsh = mu;
    const bool recompute = MODE == 1 && rscale != nullptr;
    if (MODE == 0 || MODE == 2) {  // mu = the shift K (row 0 of the tensor)
      if (VEC) {
        mu = ld_f4(a + ch);
        if (MODE == 2) { const float4 t = ld_f4(b + ch); mu.x += t.x; mu.y += t.y; mu.z += t.z; mu.w += t.w; }
      } else {
        float* pm = &mu.x;
        for (int j = 0; j < 4 && ch + j < c; ++j) pm[j] = MODE == 2 ? a[ch + j] + b[ch + j] : a[ch + j];
      }
    }
    if (MODE == 1) {
      if (VEC) {
        mu = ld_f4(mean + ch);
        if (recompute) { rsc = ld_f4(rscale + ch); rsh = ld_f4(rshift + ch); }
      } else {
        float* pm = &mu.x; float* ps = &rsc.x; float* ph = &rsh.x;
        for (int j = 0; j < 4 && ch + j < c; ++j) {
          pm[j] = mean[ch + j];
          if (recompute) { ps[j] = rscale[ch + j]; ph[j] = rshift[ch + j]; }
        }
      }
    }
    const bool masked = mask != nullptr || recompute;
    int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
    int64_t r1 = r0 + rows_per_chunk < m ? r0 + rows_per_chunk : m;
    auto accumulate = [&](float4 va, float4 vb, float4 vm) {
      if (MODE == 0 || MODE == 2) {
        va.x -= mu.x; va.y -= mu.y; va.z -= mu.z; va.w -= mu.w;
        s0.x += va.x; s0.y += va.y; s0.z += va.z; s0.w += va.w;
        s1.x = fmaf(va.x, va.x, s1.x); s1.y = fmaf(va.y, va.y, s1.y);
        s1.z = fmaf(va.z, va.z, s1.z); s1.w = fmaf(va.w, va.w, s1.w);
      } else {
        if (recompute) {  // exactly forward's expression: fmaf(x - mean, scale, beta)
          vm.x = fmaf(vb.x - mu.x, rsc.x, rsh.x); vm.y = fmaf(vb.y - mu.y, rsc.y, rsh.y);
          vm.z = fmaf(vb.z - mu.z, rsc.z, rsh.z); vm.w = fmaf(vb.w - mu.w, rsc.w, rsh.w);
        }
        if (masked) {
          va.x = vm.x > 0.f ? va.x : 0.f; va.y = vm.y > 0.f ? va.y : 0.f;
          va.z = vm.z > 0.f ? va.z : 0.f; va.w = vm.w > 0.f ? va.w : 0.f;
        }
        s0.x += va.x; s0.y += va.y; s0.z += va.z; s0.w += va.w;
        s1.x = fmaf(va.x, vb.x - mu.x, s1.x); s1.y = fmaf(va.y, vb.y - mu.y, s1.y);
        s1.z = fmaf(va.z, vb.z - mu.z, s1.z); s1.w = fmaf(va.w, vb.w - mu.w, s1.w);
      }
    };
    int64_t r = r0 + ty;
    if (VEC) {
      // U rows per iteration: all loads of the batch are issued before the first use (memory-level parallelism);
      // 8 float4 in flight per thread for the single-input statistics pass, 4 x 3 for the three-input backward pass
      constexpr int U = MODE == 0 ? 8 : 4;
      for (; r + (U - 1) * ty_n < r1; r += U * ty_n) {
        float4 va[U], vb[U], vm[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t off = (r + u * ty_n) * c + ch;
          va[u] = ld_f4_stream(a + off);
          if (MODE == 2) vb[u] = ld_f4_stream(b + off);
          if (MODE == 1) {
            vb[u] = ld_f4_stream(b + off);
            if (mask) vm[u] = ld_f4_stream(mask + off);
          }
        }
        if (MODE == 2) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            va[u].x += vb[u].x; va[u].y += vb[u].y; va[u].z += vb[u].z; va[u].w += vb[u].w;
            st_f4(sum_out + (r + u * ty_n) * c + ch, va[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) accumulate(va[u], vb[u], vm[u]);
      }
    }
    for (; r < r1; r += ty_n) {
      const int64_t off = r * c + ch;
      float4 va, vb = make_float4(0.f, 0.f, 0.f, 0.f), vm = make_float4(1.f, 1.f, 1.f, 1.f);
      if (VEC) {
        va = ld_f4_stream(a + off);
        if (MODE == 2) {
          vb = ld_f4_stream(b + off);
          va.x += vb.x; va.y += vb.y; va.z += vb.z; va.w += vb.w;
          st_f4(sum_out + off, va);
        }
        if (MODE == 1) {
          vb = ld_f4_stream(b + off);
          if (mask) vm = ld_f4_stream(mask + off);
        }
      } else {
        va = make_float4(0.f, 0.f, 0.f, 0.f);
        float* pa = &va.x; float* pb = &vb.x; float* pm = &vm.x;
        for (int j = 0; j < 4 && ch + j < c; ++j) {
          pa[j] = a[off + j];
          if (MODE == 2) {
            pa[j] += b[off + j];
            sum_out[off + j] = pa[j];
          }
          if (MODE == 1) {
            pb[j] = b[off + j];
            if (mask) pm[j] = mask[off + j];
          }
        }
      }
      accumulate(va, vb, vm);
    }
  }
  sm[ty * tx_n + tx] = s0;
  sm[(ty_n + ty) * tx_n + tx] = s1;
  __syncthreads();
  // ty == 0 row of threads finishes the block reduction in double, in fixed order
  if (ty == 0 && ch < c) {
    double d0[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0};
    for (int j = 0; j < ty_n; ++j) {
      float4 p = sm[j * tx_n + tx], q = sm[(ty_n + j) * tx_n + tx];
      d0[0] += p.x; d0[1] += p.y; d0[2] += p.z; d0[3] += p.w;
      d1[0] += q.x; d1[1] += q.y; d1[2] += q.z; d1[3] += q.w;
    }
    double* out = partials + (int64_t)blockIdx.y * 2 * c;
    if (MODE == 0 || MODE == 2) {  // shifted -> plain sums, in double: sum x = S0 + n K, sum x^2 = S1 + 2 K S0 + n K^2
      int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
      int64_t r1 = r0 + rows_per_chunk < m ? r0 + rows_per_chunk : m;
      const double nrows = r1 > r0 ? (double)(r1 - r0) : 0.0;
      for (int j = 0; j < 4 && ch + j < c; ++j) {
        const double k = MODE == 2 ? (double)(a[ch + j] + b[ch + j]) : (double)a[ch + j];
        const double t0 = d0[j], t1 = d1[j];
        d0[j] = t0 + nrows * k;
        d1[j] = t1 + 2.0 * k * t0 + nrows * k * k;
      }
    }
    for (int j = 0; j < 4 && ch + j < c; ++j) {
      out[ch + j] = d0[j];
      out[c + ch + j] = d1[j];
    }
  }
}

// Sum of the per-chunk partials of one value: blockDim = (32 values, kLanes chunk lanes); fixed order => deterministic.
// Four independent accumulators per thread keep four L2 loads in flight.  Returns the total where threadIdx.y == 0.
constexpr int kLanes = 32;
__device__ __forceinline__ double chunk_sum(const double* __restrict__ partials, int num_chunks, int stride, int idx,
                                            bool valid, double (*sm)[33]) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (valid) {
    int k = threadIdx.y;
    for (; k + 3 * kLanes < num_chunks; k += 4 * kLanes) {
      s0 += partials[(int64_t)k * stride + idx];
      s1 += partials[(int64_t)(k + kLanes) * stride + idx];
      s2 += partials[(int64_t)(k + 2 * kLanes) * stride + idx];
      s3 += partials[(int64_t)(k + 3 * kLanes) * stride + idx];
    }
    for (; k < num_chunks; k += kLanes) s0 += partials[(int64_t)k * stride + idx];
  }
  sm[threadIdx.y][threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  double t = 0.0;
  if (threadIdx.y == 0)
    for (int j = 0; j < kLanes; ++j) t += sm[j][threadIdx.x];
  __syncthreads();
  return t;
}

static int col_reduce_ctas_per_sm() {
  static int cached = 0;
  if (cached == 0) {
    int best = 8;
    const size_t smem = sizeof(float4) * 2 * kBnThreads;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, col_reduce_kernel<0, true>, kBnThreads, smem) == cudaSuccess && n > 0 && n < best) best = n;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, col_reduce_kernel<1, true>, kBnThreads, smem) == cudaSuccess && n > 0 && n < best) best = n;
    cached = best;
  }
  return cached;
}

__global__ void __launch_bounds__(1024)
reduce_partials_kernel(const double* __restrict__ partials, int num_chunks, int c2, double* __restrict__ sums) {
  pdl_entry();
  __shared__ double sm[kLanes][33];
  int i = blockIdx.x * 32 + threadIdx.x;
  double t = chunk_sum(partials, num_chunks, c2, i, i < c2, sm);
  if (threadIdx.y == 0 && i < c2) sums[i] = t;
}

// Both sums of the per-chunk partials [chunks][2][c] for 8 channels per block: blockDim = (8 channels, 128 chunk lanes).
// The finalize kernels sit between two HBM passes of every BatchNorm layer (34 launches per ResNet-18 step), so their
// latency is on the critical path: many small blocks (c / 8) and 128 lanes per channel make it ONE round of independent L2
// loads per thread (a few hundred chunks) instead of three dependent rounds per sum; fixed-order tree => deterministic.
constexpr int kFinCh = 8, kFinLanes = 128;
__device__ __forceinline__ void chunk_sum2(const double* __restrict__ partials, int num_chunks, int c, int i, bool valid,
                                           double (*sm)[2][kFinCh], double* out0, double* out1) {
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
  if (valid) {
    int k = threadIdx.y;
    for (; k + kFinLanes < num_chunks; k += 2 * kFinLanes) {
      const double* p = partials + (int64_t)k * 2 * c + i;
      const double* q = p + (int64_t)kFinLanes * 2 * c;
      a0 += p[0]; a1 += p[c];
      b0 += q[0]; b1 += q[c];
    }
    if (k < num_chunks) {
      const double* p = partials + (int64_t)k * 2 * c + i;
      a0 += p[0]; a1 += p[c];
    }
  }
  sm[threadIdx.y][0][threadIdx.x] = a0 + b0;
  sm[threadIdx.y][1][threadIdx.x] = a1 + b1;
  __syncthreads();
  for (int o = kFinLanes / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.y < o) {
      sm[threadIdx.y][0][threadIdx.x] += sm[threadIdx.y + o][0][threadIdx.x];
      sm[threadIdx.y][1][threadIdx.x] += sm[threadIdx.y + o][1][threadIdx.x];
    }
    __syncthreads();
  }
  *out0 = sm[0][0][threadIdx.x];
  *out1 = sm[0][1][threadIdx.x];
}

__global__ void __launch_bounds__(kFinCh * kFinLanes)
bn_finalize_kernel(const double* __restrict__ partials, int num_chunks, int c, BnFwdFinalize fin) {
  pdl_entry();
  __shared__ double sm[kFinLanes][2][kFinCh];
  int i = blockIdx.x * kFinCh + threadIdx.x;
  double s0, s1;
  chunk_sum2(partials, num_chunks, c, i, i < c, sm, &s0, &s1);
  if (threadIdx.y != 0 || i >= c) return;
  fin(i, s0, s1);
}

__global__ void bn_prepare_eval_kernel(const float* __restrict__ mean_in, const float* __restrict__ var_in, int c,
                                       float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float* mean, float* var_eps, float* sd, float* scale, float* shift) {
  pdl_entry();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  float mu = mean_in[i];
  float ve = __fadd_rn(var_in[i], eps);
  float s = sqrtf(ve);
  mean[i] = mu; var_eps[i] = ve; sd[i] = s;
  scale[i] = (gamma ? gamma[i] : 1.f) / s;
  shift[i] = beta ? beta[i] : 0.f;  // the additive term of y = (x - mean)*scale + beta
}

// eval-mode BatchNorm folded into the epilogue of the convolution in front of it:
//   bn(conv + b) = (conv + b - mean) * gamma / sqrt(var + eps) + beta = conv * scale + bias
__global__ void bn_fold_eval_kernel(const float* __restrict__ mean, const float* __restrict__ var, int c, float eps,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ conv_bias, float* scale, float* bias) {
  pdl_entry();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const float s = (gamma ? gamma[i] : 1.f) / sqrtf(__fadd_rn(var[i], eps));
  scale[i] = s;
  bias[i] = fmaf((conv_bias ? conv_bias[i] : 0.f) - mean[i], s, beta ? beta[i] : 0.f);
}

__global__ void __launch_bounds__(kFinCh * kFinLanes)
bn_bwd_finalize_kernel(const double* __restrict__ partials, int num_chunks, int c, BnBwdFinalize fin) {
  pdl_entry();
  __shared__ double sm[kFinLanes][2][kFinCh];
  int i = blockIdx.x * kFinCh + threadIdx.x;
  double sdy, sdyx;
  chunk_sum2(partials, num_chunks, c, i, i < c, sm, &sdy, &sdyx);
  if (threadIdx.y != 0 || i >= c) return;
  fin(i, sdy, sdyx);
}

// y = (x - mean)*scale + beta (+ReLU) - the reference's own order of operations (grad_nn.py:942-959: subtract the mean,
// divide by sd, times gamma, plus beta) with gamma/sd folded into `scale`; unlike x*scale + (beta - mean*scale) it does
// not lose digits when |mean| >> sd.  One float4 = 4 channels; channel quad = i % cq.  Two independent float4 streams
// per thread per iteration (grid-stride, 2x unrolled) keep more loads in flight.  `yh` (may be null): the same values
// rounded to bf16, co-written for the bf16 tensor-core path (the next convolution reads them instead of converting).
template <bool RELU>
__device__ __forceinline__ float4 bn_apply4(float4 v, const float* __restrict__ mean, const float* __restrict__ scale,
                                            const float* __restrict__ beta, int q) {
  float4 mu = ld_f4(mean + 4 * q), sc = ld_f4(scale + 4 * q), sh = ld_f4(beta + 4 * q);
  v.x = fmaf(v.x - mu.x, sc.x, sh.x); v.y = fmaf(v.y - mu.y, sc.y, sh.y);
  v.z = fmaf(v.z - mu.z, sc.z, sh.z); v.w = fmaf(v.w - mu.w, sc.w, sh.w);
  if (RELU) {
    v.x = v.x < 0.f ? 0.f : v.x; v.y = v.y < 0.f ? 0.f : v.y;
    v.z = v.z < 0.f ? 0.f : v.z; v.w = v.w < 0.f ? 0.f : v.w;
  }
  return v;
}

template <bool RELU, bool SHADOW>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, float* __restrict__ y, __nv_bfloat16* __restrict__ yh, int64_t n4, int cq,
                const float* __restrict__ mean, const float* __restrict__ scale, const float* __restrict__ beta) {
  pdl_entry();
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < n4; i += 2 * stride) {
    float4 a = ld_f4_stream(x + 4 * i), b = ld_f4_stream(x + 4 * (i + stride));
    a = bn_apply4<RELU>(a, mean, scale, beta, (int)(i % cq));
    b = bn_apply4<RELU>(b, mean, scale, beta, (int)((i + stride) % cq));
    st_f4(y + 4 * i, a);
    st_f4(y + 4 * (i + stride), b);
    if (SHADOW) {
      st_bf16x4(yh + 4 * i, a);
      st_bf16x4(yh + 4 * (i + stride), b);
    }
  }
  for (; i < n4; i += stride) {
    float4 a = bn_apply4<RELU>(ld_f4_stream(x + 4 * i), mean, scale, beta, (int)(i % cq));
    st_f4(y + 4 * i, a);
    if (SHADOW) st_bf16x4(yh + 4 * i, a);
  }
}

template <bool RELU>
__global__ void bn_apply_scalar_kernel(const float* __restrict__ x, float* __restrict__ y, __nv_bfloat16* __restrict__ yh,
                                       int64_t n, int c, const float* __restrict__ mean, const float* __restrict__ scale,
                                       const float* __restrict__ beta) {
  pdl_entry();
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int ch = (int)(i % c);
    float v = fmaf(x[i] - mean[ch], scale[ch], beta[ch]);
    if (RELU) v = v < 0.f ? 0.f : v;
    y[i] = v;
    if (yh) yh[i] = __float2bfloat16_rn(v);
  }
}

// y = bn(x) + residual (+ReLU): the tail of a post-activation residual block (BatchNorm -> `out += identity` -> ReLU,
// reference examples/resnet/resnet50_finetune: Bottleneck.forward) as ONE 2R+1W pass instead of the three passes
// (1R+1W, 2R+1W, 1R+1W) of the separate operators.
template <bool RELU, bool SHADOW>
__global__ void __launch_bounds__(256)
bn_apply_add_kernel(const float* __restrict__ x, const float* __restrict__ res, float* __restrict__ y,
                    __nv_bfloat16* __restrict__ yh, int64_t n4, int cq, const float* __restrict__ mean,
                    const float* __restrict__ scale, const float* __restrict__ beta) {
  pdl_entry();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 r = ld_f4_stream(res + 4 * i);
    float4 a = bn_apply4<false>(ld_f4_stream(x + 4 * i), mean, scale, beta, (int)(i % cq));
    a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
    if (RELU) {
      a.x = a.x < 0.f ? 0.f : a.x; a.y = a.y < 0.f ? 0.f : a.y;
      a.z = a.z < 0.f ? 0.f : a.z; a.w = a.w < 0.f ? 0.f : a.w;
    }
    st_f4(y + 4 * i, a);
    if (SHADOW) st_bf16x4(yh + 4 * i, a);
  }
}

// dx = c1*(g - c2 - (x-mean)*c3) [+ accum], g = dy (masked by relu_out > 0 when given).  `accum`: a gradient that already
// reached the same tensor through another branch (the residual shortcut) - added here instead of by a separate kernel.
// MASK: 0 none, 1 read the ReLU output, 2 recompute it from x (fmaf(x - mean, rscale, rbeta), see col_reduce_kernel).
// `dxh` (may be null): dx rounded to bf16, co-written for the bf16 tensor-core path (dgrad / wgrad of the producing
// convolution read it as dY).
template <int MASK, bool SHADOW>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                    const float* __restrict__ mask, const float* __restrict__ rscale, const float* __restrict__ rshift,
                    const float* __restrict__ coef, const float* __restrict__ accum, float* __restrict__ dx,
                    __nv_bfloat16* __restrict__ dxh, int64_t n4, int cq, int c) {
  pdl_entry();
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    int q = (int)(i % cq);
    float4 g = ld_f4_stream(dy + 4 * i), v = ld_f4_stream(x + 4 * i);
    float4 mu = ld_f4(mean + 4 * q);
    if (MASK) {
      float4 o;
      if (MASK == 1) {
        o = ld_f4_stream(mask + 4 * i);
      } else {
        float4 sc = ld_f4(rscale + 4 * q), sh = ld_f4(rshift + 4 * q);
        o.x = fmaf(v.x - mu.x, sc.x, sh.x); o.y = fmaf(v.y - mu.y, sc.y, sh.y);
        o.z = fmaf(v.z - mu.z, sc.z, sh.z); o.w = fmaf(v.w - mu.w, sc.w, sh.w);
      }
      g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f;
      g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    }
    float4 c1 = ld_f4(coef + 4 * q), c2 = ld_f4(coef + c + 4 * q), c3 = ld_f4(coef + 2 * c + 4 * q);
    float4 r;
    r.x = c1.x * (g.x - c2.x - (v.x - mu.x) * c3.x);
    r.y = c1.y * (g.y - c2.y - (v.y - mu.y) * c3.y);
    r.z = c1.z * (g.z - c2.z - (v.z - mu.z) * c3.z);
    r.w = c1.w * (g.w - c2.w - (v.w - mu.w) * c3.w);
    if (accum) {
      float4 a = ld_f4_stream(accum + 4 * i);
      r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
    }
    st_f4(dx + 4 * i, r);
    if (SHADOW) st_bf16x4(dxh + 4 * i, r);
  }
}

template <int MASK>
__global__ void bn_bwd_apply_scalar_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                           const float* __restrict__ mean, const float* __restrict__ mask,
                                           const float* __restrict__ rscale, const float* __restrict__ rshift,
                                           const float* __restrict__ coef, const float* __restrict__ accum,
                                           float* __restrict__ dx, __nv_bfloat16* __restrict__ dxh, int64_t n, int c) {
  pdl_entry();
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int ch = (int)(i % c);
    float g = dy[i];
    if (MASK == 1) g = mask[i] > 0.f ? g : 0.f;
    if (MASK == 2) g = fmaf(x[i] - mean[ch], rscale[ch], rshift[ch]) > 0.f ? g : 0.f;
    float r = coef[ch] * (g - coef[c + ch] - (x[i] - mean[ch]) * coef[2 * c + ch]);
    if (accum) r += accum[i];
    dx[i] = r;
    if (dxh) dxh[i] = __float2bfloat16_rn(r);
  }
}

static bool a16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int MODE>
static int launch_col_reduce(const float* a, const float* b, const float* mean, const float* mask, const float* rscale,
                             const float* rshift, int64_t m, int c, double* partials, int num_chunks, cudaStream_t st,
                             const char* what, float* sum_out = nullptr) {
  if (m <= 0 || c <= 0) return 0;
  ColGeom g = col_geom(m, c);
  TTB_REQUIRE(num_chunks == g.chunks, "%s: num_chunks=%d but ttb_bn_num_chunks gives %d", what, num_chunks, g.chunks);
  bool vec = (c % 4 == 0) && a16(a) &&
             (MODE == 0 || (MODE == 2 && a16(b) && a16(sum_out)) ||
              (MODE == 1 && a16(b) && a16(mean) && (!mask || a16(mask)) && (!rscale || (a16(rscale) && a16(rshift)))));
  dim3 grid(g.qblocks, g.chunks);
  size_t smem = sizeof(float4) * 2 * kBnThreads;
  if (vec)
    launch_k(col_reduce_kernel<MODE, true>, grid, kBnThreads, smem, st, a, b, mean, mask, rscale, rshift, m, c, g.tx, g.rows_per_chunk, partials, sum_out);
  else
    launch_k(col_reduce_kernel<MODE, false>, grid, kBnThreads, smem, st, a, b, mean, mask, rscale, rshift, m, c, g.tx, g.rows_per_chunk, partials, sum_out);
  return check_launch(what);
}

// out[i] = sum over chunks of partials[chunk][i] (i < c: the plain column sums of a MODE 0 reduction), as float
__global__ void __launch_bounds__(1024)
colsum_finish_kernel(const double* __restrict__ partials, int num_chunks, int c, float* __restrict__ out) {
  pdl_entry();
  __shared__ double sm[kLanes][33];
  int i = blockIdx.x * 32 + threadIdx.x;
  double t = chunk_sum(partials, num_chunks, 2 * c, i, i < c, sm);
  if (threadIdx.y == 0 && i < c) out[i] = (float)t;
}

// column sums of x[m][c] -> out[c] (float) with the BatchNorm statistics kernel (float4 loads, one wave of CTAs, double
// partials, fixed order); the partial buffer is a stream-ordered temporary.  Used for conv bias gradients.
int column_sums(const float* x, int64_t m, int c, float* out, cudaStream_t st) {
  if (c <= 0) return 0;
  if (m <= 0) {
    cudaMemsetAsync(out, 0, (size_t)c * sizeof(float), st);
    return 0;
  }
  const int chunks = col_geom(m, c).chunks;
  double* partials = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&partials, (size_t)chunks * 2 * c * sizeof(double), st);
  if (e != cudaSuccess) {
    set_error("column_sums: cudaMallocAsync failed: %s", cudaGetErrorString(e));
    return 1;
  }
  int rc = launch_col_reduce<0>(x, nullptr, nullptr, nullptr, nullptr, nullptr, m, c, partials, chunks, st, "column_sums");
  if (!rc) {
    launch_k(colsum_finish_kernel, (c + 31) / 32, dim3(32, kLanes), 0, st, partials, chunks, c, out);
    rc = check_launch("column_sums(finish)");
  }
  cudaFreeAsync(partials, st);
  return rc;
}

}  // namespace ttb

using namespace ttb;

extern "C" {

int ttb_bn_num_chunks(int64_t m, int c) {
  if (m <= 0 || c <= 0) return 1;
  return col_geom(m, c).chunks;
}

int ttb_bn_stats(const float* x, int64_t m, int c, double* partials, int num_chunks, void* stream) {
  return launch_col_reduce<0>(x, nullptr, nullptr, nullptr, nullptr, nullptr, m, c, partials, num_chunks, as_stream(stream),
                              "bn_stats");
}

int ttb_add_bn_stats(const float* a, const float* b, float* out, int64_t m, int c, double* partials, int num_chunks,
                     void* stream) {
  TTB_REQUIRE(out != a && out != b, "add_bn_stats: the output must not alias an input");
  return launch_col_reduce<2>(a, b, nullptr, nullptr, nullptr, nullptr, m, c, partials, num_chunks, as_stream(stream),
                              "add_bn_stats", out);
}

int ttb_bn_reduce_partials(const double* partials, int num_chunks, int c2, double* sums, void* stream) {
  if (c2 <= 0) return 0;
  launch_k(reduce_partials_kernel, (c2 + 31) / 32, dim3(32, kLanes), 0, as_stream(stream), partials, num_chunks, c2, sums);
  return check_launch("bn_reduce_partials");
}

int ttb_bn_finalize(const double* sums, int num_chunks, int64_t count, int c, float eps, float momentum, const float* gamma,
                    const float* beta, float* running_mean, float* running_var, float* mean, float* var_eps,
                    float* sd, float* scale, float* shift, void* stream) {
  if (c <= 0) return 0;
  TTB_REQUIRE(count > 0, "bn_finalize: count must be positive");
  BnFwdFinalize fin;
  fin.count = (double)count;
  fin.eps = eps;
  fin.momentum = momentum;
  bn_fwd_host_factors(count, momentum, &fin.unbias, &fin.one_minus_momentum);
  fin.gamma = gamma; fin.beta = beta; fin.running_mean = running_mean; fin.running_var = running_var;
  fin.mean = mean; fin.var_eps = var_eps; fin.sd = sd; fin.scale = scale; fin.shift = shift;
  launch_k(bn_finalize_kernel, (c + kFinCh - 1) / kFinCh, dim3(kFinCh, kFinLanes), 0, as_stream(stream), sums, num_chunks < 1 ? 1 : num_chunks, c, fin);
  return check_launch("bn_finalize");
}

int ttb_bn_prepare_eval(const float* mean_in, const float* var_in, int c, float eps, const float* gamma,
                        const float* beta, float* mean, float* var_eps, float* sd, float* scale, float* shift,
                        void* stream) {
  if (c <= 0) return 0;
  launch_k(bn_prepare_eval_kernel, (c + 127) / 128, 128, 0, as_stream(stream), mean_in, var_in, c, eps, gamma, beta, mean,
                                                                         var_eps, sd, scale, shift);
  return check_launch("bn_prepare_eval");
}

int ttb_bn_fold_eval(const float* mean, const float* var, int c, float eps, const float* gamma, const float* beta,
                     const float* conv_bias, float* scale, float* bias, void* stream) {
  if (c <= 0) return 0;
  launch_k(bn_fold_eval_kernel, (c + 127) / 128, 128, 0, as_stream(stream), mean, var, c, eps, gamma, beta, conv_bias, scale, bias);
  return check_launch("bn_fold_eval");
}

int ttb_bn_apply(const float* x, float* y, int64_t m, int c, const float* mean, const float* scale, const float* beta,
                 int relu, void* y_bf16, void* stream) {
  int64_t n = m * c;
  if (n <= 0) return 0;
  cudaStream_t st = as_stream(stream);
  __nv_bfloat16* yh = reinterpret_cast<__nv_bfloat16*>(y_bf16);
  if (c % 4 == 0 && a16(x) && a16(y) && a16(mean) && a16(scale) && a16(beta) && (reinterpret_cast<uintptr_t>(yh) & 7) == 0) {
    int grid = elementwise_grid(n / 4, 256);
    if (relu) {
      if (yh) launch_k(bn_apply_kernel<true, true>, grid, 256, 0, st, x, y, yh, n / 4, c / 4, mean, scale, beta);
      else launch_k(bn_apply_kernel<true, false>, grid, 256, 0, st, x, y, yh, n / 4, c / 4, mean, scale, beta);
    } else {
      if (yh) launch_k(bn_apply_kernel<false, true>, grid, 256, 0, st, x, y, yh, n / 4, c / 4, mean, scale, beta);
      else launch_k(bn_apply_kernel<false, false>, grid, 256, 0, st, x, y, yh, n / 4, c / 4, mean, scale, beta);
    }
  } else {
    int grid = elementwise_grid(n, 256);
    if (relu) launch_k(bn_apply_scalar_kernel<true>, grid, 256, 0, st, x, y, yh, n, c, mean, scale, beta);
    else launch_k(bn_apply_scalar_kernel<false>, grid, 256, 0, st, x, y, yh, n, c, mean, scale, beta);
  }
  return check_launch("bn_apply");
}

int ttb_bn_apply_add(const float* x, const float* residual, float* y, int64_t m, int c, const float* mean, const float* scale,
                     const float* beta, int relu, void* y_bf16, void* stream) {
  int64_t n = m * c;
  if (n <= 0) return 0;
  __nv_bfloat16* yh = reinterpret_cast<__nv_bfloat16*>(y_bf16);
  TTB_REQUIRE(c % 4 == 0 && a16(x) && a16(residual) && a16(y) && a16(mean) && a16(scale) && a16(beta) &&
                  (reinterpret_cast<uintptr_t>(yh) & 7) == 0,
              "bn_apply_add: needs a multiple of 4 channels and 16-byte aligned buffers");
  cudaStream_t st = as_stream(stream);
  const int grid = elementwise_grid(n / 4, 256);
  if (relu) {
    if (yh) launch_k(bn_apply_add_kernel<true, true>, grid, 256, 0, st, x, residual, y, yh, n / 4, c / 4, mean, scale, beta);
    else launch_k(bn_apply_add_kernel<true, false>, grid, 256, 0, st, x, residual, y, yh, n / 4, c / 4, mean, scale, beta);
  } else {
    if (yh) launch_k(bn_apply_add_kernel<false, true>, grid, 256, 0, st, x, residual, y, yh, n / 4, c / 4, mean, scale, beta);
    else launch_k(bn_apply_add_kernel<false, false>, grid, 256, 0, st, x, residual, y, yh, n / 4, c / 4, mean, scale, beta);
  }
  return check_launch("bn_apply_add");
}

int ttb_bn_bwd_reduce(const float* dy, const float* x, const float* mean, const float* relu_out, const float* relu_scale,
                      const float* relu_shift, int64_t m, int c, double* partials, int num_chunks, void* stream) {
  TTB_REQUIRE(!(relu_out && relu_scale) && (!relu_scale == !relu_shift), "bn_bwd_reduce: give relu_out OR relu_scale+relu_shift");
  return launch_col_reduce<1>(dy, x, mean, relu_out, relu_scale, relu_shift, m, c, partials, num_chunks, as_stream(stream),
                              "bn_bwd_reduce");
}

int ttb_bn_bwd_finalize(const double* sums, int num_chunks, int64_t count, int c, const float* gamma, const float* var_eps,
                        const float* sd, float* dgamma, float* dbeta, float* coef, void* stream) {
  if (c <= 0) return 0;
  BnBwdFinalize fin;
  fin.count = (double)count;
  fin.c = c;
  fin.gamma = gamma; fin.var_eps = var_eps; fin.sd = sd; fin.dgamma = dgamma; fin.dbeta = dbeta; fin.coef = coef;
  launch_k(bn_bwd_finalize_kernel, (c + kFinCh - 1) / kFinCh, dim3(kFinCh, kFinLanes), 0, as_stream(stream), sums, num_chunks < 1 ? 1 : num_chunks, c, fin);
  return check_launch("bn_bwd_finalize");
}

int ttb_bn_bwd_apply(const float* dy, const float* x, const float* mean, const float* relu_out, const float* relu_scale,
                     const float* relu_shift, const float* coef, const float* accum, float* dx, int64_t m, int c,
                     void* dx_bf16, void* stream) {
  int64_t n = m * c;
  if (n <= 0) return 0;
  TTB_REQUIRE(!(relu_out && relu_scale) && (!relu_scale == !relu_shift), "bn_bwd_apply: give relu_out OR relu_scale+relu_shift");
  cudaStream_t st = as_stream(stream);
  bool vec = c % 4 == 0 && a16(dy) && a16(x) && a16(dx) && a16(mean) && a16(coef) && (!relu_out || a16(relu_out)) &&
             (!accum || a16(accum)) && (!relu_scale || (a16(relu_scale) && a16(relu_shift)));
  const int mode = relu_out ? 1 : (relu_scale ? 2 : 0);
  __nv_bfloat16* dxh = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
  if (vec && (reinterpret_cast<uintptr_t>(dxh) & 7) == 0) {
    int grid = elementwise_grid(n / 4, 256);
#define TTB_BWD_APPLY(MASK)                                                                                             \
  do {                                                                                                                  \
    if (dxh)                                                                                                            \
      launch_k(bn_bwd_apply_kernel<MASK, true>, grid, 256, 0, st, dy, x, mean, relu_out, relu_scale, relu_shift, coef, accum, dx, \
                                                            dxh, n / 4, c / 4, c);                                      \
    else                                                                                                                \
      launch_k(bn_bwd_apply_kernel<MASK, false>, grid, 256, 0, st, dy, x, mean, relu_out, relu_scale, relu_shift, coef, accum, \
                                                             dx, dxh, n / 4, c / 4, c);                                 \
  } while (0)
    if (mode == 1) TTB_BWD_APPLY(1);
    else if (mode == 2) TTB_BWD_APPLY(2);
    else TTB_BWD_APPLY(0);
#undef TTB_BWD_APPLY
  } else {
    int grid = elementwise_grid(n, 256);
    if (mode == 1) launch_k(bn_bwd_apply_scalar_kernel<1>, grid, 256, 0, st, dy, x, mean, relu_out, relu_scale, relu_shift, coef, accum, dx, dxh, n, c);
    else if (mode == 2) launch_k(bn_bwd_apply_scalar_kernel<2>, grid, 256, 0, st, dy, x, mean, relu_out, relu_scale, relu_shift, coef, accum, dx, dxh, n, c);
    else launch_k(bn_bwd_apply_scalar_kernel<0>, grid, 256, 0, st, dy, x, mean, relu_out, relu_scale, relu_shift, coef, accum, dx, dxh, n, c);
  }
  return check_launch("bn_bwd_apply");
}

}  // extern "C"
