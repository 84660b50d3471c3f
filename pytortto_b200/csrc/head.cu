// The ops between the conv stacks and the loss, as plain CUDA kernels instead of generic
// array-library expressions: global average pool, Linear (a small fp32 GEMM), LogSoftmax, NLL loss, BCE-with-logits,
// channel concatenation / split of NHWC tensors (UNet skip connections) and the per-channel bias add.
//
// Reference: Mean  autograd/grad_fcn.py:1058-1093;  Linear nn/functional.py:54-63 (Transpose + Mm + Add);
//            LogSoftmax autograd/grad_nn.py:373-392;  NllLoss :287-349;  BinaryCrossEntropyWithLogits :236-285;
//            Cat autograd/grad_fcn.py:881-904.
// All reductions are fixed-order (deterministic, no atomics).  None of these is a roofline item: together they are
// < 1 % of a ResNet step; the point is that no library kernel is left on the training path.
#include <math.h>

#include "common.cuh"

namespace ttb {

// ---------------------------------------------------------------------------------------------------------
// global average pool over H, W of an NHWC tensor: y[n][c] = mean_{h,w} x[n][h][w][c]
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mean_hw_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int hw,
                                                          int c) {
  pdl_entry();
  // one thread per (n, c); consecutive threads = consecutive channels -> coalesced rows
  const int64_t total = (int64_t)n * c;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(t % c);
    const int64_t img = t / c;
    const float* p = x + img * hw * c + ch;
    float acc = 0.f;
    for (int i = 0; i < hw; ++i) acc += p[(int64_t)i * c];
    y[t] = acc / (float)hw;
  }
}

// dx[n][h][w][c] = dy[n][c] / (H*W)
__global__ void __launch_bounds__(256) mean_hw_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int64_t total,
                                                          int hw, int c) {
  pdl_entry();
  const float inv = 1.f / (float)hw;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(t % c);
    const int64_t img = t / ((int64_t)hw * c);
    dx[t] = dy[img * c + ch] * inv;
  }
}

// ---------------------------------------------------------------------------------------------------------
// small fp32 GEMM with arbitrary operand strides (covers x @ W^T, g @ W, g^T @ x of a Linear layer):
//   C[m][n] = sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] (+ bias[n])
// 32 x 32 tiles through shared memory, one output per thread, k summed in ascending order (deterministic).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) matmul_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                      const float* __restrict__ bias, float* __restrict__ c, int m, int n, int k,
                                                      int64_t sam, int64_t sak, int64_t sbk, int64_t sbn) {
  pdl_entry();
  __shared__ float ta[32][33], tb[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int row = blockIdx.y * 32 + ty, col = blockIdx.x * 32 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < k; k0 += 32) {
    const int ka = k0 + tx, kb = k0 + ty;
    ta[ty][tx] = (row < m && ka < k) ? a[(int64_t)row * sam + (int64_t)ka * sak] : 0.f;
    tb[ty][tx] = (kb < k && col < n) ? b[(int64_t)kb * sbk + (int64_t)col * sbn] : 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc = fmaf(ta[ty][j], tb[j][tx], acc);
    __syncthreads();
  }
  if (row < m && col < n) c[(int64_t)row * n + col] = acc + (bias ? bias[col] : 0.f);
}

// The same product for FEW outputs and a LONG reduction (the classifier head: 256 x 10 logits over 512 features, the
// 10 x 512 weight gradient over 256 rows): one WARP per output, lanes stride over k, fixed-order shuffle tree - the tiled
// kernel above would run 8 - 16 blocks with a serial k loop.
__global__ void __launch_bounds__(256) matmul_splitk_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            const float* __restrict__ bias, float* __restrict__ c, int m, int n,
                                                            int k, int64_t sam, int64_t sak, int64_t sbk, int64_t sbn) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t o = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= (int64_t)m * n) return;
  const int row = (int)(o / n), col = (int)(o % n);
  const float* pa = a + (int64_t)row * sam;
  const float* pb = b + (int64_t)col * sbn;
  float acc0 = 0.f, acc1 = 0.f;
  int kk = lane;
  for (; kk + 32 < k; kk += 64) {
    acc0 = fmaf(pa[(int64_t)kk * sak], pb[(int64_t)kk * sbk], acc0);
    acc1 = fmaf(pa[(int64_t)(kk + 32) * sak], pb[(int64_t)(kk + 32) * sbk], acc1);
  }
  if (kk < k) acc0 = fmaf(pa[(int64_t)kk * sak], pb[(int64_t)kk * sbk], acc0);
  float acc = acc0 + acc1;
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) c[o] = acc + (bias ? bias[col] : 0.f);
}

// ---------------------------------------------------------------------------------------------------------
// LogSoftmax over the last axis of a [rows][cols] matrix: one warp per row
//   aug = x - max(x);  y = aug - log(sum(exp(aug)))            (grad_nn.py:379-383)
//   dx = dy - sum(dy) * exp(y)                                  (:389-391)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) log_softmax_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int cols) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* px = x + (int64_t)r * cols;
  float mx = -INFINITY;
  for (int j = lane; j < cols; j += 32) mx = fmaxf(mx, px[j]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < cols; j += 32) s += expf(px[j] - mx);
  s = warp_sum(s);
  const float ls = logf(s);
  for (int j = lane; j < cols; j += 32) y[(int64_t)r * cols + j] = (px[j] - mx) - ls;
}

__global__ void __launch_bounds__(256) log_softmax_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                              float* __restrict__ dx, int rows, int cols) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* pg = dy + (int64_t)r * cols;
  const float* py = y + (int64_t)r * cols;
  float s = 0.f;
  for (int j = lane; j < cols; j += 32) s += pg[j];
  s = warp_sum(s);
  for (int j = lane; j < cols; j += 32) dx[(int64_t)r * cols + j] = pg[j] - s * expf(py[j]);
}

// ---------------------------------------------------------------------------------------------------------
// NLL loss on log-probabilities [rows][cols] with int64 class targets (grad_nn.py:287-349)
//   reduction 0 none: out[r] = -w_r * logp[r][t_r];  1 mean: sum / count(w);  2 sum.   w_r = (t_r != ignore_index)
//   `count` (device float) receives the number of contributing rows (the N of the mean) for backward.
// One block; per-thread partial over its rows, fixed-order tree.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nll_fwd_kernel(const float* __restrict__ logp, const long long* __restrict__ tgt, int rows,
                                                      int cols, long long ignore_index, int reduction, float* __restrict__ out,
                                                      float* __restrict__ count) {
  pdl_entry();
  __shared__ float s_sum[256], s_cnt[256];
  float acc = 0.f, cnt = 0.f;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const long long t = tgt[r];
    const bool on = t != ignore_index;
    const long long tc = t < 0 ? 0 : (t >= cols ? cols - 1 : t);
    const float v = on ? -logp[(int64_t)r * cols + tc] : 0.f;
    if (reduction == 0) out[r] = v;
    acc += v;
    cnt += on ? 1.f : 0.f;
  }
  s_sum[threadIdx.x] = acc;
  s_cnt[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
      s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (count) *count = s_cnt[0];
    if (reduction == 1) *out = s_sum[0] / s_cnt[0];
    else if (reduction == 2) *out = s_sum[0];
  }
}

// dx[r][c] = (c == t_r && t_r != ignore) ? -g_r : 0, g_r = g[0] (/ count for mean) or g[r] (reduction none)
__global__ void __launch_bounds__(256) nll_bwd_kernel(const float* __restrict__ g, const long long* __restrict__ tgt, int rows,
                                                      int cols, long long ignore_index, int reduction,
                                                      const float* __restrict__ count, float* __restrict__ dx) {
  pdl_entry();
  const int64_t total = (int64_t)rows * cols;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(t / cols), c = (int)(t % cols);
    const long long tr = tgt[r];
    float v = 0.f;
    if (tr != ignore_index && (long long)c == tr) {
      float gr = reduction == 0 ? g[r] : g[0];
      if (reduction == 1) gr = gr / *count;
      v = -gr;
    }
    dx[t] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// BCE with logits (grad_nn.py:236-285): l = max(x,0) - x*t + log1p(exp(-|x|));  dx = (sigmoid(x) - t) * g [/ n]
// two-stage fixed-order reduction (double partials)
// ---------------------------------------------------------------------------------------------------------
constexpr int kBceBlocks = 592;  // 4 per SM

__global__ void __launch_bounds__(256) bce_fwd_kernel(const float* __restrict__ x, const float* __restrict__ t, int64_t n,
                                                      float* __restrict__ elem_out, double* __restrict__ partials) {
  pdl_entry();
  __shared__ double sm[256];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xv = x[i], tv = t[i];
    const float l = fmaxf(xv, 0.f) - xv * tv + log1pf(expf(-fabsf(xv)));
    if (elem_out) elem_out[i] = l;
    acc += (double)l;
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0 && partials) partials[blockIdx.x] = sm[0];
}

__global__ void __launch_bounds__(256) sum_partials_kernel(const double* __restrict__ partials, int count, double scale,
                                                           float* __restrict__ out) {
  pdl_entry();
  __shared__ double sm[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < count; i += blockDim.x) acc += partials[i];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)(sm[0] * scale);
}

__global__ void __launch_bounds__(256) bce_bwd_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                      const float* __restrict__ g, int g_per_elem, float scale, int64_t n,
                                                      float* __restrict__ dx) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xv = x[i];
    const float sg = 1.f / (1.f + expf(-xv));
    dx[i] = (sg - t[i]) * (g_per_elem ? g[i] : g[0]) * scale;
  }
}

// ---------------------------------------------------------------------------------------------------------
// channel-range copy between NHWC tensors viewed as [rows][channels]:
//   dst[r][dst_off + c] = src[r][src_off + c],  c < c_copy      (Cat: one call per input; Split / Cat backward: per output)
// ---------------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256) copy_channels_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t rows,
                                                            int c_src, int c_dst, int src_off, int dst_off, int c_copy) {
  pdl_entry();
  const int per = VEC ? c_copy / 4 : c_copy;
  const int64_t total = rows * per;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / per;
    const int j = (int)(t % per);
    if (VEC) {
      st_f4(dst + r * c_dst + dst_off + 4 * j, ld_f4_stream(src + r * c_src + src_off + 4 * j));
    } else {
      dst[r * c_dst + dst_off + j] = src[r * c_src + src_off + j];
    }
  }
}

// y[r][c] += bias[c]
__global__ void __launch_bounds__(256) add_bias_kernel(float* __restrict__ y, const float* __restrict__ bias, int64_t total, int c) {
  pdl_entry();
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
    y[t] += bias[(int)(t % c)];
}

}  // namespace ttb

using namespace ttb;

extern "C" {

int ttb_mean_hw_fwd(const float* x, float* y, int n, int hw, int c, void* stream) {
  if ((int64_t)n * c <= 0 || hw <= 0) return 0;
  launch_k(mean_hw_fwd_kernel, elementwise_grid((int64_t)n * c, 256), 256, 0, as_stream(stream), x, y, n, hw, c);
  return check_launch("mean_hw_fwd");
}

int ttb_mean_hw_bwd(const float* dy, float* dx, int n, int hw, int c, void* stream) {
  const int64_t total = (int64_t)n * hw * c;
  if (total <= 0) return 0;
  launch_k(mean_hw_bwd_kernel, elementwise_grid(total, 256), 256, 0, as_stream(stream), dy, dx, total, hw, c);
  return check_launch("mean_hw_bwd");
}

int ttb_matmul(const float* a, const float* b, const float* bias, float* c, int m, int n, int k, int64_t sam, int64_t sak,
               int64_t sbk, int64_t sbn, void* stream) {
  if (m <= 0 || n <= 0) return 0;
  TTB_REQUIRE(k >= 0 && a && b && c, "matmul: bad arguments");
  if (k >= 128 && (int64_t)m * n <= 16384) {  // few outputs, long reduction: one warp per output
    launch_k(matmul_splitk_kernel, (unsigned)(((int64_t)m * n + 7) / 8), 256, 0, as_stream(stream), a, b, bias, c, m, n, k, sam,
             sak, sbk, sbn);
    return check_launch("matmul");
  }
  launch_k(matmul_kernel, dim3((n + 31) / 32, (m + 31) / 32), dim3(32, 32), 0, as_stream(stream), a, b, bias, c, m, n, k, sam,
           sak, sbk, sbn);
  return check_launch("matmul");
}

int ttb_log_softmax_fwd(const float* x, float* y, int rows, int cols, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  launch_k(log_softmax_fwd_kernel, (rows + 7) / 8, 256, 0, as_stream(stream), x, y, rows, cols);
  return check_launch("log_softmax_fwd");
}

int ttb_log_softmax_bwd(const float* dy, const float* y, float* dx, int rows, int cols, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  launch_k(log_softmax_bwd_kernel, (rows + 7) / 8, 256, 0, as_stream(stream), dy, y, dx, rows, cols);
  return check_launch("log_softmax_bwd");
}

int ttb_nll_loss_fwd(const float* logp, const int64_t* target, int rows, int cols, int64_t ignore_index, int reduction,
                     float* out, float* count, void* stream) {
  TTB_REQUIRE(rows > 0 && cols > 0 && reduction >= 0 && reduction <= 2, "nll_loss_fwd: bad arguments");
  launch_k(nll_fwd_kernel, 1, 256, 0, as_stream(stream), logp, reinterpret_cast<const long long*>(target), rows, cols,
           (long long)ignore_index, reduction, out, count);
  return check_launch("nll_loss_fwd");
}

int ttb_nll_loss_bwd(const float* g, const int64_t* target, int rows, int cols, int64_t ignore_index, int reduction,
                     const float* count, float* dx, void* stream) {
  TTB_REQUIRE(rows > 0 && cols > 0 && reduction >= 0 && reduction <= 2, "nll_loss_bwd: bad arguments");
  TTB_REQUIRE(reduction != 1 || count != nullptr, "nll_loss_bwd: mean reduction needs the count from forward");
  launch_k(nll_bwd_kernel, elementwise_grid((int64_t)rows * cols, 256), 256, 0, as_stream(stream), g,
           reinterpret_cast<const long long*>(target), rows, cols, (long long)ignore_index, reduction, count, dx);
  return check_launch("nll_loss_bwd");
}

size_t ttb_bce_logits_workspace_size(void) { return kBceBlocks * sizeof(double); }

int ttb_bce_logits_fwd(const float* x, const float* t, int64_t n, int reduction, float* out, void* workspace, void* stream) {
  TTB_REQUIRE(n > 0 && reduction >= 0 && reduction <= 2, "bce_logits_fwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (reduction == 0) {
    launch_k(bce_fwd_kernel, elementwise_grid(n, 256), 256, 0, st, x, t, n, out, (double*)nullptr);
    return check_launch("bce_logits_fwd");
  }
  TTB_REQUIRE(workspace != nullptr, "bce_logits_fwd: workspace needed for a reduced loss");
  int blocks = (int)(ceil_div(n, 256) < kBceBlocks ? ceil_div(n, 256) : kBceBlocks);
  launch_k(bce_fwd_kernel, blocks, 256, 0, st, x, t, n, (float*)nullptr, reinterpret_cast<double*>(workspace));
  if (check_launch("bce_logits_fwd")) return 1;
  launch_k(sum_partials_kernel, 1, 256, 0, st, reinterpret_cast<const double*>(workspace), blocks,
           reduction == 1 ? 1.0 / (double)n : 1.0, out);
  return check_launch("bce_logits_fwd(sum)");
}

int ttb_bce_logits_bwd(const float* x, const float* t, const float* g, int g_per_elem, float scale, int64_t n, float* dx,
                       void* stream) {
  if (n <= 0) return 0;
  launch_k(bce_bwd_kernel, elementwise_grid(n, 256), 256, 0, as_stream(stream), x, t, g, g_per_elem, scale, n, dx);
  return check_launch("bce_logits_bwd");
}

int ttb_copy_channels(const float* src, float* dst, int64_t rows, int c_src, int c_dst, int src_off, int dst_off, int c_copy,
                      void* stream) {
  if (rows <= 0 || c_copy <= 0) return 0;
  TTB_REQUIRE(src_off >= 0 && dst_off >= 0 && src_off + c_copy <= c_src && dst_off + c_copy <= c_dst,
              "copy_channels: channel range out of bounds");
  cudaStream_t st = as_stream(stream);
  const bool vec = c_copy % 4 == 0 && c_src % 4 == 0 && c_dst % 4 == 0 && src_off % 4 == 0 && dst_off % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
  if (vec)
    launch_k(copy_channels_kernel<true>, elementwise_grid(rows * (c_copy / 4), 256), 256, 0, st, src, dst, rows, c_src, c_dst,
             src_off, dst_off, c_copy);
  else
    launch_k(copy_channels_kernel<false>, elementwise_grid(rows * c_copy, 256), 256, 0, st, src, dst, rows, c_src, c_dst,
             src_off, dst_off, c_copy);
  return check_launch("copy_channels");
}

int ttb_add_bias(float* y, const float* bias, int64_t rows, int c, void* stream) {
  if (rows <= 0 || c <= 0) return 0;
  launch_k(add_bias_kernel, elementwise_grid(rows * c, 256), 256, 0, as_stream(stream), y, bias, rows * c, c);
  return check_launch("add_bias");
}

}  // extern "C"
