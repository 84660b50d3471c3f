// HBM-bound elementwise kernels: ReLU fwd/bwd, add, axpy, scale, fill, SGD, NCHW<->NHWC.
// All are grid-stride, 128-bit vectorised with a scalar tail; grids are sized in multiples of the SM count.
#include "common.cuh"

namespace ttb {

constexpr int kThreads = 256;

// Generic vectorised elementwise driver: `Op` gets float4s (vector body) and floats (tail / unaligned).
template <class F4, class F1>
__global__ void __launch_bounds__(kThreads) ew_kernel(int64_t n, F4 f4, F1 f1, bool vec_ok) {
  pdl_entry();
  int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok) {
    int64_t n4 = n >> 2;
    // 2x unrolled grid-stride loop: two independent 128-bit accesses in flight per thread
    int64_t i = tid;
    for (; i + stride < n4; i += 2 * stride) {
      f4(i);
      f4(i + stride);
    }
    for (; i < n4; i += stride) f4(i);
    for (int64_t j = (n4 << 2) + tid; j < n; j += stride) f1(j);
  } else {
    for (int64_t j = tid; j < n; j += stride) f1(j);
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <class F4, class F1>
static int launch_ew(const char* what, int64_t n, bool vec_ok, F4 f4, F1 f1, cudaStream_t st) {
  if (n <= 0) return 0;
  int64_t items = vec_ok ? (n + 3) / 4 : n;
  int grid = elementwise_grid(items, kThreads);
  launch_k(ew_kernel<F4, F1>, grid, kThreads, 0, st, n, f4, f1, vec_ok);
  return check_launch(what);
}

__device__ __forceinline__ float relu1(float v) { return v < 0.f ? 0.f : v; }  // NaN propagates like np.maximum

// ---------------------------------------------------------------------------------------------------------
// NCHW <-> NHWC (per image: [C][HW] <-> [HW][C]) through a 32x33 shared tile, coalesced on both sides.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                        int rows, int cols, int64_t batch_stride) {
  pdl_entry();
  // src: [batch][rows][cols] -> dst: [batch][cols][rows]
  __shared__ float tile[32][33];
  const float* s = src + (int64_t)blockIdx.z * batch_stride;
  float* d = dst + (int64_t)blockIdx.z * batch_stride;
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    int r = r0 + ty + j, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + j][tx] = s[(int64_t)r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    int c = c0 + ty + j, r = r0 + tx;
    if (r < rows && c < cols) d[(int64_t)c * rows + r] = tile[tx][ty + j];
  }
}

static int transpose_batched(const float* src, float* dst, int batch, int rows, int cols, cudaStream_t st) {
  if (batch <= 0 || rows <= 0 || cols <= 0) return 0;
  int done = 0;
  while (done < batch) {  // gridDim.z <= 65535
    int nb = batch - done < 65535 ? batch - done : 65535;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, nb);
    int64_t off = (int64_t)done * rows * cols;
    launch_k(transpose_kernel, grid, 256, 0, st, src + off, dst + off, rows, cols, (int64_t)rows * cols);
    if (check_launch("transpose")) return 1;
    done += nb;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// multi-tensor SGD: one launch updates up to kMultiMax parameter tensors (pointer table passed by value)
// ---------------------------------------------------------------------------------------------------------
constexpr int kMultiMax = 48;
constexpr int kMultiChunk = 2048;  // elements per CTA (256 threads x 2 float4)

struct SgdMulti {
  float* p[kMultiMax];
  const float* g[kMultiMax];
  float* b[kMultiMax];
  int64_t n[kMultiMax];
  int block_start[kMultiMax + 1];
  unsigned char first[kMultiMax];
  int count;
  float lr, momentum, dampening, weight_decay;
  int nesterov;
};

__device__ __forceinline__ void sgd_elem(float& p, float g, float& b, bool has_buf, bool first, float lr, float mom,
                                         float damp, float wd, int nesterov) {
  float d = g;
  if (wd != 0.f) d = __fadd_rn(d, __fmul_rn(p, wd));
  if (mom != 0.f) {
    b = first ? d : __fadd_rn(__fmul_rn(b, mom), __fmul_rn(d, 1.f - damp));
    d = nesterov ? __fadd_rn(d, __fmul_rn(b, mom)) : b;
  }
  p = __fadd_rn(p, __fmul_rn(d, -lr));
  (void)has_buf;
}

__global__ void __launch_bounds__(256) sgd_multi_kernel(const __grid_constant__ SgdMulti A) {
  pdl_entry();
  // locate the tensor this CTA works on
  int t = 0;
  while (t + 1 < A.count && (int)blockIdx.x >= A.block_start[t + 1]) ++t;
  const int64_t base = (int64_t)(blockIdx.x - A.block_start[t]) * kMultiChunk;
  const int64_t n = A.n[t];
  float* __restrict__ p = A.p[t];
  const float* __restrict__ g = A.g[t];
  float* __restrict__ b = A.b[t];
  const bool first = A.first[t] != 0;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
  if (vec) {
#pragma unroll
    for (int it = 0; it < kMultiChunk / (256 * 4); ++it) {
      int64_t i = base + (int64_t)(it * 256 + threadIdx.x) * 4;
      if (i + 3 < n) {
        float4 pv = ld_f4(p + i), gv = ld_f4(g + i);
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b && !first) bv = ld_f4(b + i);
        sgd_elem(pv.x, gv.x, bv.x, b != nullptr, first, A.lr, A.momentum, A.dampening, A.weight_decay, A.nesterov);
        sgd_elem(pv.y, gv.y, bv.y, b != nullptr, first, A.lr, A.momentum, A.dampening, A.weight_decay, A.nesterov);
        sgd_elem(pv.z, gv.z, bv.z, b != nullptr, first, A.lr, A.momentum, A.dampening, A.weight_decay, A.nesterov);
        sgd_elem(pv.w, gv.w, bv.w, b != nullptr, first, A.lr, A.momentum, A.dampening, A.weight_decay, A.nesterov);
        if (b) st_f4(b + i, bv);
        st_f4(p + i, pv);
      } else {
        for (int64_t j = i; j < n && j < i + 4; ++j) {
          float pv = p[j], bv = (b && !first) ? b[j] : 0.f;
          sgd_elem(pv, g[j], bv, b != nullptr, first, A.lr, A.momentum, A.dampening, A.weight_decay, A.nesterov);
          if (b) b[j] = bv;
          p[j] = pv;
        }
      }
    }
  } else {
    for (int64_t j = base + threadIdx.x; j < n && j < base + kMultiChunk; j += 256) {
      float pv = p[j], bv = (b && !first) ? b[j] : 0.f;
      sgd_elem(pv, g[j], bv, b != nullptr, first, A.lr, A.momentum, A.dampening, A.weight_decay, A.nesterov);
      if (b) b[j] = bv;
      p[j] = pv;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// multi-tensor Adam / AdamW (optim/_functional.py:25-68 and :71-115), same chunking as the SGD kernel.  The step count
// lives on the DEVICE (state[0] = step as float, state[1] = 1 - beta1^step, state[2] = 1 - beta2^step, advanced by
// adam_advance_kernel once per optimizer step) so a captured CUDA graph replays with the right bias corrections.
// ---------------------------------------------------------------------------------------------------------
struct AdamMulti {
  float* p[kMultiMax];
  const float* g[kMultiMax];
  float* m[kMultiMax];
  float* v[kMultiMax];
  float* vmax[kMultiMax];
  int64_t n[kMultiMax];
  int block_start[kMultiMax + 1];
  int count;
  float lr, beta1, beta2, eps, weight_decay;
  int decoupled;  // 0: Adam (L2 term added to the gradient), 1: AdamW (p *= 1 - lr*wd first)
  const float* state;
};

__global__ void adam_advance_kernel(float* state, float beta1, float beta2) {
  pdl_entry();
  const double step = (double)state[0] + 1.0;
  state[0] = (float)step;
  state[1] = (float)(1.0 - pow((double)beta1, step));
  state[2] = (float)(1.0 - pow((double)beta2, step));
}

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float* vmax, const AdamMulti& A, float bc1,
                                          float sqrt_bc2) {
  if (A.decoupled) {
    if (A.weight_decay != 0.f) p = __fmul_rn(p, 1.f - A.lr * A.weight_decay);
  } else if (A.weight_decay != 0.f) {
    g = __fadd_rn(g, __fmul_rn(p, A.weight_decay));
  }
  m = __fadd_rn(__fmul_rn(m, A.beta1), __fmul_rn(1.f - A.beta1, g));
  v = __fadd_rn(__fmul_rn(v, A.beta2), __fmul_rn(__fmul_rn(g, g), 1.f - A.beta2));
  float vv = v;
  if (vmax) {
    vv = fmaxf(*vmax, v);
    *vmax = vv;
  }
  const float denom = __fadd_rn(__fdiv_rn(sqrtf(vv), sqrt_bc2), A.eps);
  p = __fadd_rn(p, __fdiv_rn(__fmul_rn(-(A.lr / bc1), m), denom));
}

__global__ void __launch_bounds__(256) adam_multi_kernel(const __grid_constant__ AdamMulti A) {
  pdl_entry();
  int t = 0;
  while (t + 1 < A.count && (int)blockIdx.x >= A.block_start[t + 1]) ++t;
  const int64_t base = (int64_t)(blockIdx.x - A.block_start[t]) * kMultiChunk;
  const int64_t n = A.n[t];
  float* __restrict__ p = A.p[t];
  const float* __restrict__ g = A.g[t];
  float* __restrict__ m = A.m[t];
  float* __restrict__ v = A.v[t];
  float* __restrict__ vm = A.vmax[t];
  const float bc1 = A.state[1], sqrt_bc2 = sqrtf(A.state[2]);
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(vm)) & 15) == 0;
  if (vec) {
#pragma unroll
    for (int it = 0; it < kMultiChunk / (256 * 4); ++it) {
      int64_t i = base + (int64_t)(it * 256 + threadIdx.x) * 4;
      if (i + 3 < n) {
        float4 pv = ld_f4(p + i), gv = ld_f4(g + i), mv = ld_f4(m + i), vv = ld_f4(v + i);
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vm) xv = ld_f4(vm + i);
        adam_elem(pv.x, gv.x, mv.x, vv.x, vm ? &xv.x : nullptr, A, bc1, sqrt_bc2);
        adam_elem(pv.y, gv.y, mv.y, vv.y, vm ? &xv.y : nullptr, A, bc1, sqrt_bc2);
        adam_elem(pv.z, gv.z, mv.z, vv.z, vm ? &xv.z : nullptr, A, bc1, sqrt_bc2);
        adam_elem(pv.w, gv.w, mv.w, vv.w, vm ? &xv.w : nullptr, A, bc1, sqrt_bc2);
        st_f4(m + i, mv);
        st_f4(v + i, vv);
        if (vm) st_f4(vm + i, xv);
        st_f4(p + i, pv);
        continue;
      }
      for (int64_t j = i; j < n && j < i + 4; ++j) {
        float pv = p[j], mv = m[j], vv = v[j], xv = vm ? vm[j] : 0.f;
        adam_elem(pv, g[j], mv, vv, vm ? &xv : nullptr, A, bc1, sqrt_bc2);
        m[j] = mv; v[j] = vv; p[j] = pv;
        if (vm) vm[j] = xv;
      }
    }
  } else {
    for (int64_t j = base + threadIdx.x; j < n && j < base + kMultiChunk; j += 256) {
      float pv = p[j], mv = m[j], vv = v[j], xv = vm ? vm[j] : 0.f;
      adam_elem(pv, g[j], mv, vv, vm ? &xv : nullptr, A, bc1, sqrt_bc2);
      m[j] = mv; v[j] = vv; p[j] = pv;
      if (vm) vm[j] = xv;
    }
  }
}

}  // namespace ttb

using namespace ttb;

extern "C" {

int ttb_adam_advance(float* state, float beta1, float beta2, void* stream) {
  TTB_REQUIRE(state != nullptr, "adam_advance: null state");
  launch_k(adam_advance_kernel, 1, 1, 0, as_stream(stream), state, beta1, beta2);
  return check_launch("adam_advance");
}

int ttb_adam_step_multi(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                        float* const* exp_avg_sq, float* const* max_exp_avg_sq, const int64_t* sizes, const float* state,
                        float lr, float beta1, float beta2, float eps, float weight_decay, int decoupled, void* stream) {
  TTB_REQUIRE(n_tensors >= 0 && params && grads && exp_avg && exp_avg_sq && sizes && state, "adam_step_multi: null table");
  cudaStream_t st = as_stream(stream);
  int done = 0;
  while (done < n_tensors) {
    AdamMulti A;
    A.count = 0;
    A.lr = lr; A.beta1 = beta1; A.beta2 = beta2; A.eps = eps; A.weight_decay = weight_decay; A.decoupled = decoupled;
    A.state = state;
    int blocks = 0;
    while (done < n_tensors && A.count < kMultiMax) {
      if (sizes[done] > 0) {
        int i = A.count++;
        A.p[i] = params[done];
        A.g[i] = grads[done];
        A.m[i] = exp_avg[done];
        A.v[i] = exp_avg_sq[done];
        A.vmax[i] = max_exp_avg_sq ? max_exp_avg_sq[done] : nullptr;
        A.n[i] = sizes[done];
        A.block_start[i] = blocks;
        blocks += (int)ceil_div(sizes[done], kMultiChunk);
      }
      ++done;
    }
    if (A.count == 0) break;
    A.block_start[A.count] = blocks;
    launch_k(adam_multi_kernel, blocks, 256, 0, st, A);
    if (check_launch("adam_step_multi")) return 1;
  }
  return 0;
}

int ttb_sgd_step_multi(int n_tensors, float* const* params, const float* const* grads, float* const* bufs,
                       const int64_t* sizes, const unsigned char* first_step, float lr, float momentum, float dampening,
                       float weight_decay, int nesterov, void* stream) {
  TTB_REQUIRE(n_tensors >= 0 && params && grads && sizes, "sgd_step_multi: null table");
  cudaStream_t st = as_stream(stream);
  int done = 0;
  while (done < n_tensors) {
    SgdMulti A;
    A.count = 0;
    A.lr = lr; A.momentum = momentum; A.dampening = dampening; A.weight_decay = weight_decay; A.nesterov = nesterov;
    int blocks = 0;
    while (done < n_tensors && A.count < kMultiMax) {
      if (sizes[done] > 0) {
        TTB_REQUIRE(momentum == 0.f || (bufs && bufs[done]), "sgd_step_multi: momentum != 0 needs momentum buffers");
        int i = A.count++;
        A.p[i] = params[done];
        A.g[i] = grads[done];
        A.b[i] = bufs ? bufs[done] : nullptr;
        A.n[i] = sizes[done];
        A.first[i] = first_step ? first_step[done] : 0;
        A.block_start[i] = blocks;
        blocks += (int)ceil_div(sizes[done], kMultiChunk);
      }
      ++done;
    }
    if (A.count == 0) break;
    A.block_start[A.count] = blocks;
    launch_k(sgd_multi_kernel, blocks, 256, 0, st, A);
    if (check_launch("sgd_step_multi")) return 1;
  }
  return 0;
}

int ttb_nchw_to_nhwc(const float* src, float* dst, int n, int c, int h, int w, void* stream) {
  return transpose_batched(src, dst, n, c, h * w, as_stream(stream));
}

int ttb_nhwc_to_nchw(const float* src, float* dst, int n, int c, int h, int w, void* stream) {
  return transpose_batched(src, dst, n, h * w, c, as_stream(stream));
}

int ttb_relu_fwd(const float* x, float* y, int64_t n, void* y_bf16, void* stream) {
  __nv_bfloat16* yh = reinterpret_cast<__nv_bfloat16*>(y_bf16);  // optional bf16 copy of y for the bf16 tensor path
  bool v = aligned16(x) && aligned16(y) && (reinterpret_cast<uintptr_t>(yh) & 7) == 0;
  return launch_ew(
      "relu_fwd", n, v,
      [=] __device__(int64_t i) {
        float4 a = ld_f4(x + 4 * i);
        a.x = relu1(a.x); a.y = relu1(a.y); a.z = relu1(a.z); a.w = relu1(a.w);
        st_f4(y + 4 * i, a);
        if (yh) st_bf16x4(yh + 4 * i, a);
      },
      [=] __device__(int64_t i) {
        float r = relu1(x[i]);
        y[i] = r;
        if (yh) yh[i] = __float2bfloat16_rn(r);
      },
      as_stream(stream));
}

int ttb_relu_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream) {
  bool v = aligned16(dy) && aligned16(y) && aligned16(dx);
  return launch_ew(
      "relu_bwd", n, v,
      [=] __device__(int64_t i) {
        float4 g = ld_f4(dy + 4 * i), o = ld_f4(y + 4 * i);
        g.x = o.x > 0.f ? g.x : g.x * 0.f;  // dy * (y > 0): keeps NaN/Inf semantics of the multiply
        g.y = o.y > 0.f ? g.y : g.y * 0.f;
        g.z = o.z > 0.f ? g.z : g.z * 0.f;
        g.w = o.w > 0.f ? g.w : g.w * 0.f;
        st_f4(dx + 4 * i, g);
      },
      [=] __device__(int64_t i) { dx[i] = y[i] > 0.f ? dy[i] : dy[i] * 0.f; }, as_stream(stream));
}

int ttb_add(const float* a, const float* b, float* out, int64_t n, void* stream) {
  bool v = aligned16(a) && aligned16(b) && aligned16(out);
  return launch_ew(
      "add", n, v,
      [=] __device__(int64_t i) {
        float4 p = ld_f4(a + 4 * i), q = ld_f4(b + 4 * i);
        p.x += q.x; p.y += q.y; p.z += q.z; p.w += q.w;
        st_f4(out + 4 * i, p);
      },
      [=] __device__(int64_t i) { out[i] = a[i] + b[i]; }, as_stream(stream));
}

int ttb_axpy(float alpha, const float* x, float* y, int64_t n, void* stream) {
  bool v = aligned16(x) && aligned16(y);
  return launch_ew(
      "axpy", n, v,
      [=] __device__(int64_t i) {
        float4 p = ld_f4(x + 4 * i), q = ld_f4(y + 4 * i);
        q.x = fmaf(alpha, p.x, q.x); q.y = fmaf(alpha, p.y, q.y);
        q.z = fmaf(alpha, p.z, q.z); q.w = fmaf(alpha, p.w, q.w);
        st_f4(y + 4 * i, q);
      },
      [=] __device__(int64_t i) { y[i] = fmaf(alpha, x[i], y[i]); }, as_stream(stream));
}

int ttb_scale(float alpha, float* x, int64_t n, void* stream) {
  bool v = aligned16(x);
  return launch_ew(
      "scale", n, v,
      [=] __device__(int64_t i) {
        float4 p = ld_f4(x + 4 * i);
        p.x *= alpha; p.y *= alpha; p.z *= alpha; p.w *= alpha;
        st_f4(x + 4 * i, p);
      },
      [=] __device__(int64_t i) { x[i] *= alpha; }, as_stream(stream));
}

int ttb_fill(float value, float* x, int64_t n, void* stream) {
  bool v = aligned16(x);
  return launch_ew(
      "fill", n, v, [=] __device__(int64_t i) { st_f4(x + 4 * i, make_float4(value, value, value, value)); },
      [=] __device__(int64_t i) { x[i] = value; }, as_stream(stream));
}

// optim/_functional.py:4-22, one fused pass: d_p = g + wd*p; buf = first ? d_p : mom*buf + (1-damp)*d_p;
// d_p = nesterov ? d_p + mom*buf : buf; p += -lr*d_p.  Plain (non-fma-contracted) ops in the reference's order.
__device__ __forceinline__ void sgd1(float& p, float g, float* buf, float lr, float mom, float damp, float wd,
                                     int nesterov, int first) {
  float d = g;
  if (wd != 0.f) d = __fadd_rn(d, __fmul_rn(p, wd));
  if (mom != 0.f) {
    float b;
    if (first) {
      b = d;
    } else {
      b = __fadd_rn(__fmul_rn(*buf, mom), __fmul_rn(d, 1.f - damp));
    }
    *buf = b;
    d = nesterov ? __fadd_rn(d, __fmul_rn(b, mom)) : b;
  }
  p = __fadd_rn(p, __fmul_rn(d, -lr));
}

int ttb_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                 float dampening, float weight_decay, int nesterov, int first_step, void* stream) {
  TTB_REQUIRE(momentum == 0.f || momentum_buf != nullptr, "sgd_step: momentum != 0 needs a momentum buffer");
  bool v = aligned16(param) && aligned16(grad) && (momentum_buf == nullptr || aligned16(momentum_buf));
  float* mb = momentum_buf;
  return launch_ew(
      "sgd_step", n, v,
      [=] __device__(int64_t i) {
        float4 p = ld_f4(param + 4 * i), g = ld_f4(grad + 4 * i);
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mb && !first_step) b = ld_f4(mb + 4 * i);
        sgd1(p.x, g.x, &b.x, lr, momentum, dampening, weight_decay, nesterov, first_step);
        sgd1(p.y, g.y, &b.y, lr, momentum, dampening, weight_decay, nesterov, first_step);
        sgd1(p.z, g.z, &b.z, lr, momentum, dampening, weight_decay, nesterov, first_step);
        sgd1(p.w, g.w, &b.w, lr, momentum, dampening, weight_decay, nesterov, first_step);
        if (mb) st_f4(mb + 4 * i, b);
        st_f4(param + 4 * i, p);
      },
      [=] __device__(int64_t i) {
        float p = param[i];
        float b = (mb && !first_step) ? mb[i] : 0.f;
        sgd1(p, grad[i], &b, lr, momentum, dampening, weight_decay, nesterov, first_step);
        if (mb) mb[i] = b;
        param[i] = p;
      },
      as_stream(stream));
}

}  // extern "C"
