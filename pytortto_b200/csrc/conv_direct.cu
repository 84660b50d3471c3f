// Generic exact-fp32 direct convolution kernels (CUDA-core FFMA).
//
// These cover EVERY configuration of the reference's F.conv2d (any groups / dilation / stride / channel count,
// including the 3-channel network stems and depthwise convs) and are the path TTB_MATH_FP32 selects.  The
// tcgen05 implicit-GEMM kernels (conv_igemm.cu) take over whenever the problem fits tensor-core tiles; these
// stay as the fallback for shapes that do not (e.g. Cin = 3, Cin/groups < 32) and as the bisecting reference
// for the tensor path.  Accumulation is sequential fp32 FMA.
//
// Layouts: x [N,H,W,C], y/dy [N,P,Q,K], w/dw [K][R][S][Cg] with Cg = C/groups, Kg = K/groups.
// Reference semantics: /root/reference/src/tortto/autograd/grad_nn.py:595-682.
#include "common.cuh"

namespace ttb {

// ---------------------------------------------------------------------------------------------------------
// fprop: one thread per output element (m, k), k fastest (coalesced y store; x loads broadcast over k).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
direct_fprop_kernel(ttb_conv_desc d, const float* __restrict__ x, const float* __restrict__ w,
                    const float* __restrict__ bias, float* __restrict__ y, int64_t total) {
  pdl_entry();
  const int cg = d.c / d.groups, kg = d.k / d.groups;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    int k = (int)(t % d.k);
    int64_t m = t / d.k;
    int q = (int)(m % d.q);
    int64_t t2 = m / d.q;
    int p = (int)(t2 % d.p);
    int n = (int)(t2 / d.p);
    int g = k / kg;
    const float* wk = w + (int64_t)k * d.r * d.s * cg;
    float acc = 0.f;
    for (int r = 0; r < d.r; ++r) {
      int h = p * d.stride_h - d.pad_h + r * d.dil_h;
      if (h < 0 || h >= d.h) continue;
      for (int s = 0; s < d.s; ++s) {
        int ww = q * d.stride_w - d.pad_w + s * d.dil_w;
        if (ww < 0 || ww >= d.w) continue;
        const float* px = x + (((int64_t)n * d.h + h) * d.w + ww) * d.c + g * cg;
        const float* pw = wk + (r * d.s + s) * cg;
        for (int c = 0; c < cg; ++c) acc = fmaf(px[c], pw[c], acc);
      }
    }
    if (bias) acc += bias[k];
    y[t] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------
// dgrad (gather form, no zero insertion): one thread per input element (n, h, w, c), c fastest.
// dx[n,h,w,c] = sum_{r,s : (h+ph-r*dh) % sh == 0, ...} sum_{k in group(c)} dy[n,p,q,k] * w[k,r,s,c_local]
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
direct_dgrad_kernel(ttb_conv_desc d, const float* __restrict__ dy, const float* __restrict__ w,
                    float* __restrict__ dx, int64_t total) {
  pdl_entry();
  const int cg = d.c / d.groups, kg = d.k / d.groups;
  const int64_t wstride_k = (int64_t)d.r * d.s * cg;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    int c = (int)(t % d.c);
    int64_t pix = t / d.c;
    int wi = (int)(pix % d.w);
    int64_t t2 = pix / d.w;
    int hi = (int)(t2 % d.h);
    int n = (int)(t2 / d.h);
    int g = c / cg, cl = c % cg;
    float acc = 0.f;
    for (int r = 0; r < d.r; ++r) {
      int hp = hi + d.pad_h - r * d.dil_h;
      if (hp < 0 || hp % d.stride_h) continue;
      int p = hp / d.stride_h;
      if (p >= d.p) continue;
      for (int s = 0; s < d.s; ++s) {
        int wq = wi + d.pad_w - s * d.dil_w;
        if (wq < 0 || wq % d.stride_w) continue;
        int q = wq / d.stride_w;
        if (q >= d.q) continue;
        const float* pdy = dy + (((int64_t)n * d.p + p) * d.q + q) * d.k + g * kg;
        const float* pw = w + (int64_t)g * kg * wstride_k + (r * d.s + s) * cg + cl;
        for (int k = 0; k < kg; ++k) acc = fmaf(pdy[k], pw[k * wstride_k], acc);
      }
    }
    dx[t] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------
// wgrad: thread -> one dw element (k, r, s, c_local); blockIdx.y -> chunk of output pixels; per-chunk partials
// are written to the workspace [chunks][K*R*S*Cg] and summed in fixed order by wgrad_reduce_kernel.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
direct_wgrad_kernel(ttb_conv_desc d, const float* __restrict__ x, const float* __restrict__ dy,
                    float* __restrict__ partial, int64_t wsize, int64_t pixels_per_chunk) {
  pdl_entry();
  const int cg = d.c / d.groups, kg = d.k / d.groups;
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= wsize) return;
  int cl = (int)(e % cg);
  int64_t t = e / cg;
  int s = (int)(t % d.s);
  t /= d.s;
  int r = (int)(t % d.r);
  int k = (int)(t / d.r);
  int g = k / kg;
  int64_t m_total = (int64_t)d.n * d.p * d.q;
  int64_t m0 = (int64_t)blockIdx.y * pixels_per_chunk;
  int64_t m1 = m0 + pixels_per_chunk < m_total ? m0 + pixels_per_chunk : m_total;
  float acc = 0.f;
  for (int64_t m = m0; m < m1; ++m) {
    int q = (int)(m % d.q);
    int64_t t2 = m / d.q;
    int p = (int)(t2 % d.p);
    int n = (int)(t2 / d.p);
    int h = p * d.stride_h - d.pad_h + r * d.dil_h;
    int ww = q * d.stride_w - d.pad_w + s * d.dil_w;
    if (h < 0 || h >= d.h || ww < 0 || ww >= d.w) continue;
    acc = fmaf(dy[m * d.k + k], x[(((int64_t)n * d.h + h) * d.w + ww) * d.c + g * cg + cl], acc);
  }
  partial[(int64_t)blockIdx.y * wsize + e] = acc;
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int chunks, int64_t wsize,
                                    float* __restrict__ dw) {
  pdl_entry();
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= wsize) return;
  float acc = 0.f;
  for (int c = 0; c < chunks; ++c) acc += partial[(int64_t)c * wsize + e];
  dw[e] = acc;
}

static int wgrad_chunks(const ttb_conv_desc* d) {
  int64_t wsize = (int64_t)d->k * d->r * d->s * (d->c / d->groups);
  int64_t m_total = (int64_t)d->n * d->p * d->q;
  int64_t xblocks = ceil_div(wsize, 256);
  int64_t chunks = ceil_div((int64_t)sm_count() * 4, xblocks);
  if (chunks > ceil_div(m_total, 64)) chunks = ceil_div(m_total, 64);
  if (chunks < 1) chunks = 1;
  if (chunks > 1024) chunks = 1024;
  return (int)chunks;
}

size_t direct_workspace_size(const ttb_conv_desc* d, int pass) {
  if (pass != 2) return 0;
  int64_t wsize = (int64_t)d->k * d->r * d->s * (d->c / d->groups);
  return (size_t)wgrad_chunks(d) * wsize * sizeof(float);
}

bool pointwise_narrow(const ttb_conv_desc* d);
static int narrow_fprop(const ttb_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st);
static int narrow_dgrad(const ttb_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st);
static int narrow_wgrad(const ttb_conv_desc* d, const float* x, const float* dy, float* dw, void* ws, cudaStream_t st);

int direct_fprop(const ttb_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                 cudaStream_t st) {
  int64_t total = (int64_t)d->n * d->p * d->q * d->k;
  if (total <= 0) return 0;
  if (pointwise_narrow(d)) return narrow_fprop(d, x, w, bias, y, st);
  launch_k(direct_fprop_kernel, elementwise_grid(total, 256, 16), 256, 0, st, *d, x, w, bias, y, total);
  return check_launch("conv2d_fprop(direct)");
}

int direct_dgrad(const ttb_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st) {
  int64_t total = (int64_t)d->n * d->h * d->w * d->c;
  if (total <= 0) return 0;
  if (pointwise_narrow(d)) return narrow_dgrad(d, dy, w, dx, st);
  launch_k(direct_dgrad_kernel, elementwise_grid(total, 256, 16), 256, 0, st, *d, dy, w, dx, total);
  return check_launch("conv2d_dgrad(direct)");
}

int direct_wgrad(const ttb_conv_desc* d, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes,
                 cudaStream_t st) {
  int64_t wsize = (int64_t)d->k * d->r * d->s * (d->c / d->groups);
  if (wsize <= 0) return 0;
  int chunks = wgrad_chunks(d);
  TTB_REQUIRE(ws != nullptr && ws_bytes >= (size_t)chunks * wsize * sizeof(float),
              "conv2d_wgrad(direct): workspace of %zu bytes needed, %zu given", (size_t)chunks * wsize * sizeof(float),
              ws_bytes);
  int64_t m_total = (int64_t)d->n * d->p * d->q;
  if (pointwise_narrow(d)) return narrow_wgrad(d, x, dy, dw, ws, st);
  int64_t ppc = ceil_div(m_total > 0 ? m_total : 1, chunks);
  dim3 grid((unsigned)ceil_div(wsize, 256), chunks);
  launch_k(direct_wgrad_kernel, grid, 256, 0, st, *d, x, dy, (float*)ws, wsize, ppc);
  if (check_launch("conv2d_wgrad(direct)")) return 1;
  launch_k(wgrad_reduce_kernel, (unsigned)ceil_div(wsize, 256), 256, 0, st, (const float*)ws, chunks, wsize, dw);
  return check_launch("conv2d_wgrad(direct reduce)");
}

}  // namespace ttb

using namespace ttb;

namespace ttb {
int column_sums(const float* x, int64_t m, int c, float* out, cudaStream_t st);  // batchnorm.cu
}

extern "C" int ttb_bias_grad(const float* dy, float* db, int64_t m, int k, void* stream) {
  if (k <= 0) return 0;
  return ttb::column_sums(dy, m, k, db, as_stream(stream));
}

// ---------------------------------------------------------------------------------------------------------
// Pointwise (1 x 1, unit stride, no padding) convolutions with very few filters - a segmentation head such as the UNet's
// Conv2d(32, 1, 1) - are pure HBM streams over the activation; the generic kernels above give such a layer one thread per
// output element / per weight (fprop 256 us, dgrad 165 us, wgrad 2.4 ms at 8 x 512 x 512 x 32, measured).  Here a pixel is
// handled by C/4 neighbouring lanes holding one float4 of channels each: coalesced 16-byte accesses, exact fp32 FMA.
//   fprop : y[p][k]  = sum_c x[p][c] w[k][c] (+ b[k])          lane-group shuffle reduction
//   dgrad : dx[p][c] = sum_k dy[p][k] w[k][c]
//   wgrad : dw[k][c] = sum_p dy[p][k] x[p][c]                  per-thread partial -> block -> fixed-order chunk sum
// Reference semantics: grad_nn.py:595-682 with kh = kw = 1.
// ---------------------------------------------------------------------------------------------------------
namespace ttb {

constexpr int kNarrowMaxK = 4;

// (conv_api.cu keeps such a layer off the tensor path in every math mode: it is an HBM stream, not a GEMM)
bool pointwise_narrow(const ttb_conv_desc* d) {
  if (d->r != 1 || d->s != 1 || d->groups != 1 || d->stride_h != 1 || d->stride_w != 1 || d->pad_h != 0 || d->pad_w != 0)
    return false;
  if (d->k < 1 || d->k > kNarrowMaxK) return false;
  const int c = d->c;
  return c >= 4 && c <= 128 && (c & (c - 1)) == 0;  // C/4 lanes per pixel, a power of two <= 32
}

template <int K>
__global__ void __launch_bounds__(256)
narrow_fprop_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ y, int64_t pixels, int c) {
  pdl_entry();
  const int lpp = c >> 2;                       // lanes per pixel
  const int sub = threadIdx.x & (lpp - 1);      // which float4 of the pixel
  float4 wk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) wk[k] = ld_f4(w + (int64_t)k * c + sub * 4);
  const int64_t ppb = blockDim.x / lpp;         // pixels per block per iteration
  const int64_t per_sweep = (int64_t)gridDim.x * ppb;
  const int64_t sweeps = (pixels + per_sweep - 1) / per_sweep;  // the same trip count for every thread: full-warp shuffles below
  for (int64_t it = 0; it < sweeps; ++it) {
    const int64_t p = it * per_sweep + (int64_t)blockIdx.x * ppb + threadIdx.x / lpp;
    const bool live = p < pixels;
    const float4 v = live ? ld_f4_stream(x + p * c + sub * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      acc[k] = fmaf(v.x, wk[k].x, fmaf(v.y, wk[k].y, fmaf(v.z, wk[k].z, v.w * wk[k].w)));
      for (int o = lpp >> 1; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    if (live && sub == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) y[p * K + k] = acc[k] + (bias ? bias[k] : 0.f);
    }
  }
}

template <int K>
__global__ void __launch_bounds__(256)
narrow_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int64_t pixels, int c) {
  pdl_entry();
  const int lpp = c >> 2;
  const int sub = threadIdx.x & (lpp - 1);
  float4 wk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) wk[k] = ld_f4(w + (int64_t)k * c + sub * 4);
  const int64_t ppb = blockDim.x / lpp;
  for (int64_t p = (int64_t)blockIdx.x * ppb + threadIdx.x / lpp; p < pixels; p += (int64_t)gridDim.x * ppb) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float g = __ldg(dy + p * K + k);
      acc.x = fmaf(g, wk[k].x, acc.x); acc.y = fmaf(g, wk[k].y, acc.y);
      acc.z = fmaf(g, wk[k].z, acc.z); acc.w = fmaf(g, wk[k].w, acc.w);
    }
    st_f4(dx + p * c + sub * 4, acc);
  }
}

// partial[blockIdx.x][k][c]: block sums of its pixels (lane groups of one block are summed through shared memory in fixed
// order; the chunk sum over blocks is wgrad_reduce_kernel's fixed order: deterministic)
template <int K>
__global__ void __launch_bounds__(256)
narrow_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ partial, int64_t pixels,
                    int c) {
  pdl_entry();
  __shared__ float4 red[256];
  const int lpp = c >> 2;
  const int sub = threadIdx.x & (lpp - 1);
  const int grp = threadIdx.x / lpp;
  const int64_t ppb = blockDim.x / lpp;
  float4 acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t p = (int64_t)blockIdx.x * ppb + grp; p < pixels; p += (int64_t)gridDim.x * ppb) {
    const float4 v = ld_f4_stream(x + p * c + sub * 4);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float g = __ldg(dy + p * K + k);
      acc[k].x = fmaf(g, v.x, acc[k].x); acc[k].y = fmaf(g, v.y, acc[k].y);
      acc[k].z = fmaf(g, v.z, acc[k].z); acc[k].w = fmaf(g, v.w, acc[k].w);
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    red[threadIdx.x] = acc[k];
    __syncthreads();
    if (grp == 0) {
      float4 t = red[sub];
      for (int g2 = 1; g2 < (int)ppb; ++g2) {
        const float4 u = red[g2 * lpp + sub];
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      st_f4(partial + ((int64_t)blockIdx.x * K + k) * c + sub * 4, t);
    }
    __syncthreads();
  }
}

#define TTB_NARROW_K(KERNEL, ...)                                                                        \
  switch (d->k) {                                                                                        \
    case 1: launch_k(KERNEL<1>, grid, 256, 0, st, __VA_ARGS__); break;                                   \
    case 2: launch_k(KERNEL<2>, grid, 256, 0, st, __VA_ARGS__); break;                                   \
    case 3: launch_k(KERNEL<3>, grid, 256, 0, st, __VA_ARGS__); break;                                   \
    default: launch_k(KERNEL<4>, grid, 256, 0, st, __VA_ARGS__); break;                                  \
  }

static int narrow_fprop(const ttb_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
  const int64_t pixels = (int64_t)d->n * d->p * d->q;
  const int grid = elementwise_grid(pixels * (d->c / 4), 256);
  TTB_NARROW_K(narrow_fprop_kernel, x, w, bias, y, pixels, d->c)
  return check_launch("conv2d_fprop(pointwise)");
}

static int narrow_dgrad(const ttb_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st) {
  const int64_t pixels = (int64_t)d->n * d->p * d->q;
  const int grid = elementwise_grid(pixels * (d->c / 4), 256);
  TTB_NARROW_K(narrow_dgrad_kernel, dy, w, dx, pixels, d->c)
  return check_launch("conv2d_dgrad(pointwise)");
}

static int narrow_wgrad(const ttb_conv_desc* d, const float* x, const float* dy, float* dw, void* ws, cudaStream_t st) {
  const int64_t pixels = (int64_t)d->n * d->p * d->q;
  const int64_t wsize = (int64_t)d->k * d->c;
  const int grid = wgrad_chunks(d);  // (the workspace holds this many partial buffers: direct_workspace_size)
  TTB_NARROW_K(narrow_wgrad_kernel, x, dy, reinterpret_cast<float*>(ws), pixels, d->c)
  if (check_launch("conv2d_wgrad(pointwise)")) return 1;
  launch_k(wgrad_reduce_kernel, (unsigned)ceil_div(wsize, 256), 256, 0, st, (const float*)ws, grid, wsize, dw);
  return check_launch("conv2d_wgrad(pointwise reduce)");
}

}  // namespace ttb
