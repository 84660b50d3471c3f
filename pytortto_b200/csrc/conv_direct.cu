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

int direct_fprop(const ttb_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                 cudaStream_t st) {
  int64_t total = (int64_t)d->n * d->p * d->q * d->k;
  if (total <= 0) return 0;
  launch_k(direct_fprop_kernel, elementwise_grid(total, 256, 16), 256, 0, st, *d, x, w, bias, y, total);
  return check_launch("conv2d_fprop(direct)");
}

int direct_dgrad(const ttb_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st) {
  int64_t total = (int64_t)d->n * d->h * d->w * d->c;
  if (total <= 0) return 0;
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
