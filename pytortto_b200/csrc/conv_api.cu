// C-ABI entry points for convolution: validation + dispatch between the tcgen05 implicit-GEMM kernels
// (conv_igemm.cu) and the generic exact-fp32 direct kernels (conv_direct.cu).
#include "common.cuh"

namespace ttb {
// conv_direct.cu
size_t direct_workspace_size(const ttb_conv_desc* d, int pass);
int direct_fprop(const ttb_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st);
int direct_dgrad(const ttb_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st);
int direct_wgrad(const ttb_conv_desc* d, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes,
                 cudaStream_t st);
// conv_igemm.cu
bool igemm_supported(const ttb_conv_desc* d, int pass);
size_t igemm_workspace_size(const ttb_conv_desc* d, int pass);
int igemm_fprop(const ttb_conv_desc* d, const float* x, const float* w, const float* bias, float* y, void* ws,
                size_t ws_bytes, cudaStream_t st);
int igemm_dgrad(const ttb_conv_desc* d, const float* dy, const float* w, float* dx, void* ws, size_t ws_bytes,
                cudaStream_t st);
int igemm_wgrad(const ttb_conv_desc* d, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes,
                cudaStream_t st);

// ---- channel padding: problems whose input-channel count is not a multiple of 32 (network stems with C = 3, narrow
// stages with C = 16 ...) run on the tensor path over a zero-padded copy [rows][Cp], Cp = round_up(C, 32).  The
// zero channels contribute nothing; for the 3-channel CIFAR stem this replaces the CUDA-core direct kernels.
__global__ void __launch_bounds__(256)
pad_channels_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t rows, int c, int cp) {
  const int q = cp / 4;
  int64_t total = rows * q;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    int64_t r = t / q;
    int c0 = (int)(t % q) * 4;
    const float* ps = src + r * c;
    float4 v;
    v.x = c0 + 0 < c ? ps[c0 + 0] : 0.f;
    v.y = c0 + 1 < c ? ps[c0 + 1] : 0.f;
    v.z = c0 + 2 < c ? ps[c0 + 2] : 0.f;
    v.w = c0 + 3 < c ? ps[c0 + 3] : 0.f;
    *reinterpret_cast<float4*>(dst + r * cp + c0) = v;
  }
}

__global__ void __launch_bounds__(256)
unpad_channels_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t rows, int c, int cp) {
  int64_t total = rows * c;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    int64_t r = t / c;
    int ch = (int)(t % c);
    dst[t] = src[r * cp + ch];
  }
}

static inline int padded_c(const ttb_conv_desc* d) { return (d->c + 31) / 32 * 32; }
static inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

// true when pass 0 (fprop) / 2 (wgrad) should run on the tensor path over channel-padded operands
static bool pad_path(const ttb_conv_desc* d, int pass, ttb_conv_desc* padded) {
  if (d->math_mode == TTB_MATH_FP32 || d->groups != 1 || d->c % 32 == 0 || pass == 1) return false;
  ttb_conv_desc p = *d;
  p.c = padded_c(d);
  if (!igemm_supported(&p, pass)) return false;
  if (padded) *padded = p;
  return true;
}

static int validate(const ttb_conv_desc* d, const char* what) {
  TTB_REQUIRE(d != nullptr, "%s: null descriptor", what);
  TTB_REQUIRE(d->n >= 0 && d->c > 0 && d->h > 0 && d->w > 0 && d->k > 0 && d->r > 0 && d->s > 0, "%s: bad sizes", what);
  TTB_REQUIRE(d->groups > 0 && d->c % d->groups == 0 && d->k % d->groups == 0,
              "%s: channels (%d in, %d out) not divisible by groups=%d", what, d->c, d->k, d->groups);
  TTB_REQUIRE(d->stride_h > 0 && d->stride_w > 0 && d->dil_h > 0 && d->dil_w > 0 && d->pad_h >= 0 && d->pad_w >= 0,
              "%s: bad stride/dilation/padding", what);
  TTB_REQUIRE(d->p >= 0 && d->q >= 0, "%s: bad output size", what);
  return 0;
}
}  // namespace ttb

using namespace ttb;

extern "C" {

int ttb_conv2d_tensor_path_supported(const ttb_conv_desc* d, int pass) {
  if (!d || d->math_mode == TTB_MATH_FP32) return 0;
  return (igemm_supported(d, pass) || pad_path(d, pass, nullptr)) ? 1 : 0;
}

size_t ttb_conv2d_workspace_size(const ttb_conv_desc* d, int pass) {
  if (!d) return 0;
  if (d->math_mode != TTB_MATH_FP32 && igemm_supported(d, pass)) return igemm_workspace_size(d, pass);
  ttb_conv_desc p;
  if (pad_path(d, pass, &p)) {
    const size_t xb = align256((size_t)d->n * d->h * d->w * p.c * sizeof(float));
    const size_t wb = align256((size_t)d->k * d->r * d->s * p.c * sizeof(float));
    return xb + wb + align256(igemm_workspace_size(&p, pass));
  }
  return direct_workspace_size(d, pass);
}

int ttb_conv2d_fprop(const ttb_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                     void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = validate(d, "conv2d_fprop")) return rc;
  if (d->math_mode != TTB_MATH_FP32 && igemm_supported(d, 0))
    return igemm_fprop(d, x, w, bias, y, workspace, workspace_bytes, as_stream(stream));
  ttb_conv_desc p;
  if (pad_path(d, 0, &p)) {
    cudaStream_t st = as_stream(stream);
    const int64_t xrows = (int64_t)d->n * d->h * d->w, wrows = (int64_t)d->k * d->r * d->s;
    const size_t xb = align256((size_t)xrows * p.c * sizeof(float)), wb = align256((size_t)wrows * p.c * sizeof(float));
    TTB_REQUIRE(workspace != nullptr && workspace_bytes >= xb + wb, "conv2d_fprop: workspace of %zu bytes needed, %zu given",
                xb + wb, workspace_bytes);
    float* xp = reinterpret_cast<float*>(workspace);
    float* wp = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + xb);
    pad_channels_kernel<<<elementwise_grid(xrows * p.c / 4, 256), 256, 0, st>>>(x, xp, xrows, d->c, p.c);
    pad_channels_kernel<<<elementwise_grid(wrows * p.c / 4, 256), 256, 0, st>>>(w, wp, wrows, d->c, p.c);
    if (check_launch("pad_channels")) return 1;
    return igemm_fprop(&p, xp, wp, bias, y, reinterpret_cast<char*>(workspace) + xb + wb, workspace_bytes - xb - wb, st);
  }
  return direct_fprop(d, x, w, bias, y, as_stream(stream));
}

int ttb_conv2d_dgrad(const ttb_conv_desc* d, const float* dy, const float* w, float* dx, void* workspace,
                     size_t workspace_bytes, void* stream) {
  if (int rc = validate(d, "conv2d_dgrad")) return rc;
  if (d->math_mode != TTB_MATH_FP32 && igemm_supported(d, 1))
    return igemm_dgrad(d, dy, w, dx, workspace, workspace_bytes, as_stream(stream));
  return direct_dgrad(d, dy, w, dx, as_stream(stream));
}

int ttb_conv2d_wgrad(const ttb_conv_desc* d, const float* x, const float* dy, float* dw, void* workspace,
                     size_t workspace_bytes, void* stream) {
  if (int rc = validate(d, "conv2d_wgrad")) return rc;
  if (d->math_mode != TTB_MATH_FP32 && igemm_supported(d, 2))
    return igemm_wgrad(d, x, dy, dw, workspace, workspace_bytes, as_stream(stream));
  ttb_conv_desc p;
  if (pad_path(d, 2, &p)) {
    cudaStream_t st = as_stream(stream);
    const int64_t xrows = (int64_t)d->n * d->h * d->w, wrows = (int64_t)d->k * d->r * d->s;
    const size_t xb = align256((size_t)xrows * p.c * sizeof(float)), wb = align256((size_t)wrows * p.c * sizeof(float));
    TTB_REQUIRE(workspace != nullptr && workspace_bytes >= xb + wb, "conv2d_wgrad: workspace of %zu bytes needed, %zu given",
                xb + wb, workspace_bytes);
    float* xp = reinterpret_cast<float*>(workspace);
    float* dwp = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + xb);
    pad_channels_kernel<<<elementwise_grid(xrows * p.c / 4, 256), 256, 0, st>>>(x, xp, xrows, d->c, p.c);
    if (check_launch("pad_channels")) return 1;
    if (int rc = igemm_wgrad(&p, xp, dy, dwp, reinterpret_cast<char*>(workspace) + xb + wb, workspace_bytes - xb - wb, st))
      return rc;
    unpad_channels_kernel<<<elementwise_grid(wrows * d->c, 256), 256, 0, st>>>(dwp, dw, wrows, d->c, p.c);
    return check_launch("unpad_channels");
  }
  return direct_wgrad(d, x, dy, dw, workspace, workspace_bytes, as_stream(stream));
}

}  // extern "C"
