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
  return igemm_supported(d, pass) ? 1 : 0;
}

size_t ttb_conv2d_workspace_size(const ttb_conv_desc* d, int pass) {
  if (!d) return 0;
  if (d->math_mode != TTB_MATH_FP32 && igemm_supported(d, pass)) return igemm_workspace_size(d, pass);
  return direct_workspace_size(d, pass);
}

int ttb_conv2d_fprop(const ttb_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                     void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = validate(d, "conv2d_fprop")) return rc;
  if (d->math_mode != TTB_MATH_FP32 && igemm_supported(d, 0))
    return igemm_fprop(d, x, w, bias, y, workspace, workspace_bytes, as_stream(stream));
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
  return direct_wgrad(d, x, dy, dw, workspace, workspace_bytes, as_stream(stream));
}

}  // extern "C"
