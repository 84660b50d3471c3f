// placeholder: tcgen05 implicit-GEMM kernels land here
#include "common.cuh"
namespace ttb {
bool igemm_supported(const ttb_conv_desc*, int) { return false; }
size_t igemm_workspace_size(const ttb_conv_desc*, int) { return 0; }
int igemm_fprop(const ttb_conv_desc*, const float*, const float*, const float*, float*, void*, size_t, cudaStream_t) { set_error("igemm: not built"); return 3; }
int igemm_dgrad(const ttb_conv_desc*, const float*, const float*, float*, void*, size_t, cudaStream_t) { set_error("igemm: not built"); return 3; }
int igemm_wgrad(const ttb_conv_desc*, const float*, const float*, float*, void*, size_t, cudaStream_t) { set_error("igemm: not built"); return 3; }
}
