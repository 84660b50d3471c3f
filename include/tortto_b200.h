/*
 * tortto_b200.h — C ABI of libtortto_b200.so: the B200 (sm_100a) implementation of the conv / BN / ReLU / max-pool
 * hot path of samrere/pytortto (tortto v1.3.4).
 *
 * This is the drop-in boundary: plain C structs, raw device pointers, a CUDA stream handle passed
 * as void*; the CALLER owns every buffer (inputs, outputs, workspace); each call enqueues work on the given stream
 * and returns without synchronising; return value 0 = ok, non-zero = error (text via ttb_last_error()).
 * No exceptions cross the boundary, no torch / cupy types appear in any signature.
 *
 * Reference interface each entry point replaces (paths relative to /root/reference/src/tortto/):
 *   ttb_conv2d_fprop        Convolution.forward            autograd/grad_nn.py:685-715  (_conv2d :595-643)
 *   ttb_conv2d_dgrad        _conv2d_backward_x             autograd/grad_nn.py:659-682  (also TransposedConvolution.forward :755)
 *   ttb_conv2d_wgrad        _conv2d_backward_w             autograd/grad_nn.py:646-656
 *   ttb_bias_grad           gd0.sum((0,2,3))               autograd/grad_nn.py:727-728
 *   ttb_bn_*                BatchNorm.forward / backward   autograd/grad_nn.py:909-989
 *   ttb_relu_fwd / _bwd     Relu.forward / backward        autograd/grad_nn.py:50-69
 *   ttb_maxpool2d_fwd/_bwd  _max_pool2d / _backward        autograd/grad_nn.py:784-826
 *   ttb_add / ttb_axpy      Add.forward, grad accumulation tensor.py:597-599, autograd/function.py:88-93
 *   ttb_sgd_step            optim/_functional.py:4-22
 *   ttb_nchw_to_nhwc etc.   the device-memory layer, xparray.py:44-78 / ToCopy autograd/grad_fcn.py:849-878
 *
 * Data layout in HBM: 4-D activations are stored NHWC (channels innermost) as float32; the logical shape seen by
 * the Python host layer stays (N, C, H, W).  Conv weights are stored "KRSC": [Cout][kh][kw][Cin/groups] float32
 * (the same bytes as a channels-last view of the reference's (Cout, Cin/g, kh, kw) parameter).
 */
#ifndef TORTTO_B200_H
#define TORTTO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* math modes for the convolution contractions */
enum {
  TTB_MATH_FP32 = 0, /* CUDA-core FFMA, exact fp32 (generic fallback: any groups / dilation / channel count) */
  TTB_MATH_TF32 = 1, /* tcgen05.mma kind::tf32, fp32 accumulate in TMEM */
  TTB_MATH_BF16 = 2  /* tcgen05.mma kind::f16 (bf16 operands), fp32 accumulate in TMEM */
};

/* One 2-D convolution problem (cross-correlation, zero padding), reference semantics of F.conv2d
 * (nn/functional.py:80-88).  All sizes in elements. */
typedef struct ttb_conv_desc {
  int32_t n, c, h, w;          /* input  x: logical (N, C, H, W)                                        */
  int32_t k;                   /* output channels                                                       */
  int32_t r, s;                /* filter height, width                                                  */
  int32_t stride_h, stride_w;
  int32_t pad_h, pad_w;
  int32_t dil_h, dil_w;
  int32_t groups;
  int32_t p, q;                /* output y: logical (N, K, P, Q); P = floor((H+2ph-dh(r-1)-1)/sh+1)     */
  int32_t math_mode;           /* TTB_MATH_*                                                            */
} ttb_conv_desc;

typedef struct ttb_pool_desc {
  int32_t n, c, h, w;          /* input                                                                 */
  int32_t kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
  int32_t p, q;                /* output spatial size (the host computes floor/ceil_mode geometry)      */
} ttb_pool_desc;

/* ---- library ------------------------------------------------------------------------------------------- */
const char* ttb_last_error(void);            /* thread-local text of the last failure                    */
int ttb_version(void);                       /* ABI version, bumps on any signature change               */
int ttb_device_sm_count(int* out);           /* SM count of the current device                           */
/* 1 if the tcgen05 tensor path supports this problem in the requested math mode, else 0 (then the FP32 path runs) */
int ttb_conv2d_tensor_path_supported(const ttb_conv_desc* d, int pass /*0 fprop, 1 dgrad, 2 wgrad*/);

/* which kernel family a pass runs on: 0 exact fp32 direct kernels, 1 tcgen05 implicit GEMM (im2col-mode TMA), 2 tcgen05
 * flat-shift halo tile with shared-memory-resident weights (an experiment: tuning build only, never in the release library) */
int ttb_conv2d_kernel_variant(const ttb_conv_desc* d, int pass /*0 fprop, 1 dgrad, 2 wgrad*/);

/* ---- layout --------------------------------------------------------------------------------------------- */
int ttb_nchw_to_nhwc(const float* src, float* dst, int n, int c, int h, int w, void* stream);
int ttb_nhwc_to_nchw(const float* src, float* dst, int n, int c, int h, int w, void* stream);

/* ---- convolution ---------------------------------------------------------------------------------------- */
size_t ttb_conv2d_workspace_size(const ttb_conv_desc* d, int pass /*0 fprop, 1 dgrad, 2 wgrad*/);
/* y[N,P,Q,K] = conv(x[N,H,W,C], w[K,R,S,C/g]) (+ bias[K] if bias != NULL) */
int ttb_conv2d_fprop(const ttb_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                     void* workspace, size_t workspace_bytes, void* stream);
/* dx[N,H,W,C] = input gradient; every element of dx is written (zeros where no output touches it) */
int ttb_conv2d_dgrad(const ttb_conv_desc* d, const float* dy, const float* w, float* dx,
                     void* workspace, size_t workspace_bytes, void* stream);
/* dw[K,R,S,C/g] = weight gradient (overwritten, not accumulated) */
int ttb_conv2d_wgrad(const ttb_conv_desc* d, const float* x, const float* dy, float* dw,
                     void* workspace, size_t workspace_bytes, void* stream);
/* Fused epilogue of the forward convolution on the tensor path ("fused bias/BN-scale/ReLU epilogues"): what happens to an
 * accumulator element of output channel k before it is stored, in this order.  Replaces the separate full-tensor passes
 * of the reference that follow a convolution: the bias add (autograd/grad_nn.py:714-715), an eval-mode BatchNorm folded
 * to per-channel scale / shift (:942-959), the residual `Add` (tensor.py:597-599), `Relu` (:58), and the two statistics
 * reductions of the training-mode BatchNorm that reads the output next (xp.mean / xp.var, :923-924). */
typedef struct ttb_conv_epilogue {
  const float* scale;    /* [K] or NULL: v = acc * scale[k]                                                          */
  const float* bias;     /* [K] or NULL: v += bias[k]                                                                */
  const float* residual; /* [N,P,Q,K] or NULL: v += residual[same element] (may alias y: in-place accumulate)        */
  int32_t relu;          /* != 0: v = max(v, 0)                                                                      */
  double* stats;         /* NULL, or [ttb_conv2d_fprop_stats_chunks(d)][2][K]: per-chunk sum(v), sum(v*v) of the stored
                            values - the partial buffer ttb_bn_finalize / ttb_comm_bn_finalize take instead of running
                            ttb_bn_stats over y (fixed tile -> chunk assignment: deterministic)                         */
} ttb_conv_epilogue;
/* 1 if fprop of this problem runs on the tensor path (TF32 / BF16 math, groups == 1, ...), i.e. can take an epilogue */
int ttb_conv2d_fused_epilogue_supported(const ttb_conv_desc* d);
/* number of chunks (rows of the statistics partial buffer) fprop of this problem writes; 0 = no fused epilogue */
int ttb_conv2d_fprop_stats_chunks(const ttb_conv_desc* d);
/* ttb_conv2d_fprop with a fused epilogue (ep == NULL: plain store); same workspace as ttb_conv2d_fprop */
int ttb_conv2d_fprop_fused(const ttb_conv_desc* d, const float* x, const float* w, const ttb_conv_epilogue* ep, float* y,
                           void* workspace, size_t workspace_bytes, void* stream);
/* db[K] = sum over rows of dy[M,K] */
int ttb_bias_grad(const float* dy, float* db, int64_t m, int k, void* stream);

/* Multi-tensor forms of the two small per-layer helpers of the backward pass: a training step has ~20 dgrad weight
 * re-orderings and ~20 wgrad split reductions, each a latency-bound launch on its own; these do all layers in one.
 *   ttb_conv2d_dgrad_prepacked_supported  1 if dgrad of this problem can take pre-packed weights (tensor path without a
 *                                         staged copy), else use ttb_conv2d_dgrad
 *   ttb_conv2d_dgrad_pack_weights         w[i] ([K][R][S][C]) -> w_packed[i] ([C][R][S][K], same size) for count layers
 *   ttb_conv2d_dgrad_prepacked            ttb_conv2d_dgrad on weights packed by the call above (no workspace); accum (may be
 *                                         NULL, may alias dx): a gradient already pending for the same tensor, added in the
 *                                         epilogue (the engine's `grad += new`, tensor.py:597-599, without its own pass)
 *   ttb_conv2d_wgrad_partial              ttb_conv2d_wgrad without the split reduction: leaves *splits_out partial
 *                                         buffers of K*R*S*C floats at *partials_out (inside workspace); <= 1: dw is final
 *   ttb_sum_splits_multi                  outs[i][e] = sum_s partials[i][s*sizes[i] + e], fixed order, count tensors   */
int ttb_conv2d_dgrad_prepacked_supported(const ttb_conv_desc* d);
int ttb_conv2d_dgrad_pack_weights(int count, const ttb_conv_desc* const* descs, const float* const* w,
                                  float* const* w_packed, void* stream);
int ttb_conv2d_dgrad_prepacked(const ttb_conv_desc* d, const float* dy, const float* w_packed, const float* accum, float* dx,
                               void* stream);
int ttb_conv2d_wgrad_partial(const ttb_conv_desc* d, const float* x, const float* dy, float* dw, void* workspace,
                             size_t workspace_bytes, int* splits_out, const float** partials_out, void* stream);
int ttb_sum_splits_multi(int count, const float* const* partials, const int* splits, const int64_t* sizes,
                         float* const* outs, void* stream);

/* bf16-operand forms (TTB_MATH_BF16 descriptors): the operands are ALREADY bf16 in HBM - activations / gradients NHWC
 * bf16 as co-written by ttb_bn_apply / ttb_bn_bwd_apply / ttb_relu_fwd (or converted by ttb_to_bf16), weights bf16 as
 * written by ttb_conv2d_pack_weights_bf16 - so one call is exactly one tcgen05 implicit-GEMM launch; accumulation and
 * outputs stay fp32.  Replaces the same reference sites as ttb_conv2d_fprop / _dgrad / _wgrad (grad_nn.py:595-682).
 *   ttb_conv2d_bf16_supported       1 if the pass can take bf16 operands as they are (groups == 1, reduction channels in
 *                                   whole 64-channel K-blocks), else use the fp32-operand entry points
 *   ttb_conv2d_workspace_size_bf16  bytes of workspace ttb_conv2d_wgrad_bf16 needs (pixel-split partial sums)
 *   ttb_conv2d_pack_weights_bf16    w[i] fp32 [K][R][S][C] -> w_bf16[i] (same order; fprop) and wt_bf16[i] ([C][R][S][K];
 *                                   dgrad) for `count` layers in one launch; either destination entry may be NULL
 *   ttb_to_bf16                     dst[i] = bf16(src[i]), n % 4 == 0 */
int ttb_conv2d_bf16_supported(const ttb_conv_desc* d, int pass /*0 fprop, 1 dgrad, 2 wgrad*/);
size_t ttb_conv2d_workspace_size_bf16(const ttb_conv_desc* d, int pass);
int ttb_conv2d_fprop_bf16(const ttb_conv_desc* d, const void* x_bf16, const void* w_bf16, const ttb_conv_epilogue* ep /*or NULL*/,
                          float* y, void* stream);
int ttb_conv2d_dgrad_bf16(const ttb_conv_desc* d, const void* dy_bf16, const void* w_packed_bf16, const float* accum /*or NULL*/,
                          float* dx, void* stream);
int ttb_conv2d_wgrad_bf16(const ttb_conv_desc* d, const void* x_bf16, const void* dy_bf16, float* dw, void* workspace,
                          size_t workspace_bytes, void* stream);

/* dgrad fused with the statistics pass of the BatchNorm backward that consumes its output.  In a network
 * conv(relu(bn(x))) the gradient dgrad produces is what BatchNorm.backward (reference autograd/grad_nn.py:967-989) reduces
 * next: sum(g) and sum(g * (x - mean)) per channel, g = dx masked by the ReLU (recomputed as fmaf(x - mean, rscale, rshift)
 * > 0, the expression forward evaluated).  The dgrad epilogue emits those sums of what it stores, so ttb_bn_bwd_reduce (two
 * full reads) is not run; ttb_bn_bwd_finalize / ttb_comm_bn_bwd_finalize take `partials` as they take ttb_bn_bwd_reduce's.
 *   ttb_conv2d_dgrad_bn_stats_chunks   rows of `partials` ([chunks][2][C] doubles) this problem writes; 0 = not available
 *                                      (strided / grouped / staged problems, small outputs on 256-wide tiles)
 *   ttb_conv2d_dgrad_bn                dy, w_packed: fp32 ([C][R][S][K] from ttb_conv2d_dgrad_pack_weights) for TF32 math,
 *                                      bf16 (ttb_conv2d_pack_weights_bf16) for BF16 math; accum as in ttb_conv2d_dgrad_prepacked */
typedef struct ttb_dgrad_bn_stats {
  const float* x;       /* the BatchNorm's input [N,H,W,C]: same shape and layout as dx                              */
  const float* mean;    /* [C] batch mean saved by forward                                                           */
  const float* rscale;  /* [C] or NULL: scale / shift forward normalised with, when a ReLU follows the BatchNorm     */
  const float* rshift;
  double* partials;     /* [ttb_conv2d_dgrad_bn_stats_chunks(d)][2][C]                                               */
} ttb_dgrad_bn_stats;
int ttb_conv2d_dgrad_bn_stats_chunks(const ttb_conv_desc* d);
int ttb_conv2d_dgrad_bn(const ttb_conv_desc* d, const void* dy, const void* w_packed, const float* accum, float* dx,
                        const ttb_dgrad_bn_stats* bn, void* stream);
int ttb_conv2d_pack_weights_bf16(int count, const ttb_conv_desc* const* descs, const float* const* w, void* const* w_bf16,
                                 void* const* wt_bf16, void* stream);
int ttb_to_bf16(const float* src, void* dst, int64_t n, void* stream);

/* ---- batch norm (x viewed as [M = N*H*W rows][C channels]) ---------------------------------------------- */
/* number of row chunks ttb_bn_stats / ttb_bn_bwd_reduce emit partial sums for */
int ttb_bn_num_chunks(int64_t m, int c);
/* partials[chunk][2][C] (double): per-chunk sum(x), sum(x*x) */
int ttb_bn_stats(const float* x, int64_t m, int c, double* partials, int num_chunks, void* stream);
/* out = a + b (the residual `Add`, tensor.py:597-599; out must not alias a or b) AND the statistics partials of the sum, as
 * ttb_bn_stats(out, ...) would write them, in one pass: for the BatchNorm that reads the sum next */
int ttb_add_bn_stats(const float* a, const float* b, float* out, int64_t m, int c, double* partials, int num_chunks,
                     void* stream);
/* sums[2][C] = sum over chunks (double).  This is the buffer a data-parallel run all-reduces (SyncBN). */
int ttb_bn_reduce_partials(const double* partials, int num_chunks, int c2, double* sums, void* stream);
/* from sums[num_chunks][2][C] (per-chunk partials, summed here in fixed order; num_chunks = 1 for an already reduced /
 * all-reduced buffer) over `count` elements per channel: mean, var_eps = biased var + eps, sd = sqrt(var_eps)
 * (what BatchNorm.forward saves, grad_nn.py:962-963), scale = gamma/sd, shift = beta (0 without affine), and, if
 * running_mean/var != NULL, running = (1-momentum)*running + momentum*{mean, var*count/(count-1)} (:925-930). */
int ttb_bn_finalize(const double* sums, int num_chunks, int64_t count, int c, float eps, float momentum,
                    const float* gamma, const float* beta, float* running_mean, float* running_var,
                    float* mean, float* var_eps, float* sd, float* scale, float* shift, void* stream);
/* eval mode / precomputed statistics: fills var_eps, sd, scale, shift from given mean & var */
int ttb_bn_prepare_eval(const float* mean_in, const float* var_in, int c, float eps, const float* gamma,
                        const float* beta, float* mean, float* var_eps, float* sd, float* scale, float* shift,
                        void* stream);
/* eval-mode BatchNorm behind a convolution folded to the conv epilogue's per-channel scale / bias (ttb_conv_epilogue):
 * scale = gamma / sqrt(var + eps), bias = beta + (conv_bias - mean) * scale; gamma, beta, conv_bias may be NULL */
int ttb_bn_fold_eval(const float* mean, const float* var, int c, float eps, const float* gamma, const float* beta,
                     const float* conv_bias, float* scale, float* bias, void* stream);
/* y = (x - mean[c])*scale[c] + beta[c] (the reference's order of operations, grad_nn.py:942-959, with gamma/sd folded
 * into scale; `beta` = the `shift` row the finalize entry points write); relu != 0 fuses y = max(y, 0).
 * y_bf16 (may be NULL): the same values rounded to bf16, co-written for the bf16 tensor path (ttb_conv2d_*_bf16). */
int ttb_bn_apply(const float* x, float* y, int64_t m, int c, const float* mean, const float* scale, const float* beta,
                 int relu, void* y_bf16, void* stream);
/* y = bn(x) + residual, then max(y, 0) if relu != 0: the tail of a post-activation residual block (BatchNorm.forward
 * grad_nn.py:942-959 -> `Add` -> Relu.forward :58) as one pass; same arguments as ttb_bn_apply plus `residual` (shape of x);
 * needs c % 4 == 0 */
int ttb_bn_apply_add(const float* x, const float* residual, float* y, int64_t m, int c, const float* mean, const float* scale,
                     const float* beta, int relu, void* y_bf16, void* stream);
/* partials[chunk][2][C]: sum(dy), sum(dy*(x-mean)).  If relu_out != NULL dy is first masked by (relu_out > 0)
 * (fused ReLU backward, grad_nn.py:64-69).  The mask of a fused BatchNorm+ReLU node comes either from relu_out (the
 * saved ReLU output) or - one read per element cheaper - is recomputed as fmaf(x - mean, relu_scale, relu_shift) > 0
 * with the scale / shift rows forward normalised with (bit-identical to what forward tested); give one of the two or
 * neither. */
int ttb_bn_bwd_reduce(const float* dy, const float* x, const float* mean, const float* relu_out, const float* relu_scale,
                      const float* relu_shift, int64_t m, int c, double* partials, int num_chunks, void* stream);
/* from sums[num_chunks][2][C]: dgamma = sum(dy*(x-mean))/sd, dbeta = sum(dy), and the three per-channel
 * coefficients of dx = c1*(dy - c2 - (x-mean)*c3)  (grad_nn.py:984-988 re-associated) */
int ttb_bn_bwd_finalize(const double* sums, int num_chunks, int64_t count, int c, const float* gamma, const float* var_eps,
                        const float* sd, float* dgamma, float* dbeta, float* coef /*[3][C]*/, void* stream);
/* dx = c1*(dy - c2 - (x-mean)*c3) [+ accum]; accum (may be null): a gradient that already reached the same tensor
 * through another branch, i.e. the engine's `grad += new` (tensor.py:597-599) folded into this pass */
/* dx_bf16 (may be NULL): dx rounded to bf16, co-written for the bf16 tensor path (it is the dY of the producing conv) */
int ttb_bn_bwd_apply(const float* dy, const float* x, const float* mean, const float* relu_out, const float* relu_scale,
                     const float* relu_shift, const float* coef, const float* accum, float* dx, int64_t m, int c,
                     void* dx_bf16, void* stream);

/* ---- relu / elementwise ---------------------------------------------------------------------------------- */
/* y may alias x (in-place); y_bf16 (may be NULL): bf16 copy of y for the bf16 tensor path */
int ttb_relu_fwd(const float* x, float* y, int64_t n, void* y_bf16, void* stream);
int ttb_relu_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream);
int ttb_add(const float* a, const float* b, float* out, int64_t n, void* stream); /* out may alias a or b    */
int ttb_axpy(float alpha, const float* x, float* y, int64_t n, void* stream);     /* y += alpha*x            */
int ttb_scale(float alpha, float* x, int64_t n, void* stream);                    /* x *= alpha              */
int ttb_fill(float value, float* x, int64_t n, void* stream);
/* one SGD update of a flat parameter range (optim/_functional.py:4-22); first_step != 0 means the momentum buffer
 * is initialised to d_p (buf = d_p) instead of buf = momentum*buf + (1-dampening)*d_p */
int ttb_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                 float dampening, float weight_decay, int nesterov, int first_step, void* stream);

/* the same update for `n_tensors` parameter tensors in as few launches as possible (HOST arrays of device pointers,
 * sizes in elements, per-tensor first_step flags; bufs may be NULL when momentum == 0) */
int ttb_sgd_step_multi(int n_tensors, float* const* params, const float* const* grads, float* const* bufs,
                       const int64_t* sizes, const unsigned char* first_step, float lr, float momentum, float dampening,
                       float weight_decay, int nesterov, void* stream);

/* Adam / AdamW (optim/_functional.py:25-68, :71-115) for `n_tensors` parameter tensors in as few launches as possible.
 * `state` = 3 floats on the DEVICE: step, 1 - beta1^step, 1 - beta2^step; ttb_adam_advance increments the step and
 * refreshes the two bias corrections (one 1-thread kernel per optimizer step), so the update is CUDA-graph safe.
 * max_exp_avg_sq (amsgrad) may be NULL; decoupled != 0 selects AdamW's p *= 1 - lr*weight_decay. */
int ttb_adam_advance(float* state, float beta1, float beta2, void* stream);
int ttb_adam_step_multi(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                        float* const* exp_avg_sq, float* const* max_exp_avg_sq, const int64_t* sizes, const float* state,
                        float lr, float beta1, float beta2, float eps, float weight_decay, int decoupled, void* stream);

/* ---- the ops between the conv stacks and the loss --------------------------------------------- */
/* global average pool of an NHWC tensor over H, W (Mean, autograd/grad_fcn.py:1058-1093): y[n][c], dx = dy / (H*W) */
int ttb_mean_hw_fwd(const float* x, float* y, int n, int hw, int c, void* stream);
int ttb_mean_hw_bwd(const float* dy, float* dx, int n, int hw, int c, void* stream);
/* C[m][n] (dense row-major) = sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] (+ bias[n] if bias != NULL); fp32 CUDA-core GEMM
 * for the classifier head (Linear = Transpose + Mm + Add, nn/functional.py:54-63) and its two gradients */
int ttb_matmul(const float* a, const float* b, const float* bias, float* c, int m, int n, int k, int64_t sam, int64_t sak,
               int64_t sbk, int64_t sbn, void* stream);
/* LogSoftmax over the last axis of [rows][cols] (autograd/grad_nn.py:373-392) */
int ttb_log_softmax_fwd(const float* x, float* y, int rows, int cols, void* stream);
int ttb_log_softmax_bwd(const float* dy, const float* y, float* dx, int rows, int cols, void* stream);
/* NLL loss (autograd/grad_nn.py:287-349): reduction 0 none (out[rows]), 1 mean, 2 sum (out[1]); rows whose target ==
 * ignore_index do not contribute; *count (device float) = number of contributing rows, consumed by the backward of mean */
int ttb_nll_loss_fwd(const float* logp, const int64_t* target, int rows, int cols, int64_t ignore_index, int reduction,
                     float* out, float* count, void* stream);
int ttb_nll_loss_bwd(const float* g, const int64_t* target, int rows, int cols, int64_t ignore_index, int reduction,
                     const float* count, float* dx, void* stream);
/* BCE with logits (autograd/grad_nn.py:236-285): reduction 0 none (out[n]), 1 mean, 2 sum (out[1], fixed-order two-stage
 * sum in `workspace` of ttb_bce_logits_workspace_size() bytes); backward dx = (sigmoid(x) - t) * g * scale with g a
 * device scalar (g_per_elem == 0) or per element */
size_t ttb_bce_logits_workspace_size(void);
int ttb_bce_logits_fwd(const float* x, const float* t, int64_t n, int reduction, float* out, void* workspace, void* stream);
int ttb_bce_logits_bwd(const float* x, const float* t, const float* g, int g_per_elem, float scale, int64_t n, float* dx,
                       void* stream);
/* channel-range copy between NHWC tensors viewed as [rows][channels]: dst[r][dst_off + c] = src[r][src_off + c], c < c_copy
 * (Cat along the channel axis and its backward split, autograd/grad_fcn.py:881-904: UNet skip connections) */
int ttb_copy_channels(const float* src, float* dst, int64_t rows, int c_src, int c_dst, int src_off, int dst_off, int c_copy,
                      void* stream);
/* y[r][c] += bias[c] in place (the bias of ConvTranspose2d, whose contraction is a dgrad pass without a bias epilogue) */
int ttb_add_bias(float* y, const float* bias, int64_t rows, int c, void* stream);

/* ---- small-message all-reduce over NVLink peer memory (SyncBN statistics; one process per GPU) ------------------- */
/* cudaMalloc a zeroed communication buffer and export its CUDA-IPC handle (64 bytes) */
int ttb_comm_alloc(size_t bytes, void** dev_ptr, unsigned char* handle_out);
/* map a peer's buffer into this process (peer access is enabled lazily) */
int ttb_comm_open(const unsigned char* handle, void** peer_ptr);
int ttb_comm_close(void* peer_ptr);
int ttb_comm_free(void* dev_ptr);
/* bytes of one slot for exchanges of up to max_values <= 4096 doubles among <= 8 ranks: header + two buffers used alternately
 * (parity of the slot's exchange count, so a rank that runs ahead never overwrites words a slower peer is still reading),
 * each with one area per SOURCE rank - peers push their values into it and the owner polls its own memory; 0 if too large */
size_t ttb_comm_slot_bytes(int max_values);
/* out[i] = sum over ranks (rank order) of sum over chunks of that rank's partials[chunk][i], exchanged through the
 * slot at `slot_offset` of every rank's buffer.  peers_dev: DEVICE array of `world` mapped buffer bases (own buffer at
 * index `rank`).  One kernel, one one-way NVLink latency (values are pushed, polling is local); every rank must call it for
 * the same slot the same number of times. */
int ttb_comm_allreduce(const double* partials, int num_chunks, int n, void* const* peers_dev, int world, int rank,
                       size_t slot_offset, double* out, void* stream);
/* The same exchange fused with the BatchNorm finalisation that follows it - one kernel per BatchNorm layer and
 * direction instead of ttb_comm_allreduce + ttb_bn_finalize / ttb_bn_bwd_finalize.  partials = this rank's
 * [num_chunks][2][C] doubles, count = GLOBAL elements per channel; the remaining arguments are those of the finalize
 * entry points. */
int ttb_comm_bn_finalize(const double* partials, int num_chunks, void* const* peers_dev, int world, int rank, size_t slot_offset,
                         int64_t count, int c, float eps, float momentum, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, float* mean, float* var_eps, float* sd, float* scale,
                         float* shift, void* stream);
int ttb_comm_bn_bwd_finalize(const double* partials, int num_chunks, void* const* peers_dev, int world, int rank,
                             size_t slot_offset, int64_t count, int c, const float* gamma, const float* var_eps,
                             const float* sd, float* dgamma, float* dbeta, float* coef, void* stream);

/* ---- max pool --------------------------------------------------------------------------------------------- */
/* y[N,P,Q,C] = max over window (padding acts as -inf); idx[N,P,Q,C] (uint8) = r*kw+s of the FIRST maximum in
 * row-major window order (two-stage nanargmax of the reference) */
int ttb_maxpool2d_fwd(const ttb_pool_desc* d, const float* x, float* y, uint8_t* idx, void* stream);
/* dx[N,H,W,C]: reference semantics — where several windows selected the same input element the LAST window in
 * (n, p, q) raster order wins (assignment, not accumulation; grad_nn.py:820).  accumulate != 0 switches to the
 * PyTorch behaviour (sum). */
int ttb_maxpool2d_bwd(const ttb_pool_desc* d, const float* dy, const uint8_t* idx, float* dx, int accumulate,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TORTTO_B200_H */
