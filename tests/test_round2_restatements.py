"""CPU restatements (numpy) of two round-2 device algorithms, held to the oracle: what the kernels compute is checked
here without a GPU; that the kernels compute it is checked by the `-m gpu` tests (test_gpu_ops.py::test_stem_conv2d_*,
test_gpu_epilogue.py::test_dgrad_emits_the_batchnorm_backward_sums).

1. Row-packed staging of tall-filter stems (csrc/conv_api.cu::plan_tensor, pack_taps_kernel with the `pk` descriptor): the
   S*C taps of a filter ROW of pixel (n, h, q) form one staged row of round_up(S*C, 32) columns; the layer is then an R x 1
   convolution (stride / padding / dilation in H unchanged, none in W) over that [N][H][Q][kp] tensor with the weights
   viewed as [K][R][S*C -> kp].
2. The sums a dgrad epilogue emits for the BatchNorm backward that reads its output (conv_epilogue.cuh, statistics mode 2):
   sum(g), sum(g * (x - mean)) with g = dx masked by fmaf(x - mean, scale, shift) > 0; ttb_bn_bwd_finalize turns them into
   dbeta, dgamma and the three coefficients of the dx pass (bn_finalize.cuh::BnBwdFinalize)."""
import numpy as np
import pytest

from oracle import tortto_oracle as O


def _row_pack(x, s_taps, stride_w, pad_w, dil_w, q_out, blk=32):
    """x (N, C, H, W) -> staged (N, kp, H, Q): column (s, c) of row (n, h, q) = x zero-padded at (h, q*sw + s*dw - pw)"""
    n, c, h, w = x.shape
    kp = (s_taps * c + blk - 1) // blk * blk
    out = np.zeros((n, kp, h, q_out), x.dtype)
    for s in range(s_taps):
        for q in range(q_out):
            col = q * stride_w + s * dil_w - pad_w
            if 0 <= col < w:
                out[:, s * c:(s + 1) * c, :, q] = x[:, :, :, col]
    return out


@pytest.mark.parametrize("case", [(2, 3, 20, 18, 8, 7, 7, 2, 3, 1), (2, 3, 21, 40, 8, 7, 7, 2, 3, 1), (1, 4, 30, 26, 6, 7, 5, 2, 3, 1),
                                  (2, 2, 15, 15, 4, 5, 3, 1, 2, 2), (1, 3, 12, 9, 4, 3, 3, 1, 1, 1)],
                         ids=["stem7x7s2", "stem_ragged", "c4_7x5", "dilated5x3", "cifar3x3"])
def test_row_packed_stem_equals_the_convolution(case):
    n, c, h, w, k, r, s, stride, pad, dil = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((n, c, h, w)).astype(np.float32)
    wt = rng.standard_normal((k, c, r, s)).astype(np.float32)
    y = O.conv2d_forward(x, wt, None, stride, pad, dil)
    p_out, q_out = y.shape[2], y.shape[3]
    staged = _row_pack(x, s, stride, pad, dil, q_out)
    kp = staged.shape[1]
    # weights (K, C, R, S) -> channels-last [K][R][S][C] -> [K][R][S*C] zero-extended to kp -> (K, kp, R, 1)
    w_rows = np.zeros((k, r, kp), np.float32)
    w_rows[:, :, :s * c] = wt.transpose(0, 2, 3, 1).reshape(k, r, s * c)
    w_packed = w_rows.transpose(0, 2, 1)[:, :, :, None]
    y2 = O.conv2d_forward(staged, w_packed, None, (stride, 1), (pad, 0), (dil, 1))
    assert y2.shape == y.shape == (n, k, p_out, q_out)
    np.testing.assert_allclose(y2, y, rtol=1e-5, atol=1e-5)
    # wgrad over the staged tensor, cropped back to S*C columns and un-viewed, is the layer's weight gradient
    dy = rng.standard_normal(y.shape).astype(np.float32)
    dw = O.conv2d_backward_weight(x, dy, wt.shape, stride, pad, dil)
    dw2 = O.conv2d_backward_weight(staged, dy, w_packed.shape, (stride, 1), (pad, 0), (dil, 1))  # (K, kp, R, 1)
    dw2 = dw2[:, :s * c, :, 0].transpose(0, 2, 1).reshape(k, r, s, c).transpose(0, 3, 1, 2)
    np.testing.assert_allclose(dw2, dw, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("relu", [True, False])
def test_dgrad_epilogue_sums_are_the_batchnorm_backward_reductions(relu):
    rng = np.random.default_rng(3)
    n, c, h, w = 6, 8, 5, 7
    x = (rng.standard_normal((n, c, h, w)) * 1.5 + 0.3).astype(np.float32)
    gamma = np.linspace(0.5, 1.5, c, dtype=np.float32)
    beta = np.linspace(-0.4, 0.4, c, dtype=np.float32)
    y, _, _, saved = O.batch_norm_forward(x, gamma, beta, None, None, True, 0.1, 1e-5)
    mean, var_eps, sd = saved
    a = O.relu_forward(y) if relu else y
    dx_conv = rng.standard_normal(a.shape).astype(np.float32)          # what the dgrad of the conv behind it stores
    # reference order of operations: Relu.backward, then BatchNorm.backward
    g_ref = O.relu_backward(dx_conv, a) if relu else dx_conv
    dx_ref, dgamma_ref, dbeta_ref = O.batch_norm_backward(g_ref, x, gamma, saved)
    # the epilogue's version: mask recomputed from x with forward's scale / shift, two sums per channel
    scale = (gamma.reshape(mean.shape) / sd).astype(np.float32)
    cx = x - mean
    mask = (cx * scale + beta.reshape(mean.shape)) > 0 if relu else np.ones_like(x, bool)
    g = np.where(mask, dx_conv, 0).astype(np.float32)
    s0 = g.sum((0, 2, 3), dtype=np.float64)
    s1 = (g * cx).sum((0, 2, 3), dtype=np.float64)
    np.testing.assert_array_equal(g, g_ref)                             # same decisions as the ReLU took on its output
    # BnBwdFinalize: dbeta = s0, dgamma = s1 / sd, dx = c1 * (g - c2 - (x - mean) * c3)
    count = n * h * w
    np.testing.assert_allclose(s0, dbeta_ref, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(s1 / sd.reshape(-1), dgamma_ref, rtol=2e-5, atol=2e-5)
    c1 = (gamma / sd.reshape(-1)).reshape(mean.shape)
    c2 = (s0 / count).astype(np.float32).reshape(mean.shape)
    c3 = (s1 / (count * var_eps.reshape(-1).astype(np.float64))).astype(np.float32).reshape(mean.shape)
    dx = c1 * (g - c2 - cx * c3)
    np.testing.assert_allclose(dx, dx_ref, rtol=1e-4, atol=2e-5)
