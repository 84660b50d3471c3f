"""CPU: index arithmetic of the haloed-tile wgrad (pytortto_b200/csrc/conv_wgrad_halo.cu).

The kernel's geometry is restated here in numpy, step by step as the CUDA code does it - the host plan (box size, channel
slabs / filter rows per CTA, tap groups), the two TMA boxes of a step with their zero fill, the MMA operands as FLAT row
ranges of the halo tile (tap (r, s) = the tile read r*(BW+S-1) + s pixel rows later; the slabs of one MMA one pixel row
apart, so slab j is tap sg*TW + j and slabs past the filter width are garbage that must never reach dW), the epilogue's
lane -> (tap, channel) mapping - and checked against the oracle's weight gradient (reference _conv2d_backward_w,
autograd/grad_nn.py:646-656).  What this does NOT cover is the hardware side (swizzled TMA boxes, UMMA descriptors with
overlapping slabs): scripts/umma_lbo_overlap_probe.cu and the GPU tests (test_wgrad_haloed_tile_vs_oracle) do."""
import numpy as np
import pytest

from oracle import tortto_oracle as O

GARBAGE = 7e30  # what shared memory holds where no TMA box wrote


def halo_plan(n, c, h, w, k, r, s, pad_h, pad_w, bf16, kp_target=64):
    """conv_wgrad_halo.cu::halo_plan (the parts that decide the geometry)"""
    slab, tw, mma_rows = (64, 2, 16) if bf16 else (32, 4, 8)
    p, q = h + 2 * pad_h - r + 1, w + 2 * pad_w - s + 1
    assert 2 <= s <= 4 and c % slab == 0 and k % 32 == 0 and k % slab == 0 and k <= 256 and q >= mma_rows
    sg = (s + tw - 1) // tw
    cap = 512 // (sg * k)
    assert cap >= 1
    cslabs = c // slab
    nr = 3 if (r == 3 and cap >= 3) else 1
    ncs = 2 if (cap // nr >= 2 and cslabs % 2 == 0) else 1
    bw = min((q + mma_rows - 1) // mma_rows * mma_rows, 64)
    bh = min(max(kp_target // bw, 1), p)
    return dict(slab=slab, tw=tw, mma_rows=mma_rows, sg=sg, nr=nr, ncs=ncs, bw=bw, bh=bh, hw=bw + s - 1, p=p, q=q,
                rgroups=r // nr, cgroups=cslabs // ncs, boxes_w=(q + bw - 1) // bw, boxes_h=(p + bh - 1) // bh)


def halo_wgrad(x, dy, r_taps, s_taps, pad_h, pad_w, bf16, splits=3):
    """x: (N, H, W, C) NHWC, dy: (N, P, Q, K) NHWC -> dW (K, R, S, C), computed the way the kernel does"""
    n_img, h, w, c = x.shape
    k = dy.shape[3]
    g = halo_plan(n_img, c, h, w, k, r_taps, s_taps, pad_h, pad_w, bf16)
    slab, tw, rows, sg_n, nr, ncs = g["slab"], g["tw"], g["mma_rows"], g["sg"], g["nr"], g["ncs"]
    bw, bh, hw = g["bw"], g["bh"], g["hw"]
    boxes_total = n_img * g["boxes_w"] * g["boxes_h"]
    per_split = (boxes_total + splits - 1) // splits
    dw = np.zeros((k, r_taps, s_taps, c))
    for cta_y in range(g["rgroups"] * g["cgroups"]):
        cs0 = (cta_y // g["rgroups"]) * ncs
        r0 = (cta_y % g["rgroups"]) * nr
        for split in range((boxes_total + per_split - 1) // per_split):
            # accumulators of this CTA: (channel slab a, filter row r, tap group sg) -> [128 rows = (tap in group, channel)][K]
            acc = np.zeros((ncs, nr, sg_n, 128, k))
            for box in range(split * per_split, min((split + 1) * per_split, boxes_total)):
                bj = box % g["boxes_w"]
                t = box // g["boxes_w"]
                bi, n = t % g["boxes_h"], t // g["boxes_h"]
                # --- TMA box 1: x halo [ncs slabs][(bh + nr - 1) x hw pixel rows][slab channels], zero fill outside the image;
                #     laid out flat, followed by the dY tile (what the garbage taps of the last rows read into)
                hh = bh + nr - 1
                tile = np.full((ncs, hh * hw + 16, slab), GARBAGE)
                for a in range(ncs):
                    for i in range(hh):
                        for j in range(hw):
                            yy, xx = bi * bh - pad_h + r0 + i, bj * bw - pad_w + j
                            inside = 0 <= yy < h and 0 <= xx < w
                            tile[a, i * hw + j] = x[n, yy, xx, (cs0 + a) * slab:(cs0 + a + 1) * slab] if inside else 0.0
                # --- TMA box 2: dY [bh x bw pixel rows][K], zero fill past the output grid
                dyt = np.zeros((bh * bw, k))
                for i in range(bh):
                    for j in range(bw):
                        pp, qq = bi * bh + i, bj * bw + j
                        if pp < g["p"] and qq < g["q"]:
                            dyt[i * bw + j] = dy[n, pp, qq]
                # --- MMAs: K-group = `rows` consecutive pixels of one box row
                assert (bh * bw) % rows == 0 and bw % rows == 0
                for grp in range(bh * bw // rows):
                    px = grp * rows
                    gi, gj = divmod(px, bw)
                    b_op = dyt[px:px + rows]                                   # [rows][K]
                    for a in range(ncs):
                        for r in range(nr):
                            for sg in range(sg_n):
                                start = (gi + r) * hw + gj + sg * tw             # first pixel row of slab 0
                                for j in range(tw):                              # slab j starts ONE pixel row after slab j-1
                                    a_op = tile[a, start + j:start + j + rows]   # [rows][slab channels]
                                    # garbage (never-written rows) may only feed taps that the epilogue drops
                                    if sg * tw + j < s_taps:
                                        assert np.abs(a_op).max() < 1e30
                                    acc[a, r, sg, j * slab:(j + 1) * slab] += np.where(np.abs(a_op) < 1e30, a_op, 0.0).T @ b_op
            # --- epilogue: lane row = (tap in group, channel in slab); partial buffers summed over the splits
            for a in range(ncs):
                for r in range(nr):
                    for sg in range(sg_n):
                        for row in range(128):
                            s_tap = sg * tw + row // slab
                            if s_tap < s_taps:
                                dw[:, r0 + r, s_tap, (cs0 + a) * slab + row % slab] += acc[a, r, sg, row]
    return dw


# (N, C, H, W, K, R, S, pad_h, pad_w, bf16)
CASES = [
    (2, 32, 9, 10, 32, 3, 3, 1, 1, False),     # one slab, three rows per CTA
    (2, 64, 8, 8, 64, 3, 3, 1, 1, False),      # two slabs per CTA
    (1, 64, 12, 20, 32, 3, 3, 1, 1, False),    # ragged box columns (Q = 20 -> BW = 24) and rows
    (1, 32, 9, 24, 64, 3, 2, 0, 0, False),     # no padding, 2-wide filter
    (1, 32, 6, 70, 32, 3, 3, 1, 1, False),     # rows wider than a box
    (1, 64, 9, 11, 128, 1, 4, 0, 2, False),    # 4-wide filter, one filter row
    (1, 32, 10, 12, 256, 5, 3, 2, 1, False),   # one filter row per CTA (R = 5), K = 256: one accumulator per CTA... per (slab,row)
    (1, 64, 8, 16, 64, 3, 3, 1, 1, True),      # bf16: 64-channel slabs, two taps per MMA, two tap groups
    (1, 128, 5, 18, 128, 3, 3, 1, 1, True),    # bf16, ragged columns (Q = 18 -> BW = 32)
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "n{}_c{}_{}x{}_k{}_f{}x{}_p{}{}_{}".format(*c[:9], "bf16" if c[9] else "tf32"))
def test_haloed_wgrad_geometry_matches_the_oracle(case):
    n, c, h, w, k, r, s, ph, pw, bf16 = case
    rng = np.random.default_rng(abs(hash(case)) % (2 ** 31))
    x = rng.standard_normal((n, c, h, w)).astype(np.float32)
    wt_shape = (k, c, r, s)
    p, q = h + 2 * ph - r + 1, w + 2 * pw - s + 1
    dy = rng.standard_normal((n, k, p, q)).astype(np.float32)
    ref = O.conv2d_backward_weight(x, dy, wt_shape, (1, 1), (ph, pw), (1, 1))          # (K, C, R, S)
    got = halo_wgrad(x.transpose(0, 2, 3, 1).astype(np.float64), dy.transpose(0, 2, 3, 1).astype(np.float64), r, s, ph, pw, bf16)
    got = got.transpose(0, 3, 1, 2)                                                     # (K, R, S, C) -> (K, C, R, S)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 2e-5, err


def test_plan_keeps_the_accumulators_inside_tensor_memory():
    for (n, c, h, w, k, r, s, ph, pw, bf16) in CASES:
        g = halo_plan(n, c, h, w, k, r, s, ph, pw, bf16)
        assert g["ncs"] * g["nr"] * g["sg"] * k <= 512
        assert g["bw"] % g["mma_rows"] == 0 and g["bh"] >= 1
        assert (g["bh"] * g["bw"]) % g["mma_rows"] == 0
