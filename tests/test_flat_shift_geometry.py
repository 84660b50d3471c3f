"""CPU: index arithmetic of the flat-shift halo-tile prototype (pytortto_b200/csrc/conv_flat.cu).

The kernel itself could not be run when it was written (the round's GPU budget was spent), so its geometry is restated
here line by line in numpy - padded-flat tiles of 128 outputs, one strip of padded image rows per tile, every filter tap
a shift by `lead + r*Wp + s` rows, outputs at padding positions dropped, dgrad as the same correlation over dY with
flipped taps and padding R-1-pad - and checked against the oracle's convolution.  What this does NOT cover is the
hardware side (TMA boxes, swizzle, UMMA descriptors): that is what `TTB_FLAT=1 python scripts/flat_check.py` is for."""
import numpy as np
import pytest

from oracle import tortto_oracle as O

TILE = 128


def rows_max(wp, r, s):  # conv_flat.cu::flat_rows_max_of
    return (wp - 1 + TILE + s - 1 + wp - 1) // wp + (r - 1)


def flat_correlate(inp, wmat, koff, r_taps, s_taps, pad_h, pad_w, p_out, q_out, k_out):
    """inp: (N, H, W, C) NHWC; wmat: (k_out, cols) with tap t's C-slice at column koff[t]; returns (N, p_out, q_out, k_out).
    Mirrors igemm_flat_kernel: strip producer, MMA row offsets and epilogue mapping."""
    n_img, h, w, c = inp.shape
    hp, wp = h + 2 * pad_h, w + 2 * pad_w
    m_flat = n_img * hp * wp
    rmax = rows_max(wp, r_taps, s_taps)
    out = np.full((n_img, p_out, q_out, k_out), np.nan, dtype=np.float64)
    rows_total = n_img * hp
    for f0 in range(0, m_flat, TILE):
        # --- strip producer: padded rows row0..row1 (clipped at the end of the last image), one box per padded row
        row0 = f0 // wp
        last = f0 + TILE - 1 + (r_taps - 1) * wp + (s_taps - 1)
        row1 = min(last // wp, rows_total - 1)
        nrows = row1 - row0 + 1
        assert 1 <= nrows <= rmax
        strip = np.full((rmax * wp, c), 7e30)  # rows never loaded hold garbage (must only reach dropped outputs)
        for j in range(nrows):
            row = row0 + j
            n, hpad = divmod(row, hp)
            hh = hpad - pad_h
            box = np.zeros((wp, c))  # TMA box {32 ch, Wp pixels from w = -pad_w, 1 row}: out of bounds = zero fill
            if 0 <= hh < h:
                box[pad_w:pad_w + w] = inp[n, hh]
            strip[j * wp:(j + 1) * wp] = box
        # --- MMA: every tap reads 128 consecutive strip rows starting at lead + r*Wp + s
        lead = f0 - row0 * wp
        acc = np.zeros((TILE, k_out))
        for tap in range(r_taps * s_taps):
            r, s = divmod(tap, s_taps)
            off = lead + r * wp + s
            assert off + TILE <= rmax * wp
            a = strip[off:off + TILE]
            acc += a @ wmat[:, koff[tap]:koff[tap] + c].T
        # --- epilogue: flat index -> (n, p, q); padding positions are dropped
        for i in range(TILE):
            f = f0 + i
            if f >= m_flat:
                break
            q = f % wp
            t = f // wp
            p = t % hp
            n = t // hp
            if q < q_out and p < p_out:
                out[n, p, q] = acc[i]
    assert np.isfinite(out).all(), "a valid output was never written, or read a strip row that was not loaded"
    return out


CASES = [  # n, c, h, w, k, r, s, pad_h, pad_w
    (2, 4, 8, 8, 3, 3, 3, 1, 1),
    (3, 5, 12, 10, 4, 3, 3, 1, 1),     # Wp = 12, ragged last tile
    (2, 3, 32, 32, 2, 3, 3, 1, 1),     # Wp = 34: the layer-1 geometry
    (5, 2, 4, 4, 3, 3, 3, 1, 1),       # tiny images: a tile spans several images
    (2, 3, 10, 9, 2, 5, 5, 2, 2),
    (2, 3, 9, 11, 2, 3, 3, 0, 0),      # no padding
    (1, 2, 7, 20, 2, 1, 3, 0, 1),      # 1 x 3 filter
    (2, 2, 6, 6, 2, 3, 3, 2, 2),       # "full" padding
]


@pytest.mark.parametrize("case", CASES, ids=[f"n{c[0]}_{c[2]}x{c[3]}_f{c[5]}x{c[6]}_p{c[7]}{c[8]}" for c in CASES])
def test_flat_shift_fprop_and_dgrad_match_oracle(case):
    n, c, h, w, k, r, s, ph, pw = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((n, c, h, w)).astype(np.float32)
    wt = rng.standard_normal((k, c, r, s)).astype(np.float32)
    y_ref = O.conv2d_forward(x, wt, None, (1, 1), (ph, pw), (1, 1))
    p_out, q_out = y_ref.shape[2], y_ref.shape[3]
    # fprop: weights [K][R][S][C], tap t at column t*C  (conv_flat.cu::flat_fprop)
    wmat = wt.transpose(0, 2, 3, 1).reshape(k, r * s * c).astype(np.float64)
    koff = [t * c for t in range(r * s)]
    y = flat_correlate(x.transpose(0, 2, 3, 1).astype(np.float64), wmat, koff, r, s, ph, pw, p_out, q_out, k)
    np.testing.assert_allclose(y.transpose(0, 3, 1, 2), y_ref, rtol=1e-4, atol=1e-4)
    # dgrad: correlation over dY with padding R-1-pad and flipped taps on the re-ordered weights [C][R][S][K]
    # (conv_flat.cu::flat_dgrad); needs pad <= R-1
    if ph <= r - 1 and pw <= s - 1:
        dy = rng.standard_normal(y_ref.shape).astype(np.float32)
        dx_ref = O.conv2d_backward(x, wt, dy, (1, 1), (ph, pw), (1, 1))[0]
        packed = wt.transpose(1, 2, 3, 0).reshape(c, r * s * k).astype(np.float64)  # [C][R][S][K]
        koff_d = [((r - 1 - rr) * s + (s - 1 - ss)) * k for rr in range(r) for ss in range(s)]
        dx = flat_correlate(dy.transpose(0, 2, 3, 1).astype(np.float64), packed, koff_d, r, s, r - 1 - ph, s - 1 - pw, h, w, c)
        np.testing.assert_allclose(dx.transpose(0, 3, 1, 2), dx_ref, rtol=1e-4, atol=1e-4)
