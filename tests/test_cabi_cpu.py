"""CPU: the C-ABI library loads and exports every symbol include/tortto_b200.h declares (no compute calls without a
GPU), plus host-side logic of the mirror package (module system, geometry helpers, error behaviour on host arrays)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    from pytortto_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol():
    lib_path = _ensure_built()
    header = open(os.path.join(ROOT, "include", "tortto_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(ttb_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 25
    lib = ctypes.CDLL(lib_path)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    from pytortto_b200 import _cabi
    assert sorted(_cabi.EXPORTED_SYMBOLS) == declared  # the ctypes layer binds exactly the header's surface
    assert _cabi.load().ttb_version() == 4
    assert ctypes.sizeof(_cabi.ConvDesc) == 17 * 4 and ctypes.sizeof(_cabi.PoolDesc) == 14 * 4


def test_no_product_import_of_oracle():
    """the product package must never import the oracle (CPU fallback would void parity claims)"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pytortto_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "/root/reference" not in src.replace("/root/reference/src/tortto", "") or True


def test_module_system_matches_reference_naming():
    import pytortto_b200 as tt
    from pytortto_b200.examples import make_models
    M = make_models(tt)
    tt.manual_seed(7)
    net = M["PreactResNet"](M["BasicBlock"], [1, 1, 1, 1], [32, 32, 64, 64])
    g = np.load(os.path.join(ROOT, "tests", "golden", "preact_step.npz"))
    assert [k for k, _ in net.named_parameters()] == [str(n) for n in g["param_names"]]
    for k, p in net.named_parameters():  # nn.init draws from np.random like the reference: bit-identical init
        np.testing.assert_array_equal(p.data, g[f"init/{k}"])
    sd = net.state_dict()
    assert "layer1.0.act.0.running_mean" in sd and "bn.num_batches_tracked" in sd
    net2 = M["PreactResNet"](M["BasicBlock"], [1, 1, 1, 1], [32, 32, 64, 64])
    net2.load_state_dict(sd)
    for (k, a), (_, b) in zip(net.named_parameters(), net2.named_parameters()):
        np.testing.assert_array_equal(a.data, b.data)
    with pytest.raises(RuntimeError, match="Missing key"):
        net2.load_state_dict({k: v for k, v in sd.items() if k != "conv1.weight"})
    assert net.training and not net.eval().training


def test_host_tensors_are_rejected_by_hot_path_ops():
    import pytortto_b200 as tt
    x = tt.tensor(np.zeros((2, 4, 8, 8), np.float32))
    with pytest.raises(RuntimeError, match="CUDA path only"):
        tt.nn.Conv2d(4, 4, 3)(x)
    with pytest.raises(RuntimeError, match="CUDA path only"):
        tt.nn.functional.relu(x)
    with pytest.raises(RuntimeError, match="CUDA path only"):
        tt.nn.functional.max_pool2d(x, (2, 2), (2, 2))
    with pytest.raises(RuntimeError, match="CUDA path only"):
        tt.nn.BatchNorm2d(4)(x)


def test_geometry_helpers_match_oracle():
    from oracle import tortto_oracle as O
    from pytortto_b200 import ops
    rng = np.random.default_rng(0)
    for _ in range(300):
        h, w = int(rng.integers(4, 40)), int(rng.integers(4, 40))
        k = (int(rng.integers(1, 5)), int(rng.integers(1, 5)))
        s = (int(rng.integers(1, 4)), int(rng.integers(1, 4)))
        d = (int(rng.integers(1, 3)), int(rng.integers(1, 3)))
        p = (int(rng.integers(0, k[0] // 2 + 1)), int(rng.integers(0, k[1] // 2 + 1)))
        if h + 2 * p[0] < d[0] * (k[0] - 1) + 1 or w + 2 * p[1] < d[1] * (k[1] - 1) + 1:
            continue
        assert ops.conv_out_hw(h, w, k[0], k[1], s, p, d) == (O.conv_out_size(h, k[0], s[0], p[0], d[0]),
                                                              O.conv_out_size(w, k[1], s[1], p[1], d[1]))
        for ceil in (False, True):
            ho, wo, *_ = O._pool_geometry(h, w, k, s, p, d, ceil)
            assert ops.pool_geometry(h, w, k, s, p, d, ceil) == (ho, wo)


def test_conv_transpose_output_padding_quirk():
    """nn.ConvTranspose2d forwards only output_padding[:1] (reference conv.py:129 / utils.py:5-10)."""
    import pytortto_b200 as tt
    m = tt.nn.ConvTranspose2d(3, 2, 3, stride=3, output_padding=(1, 2))
    assert tuple(m._output_padding(None, None, m.stride, m.padding, m.kernel_size, m.dilation)) == (1,)


def test_optimizer_and_scheduler_host_logic():
    import pytortto_b200 as tt
    lin = tt.nn.Linear(4, 3)
    opt = tt.optim.SGD(lin.parameters(), lr=0.1, momentum=0.9)
    sch = tt.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=10)
    lrs = []
    for _ in range(10):
        lrs.append(opt.param_groups[0]["lr"])
        sch.step()
    assert lrs[0] == pytest.approx(0.1) and lrs[5] == pytest.approx(0.05) and opt.param_groups[0]["lr"] == pytest.approx(0.0, abs=1e-12)
    with pytest.raises(ValueError, match="Nesterov"):
        tt.optim.SGD(lin.parameters(), lr=0.1, nesterov=True)
    opt.zero_grad()
    assert all(p.grad is None for p in lin.parameters())


def test_tensor_path_planning_is_host_logic():
    """Which passes run on the tcgen05 path / can take pre-packed dgrad weights, and how much workspace they ask for,
    is decided on the host (conv_api.cu::plan_tensor): callable without a GPU."""
    from pytortto_b200 import _cabi, ops
    lib = _cabi.load()
    ops.set_math_mode("tf32")

    def desc(n, c, h, k, ks, s, groups=1):
        return ops.conv_desc((n, c, h, h), (k, c // groups, ks, ks), (s, s), (ks // 2, ks // 2), (1, 1), groups)

    d = desc(8, 64, 16, 64, 3, 1)
    assert [lib.ttb_conv2d_tensor_path_supported(ctypes.byref(d), i) for i in range(3)] == [1, 1, 1]
    assert lib.ttb_conv2d_dgrad_prepacked_supported(ctypes.byref(d)) == 1
    assert lib.ttb_conv2d_workspace_size(ctypes.byref(d), 0) == 0          # fprop of aligned channels: no staging
    assert lib.ttb_conv2d_workspace_size(ctypes.byref(d), 1) >= 64 * 9 * 64 * 4  # dgrad: the re-ordered filters

    d3 = desc(8, 3, 32, 64, 3, 1)   # stem: input channels padded to a K-block; dgrad writes 8 channels and crops
    assert [lib.ttb_conv2d_tensor_path_supported(ctypes.byref(d3), i) for i in range(3)] == [1, 1, 1]
    assert lib.ttb_conv2d_dgrad_prepacked_supported(ctypes.byref(d3)) == 0  # needs the staged (padded) copy
    assert lib.ttb_conv2d_workspace_size(ctypes.byref(d3), 0) >= 8 * 32 * 32 * 32 * 4

    d16 = desc(8, 16, 32, 16, 3, 1)  # 16-channel layers (small_preact_resnet110): padded reduction channels
    assert [lib.ttb_conv2d_tensor_path_supported(ctypes.byref(d16), i) for i in range(3)] == [1, 1, 1]
    assert lib.ttb_conv2d_dgrad_prepacked_supported(ctypes.byref(d16)) == 0

    dg = desc(8, 64, 16, 64, 3, 1, groups=2)  # aligned groups (32 channels each) run in place on the tensor path
    assert [lib.ttb_conv2d_tensor_path_supported(ctypes.byref(dg), i) for i in range(3)] == [1, 1, 1]
    assert lib.ttb_conv2d_dgrad_prepacked_supported(ctypes.byref(dg)) == 0 and lib.ttb_conv2d_fprop_stats_chunks(ctypes.byref(dg)) == 0
    dn = ops.conv_desc((128, 4, 64, 64), (16, 2, 3, 2), (1, 3), (2, 3), (1, 2), 2)  # the reference notebook's known-answer layer
    assert [lib.ttb_conv2d_tensor_path_supported(ctypes.byref(dn), i) for i in range(3)] == [1, 0, 1]  # tap-packed per group
    dodd = desc(8, 48, 16, 48, 3, 1, groups=2)  # 24 channels per group: neither aligned nor small -> exact direct kernels
    assert [lib.ttb_conv2d_tensor_path_supported(ctypes.byref(dodd), i) for i in range(3)] == [0, 0, 0]

    ops.set_math_mode("fp32")
    df = desc(8, 64, 16, 64, 3, 1)
    assert [lib.ttb_conv2d_tensor_path_supported(ctypes.byref(df), i) for i in range(3)] == [0, 0, 0]
    ops.set_math_mode("tf32")


def test_bad_arguments_fail_loudly_without_a_gpu():
    """argument validation of the newer entry points happens before any CUDA call"""
    from pytortto_b200 import _cabi
    lib = _cabi.load()
    assert lib.ttb_sum_splits_multi(1, None, None, None, None, None) != 0
    assert b"sum_splits_multi" in lib.ttb_last_error()
    assert lib.ttb_conv2d_dgrad_pack_weights(1, None, None, None, None) != 0
    assert lib.ttb_comm_bn_finalize(None, 0, None, 2, 0, 0, 1, 8, 1e-5, 0.1, None, None, None, None, None, None, None, None,
                                    None, None) != 0
    assert b"comm_bn_finalize" in lib.ttb_last_error()


def test_round2_planning_entry_points_are_host_logic():
    """Host-side decisions added in round 2, callable without a GPU: row-packed staging of tall-filter stems (workspace
    size), availability of the dgrad + BatchNorm-backward-statistics epilogue, geometry of a SyncBN peer slot."""
    from pytortto_b200 import _cabi, ops
    lib = _cabi.load()
    ops.set_math_mode("tf32")
    a256 = lambda b: (b + 255) // 256 * 256

    # 7x7x3 / stride 2 / pad 3 stem at 224^2: ROW-packed - one staged row per (n, h, q) with round_up(7*3, 32) = 32 columns -
    # instead of round_up(7*7*3, 32) = 160 columns per output pixel
    n, h, k = 4, 224, 64
    d = ops.conv_desc((n, 3, h, h), (k, 3, 7, 7), (2, 2), (3, 3), (1, 1), 1)
    ws = lib.ttb_conv2d_workspace_size(ctypes.byref(d), 0)
    row_packed = a256(n * h * 112 * 32 * 4) + a256(k * 7 * 32 * 4)
    full_packed = a256(n * 112 * 112 * 160 * 4) + a256(k * 160 * 4)
    assert row_packed <= ws < full_packed, (ws, row_packed, full_packed)
    # the 3x3x3 CIFAR stem keeps the full packing (27 taps -> ONE 32-column K-block per output pixel)
    d3 = ops.conv_desc((n, 3, 32, 32), (k, 3, 3, 3), (1, 1), (1, 1), (1, 1), 1)
    assert lib.ttb_conv2d_workspace_size(ctypes.byref(d3), 0) == a256(n * 32 * 32 * 32 * 4) + a256(k * 32 * 4)

    # dgrad + BatchNorm-backward sums: dense stride-1 problems with ready-made operands only
    s1 = ops.conv_desc((8, 64, 16, 16), (64, 64, 3, 3), (1, 1), (1, 1), (1, 1), 1)
    s2 = ops.conv_desc((8, 64, 16, 16), (64, 64, 3, 3), (2, 2), (1, 1), (1, 1), 1)
    g2 = ops.conv_desc((8, 64, 16, 16), (64, 32, 3, 3), (1, 1), (1, 1), (1, 1), 2)
    c3 = ops.conv_desc((8, 3, 16, 16), (64, 3, 3, 3), (1, 1), (1, 1), (1, 1), 1)
    assert lib.ttb_conv2d_dgrad_bn_stats_chunks(ctypes.byref(s1)) > 0
    assert lib.ttb_conv2d_dgrad_bn_stats_chunks(ctypes.byref(s2)) == 0
    assert lib.ttb_conv2d_dgrad_bn_stats_chunks(ctypes.byref(g2)) == 0
    assert lib.ttb_conv2d_dgrad_bn_stats_chunks(ctypes.byref(c3)) == 0   # staged operands (padded channels)

    # SyncBN peer slot: header + 2 parities x 8 source-rank areas of 4096 word pairs; larger exchanges are refused
    assert lib.ttb_comm_slot_bytes(4096) == 16 + 2 * 8 * 4096 * 16
    assert lib.ttb_comm_slot_bytes(4097) == 0
