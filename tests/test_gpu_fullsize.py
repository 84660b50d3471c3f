"""GPU: the HEADLINE configuration (BASELINE.json configs[1]: preact_resnet18, batch 256, 3x32x32) held to the numpy
oracle (oracle/resnet_oracle.StepOracle = the reference's algorithm, pinned by tests/test_oracle_golden.py), not to
itself.

What "equal" can mean at this size (measured, scripts/relu_flip_analysis.py, profiles/r2_relu_flip_analysis.txt): the
reference's OWN fp32 arithmetic is 2e-3 (rel-L2) / up to 4e-2 (max-abs / tensor max) away from a float64 evaluation of
the same formulas on every gradient upstream of the last block.  The cause is not accumulated rounding but single ReLU
decisions: one pre-activation within 1e-7 of zero lands on the other side, its whole gradient element appears or
disappears, and that one element is ~2e-3 of the L2 norm of the heavy-tailed gradient field.  TF32 / bf16 operand
rounding moves ~1e-3 / ~8e-3 of all pre-activations across zero, and the gradient error grows like the square root of
the number of flips (~0.1 / ~0.3) - in the reference's algorithm with rounded operands exactly as on the GPU.  So:

  * exact-fp32 mode: loss, log-probs and BatchNorm running statistics (forward: no decisions differ) at 1e-5 / 1e-4 /
    2e-5; EVERY parameter gradient within 3 x the largest distance the reference's own fp32 result shows from the
    float64 truth on any tensor of the step (which decisions flip differs between two summation orders, so the
    comparison is per step, not per tensor) - the implementation is held to the reference's own accuracy;
  * TF32 / bf16 modes: loss and log-probs at the north-star tolerances (2e-3 / 1e-2); per-tensor gradient bounds at
    3x / 2x the values measured on B200 (listing printed by the test, committed as profiles/r2_fullsize_parity.txt).
    What catches a wrong tap or a mis-padded K-block in ONE layer is the layer-wise replay of all 20 convolutions of
    this network at batch 256 on recorded activations (tests/test_gpu_models.py::test_preact_resnet18_layerwise_*),
    per op at 2e-3 / 1e-2 where no ReLU decision is involved.

The two oracle steps (fp32 and float64) take ~1.5 min on the box's host cores and are computed once per module."""
import numpy as np
import pytest

from gpu_util import report, require_gpu

pytestmark = pytest.mark.gpu

LAYERS, CHANNELS, BATCH = [2, 2, 2, 2], [64, 128, 256, 512], 256


@pytest.fixture(autouse=True)
def _gpu():
    require_gpu()


def _build(tt):
    from pytortto_b200.examples import make_models
    M = make_models(tt)
    return M["PreactResNet"](M["BasicBlock"], LAYERS, CHANNELS)


@pytest.fixture(scope="module")
def oracle_step():
    import pytortto_b200 as tt
    from oracle.resnet_oracle import StepOracle
    rng = np.random.default_rng(0)
    x = rng.standard_normal((BATCH, 3, 32, 32)).astype(np.float32)
    lab = rng.integers(0, 10, BATCH).astype(np.int64)
    tt.manual_seed(0)
    net = _build(tt)
    params0 = {k: np.array(p.data, copy=True) for k, p in net.named_parameters()}
    orc = StepOracle(LAYERS, CHANNELS, {k: v.copy() for k, v in params0.items()})
    loss, logp, grads = orc.forward_backward(x, lab)
    o64 = StepOracle(LAYERS, CHANNELS, {k: v.astype(np.float64) for k, v in params0.items()}, dtype=np.float64)
    loss64, logp64, grads64 = o64.forward_backward(x.astype(np.float64), lab)
    return dict(x=x, lab=lab, params0=params0, loss=float(loss), logp=logp, grads=grads, buffers=orc.buffers,
                loss64=float(loss64), logp64=logp64, grads64=grads64)


def _gpu_step(mode, o):
    import pytortto_b200 as tt
    tt.set_math_mode(mode)
    tt.manual_seed(0)
    net = _build(tt)
    for k, p in net.named_parameters():
        np.testing.assert_array_equal(p.data, o["params0"][k])
    net.cuda().train()
    logp = net(tt.tensor(o["x"]).cuda())
    loss = tt.nn.NLLLoss()(logp, tt.tensor(o["lab"], dtype=np.int64).cuda())
    loss.backward()
    grads = {k: p.grad.get() for k, p in net.named_parameters()}
    return net, loss.item(), logp.data.get(), grads


def _rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_full_size_step_fp32_vs_oracle(oracle_step):
    o = oracle_step
    net, loss, logp, grads = _gpu_step("fp32", o)
    assert abs(loss - o["loss"]) <= 1e-5 * max(1.0, abs(o["loss"])), (loss, o["loss"])
    msg, rel = report("logp", logp, o["logp"])
    print(msg)
    assert rel <= 1e-4, msg
    # WHICH ~20 of the 1.4e8 ReLU decisions land on the other side is a property of each arithmetic's summation order, so
    # the two fp32 results deviate from the truth on different tensors: each of our tensors is held to 3 x the LARGEST
    # deviation the reference's own fp32 arithmetic shows on any tensor of this step (both metrics)
    ours, ref = {}, {}
    for k, g in grads.items():
        truth = o["grads64"][k]
        ours[k] = (report(k, g, truth)[1], _rel_l2(g, truth))
        ref[k] = (report(k, o["grads"][k], truth)[1], _rel_l2(o["grads"][k], truth))
        print(f"  [fp32] {k}: vs float64 truth (max-abs/max, rel-L2) - ours {ours[k][0]:.3e} {ours[k][1]:.3e}, the reference's fp32 "
              f"algorithm {ref[k][0]:.3e} {ref[k][1]:.3e}")
    ref_max, ref_l2 = max(v[0] for v in ref.values()), max(v[1] for v in ref.values())
    our_max, our_l2 = max(v[0] for v in ours.values()), max(v[1] for v in ours.values())
    print(f"[fp32 vs float64 oracle, batch {BATCH}] loss {loss:.6f} (oracle fp32 {o['loss']:.6f}, float64 {o['loss64']:.6f}); worst "
          f"gradient error over {len(grads)} tensors (max-abs/max, rel-L2): ours {our_max:.3e} {our_l2:.3e}, reference fp32 "
          f"algorithm {ref_max:.3e} {ref_l2:.3e}")
    bad = [(k, v) for k, v in ours.items() if v[0] > 3.0 * max(ref_max, 5e-4) or v[1] > 3.0 * max(ref_l2, 5e-4)]
    assert not bad, bad
    sd = net.state_dict()
    worst_rs = 0.0
    for k, v in o["buffers"].items():
        if k.endswith("num_batches_tracked"):
            assert float(sd[k]) == float(v)
            continue
        msg, rel = report(f"buffer {k}", sd[k], v)
        worst_rs = max(worst_rs, rel)
        assert rel <= 2e-5, msg
    print(f"[fp32 vs oracle] worst BatchNorm running-statistic rel-err {worst_rs:.3e} (gate 2e-5)")


# Per-tensor gradient gates of the tensor-core modes = 3 x (tf32) / 2 x (bf16) the worst value measured on B200 for this
# exact step (measured: tf32 0.119 / 0.112, bf16 0.314 / 0.266 as max-abs / tensor max and rel-L2;
# profiles/r2_fullsize_parity.txt holds the per-tensor listing the test prints).  The error is ReLU-decision noise (module
# docstring); single-layer correctness is held at 2e-3 / 1e-2 by the layer-wise replay in tests/test_gpu_models.py.
GATES = {"tf32": dict(loss=2e-3, maxabs=0.36, l2=0.34), "bf16": dict(loss=1e-2, maxabs=0.63, l2=0.55)}


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
def test_full_size_step_tensor_modes_vs_oracle(oracle_step, mode):
    o = oracle_step
    gate = GATES[mode]
    _, loss, logp, grads = _gpu_step(mode, o)
    assert abs(loss - o["loss64"]) <= gate["loss"] * max(1.0, abs(o["loss64"])), (loss, o["loss64"])
    _, rel_logp = report("logp", logp, o["logp64"])
    worst, worst_l2, names = 0.0, 0.0, ("", "")
    for k, g in grads.items():
        assert np.isfinite(g).all(), k
        _, rel = report(k, g, o["grads64"][k])
        l2 = _rel_l2(g, o["grads64"][k])
        print(f"  [{mode}] {k}: max-abs/max {rel:.3e}  rel-L2 {l2:.3e}")
        if rel > worst:
            worst, names = rel, (k, names[1])
        if l2 > worst_l2:
            worst_l2, names = l2, (names[0], k)
    print(f"[{mode} vs oracle, batch {BATCH}] loss {loss:.6f} (oracle {o['loss']:.6f}, rel {abs(loss - o['loss']) / abs(o['loss']):.2e}), "
          f"logp rel {rel_logp:.3e}, worst grad max-abs/max {worst:.3e} ({names[0]}), worst rel-L2 {worst_l2:.3e} ({names[1]})")
    assert rel_logp <= gate["loss"]
    assert worst <= gate["maxabs"] and worst_l2 <= gate["l2"], (worst, worst_l2, names)
