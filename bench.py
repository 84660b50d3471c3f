#!/usr/bin/env python
"""bench.py - training throughput of the B200 conv / BatchNorm / ReLU / MaxPool hot path (BASELINE.json configs).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--model M] [--batch B] [--math tf32|bf16|fp32]

Default workload = BASELINE.json configs[1] (the configuration the metric is quoted on): preact_resnet18
(examples/resnet/preact_resnet18 of the reference), CIFAR-shaped synthetic 3x32x32 input, batch 256 PER GPU, one
SGD(momentum 0.9, wd 1e-4) step = forward + NLL loss + backward + (N>1: gradient all-reduce + SyncBN) + update, TF32
tensor-core math.  `--model` selects the other configurations as extra measurements (never the driver's default):
small_preact_resnet110 (cfg 1, batch 128), standard_resnet50 (cfg 3: 224x224, bf16, batch 256 per GPU), unet (cfg 4:
3x512x512, Conv + ConvTranspose2d + BatchNorm, BCE-with-logits, Adam, batch 8 per GPU).  The single-layer sweep of cfg 5 is
scripts/conv_sweep.py.

One JSON line on rank 0:
  value     images/s, whole job, inputs already resident in HBM, K steps timed with CUDA events (max over ranks)
  e2e       images/s through the public API from HOST buffers: pinned H2D of the batch + D2H of the loss every step
  roofline  the dominant kernel family (tcgen05 implicit-GEMM conv): algorithmic FLOPs of one step / summed device time
            of its launches (CUDA events around every C-ABI call, on the launching stream, with the GPU kept behind the
            host so that no launch latency is inside an interval) vs the measured bf16 peak in MEASURED_PEAKS.json;
            `hbm` sub-object: BatchNorm / ReLU / MaxPool kernels' algorithmic bytes/s vs the measured HBM peak
  bf16      (default workload only) the same step re-measured in bf16 tensor-core mode - the north-star conv target is
            quoted against the bf16 peak, which TF32 math cannot exceed half of
  cpu_baseline  the numpy oracle (port of the reference's algorithm) on the host cores, bounded sample
  dp_parity (N > 1) N-rank step on a global batch vs the single-process step on the same batch, exact-fp32 mode
--impl reference times the reference's own CPU algorithm (oracle port; the reference is pure Python+numpy and
cannot travel to the GPU box) on all host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODELS = {
    # name: builder key in examples.make_models, input HxW, default batch per GPU, default math, loss, optimizer, text
    "preact_resnet18": dict(cfg=2, hw=32, batch=256, math="tf32", loss="nll", opt="sgd",
                            oracle=dict(layers=[2, 2, 2, 2], channels=[64, 128, 256, 512]),
                            text="preact_resnet18 CIFAR-shaped 3x32x32 training step (fwd+loss+bwd+SGD)"),
    "small_preact_resnet110": dict(cfg=1, hw=32, batch=128, math="tf32", loss="nll", opt="sgd",
                                   oracle=dict(layers=[18, 18, 18], channels=[16, 32, 64]),
                                   text="small_preact_resnet110 CIFAR-shaped 3x32x32 training step (fwd+loss+bwd+SGD)"),
    "standard_resnet50": dict(cfg=3, hw=224, batch=256, math="bf16", loss="nll", opt="sgd", oracle=None,
                              text="standard_resnet50 ImageNet-shaped 3x224x224 training step (fwd+loss+bwd+SGD)"),
    "unet": dict(cfg=4, hw=512, batch=8, math="tf32", loss="bce", opt="adam", oracle=None,
                 text="UNet(3,1,[32,64,128,256]) Carvana-shaped 3x512x512 training step (fwd+BCE+bwd+Adam)"),
}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(bf16_tflops=d.get("bf16_tflops", 1590.0), bf16_sustained=d.get("bf16_tflops_sustained", 1400.0),
                    hbm_gbs=d.get("hbm_gbs", 6650.0), source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_tflops=1590.0, bf16_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the numpy oracle (port of the reference's im2col+GEMM algorithm) on host cores
# ---------------------------------------------------------------------------------------------------------------
def _all_host_threads():
    """The CPU arm uses every host core BLAS can use.  torchrun exports OMP_NUM_THREADS=1 to its workers by default, which
    would time the reference single-threaded at N > 1: lift the limit for the numpy legs (threadpoolctl)."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def oracle_images_per_sec(model, sample_batch, steps, warmup, seed=0):
    with _all_host_threads():
        return _oracle_images_per_sec(model, sample_batch, steps, warmup, seed)


def _oracle_images_per_sec(model, sample_batch, steps, warmup, seed=0):
    from oracle.resnet_oracle import StepOracle, init_params
    cfg = MODELS[model]
    oc = cfg["oracle"]
    net = StepOracle(oc["layers"], oc["channels"], init_params(oc["layers"], oc["channels"], seed=seed))
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((sample_batch, 3, cfg["hw"], cfg["hw"])).astype(np.float32)
    lab = rng.integers(0, 10, sample_batch).astype(np.int64)
    for _ in range(warmup):
        net.train_step(x, lab)
    t0 = time.perf_counter()
    for _ in range(steps):
        net.train_step(x, lab)
    dt = time.perf_counter() - t0
    return sample_batch * steps / dt, dt / steps


def cpu_sample_batch(model, batch, budget_s, steps):
    """Per-step batch of the bounded CPU sample: one probe step sizes it so that `steps` steps take about budget_s.
    The CPU arm's throughput is per image (the step's cost is linear in the batch), so the sampled batch is stated and
    the number is directly comparable with the GPU arm's images/s."""
    probe_ips, _ = oracle_images_per_sec(model, 8, 1, 0)
    return int(max(2, min(batch, probe_ips * budget_s / max(1, steps))))


def blas_threads():
    """threads the numpy legs ran on (inside _all_host_threads)"""
    try:
        from threadpoolctl import threadpool_info
        with _all_host_threads():
            return max([p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
    except Exception:
        return os.cpu_count() or 1


def workload_text(args):
    return f"{MODELS[args.model]['text']}, batch {args.batch} per GPU"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = MODELS[args.model]
    if cfg["oracle"] is None:
        print(json.dumps({"impl": "reference", "unavailable": f"no CPU oracle network for {args.model}; the CPU arm "
                          "covers the PreactResNet configurations (cfg 1, 2) and scripts/conv_sweep.py the single layers"}))
        return
    # bounded sample: the per-step batch is sized so that (warmup + steps) finish in about two minutes
    per_step = cpu_sample_batch(args.model, args.batch, 120.0, args.steps + args.warmup)
    ips, s_per_step = oracle_images_per_sec(args.model, per_step, args.steps, args.warmup)
    cores = blas_threads()
    line = {
        "impl": "reference", "metric": "ResNet train images/sec", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args), "global_batch": args.batch * args.gpus,
                   "parallelism": "host cpu (reference numpy algorithm, oracle port; rank 0 only)",
                   "batch_per_step_sampled": per_step, "math": "fp32",
                   "normalisation": "images/s = sampled batch x steps / wall time; the step cost is linear in the batch, so "
                                    "the per-image rate is the full-batch rate"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of batch {per_step} (bounded sample of the batch-{args.batch} step; "
                                   f"per-image rate)"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def build_workload(tt, args, rank):
    """-> net, criterion, optimizer, pinned host batch (x, y), label dtype"""
    import torch
    from pytortto_b200.examples import make_models
    cfg = MODELS[args.model]
    M = make_models(tt)
    tt.manual_seed(0)  # identical initial parameters on every rank
    if args.model == "unet":
        net = M["UNet"](3, 1, [32, 64, 128, 256])
    else:
        net = M[args.model]()
    net.cuda()
    B, hw = args.batch, cfg["hw"]
    rng = np.random.default_rng(1234 + rank)
    x_host = torch.from_numpy(rng.standard_normal((B, 3, hw, hw)).astype(np.float32)).pin_memory()
    if cfg["loss"] == "nll":
        crit = tt.nn.NLLLoss()
        y_host = torch.from_numpy(rng.integers(0, 10, B).astype(np.int64)).pin_memory()
        ydt = np.int64
    else:
        crit = tt.nn.BCEWithLogitsLoss()
        y_host = torch.from_numpy((rng.random((B, 1, hw, hw)) < 0.5).astype(np.float32)).pin_memory()
        ydt = np.float32
    if cfg["opt"] == "sgd":
        opt = tt.optim.SGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    else:
        opt = tt.optim.Adam(net.parameters(), lr=1e-3)
    net.train()
    return net, crit, opt, x_host, y_host, ydt


def dp_parity_check(tt, dist, world, rank):
    """N ranks on a global batch == one process on the same batch (exact-fp32 mode, full preact_resnet18, 32 images per
    rank, one step), in two variants:
      * the network as it is: loss and every BatchNorm running statistic (forward: the SyncBN exchange of every layer) to
        1e-5; every parameter gradient to 2e-2 rel-L2.  The gradients cannot be held tighter: the two runs group the
        BatchNorm sums differently, a last-bit difference can put single ReLU decisions on the other side, and one decision
        moves every gradient upstream of it by ~2e-3 (the mechanism measured in profiles/r2_relu_flip_analysis.txt);
      * the same network with its ReLUs replaced by the identity (no decisions: a smooth function of the statistics):
        every parameter gradient to 2e-3 of the tensor max (median ~3e-6) - this is the check of the gradient plumbing (flat
        buckets, in-place AVG all-reduce, 1/world of the SyncBN affine gradients, backward statistic exchange).
    Raises if either does not hold."""
    import torch
    from pytortto_b200.examples import make_models
    M = make_models(tt)
    mode0 = tt.get_math_mode()
    tt.set_math_mode("fp32")
    per = 32
    rng = np.random.default_rng(99)
    x = rng.standard_normal((per * world, 3, 32, 32)).astype(np.float32)
    lab = rng.integers(0, 10, per * world).astype(np.int64)

    def one(dp, relu):
        tt.manual_seed(3)
        net = M["preact_resnet18"]()
        if not relu:
            for m in list(net.modules()):
                for name, child in list(m._modules.items()):
                    if isinstance(child, tt.nn.ReLU):
                        m._modules[name] = tt.nn.Sequential()
        net.cuda().train()
        ddp = dist.DistributedDataParallel(net, bucket_mb=4) if dp else None
        xs, ls = dist.shard_batch(x, lab) if dp else (x, lab)
        loss = tt.nn.NLLLoss()(net(tt.tensor(xs).cuda()), tt.tensor(ls, dtype=np.int64).cuda())
        loss.backward()
        if ddp is not None:
            ddp.reduce_gradients()
            ddp.close()
        lv = torch.tensor([loss.item()], device="cuda", dtype=torch.float64)
        if dp:  # the global loss is the mean of the per-rank local-mean losses
            torch.distributed.all_reduce(lv)
            lv /= world
        grads = {k: p.grad.get() for k, p in net.named_parameters()}
        bufs = {k: np.asarray(v) for k, v in net.state_dict().items() if "running" in k}
        return float(lv.item()), grads, bufs

    def rel(a, b):
        return float(np.abs(a.astype(np.float64) - b).max() / max(float(np.abs(b).max()), 1e-30))

    def rel_l2(a, b):
        return float(np.linalg.norm(a.astype(np.float64) - b) / max(float(np.linalg.norm(b.astype(np.float64))), 1e-30))

    out = {}
    for relu in (True, False):
        l_dp, g_dp, b_dp = one(True, relu)
        with dist.single_process():
            l_1, g_1, b_1 = one(False, relu)
        fwd = max([abs(l_dp - l_1) / max(abs(l_1), 1e-30)] + [rel(b_dp[k], b_1[k]) for k in b_1])
        errs = {k: (rel_l2 if relu else rel)(g_dp[k], g_1[k]) for k in g_1}
        worst_name = max(errs, key=errs.get)
        stats = torch.tensor([fwd, errs[worst_name], float(np.median(list(errs.values())))], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(stats, op=torch.distributed.ReduceOp.MAX)
        fwd, worst, med = (float(v) for v in stats.tolist())
        # (without ReLU the activations are not halved layer by layer: plain fp32 rounding of the deep linear net is ~1e-5
        # on the forward statistics and ~2e-4 on the worst cancellation-prone BatchNorm bias gradient, measured; median 3e-6)
        tol, ftol = (2e-2, 1e-5) if relu else (2e-3, 1e-4)
        out["with_relu" if relu else "relu_as_identity"] = {
            "forward_worst_rel_err": fwd, "forward_tol": ftol, "grad_worst_err": worst, "grad_median_err": med,
            "grad_metric": "rel-L2" if relu else "max-abs / tensor max", "grad_tol": tol, "grad_worst_tensor": worst_name,
            "ok": fwd <= ftol and worst <= tol}
    tt.set_math_mode(mode0)
    res = {"ok": all(v["ok"] for v in out.values()), "mode": "fp32", "global_batch": per * world, "tensors": len(g_1),
           "running_stats": len(b_1), **out,
           "what": f"{world}-rank step (flat gradient buckets + SyncBN) vs single-process step on the same global batch: loss + "
                   f"BatchNorm running statistics (forward) and every parameter gradient"}
    if not res["ok"]:
        raise RuntimeError(f"data-parallel parity check failed: {res}")
    return res


def profile_families(step, x_dev, y_dev, steps=3):
    """Device time per C-ABI entry point per step + algorithmic work counted from the calls' own descriptors.

    CUDA events bracket every call on the stream it launches on (the wgrad fork to a second stream is switched off for
    this pass so that is always the current stream), and the GPU is first parked behind a long spin kernel so the host
    has queued the whole measured region before the device starts it: an interval then holds the kernel(s) of one call
    and nothing else (no launch latency, no host dispatch time)."""
    import torch
    import pytortto_b200 as tt
    from pytortto_b200 import _cabi
    records = []
    orig = _cabi.call
    work = {"conv_flops": 0.0, "hbm_bytes": 0.0}

    def wrapped(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *a)
        e1.record()
        records.append((name, e0, e1))
        if name.startswith("ttb_conv2d_") and hasattr(a[0], "_obj") and name.split("ttb_conv2d_")[1].split("_")[0] in (
                "fprop", "dgrad", "wgrad"):
            d = a[0]._obj  # FLOP convention of SURVEY.md §8(d): 2*N*P*Q*K*(C/g)*R*S per pass
            work["conv_flops"] += 2.0 * d.n * d.p * d.q * d.k * (d.c // d.groups) * d.r * d.s
        elif name == "ttb_bn_apply":        # BN fwd 2R+1W, bwd 4R+1W; fused ReLU credited fwd 1R+1W, bwd 2R+1W (fp32)
            m, c, relu = a[2], a[3], a[7]
            work["hbm_bytes"] += 4.0 * m * c * ((3 + 5) + ((2 + 3) if relu else 0))
        elif name == "ttb_bn_apply_add":    # bn(x) + identity (+ReLU) in one pass: credited as the three operators it replaces
            m, c, relu = a[3], a[4], a[8]
            work["hbm_bytes"] += 4.0 * m * c * ((3 + 5) + 3 + ((2 + 3) if relu else 0))
        elif name == "ttb_relu_fwd":
            work["hbm_bytes"] += 4.0 * a[2] * (2 + 3)
        elif name == "ttb_add":             # residual add / gradient accumulation: 2R + 1W
            work["hbm_bytes"] += 4.0 * a[3] * 3
        elif name == "ttb_add_bias":        # y += bias[c] in place: 1R + 1W
            work["hbm_bytes"] += 4.0 * a[2] * a[3] * 2
        elif name == "ttb_add_bn_stats":    # the residual add that also emits the next BatchNorm's statistics: 2R + 1W
            work["hbm_bytes"] += 4.0 * a[3] * a[4] * 3  # (the statistics read it replaces is credited with ttb_bn_apply)
        elif name == "ttb_maxpool2d_fwd":   # fwd reads x, writes y; bwd reads dy, writes dx (index bytes not credited)
            d = a[0]._obj
            work["hbm_bytes"] += 4.0 * 2 * (d.n * d.c * d.h * d.w + d.n * d.c * d.p * d.q)
    tt.set_wgrad_overlap(False)
    _cabi.call = wrapped
    try:
        torch.cuda.synchronize()
        step(x_dev, y_dev)  # (allocator warm-up under the changed stream assignment)
        records.clear()
        work.update(conv_flops=0.0, hbm_bytes=0.0)
        torch.cuda.synchronize()
        torch.cuda._sleep(int(2.0e9 * 0.05 * steps))  # ~50 ms per measured step: the host queues everything meanwhile
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(steps):
            step(x_dev, y_dev)
        s1.record()
        torch.cuda.synchronize()
    finally:
        _cabi.call = orig
        tt.set_wgrad_overlap(True)
    fam = {}
    for name, e0, e1 in records:
        fam[name] = fam.get(name, 0.0) + e0.elapsed_time(e1)
    fam = {k: v / steps for k, v in fam.items()}
    conv = sum(v for k, v in fam.items() if k.startswith("ttb_conv2d"))
    hbm = sum(v for k, v in fam.items() if k.startswith(("ttb_bn_", "ttb_relu", "ttb_maxpool", "ttb_comm_bn_", "ttb_add")))
    return {"conv_ms": conv, "bn_relu_pool_ms": hbm, "step_ms_serialised": s0.elapsed_time(s1) / steps,
            "conv_flops_per_step": work["conv_flops"] / steps, "hbm_bytes_per_step": work["hbm_bytes"] / steps,
            "by_entry_point": {k: round(v, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])}}


def roofline_of(prof, peaks, math, traffic=None, traffic_note=None):
    achieved_tf = prof["conv_flops_per_step"] / (prof["conv_ms"] / 1e3) / 1e12 if prof["conv_ms"] > 0 else 0.0
    achieved_gbs = prof["hbm_bytes_per_step"] / (prof["bn_relu_pool_ms"] / 1e3) / 1e9 if prof["bn_relu_pool_ms"] > 0 else 0.0
    r = {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
         "frac": achieved_tf / peaks["bf16_sustained"], "traffic": traffic,
         "kernel": "conv family: igemm_fwd_persist_kernel (fprop, dgrad) + igemm_wgrad_kernel (+ split sums, weight "
                   "re-packs) of one step",
         "algorithmic_flops_per_step": prof["conv_flops_per_step"], "family_ms_per_step": prof["conv_ms"],
         "peak_source": peaks["source"] + "; sustained bf16 figure (kernels timed inside a long step)"
                        + ("; TF32 math peaks at half of it" if math == "tf32" else ""),
         "hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                 "frac": achieved_gbs / peaks["hbm_gbs"], "algorithmic_bytes_per_step": prof["hbm_bytes_per_step"],
                 "family_ms_per_step": prof["bn_relu_pool_ms"],
                 "kernel": "ttb_bn_* (+ ttb_comm_bn_* under SyncBN) + ttb_relu_* + ttb_maxpool2d_* + ttb_add* calls of one "
                           "step; fused passes credited with the unfused algorithmic bytes (SURVEY.md 8(d)): BatchNorm "
                           "statistics emitted by a conv epilogue / the fused add count as the read they replace"}}
    if traffic_note:
        r["traffic_note"] = traffic_note
    return r


def run_ours(args):
    import torch
    import pytortto_b200 as tt
    from pytortto_b200 import _cabi
    from pytortto_b200 import distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    _cabi.load()
    if world > 1:
        dist.init_process_group("nccl", sync_bn=True)
    parity = dp_parity_check(tt, dist, world, rank) if world > 1 else None
    tt.set_math_mode(args.math)
    cfg = MODELS[args.model]
    net, crit, opt, x_host, y_host, ydt = build_workload(tt, args, rank)
    ddp = dist.DistributedDataParallel(net) if world > 1 else None
    B = args.batch
    x_dev = tt.tensor(x_host.numpy()).cuda()
    y_dev = tt.tensor(y_host.numpy(), dtype=ydt).cuda()

    def step(x, y):
        opt.zero_grad()
        loss = crit(net(x), y)
        loss.backward()
        if ddp is not None:
            ddp.reduce_gradients()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, after=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if after is not None:
            after()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item())

    def measure(steps, with_e2e):
        """-> (ms of `steps` resident steps, ms of `steps` end-to-end steps or None, C-ABI calls per step)"""
        for _ in range(max(args.warmup, 3)):
            step(x_dev, y_dev)
        launches0 = _cabi.launch_count
        step(x_dev, y_dev)
        launches = _cabi.launch_count - launches0  # kernel-launching C-ABI calls of ONE step (same count when replayed)
        if args.graph:
            # the whole step (fwd + loss + bwd [+ NCCL gradient buckets, peer-memory SyncBN] + update) recorded once and
            # replayed with one driver call
            graphed = tt.cuda_graph.GraphedStep(step, (x_dev, y_dev), modules=[net])
            run_resident = lambda: graphed(*graphed.static_inputs)
            run_from = graphed
        else:
            run_resident = lambda: step(x_dev, y_dev)
            run_from = step
        for _ in range(3):
            run_resident()
        ms_total = timed(run_resident, steps)
        if not with_e2e:
            return ms_total, None, launches
        # end to end: host batch -> pinned H2D + layout kernel -> step -> loss.item() (D2H) every step
        # (the loss of step i is read after step i+1 has been queued - every step's loss is still read inside the timed
        # region, the last one by the drain - and tt.prefetch.DevicePrefetcher issues the H2D copy + layout kernel of the
        # next batch on a second stream; every batch's copy is issued inside the timed region)
        pending = []

        def host_batches(n):
            for _ in range(n):
                yield (tt.tensor(x_host.numpy(), copy=False), tt.tensor(y_host.numpy(), dtype=ydt, copy=False))

        def e2e_run(n):
            for xb, yb in tt.prefetch.DevicePrefetcher(host_batches(n)):
                pending.append(run_from(xb, yb).item_async())
                if len(pending) > 1:
                    pending.pop(0).get()
            while pending:
                pending.pop(0).get()
        e2e_run(8)  # also lets the caching allocator reach its steady set of staging blocks on the copy stream
        ms_e2e = timed(lambda: e2e_run(steps), 1)
        return ms_total, ms_e2e, launches

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, ms_e2e, launches = measure(args.steps, True)
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel-family device time on a few extra steps (every rank runs it: the step contains collectives)
    prof = profile_families(step, x_dev, y_dev, steps=min(3, args.steps))

    # the same step in bf16 tensor-core mode (default workload only): value + conv roofline against the bf16 peak
    bf16 = None
    if args.model == "preact_resnet18" and args.math == "tf32" and args.extra_bf16:
        tt.set_math_mode("bf16")
        ms_b, _, _ = measure(args.steps, False)
        prof_b = profile_families(step, x_dev, y_dev, steps=min(3, args.steps))
        bf16 = (ms_b, prof_b)
        tt.set_math_mode(args.math)

    if rank != 0:
        return
    peaks = read_peaks()
    img_s = world * B * args.steps / (ms_total / 1e3)
    e2e_img_s = world * B * args.steps / (ms_e2e / 1e3)
    # CPU baseline (rank 0, N = 1 only): the numpy oracle on a bounded sample worth about 10-15 s of host time - a small
    # probe step sizes the batch, then 1 warm-up + 3 timed steps of that batch
    cpu_ips = None
    cpu_sample = "skipped (reported at N = 1 only)"
    if not args.cpu_baseline:
        cpu_sample = "skipped (--cpu-baseline 0)"
    elif world == 1 and cfg["oracle"] is not None:
        cpu_batch = cpu_sample_batch(args.model, args.batch, 10.5, 3)
        cpu_ips, cpu_s = oracle_images_per_sec(args.model, cpu_batch, 3, 1)
        cpu_sample = (f"3 steps of batch {cpu_batch} after 1 warm-up ({cpu_s:.1f} s per step) of the same model, numpy oracle "
                      f"(port of the reference's algorithm), all host BLAS threads; per-image rate (step cost is linear in "
                      f"the batch)")
    elif world == 1:
        cpu_sample = f"no CPU oracle network for {args.model} (cfg 1 / 2 have one; scripts/conv_sweep.py times single layers)"
    traffic, tnote = None, None
    tpath = os.path.join(ROOT, "profiles", "r2_conv_family_traffic.json")
    if os.path.exists(tpath) and args.model == "preact_resnet18" and B == 256 and args.math == "tf32":
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_total")  # DRAM bytes of the conv family per step (ncu --set full)
        tnote = ("dram__bytes_read+write summed over the conv-family launches of one step, "
                 "profiles/r2_conv_family_traffic.json (per step, like `achieved`)")
    line = {
        "metric": "ResNet train images/sec" if cfg["loss"] == "nll" else "UNet train images/sec", "value": img_s,
        "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.math, "data": "synthetic",
        "config": {"workload": workload_text(args), "baseline_config": cfg["cfg"],
                   "global_batch": B * world, "parallelism": f"dp{world}" + ("+syncbn" if world > 1 else ""),
                   "l2": "per-step working set (GBs of activations) >> 126 MB L2; no explicit flush",
                   "math": args.math, "cuda_graph": bool(args.graph)},
        "conv_tflops": prof["conv_flops_per_step"] / (prof["conv_ms"] / 1e3) / 1e12 if prof["conv_ms"] else None,
        "roofline": roofline_of(prof, peaks, args.math, traffic, tnote),
        "cpu_baseline": {"value": cpu_ips, "unit": "images/s", "cores": blas_threads(), "kind": "port",
                         "sample": cpu_sample},
        "e2e": {"value": e2e_img_s, "unit": "images/s",
                "h2d_bytes_per_step": int(x_host.numel() * x_host.element_size() + y_host.numel() * y_host.element_size()),
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches) * args.steps,
        "gpu_launches_per_step": int(launches),
        "clocks": clocks,
        "family_ms_per_step": prof,
    }
    if bf16 is not None:
        ms_b, prof_b = bf16
        line["bf16"] = {"value": world * B * args.steps / (ms_b / 1e3), "unit": "images/s", "ms_per_step": ms_b / args.steps,
                        "what": "the same workload and step count re-measured in bf16 tensor-core mode (bf16 operands "
                                "co-written by the BatchNorm kernels, fp32 accumulation and outputs; parity <= 1e-2)",
                        "roofline": roofline_of(prof_b, peaks, "bf16"), "family_ms_per_step": prof_b}
    if parity is not None:
        line["dp_parity"] = parity
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="preact_resnet18", choices=sorted(MODELS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (0: the configuration's own)")
    ap.add_argument("--math", default="", choices=["", "tf32", "bf16", "fp32"])
    ap.add_argument("--graph", type=int, default=1, help="1: replay the step as a CUDA graph, 0: eager dispatch")
    ap.add_argument("--extra-bf16", type=int, default=1, help="default workload: also measure the step in bf16 mode")
    ap.add_argument("--cpu-baseline", type=int, default=1, help="0: skip the CPU (numpy oracle) leg - for quick A/B runs")
    args = ap.parse_args()
    args.batch = args.batch or MODELS[args.model]["batch"]
    args.math = args.math or MODELS[args.model]["math"]
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
