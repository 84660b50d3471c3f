#!/usr/bin/env python
"""bench.py - ResNet training throughput of the B200 conv/BN/ReLU/pool hot path (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--model preact_resnet18] [--batch 256]

Workload: preact_resnet18 (examples/resnet/preact_resnet18 of the reference), CIFAR-shaped synthetic 3x32x32 input,
batch 256 PER GPU, one SGD(momentum 0.9, wd 1e-4) step = forward + NLL loss + backward + (N>1: gradient all-reduce
+ SyncBN) + update, TF32 tensor-core math.  One JSON line on rank 0:
  value   images/s, whole job, inputs already resident in HBM, K steps timed with CUDA events (max over ranks)
  e2e     images/s through the public API from HOST buffers: pinned H2D of the batch + D2H of the loss every step
  roofline  the dominant kernel family (tcgen05 implicit-GEMM conv), algorithmic FLOPs / CUDA-event time vs the
            measured bf16 peak in MEASURED_PEAKS.json; `hbm` sub-object: BN/ReLU kernels' algorithmic bytes/s
  cpu_baseline  the numpy oracle (port of the reference's algorithm) on the host cores, bounded sample
--impl reference times the reference's own CPU algorithm (oracle port; the reference is pure Python+numpy and
cannot travel to the GPU box) on all host cores.
"""
import argparse
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
    "preact_resnet18": dict(layers=[2, 2, 2, 2], channels=[64, 128, 256, 512], hw=32),
    "small_preact_resnet110": dict(layers=[18, 18, 18], channels=[16, 32, 64], hw=32),
}


def conv_flops_per_image(layers, channels, hw):
    """fwd + dgrad + wgrad algorithmic FLOPs per image (2*MACs; the stem has no dgrad) - SURVEY.md §8(d)."""
    total = 0

    def conv(cin, cout, k, h_out, dgrad=True):
        nonlocal total
        f = 2 * h_out * h_out * cout * cin * k * k
        total += f * (3 if dgrad else 2)
    h = hw
    conv(3, channels[0], 3, h, dgrad=False)
    cin = channels[0]
    for li, (nblk, ch) in enumerate(zip(layers, channels)):
        for b in range(nblk):
            stride = 2 if (li > 0 and b == 0) else 1
            h_out = h // stride
            conv(cin, ch, 3, h_out)
            conv(ch, ch, 3, h_out)
            if stride != 1 or cin != ch:
                conv(cin, ch, 1, h_out)
            cin, h = ch, h_out
    return total


def bn_relu_elements_per_image(layers, channels, hw):
    total, h, cin = 0, hw, channels[0]
    for li, (nblk, ch) in enumerate(zip(layers, channels)):
        for b in range(nblk):
            stride = 2 if (li > 0 and b == 0) else 1
            total += cin * h * h          # act BN+ReLU on the block input
            h //= stride
            total += ch * h * h           # BN+ReLU inside the residual branch
            cin = ch
    total += cin * h * h                  # final BN+ReLU
    return total


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
def oracle_images_per_sec(model, sample_batch, steps, warmup, seed=0):
    from oracle.resnet_oracle import StepOracle, init_params
    cfg = MODELS[model]
    net = StepOracle(cfg["layers"], cfg["channels"], init_params(cfg["layers"], cfg["channels"], seed=seed))
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


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
    except Exception:
        return os.cpu_count() or 1


def workload_text(args):
    cfg = MODELS[args.model]
    return (f"{args.model} CIFAR-shaped 3x{cfg['hw']}x{cfg['hw']} training step (fwd+loss+bwd+SGD), "
            f"batch {args.batch} per GPU")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = MODELS[args.model]
    # bounded sample: pick the per-step batch so that (warmup + steps) finish in about two minutes
    ips_probe, _ = oracle_images_per_sec(args.model, 8, 1, 0)
    budget_s = 120.0
    per_step = max(1, min(args.batch, int(ips_probe * budget_s / max(1, args.steps + args.warmup))))
    per_step = max(2, per_step)
    ips, s_per_step = oracle_images_per_sec(args.model, per_step, args.steps, args.warmup)
    cores = blas_threads()
    line = {
        "impl": "reference", "metric": "ResNet train images/sec", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args), "global_batch": args.batch * args.gpus,
                   "parallelism": "host cpu (reference numpy algorithm, oracle port; rank 0 only)",
                   "batch_per_step_sampled": per_step, "math": "fp32"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of batch {per_step} (bounded sample of the batch-{args.batch} step)"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import pytortto_b200 as tt
    from pytortto_b200 import _cabi
    from pytortto_b200 import distributed as dist
    from pytortto_b200.examples import make_models

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", sync_bn=os.environ.get("TORTTO_B200_DEBUG_NO_SYNCBN") != "1")
    tt.set_math_mode(args.math)
    cfg = MODELS[args.model]
    M = make_models(tt)
    tt.manual_seed(0)  # identical initial parameters on every rank
    net = M["PreactResNet"](M["BasicBlock"], cfg["layers"], cfg["channels"]).cuda()
    ddp = dist.DistributedDataParallel(net) if world > 1 else None
    crit = tt.nn.NLLLoss()
    opt = tt.optim.SGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    net.train()

    B, hw = args.batch, cfg["hw"]
    rng = np.random.default_rng(1234 + rank)
    x_host = torch.from_numpy(rng.standard_normal((B, 3, hw, hw)).astype(np.float32)).pin_memory()
    y_host = torch.from_numpy(rng.integers(0, 10, B).astype(np.int64)).pin_memory()
    x_dev = tt.tensor(x_host.numpy()).cuda()
    y_dev = tt.tensor(y_host.numpy(), dtype=np.int64).cuda()

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

    for _ in range(max(args.warmup, 3)):
        step(x_dev, y_dev)
    launches0 = _cabi.launch_count
    step(x_dev, y_dev)
    launches = _cabi.launch_count - launches0  # kernels-launching C-ABI calls of ONE step (same count when replayed)
    use_graph = bool(args.graph)
    if use_graph:
        # the whole step (fwd + loss + bwd [+ NCCL gradient buckets, peer-memory SyncBN] + SGD, ~270 kernels) recorded
        # once and replayed with one driver call
        graphed = tt.cuda_graph.GraphedStep(step, (x_dev, y_dev), modules=[net])
        run_resident = lambda: graphed(*graphed.static_inputs)
        run_from = graphed
    else:
        run_resident = lambda: step(x_dev, y_dev)
        run_from = step
    for _ in range(3):
        run_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(run_resident, args.steps)

    # end to end: host batch -> pinned H2D + layout kernel -> step -> loss.item() (D2H) every step
    # (the loss of step i is read after step i+1 has been queued - every step's loss is still read inside the timed
    # region, the last one by `drain` - and tt.prefetch.DevicePrefetcher issues the H2D copy + layout kernel of the next
    # batch on a second stream; every batch's copy is issued inside the timed region)
    pending = []

    def host_batches(n):
        for _ in range(n):
            yield (tt.tensor(x_host.numpy(), copy=False), tt.tensor(y_host.numpy(), dtype=np.int64, copy=False))

    def e2e_run(n):
        for xb, yb in tt.prefetch.DevicePrefetcher(host_batches(n)):
            pending.append(run_from(xb, yb).item_async())
            if len(pending) > 1:
                pending.pop(0).get()
        while pending:
            pending.pop(0).get()
    e2e_run(8)  # also lets the caching allocator reach its steady set of staging blocks on the copy stream
    ms_e2e = timed(lambda: e2e_run(args.steps), 1)
    clocks = sampler.stop() if rank == 0 else None

    # per-kernel-family device time (CUDA events around every C-ABI call) on a few extra steps
    # (every rank runs it: the step contains collectives)
    prof = profile_families(step, x_dev, y_dev, steps=min(5, args.steps))

    if rank != 0:
        return
    peaks = read_peaks()
    img_s = world * B * args.steps / (ms_total / 1e3)
    e2e_img_s = world * B * args.steps / (ms_e2e / 1e3)
    flops_step = conv_flops_per_image(cfg["layers"], cfg["channels"], hw) * B
    elems = bn_relu_elements_per_image(cfg["layers"], cfg["channels"], hw) * B
    conv_ms = prof["conv_ms"]
    bn_ms = prof["bn_relu_ms"]
    achieved_tf = flops_step / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    # BN fwd 2R+1W, BN bwd 4R+1W, ReLU fwd 1R+1W, ReLU bwd 2R+1W (fp32) - BASELINE.md §3
    bn_bytes = elems * 4 * (3 + 5 + 2 + 3)
    achieved_gbs = bn_bytes / (bn_ms / 1e3) / 1e9 if bn_ms > 0 else 0.0
    # CPU baseline (rank 0, N = 1 only): the numpy oracle on a bounded sample worth about 10-15 s of host time - a small
    # probe step sizes the batch, then 1 warm-up + 3 timed steps of that batch
    cpu_ips = cpu_s = None
    cpu_sample = "skipped (reported at N = 1 only)"
    if world == 1:
        probe_ips, _ = oracle_images_per_sec(args.model, 8, 1, 0)
        cpu_batch = int(min(args.batch, max(8, probe_ips * 3.5)))
        cpu_ips, cpu_s = oracle_images_per_sec(args.model, cpu_batch, 3, 1)
        cpu_sample = (f"3 steps of batch {cpu_batch} after 1 warm-up ({cpu_s:.1f} s per step) of the same model, numpy oracle "
                      f"(port of the reference's algorithm), all host BLAS threads")
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_conv_family_traffic.json")
    if os.path.exists(tpath) and args.model == "preact_resnet18" and B == 256 and args.math == "tf32":
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_total")  # DRAM bytes of the conv family per step (ncu --set full)
    line = {
        "metric": "ResNet train images/sec", "value": img_s, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.math, "data": "synthetic",
        "config": {"workload": workload_text(args),
                   "global_batch": B * world, "parallelism": f"dp{world}" + ("+syncbn" if world > 1 else ""),
                   "l2": "per-step working set (~2.5 GB of activations at batch 256) >> 126 MB L2; no explicit flush",
                   "math": args.math, "cuda_graph": bool(use_graph)},
        "conv_tflops": achieved_tf,
        "conv_tflops_step_share": conv_ms / (prof["step_ms"] or 1.0),
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": achieved_tf / peaks["bf16_sustained"], "traffic": traffic,
                     "traffic_note": "dram__bytes_read+write summed over the conv-family launches of one step, "
                                     "profiles/r1_conv_family_traffic.json (per step, like `achieved`)",
                     "kernel": "igemm_fwd_persist_kernel + igemm_wgrad_kernel (conv fprop+dgrad+wgrad of one step)",
                     "peak_source": peaks["source"] + "; sustained bf16 figure (kernel timed inside a long step); "
                                    "TF32 math peaks at half of it",
                     "hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": achieved_gbs / peaks["hbm_gbs"], "kernel": "bn_* + relu_* kernels of one step"}},
        "cpu_baseline": {"value": cpu_ips, "unit": "images/s", "cores": blas_threads(), "kind": "port",
                         "sample": cpu_sample},
        "e2e": {"value": e2e_img_s, "unit": "images/s", "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 8),
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches) * args.steps,
        "gpu_launches_per_step": int(launches),
        "clocks": clocks,
        "family_ms_per_step": prof,
    }
    print(json.dumps(line), flush=True)


def profile_families(step, x_dev, y_dev, steps=3):
    """Device time per kernel family per step: CUDA events around every C-ABI call (same stream)."""
    import torch
    from pytortto_b200 import _cabi
    records = []
    orig = _cabi.call

    def wrapped(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *a)
        e1.record()
        records.append((name, e0, e1))
    _cabi.call = wrapped
    import pytortto_b200.ops as ops_mod
    import pytortto_b200.xparray as xp_mod
    try:
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(steps):
            step(x_dev, y_dev)
        s1.record()
        torch.cuda.synchronize()
    finally:
        _cabi.call = orig
    fam = {}
    for name, e0, e1 in records:
        fam[name] = fam.get(name, 0.0) + e0.elapsed_time(e1)
    fam = {k: v / steps for k, v in fam.items()}
    conv = sum(v for k, v in fam.items() if k.startswith("ttb_conv2d"))
    bn = sum(v for k, v in fam.items() if k.startswith("ttb_bn_") or k.startswith("ttb_relu"))
    return {"conv_ms": conv, "bn_relu_ms": bn, "step_ms": s0.elapsed_time(s1) / steps,
            "by_entry_point": {k: round(v, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="preact_resnet18", choices=sorted(MODELS))
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--math", default="tf32", choices=["tf32", "bf16", "fp32"])
    ap.add_argument("--graph", type=int, default=1, help="1: replay the step as a CUDA graph, 0: eager dispatch")
    args = ap.parse_args()
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
