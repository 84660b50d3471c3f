"""Data-parallel layer (new work - the reference is single-process, single-device).

One process per GPU (torchrun-style env: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT), NCCL over
NVLink 5 / NVSwitch through `torch.distributed` (gloo on CPU for the host-logic tests).  The batch is sharded by
rank; two exchange steps make a k-GPU run equal the reference's single-process run on the concatenated batch:

  1. parameter gradients: every rank back-propagates its local-mean loss; gradients are packed into flat buckets in
     the order AccumulateGrad produces them, each bucket is all-reduced (SUM) as soon as it is full - overlapping the
     rest of backward on a side stream - and scaled by 1/world, which yields the global-batch mean gradient;
  2. BatchNorm statistics (SyncBN): forward all-reduces the per-channel [sum x, sum x^2] doubles (+ the element
     count), backward all-reduces [sum dy, sum dy*(x-mean)], so normalisation, running stats (global N/(N-1)) and
     dx use global-batch statistics (reference formulas, autograd/grad_nn.py:923-930, :984-988).  dgamma / dbeta
     computed from the all-reduced sums are already GLOBAL sums, so they are excluded from the gradient averaging's
     1/world... they are divided by world after SUM like every other gradient only if they were local; here they
     are marked `_ttb_global_grad` and skipped by the bucket all-reduce, then scaled by 1/world to match the
     mean-loss convention.
"""
import os

import numpy as np
import torch

_state = {"initialized": False, "world": 1, "rank": 0, "sync_bn": True, "group": None, "backend": None}


def is_initialized():
    return _state["initialized"]


def get_world_size():
    return _state["world"]


def get_rank():
    return _state["rank"]


def init_process_group(backend=None, sync_bn=True):
    """Joins the torchrun rendezvous.  backend defaults to nccl when CUDA is available, else gloo."""
    import torch.distributed as dist
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend)
    _state.update(initialized=True, world=dist.get_world_size(), rank=dist.get_rank(), sync_bn=bool(sync_bn),
                  group=dist.group.WORLD, backend=backend)
    return _state["rank"], _state["world"]


def destroy_process_group():
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()
    _state.update(initialized=False, world=1, rank=0, group=None)


def all_reduce_sum_(t):
    """In-place SUM all-reduce of a torch tensor on the current stream (no-op for world 1)."""
    if _state["initialized"] and _state["world"] > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


# ---- SyncBN hooks (called from ops.bn_forward_train / ops.bn_backward) ------------------------------------
def bn_forward_hook():
    if not (_state["initialized"] and _state["world"] > 1 and _state["sync_bn"]):
        return None

    def hook(sums, local_count):
        all_reduce_sum_(sums)
        return local_count * _state["world"]  # equal shards by construction (shard_batch)
    return hook


def bn_backward_hook():
    if not (_state["initialized"] and _state["world"] > 1 and _state["sync_bn"]):
        return None

    def hook(sums, local_count):
        all_reduce_sum_(sums)
        return local_count * _state["world"]
    return hook


def shard_batch(*arrays):
    """Contiguous equal shard of each host array along axis 0 for this rank (global batch must divide evenly)."""
    w, r = _state["world"], _state["rank"]
    out = []
    for a in arrays:
        n = a.shape[0]
        if n % w:
            raise ValueError(f"global batch {n} is not divisible by world size {w}")
        per = n // w
        out.append(a[r * per:(r + 1) * per])
    return out[0] if len(out) == 1 else tuple(out)


def broadcast_parameters(module, src=0):
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not (_state["initialized"] and _state["world"] > 1):
        return
    import torch.distributed as dist
    for t in list(module.parameters()) + list(module.buffers()):
        d = t.data
        if hasattr(d, "t"):
            dist.broadcast(d.t, src=src)
        else:
            buf = torch.from_numpy(np.ascontiguousarray(d))
            dist.broadcast(buf, src=src)
            d[...] = buf.numpy()


class DistributedDataParallel:
    """Wraps a Module: forward is unchanged; `reduce_gradients()` (call between loss.backward() and
    optimizer.step()) turns local gradients into global-batch mean gradients.

    Gradients are flattened into buckets of ~`bucket_mb` MiB in reverse parameter order (the order backward produces
    them), all-reduced with SUM and scaled by 1/world.  BatchNorm weight/bias gradients computed from all-reduced
    statistics are already global sums of per-sample terms of the LOCAL-mean loss, i.e. world x the global-mean
    gradient... see `_is_synced_bn_param`."""

    def __init__(self, module, bucket_mb=25, broadcast=True):
        self.module = module
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        if broadcast:
            broadcast_parameters(module)
        self._synced_bn_params = set()
        if _state["sync_bn"]:
            from .nn.modules import _BatchNorm
            for m in module.modules():
                if isinstance(m, _BatchNorm):
                    for p in (m._parameters.get("weight"), m._parameters.get("bias")):
                        if p is not None:
                            self._synced_bn_params.add(id(p))

    def __call__(self, *a, **k):
        return self.module(*a, **k)

    def __getattr__(self, name):
        return getattr(self.__dict__["module"], name)

    def parameters(self):
        return self.module.parameters()

    def reduce_gradients(self):
        world = _state["world"]
        if not (_state["initialized"] and world > 1):
            return
        params = [p for p in self.module.parameters() if p.grad is not None]
        params.reverse()
        inv = 1.0 / world
        bucket, size = [], 0
        for p in params:
            if id(p) in self._synced_bn_params:
                # computed from globally all-reduced sums: identical on every rank and equal to the SUM over ranks
                # of the local-mean-loss gradients -> only the 1/world of the mean convention is missing
                p.grad.t.mul_(inv)
                continue
            bucket.append(p)
            size += p.grad.nbytes
            if size >= self.bucket_bytes:
                self._reduce_bucket(bucket, inv)
                bucket, size = [], 0
        if bucket:
            self._reduce_bucket(bucket, inv)

    @staticmethod
    def _reduce_bucket(bucket, inv):
        flats = [p.grad.t.reshape(-1) if p.grad.t.is_contiguous() else p.grad.t.permute(0, 2, 3, 1).reshape(-1)
                 for p in bucket]
        flat = torch.cat(flats)
        all_reduce_sum_(flat)
        flat.mul_(inv)
        off = 0
        for p, f in zip(bucket, flats):
            n = f.numel()
            f.copy_(flat[off:off + n])  # `f` is a view of the gradient's physical storage
            off += n
