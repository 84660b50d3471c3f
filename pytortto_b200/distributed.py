"""Data-parallel layer (new work: the reference is single-process, single-device - SURVEY.md §2c, §8(e)).

One process per GPU (torchrun environment: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT), NCCL over
NVLink 5 / NVSwitch through `torch.distributed` (gloo for the CPU host-logic tests).  The global batch is sharded
by rank (`shard_batch`); two exchange steps make a k-GPU run equal the reference's single-process run on the
concatenated batch:

1. Parameter gradients.  Every rank back-propagates the mean loss of ITS shard.  `DistributedDataParallel` hooks
   `AccumulateGrad` (the leaf node of the autograd graph, reference autograd/function.py:70-93): as soon as a
   parameter's gradient is final it is appended to the current bucket (reverse parameter order = the order
   backward produces them); a full bucket (~10 MB) is flattened and all-reduced (SUM) asynchronously on NCCL's own
   stream while the rest of backward keeps the compute stream busy.  `reduce_gradients()` (called between
   `loss.backward()` and `optimizer.step()`) waits for the outstanding buckets, scales by 1/world - the mean over
   ranks of local-mean gradients IS the global-batch mean gradient - and scatters the result back into `p.grad`.

2. BatchNorm statistics (SyncBN).  BatchNorm.forward all-reduces the per-channel double sums [sum x, sum x^2]
   (2C values) and uses count = world * local count, so normalisation and the running statistics (unbiased variance
   with the GLOBAL N/(N-1), reference autograd/grad_nn.py:923-930) are those of the global batch; BatchNorm.backward
   all-reduces [sum dy, sum dy*(x-mean)] so dx follows the reference formula (:984-988) with global sums.
   dgamma / dbeta computed from those all-reduced sums are identical on every rank and equal the SUM over ranks of
   the local-loss gradients, so they skip the bucket all-reduce and only receive the 1/world scaling.
"""
import os

import numpy as np
import torch

from . import ops
from .autograd.function import AccumulateGrad

_state = {"initialized": False, "world": 1, "rank": 0, "sync_bn": True, "backend": None, "peer_comm": None}


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
                  backend=backend, peer_comm=None)
    if backend == "nccl" and sync_bn and _state["world"] > 1 and os.environ.get("TORTTO_B200_PEER_COMM", "1") != "0":
        try:
            enable_peer_comm()
        except RuntimeError as e:  # CUDA IPC not permitted between these processes: keep the NCCL path
            import warnings
            warnings.warn(f"peer-memory SyncBN path unavailable ({e}); using NCCL all-reduce")
            _state["peer_comm"] = None
    return _state["rank"], _state["world"]


def destroy_process_group():
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()
    _state.update(initialized=False, world=1, rank=0, peer_comm=None)


class single_process:
    """Context manager: inside it this rank behaves like a single-process run (no SyncBN exchange, no gradient
    all-reduce) although the process group stays up - used to compute the single-process reference of a parity check."""

    def __enter__(self):
        self._world = _state["world"]
        _state["world"] = 1
        return self

    def __exit__(self, *exc):
        _state["world"] = self._world
        return False


def all_reduce_sum_(t, async_op=False, average=False):
    """In-place SUM (or, with `average` on NCCL, AVG) all-reduce of a torch tensor (no-op for world 1).  Returns the
    work handle when async."""
    if _state["initialized"] and _state["world"] > 1:
        if os.environ.get("TORTTO_B200_DEBUG_SKIP_GRAD_ALLREDUCE") == "1":  # timing experiments only (wrong results)
            return None
        import torch.distributed as dist
        op = dist.ReduceOp.AVG if (average and _state["backend"] == "nccl") else dist.ReduceOp.SUM
        return dist.all_reduce(t, op=op, async_op=async_op)
    return None


# ---- SyncBN statistic exchange (called from ops.bn_forward_train / ops.bn_backward) ---------------------------
class PeerComm:
    """Small-message all-reduce over NVLink peer memory (csrc/comm.cu): every rank owns a cudaMalloc'ed buffer of
    fixed slots, exported through CUDA IPC and mapped by every peer.  A call site (BatchNorm layer x direction) owns
    one slot for the life of the process, so replays of a captured CUDA graph keep working."""
    MAX_VALUES = 2 * 4096   # 2C doubles, C <= 4096
    SLOTS = 256

    def __init__(self):
        import ctypes
        import torch.distributed as dist
        from . import _cabi
        self._cabi, self._ct = _cabi, ctypes
        lib = _cabi.load()
        self.world, self.rank = _state["world"], _state["rank"]
        self.slot_bytes = (int(lib.ttb_comm_slot_bytes(self.MAX_VALUES)) + 255) // 256 * 256
        nbytes = self.slot_bytes * self.SLOTS
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        _cabi.call("ttb_comm_alloc", nbytes, ctypes.byref(ptr), handle)
        self.own = ptr.value
        mine = torch.tensor(list(handle), dtype=torch.uint8, device="cuda")
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine)
        bases = []
        self._opened = []
        for r, h in enumerate(gathered):
            if r == self.rank:
                bases.append(self.own)
                continue
            hb = (ctypes.c_ubyte * 64)(*h.cpu().tolist())
            pp = ctypes.c_void_p()
            _cabi.call("ttb_comm_open", hb, ctypes.byref(pp))
            bases.append(pp.value)
            self._opened.append(pp.value)
        self.peers_dev = torch.tensor(bases, dtype=torch.int64, device="cuda")
        self.slots = {}
        dist.barrier()

    def slot_offset(self, key):
        idx = self.slots.get(key)
        if idx is None:
            idx = len(self.slots)
            if idx >= self.SLOTS:
                raise RuntimeError("PeerComm: out of SyncBN slots")
            self.slots[key] = idx
        return idx * self.slot_bytes

    def fits(self, n):
        return n <= self.MAX_VALUES

    def call(self, entry, key, partials, chunks, *rest):
        """Fused exchange + BatchNorm finalize (ttb_comm_bn_finalize / ttb_comm_bn_bwd_finalize): `rest` = the
        arguments of the entry point after slot_offset."""
        self._cabi.call(entry, partials.data_ptr(), chunks, self.peers_dev.data_ptr(), self.world, self.rank,
                        self.slot_offset(key), *rest)

    def all_reduce_partials(self, partials, chunks, n, key):
        """partials [chunks][n] doubles (this rank) -> [n] doubles summed over chunks and ranks (rank order)."""
        if n > self.MAX_VALUES:
            return None
        off = self.slot_offset(key)
        st = torch.cuda.current_stream().cuda_stream
        out = torch.empty((n,), dtype=torch.float64, device=partials.device)
        self._cabi.call("ttb_comm_allreduce", partials.data_ptr(), chunks, n, self.peers_dev.data_ptr(), self.world,
                        self.rank, off, out.data_ptr(), st)
        return out


def _sync_bn_active():
    return _state["initialized"] and _state["world"] > 1 and _state["sync_bn"]


def _make_stat_hook(key):
    def hook(partials, chunks, n, local_count):
        comm = _state.get("peer_comm")
        sums = None
        if comm is not None and key is not None and partials.is_cuda:
            sums = comm.all_reduce_partials(partials, chunks, n, key)
        if sums is None:  # generic path: collapse the chunks locally, then a library all-reduce (NCCL / gloo)
            sums = partials.reshape(chunks, n).sum(dim=0) if not partials.is_cuda else _collapse(partials, chunks, n)
            all_reduce_sum_(sums)
        return sums, local_count * _state["world"]  # equal shards by construction (shard_batch)
    # when the peer-memory path is up, ops.bn_* call the fused "exchange + finalize" kernels through these attributes
    comm = _state.get("peer_comm")
    hook.key = key
    hook.fused = comm if (comm is not None and key is not None and os.environ.get("TORTTO_B200_FUSED_SYNCBN", "1") != "0") else None
    return hook


def _collapse(partials, chunks, n):
    from . import _cabi
    sums = torch.empty((n,), dtype=torch.float64, device=partials.device)
    _cabi.call("ttb_bn_reduce_partials", partials.data_ptr(), chunks, n, sums.data_ptr(),
               torch.cuda.current_stream().cuda_stream)
    return sums


def bn_forward_hook(key=None):
    return _make_stat_hook(key) if _sync_bn_active() else None


def bn_backward_hook(key=None):
    return _make_stat_hook(key) if _sync_bn_active() else None


# ids of BatchNorm weight / bias Tensors whose gradient of the CURRENT backward pass was computed from all-reduced sums
# (a SyncBN layer in training mode).  Decided per step, not per module: a BatchNorm layer in eval mode (frozen
# statistics) produces rank-local dgamma / dbeta, which must go through the gradient buckets like any other parameter.
_synced_bn_step = set()


def note_synced_bn_params(ids):
    for i in ids:
        if i is not None:
            _synced_bn_step.add(i)


def enable_peer_comm():
    """Switch the SyncBN statistic exchange to the NVLink peer-memory path (needs CUDA IPC between the ranks)."""
    if _state.get("peer_comm") is None and _state["initialized"] and _state["world"] > 1 and _state["backend"] == "nccl":
        _state["peer_comm"] = PeerComm()
    return _state.get("peer_comm") is not None


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
            d._touched()
        else:
            buf = torch.from_numpy(np.array(d, copy=True))
            dist.broadcast(buf, src=src)
            d[...] = buf.numpy()


def _flat_view(t):
    """1-D view of a gradient's PHYSICAL storage (4-D arrays are NHWC in memory)."""
    if t.dim() == 4:
        return t.permute(0, 2, 3, 1).reshape(-1)
    return t.reshape(-1)


class _Bucket:
    __slots__ = ("params", "flat", "work", "averaged")

    def __init__(self, params):
        self.params = params
        self.flat = torch.cat([_flat_view(p.grad.t) for p in params])
        self.averaged = _state["backend"] == "nccl"  # NCCL divides by the world size inside the collective
        self.work = all_reduce_sum_(self.flat, async_op=True, average=True)

    def finish(self, inv_world, skip=()):
        if self.work is not None:
            self.work.wait()  # makes the current stream wait for the collective
        if not self.averaged:
            self.flat.mul_(inv_world)
        off = 0
        dsts, srcs = [], []
        for p in self.params:
            v = _flat_view(p.grad.t)
            n = v.numel()
            if id(p) not in skip:
                dsts.append(v)
                srcs.append(self.flat[off:off + n])
            off += n
        if dsts:
            torch._foreach_copy_(dsts, srcs)  # one multi-tensor launch instead of one copy per parameter


class DistributedDataParallel:
    """Wraps a Module.  Forward is unchanged; gradients are all-reduced in buckets that start during backward
    (AccumulateGrad hook) and complete in `reduce_gradients()`."""

    def __init__(self, module, bucket_mb=10, broadcast=True, overlap=True):
        self.module = module
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self.overlap = overlap
        if broadcast:
            broadcast_parameters(module)
        self._param_ids = {id(p) for p in module.parameters()}
        # BatchNorm affine parameters whose gradient came out of a SyncBN backward THIS step (see _synced_bn_step)
        self._synced_bn_params = _synced_bn_step
        self._pending, self._pending_bytes = [], 0
        self._buckets, self._seen = [], set()
        AccumulateGrad.post_hooks.append(self._on_grad_ready)

    def close(self):
        if self._on_grad_ready in AccumulateGrad.post_hooks:
            AccumulateGrad.post_hooks.remove(self._on_grad_ready)

    def __call__(self, *a, **k):
        return self.module(*a, **k)

    def __getattr__(self, name):
        return getattr(self.__dict__["module"], name)

    def parameters(self):
        return self.module.parameters()

    # ---- backward-time hook ---------------------------------------------------------------------------------
    def _on_grad_ready(self, p):
        if not (self.overlap and _state["initialized"] and _state["world"] > 1):
            return
        pid = id(p)
        if pid not in self._param_ids or pid in self._synced_bn_params:
            return
        if pid in self._seen:  # a parameter used twice in the graph: its gradient is not final yet - reduce at the end
            self._seen.add(("dirty", pid))
            return
        self._seen.add(pid)
        self._pending.append(p)
        self._pending_bytes += p.grad.nbytes
        if self._pending_bytes >= self.bucket_bytes:
            self._launch_pending()

    def _launch_pending(self):
        if self._pending:
            ops.join_wgrad()  # weight gradients still in flight on the wgrad stream are packed below
            self._buckets.append(_Bucket(self._pending))
            self._pending, self._pending_bytes = [], 0

    # ---- after backward ----------------------------------------------------------------------------------------
    def reduce_gradients(self):
        world = _state["world"]
        if not (_state["initialized"] and world > 1):
            return
        inv = 1.0 / world
        dirty = {k[1] for k in self._seen if isinstance(k, tuple)}
        self._launch_pending()
        done = set()
        for b in self._buckets:
            b.finish(inv, skip=dirty)
            done.update(id(p) for p in b.params)
        # whatever was not bucketed during backward (overlap off, re-used parameters, grads set by hand)
        rest, size = [], 0
        bn_grads = []
        for p in reversed(list(self.module.parameters())):
            if p.grad is None:
                continue
            pid = id(p)
            if pid in self._synced_bn_params:
                bn_grads.append(p.grad.t)  # identical on all ranks and already the sum over ranks (module docstring)
                continue
            if pid in done and pid not in dirty:
                continue
            rest.append(p)
            size += p.grad.nbytes
            if size >= self.bucket_bytes:
                _Bucket(rest).finish(inv)
                rest, size = [], 0
        if rest:
            _Bucket(rest).finish(inv)
        if bn_grads:
            torch._foreach_mul_(bn_grads, inv)  # one multi-tensor launch for the ~2 x (#BatchNorm layers) tiny tensors
        self._buckets, self._seen = [], set()
        _synced_bn_step.clear()
