"""Data-parallel layer (new work: the reference is single-process, single-device - SURVEY.md §2c, §8(e)).

One process per GPU (torchrun environment: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT), NCCL over
NVLink 5 / NVSwitch through `torch.distributed` (gloo for the CPU host-logic tests).  The global batch is sharded
by rank (`shard_batch`); two exchange steps make a k-GPU run equal the reference's single-process run on the
concatenated batch:

1. Parameter gradients.  Every rank back-propagates the mean loss of ITS shard.  The parameters are laid out once in
   static flat buckets (~10 MB, reverse parameter order = the order backward produces them); the wgrad / BatchNorm
   backward kernels write each gradient straight into its slot of the bucket (`p._grad_slot`), and
   `DistributedDataParallel`'s hook on `AccumulateGrad` (the leaf node of the autograd graph, reference
   autograd/function.py:70-93) launches NCCL's all-reduce (AVG) IN PLACE on a bucket as soon as its last gradient is
   final - from the wgrad stream, so it overlaps the rest of backward; no pack / unpack passes.  `reduce_gradients()`
   (between `loss.backward()` and `optimizer.step()`) waits for the outstanding buckets; the mean over ranks of
   local-mean gradients IS the global-batch mean gradient.

2. BatchNorm statistics (SyncBN).  BatchNorm.forward all-reduces the per-channel double sums [sum x, sum x^2]
   (2C values) and uses count = world * local count, so normalisation and the running statistics (unbiased variance
   with the GLOBAL N/(N-1), reference autograd/grad_nn.py:923-930) are those of the global batch; BatchNorm.backward
   all-reduces [sum dy, sum dy*(x-mean)] so dx follows the reference formula (:984-988) with global sums.
   dgamma / dbeta computed from those all-reduced sums are identical on every rank and equal the SUM over ranks of
   the local-loss gradients, so they skip the bucket all-reduce and only receive the 1/world scaling (decided per step:
   only for layers that really ran synced).  On one NVSwitch box (<= 8 ranks) the exchange does not go through NCCL:
   `PeerComm` / csrc/comm.cu is a one-kernel all-reduce over NVLink peer memory, fused with the BatchNorm finalisation.
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
    if backend == "nccl" and sync_bn and 1 < _state["world"] <= PeerComm.MAX_WORLD and \
            os.environ.get("TORTTO_B200_PEER_COMM", "1") != "0":
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
    MAX_VALUES = 2 * 2048   # 2C doubles, C <= 2048 (wider layers take the library all-reduce)
    MAX_WORLD = 8           # one NVSwitch box: every rank's slot has an area per source rank (csrc/comm.cu)
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
        self.slots = {}                      # key -> slot index, assigned in first-call order (the same on every rank)
        self.free = []                       # indices released by `release` (deterministic program points only)
        self.next = 0
        dist.barrier()

    def slot_index(self, key):
        """Slot of a call site, or None when the buffer is exhausted (the caller then takes the library all-reduce).
        Keys are (BatchNorm module construction index, direction): identical on every rank by construction."""
        idx = self.slots.get(key)
        if idx is None:
            if self.free:
                idx = self.free.pop()
            elif self.next < self.SLOTS:
                idx = self.next
                self.next += 1
            else:
                return None
            self.slots[key] = idx
        return idx

    def slot_offset(self, key):
        idx = self.slot_index(key)
        if idx is None:
            raise RuntimeError("PeerComm: out of SyncBN slots")
        return idx * self.slot_bytes

    def release(self, sync_ids):
        """Give the slots of these BatchNorm layers back (called at the same program point on every rank, e.g.
        DistributedDataParallel.close(), in module order - never from a finaliser, whose timing differs between ranks)."""
        for sid in sync_ids:
            for d in ('f', 'b'):
                idx = self.slots.pop((sid, d), None)
                if idx is not None:
                    self.free.append(idx)

    def fits(self, n, key=None):
        return n <= self.MAX_VALUES and (key is None or self.slot_index(key) is not None)

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
        if comm is not None and key is not None and partials.is_cuda and comm.fits(n, key):
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


class _FlatBucket:
    """A contiguous fp32 buffer holding the gradients of a fixed set of parameters.  Every parameter of the bucket owns a
    `cparray` VIEW of its slice (`p._grad_slot`, physical layout of the parameter itself: NHWC for conv weights): the
    kernels that produce parameter gradients (wgrad, bias / BatchNorm / Linear gradients) write straight into it, the
    bucket's all-reduce runs in place on the buffer, and the optimizer reads `p.grad` = the same view - no pack, no
    unpack, no copy."""
    __slots__ = ("params", "flat", "slots", "ready", "work", "dirty")

    def __init__(self, params, device):
        from .xparray import cparray
        self.params = params
        total = sum(p.data.size for p in params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.slots = {}
        off = 0
        for p in params:
            n = p.data.size
            shp = tuple(p.data.shape)
            chunk = self.flat[off:off + n]
            if len(shp) == 4:
                k, c, r, s_ = shp
                view = chunk.view(k, r, s_, c).permute(0, 3, 1, 2)  # logical (K, C, R, S) over [K][R][S][C] storage
            else:
                view = chunk.view(shp)
            slot = cparray(view)
            assert slot.t.data_ptr() == chunk.data_ptr(), "gradient slot must alias the flat buffer"
            self.slots[id(p)] = slot
            p._grad_slot = slot
            off += n
        self.ready, self.work, self.dirty = set(), None, False

    def reset(self):
        self.ready, self.work, self.dirty = set(), None, False


class DistributedDataParallel:
    """Wraps a Module.  Forward is unchanged.  Parameters are assigned once to flat gradient buckets (`bucket_mb`, reverse
    parameter order = the order backward produces them); a bucket's all-reduce (NCCL AVG, in place on the flat buffer)
    starts from the AccumulateGrad hook as soon as its last gradient exists - issued from the wgrad side stream so that
    the collective waits for the weight-gradient kernels without the main stream having to - and `reduce_gradients()`
    (between `loss.backward()` and `optimizer.step()`) only waits for the outstanding collectives."""

    def __init__(self, module, bucket_mb=10, broadcast=True, overlap=True):
        self.module = module
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self.overlap = overlap
        if broadcast:
            broadcast_parameters(module)
        params = [p for p in module.parameters() if p.requires_grad]
        self._param_ids = {id(p) for p in params}
        # BatchNorm affine parameters whose gradient came out of a SyncBN backward THIS step (see _synced_bn_step)
        self._synced_bn_params = _synced_bn_step
        self._buckets, self._bucket_of = [], {}
        cur, size = [], 0
        for p in reversed(params):
            cur.append(p)
            size += p.data.nbytes
            if size >= self.bucket_bytes:
                self._add_bucket(cur)
                cur, size = [], 0
        if cur:
            self._add_bucket(cur)
        AccumulateGrad.post_hooks.append(self._on_grad_ready)

    def _add_bucket(self, params):
        d0 = params[0].data
        dev = d0.t.device if hasattr(d0, "t") else torch.device("cpu")  # (host parameters: the gloo host-logic tests)
        b = _FlatBucket(params, dev)
        self._buckets.append(b)
        for p in params:
            self._bucket_of[id(p)] = b

    def close(self):
        if self._on_grad_ready in AccumulateGrad.post_hooks:
            AccumulateGrad.post_hooks.remove(self._on_grad_ready)
        comm = _state.get("peer_comm")
        if comm is not None:  # this module's SyncBN slots can serve the next model (same order on every rank)
            from .nn.modules import _BatchNorm
            comm.release([m._sync_id for m in self.module.modules() if isinstance(m, _BatchNorm)])
        for b in self._buckets:
            for p in b.params:
                if getattr(p, "_grad_slot", None) is b.slots[id(p)]:
                    p._grad_slot = None

    def __call__(self, *a, **k):
        return self.module(*a, **k)

    def __getattr__(self, name):
        return getattr(self.__dict__["module"], name)

    def parameters(self):
        return self.module.parameters()

    # ---- backward-time hook ---------------------------------------------------------------------------------
    def _on_grad_ready(self, p):
        if not (_state["initialized"] and _state["world"] > 1):
            return
        pid = id(p)
        b = self._bucket_of.get(pid)
        if b is None:
            return
        slot = b.slots[pid]
        if p.grad is not slot:  # produced by an op that does not write into the slot (or accumulated): move it there
            if p.grad.t.is_cuda:
                ops.join_wgrad()
            slot.t.copy_(p.grad.t)
            p.grad = slot
        if pid in b.ready or b.work is not None:  # a second gradient for the same parameter / after the bucket went out
            b.dirty = True
            return
        b.ready.add(pid)
        if self.overlap and len(b.ready) == len(b.params):
            self._launch(b)

    def _launch(self, b):
        """all-reduce the bucket in place; the collective is ordered after BOTH compute streams without stalling the main one"""
        if b.flat.is_cuda:
            side = ops._side_stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                b.work = all_reduce_sum_(b.flat, async_op=True, average=True)
            ops.mark_side_stream_used()
        else:
            b.work = all_reduce_sum_(b.flat, async_op=True, average=True)

    # ---- after backward ----------------------------------------------------------------------------------------
    def reduce_gradients(self):
        world = _state["world"]
        if not (_state["initialized"] and world > 1):
            return
        inv = 1.0 / world
        averaged = _state["backend"] == "nccl"  # NCCL divides by the world size inside the collective
        leftovers = []
        for b in self._buckets:
            if b.work is None and not b.dirty and len(b.ready) == len(b.params):
                self._launch(b)  # overlap off
            if b.work is not None:
                b.work.wait()  # makes the current stream wait for the collective
                if not averaged:
                    b.flat.mul_(inv)
            if b.work is None or b.dirty:
                # incomplete bucket (a parameter without a gradient this step) or a gradient that changed after the
                # bucket went out: reduce what exists parameter by parameter (rare path)
                leftovers.extend(p for p in b.params if p.grad is not None and (b.work is None or b.dirty))
        if leftovers:
            ops.join_wgrad()
        for p in leftovers:
            g = _flat_view(p.grad.t)
            all_reduce_sum_(g, average=True)
            if not averaged:
                g.mul_(inv)
        ops.join_wgrad()
        # dgamma / dbeta of SyncBN layers were computed from all-reduced sums: identical on every rank and equal to the SUM
        # over ranks of the local-loss gradients - the AVG all-reduce left them unchanged, they still need the 1/world
        bn_grads = [p.grad.t for b in self._buckets for p in b.params
                    if id(p) in self._synced_bn_params and p.grad is not None]
        if bn_grads:
            torch._foreach_mul_(bn_grads, inv)  # one multi-tensor launch for the ~2 x (#BatchNorm layers) tiny tensors
        for b in self._buckets:
            b.reset()
        _synced_bn_step.clear()
