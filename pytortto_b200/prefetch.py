"""Host -> device input pipeline (the reference's `tt.tensor(batch).cuda()` idiom is a
synchronous copy on the compute stream every step, README.md:40).

`DevicePrefetcher(batches)` wraps any iterable of tuples of HOST tensors (numpy-backed `tt.Tensor`s; keep the numpy
arrays in pinned memory - e.g. `torch.from_numpy(a).pin_memory().numpy()` - for truly asynchronous copies) and yields
the same tuples on the device.  The copy of batch i+1 (pinned H2D + the NCHW->NHWC layout kernel) is issued on a
second stream while the caller is still computing on batch i; the consumer stream only waits for the copy's event.

    for x, y in tt.prefetch.DevicePrefetcher(host_batches):
        loss = step(x, y)
"""
import collections

import torch


_copy_streams = {}  # one copy stream per device for the life of the process: the caching allocator keeps its free
                    # blocks per stream, a fresh stream per prefetcher would cudaMalloc its staging buffers again


def _copy_stream():
    dev = torch.cuda.current_device()
    st = _copy_streams.get(dev)
    if st is None:
        st = _copy_streams[dev] = torch.cuda.Stream(device=dev)
    return st


class DevicePrefetcher:
    def __init__(self, batches, depth=1):
        self.batches = batches
        self.depth = max(1, int(depth))
        self.stream = None

    def __iter__(self):
        if self.stream is None:
            self.stream = _copy_stream()
        it = iter(self.batches)
        pending = collections.deque()

        def issue():
            batch = next(it, None)
            if batch is None:
                return
            with torch.cuda.stream(self.stream):
                dev = tuple(t.cuda() for t in batch)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            pending.append((dev, ev))

        for _ in range(self.depth):
            issue()
        while pending:
            dev, ev = pending.popleft()
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            for t in dev:  # allocated on the copy stream, consumed on this one: tell the caching allocator
                if t.is_cuda:
                    t.data.t.record_stream(cur)
            issue()  # the next copy overlaps whatever the caller does with this batch
            yield dev
