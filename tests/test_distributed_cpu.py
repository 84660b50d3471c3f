"""CPU, gloo, world_size 2: host-side logic of the data-parallel layer (pytortto_b200/distributed.py) -
sharding, bucketed gradient averaging through the AccumulateGrad hook, SyncBN statistic exchange, parameter
broadcast.  The device arrays are cparrays wrapping CPU torch tensors (the wrapper itself needs no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        import pytortto_b200 as tt
        from pytortto_b200 import distributed as dist
        from pytortto_b200.autograd.function import AccumulateGrad
        from pytortto_b200.xparray import cparray
        from oracle import tortto_oracle as O

        assert dist.init_process_group("gloo") == (rank, world)
        # ---- shard_batch -------------------------------------------------------------------------------------
        rng = np.random.default_rng(0)
        x = rng.standard_normal((8, 6, 5, 5)).astype(np.float32)
        lab = np.arange(8)
        xs, ls = dist.shard_batch(x, lab)
        assert xs.shape[0] == 4 and np.array_equal(ls, lab[rank * 4:(rank + 1) * 4])
        with pytest.raises(ValueError):
            dist.shard_batch(np.zeros((7, 2)))

        # ---- SyncBN statistics: all-reduced double sums == global-batch statistics (oracle) ----------------------
        hook = dist.bn_forward_hook()
        assert hook is not None
        # two per-chunk partial rows per rank, like the stats kernel emits them: [chunks][2C]
        halves = [xs[:2], xs[2:]]
        partials = torch.from_numpy(np.stack([np.concatenate([h.sum((0, 2, 3), dtype=np.float64),
                                                              (h.astype(np.float64) ** 2).sum((0, 2, 3))]) for h in halves]))
        sums, count = hook(partials, 2, 12, xs.shape[0] * 25)
        assert count == 8 * 25 and tuple(sums.shape) == (12,)
        sums = sums.reshape(2, 6)
        mean = sums[0].numpy() / count
        var = sums[1].numpy() / count - mean ** 2
        _, rm, rv, saved = O.batch_norm_forward(x, None, None, np.zeros(6, np.float32), np.ones(6, np.float32), True, 0.1, 1e-5)
        np.testing.assert_allclose(mean, saved[0].reshape(-1), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(var + 1e-5, saved[1].reshape(-1), rtol=1e-5)
        np.testing.assert_allclose(0.9 + 0.1 * var * count / (count - 1), rv, rtol=1e-5)  # global N/(N-1)

        # ---- broadcast + bucketed gradient averaging through the AccumulateGrad hook -----------------------------
        tt.manual_seed(100 + rank)  # different init per rank on purpose
        # net[1]: a BatchNorm whose backward ran the SyncBN exchange this step (its dgamma / dbeta are already global
        # sums: scaled by 1/world only); net[2]: a BatchNorm in eval mode (frozen statistics, no exchange) - its affine
        # gradients are rank-local and must be averaged through the buckets like any other parameter
        net = tt.nn.Sequential(tt.nn.Conv2d(3, 4, 3, bias=True), tt.nn.BatchNorm2d(4), tt.nn.BatchNorm2d(4),
                               tt.nn.Linear(5, 2))
        ddp = dist.DistributedDataParallel(net, bucket_mb=1e-4)  # tiny buckets: several collectives in flight
        ref = [np.array(p.data, copy=True) for p in net.parameters()]
        gathered = [torch.zeros(sum(r.size for r in ref)) for _ in range(world)]
        torch.distributed.all_gather(gathered, torch.from_numpy(np.concatenate([r.ravel() for r in ref])))
        assert torch.equal(gathered[0], gathered[1]), "broadcast_parameters must equalise the ranks"

        params = list(net.parameters())
        dist.note_synced_bn_params((id(net[1].weight), id(net[1].bias)))  # what BatchNorm.backward does under SyncBN
        bn_ids = {id(net[1].weight), id(net[1].bias)}
        local = []
        for i, p in enumerate(reversed(params)):  # backward order
            # (SyncBN affine gradients come out of all-reduced sums: bit-identical on every rank)
            g = np.random.default_rng(i if id(p) in bn_ids else 10 * rank + i).standard_normal(p.shape).astype(np.float32)
            local.append(g)
            acc = AccumulateGrad()
            acc.variable = p
            t = torch.from_numpy(g.copy())
            if t.dim() == 4:
                t = t.contiguous(memory_format=torch.channels_last)
            acc.grad = [cparray(t)]
            acc.apply(acc.grad[0])  # fires the DDP hook exactly like the engine does
        ddp.reduce_gradients()
        for b in ddp._buckets:  # gradients live in the flat buckets: p.grad IS the parameter's view of the bucket buffer
            for p in b.params:
                assert p.grad is b.slots[id(p)]
        for i, p in enumerate(reversed(params)):
            other = np.random.default_rng(10 * (1 - rank) + i).standard_normal(p.shape).astype(np.float32)
            want = local[i] / world if id(p) in bn_ids else (local[i] + other) / world
            np.testing.assert_allclose(p.grad.t.numpy(), want, rtol=1e-6, atol=1e-7)
        assert not dist._synced_bn_step, "the per-step set must be cleared by reduce_gradients"
        ddp.close()
        assert not AccumulateGrad.post_hooks
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
        raise


def test_data_parallel_host_logic_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in results), results


def test_single_process_is_a_noop():
    from pytortto_b200 import distributed as dist
    assert not dist.is_initialized() and dist.get_world_size() == 1
    assert dist.bn_forward_hook() is None and dist.bn_backward_hook() is None
    a = np.arange(6).reshape(3, 2)
    assert dist.shard_batch(a) is not None and np.array_equal(dist.shard_batch(a), a)
    t = torch.ones(3)
    assert dist.all_reduce_sum_(t) is None and torch.equal(t, torch.ones(3))
