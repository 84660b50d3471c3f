"""GPU: a CUDA-graph replay of the training step (pytortto_b200.cuda_graph.GraphedStep) is bit-identical to the
eager step - same parameters, momentum buffers and BN statistics after several steps on changing inputs."""
import numpy as np
import pytest

from gpu_util import require_gpu

pytestmark = pytest.mark.gpu


def _run(graphed_mode):
    import pytortto_b200 as tt
    from pytortto_b200.examples import make_models
    tt.set_math_mode("tf32")
    M = make_models(tt)
    tt.manual_seed(9)
    net = M["PreactResNet"](M["BasicBlock"], [1, 1, 1, 1], [32, 32, 64, 64]).cuda()
    opt = tt.optim.SGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    crit = tt.nn.NLLLoss()

    def step(x, y):
        opt.zero_grad()
        loss = crit(net(x), y)
        loss.backward()
        opt.step()
        return loss

    rng = np.random.default_rng(9)
    batches = [(rng.standard_normal((16, 3, 16, 16)).astype(np.float32), rng.integers(0, 10, 16).astype(np.int64))
               for _ in range(8)]
    dev = [(tt.tensor(x).cuda(), tt.tensor(y, dtype=np.int64).cuda()) for x, y in batches]
    losses = []
    if graphed_mode:
        # GraphedStep runs 3 eager warm-up steps on its example inputs (batch 0); the recording itself executes nothing
        g = tt.cuda_graph.GraphedStep(step, dev[0], modules=[net], warmup=3)
        for x, y in dev[1:5]:
            losses.append(g(x, y).item())
    else:
        for _ in range(3):
            step(*dev[0])
        for x, y in dev[1:5]:
            losses.append(step(x, y).item())
    return losses, net.state_dict()


def test_graphed_step_matches_eager():
    require_gpu()
    eager_losses, eager_sd = _run(False)
    graph_losses, graph_sd = _run(True)
    assert eager_losses == graph_losses
    for k in eager_sd:
        np.testing.assert_array_equal(eager_sd[k], graph_sd[k], err_msg=k)


def test_item_async_matches_item():
    """Tensor.item_async(): the pinned D2H read-back used by pipelined training loops returns what item() returns."""
    import pytortto_b200 as tt
    vals = [tt.tensor(np.array([v], dtype=np.float32)).cuda() for v in (1.5, -2.25, 3.0)]
    handles = [(v + v).item_async() for v in vals]  # all queued before the first one is read
    assert [h.get() for h in handles] == [3.0, -4.5, 6.0]
    assert tt.tensor(np.array([7.0], dtype=np.float32)).item_async().get() == 7.0  # host tensor: immediate


def test_device_prefetcher_yields_batches_in_order():
    """tt.prefetch.DevicePrefetcher: copies issued on a second stream arrive intact and in order."""
    import torch
    import pytortto_b200 as tt
    rng = np.random.default_rng(3)
    xs = [torch.from_numpy(rng.standard_normal((4, 3, 8, 8)).astype(np.float32)).pin_memory() for _ in range(5)]
    ys = [torch.from_numpy(rng.integers(0, 10, 4).astype(np.int64)).pin_memory() for _ in range(5)]
    host = [(tt.tensor(x.numpy(), copy=False), tt.tensor(y.numpy(), dtype=np.int64, copy=False)) for x, y in zip(xs, ys)]
    seen = 0
    for i, (xb, yb) in enumerate(tt.prefetch.DevicePrefetcher(host, depth=2)):
        assert xb.is_cuda and yb.is_cuda
        z = xb + xb  # consume on the current stream
        np.testing.assert_array_equal(z.data.get(), 2 * xs[i].numpy())
        np.testing.assert_array_equal(yb.data.get(), ys[i].numpy())
        seen += 1
    assert seen == 5


def test_cumulative_average_batchnorm_refuses_graph_capture():
    """BatchNorm2d(momentum=None) averages its running statistics with the HOST factor 1/num_batches_tracked, which changes
    every step: a graph replay would keep the captured step's factor.  Capturing such a step fails loudly instead."""
    require_gpu()
    import torch
    import pytortto_b200 as tt
    tt.set_math_mode("tf32")
    bn = tt.nn.BatchNorm2d(8, momentum=None).cuda()
    x = tt.tensor(np.random.default_rng(0).standard_normal((4, 8, 6, 6)).astype(np.float32)).cuda()

    def step(inp):
        return bn(inp).sum()

    with pytest.raises(RuntimeError, match="cumulative moving average"):
        tt.cuda_graph.GraphedStep(step, (x,), modules=[bn], warmup=1)
    torch.cuda.synchronize()
    bn(x)  # eager execution is unaffected
