"""torchrun worker: k-GPU data-parallel run on a global batch == single-process run on the same batch.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_parity.py
Prints 'DP_PARITY_OK worst=<rel>' on rank 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pytortto_b200 as tt
from pytortto_b200 import distributed as dist
from pytortto_b200.examples import make_models

mode = os.environ.get("DP_MATH", "fp32")
tol = 2e-4 if mode == "fp32" else 0.3
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
tt.set_math_mode(mode)
M = make_models(tt)
rng = np.random.default_rng(5)
GB = 32
x = rng.standard_normal((GB, 3, 16, 16)).astype(np.float32)
lab = rng.integers(0, 10, GB).astype(np.int64)


def run(ddp_mode):
    tt.manual_seed(11)
    net = M["PreactResNet"](M["BasicBlock"], [1, 1, 1, 1], [32, 32, 64, 64]).cuda()
    ddp = dist.DistributedDataParallel(net, bucket_mb=0.05) if ddp_mode else None
    opt = tt.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    xs, ls = (dist.shard_batch(x, lab) if ddp_mode else (x, lab))
    losses = []
    for _ in range(2):
        opt.zero_grad()
        loss = tt.nn.NLLLoss()(net(tt.tensor(xs).cuda()), tt.tensor(ls, dtype=np.int64).cuda())
        loss.backward()
        if ddp is not None:
            ddp.reduce_gradients()
        opt.step()
        losses.append(loss.item())
    if ddp is not None:
        ddp.close()
    return losses, net.state_dict()


single_losses, single_sd = run(False)          # before the process group exists: plain single-GPU run
rank, world = dist.init_process_group("nccl")
dp_losses, dp_sd = run(True)
# the global loss is the mean of the per-rank local-mean losses
t = torch.tensor(dp_losses, device="cuda", dtype=torch.float64)
torch.distributed.all_reduce(t)
dp_global = (t / world).tolist()
worst = 0.0
for k in single_sd:
    a, b = np.asarray(dp_sd[k], np.float64), np.asarray(single_sd[k], np.float64)
    worst = max(worst, float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)))
ok = worst < tol and all(abs(a - b) < tol * max(1, abs(b)) for a, b in zip(dp_global, single_losses))
if rank == 0:
    print(f"losses single {single_losses} dp {dp_global}")
    print(("DP_PARITY_OK" if ok else "DP_PARITY_FAIL") + f" world={world} mode={mode} worst={worst:.3e}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
