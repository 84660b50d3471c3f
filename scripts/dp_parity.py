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
# After step 1 the forward quantities (loss, BatchNorm running statistics = the SyncBN exchange of every layer) must agree
# to 1e-5.  Parameters after two SGD steps are held to 2e-2 in fp32 mode: the two runs group the BatchNorm sums
# differently, a last-bit difference can flip one ReLU decision, and one decision moves the gradients upstream of it by
# ~1e-3 at this size (profiles/r2_relu_flip_analysis.txt); a wrong scale / missing all-reduce is an O(1) error.
tol = 2e-2 if mode == "fp32" else 0.3
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
    first_stats = None
    for it in range(2):
        opt.zero_grad()
        loss = tt.nn.NLLLoss()(net(tt.tensor(xs).cuda()), tt.tensor(ls, dtype=np.int64).cuda())
        loss.backward()
        if ddp is not None:
            ddp.reduce_gradients()
        opt.step()
        losses.append(loss.item())
        if it == 0:
            first_stats = {k: np.array(v, copy=True) for k, v in net.state_dict().items() if "running" in k}
    if ddp is not None:
        ddp.close()
    return losses, net.state_dict(), first_stats


single_losses, single_sd, single_first = run(False)          # before the process group exists: plain single-GPU run
rank, world = dist.init_process_group("nccl")
dp_losses, dp_sd, dp_first = run(True)
# the global loss is the mean of the per-rank local-mean losses
t = torch.tensor(dp_losses, device="cuda", dtype=torch.float64)
torch.distributed.all_reduce(t)
dp_global = (t / world).tolist()
worst = 0.0
for k in single_sd:
    a, b = np.asarray(dp_sd[k], np.float64), np.asarray(single_sd[k], np.float64)
    worst = max(worst, float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)))
fwd = max(float(np.abs(np.asarray(dp_first[k], np.float64) - single_first[k]).max() / max(np.abs(single_first[k]).max(), 1e-30))
          for k in single_first)
fwd = max(fwd, abs(dp_global[0] - single_losses[0]) / abs(single_losses[0]))
fwd_tol = 1e-5 if mode == "fp32" else 2e-3
ok = worst < tol and fwd < fwd_tol and all(abs(a - b) < tol * max(1, abs(b)) for a, b in zip(dp_global, single_losses))
if rank == 0:
    print(f"losses single {single_losses} dp {dp_global}")
    print(("DP_PARITY_OK" if ok else "DP_PARITY_FAIL") + f" world={world} mode={mode} forward(step 1)={fwd:.3e} (tol {fwd_tol:g}) "
          f"state after 2 steps worst={worst:.3e} (tol {tol:g})", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
