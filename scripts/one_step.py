"""Runs a few preact_resnet18 training steps (batch 256) - the command profiled under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pytortto_b200 as tt
from pytortto_b200.examples import make_models
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
tt.set_math_mode(os.environ.get("MATH", "tf32"))
M = make_models(tt)
tt.manual_seed(0)
net = M["preact_resnet18"]().cuda()
opt = tt.optim.SGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
rng = np.random.default_rng(0)
x = tt.tensor(rng.standard_normal((batch, 3, 32, 32)).astype(np.float32)).cuda()
y = tt.tensor(rng.integers(0, 10, batch).astype(np.int64), dtype=np.int64).cuda()
for i in range(steps):
    opt.zero_grad()
    loss = tt.nn.NLLLoss()(net(x), y)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", loss.item())
