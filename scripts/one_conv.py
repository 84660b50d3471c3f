"""One conv layer fwd + bwd on the tensor path (profiling target).  args: N C H W K ksize stride pad [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pytortto_b200 as tt
n, c, h, w, k, ks, s, p = [int(v) for v in sys.argv[1:9]]
iters = int(sys.argv[9]) if len(sys.argv) > 9 else 2
rng = np.random.default_rng(0)
x = tt.nn.Parameter(tt.tensor(rng.standard_normal((n, c, h, w)).astype(np.float32)).cuda())
conv = tt.nn.Conv2d(c, k, ks, stride=s, padding=p, bias=False).cuda()
for _ in range(iters):
    x.grad = None; conv.weight.grad = None
    y = conv(x)
    y.backward(tt.tensor(np.ones(y.shape, np.float32)).cuda())
torch.cuda.synchronize()
print("ok", y.shape)
