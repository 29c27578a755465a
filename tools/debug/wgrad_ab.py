"""A/B the weight-gradient kernels through a whole branch backward (development helper): run with TFSWA_WGRAD=mma / unset."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import tfswa_unet_b200 as T
from tfswa_unet_b200 import ops
from oracle import tfswa_oracle as O

T.set_precision("bf16")
m = T.TemporalSequenceAttention(32, 8)
sd = m.state_dict(); O.randomize_state_(sd, 3, 1.0); m.load_state_dict(sd); m = m.eval().cuda()
g = torch.Generator().manual_seed(1)
x = torch.randn((2, 32, 37, 21), generator=g).cuda().requires_grad_(True)
calls = []
orig = ops.linear_wgrad
def spy(x_, g_, **kw):
    calls.append((tuple(x_.shape), x_.stride(), tuple(g_.shape), g_.stride(), kw.get("prologue"), kw.get("row_stats") is not None and tuple(kw["row_stats"].shape), kw.get("want_bias")))
    return orig(x_, g_, **kw)
ops.linear_wgrad = spy
y = m(x)
w = torch.randn(y.shape, generator=g).cuda()
(y.float() * w).sum().backward()
torch.cuda.synchronize()
for c in calls: print(c)
torch.save({k: p.grad.cpu() for k, p in m.named_parameters()}, sys.argv[1])
