"""Kernel list of ONE eager SGLD step (cfg 5 shape) -- what the CUDA graph of BNN_SGDMC replays.  python tools/sgld_profile.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from torch.profiler import profile, ProfilerActivity
from bnn_b200.BNN_SGDMC import BNN_SGDMC
from bnn_b200 import ops
torch.manual_seed(0)
dev = torch.device("cuda")
m = BNN_SGDMC(90, nn.Tanh(), [512, 512, 512], nout=1, conf={"batch_size": 128, "device": "cuda"})
X = torch.randn(100000, 90, device=dev); y = torch.randn(100000, 1, device=dev)
m._chain_init(dev)
idx = torch.arange(128, device=dev)
step_dev = torch.zeros((), dtype=torch.long, device=dev)
def one():
    loss, grad = m._grad_into(X, y, idx, 100000)
    ops.psgld_step_dev(m._theta, grad.contiguous(), m._V, step_dev, m.lr_weight, lr_tail=m.lr_noise, n_head=m.arch.stride,
                       alpha=m.psgld_alpha, lam=m.psgld_lambda, seed=1)
for _ in range(5): one()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    one(); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time for e in ev)
print(f"{len(ev)} kernels / memcpys, {tot:.1f} us of device time in one step")
import collections
agg = collections.OrderedDict()
for e in ev:
    k = e.name[:70]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e.device_time
print("per-launch us of our kernels:", [round(e.device_time, 1) for e in ev if "bnn::" in e.name])
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{t:8.1f} us  x{n:<3d} {k}")
