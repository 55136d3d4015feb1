"""One launch of each streaming kernel at a bandwidth-bound size, for ncu:
    ncu --set full --clock-control none -k regex:'psgld_kernel|moments_kernel|masks_kernel|metrics_kernel|apply_masks_kernel' -o gpurun_out/r02_streaming python tools/prof_streaming.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bnn_b200 import ops
from bnn_b200.layout import Arch

dev = "cuda"
n = 64 << 20
th, g, V = torch.randn(n, device=dev), torch.randn(n, device=dev), torch.rand(n, device=dev) + 0.1
ops.psgld_step(th, g, V, 1e-3, seed=1, step=3)                      # (c) 20 B/param
n5 = 572_420
ops.psgld_step(th[:n5], g[:n5], V[:n5], 1e-3, seed=1, step=4)       # cfg 5 size (latency bound)
del th, g, V
pred = torch.randn(1000, 1_000_000, device=dev)
mean, var = ops.moments(pred)                                       # (e) 4*S*M read
y = torch.randn(1_000_000, device=dev)
ops.metrics(mean, var, y)                                           # 12 B/point
del pred
a = Arch([9, 100, 100, 1], "tanh")
cfg = ops.MaskConfig("bernoulli", [True, True, True], p_drop=[0.05] * 3, seed=3)
m = ops.dropout_masks(a, cfg, 200_000, dev)                         # (d) 4*S*209 written
cfg = ops.MaskConfig("concrete", [True, True, True], p_drop=[0.1, 0.5, 0.7], seed=3)
m2 = ops.dropout_masks(a, cfg, 200_000, dev)
base = torch.randn(a.stride, device=dev)
W = ops.apply_masks(a, [True, True, True], base, m[:20_000].contiguous())
torch.cuda.synchronize()
print("ok")
