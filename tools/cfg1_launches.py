"""Five cfg 1 predictive calls (13-50-1, S=100, N=51, FP32 SIMT path) for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bnn_b200 import ops
from bnn_b200.layout import Arch
dev = torch.device("cuda:0")
arch = Arch([13, 50, 1], "relu")
g = torch.Generator().manual_seed(0)
W = (torch.randn(100, arch.stride, generator=g) * 0.3 * arch.valid_mask()).to(dev)
X = torch.randn(51, 13, generator=g).to(dev)
for _ in range(5):
    out = ops.forward(arch, X, W, precision="fp32")
torch.cuda.synchronize()
print("ok", bool(torch.isfinite(out["mean"]).all()))
