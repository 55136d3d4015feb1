"""A few launches of the single-SM SGLD kernel at the cfg 1 shape (13-50-1, batch 32) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bnn_b200 import ops
from bnn_b200.layout import Arch
dev = torch.device("cuda:0")
dims, act, B, rows = [13, 50, 1], "relu", 32, 455
arch = Arch(dims, act)
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn(rows, dims[0], generator=g, device=dev)
y = torch.randn(rows, 1, generator=g, device=dev)
chain = ops.SgldChain(arch, B, dev, 2e-2, 1e-1, 1e-2, 1e-1, seed=1)
net = [(torch.randn(o, i) * (2.0 / (i + o)) ** 0.5, torch.zeros(o)) for i, o in zip(dims[:-1], dims[1:])]
chain.set_state(arch.pack([net])[0], torch.zeros(1))
chain.bind(X, y)
perm = torch.randperm(rows, device=dev)
for r in range(4):
    chain.run(perm, 0, 14)
torch.cuda.synchronize()
print("ok", bool(torch.isfinite(chain.state).all()))
