"""A few launches of the three fused training kernels (rows f2) for ncu: bbb_train_kernel, dropout_train_kernel<false/true>.
    python tools/train_kernels_ncu.py            # 2 launches of each, one 32-step run per launch"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bnn_b200 import ops
from bnn_b200.layout import Arch

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
N, B, steps = 4096, 32, 32
# MC dropout / concrete dropout: cfg 3 shape
dims = [9, 100, 100, 1]
arch = Arch(dims, "relu")
X = torch.randn(N, 9, generator=g)
y = torch.sin(X[:, 0]) + 0.5 * X[:, 1]
net = [(torch.empty(o, i).uniform_(-1, 1, generator=g) / i ** 0.5, torch.zeros(o)) for i, o in zip(dims[:-1], dims[1:])]
perm = torch.randperm(N, generator=g).to(dev)
do = ops.DropoutTrainer(arch, B, dev, p_drop=0.05, seed=1)
do.set_state(arch.pack([net])[0])
do.bind(X.to(dev), y.reshape(-1, 1).to(dev))
cd = ops.CDropoutTrainer(arch, B, dev, seed=1)
cd.set_state(arch.pack([net])[0], torch.zeros(3))
cd.bind(X.to(dev), y.reshape(-1, 1).to(dev))
# Bayes by Backprop: cfg 2 shape
dims2 = [1, 50, 50, 1]
arch2 = Arch(dims2, "tanh")
X2 = torch.rand(N, 1, generator=g) * 6 - 3
mus = [torch.empty(o, i + 1).uniform_(-0.3, 0.3, generator=g) for i, o in zip(dims2[:-1], dims2[1:])]
rhos = [torch.full((o, i + 1), -5.0) for i, o in zip(dims2[:-1], dims2[1:])]
bb = ops.BbbTrainer(arch2, B, dev, seed=1)
bb.set_state(arch2.pack_bias_last(mus), arch2.pack_bias_last(rhos))
bb.bind(X2.to(dev), torch.sin(X2).to(dev))
for r in range(2):
    do.run(perm, r * steps * B, steps)
    cd.run(perm, r * steps * B, steps)
    bb.run(perm, r * steps * B, steps)
torch.cuda.synchronize()
print("ok", bool(torch.isfinite(do.W).all() and torch.isfinite(cd.W).all() and torch.isfinite(bb.mu).all()))
