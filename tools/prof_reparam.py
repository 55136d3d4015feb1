"""One reparam (BBB draw) launch on a cfg4-sized bank, for ncu:  ncu --set full -k regex:reparam_kernel python tools/prof_reparam.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bnn_b200 import ops
from bnn_b200.layout import Arch
a = Arch([16, 200, 200, 1], "tanh")
mu, rho = torch.randn(a.stride, device="cuda") * a.valid_mask().cuda(), torch.full((a.stride,), -5.0, device="cuda")
for _ in range(3):
    out = ops.reparam_kl(a, mu, rho, 1250, seed=1)
torch.cuda.synchronize()
print("ok")
