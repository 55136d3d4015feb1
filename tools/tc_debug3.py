import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import random_nets
from oracle import bnn_oracle as O
O.warmup()
from bnn_b200 import ops
from bnn_b200.layout import Arch
dims, act, S, N = [9, 100, 100, 1], "tanh", 40, 4573
prec = os.environ.get("PREC", "tf32x3")
nets = random_nets(dims, S, seed=5)
X = torch.randn(N, dims[0], generator=torch.Generator().manual_seed(1))
arch = Arch(dims, act)
ref = O.sample_predict(nets, X, act, nout=1)          # CPU first, like tc_check does it AFTER; here BEFORE the launch
Xd, bank = X.cuda(), arch.pack(nets).cuda()
o = ops.forward(arch, Xd, bank, want_pred=True, want_moments=True, precision=prec)
torch.cuda.synchronize()
refd = ref.cuda()
err_dev = float(((o["pred"] - refd).abs() / refd.abs().max()).max())     # computed on the device, before any D2H of pred
pa = o["pred"].cpu()
pb = o["pred"].clone().cpu()
err_a = float(((pa - ref).abs() / ref.abs().max()).max())
err_b = float(((pb - ref).abs() / ref.abs().max()).max())
ref2 = O.sample_predict(nets, X, act, nout=1)          # CPU reference recomputed AFTER the launch
print(f"{prec}: err_dev {err_dev:.2e} err_d2h {err_a:.2e} err_d2h_clone {err_b:.2e}  ref recomputed equal: {torch.equal(ref, ref2)}  max|ref-ref2| {float((ref-ref2).abs().max()):.2e}")
