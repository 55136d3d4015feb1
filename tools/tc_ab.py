import os, sys, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from bnn_b200 import ops
from bnn_b200.layout import Arch
from helpers import random_nets, max_rel
from oracle import bnn_oracle as O
O.warmup()
arch = Arch([16, 200, 200, 1], "relu")
# parity on a small case incl. ragged tiles
nets = random_nets([16, 200, 200, 1], 6, seed=1)
X = torch.randn(700, 16, generator=torch.Generator().manual_seed(2))
out = ops.forward(arch, X.cuda(), arch.pack(nets).cuda(), want_pred=True, want_moments=True, precision="tf32_bf16")
torch.cuda.synchronize()
ref = O.sample_predict(nets, X, "relu", nout=1)
print("parity pred", max_rel(out["pred"].cpu(), ref), "mean", max_rel(out["mean"].cpu(), O.predict_mv(ref, 700)[0]))
g = torch.Generator(device="cuda").manual_seed(0)
Xb = torch.randn(1_000_000, 16, generator=g, device="cuda")
bank = (0.1 * torch.randn(128, arch.stride, generator=g, device="cuda")) * arch.valid_mask().cuda()
for prec in ("tf32_bf16", "tf32x3", "tf32"):
    for _ in range(2):
        ops.forward(arch, Xb, bank, precision=prec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.forward(arch, Xb, bank, precision=prec)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{prec}: {ms:.2f} ms  {128e6 / ms / 1e6:.3f}e9 evals/s  V1={os.environ.get('BNN_TC_V1')}")
