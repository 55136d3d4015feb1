"""On a failing first launch, identify which hi/lo operand of the 3xTF32 scheme was ineffective."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import random_nets
from bnn_b200 import ops
from bnn_b200.layout import Arch

def hi(x):   # tf32 round-to-nearest (ties away), like cvt.rna.tf32.f32
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)

def model(net, X, act, drop):
    f = {"tanh": torch.tanh, "relu": torch.relu}[act]
    (W1, b1), (W2, b2), (W3, b3) = net
    def mm(A, B, dropA, dropB):   # A @ B^T with 3xTF32 semantics, optionally without a lo term
        Ah, Bh = hi(A), hi(B)
        Al, Bl = A - Ah, B - Bh
        r = Ah.double() @ Bh.double().t()
        if not dropA: r = r + Al.double() @ Bh.double().t()
        if not dropB: r = r + Ah.double() @ Bl.double().t()
        return r.float()
    h1 = f(mm(X, W1, "x" in drop, "w1" in drop) + b1)
    h2 = f(mm(h1, W2, "h1" in drop, "w2" in drop) + b2)
    return (h2.double() @ W3.double().t() + b3.double()).float()

dims, act, S, N = [9, 100, 100, 1], "tanh", 40, 4573
nets = random_nets(dims, S, seed=5)
X = torch.randn(N, dims[0], generator=torch.Generator().manual_seed(1))
arch = Arch(dims, act)
import time
Xd, bank = X.cuda(), arch.pack(nets).cuda()
if os.environ.get("SYNC_FIRST"):
    torch.cuda.synchronize(); time.sleep(0.05)
o = ops.forward(arch, Xd, bank, want_pred=True, want_moments=True, precision="tf32x3")
torch.cuda.synchronize()
o2 = ops.forward(arch, Xd, bank, want_pred=True, want_moments=True, precision="tf32x3")
torch.cuda.synchronize()
print("second launch identical to first:", torch.equal(o["pred"], o2["pred"]), " max |first-second| rel", float((o["pred"]-o2["pred"]).abs().max() / o2["pred"].abs().max()))
pred = o["pred"].cpu().squeeze(-1)
full = model(nets[0], X, act, set()).squeeze(-1)
err = (pred[0] - full).abs() / full.abs().max()
bad = (err > 1e-5).nonzero().squeeze(-1)
print(f"sample 0: max err vs full model {float(err.max()):.2e}, bad rows {len(bad)}")
if len(bad):
    for drop in (["x"], ["w1"], ["h1"], ["w2"], ["x", "w1"], ["h1", "w2"]):
        m = model(nets[0], X, act, set(drop)).squeeze(-1)
        d = (pred[0][bad] - m[bad]).abs().max() / full.abs().max()
        print(f"   hypothesis drop {drop}: residual on bad rows {float(d):.2e}")
    tiles = sorted(set(int(b) // 128 for b in bad))
    for t in tiles[:3]:
        rows = torch.arange(t * 128, min(N, t * 128 + 128))
        e = ((pred[0][rows] - full[rows]).abs() / full.abs().max())
        print(f"   tile {t}: rows with err>3e-6: {int((e > 3e-6).sum())}/128; per-quarter max {[f'{float(e[q*32:(q+1)*32].max()):.1e}' for q in range(4) if len(e) >= (q+1)*32]}")
