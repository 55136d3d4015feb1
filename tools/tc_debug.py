import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import random_nets
from oracle import bnn_oracle as O
from bnn_b200 import ops
from bnn_b200.layout import Arch

def run(dims, act, S, N, reps):
    nets = random_nets(dims, S, seed=5)
    X = torch.randn(N, dims[0], generator=torch.Generator().manual_seed(1))
    arch = Arch(dims, act)
    bank = arch.pack(nets).cuda(); Xd = X.cuda()
    ref = O.sample_predict(nets, X, act, nout=dims[-1])
    nbad = 0
    for rep in range(reps):
        o = ops.forward(arch, Xd, bank, want_pred=True, want_moments=True, precision="tf32x3")
        torch.cuda.synchronize()
        pred = o["pred"].cpu()
        err = ((pred - ref).abs() / ref.abs().max()).squeeze(-1)
        if float(err.max()) > 1e-5:
            nbad += 1
            bad = (err > 1e-5).nonzero()
            ss = sorted(set(int(b[0]) for b in bad)); nn = sorted(set(int(b[1]) for b in bad))
            print(f"  rep {rep}: max {float(err.max()):.2e} #bad {len(bad)} samples {ss[:10]} points n={len(nn)} [{nn[0]}..{nn[-1]}] tiles128 {sorted(set(n // 128 for n in nn))[:10]}")
            s = ss[0]
            rows = [int(b[1]) for b in bad if int(b[0]) == s]
            print(f"     sample {s}: bad rows mod 128: {sorted(set(r % 128 for r in rows))[:40]}")
    print(f"{dims} {act} S={S} N={N}: {nbad}/{reps} bad launches")

reps = int(os.environ.get("REPS", 30))
which = os.environ.get("CASE", "all")
if which in ("all", "a"): run([9, 100, 100, 1], "tanh", 40, 4573, reps)
if which in ("all", "b"): run([16, 200, 200, 1], "tanh", 9, 70000, reps)
if which in ("all", "c"): run([16, 200, 200, 1], "relu", 9, 70000, reps)
