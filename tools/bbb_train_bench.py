"""Fused Bayes-by-Backprop training kernel (bnn_bbb_train_run) at the cfg 2 shape (1-50-50-1 tanh, batch 32, 1000 rows):
steps/s of the fused kernel (one launch per epoch), of the same recipe in plain torch on the GPU (the fallback path of
bnn_b200.BNN_BBB.train) and of the reference's recipe on the CPU (oracle port).   python tools/bbb_train_bench.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from bnn_b200 import ops
from bnn_b200.layout import Arch
from bnn_b200.BNN_BBB import BNN_BBB

dims, B, N = [1, 50, 50, 1], 32, 1000
arch = Arch(dims, "tanh")
g = torch.Generator().manual_seed(0)
X = torch.rand(N, 1, generator=g) * 6 - 3
y = torch.sin(X)
mus = [torch.empty(o, i + 1).uniform_(-0.3, 0.3, generator=g) for i, o in zip(dims[:-1], dims[1:])]
rhos = [torch.full((o, i + 1), -5.0) for i, o in zip(dims[:-1], dims[1:])]
tr = ops.BbbTrainer(arch, B, "cuda", seed=1)
tr.set_state(arch.pack_bias_last(mus), arch.pack_bias_last(rhos))
tr.bind(X.cuda(), y.cuda())
steps = (N + B - 1) // B
perms = [torch.randperm(N, generator=g).cuda() for _ in range(8)]
for p in perms[:2]:
    tr.run(p, 0, steps)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 200
e0.record()
for r in range(reps):
    tr.run(perms[r % 8], 0, steps)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (reps * steps)
out = {"fused_us_per_step": us, "fused_steps_per_s": 1e6 / us}
print(f"fused kernel: {us:.2f} us/step = {1e6 / us:.0f} steps/s ({steps} steps per launch)")

for dev in ("cuda", "cpu"):
    m = BNN_BBB(1, nn.Tanh(), [50, 50], conf={"device": dev, "num_epochs": 3 if dev == "cuda" else 2, "print_every": 10 ** 6,
                                               "fused_train": False}) if dev == "cuda" else None
    if dev == "cuda":
        m.train(X, y.view(-1))                      # warm-up epoch(s) included; time a second call
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m.train(X, y.view(-1))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        n = 3 * steps
    else:
        from oracle import bnn_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        batches = [torch.randperm(N, generator=g)[:B] for _ in range(200)]
        eps = [[torch.randn(B, o) for o in dims[1:]] for _ in range(200)]
        t0 = time.perf_counter()
        O.bbb_train_steps(mus, rhos, X, y, batches, "tanh", eps)
        dt = time.perf_counter() - t0
        n = 200
    out[f"torch_{dev}_steps_per_s"] = n / dt
    print(f"torch on {dev}: {dt / n * 1e6:.1f} us/step = {n / dt:.0f} steps/s")
print(json.dumps(out))
