"""Fused SGLD step (bnn_sgld_run): steps/s of one chain at cfg 5 (90-512-512-512-1 tanh, batch 128) and cfg 1 shapes,
for a few launch lengths.   python tools/sgld_bench.py [rows]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bnn_b200 import ops
from bnn_b200.layout import Arch

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = "cuda"
out = []
for dims, act, B in (([90, 512, 512, 512, 1], "tanh", 128), ([90, 512, 512, 512, 1], "tanh", 32), ([13, 50, 1], "relu", 32),
                     ([9, 100, 100, 1], "tanh", 128)):
    arch = Arch(dims, act)
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.randn(rows, dims[0], generator=g, device=dev)
    y = torch.randn(rows, dims[-1], generator=g, device=dev)
    chain = ops.SgldChain(arch, B, dev, 2e-2, 1e-1, 1e-2, 1e-1, seed=1)
    net = [(torch.randn(o, i) * (2.0 / (i + o)) ** 0.5, torch.zeros(o)) for i, o in zip(dims[:-1], dims[1:])]
    chain.set_state(arch.pack([net])[0], torch.zeros(dims[-1]))
    chain.bind(X, y)
    for n in (1, 50, 500):
        perm = ops.sgld_perm(1, 0, rows, n * B, dev)
        chain.run(perm, 0, n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(1, 2000 // n)
        e0.record()
        for r in range(reps):
            perm = ops.sgld_perm(1, r + 1, rows, n * B, dev)
            chain.run(perm, 0, n)
        e1.record()
        torch.cuda.synchronize()
        chain.check()
        us = e0.elapsed_time(e1) * 1e3 / (reps * n)
        print(f"{'-'.join(map(str, dims))} {act} B={B}: {n:4d} steps/launch -> {us:8.2f} us/step = {1e6 / us:9.0f} steps/s  (finite: {bool(torch.isfinite(chain.state).all())})")
        out.append({"dims": dims, "B": B, "steps_per_launch": n, "us_per_step": us})
    del X, y
print(json.dumps(out))
