"""cfg 1's training job on the single-SM SGLD kernel and on the dataflow kernel (bench.py's sgld_cfg1 block) + raw step rates."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bnn_b200 import ops
from bnn_b200.layout import Arch
dev = torch.device("cuda:0")
print(json.dumps(bench.sgld_cfg1_block(dev)))
for dims, act, B in (([13, 50, 1], "relu", 32), ([13, 50, 1], "relu", 128), ([1, 50, 50, 1], "tanh", 32), ([8, 100, 1], "relu", 32), ([9, 100, 100, 1], "tanh", 32)):
    for env in ("2", "0"):
        os.environ["BNN_SGLD_SMALL"] = env
        arch = Arch(dims, act)
        rows = 100_000
        g = torch.Generator(device=dev).manual_seed(0)
        X = torch.randn(rows, dims[0], generator=g, device=dev)
        y = torch.randn(rows, 1, generator=g, device=dev)
        chain = ops.SgldChain(arch, B, dev, 2e-2, 1e-1, 1e-2, 1e-1, seed=1)
        net = [(torch.randn(o, i) * (2.0 / (i + o)) ** 0.5, torch.zeros(o)) for i, o in zip(dims[:-1], dims[1:])]
        chain.set_state(arch.pack([net])[0], torch.zeros(1))
        chain.bind(X, y)
        for n in (50,):
            perm = ops.sgld_perm(1, 0, rows, n * B, dev)
            chain.run(perm, 0, n)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(1, 3000 // n)
            e0.record()
            for r in range(reps):
                chain.run(perm, 0, n)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (reps * n)
            print(f"{'-'.join(map(str, dims))} B={B} SMALL={env}: {n:4d} steps/launch -> {us:7.2f} us/step  finite {bool(torch.isfinite(chain.state).all())}")
os.environ.pop("BNN_SGLD_SMALL", None)
