"""Three launches of 50 fused SGLD steps at the cfg 5 shape (the ncu target: `-k regex:sgld_step_kernel -s 1 -c 1`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bnn_b200 import ops
from bnn_b200.layout import Arch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dims, dev, rows = [90, 512, 512, 512, 1], "cuda", 200_000
arch = Arch(dims, "tanh")
X, y = torch.randn(rows, dims[0], device=dev), torch.randn(rows, dims[-1], device=dev)
chain = ops.SgldChain(arch, B, dev, 2e-2, 1e-1, 1e-2, 1e-1, seed=1)
net = [(torch.randn(o, i) * (2.0 / (i + o)) ** 0.5, torch.zeros(o)) for i, o in zip(dims[:-1], dims[1:])]
chain.set_state(arch.pack([net])[0], torch.zeros(dims[-1]))
chain.bind(X, y)
for r in range(3):
    perm = ops.sgld_perm(1, r, rows, 50 * B, dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    chain.run(perm, 0, 50)
    e1.record()
    torch.cuda.synchronize()
    print(f"launch {r}: {e0.elapsed_time(e1) * 1e3 / 50:.2f} us/step")
chain.check()
