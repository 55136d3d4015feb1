"""Per-phase timeline of the fused SGLD step (cfg 5 shape by default): for every phase, the distribution over CTAs of
(work time, barrier wait) in cycles.   python tools/sgld_trace.py [B]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bnn_b200 import ops, _lib
from bnn_b200.layout import Arch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dims = [int(v) for v in sys.argv[2].split("-")] if len(sys.argv) > 2 else [90, 512, 512, 512, 1]
dev, rows = "cuda", 200_000
arch = Arch(dims, "tanh")
X, y = torch.randn(rows, dims[0], device=dev), torch.randn(rows, dims[-1], device=dev)
chain = ops.SgldChain(arch, B, dev, 2e-2, 1e-1, 1e-2, 1e-1, seed=1)
net = [(torch.randn(o, i) * (2.0 / (i + o)) ** 0.5, torch.zeros(o)) for i, o in zip(dims[:-1], dims[1:])]
chain.set_state(arch.pack([net])[0], torch.zeros(dims[-1]))
chain.bind(X, y)
L = _lib.lib()
perm = ops.sgld_perm(1, 0, rows, 50 * B, dev)
chain.run(perm, 0, 50)
torch.cuda.synchronize()
L.bnn_sgld_trace(1, None, None, None)
chain.run(perm, 0, 50)
torch.cuda.synchronize()
grid, ph = C.c_int32(0), C.c_int32(0)
buf = np.zeros(1024 * 64, dtype=np.int64)
L.bnn_sgld_trace(1, buf.ctypes.data_as(C.c_void_p), C.byref(grid), C.byref(ph))
grid, ph = grid.value, ph.value
t = buf[: grid * ph * 2].reshape(grid, ph, 2)
nl = ph // 2
names = [f"F{l}" for l in range(nl)] + [f"B{l}" for l in range(nl - 1, -1, -1)]
print(f"grid {grid}, {ph} phases; cycles per CTA: work = barrier-entry minus previous barrier-exit, wait = barrier exit minus entry")
tot = 0
for k in range(ph):
    wait = t[:, k, 1] - t[:, k, 0]
    work = t[:, k, 0] - t[:, k - 1, 1] if k > 0 else None
    line = f"{names[k]:4s} wait min/med/max {wait.min():7d} {int(np.median(wait)):7d} {wait.max():7d}"
    if work is not None:
        line += f"   work min/med/max {work.min():7d} {int(np.median(work)):7d} {work.max():7d}"
    print(line)
pr = buf[grid * ph * 2: grid * ph * 2 + ph * 8].reshape(ph, 8)
print("CTA 1, first job of each phase: phase-start->job-start | prologue issue | K loop | reduction | epilogue+rest -> barrier entry")
for k in range(ph):
    a = pr[k]
    if a[0] == 0:
        print(f"{names[k]:4s} (no job)"); continue
    print(f"{names[k]:4s} {a[0] - a[5]:7d} | {a[1] - a[0]:7d} | {a[2] - a[1]:7d} | {a[3] - a[2]:7d} | {a[4] - a[3]:7d}")
span = t[:, ph - 1, 1] - t[:, 0, 0]
print("F0-entry .. last barrier exit (cycles), median over CTAs:", int(np.median(span)))
