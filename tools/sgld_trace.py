"""Job timeline of the fused SGLD step (cfg 5 shape by default): per job kind, how long the jobs of the launch's last step
took, how long they waited for their inputs, and how busy the CTAs were.   python tools/sgld_trace.py [B] [dims]"""
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
grid, recs = C.c_int32(0), C.c_int32(0)
buf = np.zeros(1024 * 96 * 12, dtype=np.int64)
L.bnn_sgld_trace(1, buf.ctypes.data_as(C.c_void_p), C.byref(grid), C.byref(recs))
grid, recs = grid.value, recs.value
t = buf[: grid * recs * 12].reshape(grid, recs, 12)
ok = t[:, :, 3] > 0
code, grab, ready, done, ggrab, gdone = (t[..., i][ok] for i in range(6))
probes = t[..., 6:10][ok]
probes2 = t[..., 10:12][ok]
cta = np.broadcast_to(np.arange(grid)[:, None], ok.shape)[ok]
it, typ, lay = code >> 32, (code >> 24) & 0xff, (code >> 16) & 0xff
last = it.max()
names = {0: "FWD", 1: "HEAD", 2: "DH", 3: "DW", 4: "LOGP", 5: "LOSS"}
print(f"grid {grid}; jobs recorded {ok.sum()}; steps {it.min()}..{last}")
for step in (last - 1, last):
    m = it == step
    if not m.any():
        continue
    t0 = ggrab[m].min()
    print(f"step {step}: span {gdone[m].max() - t0} ns (first grab .. last done), {m.sum()} jobs")
    print("  kind  layer  jobs   run med/max (cyc)  idle-before-start med/max (cyc)  first grab .. last done (ns from the step's first grab)")
    for ty in range(6):
        for l in sorted(set(lay[m & (typ == ty)])):
            q = m & (typ == ty) & (lay == l)
            run = done[q] - ready[q]
            wait = np.where(probes[q][:, 3] > 0, ready[q] - probes[q][:, 3], 0)     # compute warps idle since the previous job ended
            print(f"  {names[ty]:5s} {l:3d} {q.sum():6d}   {int(np.median(run)):7d} {run.max():7d}      {int(np.median(wait)):7d} {wait.max():7d}"
                  f"          {ggrab[q].min() - t0:7d} .. {gdone[q].max() - t0:7d}")
    busy = np.zeros(grid)
    for c in range(grid):
        q = m & (cta == c)
        busy[c] = (done[q] - ready[q]).sum()
    print(f"  CTA busy cycles (running jobs) min/med/max: {int(busy.min())} {int(np.median(busy))} {int(busy.max())}")
# step period: distance between the first FWD grabs of the last two steps
a, b = ggrab[(it == last - 1) & (typ == 0)], ggrab[(it == last) & (typ == 0)]
if len(a) and len(b):
    print("step period (first FWD grab to first FWD grab):", b.min() - a.min(), "ns")

q = (it == last) & (typ == 1)
print("HEAD jobs of the last step: start -> operands staged | y + head loop | dz of the last hidden layer -> done")
for r, pz, d in zip(ready[q], probes[q], done[q]):
    print("  ", pz[0] - r, pz[1] - pz[0], d - pz[1])
m = (it == last) & (probes[:, 3] > 0)
idle = (ready - probes[:, 3])[m]
print(f"last step: compute-warp idle time between jobs: median {int(np.median(idle))}, mean {int(idle.mean())}, total per CTA {int(idle.sum() / grid)} cycles;"
      f" scheduler grab -> inputs complete: median {int(np.median((probes[:, 2] - grab)[m]))}")
# does a job run faster when the CTA's previous job had the same kind (same code, warm instruction cache)?
order = np.lexsort((grab, cta))
prev_same = np.zeros(len(code), dtype=bool)
for a_, b_ in zip(order[:-1], order[1:]):
    if cta[a_] == cta[b_]:
        prev_same[b_] = typ[a_] == typ[b_]
run_all = done - ready
for ty in (0, 2, 3):
    for same in (True, False):
        q = (typ == ty) & (prev_same == same) & (lay != 0)
        if q.any():
            print(f"{names[ty]} (layers >= 1), previous job on the CTA same kind = {same}: n = {q.sum()}, run median {int(np.median(run_all[q]))}")

print("tile jobs of the last step, medians (cycles): start -> core entry | prologue issue | K loop | slice reduction | epilogue -> done")
for ty, lsel, nm in ((0, 0, "FWD 0"), (0, 1, "FWD 1+"), (2, 1, "DH"), (3, 0, "DW 0"), (3, 1, "DW 1+")):
    q = (it == last) & (typ == ty) & ((lay == 0) if lsel == 0 else (lay > 0))
    if not q.any():
        continue
    p0, p1, p2, p3 = probes[q][:, 0], probes[q][:, 1], probes2[q][:, 0], probes2[q][:, 1]
    med = lambda v: int(np.median(v))
    print(f"  {nm:7s} {med(p0 - ready[q]):6d} | {med(p1 - p0):6d} | {med(p2 - p1):6d} | {med(p3 - p2):6d} | {med(done[q] - p3):6d}")
