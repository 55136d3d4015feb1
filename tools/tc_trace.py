"""Dump the in-kernel timeline of the tensor-core forward (BNN_TC_TRACE=1): per K-step timestamps of the MMA issuer
and of one producer warp of each WG-A set."""
import os, sys
os.environ.setdefault("BNN_TC_TRACE", "7")   # bit 0 MMA per-stage events, bit 1 WG-A, bit 2 WG-B; tile ends always
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the probes are compiled in only in the trace build (python -m bnn_b200.build --trace)
os.environ["BNN_B200_LIB"] = os.path.join(ROOT, "bnn_b200", "libbnn_b200_trace.so")
import numpy as np, torch
from bnn_b200 import ops, _lib
from bnn_b200.layout import Arch
dims = [int(v) for v in os.environ.get("DIMS", "16,200,200,1").split(",")]
arch = Arch(dims, "relu")
S, N = 4, int(os.environ.get("NPTS", 75776))
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(N, dims[0], generator=g, device="cuda")
bank = (0.1 * torch.randn(S, arch.stride, generator=g, device="cuda")) * arch.valid_mask().cuda()
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
if os.environ.get("MASKED"):     # dropout family: ONE shared network + per-sample external masks (cfg 3)
    S = int(os.environ.get("NS", 64))
    masks = (torch.rand(S, sum(dims[:-1]), generator=g, device="cuda") > 0.05).float()
    cfg = ops.MaskConfig("external", [True] * (len(dims) - 1), mask=masks)
    run = lambda: ops.forward(arch, X, bank[0].contiguous(), S=S, mask=cfg, precision=prec)
else:
    run = lambda: ops.forward(arch, X, bank, precision=prec)
run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
buf = np.zeros(5 * 4096 * 2, dtype=np.uint64)
_lib.check(_lib.lib().bnn_tc_trace_dump(buf.ctypes.data), "dump")
buf = buf.reshape(5, 4096, 2)
names = {4: "mma:tf32_issued", 5: "mma:bf16_issued", 36: "B:issued", 33: "B:ldwait", 34: "B:ldok", 35: "B:reduced", 30: "B:wait_d2full", 31: "B:got_d2full", 32: "B:drained", 20: "start", 21: "tile_end", 1: "mma:wait_full", 2: "mma:got_full", 3: "mma:committed", 10: "ld_start", 11: "ld_done", 12: "wait_empty", 13: "got_empty", 14: "signalled"}
t0 = min(int(buf[w, 0, 1]) for w in range(5) if buf[w, 0, 1])
lo, hi = int(os.environ.get("IT0", 30)), int(os.environ.get("IT1", 42))
ev = []
tl, th = int(os.environ.get("T0", 0)), int(os.environ.get("T1", 0))
n_epi = 0
for w in range(5):
    for tag_it, clk in buf[w]:
        if clk == 0: break
        tag, it = int(tag_it) >> 32, int(tag_it) & 0xffffffff
        if tag in (33, 34, 35, 36):
            if os.environ.get('EPI2') and w == 3 and tl <= it < th: ev.append((int(clk) - t0, 'B0', names[tag], it))
        elif tag >= 30:
            if tl <= it < th: ev.append((int(clk) - t0, ["MMA", "A0", "A1", "B0", "B1"][w], names[tag], it))
        elif tag in (20, 21):
            if tl <= it < th: ev.append((int(clk) - t0, "MMA", names[tag], it))
        elif lo <= it < hi: ev.append((int(clk) - t0, ["MMA", "A0", "A1", "B0", "B1"][w], names[tag], it))
for e in sorted(ev): print(f"{e[0]:8d}  {e[1]:4s} {e[2]:14s} it={e[3]}")
# per K-step period of the MMA thread
m = [(int(t) & 0xffffffff, int(c)) for t, c in buf[0] if c and (int(t) >> 32) == 3]
d = np.diff([c for _, c in m])
tiles = [int(c) for t, c in buf[0] if c and (int(t) >> 32) == 21]
start = [int(c) for t, c in buf[0] if c and (int(t) >> 32) == 20][0]
print("tile periods (clk):", np.diff(tiles)[:12], " total clk %d over %.3f ms (incl. prep+finalize) -> >= %.0f MHz" % (tiles[-1] - start, ms, (tiles[-1] - start) / ms / 1e3))
if len(d): print("MMA commit-to-commit period: median %.0f  p10 %.0f  p90 %.0f  (n=%d)" % (np.median(d), np.percentile(d, 10), np.percentile(d, 90), len(d)))
