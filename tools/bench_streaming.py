"""HBM-bound kernels (b), (c), (e): achieved GB/s (ALGORITHMIC bytes / CUDA-event time) vs the measured copy peak.
    python tools/bench_streaming.py            -> prints a table and one JSON line
Algorithmic bytes: pSGLD 20 B/param/step; reparam 4*S*P written (+8P read); moments 4*S*M read + 8*M."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bnn_b200 import ops
from bnn_b200.layout import Arch

PEAK = 6506.6
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, reps=20, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()                      # > L2: evict
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []

# (c) pSGLD update: cfg5 size (launch-bound) and a large parameter vector (bandwidth-bound)
for n in (572_420, 64 << 20):
    th, g, V = torch.randn(n, device=dev), torch.randn(n, device=dev), torch.rand(n, device=dev) + 0.1
    ms = timeit(lambda: ops.psgld_step(th, g, V, 1e-3, seed=1, step=3), flush=flush if n > (8 << 20) else None)
    rows.append(("psgld_kernel", f"n={n}", 20.0 * n, ms))

# (b) reparam: cfg2 (S=1000, P=2.7k) and a cfg4-sized bank (S=1250, stride 45k)
for dims, S in (([1, 50, 50, 1], 1000), ([16, 200, 200, 1], 1250)):
    a = Arch(dims, "tanh")
    mu, rho = torch.randn(a.stride, device=dev) * a.valid_mask().to(dev), torch.full((a.stride,), -5.0, device=dev)
    ms = timeit(lambda: ops.reparam_kl(a, mu, rho, S, seed=1), flush=flush if S * a.stride * 4 > (64 << 20) else None)
    rows.append(("reparam_kernel (Philox)", f"S={S} stride={a.stride}", 4.0 * S * a.stride + 8.0 * a.stride, ms))

# (e) standalone moments over a materialised pred[S, M]
for S, M in ((1000, 4573), (1000, 1_000_000)):
    pred = torch.randn(S, M, device=dev)
    ms = timeit(lambda: ops.moments(pred), reps=10)
    rows.append(("moments_kernel + finalize", f"S={S} M={M}", 4.0 * S * M + 8.0 * M, ms))
    del pred

# (d) dropout masks: 4 * S * mask_len bytes written
a3 = Arch([9, 100, 100, 1], "tanh")
for kind, pd in (("bernoulli", [0.05] * 3), ("concrete", [0.1, 0.5, 0.7])):
    for S in (1000, 200_000):
        cfg = ops.MaskConfig(kind, [True, True, True], p_drop=pd, seed=3)
        ms = timeit(lambda: ops.dropout_masks(a3, cfg, S, dev))
        rows.append((f"masks_kernel ({kind})", f"S={S} mask_len=209", 4.0 * S * 209, ms))

out = []
print(f"{'kernel':28s} {'size':26s} {'MB':>10s} {'ms':>9s} {'GB/s':>9s} {'frac of %.0f GB/s' % PEAK:>18s}")
for k, sz, b, ms in rows:
    gbs = b / ms / 1e6
    print(f"{k:28s} {sz:26s} {b / 1e6:10.2f} {ms:9.4f} {gbs:9.1f} {gbs / PEAK:18.3f}")
    out.append({"kernel": k, "size": sz, "alg_bytes": b, "ms": ms, "gbs": gbs, "frac": gbs / PEAK})
print(json.dumps({"hbm_peak_gbs": PEAK, "rows": out}))
