"""All BASELINE.json predictive configs through the DROP-IN classes (sample + predict_mv), next to the reference's own
CPU loop (oracle port) on the same inputs.  Prints a markdown table + one JSON line.   python tools/bench_configs.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn as nn
from oracle import bnn_oracle as O
from helpers import max_rel
O.warmup()
torch.set_num_threads(os.cpu_count() or 1)


def gpu_time(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return sorted(ts)[len(ts) // 2]


def cpu_time(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts)


rows = []

# cfg 1: BNN_SGDMC, 13-50-1 ReLU, S = 100 chain samples, N = 51 (bostonHousing test split size) and 455
from bnn_b200.BNN_SGDMC import BNN_SGDMC
from bnn_b200.bank import SampleBank
for N in (51, 455):
    m = BNN_SGDMC(13, nn.ReLU(), [50], nout=1, conf={})
    base = m.arch.pack([m.nn.nn])[0]
    bank = (base[None] + 0.05 * torch.randn(100, m.arch.stride)) * m.arch.valid_mask()
    m.nns = SampleBank(m.arch, bank.cuda())
    X = torch.randn(N, 13)
    Xd = X.cuda()
    nets = [m.arch.unpack(bank[s]) for s in range(100)]
    tg = gpu_time(lambda: m.predict_mv(Xd, m.nns))
    tc = cpu_time(lambda: O.predict_mv(O.sample_predict(nets, X, "relu", nout=1), N))
    err = max_rel(m.predict_mv(Xd, m.nns)[0].cpu(), O.predict_mv(O.sample_predict(nets, X, "relu", nout=1), N)[0])
    rows.append((f"cfg1 SGDMC 13-50-1 relu, S=100, N={N}", 100 * N, tg, tc, err, "fp32"))

# cfg 2: BNN_BBB, 1-50-50-1 tanh, 1000 reparameterised draws per predict, N = 1000 grid (sampling INCLUDED)
from bnn_b200.BNN_BBB import BNN_BBB
m = BNN_BBB(1, nn.Tanh(), [50, 50], conf={})
X = torch.linspace(-3, 3, 1000).view(-1, 1); Xd = X.cuda()
def bbb_gpu():
    nns = m.sample(1000); return m.predict_mv(Xd, nns)
m.precision = "fp32"
gls = m.nn.gaussian_layers()
def bbb_cpu():
    nets = []
    for _ in range(1000):
        nets.append([O.bbb_split(O.bbb_rsample(l.mu.detach().cpu(), l.rho.detach().cpu(), torch.randn(l.mu.shape))) for l in gls])
    return O.predict_mv(O.sample_predict(nets, X, "tanh"), 1000)
tg, tc = gpu_time(bbb_gpu), cpu_time(bbb_cpu, reps=1)
rows.append(("cfg2 BBB 1-50-50-1 tanh, S=1000 draws + predict, N=1000 (fp32)", 1000 * 1000, tg, tc, float("nan"), "fp32"))
m.precision = "tf32_bf16"
rows.append(("cfg2 BBB, same (tf32_bf16, tensor cores)", 1000 * 1000, gpu_time(bbb_gpu), tc, float("nan"), "tf32_bf16"))

# cfg 3: BNN_Dropout / BNN_CDropout, 9-100-100-1 tanh, 1000 MC-dropout passes, N = 4573 (protein test split size)
from bnn_b200.BNN_Dropout import BNN_Dropout
from bnn_b200.BNN_CDropout import BNN_CDropout
X = torch.randn(4573, 9); Xd = X.cuda()
for cls, name in ((BNN_Dropout, "Dropout p=0.05"), (BNN_CDropout, "CDropout")):
    m = cls(9, nn.Tanh(), [100, 100], conf={})
    m.precision = "fp32"
    def do_gpu():
        nns = m.sample(1000); return m.predict_mv(Xd, nns)
    nns = m.sample(64)
    def do_cpu():   # the reference's loop on 64 materialised nets (scaled to 1000 below)
        nets = [m.arch.unpack(nns.packed(i).cpu()) for i in range(64)]
        return O.predict_mv(O.sample_predict(nets, X, "tanh"), 4573)
    tg, tc = gpu_time(do_gpu), cpu_time(do_cpu, reps=1) * (1000 / 64)
    rows.append((f"cfg3 {name} 9-100-100-1 tanh, S=1000 masks + predict, N=4573 (fp32)", 1000 * 4573, tg, tc, float("nan"), "fp32"))
    m.precision = "tf32_bf16"
    rows.append((f"cfg3 {name}, same (tf32_bf16, tensor cores)", 1000 * 4573, gpu_time(do_gpu), tc, float("nan"), "tf32_bf16"))

# cfg 4 slice: 16-200-200-1 relu, S = 1250 on one GPU, N = 100k slice (full N = 1M is bench.py); tensor-core path
from bnn_b200 import ops
from bnn_b200.layout import Arch
arch = Arch([16, 200, 200, 1], "relu")
g = torch.Generator().manual_seed(0)
bank = (0.1 * torch.randn(64, arch.stride, generator=g)) * arch.valid_mask()
X = torch.randn(100_000, 16, generator=g); Xd, bd = X.cuda(), bank.cuda()
nets = [arch.unpack(bank[s]) for s in range(8)]
for prec in ("tf32_bf16", "tf32x3", "fp32"):
    tg = gpu_time(lambda: ops.forward(arch, Xd, bd, precision=prec), reps=10)
    tc = cpu_time(lambda: O.predict_mv(O.sample_predict(nets, X, "relu", nout=1), 100_000), reps=1) * (64 / 8)
    err = max_rel(ops.forward(arch, Xd, bd[:8].contiguous(), precision=prec)["mean"].cpu(),
                  O.predict_mv(O.sample_predict(nets, X, "relu", nout=1), 100_000)[0])
    rows.append((f"cfg4 slice 16-200-200-1 relu, S=64, N=100k ({prec})", 64 * 100_000, tg, tc, err, prec))

print("| config | evals | GPU ms | GPU evals/s | CPU ms (%d thr) | CPU evals/s | speed-up | max-norm err of mean |" % (os.cpu_count() or 1))
print("|---|---|---|---|---|---|---|---|")
out = []
for name, ev, tg, tc, err, prec in rows:
    print(f"| {name} | {ev:,} | {tg * 1e3:.3f} | {ev / tg:.3e} | {tc * 1e3:.1f} | {ev / tc:.3e} | {tc / tg:.0f}x | {err:.1e} |")
    out.append({"config": name, "evals": ev, "gpu_ms": tg * 1e3, "cpu_ms": tc * 1e3, "speedup": tc / tg, "err": err})
print(json.dumps(out))
