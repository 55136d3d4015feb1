"""cfg 1 predictive call (13-50-1, S=100, N=51): where do its ~80 us go?  Host issue time per call vs device time per call
(back-to-back calls), and the device time of the library call alone (C ABI through ops.forward) vs through the class."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from bnn_b200 import ops
from bnn_b200.BNN_SGDMC import BNN_SGDMC
from bnn_b200.bank import SampleBank
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = BNN_SGDMC(13, nn.ReLU(), [50], nout=1, conf={"device": "cuda:0"})
base = m.arch.pack([m.nn.nn])[0]
m.nns = SampleBank(m.arch, ((base[None] + 0.05 * torch.randn(100, m.arch.stride)) * m.arch.valid_mask()).to(dev))
X = torch.randn(51, 13, device=dev)
W = m.nns.weights


def measure(fn, n=300):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    return host, e0.elapsed_time(e1) * 1e3 / n


for name, fn in (("class predict_mv", lambda: m.predict_mv(X, m.nns)),
                 ("ops.forward fp32", lambda: ops.forward(m.arch, X, W, precision="fp32")),
                 ("ops.forward auto", lambda: ops.forward(m.arch, X, W, precision="auto")),
                 ("torch.empty x3", lambda: (torch.empty(51, 1, device=dev), torch.empty(51, 1, device=dev), torch.empty(4096, dtype=torch.uint8, device=dev)))):
    h, d = measure(fn)
    print(f"{name:20s}: host {h:7.1f} us/call issue, device {d:7.1f} us/call back to back")
# single isolated call, device time between events right around it
ts = []
for _ in range(50):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.forward(m.arch, X, W, precision="fp32"); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
print("isolated ops.forward fp32: median", ts[len(ts) // 2], "us, min", ts[0])
