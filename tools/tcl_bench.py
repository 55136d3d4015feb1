"""cfg 5's predictive pass (90-512-512-512-1 tanh) on the streaming tensor-core kernel vs the FP32 SIMT kernel.
    python tools/tcl_bench.py [S] [N]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bnn_b200 import ops
from bnn_b200.layout import Arch
S = int(sys.argv[1]) if len(sys.argv) > 1 else 50
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
dims = [int(v) for v in sys.argv[3].split("-")] if len(sys.argv) > 3 else [90, 512, 512, 512, 1]
arch = Arch(dims, "tanh")
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(N, dims[0], generator=g, device="cuda")
bank = (0.05 * torch.randn(S, arch.stride, generator=g, device="cuda")) * arch.valid_mask().cuda()
res = {}
for prec in ("tf32_bf16", "fp32"):
    for _ in range(2):
        out = ops.forward(arch, X, bank, precision=prec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        out = ops.forward(arch, X, bank, precision=prec)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = arch.flops_per_eval * S * N / (ms * 1e-3) / 1e12
    res[prec] = out["mean"].clone()
    print(f"{'-'.join(map(str, dims))} S={S} N={N} {prec}: {ms:.2f} ms  {S * N / ms / 1e6:.4f}e9 evals/s  {tf:.1f} TFLOP/s algorithmic")
print("max |mean_tc - mean_fp32| / max|mean|:", float((res["tf32_bf16"] - res["fp32"]).abs().max() / res["fp32"].abs().max()))
