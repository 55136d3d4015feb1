"""Small, fast cfg4-shaped launch for ncu (one wave-ish of CTAs, few samples)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bnn_b200 import ops
from bnn_b200.layout import Arch

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 8
N = int(sys.argv[3]) if len(sys.argv) > 3 else 64 * 148 * 2
dims = [int(v) for v in os.environ.get("DIMS", "16,200,200,1").split(",")]
arch = Arch(dims, "relu")
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(N, dims[0], generator=g, device="cuda")
bank = (0.1 * torch.randn(S, arch.stride, generator=g, device="cuda")) * arch.valid_mask().cuda()
for _ in range(3):
    out = ops.forward(arch, X, bank, precision=prec)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = ops.forward(arch, X, bank, precision=prec)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"dims={dims} ring={os.environ.get('BNN_TC_RING','max')} prec={prec} S={S} N={N}: {ms:.3f} ms, {S*N/ms*1e3:.3e} evals/s, {S*N*arch.flops_per_eval/ms*1e-9:.2f} TFLOP/s")
