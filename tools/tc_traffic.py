"""DRAM traffic probe for the resident tensor-core forward: S samples x 1M points of cfg 4, moments only.
Run plain for the time, under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
-k regex:fwd_tc_kernel` for the bytes (BNN_TC_L2PERSIST=0 switches the L2 set-aside for the Welford state off)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bnn_b200 import ops
from bnn_b200.layout import Arch

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
arch = Arch([16, 200, 200, 1], "relu")
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(N, 16, generator=g, device="cuda")
bank = (0.1 * torch.randn(S, arch.stride, generator=g, device="cuda")) * arch.valid_mask().cuda()
ops.forward(arch, X, bank, precision="tf32_bf16")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = ops.forward(arch, X, bank, precision="tf32_bf16")
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"S={S} N={N} persist={os.environ.get('BNN_TC_L2PERSIST', '1')}: {ms:.2f} ms  {S * N / ms / 1e6:.3f}e9 evals/s  mean[0]={float(out['mean'][0]):.6f}")
