"""Per-instruction cost of tcgen05.mma variants on one SM (debug probe, see bnn_tc_mma_bench)."""
import sys
import torch
sys.path.insert(0, ".")
from bnn_b200 import _lib
names = ["tf32 SS K=8", "tf32 TS K=8", "bf16 SS K=16", "bf16 TS K=16", "mixed stage (2 tf32 + 2 bf16, TS)", "3xTF32 stage (TS)"]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for N in (208, 256, 112, 64, 32, 16):
    for mode in range(6):
        for iters in (240,):
            _lib.check(_lib.lib().bnn_tc_mma_bench(mode, N, iters, out.data_ptr(), None), "bench")
            torch.cuda.synchronize()
            _lib.check(_lib.lib().bnn_tc_mma_bench(mode, N, iters, out.data_ptr(), None), "bench")
            torch.cuda.synchronize()
            a, b = out.tolist()
            print(f"N={N:3d} {names[mode]:36s} iters={iters}: issue {a/iters:6.1f} clk/MMA, complete {b/iters:6.1f} clk/MMA")
