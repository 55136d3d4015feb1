import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from bnn_b200.BNN_SGDMC import BNN_SGDMC
N = int(sys.argv[1]); dims = [int(v) for v in sys.argv[2].split(",")]
dev = torch.device("cuda", 0)
X = torch.randn(N, dims[0], device=dev); y = torch.randn(N, 1, device=dev)
m = BNN_SGDMC(dims[0], nn.Tanh(), dims[1:-1], nout=1, conf={"device": "cuda:0", "seed": 1, "batch_size": 128})
m._chain_init(dev); m._perm, m._perm_pos = None, 0
gen = torch.Generator(device=dev).manual_seed(1)
m.sgld_steps(50, X, y, gen); torch.cuda.synchronize(); print("warm ok")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); m.sgld_steps(2000, X, y, gen); e1.record(); torch.cuda.synchronize()
print(f"N={N} dims={dims}: {2000 / (e0.elapsed_time(e1) * 1e-3):.0f} steps/s, loss {float(m._g_loss):.1f}")
