import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from bnn_b200.BNN_SGDMC import BNN_SGDMC
N, B = int(sys.argv[1]), int(sys.argv[3]); dims = [int(v) for v in sys.argv[2].split(",")]
dev = torch.device("cuda", 0)
X = torch.randn(N, dims[0], device=dev); y = torch.randn(N, 1, device=dev)
m = BNN_SGDMC(dims[0], nn.Tanh(), dims[1:-1], nout=1, conf={"device": "cuda:0", "seed": 1, "batch_size": B})
m._chain_init(dev); m._perm, m._perm_pos = None, 0
m._build_graph(X, y, N)
torch.cuda.synchronize()
print("built; off", int(m._g_off), "step", int(m._g_step), "perm max", int(m._g_perm.max()), "theta finite", bool(torch.isfinite(m._theta).all()), "V min", float(m._V.min()))
gen = torch.Generator(device=dev).manual_seed(1)
perm = torch.randperm(N, generator=gen, device=dev)
print("perm dtype", perm.dtype, "min", int(perm.min()), "max", int(perm.max()), "unique", int(perm.unique().numel()))
m._g_perm.copy_(perm); m._g_off.zero_()
torch.cuda.synchronize()
m._graph.replay(); torch.cuda.synchronize()
print("replay 1 ok; off", int(m._g_off), "step", int(m._g_step), "loss", float(m._g_loss))

m._perm, m._perm_pos = None, 0
class G:
    def __init__(self, g): self.g, self.n = g, 0
    def replay(self):
        torch.cuda.synchronize()
        print("  replay", self.n, "off", int(m._g_off), "step", int(m._g_step), "perm[off:off+3]", m._g_perm[int(m._g_off):int(m._g_off)+3].tolist(), "perm max", int(m._g_perm.max()), "numel", m._g_perm.numel(), flush=True)
        self.g.replay(); torch.cuda.synchronize(); self.n += 1
m._graph = G(m._graph)
m.sgld_steps(20, X, y, gen); torch.cuda.synchronize(); print("sgld_steps ok", int(m._g_off), int(m._g_step), m._t)
