"""Error of the tcgen05 self-test GEMM vs float64 for each pass scheme (1 = TF32, 2 = TF32 + 2 BF16 terms, 3 = 3xTF32; +16 = A in TMEM)."""
import torch, sys
sys.path.insert(0, '.')
from bnn_b200 import _lib
for K, N, passes in [(16,64,3),(16,64,2),(64,208,3),(64,208,2),(64,208,18),(64,208,19),(64,208,1)]:
    g = torch.Generator().manual_seed(K*1000+N)
    A, B = torch.randn(128, K, generator=g), torch.randn(N, K, generator=g)
    Ad, Bd, D = A.cuda(), B.cuda(), torch.zeros(128, N, device="cuda")
    _lib.check(_lib.lib().bnn_tc_selftest(Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), K, N, passes, None), "selftest")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    e = (D.cpu().double() - ref).abs()
    f32 = (A @ B.t()).double()
    print(K, N, passes, "max err/max|ref| %.3e  rms err/rms ref %.3e   (fp32 matmul: %.3e)" % (e.max()/ref.abs().max(), (e.pow(2).mean().sqrt()/ref.pow(2).mean().sqrt()), (f32-ref).abs().max()/ref.abs().max()))
