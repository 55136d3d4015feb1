"""GPU parity of kernel (a)+(e): fused forward + moments vs the oracle and the golden fixtures.
Tolerances: FP32 path, 1e-5 relative (north_star); var bound per helpers.var_bound."""
import pytest
import torch

from helpers import golden, flat_to_nets, random_nets, max_rel, var_bound
from oracle import bnn_oracle as O

pytestmark = pytest.mark.gpu


def run_cuda(dims, act, nets, X, **kw):
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    arch = Arch(dims, act)
    bank = arch.pack(nets).cuda()
    out = ops.forward(arch, X.cuda().contiguous(), bank, want_pred=True, want_moments=True, **kw)
    torch.cuda.synchronize()
    return {k: v.cpu() for k, v in out.items() if not k.startswith("_")}


def check_against(out, pred_ref, N, depth=2):
    S = pred_ref.shape[0]
    pred_ref = pred_ref.reshape(S, N, -1)
    assert out["pred"].shape == pred_ref.shape
    assert max_rel(out["pred"], pred_ref) <= 1e-5
    mean_ref, var_ref = O.predict_mv(pred_ref, N)
    assert max_rel(out["mean"], mean_ref) <= 1e-5
    if S > 1:
        dv = (out["var"].double() - var_ref.double()).abs()
        vb = var_bound(pred_ref, depth=depth)
        assert bool((dv <= vb).all()), float((dv / vb).max())
    else:
        assert bool(torch.isnan(out["var"]).all())      # torch.var of one sample is NaN (BNN.py:37)


def test_golden_sgdmc_boston():
    g = golden("sgdmc_boston.npz")
    dims = [int(d) for d in g["dims"]]
    nets = flat_to_nets(g["W"], dims)
    X = torch.from_numpy(g["X"])
    out = run_cuda(dims, "relu", nets, X)
    check_against(out, torch.from_numpy(g["pred"]), X.shape[0])
    assert max_rel(out["mean"], g["mean"]) <= 1e-5


def test_golden_sgdmc_multiout():
    g = golden("sgdmc_multiout.npz")
    dims = [int(d) for d in g["dims"]]
    nets = flat_to_nets(g["W"], dims)
    X = torch.from_numpy(g["X"])
    out = run_cuda(dims, "tanh", nets, X)
    check_against(out, torch.from_numpy(g["pred"]), X.shape[0])


@pytest.mark.parametrize("dims,act,S,N", [
    ([13, 50, 1], "relu", 100, 51),          # cfg 1
    ([1, 50, 50, 1], "tanh", 64, 333),       # cfg 2 shape
    ([9, 100, 100, 1], "tanh", 24, 1000),    # cfg 3 shape
    ([16, 200, 200, 1], "relu", 6, 700),     # cfg 4 shape
    ([16, 200, 200, 1], "tanh", 3, 129),
    ([90, 512, 512, 512, 1], "tanh", 2, 100),  # cfg 5 shape (TN=32 path, 2 rounds)
    ([7, 33, 5], "sigmoid", 5, 65),          # ragged everything
    ([3, 1, 2], "identity", 4, 17),          # width-1 hidden layer
    ([1000, 17, 3], "relu", 2, 40),          # wide input
    ([5, 1024, 2], "relu", 2, 20),           # widest layer (TN=16 path)
    ([4, 8, 8, 8, 8, 8, 8, 8, 2], "tanh", 3, 50),  # 8 Linear layers
    ([2, 20, 1], "relu", 1, 10),             # S = 1 -> var NaN
])
def test_random_shapes(dims, act, S, N):
    nets = random_nets(dims, S, seed=sum(dims) + S)
    X = torch.randn(N, dims[0], generator=torch.Generator().manual_seed(N))
    ref = O.sample_predict(nets, X, act, nout=dims[-1])
    out = run_cuda(dims, act, nets, X)
    check_against(out, ref, N, depth=len(dims) - 1)


def test_sample_groups_and_partials():
    """Few points, many samples: the S axis is split over CTAs and Chan-merged; result must not care."""
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    dims, S, N = [6, 40, 1], 700, 9
    nets = random_nets(dims, S, seed=3)
    X = torch.randn(N, 6, generator=torch.Generator().manual_seed(1))
    ref = O.sample_predict(nets, X, "tanh", nout=1)
    arch = Arch(dims, "tanh")
    out = ops.forward(arch, X.cuda(), arch.pack(nets).cuda(), want_pred=False, want_moments=True, want_partials=True)
    mean_ref, var_ref = O.predict_mv(ref, N)
    assert max_rel(out["mean"].cpu(), mean_ref) <= 1e-5
    dv = (out["var"].cpu().double() - var_ref.double()).abs()
    assert bool((dv <= var_bound(ref)).all())
    n, m64, m2 = O.shard_moments(ref.double().reshape(S, -1).numpy())
    assert max_rel(out["part_mean"].cpu().reshape(-1), m64) <= 1e-6
    assert max_rel(out["part_m2"].cpu().reshape(-1), m2) <= 1e-5


def test_variance_stress():
    """mean 5, sigma 0.01: a naive FP32 (sum, sum^2) loses ~13% of the variance."""
    from bnn_b200 import ops
    S, M = 1000, 257
    g = torch.Generator().manual_seed(0)
    pred = (5.0 + 0.01 * torch.randn(S, M, generator=g)).float()
    mean, var = ops.moments(pred.cuda())
    assert max_rel(var.cpu(), pred.var(dim=0)) <= 1e-5
    assert max_rel(mean.cpu(), pred.mean(dim=0)) <= 1e-6


@pytest.mark.parametrize("precision", ["fp32", "tf32_bf16"])
def test_forward_host_pipelined_matches_device_call(precision):
    """bnn_forward_host (host pointers in, host pointers out): banks of >= 768 samples are uploaded in chunks behind the
    evaluation and Chan-merged; smaller ones go through in one piece.  Both must agree with the device-pointer call."""
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    dims = [16, 200, 200, 1] if precision != "fp32" else [6, 24, 1]
    arch = Arch(dims, "tanh")
    g = torch.Generator().manual_seed(5)
    for S, N in ((900, 300), (40, 300)):
        bank = ((0.3 * torch.randn(S, arch.stride, generator=g)) * arch.valid_mask()).contiguous()
        X = torch.randn(N, dims[0], generator=g)
        ref = ops.forward(arch, X.cuda(), bank.cuda(), precision=precision)
        out = ops.forward_host(arch, X.pin_memory(), bank.pin_memory(), precision=precision)
        assert max_rel(out["mean"], ref["mean"].cpu()) <= 1e-6
        assert max_rel(out["var"], ref["var"].cpu()) <= 1e-5
