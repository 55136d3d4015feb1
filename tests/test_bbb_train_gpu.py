"""GPU: the fused Bayes-by-Backprop training kernel (bnn_bbb_train_run, csrc/bbb_train.cu).

* three steps with the reference's own noise injected against the REFERENCE's parameters after the same three steps
  (tests/golden/bbb_train.npz: model.loss + backward + Adam, short last batch included) -- pinned;
* several launches == one launch; in-kernel Philox noise: statistics and determinism;
* BNN_BBB.train end to end on the sinc toy (the fused path is the one that runs).
Tolerance: 2e-5 relative on (mu, rho) after three steps (FP32 sums in a different order than ATen's), 1e-5 on mse / KL."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from helpers import bbb_train_case
from oracle import bnn_oracle as O

pytestmark = pytest.mark.gpu


def make_trainer(c, seed=0):
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    arch = Arch(c["dims"], c["act"])
    tr = ops.BbbTrainer(arch, c["batch"], "cuda", lr=c["lr"], kl_factor=c["kl_factor"], weight_std=c["weight_std"], seed=seed)
    tr.set_state(arch.pack_bias_last(c["mu0"]), arch.pack_bias_last(c["rho0"]))
    tr.bind(c["X"].cuda(), c["y"].reshape(-1, 1).cuda())
    return arch, tr


def unpack_bias_last(arch, vec):
    return [arch.layer_view(vec, l)[:, : arch.ins[l] + 1].cpu() for l in range(arch.n_linear)]


def flat_eps(c, steps):
    return torch.stack([torch.cat([c["eps"][i][l].reshape(-1) if c["eps"][i][l].shape[0] == c["batch"] else
                                   torch.cat([c["eps"][i][l], torch.zeros(c["batch"] - c["eps"][i][l].shape[0], c["eps"][i][l].shape[1])]).reshape(-1)
                                   for l in range(len(c["dims"]) - 1)]) for i in steps]).cuda()


def test_three_steps_match_the_reference():
    c = bbb_train_case()
    arch, tr = make_trainer(c)
    mse, kl = tr.run(c["perm"].cuda(), 0, 3, want_stats=True, eps=flat_eps(c, range(3)))
    torch.cuda.synchronize()
    assert np.allclose(mse.cpu().numpy(), c["mse"], rtol=1e-5), (mse.cpu().numpy(), c["mse"])
    assert np.allclose(kl.cpu().numpy(), c["kl"], rtol=1e-5), (kl.cpu().numpy(), c["kl"])
    for got, ref in zip(unpack_bias_last(arch, tr.mu), c["mu1"]):
        assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    for got, ref in zip(unpack_bias_last(arch, tr.rho), c["rho1"]):
        assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    assert tr.t == 3
    assert not bool(tr.mu[~arch.valid_mask().cuda()].any())          # layout padding untouched


def test_single_step_launches_equal_one_launch():
    c = bbb_train_case()
    arch, a = make_trainer(c)
    _, b = make_trainer(c)
    perm = c["perm"].cuda()
    a.run(perm, 0, 3, eps=flat_eps(c, range(3)))
    for i in range(3):
        b.run(perm, i * c["batch"], 1, eps=flat_eps(c, [i]))
    torch.cuda.synchronize()
    assert torch.allclose(a.mu, b.mu, rtol=1e-6, atol=1e-8) and torch.allclose(a.rho, b.rho, rtol=1e-6, atol=1e-8)
    assert torch.allclose(a.adam, b.adam, rtol=1e-6, atol=1e-12)


def test_philox_noise_is_deterministic_and_trains():
    """In-kernel noise: same seed -> same trajectory, different seed -> different; 300 steps on a learnable target lower the MSE."""
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    g = torch.Generator().manual_seed(0)
    dims, B, N = [1, 50, 50, 1], 32, 256
    arch = Arch(dims, "tanh")
    X = (torch.rand(N, 1, generator=g) * 6 - 3)
    y = torch.sin(X)
    mus = [torch.empty(o, i + 1).uniform_(-0.3, 0.3, generator=g) for i, o in zip(dims[:-1], dims[1:])]
    rhos = [torch.full((o, i + 1), -5.0) for i, o in zip(dims[:-1], dims[1:])]
    outs = []
    for seed in (1, 1, 2):
        tr = ops.BbbTrainer(arch, B, "cuda", lr=1e-2, kl_factor=1e-3, weight_std=0.1, seed=seed)
        tr.set_state(arch.pack_bias_last(mus), arch.pack_bias_last(rhos))
        tr.bind(X.cuda(), y.cuda())
        first = last = None
        for epoch in range(40):
            perm = torch.randperm(N, generator=torch.Generator().manual_seed(epoch)).cuda()
            mse, _ = tr.run(perm, 0, N // B, want_stats=True)
            if epoch == 0:
                first = float(mse.mean())
            last = float(mse.mean())
        torch.cuda.synchronize()
        assert bool(torch.isfinite(tr.mu).all() and torch.isfinite(tr.rho).all())
        assert last < 0.5 * first, (first, last)
        outs.append(tr.mu.clone())
    assert torch.equal(outs[0], outs[1])
    assert not torch.equal(outs[0], outs[2])


def test_unsupported_size_is_refused():
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    assert ops.bbb_train_supported(Arch([1, 50, 50, 1], "tanh"), 32)
    assert not ops.bbb_train_supported(Arch([90, 512, 512, 1], "tanh"), 32)
    with pytest.raises(ValueError):
        ops.BbbTrainer(Arch([90, 512, 512, 1], "tanh"), 32, "cuda")


def test_bnn_bbb_train_uses_the_fused_kernel():
    from bnn_b200.BNN_BBB import BNN_BBB
    torch.manual_seed(0)
    X = torch.linspace(-3, 3, 200).view(-1, 1)
    y = torch.sin(X).view(-1)
    m = BNN_BBB(1, nn.Tanh(), [50, 50], conf={"device": "cuda", "num_epochs": 150, "print_every": 1000, "kl_factor": 1e-4})
    m.train(X, y)
    assert m.trained_with == "fused"
    py, pv = m.predict_mv(X, m.sample(64))
    rmse = float(((py.cpu().view(-1) - y) ** 2).mean().sqrt())
    assert rmse < 0.25, rmse
