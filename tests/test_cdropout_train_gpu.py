"""GPU: the fused concrete-dropout training kernel (bnn_cdropout_train_run, csrc/dropout_train.cu with CONCRETE = true; SURVEY 8
row f2).

* three steps with the reference's own torch.rand_like draws injected against the REFERENCE's network and p_logits after the same
  three steps (tests/golden/cdropout_train.npz: NN_CDropout.forward + MSELoss(sum) + reg() + backward + Adam, short last batch,
  distinct p_logits, noise_level 0.7) -- pinned;
* several launches == one launch; the epoch-end noise_level update; in-kernel Philox uniforms: determinism, training lowers the
  loss and moves the dropout probabilities;
* BNN_CDropout.train end to end on the sinc toy (the fused path is the one that runs).
Tolerance: 2e-5 relative on the parameters after three steps, 1e-5 on the per-step statistics."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from helpers import cdropout_train_case

pytestmark = pytest.mark.gpu


def make_trainer(c, seed=0):
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    arch = Arch(c["dims"], c["act"])
    tr = ops.CDropoutTrainer(arch, c["batch"], "cuda", lr=c["lr"], lscale=c["lscale"], dr=c["dr"], noise_level=c["noise_level"],
                             seed=seed)
    tr.set_state(arch.pack([c["net0"]])[0], c["p_logit0"])
    tr.bind(c["X"].cuda(), c["y"].reshape(-1, 1).cuda())
    return arch, tr


def flat_us(c, steps):
    """[len(steps), u_per_step]: per layer [batch][in], short batches padded with 0.5, each block padded to 4 floats."""
    B = c["batch"]
    rows = []
    for i in steps:
        blocks = []
        for u in c["us"][i]:
            full = torch.full((B, u.shape[1]), 0.5)
            full[: u.shape[0]] = u
            flat = full.reshape(-1)
            blocks.append(torch.cat([flat, torch.zeros((-flat.numel()) % 4)]))
        rows.append(torch.cat(blocks))
    return torch.stack(rows).cuda()


def test_three_steps_match_the_reference():
    c = cdropout_train_case()
    arch, tr = make_trainer(c)
    assert tr.u_per_step == flat_us(c, [0]).shape[1]
    stats = tr.run(c["perm"].cuda(), 0, 3, want_stats=True, uniforms=flat_us(c, range(3)))
    torch.cuda.synchronize()
    assert np.allclose(stats.cpu().numpy(), c["stats"], rtol=1e-5), (stats.cpu().numpy(), c["stats"])
    for (W, b), (W1, b1) in zip(arch.unpack(tr.W.cpu()), c["net1"]):
        assert float((W - W1).abs().max()) <= 2e-5 * float(W1.abs().max())
        assert float((b - b1).abs().max()) <= 2e-5 * float(b1.abs().max())
    assert float((tr.p_logit.cpu() - c["p_logit1"]).abs().max()) <= 2e-5 * float(c["p_logit1"].abs().max())
    assert tr.t == 3
    assert not bool(tr.W[~arch.valid_mask().cuda()].any())            # layout padding untouched
    assert float(tr.noise_level) == pytest.approx(c["noise_level"])  # no epoch end requested


def test_single_step_launches_equal_one_launch_and_epoch_end():
    c = cdropout_train_case()
    arch, a = make_trainer(c)
    _, b = make_trainer(c)
    perm = c["perm"].cuda()
    sa = a.run(perm, 0, 3, epoch_end=True, want_stats=True, uniforms=flat_us(c, range(3)))
    for i in range(3):
        b.run(perm, i * c["batch"], 1, uniforms=flat_us(c, [i]))
    torch.cuda.synchronize()
    assert torch.allclose(a.W, b.W, rtol=1e-6, atol=1e-8)
    assert torch.allclose(a.p_logit, b.p_logit, rtol=1e-6, atol=1e-8)
    assert torch.allclose(a.adam, b.adam, rtol=1e-6, atol=1e-12)
    assert torch.allclose(a.p_adam, b.p_adam, rtol=1e-6, atol=1e-12)
    # BNN_CDropout.py:134: noise_level = max(min_noise, sqrt(epoch_mse / num_train))
    want = float(np.sqrt(c["stats"][:, 0].sum() / c["X"].shape[0]))
    assert float(a.noise_level) == pytest.approx(want, rel=1e-5)
    assert float(sa[:, 0].sum()) == pytest.approx(float(c["stats"][:, 0].sum()), rel=1e-5)
    assert float(b.noise_level) == pytest.approx(c["noise_level"])


def test_philox_uniforms_are_deterministic_and_train():
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    g = torch.Generator().manual_seed(0)
    dims, B, N = [9, 100, 100, 1], 32, 256
    arch = Arch(dims, "tanh")
    X = torch.randn(N, 9, generator=g)
    y = torch.sin(X[:, 0]) + 0.5 * X[:, 1]
    net = [(torch.empty(o, i).uniform_(-1, 1, generator=g) / i ** 0.5, torch.zeros(o)) for i, o in zip(dims[:-1], dims[1:])]
    outs = []
    for seed in (1, 1, 2):
        tr = ops.CDropoutTrainer(arch, B, "cuda", lr=1e-2, seed=seed)
        tr.set_state(arch.pack([net])[0], torch.zeros(3))             # p = 0.5, BNN_CDropout.py:14-16
        tr.bind(X.cuda(), y.reshape(-1, 1).cuda())
        first = last = None
        for epoch in range(60):
            perm = torch.randperm(N, generator=torch.Generator().manual_seed(epoch)).cuda()
            mse = float(tr.run(perm, 0, N // B, epoch_end=True, want_stats=True)[:, 0].sum())
            first = mse if epoch == 0 else first
            last = mse
        assert bool(torch.isfinite(tr.W).all() and torch.isfinite(tr.p_logit).all())
        assert last < 0.5 * first, (first, last)
        assert float(tr.noise_level) == pytest.approx(np.sqrt(last / N), rel=1e-4)
        # the data term pushes the dropout probabilities below 0.5 (the reference on this problem: p_logit = -2.6, -0.10, -0.04)
        assert float(tr.p_logit[0]) < -0.5 and bool((tr.p_logit < 0).all()), tr.p_logit
        outs.append((tr.W.clone(), tr.p_logit.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert not torch.equal(outs[0][0], outs[2][0])


def test_unsupported_size_is_refused():
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    assert ops.cdropout_train_supported(Arch([9, 100, 100, 1], "tanh"), 32)
    assert not ops.cdropout_train_supported(Arch([90, 512, 512, 1], "tanh"), 32)
    with pytest.raises(ValueError):
        ops.CDropoutTrainer(Arch([90, 512, 512, 1], "tanh"), 32, "cuda")


def test_bnn_cdropout_train_uses_the_fused_kernel():
    from bnn_b200.BNN_CDropout import BNN_CDropout
    torch.manual_seed(0)
    X = torch.linspace(-3, 3, 200).view(-1, 1)
    y = torch.sin(X).view(-1)
    m = BNN_CDropout(1, nn.Tanh(), [50, 50], conf={"device": "cuda", "num_epochs": 200, "print_every": 100})
    m.train(X, y)
    assert m.trained_with == "fused"
    assert 0 < m.noise_level < 0.6
    assert all(0 < p < 0.5 for p in m.drop_probs()), m.drop_probs()
    py, pv = m.predict_mv(X, m.sample(64))
    rmse = float(((py.cpu().view(-1) - y) ** 2).mean().sqrt())
    assert rmse < 0.35, rmse
