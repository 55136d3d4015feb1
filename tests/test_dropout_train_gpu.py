"""GPU: the fused MC-dropout training kernel (bnn_dropout_train_run, csrc/dropout_train.cu).

* three steps with the reference's own F.dropout draws injected against the REFERENCE's network after the same three steps
  (tests/golden/dropout_train.npz: NN_Dropout.forward + MSELoss(sum) + backward + Adam with the two weight-decay groups, short
  last batch included) -- pinned;
* several launches == one launch; in-kernel Philox masks: keep rate, determinism, training lowers the loss;
* BNN_Dropout.train end to end on the sinc toy (the fused path is the one that runs).
Tolerance: 2e-5 relative on the parameters after three steps (FP32 sums in a different order than ATen's), 1e-5 on the loss."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from helpers import dropout_train_case

pytestmark = pytest.mark.gpu


def make_trainer(c, seed=0):
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    arch = Arch(c["dims"], c["act"])
    tr = ops.DropoutTrainer(arch, c["batch"], "cuda", lr=c["lr"], l2_reg=c["l2_reg"], p_drop=c["p_drop"], seed=seed)
    tr.set_state(arch.pack([c["net0"]])[0])
    tr.bind(c["X"].cuda(), c["y"].reshape(-1, 1).cuda())
    return arch, tr


def flat_masks(c, steps):
    """[len(steps), mask_per_step]: per masked layer [batch][in], short batches zero-padded, each block padded to 4 floats."""
    B = c["batch"]
    rows = []
    for i in steps:
        blocks = []
        for k in c["keeps"][i]:
            full = torch.zeros(B, k.shape[1])
            full[: k.shape[0]] = k
            flat = full.reshape(-1)
            blocks.append(torch.cat([flat, torch.zeros((-flat.numel()) % 4)]))
        rows.append(torch.cat(blocks))
    return torch.stack(rows).cuda()


def test_three_steps_match_the_reference():
    c = dropout_train_case()
    arch, tr = make_trainer(c)
    assert tr.mask_per_step == flat_masks(c, [0]).shape[1]
    loss = tr.run(c["perm"].cuda(), 0, 3, want_loss=True, masks=flat_masks(c, range(3)))
    torch.cuda.synchronize()
    assert np.allclose(loss.cpu().numpy(), c["loss"], rtol=1e-5), (loss.cpu().numpy(), c["loss"])
    for (W, b), (W1, b1) in zip(arch.unpack(tr.W.cpu()), c["net1"]):
        assert float((W - W1).abs().max()) <= 2e-5 * float(W1.abs().max())
        assert float((b - b1).abs().max()) <= 2e-5 * float(b1.abs().max())
    assert tr.t == 3
    assert not bool(tr.W[~arch.valid_mask().cuda()].any())            # layout padding untouched


def test_single_step_launches_equal_one_launch():
    c = dropout_train_case()
    arch, a = make_trainer(c)
    _, b = make_trainer(c)
    perm = c["perm"].cuda()
    a.run(perm, 0, 3, masks=flat_masks(c, range(3)))
    for i in range(3):
        b.run(perm, i * c["batch"], 1, masks=flat_masks(c, [i]))
    torch.cuda.synchronize()
    assert torch.allclose(a.W, b.W, rtol=1e-6, atol=1e-8)
    assert torch.allclose(a.adam, b.adam, rtol=1e-6, atol=1e-12)


def test_philox_masks_are_deterministic_and_train():
    """In-kernel masks: same seed -> same trajectory, different seed -> different; 40 epochs on a learnable target lower the loss;
    with p_drop = 0 the kernel equals plain Adam regression (the oracle with all-ones keep factors)."""
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    from oracle import bnn_oracle as O
    g = torch.Generator().manual_seed(0)
    dims, B, N = [9, 100, 100, 1], 32, 256
    arch = Arch(dims, "relu")
    X = torch.randn(N, 9, generator=g)
    y = torch.sin(X[:, 0]) + 0.5 * X[:, 1]
    net = [(torch.empty(o, i).uniform_(-1, 1, generator=g) / i ** 0.5, torch.zeros(o)) for i, o in zip(dims[:-1], dims[1:])]
    outs = []
    for seed in (1, 1, 2):
        tr = ops.DropoutTrainer(arch, B, "cuda", lr=1e-2, l2_reg=1e-6, p_drop=0.05, seed=seed)
        tr.set_state(arch.pack([net])[0])
        tr.bind(X.cuda(), y.reshape(-1, 1).cuda())
        first = last = None
        for epoch in range(40):
            perm = torch.randperm(N, generator=torch.Generator().manual_seed(epoch)).cuda()
            loss = float(tr.run(perm, 0, N // B, want_loss=True).sum())
            first = loss if epoch == 0 else first
            last = loss
        assert bool(torch.isfinite(tr.W).all())
        assert last < 0.3 * first, (first, last)
        outs.append(tr.W.clone())
    assert torch.equal(outs[0], outs[1])
    assert not torch.equal(outs[0], outs[2])
    # p_drop = 0: no randomness left -- three steps against the oracle
    tr = ops.DropoutTrainer(arch, B, "cuda", lr=1e-2, l2_reg=1e-4, p_drop=0.0, seed=3)
    tr.set_state(arch.pack([net])[0])
    tr.bind(X.cuda(), y.reshape(-1, 1).cuda())
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(7))
    loss = tr.run(perm.cuda(), 0, 3, want_loss=True)
    batches = [perm[i * B:(i + 1) * B] for i in range(3)]
    keeps = [[torch.ones(B, i) for i in dims[:-1]] for _ in range(3)]
    o_net, o_loss = O.dropout_train_steps(net, X, y, batches, "relu", keeps, lr=1e-2, l2_reg=1e-4)
    assert np.allclose(loss.cpu().numpy(), o_loss, rtol=1e-5)
    for (W, b), (W1, b1) in zip(arch.unpack(tr.W.cpu()), o_net):
        assert float((W - W1).abs().max()) <= 5e-5 * float(W1.abs().max())
        assert float((b - b1).abs().max()) <= 5e-5 * float(b1.abs().max()) + 1e-7


def test_keep_rate_of_in_kernel_masks():
    """A one-layer 'network' y = w.x with w = 1, lr tiny: the step's loss reveals the mask -- E[sum_k keep_k x_k] = (1 - p) sum x."""
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    D, B, N, p = 64, 64, 64, 0.25
    arch = Arch([D, 1], "relu")
    X = torch.ones(N, D)
    y = torch.zeros(N)
    tr = ops.DropoutTrainer(arch, B, "cuda", lr=1e-12, l2_reg=0.0, p_drop=p, seed=11)
    tr.set_state(arch.pack([[(torch.ones(1, D), torch.zeros(1))]])[0])
    tr.bind(X.cuda(), y.reshape(-1, 1).cuda())
    perm = torch.arange(N).cuda()
    est = []
    for _ in range(20):
        loss = tr.run(perm, 0, 1, want_loss=True)          # sum_b (sum_k keep)^2, keep = Bernoulli(1 - p) (no 1 / (1 - p) left)
        est.append(float(loss))
    # E[(sum keep)^2] = D p (1 - p) + (D (1 - p))^2 per row
    expect = B * (D * p * (1 - p) + (D * (1 - p)) ** 2)
    assert abs(np.mean(est) / expect - 1) < 0.02, (np.mean(est), expect)


def test_unsupported_size_is_refused():
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    assert ops.dropout_train_supported(Arch([9, 100, 100, 1], "relu"), 32)
    assert not ops.dropout_train_supported(Arch([90, 512, 512, 1], "tanh"), 32)
    with pytest.raises(ValueError):
        ops.DropoutTrainer(Arch([90, 512, 512, 1], "tanh"), 32, "cuda")


def test_bnn_dropout_train_uses_the_fused_kernel():
    from bnn_b200.BNN_Dropout import BNN_Dropout
    torch.manual_seed(0)
    X = torch.linspace(-3, 3, 200).view(-1, 1)
    y = torch.sin(X).view(-1)
    m = BNN_Dropout(1, nn.Tanh(), [50, 50], conf={"device": "cuda", "num_epochs": 150, "print_every": 1000, "dropout_rate": 0.02})
    m.train(X, y)
    assert m.trained_with == "fused"
    assert 0 < m.noise_level < 0.5
    py, pv = m.predict_mv(X, m.sample(64))
    rmse = float(((py.cpu().view(-1) - y) ** 2).mean().sqrt())
    assert rmse < 0.25, rmse
