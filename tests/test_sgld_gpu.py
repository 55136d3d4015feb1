"""GPU: the fused SGLD step kernel (bnn_sgld_run, csrc/sgld_step.cu).

* gradient and loss of ONE step against the REFERENCE's own log_prior / log_lik / loss.backward()
  (BNN_SGDMC.py:61-74,83-85) recorded in tests/golden/sgdmc_energy.npz -- pinned;
* the update epilogue against the standalone update kernel (bnn_psgld_step) and the oracle's restatement with the Philox
  twin's noise (the update RULE itself is unpinned: pybnn is absent, see oracle/__init__.py);
* minibatch plumbing: permutation kernel is a bijection, short last batch of an epoch, n steps in one launch == n launches;
* BNN_SGDMC.train end to end, incl. tiny training sets (num_train < 3 * batch) and the warm-start semantics.
Tolerances: gradient 2e-5 max-normalised per tensor (FP32 sums in a different order than ATen's), loss 1e-5 relative."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from helpers import energy_cases, check_energy_grads
from oracle import bnn_oracle as O
from oracle import philox as P

pytestmark = pytest.mark.gpu


def make_chain(dims, act, B, net, log_precs, conf=None, seed=3):
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    conf = conf or {}
    arch = Arch(dims, act)
    chain = ops.SgldChain(arch, B, "cuda", conf.get("lr_weight", 2e-2), conf.get("lr_noise", 1e-1), conf.get("alpha_n", 1e-2),
                          conf.get("beta_n", 1e-1), seed=seed)
    chain.set_state(arch.pack([net])[0], log_precs)
    return arch, chain


def unpack_grad(arch, g):
    pairs = arch.unpack(g[: arch.stride].cpu())
    return pairs, g[arch.stride: arch.stride + arch.nout].cpu()


@pytest.mark.parametrize("case", energy_cases(), ids=lambda c: c["name"])
def test_step_gradient_matches_reference_fixture(case):
    arch, chain = make_chain(case["dims"], case["act"], case["B"], case["net"], case["log_precs"],
                             {"alpha_n": float(case["alpha_n"]), "beta_n": float(case["beta_n"])})
    X, y = case["X"].cuda(), case["y"].cuda()
    chain.bind(X, y, num_train=case["num_train"])
    perm = torch.arange(case["B"], dtype=torch.int64, device="cuda")
    g, loss = chain.grad(perm)
    torch.cuda.synchronize()
    chain.check()
    assert abs(float(loss) - float(case["loss"])) <= 1e-5 * abs(float(case["loss"])), (float(loss), float(case["loss"]))
    pairs, glp = unpack_grad(arch, g)
    check_energy_grads(case, pairs, glp)
    assert not bool(g[: arch.stride][~arch.valid_mask().cuda()].any())            # layout padding carries no gradient


def test_short_last_batch_and_permuted_rows():
    """Second batch of an epoch of 40 rows at batch 32 has 8 rows: N/|b| = 40/8, rows picked through the permutation."""
    dims, act, B, N = [5, 24, 16, 2], "tanh", 32, 40
    g = torch.Generator().manual_seed(0)
    net = [((torch.rand(o, i, generator=g) - 0.5), (torch.rand(o, generator=g) - 0.5) * 0.2) for i, o in zip(dims[:-1], dims[1:])]
    lp = torch.tensor([0.2, -0.3])
    arch, chain = make_chain(dims, act, B, net, lp)
    X, y = torch.randn(N, 5, generator=g), torch.randn(N, 2, generator=g)
    chain.bind(X.cuda(), y.cuda())
    perm = torch.randperm(N, generator=g)
    for off, rows in ((0, 32), (32, 8)):
        gd, loss = chain.grad(perm.cuda(), perm_off=off)
        idx = perm[off: off + rows]
        ps = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in net]
        lpr = lp.clone().requires_grad_(True)
        U = O.sgdmc_neg_log_post(ps, lpr, X[idx], y[idx], act, N, 1e-2, 1e-1)
        gs = torch.autograd.grad(U, [q for pr in ps for q in pr] + [lpr])
        assert abs(float(loss) - float(U)) <= 1e-5 * abs(float(U))
        pairs, glp = unpack_grad(arch, gd)
        for l, (gW, gb) in enumerate(pairs):
            for got, ref in ((gW, gs[2 * l]), (gb, gs[2 * l + 1])):
                assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
        assert float((glp - gs[-1]).abs().max()) <= 2e-5 * float(gs[-1].abs().max())


@pytest.mark.parametrize("dims,act,B", [
    ([6, 1000, 40, 1], "tanh", 48),        # K = 1004 > 800: the forward / dh tiles take two K rounds
    ([11, 72, 72, 2], "relu", 512),        # K = batch = 512 > 384: the dW tiles take two K rounds; 16 row blocks
    ([9, 33, 12], "tanh", 20),             # 12 outputs: the head's o >= 8 path; ragged sizes everywhere
    ([5, 3], "relu", 7),                   # no hidden layer: head straight from the gathered rows
    ([300, 64, 64, 1], "sigmoid", 100),    # wide input: several 128-column rounds of the gather
])
def test_gradient_against_autograd_on_the_uncommon_paths(dims, act, B):
    """One gradient-only step against torch autograd of the oracle's energy (the restatement pinned by the reference fixture)."""
    g = torch.Generator().manual_seed(5)
    net = [((torch.rand(o, i, generator=g) - 0.5) * (6.0 / (i + o)) ** 0.5, (torch.rand(o, generator=g) - 0.5) * 0.2)
           for i, o in zip(dims[:-1], dims[1:])]
    lp = 0.2 * torch.randn(dims[-1], generator=g)
    arch, chain = make_chain(dims, act, B, net, lp)
    N = 3 * B + 5
    X, y = torch.randn(N, dims[0], generator=g), torch.randn(N, dims[-1], generator=g)
    chain.bind(X.cuda(), y.cuda())
    perm = torch.randperm(N, generator=g)
    for off, rows in ((B, B), (3 * B, 5)):                           # a full batch and the short last one
        gd, loss = chain.grad(perm.cuda(), perm_off=off)
        torch.cuda.synchronize()
        chain.check()
        idx = perm[off: off + rows]
        ps = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in net]
        lpr = lp.clone().requires_grad_(True)
        U = O.sgdmc_neg_log_post(ps, lpr, X[idx], y[idx], act, N, 1e-2, 1e-1)
        gs = torch.autograd.grad(U, [q for pr in ps for q in pr] + [lpr])
        assert abs(float(loss) - float(U.detach())) <= 2e-5 * abs(float(U.detach()))
        pairs, glp = unpack_grad(arch, gd)
        for l, (gW, gb) in enumerate(pairs):
            for got, ref in ((gW, gs[2 * l]), (gb, gs[2 * l + 1])):
                assert float((got - ref).abs().max()) <= 3e-5 * float(ref.abs().max()), (l, tuple(ref.shape))
        assert float((glp - gs[-1]).abs().max()) <= 3e-5 * float(gs[-1].abs().max())
    # and a few real steps stay finite and keep the layout padding at zero
    chain.run(perm.cuda(), 0, 3)
    torch.cuda.synchronize()
    chain.check()
    st = chain.state
    assert bool(torch.isfinite(st).all())
    assert not bool(st[: arch.stride][~arch.valid_mask().cuda()].any())


@pytest.mark.parametrize("dims,act,B", [([90, 512, 512, 512, 1], "tanh", 128), ([13, 50, 1], "relu", 32), ([7, 20, 3], "sigmoid", 10)])
def test_update_epilogue_equals_standalone_kernel(dims, act, B):
    """One fused step == gradient-only step + bnn_psgld_step with the same (seed, t) noise; thetaT stays the transpose."""
    from bnn_b200 import ops
    g = torch.Generator().manual_seed(1)
    net = [((torch.rand(o, i, generator=g) - 0.5) * (4.0 / (i + o)) ** 0.5, (torch.rand(o, generator=g) - 0.5) * 0.2)
           for i, o in zip(dims[:-1], dims[1:])]
    lp = 0.1 * torch.randn(dims[-1], generator=g)
    arch, chain = make_chain(dims, act, B, net, lp, seed=11)
    N = 4 * B
    X, y = torch.randn(N, dims[0], generator=g).cuda(), torch.randn(N, dims[-1], generator=g).cuda()
    chain.bind(X, y)
    perm = torch.randperm(N, generator=g).cuda()
    chain.t = 7                                                      # a later point of the LR schedule
    gd, _ = chain.grad(perm, perm_off=B)
    theta0, V0 = chain.state.clone(), chain.V.clone()
    f = float(O.lr_factor(7))
    th_ref, V_ref = theta0.clone(), V0.clone()
    ops.psgld_step(th_ref, gd, V_ref, 2e-2 * f, lr_tail=1e-1 * f, n_head=arch.stride, alpha=0.99, lam=1e-5, seed=11, step=7)
    valid = torch.zeros(chain.n_theta, device="cuda")
    valid[: arch.stride] = arch.valid_mask().cuda().float()
    valid[arch.stride: arch.stride + arch.nout] = 1
    chain.run(perm, B, 1)
    torch.cuda.synchronize()
    chain.check()
    th_new = chain.state
    assert chain.t == 8 and chain.cur == 1
    assert torch.allclose(th_new * valid, th_ref * valid, rtol=2e-6, atol=1e-7)
    assert torch.allclose(chain.V * valid, V_ref * valid, rtol=2e-6, atol=1e-9)
    assert not bool((th_new * (1 - valid)).any())                    # padding stays zero: no noise lands there
    assert torch.equal(chain.V * (1 - valid), V0 * (1 - valid))
    # transposed weight copies of Linear 1.. follow the update
    L = _lib_lengths(arch)
    off = 0
    for l in range(1, arch.n_linear):
        ldT = (arch.outs[l] + 3) // 4 * 4
        blk = chain.thetaT[chain.cur][off: off + arch.ins[l] * ldT].view(arch.ins[l], ldT)
        W = arch.layer_view(th_new[: arch.stride], l)[:, : arch.ins[l]]
        assert torch.equal(blk[:, : arch.outs[l]], W.t())
        off += arch.ins[l] * ldT
    assert off == L or arch.n_linear == 1


def _lib_lengths(arch):
    import ctypes as C
    from bnn_b200 import _lib
    d = arch.desc()
    return int(_lib.lib().bnn_sgld_theta_t_len(C.byref(d)))


def test_steps_in_one_launch_equal_single_step_launches():
    dims, act, B, N = [9, 64, 64, 1], "tanh", 32, 300
    g = torch.Generator().manual_seed(2)
    net = [((torch.rand(o, i, generator=g) - 0.5) * 0.5, torch.zeros(o)) for i, o in zip(dims[:-1], dims[1:])]
    X, y = torch.randn(N, 9, generator=g).cuda(), torch.randn(N, 1, generator=g).cuda()
    perm = torch.randperm(N, generator=g).cuda()
    outs = []
    for split in (False, True):
        arch, chain = make_chain(dims, act, B, net, torch.zeros(1), seed=5)
        chain.bind(X, y)
        if split:
            for i in range(7):
                loss = chain.run(perm, i * B, 1)
        else:
            loss = chain.run(perm, 0, 7)
        torch.cuda.synchronize()
        chain.check()
        outs.append((chain.state.clone(), chain.V.clone(), float(loss)))
    assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-6, atol=1e-7)
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-6, atol=1e-9)
    assert abs(outs[0][2] - outs[1][2]) <= 1e-6 * abs(outs[1][2])


def test_three_steps_against_the_oracle_chain():
    """Fused steps vs the CPU restatement: oracle energy + autograd gradient + oracle pSGLD update fed with the Philox
    twin's normals and the LambdaLR factor (BNN_SGDMC.py:76-91,101)."""
    dims, act, B, N, seed = [6, 32, 32, 2], "tanh", 16, 100, 21
    g = torch.Generator().manual_seed(3)
    net = [((torch.rand(o, i, generator=g) - 0.5) * 0.8, (torch.rand(o, generator=g) - 0.5) * 0.1) for i, o in zip(dims[:-1], dims[1:])]
    lp = torch.tensor([0.3, -0.2])
    arch, chain = make_chain(dims, act, B, net, lp, {"lr_weight": 1e-2, "lr_noise": 5e-2}, seed=seed)
    X, y = torch.randn(N, 6, generator=g), torch.randn(N, 2, generator=g)
    chain.bind(X.cuda(), y.cuda())
    perm = torch.randperm(N, generator=g)
    chain.run(perm.cuda(), 0, 3)
    torch.cuda.synchronize()
    chain.check()
    # CPU chain on the packed vector
    theta = torch.zeros(chain.n_theta)
    theta[: arch.stride] = arch.pack([net])[0]
    theta[arch.stride: arch.stride + 2] = lp
    V = torch.ones(chain.n_theta)
    valid = torch.zeros(chain.n_theta, dtype=torch.bool)
    valid[: arch.stride] = arch.valid_mask()
    valid[arch.stride: arch.stride + 2] = True
    for t in range(3):
        idx = perm[t * B: (t + 1) * B]
        th = theta.clone().requires_grad_(True)
        pairs = [(arch.layer_view(th[: arch.stride], l)[:, : arch.ins[l]], arch.layer_view(th[: arch.stride], l)[:, arch.ins[l]])
                 for l in range(arch.n_linear)]
        U = O.sgdmc_neg_log_post(pairs, th[arch.stride: arch.stride + 2], X[idx], y[idx], act, N, 1e-2, 1e-1)
        grad, = torch.autograd.grad(U, th)
        z = torch.from_numpy(P.stream_normal(seed, t, P.STREAM_SGLD, chain.n_theta))
        f = float(O.lr_factor(t))
        lr = torch.full((chain.n_theta,), 1e-2 * f)
        lr[arch.stride:] = 5e-2 * f
        th2, V2 = O.psgld_step(theta, grad * valid, V, lr, z)
        theta = torch.where(valid, th2, theta)
        V = torch.where(valid, V2, V)
    assert float((chain.state.cpu() - theta).abs().max()) <= 2e-5 * float(theta.abs().max())
    assert float((chain.V.cpu() - V).abs().max()) <= 2e-5 * float(V.abs().max())


def test_epoch_permutation_is_a_bijection():
    from bnn_b200 import ops
    for N in (1, 2, 40, 455, 4573, 100_000):
        p0 = ops.sgld_perm(7, 0, N, N, "cuda")
        assert torch.equal(torch.sort(p0).values, torch.arange(N, device="cuda"))
        if N >= 455:
            p1 = ops.sgld_perm(7, 1, N, N, "cuda")
            assert not torch.equal(p0, p1)                                      # a new shuffle every epoch
            assert torch.equal(ops.sgld_perm(7, 0, N, 100, "cuda", offset=50), p0[50:150])   # any window, same permutation
            # no obvious structure: mean displacement of a random permutation is ~N/3
            disp = (p0 - torch.arange(N, device="cuda")).abs().double().mean()
            assert 0.25 * N < float(disp) < 0.42 * N
    big = ops.sgld_perm(9, 3, 10_000_000, 6400, "cuda")                        # cfg 5: a prefix of a 10M-row epoch
    assert int(big.min()) >= 0 and int(big.max()) < 10_000_000 and big.unique().numel() == 6400


@pytest.mark.parametrize("num_train", [40, 70, 200])
def test_train_small_sets(num_train):
    """BNN_SGDMC.train on training sets of 32..95 points at batch 32 (the range BO.py's loop passes through)."""
    from bnn_b200.BNN_SGDMC import BNN_SGDMC
    torch.manual_seed(0)
    X = torch.linspace(-1, 1, num_train).view(-1, 1)
    y = torch.sin(3 * X).squeeze() + 0.05 * torch.randn(num_train)
    conf = dict(steps_burnin=600, steps=200, keep_every=20, batch_size=32, lr_weight=1e-1, lr_noise=1e-1, seed=1)
    model = BNN_SGDMC(1, nn.Tanh(), [32], nout=1, conf=conf)
    model.train(X, y)
    assert len(model.nns) == 10
    rmse, nll = model.validate(X, y, num_samples=10)
    assert torch.isfinite(rmse).all() and float(rmse) < (0.6 if num_train < 100 else 0.3)    # short chains on 40..70 points: sanity only


def test_warm_start_restarts_optimizer_and_continues_from_the_module():
    """BNN_SGDMC.py:96-101,114-115: every train() builds a fresh pSGLD (V = 1) and LambdaLR (t = 0); warm_start only skips
    init_nn(), so the chain continues from the CURRENT self.nn / self.log_precs (including edits made by the user)."""
    from bnn_b200.BNN_SGDMC import BNN_SGDMC
    torch.manual_seed(1)
    X = torch.randn(64, 2)
    y = X[:, 0] - X[:, 1]
    conf = dict(steps_burnin=40, steps=40, keep_every=20, batch_size=32, warm_start=True, seed=4)
    model = BNN_SGDMC(2, nn.Tanh(), [8], nout=1, conf=conf)
    model.train(X, y)
    assert model._chain.t == 80
    with torch.no_grad():
        for q in model.nn.nn.parameters():
            q.mul_(0.5)                                               # user edit between the BO iterations
    w_edit = model.arch.pack([model.nn.nn])[0]
    model.steps_burnin, model.steps = 0, 20
    model.train(X, y)
    assert model._chain.t == 20                                       # scheduler restarted
    # a cold start re-initialises instead
    model.warm_start = False
    model.steps = 0
    before = model.arch.pack([model.nn.nn])[0].clone()
    model.nns = []
    try:
        model.train(X, y)
    except Exception:
        pass
    assert not torch.equal(model.arch.pack([model.nn.nn])[0], before)
    assert w_edit.shape == before.shape


@pytest.mark.parametrize("dims,act,B,N", [([13, 50, 1], "relu", 32, 455), ([9, 64, 64, 2], "tanh", 24, 200), ([5, 3], "relu", 8, 50)])
def test_single_sm_kernel_equals_dataflow_kernel(dims, act, B, N, monkeypatch):
    """Networks that fit one SM run in sgld_small_kernel (chain state in shared memory, one CTA); BNN_SGLD_SMALL=0 forces the
    multi-CTA dataflow kernel.  Same arithmetic, same Philox stream, same LambdaLR factors: gradients, nine steps (short last
    batch included), the loss, and a hand-over from one kernel to the other in the middle of a chain must agree."""
    g = torch.Generator().manual_seed(11)
    net = [((torch.rand(o, i, generator=g) - 0.5) * 0.6, (torch.rand(o, generator=g) - 0.5) * 0.1) for i, o in zip(dims[:-1], dims[1:])]
    lp = torch.linspace(-0.2, 0.3, dims[-1])
    X, y = torch.randn(N, dims[0], generator=g).cuda(), torch.randn(N, dims[-1], generator=g).cuda()
    perm = torch.randperm(N, generator=g).cuda()
    last = ((N + B - 1) // B - 1) * B                        # offset of the epoch's (possibly short) last batch
    res = {}
    for mode in ("small", "dataflow", "mixed"):
        arch, chain = make_chain(dims, act, B, net, lp, seed=9)
        chain.bind(X, y)
        monkeypatch.setenv("BNN_SGLD_SMALL", "0" if mode == "dataflow" else "2")
        grad, gloss = chain.grad(perm, 0)
        chain.run(perm, 0, 4)
        if mode == "mixed":
            monkeypatch.setenv("BNN_SGLD_SMALL", "0")
        chain.run(perm, 4 * B, 3)
        if mode == "mixed":
            monkeypatch.setenv("BNN_SGLD_SMALL", "2")
        loss = chain.run(perm, last, 1)
        chain.run(perm, 0, 1)
        torch.cuda.synchronize()
        chain.check()
        res[mode] = (grad, chain.state.clone(), chain.V.clone(), float(loss), float(gloss))
    ref = res["dataflow"]
    for mode in ("small", "mixed"):
        got = res[mode]
        assert float((got[0] - ref[0]).abs().max()) <= 2e-6 * float(ref[0].abs().max()), mode
        assert float((got[1] - ref[1]).abs().max()) <= 5e-6 * float(ref[1].abs().max()), mode
        assert float((got[2] - ref[2]).abs().max()) <= 5e-6 * float(ref[2].abs().max()), mode
        assert abs(got[3] - ref[3]) <= 1e-5 * abs(ref[3]), mode
        assert abs(got[4] - ref[4]) <= 1e-5 * abs(ref[4]), mode


def test_capped_grid_and_two_streams_give_the_same_chains():
    """bnn_sgld_set_grid_cap: two chains with half the CTAs each, launched on two streams so that they run side by side, end in
    exactly the state each reaches alone on the whole GPU (the job list is the same, only the number of workers changes)."""
    from bnn_b200 import ops
    dims, act, B, N = [20, 256, 256, 1], "tanh", 64, 2000
    g = torch.Generator().manual_seed(4)
    X, y = torch.randn(N, 20, generator=g).cuda(), torch.randn(N, 1, generator=g).cuda()
    perm = torch.randperm(N, generator=g).cuda()
    nets = [[((torch.rand(o, i, generator=g) - 0.5) * 0.3, torch.zeros(o)) for i, o in zip(dims[:-1], dims[1:])] for _ in range(2)]

    def run(capped):
        chains = []
        for c in range(2):
            _, ch = make_chain(dims, act, B, nets[c], torch.zeros(1), seed=30 + c)
            ch.bind(X, y)
            chains.append(ch)
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream() for _ in range(2)] if capped else [torch.cuda.current_stream()] * 2
        ops.sgld_set_grid_cap(torch.cuda.get_device_properties(0).multi_processor_count // 2 if capped else 0)
        try:
            for r in range(3):
                for c in range(2):
                    with torch.cuda.stream(streams[c]):
                        chains[c].run(perm, r * 5 * B, 5)
            torch.cuda.synchronize()
        finally:
            ops.sgld_set_grid_cap(0)
        for ch in chains:
            ch.check()
        return [(ch.state.clone(), ch.V.clone()) for ch in chains]

    alone, side = run(False), run(True)
    for (s0, v0), (s1, v1) in zip(alone, side):
        assert torch.allclose(s0, s1, rtol=1e-6, atol=1e-7)
        assert torch.allclose(v0, v1, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("hidden", [[64, 64], [50]], ids=["dataflow_kernel", "single_sm_kernel"])
def test_num_chains_pools_the_samples_of_chains_trained_side_by_side(hidden):
    """conf['num_chains'] = 2: two chains (seeds seed, seed + 1) share the GPU (one stream each; capped grids on the dataflow
    kernel, one SM each on the single-SM kernel); the pooled bank holds both chains' thinned samples, chain 0's half equals what
    a single-chain model with the same seed produces."""
    from bnn_b200.BNN_SGDMC import BNN_SGDMC
    g = torch.Generator().manual_seed(0)
    X = torch.randn(300, 6, generator=g)
    y = torch.sin(X[:, 0]) + 0.1 * torch.randn(300, generator=g)
    conf = {"device": "cuda", "seed": 7, "steps_burnin": 100, "steps": 200, "keep_every": 50, "batch_size": 32}
    banks = []
    for C in (1, 2):
        torch.manual_seed(1)                                    # same initial network (xavier init draws from torch's global stream)
        m = BNN_SGDMC(6, nn.Tanh(), hidden, nout=1, conf=dict(conf, num_chains=C))
        m.train(X, y)
        assert len(m.nns) == 4 * C
        py, pv = m.predict_mv(X, m.sample(4 * C))
        assert bool(torch.isfinite(py).all() and (pv > 0).all())
        banks.append(m.nns.weights.clone())
    assert torch.allclose(banks[1][:4], banks[0], rtol=1e-6, atol=1e-7)      # chain 0 of the pair == the chain alone
    assert not torch.allclose(banks[1][4:], banks[0], rtol=1e-3, atol=1e-4)  # chain 1 is a different chain
