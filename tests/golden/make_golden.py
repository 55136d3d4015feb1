#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference at /root/reference.

Run in the authoring container only (the reference does not travel to the GPU
box):   python tests/golden/make_golden.py

What is recorded: inputs (weights / mu,rho / base weights, data rows), the
random draws the reference consumed (captured by wrapping torch.randn /
torch.rand / F.dropout while the reference's own ``sample()`` runs -- the
reference code itself is not modified), and the reference's outputs
(sample_predict, predict_mv, validate, KL).  While generating, the CPU oracle is
checked against the reference (asserts below), which is what pins it.

``pybnn`` is absent offline; BNN_SGDMC imports it at module import only
(BNN_SGDMC.py:14-17), so four empty stand-in classes are registered -- no
arithmetic is stubbed; ``train()`` (the only user) is never called here.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

for name, cls in [("pybnn", None), ("pybnn.sampler", None),
                  ("pybnn.sampler.sgld", "SGLD"),
                  ("pybnn.sampler.preconditioned_sgld", "PreconditionedSGLD"),
                  ("pybnn.sampler.sghmc", "SGHMC"),
                  ("pybnn.sampler.adaptive_sghmc", "AdaptiveSGHMC")]:
    m = types.ModuleType(name)
    if cls:
        setattr(m, cls, type(cls, (), {}))
    sys.modules[name] = m

import BNN_SGDMC as ref_sgdmc      # noqa: E402
import BNN_BBB as ref_bbb          # noqa: E402
import BNN_Dropout as ref_do       # noqa: E402
import BNN_CDropout as ref_cdo     # noqa: E402
from util import normalize         # noqa: E402

from oracle import bnn_oracle as O  # noqa: E402
sys.path.insert(0, os.path.dirname(HERE))
from helpers import energy_inputs   # noqa: E402  (shared with the tests, which rebuild the same inputs)


def seq_to_net(seq):
    return [(l.weight.detach().clone(), l.bias.detach().clone()) for l in seq if isinstance(l, nn.Linear)]


def flat(net):
    return np.concatenate([np.concatenate([W.numpy().ravel(), b.numpy().ravel()]) for W, b in net]).astype(np.float32)


def uci_split(dataset, split):
    d = np.loadtxt(f"{REF}/UCI_Datasets/{dataset}/data/data.txt")
    tr = np.loadtxt(f"{REF}/UCI_Datasets/{dataset}/data/index_train_{split}.txt", dtype=np.int64)
    te = np.loadtxt(f"{REF}/UCI_Datasets/{dataset}/data/index_test_{split}.txt", dtype=np.int64)
    f = lambda a: torch.FloatTensor(a)
    return f(d[tr, :-1]), f(d[tr, -1]), f(d[te, :-1]), f(d[te, -1])


def save(name, **kw):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in kw.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


# --------------------------------------------------------------------------- #
def gen_sgdmc_boston():
    """cfg 1: BNN_SGDMC, bostonHousing split 0, 1x50 ReLU, 100 posterior samples."""
    torch.manual_seed(11); np.random.seed(11)
    trx, try_, tex, tey = uci_split("bostonHousing", 0)
    xm, xs, ym, ys = normalize(trx, try_)                       # experiments/SGDMC.py:55-58
    tex_n = (tex - xm) / xs
    model = ref_sgdmc.BNN_SGDMC(13, nn.ReLU(), [50], nout=1, conf={})
    model.nns = []
    from copy import deepcopy
    for _ in range(100):                                        # stand-in for a thinned chain
        net = deepcopy(model.nn)
        for p in net.parameters():
            p.data += 0.05 * torch.randn_like(p)
        model.nns.append(net)
    model.log_precs.data.fill_(1.3)
    nets = [seq_to_net(n.nn) for n in model.nns]
    with torch.no_grad():
        pred = model.sample_predict(model.nns, tex_n)           # [100,51,1]
        py, pv = model.predict_mv(tex_n, model.nns)
        # BNN.validate with a recorded permutation (BNN_SGDMC.py:127-129 uses numpy's global RNG)
        np.random.seed(5)
        perm = np.random.permutation(100)[:40]
        np.random.seed(5)
        rmse, nll = model.validate(tex_n, (tey - ym) / ys, num_samples=40)
        # experiments/SGDMC.py:81-93 variant
        preds_d = pred * ys + ym
        py_d = preds_d.mean(dim=0).squeeze()
        pv_d = preds_d.var(dim=0).squeeze() + ys ** 2 / model.log_precs.exp().squeeze().detach()
        rmse_d = torch.mean((py_d - tey) ** 2).sqrt()
        nll_d = -1 * torch.distributions.Normal(py_d, pv_d.sqrt()).log_prob(tey.squeeze()).mean()
        # pin the oracle
        o_pred = O.sample_predict(nets, tex_n, "relu", nout=1)
        assert torch.equal(o_pred, pred)
        o_py, o_pv = O.predict_mv(o_pred, 51)
        assert torch.equal(o_py, py) and torch.equal(o_pv, pv)
        sub = [nets[i] for i in perm]
        s_py, s_pv = O.predict_mv(O.sample_predict(sub, tex_n, "relu", nout=1), 51)
        o_rmse, o_nll = O.validate_metrics(s_py, s_pv, (tey - ym) / ys)
        assert torch.allclose(o_rmse, rmse, rtol=1e-6) and torch.allclose(o_nll, nll, rtol=1e-6)
    save("sgdmc_boston.npz", dims=[13, 50, 1], act="relu",
         W=np.stack([flat(n) for n in nets]), X=tex_n, y_norm=(tey - ym) / ys, y_raw=tey,
         ym=ym, ys=ys, log_prec=model.log_precs.detach(),
         pred=pred, mean=py, var=pv, perm=perm, rmse=rmse, nll=nll,
         mean_denorm=py_d, var_denorm=pv_d, rmse_denorm=rmse_d, nll_denorm=nll_d)


def gen_sgdmc_multiout():
    """nout > 1 (experiments/MO_SGDMC.py shape family): 10 -> 30 -> 30 -> 15, tanh."""
    torch.manual_seed(12)
    model = ref_sgdmc.BNN_SGDMC(10, nn.Tanh(), [30, 30], nout=15, conf={})
    from copy import deepcopy
    model.nns = []
    for _ in range(12):
        net = deepcopy(model.nn)
        for p in net.parameters():
            p.data += 0.1 * torch.randn_like(p)
        model.nns.append(net)
    X = torch.randn(37, 10)
    y = torch.randn(37, 15)
    nets = [seq_to_net(n.nn) for n in model.nns]
    with torch.no_grad():
        pred = model.sample_predict(model.nns, X)
        py, pv = model.predict_mv(X, model.nns)
        rmse = torch.mean((py - y) ** 2, dim=0).sqrt()                       # BNN.py:30
        nll = -1 * torch.distributions.Normal(py, pv.sqrt()).log_prob(y).mean(dim=0)  # BNN.py:31
        o_pred = O.sample_predict(nets, X, "tanh", nout=15)
        assert torch.equal(o_pred, pred)
        o_rmse, o_nll = O.validate_metrics(*O.predict_mv(o_pred, 37), y)
        assert torch.allclose(o_rmse, rmse, rtol=1e-6) and torch.allclose(o_nll, nll, rtol=1e-6)
    save("sgdmc_multiout.npz", dims=[10, 30, 30, 15], act="tanh",
         W=np.stack([flat(n) for n in nets]), X=X, y=y, pred=pred, mean=py, var=pv, rmse=rmse, nll=nll)


def gen_bbb_sinc():
    """cfg 2: BNN_BBB, 1-D sinc toy, 2x50 tanh; eps captured from torch.randn."""
    torch.manual_seed(13)
    model = ref_bbb.BNN_BBB(1, nn.Tanh(), [50, 50], conf={"weight_std": 0.1})
    gls = [l for l in model.nn.nn if isinstance(l, ref_bbb.GaussianLinear)]
    for l in gls:                                   # move off the init point; cover the clamp and the threshold
        l.mu.data += 0.3 * torch.randn_like(l.mu)
        l.rho.data += 1.5 * torch.randn_like(l.rho)
    gls[0].rho.data[0, 0] = -50.                    # exercises clamp(min=-35)
    gls[0].rho.data[1, 0] = 25.                     # exercises softplus threshold 20
    gls[1].rho.data[3, 4] = 19.5
    S = 8
    rec = []
    orig = torch.randn
    torch.randn = lambda *a, **k: (rec.append(orig(*a, **k)) or rec[-1])
    try:
        torch.manual_seed(99)
        nns = model.sample(S)                       # BNN_BBB.py:139-141
    finally:
        torch.randn = orig
    assert len(rec) == S * len(gls)
    X = torch.linspace(-3, 3, 64).view(-1, 1)
    y = torch.sin(3 * X) / (3 * X)
    with torch.no_grad():
        pred = model.sample_predict(nns, X)         # [S,64]
        py, pv = model.predict_mv(X, nns)
        mse, kld = model.loss(X, y)                 # kld is deterministic (BNN_BBB.py:107-110)
        # pin the oracle draw + KL + forward
        nets = []
        for s in range(S):
            net = []
            for li, l in enumerate(gls):
                wb = O.bbb_rsample(l.mu.detach(), l.rho.detach(), rec[s * len(gls) + li])
                net.append(O.bbb_split(wb))
            nets.append(net)
        ref_nets = [seq_to_net(n) for n in nns]
        for a, b in zip(nets, ref_nets):
            for (W1, b1), (W2, b2) in zip(a, b):
                assert torch.equal(W1, W2) and torch.equal(b1, b2)
        assert torch.equal(O.sample_predict(nets, X, "tanh"), pred)
        o_kl = O.bbb_kl([l.mu.detach() for l in gls], [l.rho.detach() for l in gls], 0.1)
        assert torch.allclose(o_kl, kld.detach(), rtol=1e-6), (o_kl, kld)
    kw = {}
    for li, l in enumerate(gls):
        kw[f"mu{li}"] = l.mu.detach()
        kw[f"rho{li}"] = l.rho.detach()
        kw[f"eps{li}"] = torch.stack([rec[s * len(gls) + li] for s in range(S)])
        kw[f"wb{li}"] = torch.stack([torch.cat([ref_nets[s][li][0], ref_nets[s][li][1][:, None]], dim=1) for s in range(S)])
    save("bbb_sinc.npz", dims=[1, 50, 50, 1], act="tanh", weight_std=0.1, X=X, y=y,
         pred=pred, mean=py, var=pv, kl=kld.detach(), **kw)


def gen_dropout_protein():
    """cfg 3a: BNN_Dropout, protein split 0 rows, 2x100 tanh, p = 0.05; masks captured from F.dropout."""
    torch.manual_seed(14)
    trx, try_, tex, tey = uci_split("protein-tertiary-structure", 0)
    xm, xs, _, _ = normalize(trx, try_)
    X = ((tex - xm) / xs)[:96]
    y = tey[:96]
    p = 0.05
    model = ref_do.BNN_Dropout(9, nn.Tanh(), [100, 100], conf={"dropout_rate": p})
    for q in model.nn.parameters():
        q.data += 0.05 * torch.randn_like(q)
    S = 6
    rec = []
    orig = F.dropout
    F.dropout = lambda *a, **k: (rec.append(orig(*a, **k)) or rec[-1])
    try:
        torch.manual_seed(77)
        nns = model.sample(S)                       # BNN_Dropout.py:67-76
    finally:
        F.dropout = orig
    assert len(rec) == 3 * S                        # every Linear has in > 1 here
    masks = [[(rec[3 * s + l] * (1 - p)) for l in range(3)] for s in range(S)]
    for s in range(S):
        for m in masks[s]:
            assert set(np.unique(m.numpy())).issubset({0.0, 1.0})
    base = seq_to_net(model.nn.nn)
    with torch.no_grad():
        pred = model.sample_predict(nns, X)
        py, pv = model.predict_mv(X, nns)
        nets = [O.scale_columns(base, masks[s], skip_unit_inputs=True) for s in range(S)]
        for a, b in zip(nets, [seq_to_net(n) for n in nns]):
            for (W1, b1), (W2, b2) in zip(a, b):
                assert torch.equal(W1, W2) and torch.equal(b1, b2)
        assert torch.equal(O.sample_predict(nets, X, "tanh"), pred)
    save("dropout_protein.npz", dims=[9, 100, 100, 1], act="tanh", p=p, base=flat(base), X=X, y=y,
         masks=np.stack([np.concatenate([m.numpy() for m in masks[s]]) for s in range(S)]),
         pred=pred, mean=py, var=pv)


def gen_cdropout_protein():
    """cfg 3b: BNN_CDropout, 2x100 tanh; uniforms captured from torch.rand."""
    torch.manual_seed(15)
    trx, try_, tex, _ = uci_split("protein-tertiary-structure", 0)
    xm, xs, _, _ = normalize(trx, try_)
    X = ((tex - xm) / xs)[100:180]
    model = ref_cdo.BNN_CDropout(9, nn.Tanh(), [100, 100], conf={})
    cls = [l for l in model.nn.nn if isinstance(l, ref_cdo.CDropoutLinear)]
    for i, l in enumerate(cls):
        l.layer[0].p_logit.data.fill_([-2.0, 0.0, 1.0][i])
        l.layer[1].weight.data += 0.05 * torch.randn_like(l.layer[1].weight)
    S = 5
    rec = []
    orig = torch.rand
    torch.rand = lambda *a, **k: (rec.append(orig(*a, **k)) or rec[-1])
    try:
        torch.manual_seed(55)
        nns = model.sample(S)                       # BNN_CDropout.py:154-156
    finally:
        torch.rand = orig
    assert len(rec) == 3 * S
    base = [(l.layer[1].weight.detach().clone(), l.layer[1].bias.detach().clone()) for l in cls]
    p_logits = [float(l.layer[0].p_logit) for l in cls]
    with torch.no_grad():
        pred = model.sample_predict(nns, X)
        py, pv = model.predict_mv(X, nns)
        masks = [[O.concrete_keep_mask(rec[3 * s + l], p_logits[l]) for l in range(3)] for s in range(S)]
        nets = [O.scale_columns(base, masks[s], skip_unit_inputs=False) for s in range(S)]
        for a, b in zip(nets, [seq_to_net(n) for n in nns]):
            for (W1, b1), (W2, b2) in zip(a, b):
                assert torch.equal(W1, W2) and torch.equal(b1, b2)
        assert torch.equal(O.sample_predict(nets, X, "tanh"), pred)
    save("cdropout_protein.npz", dims=[9, 100, 100, 1], act="tanh", p_logits=p_logits, base=flat(base), X=X,
         uniforms=np.stack([np.concatenate([rec[3 * s + l].numpy() for l in range(3)]) for s in range(S)]),
         masks=np.stack([np.concatenate([m.numpy() for m in masks[s]]) for s in range(S)]),
         pred=pred, mean=py, var=pv)


# --------------------------------------------------------------------------- #
ENERGY_CASES = [
    # name, dims, act, B, num_train, conf
    ("relu_1x50", [13, 50, 1], "relu", 32, 455, {}),
    ("relu_1x50_uci", [13, 50, 1], "relu", 128, 455, {"alpha_n": 0.5, "beta_n": 0.4}),       # experiments/SGDMC.py:63-72
    ("tanh_3x512", [90, 512, 512, 512, 1], "tanh", 128, 10_000_000, {}),
    ("tanh_2x30_mo3", [10, 30, 30, 3], "tanh", 37, 2000, {"alpha_n": 0.5, "beta_n": 0.4}),  # nout = 3, ragged batch
    ("tanh_3x512_mo3", [90, 512, 512, 512, 3], "tanh", 128, 10_000_000, {}),
    ("tanh_1x50_noise", [1, 50, 1], "tanh", 20, 20, {"noise_level": 0.01}),                 # BNN_SGDMC.py:39-45
]
GRAD_STRIDE = 61      # big gradients are stored as every 61st element + their sum and L2 norm


def gen_sgdmc_energy():
    """a12 energy: the reference's own log_prior(), log_lik(bx, by), loss and autograd gradient
    (BNN_SGDMC.py:61-74, :83-85) on deterministic inputs -- pins oracle.sgdmc_neg_log_post and, on the GPU, the fused
    SGLD step's gradient."""
    out = {"names": np.array([c[0] for c in ENERGY_CASES]), "grad_stride": GRAD_STRIDE}
    for ci, (name, dims, act, B, num_train, conf) in enumerate(ENERGY_CASES):
        nout = dims[-1]
        model = ref_sgdmc.BNN_SGDMC(dims[0], {"relu": nn.ReLU(), "tanh": nn.Tanh()}[act], dims[1:-1], nout=nout, conf=conf)
        net, log_precs, X, y = energy_inputs(dims, nout, B, 1000 + ci)
        lins = [l for l in model.nn.nn if isinstance(l, nn.Linear)]
        for lin, (W, b) in zip(lins, net):
            lin.weight.data, lin.bias.data = W.clone(), b.clone()
        model.log_precs.data = log_precs.clone()
        log_prior = model.log_prior()                                    # BNN_SGDMC.py:61-67
        log_lik = model.log_lik(X, y)                                    # :69-74
        loss = -1 * (log_lik * (num_train / X.shape[0]) + log_prior)     # :83
        model.zero_grad()
        loss.backward()                                                  # :85
        # pin the oracle restatement: value and gradient
        ps = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in net]
        lp = log_precs.clone().requires_grad_(True)
        U = O.sgdmc_neg_log_post(ps, lp, X, y, act, num_train, float(model.alpha_n), float(model.beta_n))
        assert torch.allclose(U.detach(), loss.detach(), rtol=2e-6), (name, float(U), float(loss))
        gs = torch.autograd.grad(U, [t for pr in ps for t in pr] + [lp])
        ref_gs = [t.grad for lin in lins for t in (lin.weight, lin.bias)] + [model.log_precs.grad]
        for a, b_ in zip(gs, ref_gs):
            assert float((a - b_).abs().max()) <= 2e-5 * float(b_.abs().max() + 1e-30), name
        out[f"{name}.dims"] = np.array(dims)
        out[f"{name}.act"] = np.array(act)
        out[f"{name}.B"] = B
        out[f"{name}.num_train"] = num_train
        out[f"{name}.seed"] = 1000 + ci
        out[f"{name}.alpha_n"] = float(model.alpha_n)
        out[f"{name}.beta_n"] = float(model.beta_n)
        out[f"{name}.log_prior"] = log_prior.detach()
        out[f"{name}.log_lik"] = log_lik.detach()
        out[f"{name}.loss"] = loss.detach()
        out[f"{name}.grad_log_precs"] = model.log_precs.grad
        for li, lin in enumerate(lins):
            for tag, t in (("W", lin.weight.grad), ("b", lin.bias.grad)):
                v = t.detach().reshape(-1)
                out[f"{name}.g{tag}{li}.sub"] = v[::GRAD_STRIDE] if v.numel() > 4096 else v
                out[f"{name}.g{tag}{li}.sum"] = v.double().sum()
                out[f"{name}.g{tag}{li}.norm"] = v.double().norm()
    save("sgdmc_energy.npz", **out)



def gen_bbb_train():
    """f2: three minibatch steps of the REFERENCE's BNN_BBB training recipe -- model.loss (GaussianLinear.forward with local
    reparameterisation, MSELoss, KL; BNN_BBB.py:29-37,101-110), loss.backward(), torch.optim.Adam(lr).step() (:114,121-126) --
    on fixed batches, with the torch.randn draws of the forward captured.  Pins oracle.bbb_train_steps and, on the GPU, the fused
    training kernel."""
    torch.manual_seed(21)
    conf = {"weight_std": 0.1, "lr": 1e-2, "kl_factor": 1e-3}
    model = ref_bbb.BNN_BBB(3, nn.Tanh(), [20, 12], conf=conf)
    gls = [l for l in model.nn.nn if isinstance(l, ref_bbb.GaussianLinear)]
    for l in gls:
        l.mu.data += 0.2 * torch.randn_like(l.mu)
        l.rho.data += 0.5 * torch.randn_like(l.rho)
    gls[0].rho.data[0, 0] = -50.                    # clamp(min=-35): no gradient through it
    N, B = 37, 16
    X = torch.randn(N, 3)
    y = torch.sin(X.sum(dim=1))
    perm = torch.randperm(N)
    batches = [perm[0:16], perm[16:32], perm[32:37]]            # the last batch of the epoch is short
    mu0 = [l.mu.detach().clone() for l in gls]
    rho0 = [l.rho.detach().clone() for l in gls]
    opt = torch.optim.Adam(model.nn.parameters(), model.lr)    # BNN_BBB.py:114
    rec = []
    orig = torch.randn
    torch.randn = lambda *a, **k: (rec.append(orig(*a, **k)) or rec[-1])
    mses, kls = [], []
    try:
        for idx in batches:                                      # BNN_BBB.py:121-126
            opt.zero_grad()
            mse, kl_term = model.loss(X[idx], y[idx])
            loss = mse + model.kl_factor * kl_term
            loss.backward()
            opt.step()
            mses.append(float(mse.detach())); kls.append(float(kl_term.detach()))
    finally:
        torch.randn = orig
    assert len(rec) == len(batches) * len(gls)
    eps = [[rec[i * len(gls) + l] for l in range(len(gls))] for i in range(len(batches))]
    # pin the oracle restatement
    o_mu, o_rho, o_mse, o_kl = O.bbb_train_steps(mu0, rho0, X, y, batches, "tanh", eps, lr=1e-2, kl_factor=1e-3, weight_std=0.1)
    for l, g in enumerate(gls):
        assert torch.allclose(o_mu[l], g.mu.detach(), rtol=1e-6, atol=1e-7), l
        assert torch.allclose(o_rho[l], g.rho.detach(), rtol=1e-6, atol=1e-7), l
    assert np.allclose(o_mse, mses, rtol=1e-6) and np.allclose(o_kl, kls, rtol=1e-6)
    kw = {}
    for l, g in enumerate(gls):
        kw[f"mu0_{l}"], kw[f"rho0_{l}"] = mu0[l], rho0[l]
        kw[f"mu1_{l}"], kw[f"rho1_{l}"] = g.mu.detach(), g.rho.detach()
        for i in range(len(batches)):
            kw[f"eps{i}_{l}"] = eps[i][l]
    save("bbb_train.npz", dims=[3, 20, 12, 1], act="tanh", weight_std=0.1, lr=1e-2, kl_factor=1e-3, batch=B, X=X, y=y,
         perm=perm.numpy().astype(np.int64), mse=np.array(mses, dtype=np.float32), kl=np.array(kls, dtype=np.float32), **kw)


def gen_dropout_train():
    """f2: three minibatch steps of the REFERENCE's BNN_Dropout training recipe -- NN_Dropout.forward in training mode, MSELoss(sum),
    loss.backward(), Adam with the two weight-decay groups (BNN_Dropout.py:15-21,43-62) -- on fixed batches, with the keep factors
    of every F.dropout call captured (F.dropout(ones) * (1 - p): the same generator state as the reference's own call).  Pins
    oracle.dropout_train_steps and, on the GPU, the fused training kernel."""
    torch.manual_seed(22)
    p = 0.2
    model = ref_do.BNN_Dropout(5, nn.Tanh(), [24, 16], conf={"dropout_rate": p, "l2_reg": 1e-3, "lr": 1e-2})
    N, B = 41, 16
    X = torch.randn(N, 5)
    y = torch.sin(X.sum(dim=1))
    perm = torch.randperm(N)
    batches = [perm[0:16], perm[16:32], perm[32:41]]             # the last batch of the epoch is short
    net0 = seq_to_net(model.nn.nn)
    dict_decay = {'params': [], 'weight_decay': model.l2_reg}      # BNN_Dropout.py:43-52
    dict_no_decay = {'params': [], 'weight_decay': 0.}
    for name, param in model.nn.named_parameters():
        (dict_no_decay if "bias" in name else dict_decay)['params'].append(param)
    opt = torch.optim.Adam([dict_decay, dict_no_decay], lr=model.lr)
    crit = nn.MSELoss(reduction='sum')
    rec = []
    orig = F.dropout

    def wrapped(x, p=0.5, training=True, inplace=False):
        m = orig(torch.ones_like(x), p=p, training=training)      # consumes the generator exactly like orig(x, ...)
        rec.append(m * (1 - p))
        return x * m
    # (1) the UNMODIFIED reference from this generator state: the fixture's targets
    import copy
    rng = torch.get_rng_state()
    ref_model = copy.deepcopy(model)
    ref_decay = {'params': [q for n, q in ref_model.nn.named_parameters() if "bias" not in n], 'weight_decay': model.l2_reg}
    ref_no = {'params': [q for n, q in ref_model.nn.named_parameters() if "bias" in n], 'weight_decay': 0.}
    ref_opt = torch.optim.Adam([ref_decay, ref_no], lr=model.lr)
    ref_model.nn.train()
    ref_losses = []
    for idx in batches:                                            # BNN_Dropout.py:57-62, verbatim
        ref_opt.zero_grad()
        loss = crit(ref_model.nn(X[idx]).squeeze(), y[idx])
        loss.backward()
        ref_opt.step()
        ref_losses.append(float(loss.detach()))
    # (2) the same run again from the same generator state with F.dropout wrapped to RECORD its draws
    torch.set_rng_state(rng)
    model.nn.train()
    losses = []
    F.dropout = wrapped
    try:
        for idx in batches:
            opt.zero_grad()
            pred_y = model.nn(X[idx]).squeeze()
            loss = crit(pred_y, y[idx])
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
    finally:
        F.dropout = orig
    for (W, b), (W1, b1) in zip(seq_to_net(model.nn.nn), seq_to_net(ref_model.nn.nn)):     # recording did not change the run
        assert torch.allclose(W, W1, rtol=1e-6, atol=1e-8) and torch.allclose(b, b1, rtol=1e-6, atol=1e-8)
    assert np.allclose(losses, ref_losses, rtol=1e-6)
    model = ref_model
    losses = ref_losses
    n_masked = 3                                                   # every Linear of 5-24-16-1 has an input wider than one unit
    assert len(rec) == len(batches) * n_masked
    keeps = [[rec[i * n_masked + l] for l in range(n_masked)] for i in range(len(batches))]
    o_net, o_loss = O.dropout_train_steps(net0, X, y, batches, "tanh", keeps, lr=1e-2, l2_reg=1e-3)
    for (W, b), (W1, b1) in zip(o_net, seq_to_net(model.nn.nn)):
        assert torch.allclose(W, W1, rtol=1e-5, atol=1e-7) and torch.allclose(b, b1, rtol=1e-5, atol=1e-7)
    assert np.allclose(o_loss, losses, rtol=1e-6)
    kw = {}
    for l, ((W0, b0), (W1, b1)) in enumerate(zip(net0, seq_to_net(model.nn.nn))):
        kw[f"W0_{l}"], kw[f"b0_{l}"], kw[f"W1_{l}"], kw[f"b1_{l}"] = W0, b0, W1, b1
        for i in range(len(batches)):
            kw[f"keep{i}_{l}"] = keeps[i][l]
    save("dropout_train.npz", dims=[5, 24, 16, 1], act="tanh", p_drop=p, l2_reg=1e-3, lr=1e-2, batch=B, X=X, y=y,
         perm=perm.numpy().astype(np.int64), loss=np.array(losses, dtype=np.float32), **kw)


def gen_cdropout_train():
    """f2: three minibatch steps of the REFERENCE's BNN_CDropout training recipe -- NN_CDropout.forward in training mode, MSELoss(sum),
    model.reg(), loss.backward(), Adam on every parameter (BNN_CDropout.py:114-133) -- on fixed batches, with the uniforms of every
    torch.rand_like call captured.  Distinct p_logits per layer and noise_level = 0.7 so that every term of the regulariser shows.
    Pins oracle.cdropout_train_steps and, on the GPU, the fused training kernel."""
    import copy
    torch.manual_seed(23)
    model = ref_cdo.BNN_CDropout(5, nn.Tanh(), [24, 16], conf={"lr": 1e-2, "lscale": 0.3, "dr": 1.5})
    model.noise_level = 0.7
    cls = [l for l in model.nn.nn if isinstance(l, ref_cdo.CDropoutLinear)]
    with torch.no_grad():
        for l, pl in zip(cls, (-2.0, -0.5, 0.4)):
            l.layer[0].p_logit.fill_(pl)
    N, B = 41, 16
    X = torch.randn(N, 5)
    y = torch.sin(X.sum(dim=1))
    perm = torch.randperm(N)
    batches = [perm[0:16], perm[16:32], perm[32:41]]
    net0 = [(l.layer[1].weight.detach().clone(), l.layer[1].bias.detach().clone()) for l in cls]
    pl0 = [l.layer[0].p_logit.detach().clone() for l in cls]
    num_train = N
    rng = torch.get_rng_state()

    def run(m, record):
        opt = torch.optim.Adam(m.nn.parameters(), lr=m.lr)              # BNN_CDropout.py:116
        crit = nn.MSELoss(reduction='sum')
        stats = []
        for idx in batches:                                            # BNN_CDropout.py:123-133, verbatim
            bx, by = X[idx], y[idx]
            opt.zero_grad()
            py = m.nn(bx).squeeze()
            mse_loss = crit(py, by)
            wreg, ent = m.reg()
            wreg *= bx.shape[0] / num_train
            ent *= bx.shape[0] / num_train
            loss = mse_loss + (wreg - ent)
            loss.backward()
            opt.step()
            stats.append((float(mse_loss.detach()), float(wreg.detach()), float(ent.detach())))
        return stats
    ref_model = copy.deepcopy(model)
    ref_stats = run(ref_model, False)                                  # (1) the unmodified reference: the fixture's targets
    torch.set_rng_state(rng)                                           # (2) the same run with torch.rand_like wrapped to record
    rec = []
    orig = torch.rand_like

    def wrapped(t, *a, **k):
        u = orig(t, *a, **k)
        rec.append(u.clone())
        return u
    torch.rand_like = wrapped
    try:
        stats = run(model, True)
    finally:
        torch.rand_like = orig
    assert stats == ref_stats
    rcls = [l for l in ref_model.nn.nn if isinstance(l, ref_cdo.CDropoutLinear)]
    net1 = [(l.layer[1].weight.detach().clone(), l.layer[1].bias.detach().clone()) for l in rcls]
    pl1 = [l.layer[0].p_logit.detach().clone() for l in rcls]
    for (W, b), l in zip(net1, cls):
        assert torch.equal(W, l.layer[1].weight) and torch.equal(b, l.layer[1].bias)
    nl = len(cls)
    assert len(rec) == len(batches) * nl
    us = [[rec[i * nl + l] for l in range(nl)] for i in range(len(batches))]
    o_net, o_pl, o_stats = O.cdropout_train_steps(net0, pl0, X, y, batches, "tanh", us, num_train, noise_level=0.7, lr=1e-2,
                                                  lscale=0.3, dr=1.5)
    for (W, b), (W1, b1) in zip(o_net, net1):
        assert torch.allclose(W, W1, rtol=1e-5, atol=1e-7) and torch.allclose(b, b1, rtol=1e-5, atol=1e-7)
    assert torch.allclose(torch.stack(o_pl), torch.stack(pl1), rtol=1e-5, atol=1e-7)
    assert np.allclose(np.array(o_stats), np.array(ref_stats), rtol=1e-6)
    kw = {}
    for l in range(nl):
        kw[f"W0_{l}"], kw[f"b0_{l}"], kw[f"W1_{l}"], kw[f"b1_{l}"] = net0[l][0], net0[l][1], net1[l][0], net1[l][1]
        for i in range(len(batches)):
            kw[f"u{i}_{l}"] = us[i][l]
    save("cdropout_train.npz", dims=[5, 24, 16, 1], act="tanh", lr=1e-2, lscale=0.3, dr=1.5, noise_level=0.7, batch=B, X=X, y=y,
         perm=perm.numpy().astype(np.int64), stats=np.array(ref_stats, dtype=np.float32),
         p_logit0=torch.stack(pl0), p_logit1=torch.stack(pl1), **kw)


if __name__ == "__main__":
    gen_sgdmc_boston()
    gen_sgdmc_multiout()
    gen_bbb_sinc()
    gen_dropout_protein()
    gen_cdropout_protein()
    gen_sgdmc_energy()
    gen_bbb_train()
    gen_dropout_train()
    gen_cdropout_train()
