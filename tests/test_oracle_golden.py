"""CPU: the oracle reproduces every golden fixture recorded from the unmodified reference
(tests/golden/make_golden.py) -- this is what pins the oracle."""
import numpy as np
import torch

from helpers import golden, flat_to_nets
from oracle import bnn_oracle as O
from oracle import philox as P


def t(a):
    return torch.from_numpy(np.asarray(a))


def test_sgdmc_boston():
    g = golden("sgdmc_boston.npz")
    nets = flat_to_nets(g["W"], [13, 50, 1])
    X = t(g["X"])
    pred = O.sample_predict(nets, X, "relu", nout=1)
    assert torch.equal(pred, t(g["pred"]))
    py, pv = O.predict_mv(pred, 51)
    assert torch.equal(py, t(g["mean"])) and torch.equal(pv, t(g["var"]))
    sub = [nets[i] for i in g["perm"]]
    spy, spv = O.predict_mv(O.sample_predict(sub, X, "relu", nout=1), 51)
    rmse, nll = O.validate_metrics(spy, spv, t(g["y_norm"]))
    assert torch.allclose(rmse, t(g["rmse"]), rtol=1e-6) and torch.allclose(nll, t(g["nll"]), rtol=1e-6)
    # experiments/SGDMC.py:81-93 recipe
    ys, ym = float(g["ys"]), float(g["ym"])
    pd = pred * ys + ym
    pyd, pvd = O.predict_mv(pd, 51)
    nv = torch.tensor([ys ** 2 / float(np.exp(g["log_prec"][0]))])
    rmse_d, nll_d = O.validate_metrics(pyd, pvd, t(g["y_raw"]), noise_var=nv)
    assert torch.allclose(rmse_d, t(g["rmse_denorm"]), rtol=1e-5) and torch.allclose(nll_d, t(g["nll_denorm"]), rtol=1e-5)


def test_sgdmc_multiout():
    g = golden("sgdmc_multiout.npz")
    nets = flat_to_nets(g["W"], [10, 30, 30, 15])
    pred = O.sample_predict(nets, t(g["X"]), "tanh", nout=15)
    assert torch.equal(pred, t(g["pred"]))
    rmse, nll = O.validate_metrics(*O.predict_mv(pred, 37), t(g["y"]))
    assert torch.allclose(rmse, t(g["rmse"]), rtol=1e-6) and torch.allclose(nll, t(g["nll"]), rtol=1e-6)


def test_bbb():
    g = golden("bbb_sinc.npz")
    S = g["eps0"].shape[0]
    nets = []
    for s in range(S):
        nets.append([O.bbb_split(O.bbb_rsample(t(g[f"mu{l}"]), t(g[f"rho{l}"]), t(g[f"eps{l}"][s]))) for l in range(3)])
        for l in range(3):
            W, b = nets[-1][l]
            assert torch.equal(torch.cat([W, b[:, None]], dim=1), t(g[f"wb{l}"][s]))
    assert torch.equal(O.sample_predict(nets, t(g["X"]), "tanh"), t(g["pred"]))
    kl = O.bbb_kl([t(g[f"mu{l}"]) for l in range(3)], [t(g[f"rho{l}"]) for l in range(3)], float(g["weight_std"]))
    assert torch.allclose(kl, t(g["kl"]), rtol=1e-6)


def test_dropout_and_cdropout():
    g = golden("dropout_protein.npz")
    base = flat_to_nets(g["base"][None], [9, 100, 100, 1])[0]
    nets = [O.scale_columns(base, [t(g["masks"][s, o:o + n]) for o, n in [(0, 9), (9, 100), (109, 100)]], True)
            for s in range(g["masks"].shape[0])]
    assert torch.equal(O.sample_predict(nets, t(g["X"]), "tanh"), t(g["pred"]))
    c = golden("cdropout_protein.npz")
    base = flat_to_nets(c["base"][None], [9, 100, 100, 1])[0]
    nets = []
    for s in range(c["uniforms"].shape[0]):
        ms = [O.concrete_keep_mask(t(c["uniforms"][s, o:o + n]), float(pl)) for (o, n), pl in
              zip([(0, 9), (9, 100), (109, 100)], c["p_logits"])]
        assert torch.equal(torch.cat(ms), t(c["masks"][s]))
        nets.append(O.scale_columns(base, ms, False))
    assert torch.equal(O.sample_predict(nets, t(c["X"]), "tanh"), t(c["pred"]))


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    h = lambda v: [int(x) for x in v]
    assert h(P.philox4x32_10(0, 0, 0, 0, 0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert h(P.philox4x32_10(f, f, f, f, f, f)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert h(P.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    z = P.stream_normal(7, 3, P.STREAM_REPARAM, 200000)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    k = P.bernoulli_keep(1, 2, 100000, 0.05)
    assert abs(k.mean() - 0.95) < 0.003


def test_merge_moments_equals_single_pass():
    g = torch.Generator().manual_seed(0)
    x = (5 + 0.01 * torch.randn(1000, 33, generator=g)).double().numpy()     # the SURVEY 0.2 stress case
    parts = [O.shard_moments(x[a:b]) for a, b in [(0, 1), (1, 400), (400, 1000)]]
    n, mean, m2 = O.merge_moments(parts)
    assert n == 1000
    assert np.allclose(mean, x.mean(0), rtol=1e-14) and np.allclose(m2 / (n - 1), x.var(0, ddof=1), rtol=1e-10)


def test_psgld_restatement_and_schedule():
    th, V = torch.tensor([1.0, -2.0]), torch.tensor([1.0, 4.0])
    g, xi = torch.tensor([0.5, 0.5]), torch.tensor([0.0, 1.0])
    th2, V2 = O.psgld_step(th, g, V, 0.01, xi, alpha=0.99, lam=1e-5)
    Vexp = 0.99 * V + 0.01 * g * g
    G = 1 / (1e-5 + Vexp.sqrt())
    assert torch.allclose(V2, Vexp) and torch.allclose(th2, th - (0.005 * G * g + (0.01 * G).sqrt() * xi))
    assert O.lr_factor(0) == np.float32(1.0) and abs(float(O.lr_factor(9)) - 10 ** -0.33) < 1e-7


def test_sgdmc_energy_terms_match_oracle():
    """BNN_SGDMC.log_prior / log_lik (kept for API parity with BNN_SGDMC.py:61-74) against the oracle's energy:
    U = -(log_lik * N / |b| + log_prior)."""
    import torch.nn as nn
    from bnn_b200.BNN_SGDMC import BNN_SGDMC
    torch.manual_seed(4)
    m = BNN_SGDMC(5, nn.Tanh(), [7, 6], nout=2, conf={})
    m.log_precs.data = torch.tensor([0.3, -0.4])
    X, y = torch.randn(11, 5), torch.randn(11, 2)
    net = [(l.weight.detach(), l.bias.detach()) for l in m.nn.nn if isinstance(l, nn.Linear)]
    u_ref = O.sgdmc_neg_log_post(net, m.log_precs.detach(), X, y, "tanh", num_train=40, alpha_n=m.alpha_n, beta_n=m.beta_n)
    u = -(m.log_lik(X, y) * (40 / 11) + m.log_prior())
    assert abs(float(u) - float(u_ref)) <= 1e-5 * abs(float(u_ref))


def test_sgdmc_energy_golden():
    """oracle.sgdmc_neg_log_post and its autograd gradient against the reference's own log_prior / log_lik / loss /
    loss.backward() (BNN_SGDMC.py:61-74,83-85) recorded in sgdmc_energy.npz -- pins the energy half of row a12."""
    from helpers import energy_cases, check_energy_grads
    for c in energy_cases():
        ps = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in c["net"]]
        lp = c["log_precs"].clone().requires_grad_(True)
        U = O.sgdmc_neg_log_post(ps, lp, c["X"], c["y"], c["act"], c["num_train"], float(c["alpha_n"]), float(c["beta_n"]))
        assert abs(float(U) - float(c["loss"])) <= 2e-6 * abs(float(c["loss"])), c["name"]
        gs = torch.autograd.grad(U, [q for pr in ps for q in pr] + [lp])
        check_energy_grads(c, [(gs[2 * i], gs[2 * i + 1]) for i in range(len(ps))], gs[-1])


def test_bbb_train_steps_golden():
    """f2: the oracle's restatement of the BNN_BBB training step against the reference's own three steps."""
    from helpers import bbb_train_case
    c = bbb_train_case()
    mu, rho, mse, kl = O.bbb_train_steps(c["mu0"], c["rho0"], c["X"], c["y"], c["batches"], c["act"], c["eps"], lr=c["lr"],
                                         kl_factor=c["kl_factor"], weight_std=c["weight_std"])
    for l in range(len(mu)):
        assert torch.allclose(mu[l], c["mu1"][l], rtol=1e-6, atol=1e-7)
        assert torch.allclose(rho[l], c["rho1"][l], rtol=1e-6, atol=1e-7)
    assert np.allclose(mse, c["mse"], rtol=1e-6) and np.allclose(kl, c["kl"], rtol=1e-6)


def test_dropout_train_steps_golden():
    """f2: the oracle's restatement of the BNN_Dropout training step against the reference's own three steps."""
    from helpers import dropout_train_case
    c = dropout_train_case()
    net, loss = O.dropout_train_steps(c["net0"], c["X"], c["y"], c["batches"], c["act"], c["keeps"], lr=c["lr"], l2_reg=c["l2_reg"])
    for (W, b), (W1, b1) in zip(net, c["net1"]):
        assert torch.allclose(W, W1, rtol=1e-6, atol=1e-7) and torch.allclose(b, b1, rtol=1e-6, atol=1e-7)
    assert np.allclose(loss, c["loss"], rtol=1e-6)
    for i, ks in enumerate(c["keeps"]):                      # F.dropout(x, p) * (1 - p): keep factors are 0 or (an ulp from) 1
        for k in ks:
            assert k.shape[0] == c["batches"][i].shape[0]
            assert bool(((k == 0) | ((k - 1).abs() < 1e-6)).all())


def test_cdropout_train_steps_golden():
    """f2: the oracle's restatement of the BNN_CDropout training step against the reference's own three steps."""
    from helpers import cdropout_train_case
    c = cdropout_train_case()
    net, pl, stats = O.cdropout_train_steps(c["net0"], list(c["p_logit0"]), c["X"], c["y"], c["batches"], c["act"], c["us"],
                                            c["X"].shape[0], noise_level=c["noise_level"], lr=c["lr"], lscale=c["lscale"], dr=c["dr"])
    for (W, b), (W1, b1) in zip(net, c["net1"]):
        assert torch.allclose(W, W1, rtol=1e-6, atol=1e-7) and torch.allclose(b, b1, rtol=1e-6, atol=1e-7)
    assert torch.allclose(torch.stack(pl), c["p_logit1"], rtol=1e-6, atol=1e-7)
    assert np.allclose(np.array(stats), c["stats"], rtol=1e-6)


def test_cdropout_closed_form_gradients_match_the_reference():
    """The closed-form gradients of the fused concrete-dropout kernel (keep factors, regulariser, p_logit; no autograd), restated in
    float64, reproduce the reference's three recorded steps: pins the kernel's derivation on the CPU."""
    from helpers import cdropout_train_case
    c = cdropout_train_case()
    net, pl, stats = O.cdropout_train_steps_closed_form(c["net0"], list(c["p_logit0"]), c["X"], c["y"], c["batches"], c["act"], c["us"],
                                                        c["X"].shape[0], noise_level=c["noise_level"], lr=c["lr"], lscale=c["lscale"],
                                                        dr=c["dr"])
    for (W, b), (W1, b1) in zip(net, c["net1"]):
        assert float((W.float() - W1).abs().max()) <= 2e-6 * float(W1.abs().max())
        assert float((b.float() - b1).abs().max()) <= 2e-6 * float(b1.abs().max()) + 1e-8
    assert float((pl.float() - c["p_logit1"]).abs().max()) <= 2e-6 * float(c["p_logit1"].abs().max())
    assert np.allclose(np.array(stats), c["stats"], rtol=2e-6)


def test_bbb_closed_form_gradients_match_the_reference():
    """The closed-form gradients of the fused Bayes-by-Backprop kernel (local reparameterisation backward, KL gradient; no autograd),
    restated in float64, reproduce the reference's three recorded steps."""
    from helpers import bbb_train_case
    c = bbb_train_case()
    mu, rho, mse, kl = O.bbb_train_steps_closed_form(c["mu0"], c["rho0"], c["X"], c["y"], c["batches"], c["act"], c["eps"], lr=c["lr"],
                                                     kl_factor=c["kl_factor"], weight_std=c["weight_std"])
    for l in range(len(mu)):
        assert float((mu[l].float() - c["mu1"][l]).abs().max()) <= 2e-6 * float(c["mu1"][l].abs().max())
        assert float((rho[l].float() - c["rho1"][l]).abs().max()) <= 2e-6 * float(c["rho1"][l].abs().max())
    assert np.allclose(mse, c["mse"], rtol=2e-6) and np.allclose(kl, c["kl"], rtol=2e-6)
