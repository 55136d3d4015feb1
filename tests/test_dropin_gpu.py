"""The drop-in classes (bnn_b200.BNN_*) against the reference's recorded behaviour (tests/golden):
same constructor, same method names, same shapes, results within 1e-5 (FP32 path)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from helpers import golden, flat_to_nets, max_rel
from oracle import bnn_oracle as O
from oracle import philox as P

pytestmark = pytest.mark.gpu


def test_sgdmc_predict_validate_golden():
    from bnn_b200.BNN_SGDMC import BNN_SGDMC
    from bnn_b200.bank import SampleBank
    g = golden("sgdmc_boston.npz")
    model = BNN_SGDMC(13, nn.ReLU(), [50], nout=1, conf={})
    nets = flat_to_nets(g["W"], [13, 50, 1])
    model.nns = SampleBank(model.arch, model.arch.pack(nets).cuda())
    X = torch.from_numpy(g["X"])
    pred = model.sample_predict(model.nns, X)
    assert pred.shape == (100, 51, 1) and pred.device.type == "cpu"
    assert max_rel(pred, g["pred"]) <= 1e-5
    py, pv = model.predict_mv(X, model.nns)
    assert py.shape == (51, 1) and max_rel(py, g["mean"]) <= 1e-5 and max_rel(pv, g["var"]) <= 1e-5
    np.random.seed(5)                                         # same numpy stream as the recorded validate()
    rmse, nll = model.validate(X, torch.from_numpy(g["y_norm"]), num_samples=40)
    assert abs(float(rmse) - float(g["rmse"])) <= 1e-5 * abs(float(g["rmse"]))
    assert abs(float(nll) - float(g["nll"])) <= 1e-5 * abs(float(g["nll"]))
    with pytest.raises(AssertionError):                       # BNN_SGDMC.py:128
        model.sample(101)
    # elements are callable networks (BO.py:32,82)
    net7 = model.nns[7]
    with torch.no_grad():
        assert max_rel(net7(X), g["pred"][7]) <= 1e-5
    assert len(model.sample(10)) == 10


def test_sgdmc_accepts_reference_style_module_lists():
    from bnn_b200.BNN_SGDMC import BNN_SGDMC
    g = golden("sgdmc_multiout.npz")
    dims = [int(d) for d in g["dims"]]
    model = BNN_SGDMC(10, nn.Tanh(), [30, 30], nout=15, conf={})
    mods = [model.arch.to_module(v) for v in model.arch.pack(flat_to_nets(g["W"], dims))]
    X = torch.from_numpy(g["X"])
    pred = model.sample_predict(mods, X)                      # a plain Python list of nn.Sequential, as the reference passes
    assert pred.shape == (12, 37, 15) and max_rel(pred, g["pred"]) <= 1e-5
    arr = np.empty(len(mods), dtype=object)                   # ... or the object ndarray np.random.permutation returns
    for i, m in enumerate(mods):
        arr[i] = m
    py, pv = model.predict_mv(X, arr[np.random.permutation(len(mods))])
    assert max_rel(py, g["mean"]) <= 1e-5 and max_rel(pv, g["var"]) <= 1e-5


def test_bbb_golden():
    from bnn_b200.BNN_BBB import BNN_BBB
    g = golden("bbb_sinc.npz")
    model = BNN_BBB(1, nn.Tanh(), [50, 50], conf={"weight_std": float(g["weight_std"])})
    for l, gl in enumerate(model.nn.gaussian_layers()):
        gl.mu.data = torch.from_numpy(g[f"mu{l}"]); gl.rho.data = torch.from_numpy(g[f"rho{l}"])
    eps = model.arch.pack_bias_last([torch.from_numpy(g[f"eps{l}"]) for l in range(3)]).cuda()
    nns = model.sample(eps.shape[0], eps=eps)
    X = torch.from_numpy(g["X"])
    pred = model.sample_predict(nns, X)
    assert pred.shape == (8, 64) and max_rel(pred, g["pred"]) <= 1e-5
    py, pv = model.predict_mv(X, nns)
    assert max_rel(py, g["mean"]) <= 1e-5 and max_rel(pv, g["var"]) <= 1e-5
    assert abs(float(model.kl()) - float(g["kl"])) <= 1e-5 * abs(float(g["kl"]))
    # in-kernel Philox draws: fresh every call, reproducible from (seed, counter)
    a, b = model.sample(4), model.sample(4)
    assert not torch.equal(a.weights, b.weights)
    z = torch.from_numpy(P.stream_normal(model.seed, 8, P.STREAM_REPARAM, model.arch.stride))
    mu, rho = model._packed_mu_rho()
    ref = O.bbb_rsample(mu.cpu(), rho.cpu(), z) * model.arch.valid_mask()
    assert float((a.weights[0].cpu() - ref).abs().max()) < 2e-5


def test_dropout_golden():
    from bnn_b200.BNN_Dropout import BNN_Dropout
    g = golden("dropout_protein.npz")
    model = BNN_Dropout(9, nn.Tanh(), [100, 100], conf={"dropout_rate": float(g["p"])})
    base = flat_to_nets(g["base"][None], [9, 100, 100, 1])[0]
    for lin, (W, b) in zip([m for m in model.nn.nn if isinstance(m, nn.Linear)], base):
        lin.weight.data, lin.bias.data = W, b
    nns = model.sample(6, masks=torch.from_numpy(g["masks"]))
    X = torch.from_numpy(g["X"])
    pred = model.sample_predict(nns, X)
    assert pred.shape == (6, 96) and max_rel(pred, g["pred"]) <= 1e-5
    py, pv = model.predict_mv(X, nns)
    assert max_rel(py.squeeze(-1), g["mean"].squeeze(-1)) <= 1e-5
    # a materialised element equals the reference's sampled network (column-zeroed copy)
    with torch.no_grad():
        assert max_rel(nns[2](X).squeeze(-1), g["pred"][2]) <= 1e-5
    # own draws: Bernoulli(1-p) 0/1, correct keep rate, layer with in == 1 would be skipped
    own = model.sample(2000)
    assert set(np.unique(own.masks.cpu().numpy())) <= {0.0, 1.0}
    assert abs(float(own.masks.mean()) - (1 - float(g["p"]))) < 0.005
    rmse, nll = model.validate(X, torch.from_numpy(g["y"]), num_samples=50)
    assert rmse.shape == (1,) and torch.isfinite(rmse).all() and torch.isfinite(nll).all()


def test_cdropout_golden():
    from bnn_b200.BNN_CDropout import BNN_CDropout
    g = golden("cdropout_protein.npz")
    model = BNN_CDropout(9, nn.Tanh(), [100, 100], conf={})
    base = flat_to_nets(g["base"][None], [9, 100, 100, 1])[0]
    for cl, (W, b), pl in zip(model.nn.cd_layers(), base, g["p_logits"]):
        cl.layer[1].weight.data, cl.layer[1].bias.data = W, b
        cl.layer[0].p_logit.data.fill_(float(pl))
    nns = model.sample(5, uniforms=torch.from_numpy(g["uniforms"]))
    X = torch.from_numpy(g["X"])
    pred = model.sample_predict(nns, X)
    assert pred.shape == (5, 80) and max_rel(pred, g["pred"]) <= 2e-5     # soft masks: 1/T = 10 amplifies logf ulps
    own = model.sample(500)
    m = own.masks.cpu()
    assert float(m.min()) > 0 and float(m.max()) < 1
    keep = [1 - p for p in model.drop_probs()]
    for (o, n), k in zip([(0, 9), (9, 100), (109, 100)], keep):
        assert abs(float(m[:, o:o + n].mean()) - k) < 0.03               # relaxed Bernoulli mean ~ keep prob at T=0.1


def test_svi_draws():
    from bnn_b200.BNN_SVI import BNN_SVI
    model = BNN_SVI(4, nn.ReLU(), [20], conf={"seed": 11})
    nns = model.sample(3)
    for s in range(3):
        z = torch.from_numpy(P.stream_normal(11, s, P.STREAM_REPARAM, model.arch.stride))
        ref = O.svi_draw(model.mu, model.sigma, z) * model.arch.valid_mask()
        assert float((nns.weights[s].cpu() - ref).abs().max()) < 2e-5
    X = torch.randn(10, 4)
    assert model.sample_predict(nns, X).shape == (3, 10)


def test_sgdmc_train_short_chain():
    """End-to-end: the pSGLD chain (autograd gradient + fused update kernel) fits a toy regression."""
    from bnn_b200.BNN_SGDMC import BNN_SGDMC
    torch.manual_seed(0)
    X = torch.linspace(-1, 1, 200).view(-1, 1)
    y = torch.sin(3 * X).squeeze() + 0.05 * torch.randn(200)
    conf = dict(steps_burnin=1500, steps=500, keep_every=25, batch_size=32, lr_weight=1e-1, lr_noise=1e-1, seed=1)
    model = BNN_SGDMC(1, nn.Tanh(), [32], nout=1, conf=conf)
    model.train(X, y)
    assert len(model.nns) == 20
    rmse, nll = model.validate(X, y, num_samples=20)
    assert float(rmse) < 0.15          # the CPU oracle chain with the same recipe reaches ~0.07; y has std ~0.7
