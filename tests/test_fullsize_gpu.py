"""GPU parity at the FULL sizes of BASELINE.json configs 2, 3 and 5 (predictive), with the reference's draws injected.

cfg 2: BNN_BBB, 1-50-50-1 tanh, S = 1000 reparameterised draws, N = 1000 grid points       (BNN_BBB.py:24-27,139-149)
cfg 3: BNN_Dropout / BNN_CDropout, 9-100-100-1 tanh, S = 1000 masks, N = 4573 points        (BNN_Dropout.py:67-84,
                                                                                             BNN_CDropout.py:59-65,154-164)
cfg 5 (predictive): 90-512-512-512-1 tanh, S = 50 chain samples, N = 10,000 points          (BNN_SGDMC.py:131-137)

The draws (eps / Bernoulli masks / uniforms) are generated on the CPU in the order the reference consumes them (sample-major,
then layer, then row-major element: tests/golden/make_golden.py records exactly this order from the reference's own
sample()), handed to the drop-in classes through their injection arguments and to the oracle, and both run the same S x N
problem.  Tolerance: 1e-5 max-normalised on pred / mean for `fp32` and `tf32_bf16` alike (2e-5 on pred for the concrete masks,
whose 1/T = 10 amplifies 1-ulp logf differences); variance per helpers.var_bound.  The measured errors are appended to
gpurun_out/fullsize_errors.json (copied into profiles/ by hand)."""
import json
import os

import pytest
import torch
import torch.nn as nn

from helpers import max_rel, var_bound
from oracle import bnn_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(name, **errs):
    d = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(d):
        return
    path = os.path.join(d, "fullsize_errors.json")
    try:
        data = json.load(open(path))
    except Exception:
        data = {}
    data[name] = errs
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)


def check(name, pred, mean, var, pred_ref, N, tol_pred=1e-5, depth=2):
    S = pred_ref.shape[0]
    pred_ref = pred_ref.reshape(S, N, -1)
    mean_ref, var_ref = O.predict_mv(pred_ref, N)
    e_pred = max_rel(pred.reshape(S, N, -1), pred_ref)
    e_mean = max_rel(mean.reshape(N, -1), mean_ref)
    dv = (var.reshape(N, -1).double() - var_ref.double()).abs()
    vb = var_bound(pred_ref, depth=depth)
    e_var = float((dv / vb).max())
    record(name, pred=e_pred, mean=e_mean, var_over_bound=e_var, var_max_norm=max_rel(var.reshape(N, -1), var_ref))
    assert e_pred <= tol_pred, (name, e_pred)
    assert e_mean <= 1e-5, (name, e_mean)
    assert e_var <= 1.0, (name, e_var)


@pytest.mark.parametrize("precision", ["fp32", "tf32_bf16"])
def test_cfg2_bbb_full(precision):
    from bnn_b200.BNN_BBB import BNN_BBB
    S, N = 1000, 1000
    torch.manual_seed(21)
    model = BNN_BBB(1, nn.Tanh(), [50, 50], conf={"precision": precision})
    gls = model.nn.gaussian_layers()
    for l in gls:                                     # off the init point (a trained posterior is not at rho = -5 either)
        l.mu.data += 0.3 * torch.randn_like(l.mu)
        l.rho.data += 1.0 * torch.randn_like(l.rho)
    g = torch.Generator().manual_seed(22)
    eps = [[torch.randn(l.mu.shape, generator=g) for l in gls] for _ in range(S)]          # sample-major, then layer
    X = torch.linspace(-3, 3, N).view(-1, 1)
    nets = [[O.bbb_split(O.bbb_rsample(l.mu.detach(), l.rho.detach(), eps[s][li])) for li, l in enumerate(gls)] for s in range(S)]
    pred_ref = O.sample_predict(nets, X, "tanh")
    eps_packed = model.arch.pack_bias_last([torch.stack([eps[s][li] for s in range(S)]) for li in range(3)]).cuda()
    nns = model.sample(S, eps=eps_packed)
    pred = model.sample_predict(nns, X)
    mean, var = model.predict_mv(X, nns)
    assert pred.shape == (S, N)
    check(f"cfg2_bbb_S{S}_N{N}_{precision}", pred, mean, var, pred_ref, N)


@pytest.mark.parametrize("precision", ["fp32", "tf32_bf16"])
def test_cfg3_dropout_full(precision):
    from bnn_b200.BNN_Dropout import BNN_Dropout
    S, N, p = 1000, 4573, 0.05
    torch.manual_seed(31)
    model = BNN_Dropout(9, nn.Tanh(), [100, 100], conf={"dropout_rate": p, "precision": precision})
    g = torch.Generator().manual_seed(32)
    X = torch.randn(N, 9, generator=g)
    ins = [9, 100, 100]
    # F.dropout(ones(in), p) * (1 - p) == 0/1 keep mask, one per (sample, layer), layer order (BNN_Dropout.py:70-75)
    masks = [[torch.empty(i).bernoulli_(1 - p, generator=g) for i in ins] for _ in range(S)]
    base = [(l.weight.detach().clone(), l.bias.detach().clone()) for l in model.nn.nn if isinstance(l, nn.Linear)]
    nets = [O.scale_columns(base, masks[s], skip_unit_inputs=True) for s in range(S)]
    pred_ref = O.sample_predict(nets, X, "tanh")
    nns = model.sample(S, masks=torch.stack([torch.cat(m) for m in masks]))
    pred = model.sample_predict(nns, X)
    mean, var = model.predict_mv(X, nns)
    check(f"cfg3_dropout_S{S}_N{N}_{precision}", pred, mean, var, pred_ref, N)


@pytest.mark.parametrize("precision", ["fp32", "tf32_bf16"])
def test_cfg3_cdropout_full(precision):
    from bnn_b200.BNN_CDropout import BNN_CDropout
    S, N = 1000, 4573
    torch.manual_seed(41)
    model = BNN_CDropout(9, nn.Tanh(), [100, 100], conf={"precision": precision})
    cls = model.nn.cd_layers()
    for i, l in enumerate(cls):
        l.layer[0].p_logit.data.fill_([-2.0, 0.0, 1.0][i])
    g = torch.Generator().manual_seed(42)
    X = torch.randn(N, 9, generator=g)
    ins = [9, 100, 100]
    uniforms = [[torch.rand(i, generator=g) for i in ins] for _ in range(S)]          # torch.rand inside RelaxedBernoulli.sample
    base = [(l.layer[1].weight.detach().clone(), l.layer[1].bias.detach().clone()) for l in cls]
    p_logits = [float(l.layer[0].p_logit) for l in cls]
    nets = [O.scale_columns(base, [O.concrete_keep_mask(uniforms[s][l], p_logits[l]) for l in range(3)], skip_unit_inputs=False)
            for s in range(S)]
    pred_ref = O.sample_predict(nets, X, "tanh")
    nns = model.sample(S, uniforms=torch.stack([torch.cat(u) for u in uniforms]))
    pred = model.sample_predict(nns, X)
    mean, var = model.predict_mv(X, nns)
    check(f"cfg3_cdropout_S{S}_N{N}_{precision}", pred, mean, var, pred_ref, N, tol_pred=2e-5)


@pytest.mark.parametrize("precision", ["fp32", "auto"])
def test_cfg5_predictive(precision):
    """cfg 5's own network through BNN_SGDMC.sample_predict / predict_mv: 50 chain states, 10k points."""
    from bnn_b200.BNN_SGDMC import BNN_SGDMC
    from bnn_b200.bank import SampleBank
    S, N = 50, 10_000
    torch.manual_seed(51)
    model = BNN_SGDMC(90, nn.Tanh(), [512, 512, 512], nout=1, conf={"precision": precision})
    base = model.arch.pack([model.nn.nn])[0]
    g = torch.Generator().manual_seed(52)
    bank = (base[None] + 0.02 * torch.randn(S, model.arch.stride, generator=g)) * model.arch.valid_mask()
    model.nns = SampleBank(model.arch, bank.cuda())
    X = torch.randn(N, 90, generator=g)
    nets = [model.arch.unpack(bank[s]) for s in range(S)]
    pred_ref = O.sample_predict(nets, X, "tanh", nout=1)
    pred = model.sample_predict(model.nns, X)
    mean, var = model.predict_mv(X, model.nns)
    assert pred.shape == (S, N, 1)
    check(f"cfg5_predictive_S{S}_N{N}_{precision}", pred, mean, var, pred_ref, N, depth=4)
