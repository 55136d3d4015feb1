"""GPU: the driver layer end to end on small synthetic problems, with its metric formulas checked
against the reference's own expressions evaluated on the same predictions (experiments/SGDMC.py:81-96, MO_SGDMC.py:48-63)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from test_drivers_cpu import make_uci_dir

pytestmark = pytest.mark.gpu


def test_uci_sgdmc_driver_metrics(tmp_path):
    from bnn_b200.experiments import uci_sgdmc, get_data
    from bnn_b200.util import normalize
    make_uci_dir(str(tmp_path), n=200, d=4)
    res = uci_sgdmc("toy", 0, str(tmp_path), conf=dict(steps_burnin=400, steps=200, keep_every=20, batch_size=32, seed=3))
    model = res["model"]
    assert len(model.nns) == 10 and torch.isfinite(res["rmse"]).all() and torch.isfinite(res["nll"]).all()
    # the reference's recipe on the same samples, in torch on the CPU (experiments/SGDMC.py:81-96)
    trx, try_, tex, tey, *_ = get_data("toy", 0, str(tmp_path))
    xm, xs, ym, ys = normalize(trx, try_)
    preds = model.sample_predict(model.nns, (tex - xm) / xs) * ys + ym
    py = preds.mean(dim=0).squeeze()
    pv = preds.var(dim=0).squeeze() + ys ** 2 / model.log_precs.exp().squeeze()
    rmse = torch.mean((py - tey) ** 2).sqrt()
    nll = -1 * torch.distributions.Normal(py, pv.sqrt()).log_prob(tey).mean()
    smse = rmse ** 2 / torch.mean((tey - try_.mean()) ** 2)
    assert abs(float(res["rmse"]) - float(rmse)) <= 1e-5 * abs(float(rmse))
    assert abs(float(res["nll"]) - float(nll)) <= 1e-4 * max(1.0, abs(float(nll)))
    assert abs(float(res["smse"]) - float(smse)) <= 1e-4 * abs(float(smse))
    assert float(res["rmse"]) < 0.6 * float(tey.std()) + 0.3             # the chain learned the linear toy problem


def test_mo_metrics_match_reference_formulas():
    from bnn_b200.BNN_SGDMC import BNN_SGDMC
    from bnn_b200.bank import SampleBank
    from bnn_b200.experiments import mo_metrics
    torch.manual_seed(0)
    model = BNN_SGDMC(dim=6, nout=5, act=nn.Tanh(), num_hiddens=[20, 20], conf={})
    base = model.arch.pack([model.nn.nn])[0]
    model.nns = SampleBank(model.arch, ((base[None] + 0.1 * torch.randn(30, model.arch.stride)) * model.arch.valid_mask()).cuda())
    test_x, test_y = torch.randn(80, 6), torch.randn(80, 5) * 2 + 1
    y_mean, y_std = torch.randn(5), torch.rand(5) + 0.5
    res = mo_metrics(model, model.nns, test_x, test_y, y_mean, y_std)
    pred = model.sample_predict(model.nns, test_x) * y_std + y_mean              # MO_SGDMC.py:48-63
    py, ps = pred.mean(dim=0), pred.std(dim=0)
    mse = torch.mean((py - test_y) ** 2, dim=0)
    smse = mse / torch.mean((test_y - y_mean) ** 2, dim=0)
    ll = torch.distributions.Normal(py, ps).log_prob(test_y).mean(dim=0)
    ll_ref = torch.distributions.Normal(y_mean, y_std).log_prob(test_y).mean(dim=0)
    assert torch.allclose(res["smse"], smse, rtol=1e-4) and torch.allclose(res["nll"], ll_ref - ll, rtol=1e-4, atol=1e-4)
    assert torch.allclose(res["py"].cpu(), py, rtol=1e-4, atol=1e-5) and torch.allclose(res["ps"].cpu(), ps, rtol=1e-4, atol=1e-6)


def test_uci_validate_dropout(tmp_path):
    from bnn_b200.BNN_Dropout import BNN_Dropout
    from bnn_b200.experiments import uci_validate
    make_uci_dir(str(tmp_path), n=120, d=3)
    res = uci_validate(BNN_Dropout, "toy", 1, str(tmp_path), conf=dict(num_epochs=30, batch_size=32), num_samples=50)
    assert torch.isfinite(res["rmse"]).all() and torch.isfinite(res["nll"]).all() and float(res["smse"]) < 1.5


def test_bo_loop_improves_a_quadratic():
    """BO.OSFTA on f(x) = (x - 0.5)^2: every iteration trains the chain, minimises sampled nets with batched NSGA-II generations
    through the CUDA forward, and appends the suggestion."""
    from bnn_b200.BO import BO
    torch.manual_seed(0)
    f = lambda X: ((X - 0.5) ** 2).sum(dim=1, keepdim=True)
    conf = dict(steps_burnin=300, steps=100, keep_every=20, batch_size=8, lr_weight=5e-2, lr_noise=5e-2, seed=2,
                nsga_population=20, nsga_evals=400)
    bo = BO(f, dim=1, nobj=1, ncons=0, max_eval=3, num_init=6, act=nn.Tanh(), num_hiddens=[16], conf=conf)
    y0 = float(bo.y.min())
    bo.OSFTA()
    assert bo.X.shape == (9, 1) and bo.y.shape == (9, 1)
    assert float(bo.X.min()) >= -1.7321 and float(bo.X.max()) <= 1.7321
    assert float(bo.y.min()) <= y0 + 1e-6
    xs, curves = bo.see(torch.tensor([0]))
    assert curves.shape == (len(bo.bnn.nns), 100)
