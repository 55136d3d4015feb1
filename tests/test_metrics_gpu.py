"""GPU parity of kernel (e): standalone moments, Chan merge, rmse/nll (both variants)."""
import numpy as np
import pytest
import torch

from helpers import golden, max_rel
from oracle import bnn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S,M", [(100, 51), (1000, 4573), (7, 1), (3, 1000003), (2000, 8), (1, 5)])
def test_moments_vs_torch(S, M):
    from bnn_b200 import ops
    g = torch.Generator().manual_seed(S * 7 + M)
    pred = torch.randn(S, M, generator=g) * 0.3 + torch.randn(M, generator=g)
    mean, var = ops.moments(pred.cuda())
    assert max_rel(mean.cpu(), pred.mean(dim=0)) <= 1e-6
    if S > 1:
        assert max_rel(var.cpu(), pred.var(dim=0)) <= 1e-5
    else:
        assert bool(torch.isnan(var).all())


def test_merge_matches_single_shard():
    from bnn_b200 import ops
    S, M = 103, 777
    pred = torch.randn(S, M, generator=torch.Generator().manual_seed(2)) + 3.0
    cuts = [0, 10, 11, 60, 103]
    pm, pm2, cnt = [], [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        _, _, m, m2 = ops.moments(pred[a:b].cuda(), want_partials=True)
        pm.append(m); pm2.append(m2); cnt.append(b - a)
    mean, var = ops.moments_merge(torch.stack(pm), torch.stack(pm2), cnt)
    assert max_rel(mean.cpu(), pred.mean(dim=0)) <= 1e-6
    assert max_rel(var.cpu(), pred.var(dim=0)) <= 1e-5
    # oracle's Chan merge on the same shards
    parts = [O.shard_moments(pred[a:b].double().numpy()) for a, b in zip(cuts[:-1], cuts[1:])]
    n, mu, m2 = O.merge_moments(parts)
    assert n == S and max_rel(var.cpu(), m2 / (n - 1)) <= 1e-6


def test_metrics_golden_boston():
    from bnn_b200 import ops
    g = golden("sgdmc_boston.npz")
    perm = g["perm"]
    pred = torch.from_numpy(g["pred"])[perm]                    # BNN.validate draws 40 of the 100 nets
    mean, var = ops.moments(pred.cuda())
    rmse, nll = ops.metrics(mean, var, torch.from_numpy(g["y_norm"]).cuda())
    assert abs(float(rmse) - float(g["rmse"])) <= 1e-5 * abs(float(g["rmse"]))
    assert abs(float(nll) - float(g["nll"])) <= 1e-5 * abs(float(g["nll"]))
    # experiments/SGDMC.py:81-93: de-normalised predictions + learned noise variance
    ys, ym = float(g["ys"]), float(g["ym"])
    pred_d = torch.from_numpy(g["pred"]) * ys + ym
    mean_d, var_d = ops.moments(pred_d.cuda())
    noise_var = ys ** 2 / float(np.exp(g["log_prec"][0]))
    rmse_d, nll_d = ops.metrics(mean_d, var_d, torch.from_numpy(g["y_raw"]).cuda(), noise_var=[noise_var])
    assert abs(float(rmse_d) - float(g["rmse_denorm"])) <= 1e-5 * abs(float(g["rmse_denorm"]))
    assert abs(float(nll_d) - float(g["nll_denorm"])) <= 1e-5 * abs(float(g["nll_denorm"]))


def test_metrics_multiout_and_error():
    from bnn_b200 import ops
    g = golden("sgdmc_multiout.npz")
    mean, var = torch.from_numpy(g["mean"]).cuda(), torch.from_numpy(g["var"]).cuda()
    rmse, nll = ops.metrics(mean, var, torch.from_numpy(g["y"]).cuda())
    assert max_rel(rmse, g["rmse"]) <= 1e-5 and max_rel(nll, g["nll"]) <= 1e-5
    with pytest.raises(ValueError):                              # BNN.py:31 raises on scale <= 0
        ops.metrics(mean, torch.zeros_like(var), torch.from_numpy(g["y"]).cuda())
