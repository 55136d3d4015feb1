"""Data / driver layer of the reference's experiments, on top of the CUDA hot path.

Mirrors, function by function:
  * ``get_data``            experiments/SGDMC.py:23-47 (identical copies in BBB/Dropout/CDropout/SVI.py): whitespace text matrices
                            under ``<root>/<dataset>/data`` -- data.txt (last column = target), index_train_k.txt / index_test_k.txt,
                            n_splits.txt, n_hidden.txt, n_epochs.txt -- and ``results/test_tau_100_xepochs_1_hidden_layers.txt``;
  * ``uci_sgdmc``           experiments/SGDMC.py:50-99: normalise X and y, train the pSGLD chain, predict with ALL chain samples,
                            de-normalise (``* ys + ym``), add the learned noise variance ``ys**2 / exp(log_prec)``, RMSE / NLL / SMSE;
  * ``uci_validate``        experiments/{Dropout,BBB,CDropout,SVI}.py:50-76: normalise X only, train, ``validate(num_samples=100)``;
  * ``mo_sgdmc``            experiments/MO_SGDMC.py:14-65: multi-output regression (nout > 1), per-output SMSE and normalised LL.
The predictive statistics never materialise on the host: the forward kernel reduces pred[S, N, nout] to per-point moments on the
device, the affine de-normalisation is applied to the moments (mean * ys + ym, var * ys**2 -- the same quantities the reference gets
by transforming every prediction first) and RMSE / NLL come from the metrics kernel.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .util import normalize


def get_data(dataset, split_id, root):
    """experiments/SGDMC.py:23-47.  Returns (train_x, train_y, test_x, test_y, tau, n_hidden, n_splits, n_epochs); the first
    six are None when split_id >= n_splits, exactly like the reference."""
    d = os.path.join(root, dataset)
    data = np.loadtxt(os.path.join(d, "data", "data.txt"))
    xs, ys = data[:, :-1], data[:, -1]
    n_splits = int(np.loadtxt(os.path.join(d, "data", "n_splits.txt"), dtype=np.int64))
    if split_id >= n_splits:
        return None, None, None, None, None, None, n_splits, None
    train_id = np.loadtxt(os.path.join(d, "data", "index_train_%d.txt" % split_id), dtype=np.int64)
    test_id = np.loadtxt(os.path.join(d, "data", "index_test_%d.txt" % split_id), dtype=np.int64)
    tau_path = os.path.join(d, "results", "test_tau_100_xepochs_1_hidden_layers.txt")
    tau = float(np.atleast_1d(np.loadtxt(tau_path))[split_id]) if os.path.exists(tau_path) else float("nan")
    n_hidden = int(np.loadtxt(os.path.join(d, "data", "n_hidden.txt"), dtype=np.int64))
    n_epochs = int(np.loadtxt(os.path.join(d, "data", "n_epochs.txt"), dtype=np.int64))
    f = lambda a: torch.FloatTensor(a)
    return f(xs[train_id]), f(ys[train_id]), f(xs[test_id]), f(ys[test_id]), tau, n_hidden, n_splits, n_epochs


def sgdmc_conf():
    """The UCI recipe of experiments/SGDMC.py:61-72."""
    return dict(batch_size=128, steps_burnin=2500, steps=2500, keep_every=50, lr_weight=1e-1, lr_noise=2e-2, alpha_n=0.5, beta_n=0.4)


def denormalised_metrics(model, nns, test_x, test_y, ym, ys, noise_var=None):
    """experiments/SGDMC.py:81-93 on the device: moments of ``pred * ys + ym`` over the samples, plus ``noise_var`` (already in
    target units, per output) inside the NLL only.  Returns (rmse[nout], nll[nout], mean[N, nout], var[N, nout])."""
    out = model._forward(nns, test_x, want_pred=False, want_moments=True)
    dev = out["mean"].device
    ys_t = torch.as_tensor(ys, dtype=torch.float32, device=dev).reshape(1, -1)
    ym_t = torch.as_tensor(ym, dtype=torch.float32, device=dev).reshape(1, -1)
    mean = (out["mean"] * ys_t + ym_t).contiguous()
    var = (out["var"] * ys_t * ys_t).contiguous()
    y = model._to_dev(test_y).reshape(mean.shape[0], -1)
    rmse, nll = ops.metrics(mean, var, y, noise_var=noise_var)
    return rmse, nll, mean, var


def uci_sgdmc(dataset, split_id, root, conf=None, act=None, model_conf=None):
    """experiments/SGDMC.py:50-99.  Returns dict(rmse, nll, smse, precision, model)."""
    from .BNN_SGDMC import BNN_SGDMC
    train_x, train_y, test_x, test_y, tau, n_hiddens, n_splits, n_epochs = get_data(dataset, split_id, root)
    if split_id >= n_splits:
        print("Invalid split_id")
        return dict(rmse=np.nan, nll=np.nan, smse=np.nan, precision=np.nan, model=None)
    print('Dataset %s, split: %d, n_hiddens: %d, prec: %g' % (dataset, split_id, n_hiddens, tau))
    xm, xs, ym, ys = normalize(train_x, train_y)
    train_xn, test_xn, train_yn = (train_x - xm) / xs, (test_x - xm) / xs, (train_y - ym) / ys
    c = sgdmc_conf()
    c['num_epochs'] = 100 * n_epochs
    c.update(conf or {})
    c.update(model_conf or {})
    model = BNN_SGDMC(train_x.shape[1], act if act is not None else nn.Tanh(), num_hiddens=[n_hiddens, n_hiddens], conf=c)
    model.train(train_xn, train_yn)
    model.report()
    noise_var = float(ys) ** 2 / model.log_precs.exp().reshape(-1)            # ys**2 / exp(log_prec), :84
    rmse, nll, _, _ = denormalised_metrics(model, model.nns, test_xn, test_y, ym, ys, noise_var=noise_var)
    smse = rmse ** 2 / torch.mean((test_y - train_y.mean()) ** 2)             # :96
    print('RMSE = %g, SMSE = %g, NLL = %6.3f' % (float(rmse), float(smse), float(nll)), flush=True)
    print('Model  precision: %g' % float(model.log_precs.exp().mean()))
    return dict(rmse=rmse, nll=nll, smse=smse, precision=model.log_precs.exp(), model=model)


def uci_validate(model_cls, dataset, split_id, root, conf=None, act=None, num_samples=100):
    """experiments/Dropout.py:50-76 (same shape in BBB / CDropout / SVI.py): X normalised, y raw, validate(num_samples=100)."""
    train_x, train_y, test_x, test_y, tau, n_hiddens, n_splits, n_epochs = get_data(dataset, split_id, root)
    if split_id >= n_splits:
        print("Invalid split_id")
        return dict(rmse=np.nan, nll=np.nan, smse=np.nan, model=None)
    print('Dataset %s, split: %d, n_hiddens: %d, prec: %g' % (dataset, split_id, n_hiddens, tau))
    xm, xs, _, _ = normalize(train_x, train_y)
    train_xn, test_xn = (train_x - xm) / xs, (test_x - xm) / xs
    c = dict(num_epochs=100 * n_epochs, batch_size=128, print_every=100)
    if tau == tau:
        c['min_noise'] = float(np.sqrt(1 / tau))
    c.update(conf or {})
    model = model_cls(train_x.shape[1], act if act is not None else nn.Tanh(), num_hiddens=[n_hiddens, n_hiddens], conf=c)
    model.train(train_xn, train_y)
    model.report()
    rmse, nll = model.validate(test_xn, test_y, num_samples=num_samples)
    smse = rmse ** 2 / torch.mean((test_y - train_y.mean()) ** 2)
    print('RMSE = %g, SMSE = %g, NLL = %6.3f' % (float(rmse), float(smse), float(nll)), flush=True)
    return dict(rmse=rmse, nll=nll, smse=smse, model=model)


def mo_metrics(model, nns, test_xn, test_y, y_mean, y_std):
    """experiments/MO_SGDMC.py:48-63 given a trained multi-output model: predictive mean / std of the de-normalised samples per
    output, SMSE and normalised log-likelihood ``ll_ref - ll`` (both per output).  The reference's predictive std is the sample
    std only (no noise term) -- kept."""
    out = model._forward(nns, test_xn, want_pred=False, want_moments=True)
    dev = out["mean"].device
    ys_t = torch.as_tensor(y_std, dtype=torch.float32, device=dev).reshape(1, -1)
    ym_t = torch.as_tensor(y_mean, dtype=torch.float32, device=dev).reshape(1, -1)
    py = (out["mean"] * ys_t + ym_t).contiguous()
    pv = (out["var"] * ys_t * ys_t).contiguous()
    y = model._to_dev(test_y).reshape(py.shape[0], -1)
    rmse, nll = ops.metrics(py, pv, y)                                           # -mean log N(y | py, ps) per output
    mse = rmse ** 2
    test_y_c = torch.as_tensor(test_y, dtype=torch.float32)
    y_mean_c, y_std_c = torch.as_tensor(y_mean, dtype=torch.float32), torch.as_tensor(y_std, dtype=torch.float32)
    smse = mse / torch.mean((test_y_c - y_mean_c) ** 2, dim=0)
    ll_ref = torch.distributions.Normal(y_mean_c, y_std_c).log_prob(test_y_c).mean(dim=0)
    return dict(py=py, ps=pv.sqrt(), rmse=rmse, smse=smse, nll=ll_ref - (-nll))   # MO_SGDMC.py:61-63: nll = ll_ref - ll


def mo_sgdmc(train_x, train_y, test_x, test_y, conf=None, num_hiddens=(50, 50), act=None):
    """experiments/MO_SGDMC.py:14-65 with the data passed in (the reference reads MultiTask/OpAmp text files)."""
    from .BNN_SGDMC import BNN_SGDMC
    x_mean, x_std = train_x.mean(dim=0), train_x.std(dim=0)
    y_mean, y_std = train_y.mean(dim=0), train_y.std(dim=0)
    train_xn, train_yn, test_xn = (train_x - x_mean) / x_std, (train_y - y_mean) / y_std, (test_x - x_mean) / x_std
    c = dict(noise_level=0.01, steps_burnin=2500, steps=2500, keep_every=50, lr_weight=1e-2, lr_noise=3e-1)
    c.update(conf or {})
    model = BNN_SGDMC(dim=train_x.shape[1], nout=train_y.shape[1], act=act if act is not None else nn.Tanh(),
                      num_hiddens=list(num_hiddens), conf=c)
    model.train(train_xn, train_yn)
    model.report()
    res = mo_metrics(model, model.nns, test_xn, test_y, y_mean, y_std)
    res["model"] = model
    print(res["smse"])
    print(res["nll"])
    return res
