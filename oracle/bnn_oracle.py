"""CPU restatement of the reference's predictive hot path (TEST INFRASTRUCTURE).

Every function cites the reference file:line (under /root/reference) whose
arithmetic it restates.  All tensors are torch CPU float32 unless said
otherwise, because that is what the reference computes in; the per-sample loop
is kept as a loop on purpose -- it is also the timed CPU baseline.

A "net" here is a plain list of ``(W, b)`` pairs (``W`` is ``[out, in]``,
``b`` is ``[out]``), one pair per ``nn.Linear`` of the reference's
``nn.Sequential`` (util.py:21-29); the activation sits between pairs and not
after the last one.

Parity status: pinned by tests/golden/*.npz (generated from the unmodified
reference by tests/golden/make_golden.py) for everything except
``psgld_step`` and ``svi_draw`` -- see oracle/__init__.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

F32_EPS = float(torch.finfo(torch.float32).eps)    # 2**-23
F32_TINY = float(torch.finfo(torch.float32).tiny)  # 2**-126


def warmup():
    """Run torch's CPU kernels once before the oracle is used as a checker.

    Observed on the GPU boxes (16-core Xeon, torch 2.11 CPU): the FIRST multi-threaded tanh/addmm call of a fresh
    process occasionally returns values that differ by up to ~7e-5 on one thread's row chunk from every later
    call with the same inputs (the device results agree with the later, stable values to 9e-7).  Tests, smoke() and
    bench call this once so that the checker is deterministic."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4096, 64, generator=g)
    w = torch.randn(128, 64, generator=g)
    b = torch.zeros(128)
    for _ in range(3):
        for f in (torch.tanh, torch.relu, torch.sigmoid):
            f(torch.addmm(b, x, w.t()))
        F.softplus(x).sqrt().log1p()
    return True


# --------------------------------------------------------------------------- #
# a1-a3: the MLP and the per-sample loop
# --------------------------------------------------------------------------- #
def act_fn(name):
    """Activation module shared by all layers (util.py:12, :25)."""
    return {"relu": torch.relu, "tanh": torch.tanh,
            "sigmoid": torch.sigmoid, "identity": (lambda t: t)}[name]


def mlp_forward(net, x, act):
    """util.py:31-33 -> nn.Sequential of Linear/act pairs built at util.py:21-29.

    Linear is ``addmm(b, x, W^T)``; the activation follows every Linear but the
    last.
    """
    f = act_fn(act)
    h = x
    last = len(net) - 1
    for i, (W, b) in enumerate(net):
        h = torch.addmm(b, h, W.t())
        if i != last:
            h = f(h)
    return h


def sample_predict(nets, X, act, nout=None):
    """The hot loop.

    nout given  -> BNN_SGDMC.sample_predict (BNN_SGDMC.py:131-137): pred [S,N,nout]
    nout None   -> BNN_BBB.py:143-149 / BNN_Dropout.py:78-84 /
                   BNN_CDropout.py:158-164 / BNN_SVI.py:70-76: pred [S,N] (squeeze)
    """
    S, N = len(nets), X.shape[0]
    if nout is not None:
        pred = torch.empty(S, N, nout)
        for i in range(S):
            pred[i] = mlp_forward(nets[i], X, act)
    else:
        pred = torch.zeros(S, N)
        for i in range(S):
            pred[i] = mlp_forward(nets[i], X, act).squeeze()
    return pred


def predict_mv(pred, num_test):
    """BNN.py:34-37: view [S,N,-1]; mean over S and *unbiased* var over S."""
    preds = pred.view(pred.shape[0], num_test, -1)
    return preds.mean(dim=0), preds.var(dim=0)


def normal_log_prob(value, loc, scale):
    """torch.distributions.Normal.log_prob as called at BNN.py:31."""
    var = scale ** 2
    return -((value - loc) ** 2) / (2 * var) - scale.log() - math.log(math.sqrt(2 * math.pi))


def validate_metrics(py, pv, y, noise_var=None):
    """BNN.py:25-32 (epistemic variance only) and, with ``noise_var`` (per
    output, already in target units), the SGDMC experiment's variant
    experiments/SGDMC.py:83-93 (pv + ys**2/exp(log_prec))."""
    y = y.view(py.shape[0], -1)
    v = pv if noise_var is None else pv + noise_var
    if bool((v <= 0).any()):
        # torch.distributions.Normal validates scale > 0 (BNN.py:31 raises)
        raise ValueError("predictive variance is not positive")
    rmse = torch.mean((py - y) ** 2, dim=0).sqrt()
    nll = -1 * normal_log_prob(y, py, v.sqrt()).mean(dim=0)
    return rmse, nll


# --------------------------------------------------------------------------- #
# a5-a6: Bayes-by-Backprop draw and KL
# --------------------------------------------------------------------------- #
def stable_noise_var(rho):
    """util.py:52-53: softplus(clamp(rho, min=-35)); softplus threshold 20."""
    return F.softplus(torch.clamp(rho, min=-35.))


def bbb_rsample(mu, rho, eps):
    """BNN_BBB.py:24-27: wb = mu + sqrt(softplus(clamp(rho,-35))) * eps,
    mu/rho/eps all ``[out, in+1]`` with the bias in the last column."""
    return mu + stable_noise_var(rho).sqrt() * eps


def bbb_split(wb):
    """BNN_BBB.py:44-51: weight = wb[:, :-1], bias = wb[:, -1]."""
    return wb[:, :-1].clone(), wb[:, -1].clone()


def bbb_kl(mus, rhos, weight_std):
    """BNN_BBB.py:107-110: sum over layers of KL(N(mu, sigma) || N(0, weight_std))
    with the closed form torch's _kl_normal_normal uses:
    0.5 * (var_ratio + t1 - 1 - log(var_ratio))."""
    kld = torch.zeros(())
    for mu, rho in zip(mus, rhos):
        sigma = stable_noise_var(rho).sqrt()
        var_ratio = (sigma / weight_std) ** 2
        t1 = ((mu - 0.) / weight_std) ** 2
        kld = kld + (0.5 * (var_ratio + t1 - 1 - var_ratio.log())).sum()
    return kld


def svi_draw(mu, sigma_raw, eps):
    """BNN_SVI.py:42-46: theta = mu + softplus(sigma_raw) * eps for every
    parameter tensor (weights and biases).  UNPINNED (pyro absent)."""
    return mu + F.softplus(sigma_raw) * eps


# --------------------------------------------------------------------------- #
# a7-a8: dropout draws, applied as column scaling of the weights
# --------------------------------------------------------------------------- #
def scale_columns(net, masks, skip_unit_inputs):
    """W[:, j] *= m_j per Linear.

    skip_unit_inputs=True  -> BNN_Dropout.py:70-75 (only layers with in > 1)
    skip_unit_inputs=False -> BNN_CDropout.py:59-65 (every layer)
    ``masks`` has one 1-D tensor per *masked* layer, in layer order.
    """
    out, it = [], iter(masks)
    for (W, b) in net:
        if skip_unit_inputs and W.shape[1] <= 1:
            out.append((W.clone(), b.clone()))
            continue
        m = next(it)
        out.append((W * m, b.clone()))   # broadcasts over rows: column j scaled by m[j]
    return out


def bernoulli_keep_mask(u, p):
    """BNN_Dropout.py:74: F.dropout(ones(in), p) * (1 - p) is a pure 0/1 keep
    mask with P(keep) = 1 - p; ``u`` are the uniforms in [0, 1)."""
    return (u < (1.0 - p)).to(torch.float32)


def concrete_keep_mask(u, p_logit, temperature=0.1):
    """BNN_CDropout.py:18-19, :59-65 + torch relaxed_bernoulli.py (LogitRelaxed
    Bernoulli.rsample) + transforms._clipped_sigmoid.

    p = clamp(sigmoid(p_logit)), q = clamp(1 - p), u = clamp(u);
    z = clip(sigmoid((log u - log1p(-u) + log q - log1p(-q)) / T), tiny, 1-eps).
    """
    p = torch.sigmoid(torch.as_tensor(p_logit, dtype=torch.float32)).clamp(F32_EPS, 1 - F32_EPS)
    q = (1 - p).clamp(F32_EPS, 1 - F32_EPS)
    u = u.clamp(F32_EPS, 1 - F32_EPS)
    logits = (u.log() - (-u).log1p() + q.log() - (-q).log1p()) / temperature
    return torch.clamp(torch.sigmoid(logits), min=F32_TINY, max=1.0 - F32_EPS)


# --------------------------------------------------------------------------- #
# a12: pSGLD update (UNPINNED -- pybnn absent; Li et al. 2016 Alg. 1, no Gamma term)
# --------------------------------------------------------------------------- #
def psgld_step(theta, grad, V, lr, noise, alpha=0.99, lam=1e-5):
    """V <- a V + (1-a) g^2 ; G = 1/(lam + sqrt(V)) ;
    theta <- theta - (0.5 lr G g + sqrt(lr G) xi).   SURVEY.md App. C."""
    V = alpha * V + (1.0 - alpha) * grad * grad
    G = 1.0 / (lam + V.sqrt())
    theta = theta - (0.5 * lr * G * grad + (lr * G).sqrt() * noise)
    return theta, V


def lr_factor(t):
    """BNN_SGDMC.py:101: np.float32((1 + it) ** -0.33)."""
    return np.float32((1 + t) ** -0.33)


def sgdmc_neg_log_post(net, log_precs, X, y, act, num_train, alpha_n, beta_n, gain=5. / 3):
    """U = -(loglik * N/|b| + logprior): BNN_SGDMC.py:61-74, :83.

    log p(l) for l = log-precision: Gamma(alpha, beta).log_prob(exp(l)) + l
    (TransformedDistribution with ExpTransform().inv, BNN_SGDMC.py:47);
    weights only get a Normal(0, gain*sqrt(2/(fan_in+fan_out))) prior.
    """
    nout = log_precs.shape[0]
    yv = y.view(-1, nout)
    out = mlp_forward(net, X, act).view(-1, nout)
    precs = log_precs.exp()
    log_lik = (-0.5 * precs * (yv - out) ** 2 + 0.5 * log_precs - 0.5 * np.log(2 * np.pi)).sum()
    a, b = torch.as_tensor(alpha_n, dtype=torch.float32), torch.as_tensor(beta_n, dtype=torch.float32)
    x = log_precs.exp()
    log_gamma = a * b.log() + (a - 1) * x.log() - b * x - torch.lgamma(a)
    log_p = (log_gamma + log_precs).sum()
    for (W, _) in net:
        std = gain * np.sqrt(2. / (W.shape[0] + W.shape[1]))
        log_p = log_p + normal_log_prob(W, torch.zeros(()), torch.as_tensor(std, dtype=torch.float32)).sum()
    return -1 * (log_lik * (num_train / X.shape[0]) + log_p)


# --------------------------------------------------------------------------- #
# multi-GPU merge of per-shard moments (Chan et al.); float64 on purpose
# --------------------------------------------------------------------------- #
def shard_moments(pred64):
    """(count, mean, M2) of one shard of samples; pred64 is [S_shard, M] float64."""
    n = pred64.shape[0]
    mean = pred64.mean(axis=0)
    m2 = ((pred64 - mean) ** 2).sum(axis=0)
    return n, mean, m2


def merge_moments(parts):
    """Pairwise Chan merge of (n, mean, M2) triples, left to right."""
    n, mean, m2 = parts[0]
    for (nb, mb, m2b) in parts[1:]:
        d = mb - mean
        nn = n + nb
        mean = mean + d * (nb / nn)
        m2 = m2 + m2b + d * d * (n * nb / nn)
        n = nn
    return n, mean, m2


# ---- Bayes-by-Backprop training step -------------------------------------------------------------------------
def bbb_local_forward(mus, rhos, x, act, eps):
    """BayesianNN.forward over GaussianLinear layers with local reparameterisation (BNN_BBB.py:29-37, :70-72) and INJECTED noise:
    out = x mu_w^T + mu_b + eps * sqrt(x^2 s2_w^T + s2_b), s2 = stable_noise_var(rho) (util.py:52-53).  mus / rhos: per layer
    [out, in + 1] (bias last, BNN_BBB.py:19-20); eps: per layer [B, out]."""
    f = act_fn(act)
    h = x
    for l, (mu, rho) in enumerate(zip(mus, rhos)):
        w_s2, b_s2 = stable_noise_var(rho[:, :-1]), stable_noise_var(rho[:, -1])
        out_mu = torch.nn.functional.linear(h, mu[:, :-1], mu[:, -1])
        out_std = torch.nn.functional.linear(h ** 2, w_s2, b_s2).sqrt()
        h = out_mu + eps[l] * out_std
        if l + 1 < len(mus):
            h = f(h)
    return h


def bbb_train_steps(mus, rhos, X, y, batches, act, eps, lr=1e-2, kl_factor=1e-3, weight_std=0.1):
    """The minibatch loop body of BNN_BBB.train (BNN_BBB.py:121-126): loss = MSELoss(mean) + kl_factor * KL (:104-110),
    loss.backward(), torch.optim.Adam(lr).step() (:114).  batches: list of index tensors; eps[i][l]: the step's noise.
    Returns (mus, rhos, [mse_i], [kl_i])."""
    mus = [m.clone().requires_grad_(True) for m in mus]
    rhos = [r.clone().requires_grad_(True) for r in rhos]
    # parameter order of the reference's nn.Module: (mu, rho) per layer
    opt = torch.optim.Adam([q for pair in zip(mus, rhos) for q in pair], lr)
    mses, kls = [], []
    for i, idx in enumerate(batches):
        opt.zero_grad()
        pred = bbb_local_forward(mus, rhos, X[idx].reshape(len(idx), -1), act, eps[i])
        mse = torch.nn.MSELoss(reduction="mean")(pred, y[idx].reshape(len(idx), -1))
        kl = bbb_kl(mus, rhos, weight_std)
        (mse + kl_factor * kl).backward()
        opt.step()
        mses.append(float(mse))
        kls.append(float(kl))
    return [m.detach() for m in mus], [r.detach() for r in rhos], mses, kls


# ---- MC-dropout training step --------------------------------------------------------------------------------
def dropout_train_forward(net, x, act, keeps):
    """NN_Dropout.forward in training mode (BNN_Dropout.py:15-21) with INJECTED keep factors: before every Linear whose input is
    wider than one unit, x <- F.dropout(x, p) * (1 - p) = x * keep, keep in {0, ~1}.  net: [(W, b), ...]; keeps: one [B, in] tensor
    per masked layer, in layer order."""
    f = act_fn(act)
    ki = 0
    for l, (W, b) in enumerate(net):
        if x.shape[1] > 1:
            x = x * keeps[ki]
            ki += 1
        x = torch.nn.functional.linear(x, W, b)
        if l + 1 < len(net):
            x = f(x)
    return x


def dropout_train_steps(net, X, y, batches, act, keeps, lr=1e-2, l2_reg=1e-6):
    """The minibatch loop body of BNN_Dropout.train (BNN_Dropout.py:57-62): MSELoss(reduction='sum') of the dropout forward,
    loss.backward(), Adam with weight decay l2_reg on the weights and none on the biases (:43-52).  Returns (net, [loss_i])."""
    Ws = [W.clone().requires_grad_(True) for W, _ in net]
    bs = [b.clone().requires_grad_(True) for _, b in net]
    opt = torch.optim.Adam([{"params": Ws, "weight_decay": l2_reg}, {"params": bs, "weight_decay": 0.}], lr=lr)
    losses = []
    for i, idx in enumerate(batches):
        opt.zero_grad()
        pred = dropout_train_forward(list(zip(Ws, bs)), X[idx], act, keeps[i]).squeeze(-1)
        loss = torch.nn.MSELoss(reduction="sum")(pred, y[idx])
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return [(W.detach(), b.detach()) for W, b in zip(Ws, bs)], losses


# ---- concrete-dropout training step --------------------------------------------------------------------------
def cdropout_rate(p_logit):
    """CDropout.dropout_rate (BNN_CDropout.py:18-19)."""
    return torch.distributions.utils.clamp_probs(torch.sigmoid(p_logit))


def cdropout_train_forward(net, p_logits, x, act, us):
    """NN_CDropout.forward in training mode (BNN_CDropout.py:21-34,45-46,92-93) with INJECTED uniforms: the input of EVERY Linear is
    scaled by 1 - sigmoid((log(p + e) - log(1 - p + e) + log(u + e) - log(1 - u + e)) / 0.1), e = 1e-7.  us: one [B, in] per layer."""
    f = act_fn(act)
    eps, temp = 1e-7, 0.1
    for l, (W, b) in enumerate(net):
        p = cdropout_rate(p_logits[l])
        u = us[l]
        drop = torch.log(p + eps) - torch.log(1 - p + eps) + torch.log(u + eps) - torch.log(1 - u + eps)
        x = x * (1 - torch.sigmoid(drop / temp))
        x = torch.nn.functional.linear(x, W, b)
        if l + 1 < len(net):
            x = f(x)
    return x


def cdropout_reg(net, p_logits, noise_level, lscale=1e-1, dr=2.):
    """BNN_CDropout.reg (BNN_CDropout.py:141-151) over CDropoutLinear.reg (:53-57): (weight_reg, entropy)."""
    noise_var = noise_level ** 2
    wreg, ent = 0., 0.
    for (W, _), pl in zip(net, p_logits):
        p = cdropout_rate(pl)
        entropy = -1 * (p * p.log() + (1 - p) * (1 - p).log())
        wreg = wreg + lscale ** 2 * (1 - p) * noise_var * torch.sum(W ** 2)
        ent = ent + dr * 2 * noise_var * (1. * W.shape[1]) * entropy
    return wreg, ent


def cdropout_train_steps(net, p_logits, X, y, batches, act, us, num_train, noise_level=1., lr=1e-2, lscale=1e-1, dr=2.):
    """The minibatch loop body of BNN_CDropout.train (BNN_CDropout.py:121-133): MSELoss(sum) + (wreg - ent) * rows / num_train,
    backward, Adam on W, b and p_logit (:116).  Returns (net, p_logits, [(mse, wreg, ent)_i])."""
    Ws = [W.clone().requires_grad_(True) for W, _ in net]
    bs = [b.clone().requires_grad_(True) for _, b in net]
    pls = [torch.as_tensor(q, dtype=torch.float32).clone().requires_grad_(True) for q in p_logits]
    opt = torch.optim.Adam(Ws + bs + pls, lr=lr)
    stats = []
    for i, idx in enumerate(batches):
        opt.zero_grad()
        cur = list(zip(Ws, bs))
        pred = cdropout_train_forward(cur, pls, X[idx], act, us[i]).squeeze(-1)
        mse = torch.nn.MSELoss(reduction="sum")(pred, y[idx])
        wreg, ent = cdropout_reg(cur, pls, noise_level, lscale, dr)
        wreg = wreg * (idx.shape[0] / num_train)
        ent = ent * (idx.shape[0] / num_train)
        (mse + (wreg - ent)).backward()
        opt.step()
        stats.append((float(mse.detach()), float(wreg.detach()), float(ent.detach())))
    return [(W.detach(), b.detach()) for W, b in zip(Ws, bs)], [q.detach() for q in pls], stats


def cdropout_train_steps_closed_form(net, p_logits, X, y, batches, act, us, num_train, noise_level=1., lr=1e-2, lscale=1e-1, dr=2.,
                                     betas=(0.9, 0.999), adam_eps=1e-8):
    """The SAME three steps without autograd, in float64: the closed-form gradients the fused kernel uses (csrc/dropout_train.cu) --
    d keep / d z = -(1 - keep) keep / T, so d L / d p_logit_l = sum_{b,k} -g[b,k] xm[b,k] (1 - keep[b,k]) / T * dz/dp * dp/dlogit plus the
    regulariser's -(rows / N) noise^2 (lscale^2 sum(W^2) + 2 dr in log((1 - p) / p)) dp/dlogit; wreg acts on W as the per-layer decay
    2 (rows / N) lscale^2 (1 - p) noise^2 -- and torch.optim.Adam restated.  Checked against the reference-recorded fixture on the
    CPU, this pins the kernel's derivation independently of the GPU.  tanh / relu / sigmoid activations."""
    dt = torch.float64
    net = [(W.to(dt).clone(), b.to(dt).clone()) for W, b in net]
    pl = torch.stack([torch.as_tensor(q, dtype=dt).reshape(()) for q in p_logits]).clone()
    X, y = X.to(dt), y.to(dt)
    nl = len(net)
    nv, eps, T = float(noise_level) ** 2, 1e-7, 0.1
    peps = float(torch.finfo(torch.float32).eps)
    m = [[torch.zeros_like(W), torch.zeros_like(b)] for W, b in net]
    v = [[torch.zeros_like(W), torch.zeros_like(b)] for W, b in net]
    pm, pv = torch.zeros(nl, dtype=dt), torch.zeros(nl, dtype=dt)
    b1, b2 = betas
    f = act_fn(act)

    def dact(a):                                            # derivative through the activation's OUTPUT
        return {"tanh": 1 - a * a, "relu": (a > 0).to(dt), "sigmoid": a * (1 - a)}[act]
    stats = []
    for it, idx in enumerate(batches):
        rows = idx.shape[0]
        frac = rows / num_train
        sg = torch.sigmoid(pl)
        p = sg.clamp(peps, 1 - peps)
        lpr = torch.log(p + eps) - torch.log(1 - p + eps)
        w2 = [(W ** 2).sum() for W, _ in net]
        xs, keeps = [], []
        x = X[idx]
        for l, (W, b) in enumerate(net):
            u = us[it][l].to(dt)
            keep = 1 - torch.sigmoid((lpr[l] + torch.log(u + eps) - torch.log(1 - u + eps)) / T)
            keeps.append(keep)
            xm = x * keep
            xs.append(xm)
            x = xm @ W.T + b
            if l + 1 < nl:
                x = f(x)
        r = x.squeeze(-1) - y[idx]
        wreg = sum(lscale ** 2 * (1 - p[l]) * nv * w2[l] for l in range(nl)) * frac
        ent = sum(dr * 2 * nv * net[l][0].shape[1] * -(p[l] * p[l].log() + (1 - p[l]) * (1 - p[l]).log()) for l in range(nl)) * frac
        stats.append((float((r ** 2).sum()), float(wreg), float(ent)))
        g0 = (2 * r).unsqueeze(-1)
        step = it + 1
        ss, bc2 = lr / (1 - b1 ** step), (1 - b2 ** step) ** 0.5
        gp = torch.zeros(nl, dtype=dt)
        new = [None] * nl
        for l in range(nl - 1, -1, -1):
            W, b = net[l]
            gxm = g0 @ W                                     # gradient with respect to the layer's MASKED input
            dpdl = sg[l] * (1 - sg[l]) if peps <= float(sg[l]) <= 1 - peps else 0.0
            dzdp = 1 / (p[l] + eps) + 1 / (1 - p[l] + eps)
            greg = frac * nv * (-lscale ** 2 * w2[l] - 2 * dr * W.shape[1] * (torch.log(1 - p[l]) - torch.log(p[l])))
            gp[l] = (-(gxm * xs[l] * (1 - keeps[l])).sum() / T * dzdp + greg) * dpdl
            if l > 0:
                a = torch.where(keeps[l] != 0, xs[l] / keeps[l], torch.zeros_like(xs[l]))
                g1 = gxm * keeps[l] * dact(a)
            grads = (g0.T @ xs[l] + 2 * frac * lscale ** 2 * (1 - p[l]) * nv * W, g0.sum(0))
            upd = []
            for q, (par, gr) in enumerate(zip((W, b), grads)):
                m[l][q] = b1 * m[l][q] + (1 - b1) * gr
                v[l][q] = b2 * v[l][q] + (1 - b2) * gr * gr
                upd.append(par - ss * m[l][q] / (v[l][q].sqrt() / bc2 + adam_eps))
            new[l] = tuple(upd)
            if l > 0:
                g0 = g1
        net = new
        pm = b1 * pm + (1 - b1) * gp
        pv = b2 * pv + (1 - b2) * gp * gp
        pl = pl - ss * pm / (pv.sqrt() / bc2 + adam_eps)
    return net, pl, stats


def bbb_train_steps_closed_form(mus, rhos, X, y, batches, act, eps, lr=1e-2, kl_factor=1e-3, weight_std=0.1, betas=(0.9, 0.999),
                                adam_eps=1e-8):
    """bbb_train_steps without autograd, in float64: what the fused kernel computes (csrc/bbb_train.cu).  With a = (h | 1) the
    augmented input, out = a mu^T + eps sqrt(a^2 s2^T): d L / d var = d L / d out * eps / (2 sqrt(var)); d mu = (dL/dout)^T a;
    d s2 = (dL/dvar)^T a^2; d a = dL/dout mu + 2 a * (dL/dvar s2), through the activation below; KL(N(mu, s) || N(0, w0)) adds
    kl_factor mu / w0^2 and kl_factor (1 / w0^2 - 1 / s2) / 2; s2 = softplus(max(rho, -35)) -> d s2 / d rho = sigmoid(rho) above the
    clamp; torch.optim.Adam restated.  Pins the kernel's derivation on the CPU against the reference-recorded fixture."""
    dt = torch.float64
    mus = [m.to(dt).clone() for m in mus]
    rhos = [r.to(dt).clone() for r in rhos]
    X, y = X.to(dt), y.to(dt)
    nl = len(mus)
    iw2 = 1.0 / weight_std ** 2
    st = [[torch.zeros_like(m) for _ in range(4)] for m in mus]       # Adam: m_mu, v_mu, m_rho, v_rho
    b1, b2 = betas
    f = act_fn(act)

    def dact(a):
        return {"tanh": 1 - a * a, "relu": (a > 0).to(dt), "sigmoid": a * (1 - a)}[act]
    mses, kls = [], []
    for it, idx in enumerate(batches):
        rows = idx.shape[0]
        s2 = [F.softplus(torch.clamp(r, min=-35.)) for r in rhos]
        kls.append(float(sum((0.5 * (v * iw2 + m * m * iw2 - 1 - torch.log(v * iw2))).sum() for m, v in zip(mus, s2))))
        h = X[idx].reshape(rows, -1)
        A, SD = [], []
        for l in range(nl):
            a = torch.cat([h, torch.ones(rows, 1, dtype=dt)], dim=1)
            sd = (a * a @ s2[l].T).sqrt()
            A.append(a)
            SD.append(sd)
            h = a @ mus[l].T + eps[it][l].to(dt) * sd
            if l + 1 < nl:
                h = f(h)
        r = h - y[idx].reshape(rows, -1)
        mses.append(float((r * r).mean()))
        g0 = 2 * r / r.numel()
        step = it + 1
        ss, bc2 = lr / (1 - b1 ** step), (1 - b2 ** step) ** 0.5
        for l in range(nl - 1, -1, -1):
            a, sd = A[l], SD[l]
            gv = g0 * eps[it][l].to(dt) / (2 * sd)
            g_mu = g0.T @ a + kl_factor * mus[l] * iw2
            g_s2 = gv.T @ (a * a) + kl_factor * 0.5 * (iw2 - 1 / s2[l])
            g_rho = g_s2 * torch.where(rhos[l] > -35., torch.sigmoid(rhos[l]), torch.zeros_like(rhos[l]))
            if l > 0:
                ga = g0 @ mus[l] + 2 * a * (gv @ s2[l])
                g0 = ga[:, :-1] * dact(a[:, :-1])
            for q, (par, gr) in enumerate(((mus, g_mu), (rhos, g_rho))):
                st[l][2 * q] = b1 * st[l][2 * q] + (1 - b1) * gr
                st[l][2 * q + 1] = b2 * st[l][2 * q + 1] + (1 - b2) * gr * gr
                par[l] = par[l] - ss * st[l][2 * q] / (st[l][2 * q + 1].sqrt() / bc2 + adam_eps)
    return mus, rhos, mses, kls
