"""Shared test helpers: fixtures loading, random nets, oracle adapters."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def flat_to_nets(flat, dims):
    """[S, n_params] in the reference's parameter order -> list of [(W, b), ...]."""
    flat = torch.as_tensor(flat, dtype=torch.float32)
    nets = []
    for s in range(flat.shape[0]):
        o, net = 0, []
        for i, j in zip(dims[:-1], dims[1:]):
            W = flat[s, o:o + i * j].reshape(j, i).clone(); o += i * j
            b = flat[s, o:o + j].clone(); o += j
            net.append((W, b))
        nets.append(net)
    return nets


def random_nets(dims, S, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    nets = []
    for _ in range(S):
        net = []
        for i, j in zip(dims[:-1], dims[1:]):
            bound = scale * (6.0 / (i + j)) ** 0.5
            W = (torch.rand(j, i, generator=g) * 2 - 1) * bound
            b = (torch.rand(j, generator=g) * 2 - 1) / max(i, 1) ** 0.5
            net.append((W, b))
        nets.append(net)
    return nets


def max_rel(a, b):
    """max |a-b| normalised by max |b| (the survey's 'max-normalised' error)."""
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def var_bound(pred_ref, rel=1e-5, c=32.0, depth=2):
    """Per-point bound on |dvar|:  rel*var + c*eps_f32*max_s|pred|*sqrt(var)  (+ the same in var).

    The forward's own rounding (~eps*|pred| per sample, more for deep nets) enters the variance as
    ~2*eps*|pred|*sigma, which dominates rel*var when sigma << |pred|."""
    eps = float(np.finfo(np.float32).eps)
    p = torch.as_tensor(pred_ref, dtype=torch.float64)
    var = p.var(dim=0) if p.shape[0] > 1 else torch.zeros_like(p[0])
    scale = p.abs().amax(dim=0)
    c = c * max(1.0, depth / 2.0)      # forward rounding accumulates layer by layer
    return rel * var + c * eps * (scale * var.sqrt() + var) + 1e-30


def energy_inputs(dims, nout, B, seed):
    """Deterministic inputs of the SGDMC energy fixture, from numpy's frozen legacy stream (RandomState), so that the
    tests can rebuild them bit for bit on any box without storing megabytes of weights:
    per Linear W[out,in] ~ 0.5 * xavier-uniform(gain 5/3), b ~ U(-0.1, 0.1); log_precs ~ N(0.3, 0.2); X ~ N(0,1); y ~ N(0,1)."""
    rs = np.random.RandomState(seed)
    net = []
    for i, o in zip(dims[:-1], dims[1:]):
        bound = 0.5 * (5. / 3) * np.sqrt(6. / (i + o))
        W = rs.uniform(-bound, bound, size=(o, i)).astype(np.float32)
        b = rs.uniform(-0.1, 0.1, size=(o,)).astype(np.float32)
        net.append((torch.from_numpy(W), torch.from_numpy(b)))
    log_precs = torch.from_numpy((0.3 + 0.2 * rs.standard_normal(nout)).astype(np.float32))
    X = torch.from_numpy(rs.standard_normal((B, dims[0])).astype(np.float32))
    y = torch.from_numpy(rs.standard_normal((B, nout)).astype(np.float32))
    return net, log_precs, X, y


def energy_cases():
    """The reference-recorded SGDMC energy / gradient fixture (tests/golden/sgdmc_energy.npz) as a list of dicts with the
    inputs rebuilt from their seed."""
    g = golden("sgdmc_energy.npz")
    cases = []
    for name in [str(n) for n in g["names"]]:
        c = {k[len(name) + 1:]: g[k] for k in g if k.startswith(name + ".")}
        c["name"] = name
        c["dims"] = [int(d) for d in c["dims"]]
        c["act"] = str(c["act"])
        c["B"], c["num_train"], c["seed"] = int(c["B"]), int(c["num_train"]), int(c["seed"])
        c["grad_stride"] = int(g["grad_stride"])
        c["net"], c["log_precs"], c["X"], c["y"] = energy_inputs(c["dims"], c["dims"][-1], c["B"], c["seed"])
        cases.append(c)
    return cases


def check_energy_grads(case, grads, grad_log_precs, rtol=2e-5):
    """grads: [(gW, gb), ...] per Linear (any float tensors); compares with the reference's autograd gradient recorded in
    the fixture (every grad_stride-th element of the big tensors, plus sum and L2 norm of each)."""
    stride = case["grad_stride"]
    for li, pair in enumerate(grads):
        for tag, t in zip(("W", "b"), pair):
            v = torch.as_tensor(t).detach().cpu().double().reshape(-1)
            sub = torch.from_numpy(case[f"g{tag}{li}.sub"]).double()
            got = v[::stride] if v.numel() > 4096 else v
            scale = float(sub.abs().max()) + 1e-30
            assert float((got - sub).abs().max()) <= rtol * scale, (case["name"], tag, li, float((got - sub).abs().max()) / scale)
            norm = float(case[f"g{tag}{li}.norm"])
            assert abs(float(v.norm()) - norm) <= rtol * (norm + 1e-30), (case["name"], tag, li, "norm")
    glp = torch.as_tensor(grad_log_precs).detach().cpu().double().reshape(-1)
    ref = torch.from_numpy(case["grad_log_precs"]).double().reshape(-1)
    assert float((glp - ref).abs().max()) <= rtol * float(ref.abs().max() + 1e-30), (case["name"], "log_precs")


def bbb_train_case():
    """The reference-generated fixture of three BNN_BBB training steps (tests/golden/bbb_train.npz; make_golden.gen_bbb_train)."""
    g = golden("bbb_train.npz")
    dims = [int(v) for v in g["dims"]]
    nl = len(dims) - 1
    t = lambda a: torch.as_tensor(a)
    case = {"dims": dims, "act": str(g["act"]), "weight_std": float(g["weight_std"]), "lr": float(g["lr"]),
            "kl_factor": float(g["kl_factor"]), "batch": int(g["batch"]), "X": t(g["X"]), "y": t(g["y"]),
            "perm": t(g["perm"]).long(), "mse": g["mse"], "kl": g["kl"],
            "mu0": [t(g[f"mu0_{l}"]) for l in range(nl)], "rho0": [t(g[f"rho0_{l}"]) for l in range(nl)],
            "mu1": [t(g[f"mu1_{l}"]) for l in range(nl)], "rho1": [t(g[f"rho1_{l}"]) for l in range(nl)]}
    n_steps = len(case["mse"])
    case["eps"] = [[t(g[f"eps{i}_{l}"]) for l in range(nl)] for i in range(n_steps)]
    B, N = case["batch"], case["X"].shape[0]
    case["batches"] = [case["perm"][i * B: min(N, (i + 1) * B)] for i in range(n_steps)]
    return case


def dropout_train_case():
    """The reference-generated fixture of three BNN_Dropout training steps (tests/golden/dropout_train.npz;
    make_golden.gen_dropout_train)."""
    g = golden("dropout_train.npz")
    dims = [int(v) for v in g["dims"]]
    nl = len(dims) - 1
    t = lambda a: torch.as_tensor(a)
    case = {"dims": dims, "act": str(g["act"]), "p_drop": float(g["p_drop"]), "l2_reg": float(g["l2_reg"]), "lr": float(g["lr"]),
            "batch": int(g["batch"]), "X": t(g["X"]), "y": t(g["y"]), "perm": t(g["perm"]).long(), "loss": g["loss"],
            "net0": [(t(g[f"W0_{l}"]), t(g[f"b0_{l}"])) for l in range(nl)],
            "net1": [(t(g[f"W1_{l}"]), t(g[f"b1_{l}"])) for l in range(nl)]}
    n_steps = len(case["loss"])
    case["keeps"] = [[t(g[f"keep{i}_{l}"]) for l in range(nl)] for i in range(n_steps)]
    B, N = case["batch"], case["X"].shape[0]
    case["batches"] = [case["perm"][i * B: min(N, (i + 1) * B)] for i in range(n_steps)]
    return case


def cdropout_train_case():
    """The reference-generated fixture of three BNN_CDropout training steps (tests/golden/cdropout_train.npz;
    make_golden.gen_cdropout_train)."""
    g = golden("cdropout_train.npz")
    dims = [int(v) for v in g["dims"]]
    nl = len(dims) - 1
    t = lambda a: torch.as_tensor(a)
    case = {"dims": dims, "act": str(g["act"]), "lr": float(g["lr"]), "lscale": float(g["lscale"]), "dr": float(g["dr"]),
            "noise_level": float(g["noise_level"]), "batch": int(g["batch"]), "X": t(g["X"]), "y": t(g["y"]),
            "perm": t(g["perm"]).long(), "stats": g["stats"], "p_logit0": t(g["p_logit0"]), "p_logit1": t(g["p_logit1"]),
            "net0": [(t(g[f"W0_{l}"]), t(g[f"b0_{l}"])) for l in range(nl)],
            "net1": [(t(g[f"W1_{l}"]), t(g[f"b1_{l}"])) for l in range(nl)]}
    n_steps = case["stats"].shape[0]
    case["us"] = [[t(g[f"u{i}_{l}"]) for l in range(nl)] for i in range(n_steps)]
    B, N = case["batch"], case["X"].shape[0]
    case["batches"] = [case["perm"][i * B: min(N, (i + 1) * B)] for i in range(n_steps)]
    return case
