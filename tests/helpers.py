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
