"""CPU: host-side glue of the driver layer: the UCI loader against a synthetic dataset directory (and
against the reference's own get_data when /root/reference is present, i.e. in the authoring container only), and the batched
NSGA-II restatement on analytic problems."""
import os
import sys

import numpy as np
import pytest
import torch


def make_uci_dir(root, name="toy", n=60, d=4, n_splits=2, seed=0):
    rs = np.random.RandomState(seed)
    base = os.path.join(root, name)
    os.makedirs(os.path.join(base, "data"))
    os.makedirs(os.path.join(base, "results"))
    X = rs.standard_normal((n, d))
    y = X @ rs.standard_normal(d) + 0.1 * rs.standard_normal(n)
    np.savetxt(os.path.join(base, "data", "data.txt"), np.column_stack([X, y]))
    for k in range(n_splits):
        perm = rs.permutation(n)
        np.savetxt(os.path.join(base, "data", f"index_train_{k}.txt"), perm[: int(0.9 * n)], fmt="%d")
        np.savetxt(os.path.join(base, "data", f"index_test_{k}.txt"), perm[int(0.9 * n):], fmt="%d")
    np.savetxt(os.path.join(base, "data", "n_splits.txt"), [n_splits], fmt="%d")
    np.savetxt(os.path.join(base, "data", "n_hidden.txt"), [16], fmt="%d")
    np.savetxt(os.path.join(base, "data", "n_epochs.txt"), [4], fmt="%d")
    np.savetxt(os.path.join(base, "results", "test_tau_100_xepochs_1_hidden_layers.txt"), [0.5 + k for k in range(n_splits)])
    return X, y


def test_get_data_layout(tmp_path):
    from bnn_b200.experiments import get_data
    X, y = make_uci_dir(str(tmp_path))
    trx, try_, tex, tey, tau, nh, ns, ne = get_data("toy", 1, str(tmp_path))
    tr = np.loadtxt(tmp_path / "toy" / "data" / "index_train_1.txt", dtype=np.int64)
    te = np.loadtxt(tmp_path / "toy" / "data" / "index_test_1.txt", dtype=np.int64)
    assert trx.dtype == torch.float32 and trx.shape == (54, 4) and tex.shape == (6, 4)
    assert np.allclose(trx.numpy(), X[tr].astype(np.float32)) and np.allclose(tey.numpy(), y[te].astype(np.float32))
    assert tau == 1.5 and nh == 16 and ns == 2 and ne == 4
    out = get_data("toy", 2, str(tmp_path))                     # split_id >= n_splits: Nones, like the reference
    assert out[0] is None and out[6] == 2


@pytest.mark.skipif(not os.path.isdir("/root/reference/UCI_Datasets"), reason="reference tree only exists in the authoring container")
def test_get_data_matches_reference_loader():
    from bnn_b200.experiments import get_data
    root = "/root/reference/UCI_Datasets"
    ours = get_data("bostonHousing", 0, root)
    # the reference's loader, executed from its own experiments/ directory (it uses relative paths)
    import subprocess
    code = ("import sys, types; sys.path.insert(0, '..');\n"
            "import numpy as np, torch\n"
            "src = open('SGDMC.py').read().split('def uci')[0].split('def get_data')[1]\n"
            "ns = {'np': np, 'torch': torch}\n"
            "exec('def get_data' + src, ns)\n"
            "r = ns['get_data']('bostonHousing', 0)\n"
            "print(r[0].shape, float(r[0].sum()), float(r[3].sum()), float(r[4]), int(r[5]), int(r[6]), int(r[7]))\n")
    r = subprocess.run([sys.executable, "-c", code], cwd="/root/reference/experiments", capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    want = r.stdout.strip().splitlines()[-1]
    got = f"{ours[0].shape} {float(ours[0].sum())} {float(ours[3].sum())} {float(ours[4])} {int(ours[5])} {int(ours[6])} {int(ours[7])}"
    assert got == want


def test_nsga2_single_objective_and_constraint():
    from bnn_b200.BO import nsga2
    rng = np.random.default_rng(1)
    X, Y = nsga2(lambda P: ((P - 0.3) ** 2).sum(axis=1, keepdims=True), 3, -1.7, 1.7, nobj=1, ncons=0, population=40,
                 max_evals=4000, rng=rng)
    assert Y.min() < 1e-3 and np.abs(X[np.argmin(Y[:, 0])] - 0.3).max() < 0.05
    # minimise x0 subject to 0.5 - x0 <= 0: optimum at x0 = 0.5
    f = lambda P: np.column_stack([P[:, 0], 0.5 - P[:, 0]])
    X, Y = nsga2(f, 1, -1.7, 1.7, nobj=1, ncons=1, population=40, max_evals=3000, rng=rng)
    assert np.all(Y[:, 1] <= 1e-9) and abs(Y[:, 0].min() - 0.5) < 1e-2


def test_nsga2_two_objectives_front():
    from bnn_b200.BO import nsga2
    f = lambda P: np.column_stack([P[:, 0] ** 2, (P[:, 0] - 1.0) ** 2])       # Schaffer: Pareto set x in [0, 1]
    X, Y = nsga2(f, 1, -1.7, 1.7, nobj=2, population=40, max_evals=3000, rng=np.random.default_rng(2))
    assert len(X) >= 10 and X.min() > -0.05 and X.max() < 1.05
