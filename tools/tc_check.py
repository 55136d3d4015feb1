"""Staged checks of the tensor-core path; each stage runs in its own process under a timeout so that a hang
in one stage cannot take the others (or the box) down:  python tools/tc_check.py [stage ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def selftest(K, N, passes):
    import ctypes as C, torch
    from bnn_b200 import _lib
    g = torch.Generator().manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=g); B = torch.randn(N, K, generator=g)
    Ad, Bd = A.cuda(), B.cuda()
    D = torch.zeros(128, N, device="cuda")
    _lib.check(_lib.lib().bnn_tc_selftest(Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), K, N, passes, None), "selftest")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    err = float((D.cpu().double() - ref).abs().max() / ref.abs().max())
    print(f"selftest K={K} N={N} passes={passes}: max-normalised err {err:.3e}")
    return err


def fused(dims, act, S, N, prec):
    import torch
    from helpers import random_nets, max_rel
    from oracle import bnn_oracle as O
    O.warmup()
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    nets = random_nets(dims, S, seed=5)
    X = torch.randn(N, dims[0], generator=torch.Generator().manual_seed(1))
    arch = Arch(dims, act)
    out = ops.forward(arch, X.cuda(), arch.pack(nets).cuda(), want_pred=True, want_moments=True, precision=prec)
    torch.cuda.synchronize()
    ref = O.sample_predict(nets, X, act, nout=dims[-1])
    mean_ref, var_ref = O.predict_mv(ref, N)
    e = (max_rel(out["pred"].cpu(), ref), max_rel(out["mean"].cpu(), mean_ref), max_rel(out["var"].cpu(), var_ref) if S > 1 else 0.0)
    print(f"fused {dims} {act} S={S} N={N} {prec}: pred {e[0]:.3e} mean {e[1]:.3e} var {e[2]:.3e}")
    return e


STAGES = {
    "self1": lambda: selftest(16, 64, 1),
    "self3": lambda: selftest(16, 64, 3),
    "self3b": lambda: selftest(64, 208, 3),
    "f1": lambda: fused([16, 100, 100, 1], "relu", 3, 300, "tf32x3"),
    "f1b": lambda: fused([9, 100, 100, 1], "tanh", 40, 4573, "tf32x3"),
    "f1c": lambda: fused([13, 50, 64, 3], "tanh", 7, 129, "tf32"),
    "p1": lambda: fused([16, 200, 200, 1], "relu", 3, 700, "tf32x3"),
    "p2": lambda: fused([16, 200, 200, 1], "tanh", 9, 70000, "tf32x3"),
    "p3": lambda: fused([16, 200, 200, 1], "relu", 2, 300, "tf32"),
    "m1": lambda: fused([16, 100, 100, 1], "relu", 3, 300, "tf32_bf16"),
    "m1b": lambda: fused([9, 100, 100, 1], "tanh", 40, 4573, "tf32_bf16"),
    "m1c": lambda: fused([13, 50, 64, 3], "tanh", 7, 129, "tf32_bf16"),
    "mp1": lambda: fused([16, 200, 200, 1], "relu", 3, 700, "tf32_bf16"),
    "mp2": lambda: fused([16, 200, 200, 1], "tanh", 9, 70000, "tf32_bf16"),
    "mp3": lambda: fused([16, 200, 200, 1], "relu", 64, 20000, "tf32_bf16"),
}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        STAGES[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(STAGES)
    for nm in names:
        try:
            r = subprocess.run([sys.executable, __file__, "--run", nm], timeout=90, capture_output=True, text=True)
            tail = (r.stdout.strip().splitlines() or [""])[-1]
            err = (r.stderr.strip().splitlines() or [""])[-1] if r.returncode else ""
            print(f"[{nm}] rc={r.returncode} {tail} {err}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"[{nm}] TIMEOUT (hang)", flush=True)
            break      # a hung kernel may have wedged the context; stop here
