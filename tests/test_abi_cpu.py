"""CPU: the C-ABI library loads and exports every symbol include/bnn_b200.h declares; the host-side
layout arithmetic (no GPU needed) agrees between the C library and the Python mirror; argument
errors come back as status codes + messages, never aborts."""
import ctypes as C
import os
import re

import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "bnn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bnn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_everything_the_header_declares():
    from bnn_b200 import _lib
    names = declared_functions()
    assert len(names) >= 20
    handle = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/bnn_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names          # the ctypes table mirrors the header one to one
    assert _lib.lib().bnn_abi_version() == 2


@pytest.mark.parametrize("dims", [[13, 50, 1], [1, 50, 50, 1], [9, 100, 100, 1], [16, 200, 200, 1], [90, 512, 512, 512, 1], [7, 3, 5]])
def test_layout_c_vs_python(dims):
    from bnn_b200 import _lib
    from bnn_b200.layout import Arch
    a = Arch(dims, "tanh")
    d = a.desc()
    L = _lib.lib()
    assert L.bnn_param_stride(C.byref(d)) == a.stride and a.stride % 4 == 0
    for l in range(a.n_linear):
        assert L.bnn_layer_offset(C.byref(d), l) == a.offsets[l] and a.offsets[l] % 4 == 0
        assert L.bnn_layer_ld(C.byref(d), l) == a.lds[l]
    lm = (C.c_int32 * 8)(*([1] * a.n_linear + [0] * (8 - a.n_linear)))
    assert L.bnn_mask_len(C.byref(d), lm) == sum(a.ins)
    assert a.n_params == sum((i + 1) * o for i, o in zip(dims[:-1], dims[1:]))


def test_pack_unpack_roundtrip():
    from bnn_b200.layout import Arch
    from bnn_b200.util import NN
    a = Arch([5, 7, 3], "relu")
    nets = [NN(5, nn.ReLU(), [7], nout=3) for _ in range(4)]
    bank = a.pack(nets)
    assert bank.shape == (4, a.stride)
    assert not bool(bank[:, ~a.valid_mask()].any())
    x = torch.randn(11, 5)
    for s in range(4):
        with torch.no_grad():
            assert torch.equal(a.to_module(bank[s])(x), nets[s](x))
    flat = torch.stack([torch.cat([p.detach().reshape(-1) for p in n.parameters()]) for n in nets])
    assert torch.equal(a.pack_flat(flat), bank)


def test_errors_are_status_codes():
    from bnn_b200 import _lib
    from bnn_b200.layout import Arch
    L = _lib.lib()
    d = Arch([4, 4, 1]).desc()
    d.n_linear = 99
    assert L.bnn_param_stride(C.byref(d)) == -1
    assert b"n_linear" in L.bnn_last_error()
    with pytest.raises(ValueError):
        Arch([4, 5000, 1])
    with pytest.raises(NotImplementedError):
        Arch([4, 4, 1], nn.GELU())
    a = _lib.ForwardArgs()
    a.desc = Arch([4, 4, 1]).desc()
    assert L.bnn_forward(C.byref(a), None) == -1            # N = 0, null pointers: rejected before any CUDA call
    assert L.bnn_moments(None, 0, 0, None, None, None, None, None, 0, None) == -1


def test_no_cpu_fallback():
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    a = Arch([3, 4, 1])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.forward(a, torch.zeros(2, 3), torch.zeros(1, a.stride))
    if not torch.cuda.is_available():
        from bnn_b200.BNN_Dropout import BNN_Dropout
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            BNN_Dropout(3, nn.ReLU(), [4]).sample(2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bnn_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_plain_c_consumer(tmp_path):
    """include/bnn_b200.h compiles as C99 and a plain-C program links against the library and gets the documented layout
    numbers and error convention out of the host-only entry points."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from bnn_b200 import _lib
    _lib.lib()
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "c", "abi_smoke.c"),
                    "-o", exe, "-L", libdir, "-l:" + os.path.basename(_lib.LIB_PATH), "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.returncode, r.stdout, r.stderr)


@pytest.mark.parametrize("dims,batch", [([9, 100, 100, 1], 32), ([1, 50, 50, 1], 32), ([5, 24, 16, 1], 16), ([13, 50, 1], 10), ([3, 7, 2], 1)])
def test_training_kernel_plans_host_side(dims, batch):
    """The single-SM training kernels' host-side plans (no GPU): per-step draw counts follow the documented layout -- per layer
    [batch][in] (MC dropout: only layers whose input is wider than one unit; concrete dropout: every layer), each block padded to
    a multiple of four floats; local-reparameterisation noise [batch][out] per layer -- and oversized networks are refused."""
    from bnn_b200 import _lib
    from bnn_b200.layout import Arch
    a = Arch(dims, "tanh")
    d = a.desc()
    L = _lib.lib()
    r4 = lambda n: (n + 3) // 4 * 4
    assert L.bnn_dropout_train_supported(C.byref(d), batch) == 1
    assert L.bnn_cdropout_train_supported(C.byref(d), batch) == 1
    assert L.bnn_dropout_train_mask_per_step(C.byref(d), batch) == sum(r4(batch * i) for i in a.ins if i > 1)
    assert L.bnn_cdropout_train_u_per_step(C.byref(d), batch) == sum(r4(batch * i) for i in a.ins)
    # Bayes by Backprop keeps nine packed vectors in shared memory: 9-100-100-1 (11,704 parameters) does not fit one SM
    assert L.bnn_bbb_train_supported(C.byref(d), batch) == (0 if dims == [9, 100, 100, 1] else 1)
    assert L.bnn_bbb_train_eps_per_step(C.byref(d), batch) == sum(batch * o for o in a.outs)
    big = Arch([90, 512, 512, 1], "tanh").desc()
    assert L.bnn_dropout_train_supported(C.byref(big), 32) == 0
    assert L.bnn_cdropout_train_supported(C.byref(big), 32) == 0
    assert L.bnn_bbb_train_supported(C.byref(big), 32) == 0
    assert L.bnn_dropout_train_supported(C.byref(d), 0) == 0          # bad batch: refused, not a crash
    assert L.bnn_sgld_set_grid_cap(-1) == -1 and b"ctas" in L.bnn_last_error()
    assert L.bnn_sgld_set_grid_cap(0) == 0
