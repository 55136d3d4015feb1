"""GPU parity of the tensor-core (tcgen05) forward.  Stated tolerances (DESIGN.md "Precision"):
  tf32x3 (3-pass error-compensated, FP32-grade): 1e-5 relative, same as the FP32 SIMT path;
  tf32_bf16 (TF32 main term + two BF16 correction terms, FP32-grade): 1e-5 relative as well;
  tf32   (single pass, 10-bit mantissa operands) : 5e-3 relative (looser bound, never the default)."""
import ctypes as C

import pytest
import torch

from helpers import golden, flat_to_nets, random_nets, max_rel, var_bound
from oracle import bnn_oracle as O

pytestmark = pytest.mark.gpu


def run(dims, act, nets, X, prec, **kw):
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    arch = Arch(dims, act)
    out = ops.forward(arch, X.cuda().contiguous(), arch.pack(nets).cuda(), want_pred=True, want_moments=True, precision=prec, **kw)
    torch.cuda.synchronize()
    return {k: v.cpu() for k, v in out.items() if not k.startswith("_")}


@pytest.mark.parametrize("K,N,passes,tol", [(8, 16, 1, 2e-3), (16, 64, 3, 2e-6), (64, 208, 3, 2e-6), (64, 256, 1, 2e-3),
                                             (16, 64, 16 + 3, 2e-6), (64, 208, 16 + 3, 2e-6), (24, 112, 16 + 1, 2e-3),
                                             (16, 64, 2, 6e-6), (64, 208, 2, 6e-6), (32, 112, 16 + 2, 6e-6), (64, 208, 16 + 2, 6e-6)])
def test_tcgen05_selftest_gemm(K, N, passes, tol):
    """The kernels' operand layout / descriptors / TMEM read-back on a bare GEMM tile vs float64 (passes + 16: the A
    operand is written to tensor memory with tcgen05.st and read from there by the MMA; passes 2: TF32 main term +
    two BF16 correction terms)."""
    from bnn_b200 import _lib
    g = torch.Generator().manual_seed(K * 1000 + N)
    A, B = torch.randn(128, K, generator=g), torch.randn(N, K, generator=g)
    Ad, Bd, D = A.cuda(), B.cuda(), torch.zeros(128, N, device="cuda")
    _lib.check(_lib.lib().bnn_tc_selftest(Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), K, N, passes, None), "selftest")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    assert float((D.cpu().double() - ref).abs().max() / ref.abs().max()) < tol


@pytest.mark.parametrize("dims,act,S,N", [
    ([16, 200, 200, 1], "relu", 5, 1000),      # cfg 4 shape: CTA pair (cta_group::2), ragged last tile
    ([16, 200, 200, 1], "tanh", 9, 70000),     # several tiles per CTA pair, several samples
    ([9, 100, 100, 1], "tanh", 40, 4573),      # cfg 3 shape: single CTA, sample groups
    ([1, 50, 50, 1], "tanh", 33, 1000),        # cfg 2 shape (D = 1)
    ([13, 50, 64, 3], "sigmoid", 7, 129),      # nout = 3, uneven widths
    ([32, 128, 160, 4], "relu", 2, 300),       # wide input, nout = 4, CTA pair because of shared-memory footprint
    ([64, 96, 64, 2], "relu", 3, 260),         # D = 64
    ([5, 8, 24, 2], "identity", 3, 50),        # tiny
    ([16, 128, 128, 1], "relu", 1, 200),       # S = 1 -> var NaN
])
@pytest.mark.parametrize("prec", ["tf32x3", "tf32_bf16"])
def test_fp32_grade_tc_matches_oracle(dims, act, S, N, prec):
    nets = random_nets(dims, S, seed=sum(dims) + S)
    X = torch.randn(N, dims[0], generator=torch.Generator().manual_seed(N))
    ref = O.sample_predict(nets, X, act, nout=dims[-1])
    out = run(dims, act, nets, X, prec)
    assert max_rel(out["pred"], ref) <= 1e-5
    mean_ref, var_ref = O.predict_mv(ref, N)
    assert max_rel(out["mean"], mean_ref) <= 1e-5
    if S > 1:
        dv = (out["var"].double() - var_ref.double()).abs()
        vb = var_bound(ref, c=256.0, depth=3)      # compensated products carry ~2^-21 relative error each (vs 2^-24 FP32)
        assert bool((dv <= vb).all()), float((dv / vb).max())
    else:
        assert bool(torch.isnan(out["var"]).all())


def test_tf32_single_pass_looser_bound():
    dims, S, N = [16, 200, 200, 1], 4, 600
    nets = random_nets(dims, S, seed=1)
    X = torch.randn(N, 16, generator=torch.Generator().manual_seed(2))
    ref = O.sample_predict(nets, X, "relu", nout=1)
    out = run(dims, "relu", nets, X, "tf32")
    assert 1e-6 < max_rel(out["pred"], ref) <= 5e-3


def test_tc_equals_simt_on_moments_partials():
    """tensor-core and SIMT paths agree with each other (and expose the same shard moments for the multi-GPU merge)."""
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    dims, S, N = [16, 200, 200, 1], 12, 3000
    arch = Arch(dims, "relu")
    bank = arch.pack(random_nets(dims, S, seed=9)).cuda()
    X = torch.randn(N, 16, generator=torch.Generator().manual_seed(4)).cuda()
    a = ops.forward(arch, X, bank, want_partials=True, precision="fp32")
    b = ops.forward(arch, X, bank, want_partials=True, precision="tf32_bf16")
    assert max_rel(b["mean"].cpu(), a["mean"].cpu()) <= 1e-5
    assert max_rel(b["part_mean"].cpu(), a["part_mean"].cpu()) <= 1e-5
    assert max_rel(b["part_m2"].cpu(), a["part_m2"].cpu()) <= 1e-4


def test_unsupported_shapes_are_refused_loudly():
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    for dims in ([13, 50, 1], [90, 512, 512, 512, 1], [16, 300, 300, 1], [64, 256, 256, 4]):
        arch = Arch(dims, "relu")
        bank = arch.pack(random_nets(dims, 2, seed=1)).cuda()
        X = torch.randn(10, dims[0]).cuda()
        with pytest.raises(ValueError, match="tensor-core"):
            ops.forward(arch, X, bank, precision="tf32x3")          # no silent fallback to the SIMT kernel
        assert ops.forward(arch, X, bank, precision="fp32")["mean"].shape == (10, dims[-1])


@pytest.mark.parametrize("prec", ["tf32x3", "tf32_bf16"])
@pytest.mark.parametrize("fixture,kind", [("dropout_protein.npz", "external"), ("cdropout_protein.npz", "external")])
def test_tc_masked_shared_weights_golden(fixture, kind, prec):
    """Dropout family on the tensor-core path: ONE base network + per-sample input-unit masks vs the reference's
    recorded predictions of S column-scaled deep copies."""
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    g = golden(fixture)
    dims = [int(d) for d in g["dims"]]
    arch = Arch(dims, "tanh")
    base = arch.pack(flat_to_nets(g["base"][None], dims))[0].cuda()
    masks = torch.from_numpy(g["masks"]).cuda()
    X = torch.from_numpy(g["X"]).cuda()
    cfg = ops.MaskConfig("external", [True, True, True], mask=masks)
    out = ops.forward(arch, X, base, S=masks.shape[0], mask=cfg, want_pred=True, precision=prec)
    assert max_rel(out["pred"].cpu().squeeze(-1), g["pred"]) <= 2e-5
    assert max_rel(out["mean"].cpu().squeeze(-1), g["mean"].squeeze(-1)) <= 1e-5


@pytest.mark.parametrize("prec", ["tf32x3", "tf32_bf16"])
@pytest.mark.parametrize("kind", ["bernoulli", "concrete"])
def test_tc_inkernel_masks_match_simt(kind, prec):
    """In-kernel Philox masks: tensor-core and SIMT forwards draw the same masks (same counters), so they agree to
    FP32-grade accuracy for any sharding; per-sample bank + masks and several tiles/samples per CTA are exercised."""
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    dims = [9, 100, 100, 1]
    arch = Arch(dims, "tanh")
    base = arch.pack(random_nets(dims, 1, seed=8))[0].cuda()
    X = torch.randn(20000, 9, generator=torch.Generator().manual_seed(3)).cuda()
    S = 23
    cfg = ops.MaskConfig(kind, [True, True, True], p_drop=[0.2, 0.4, 0.1], temperature=0.1, seed=123, sample_offset=50)
    a = ops.forward(arch, X, base, S=S, mask=cfg, want_pred=True, precision="fp32")
    b = ops.forward(arch, X, base, S=S, mask=cfg, want_pred=True, precision=prec)
    assert max_rel(b["pred"].cpu(), a["pred"].cpu()) <= 1e-5
    assert max_rel(b["mean"].cpu(), a["mean"].cpu()) <= 1e-5
    bank = arch.pack(random_nets(dims, S, seed=9)).cuda()          # per-sample weights AND masks
    a = ops.forward(arch, X, bank, mask=cfg, want_pred=True, precision="fp32")
    b = ops.forward(arch, X, bank, mask=cfg, want_pred=True, precision=prec)
    assert max_rel(b["pred"].cpu(), a["pred"].cpu()) <= 1e-5


def test_tc_pred_only_call():
    """A pred-only call (no moments) still needs the operand-image scratch on the tensor-core path."""
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    dims, act = [16, 200, 200, 1], "relu"
    nets, X = random_nets(dims, 3, seed=5), torch.randn(300, 16, generator=torch.Generator().manual_seed(6))
    arch = Arch(dims, act)
    out = ops.forward(arch, X.cuda(), arch.pack(nets).cuda(), want_pred=True, want_moments=False, precision="tf32x3")
    torch.cuda.synchronize()
    ref = O.sample_predict(nets, X, act, nout=1)
    assert max_rel(out["pred"].cpu(), ref) < 1e-5


def test_full_size_cfg4_properties():
    """BASELINE cfg 4 at its full point count (N = 1,000,000; S reduced to 48 so the oracle spot-check stays in seconds):
    size-independent properties of the tensor-core path --
      * the moments of two sample shards, Chan-merged, equal the moments of the unsharded call (the multi-GPU path);
      * permuting the samples does not change mean / var beyond FP32 rounding;
      * a random subset of 1500 points agrees with the CPU oracle to the stated 1e-5."""
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    dims, act, S, N = [16, 200, 200, 1], "relu", 48, 1_000_000
    arch = Arch(dims, act)
    g = torch.Generator().manual_seed(11)
    nets = random_nets(dims, S, seed=12)
    bank = arch.pack(nets).cuda()
    X = torch.randn(N, 16, generator=g)
    Xd = X.cuda()
    full = ops.forward(arch, Xd, bank, want_partials=True, precision="tf32_bf16")
    a = ops.forward(arch, Xd, bank[:20].contiguous(), want_partials=True, precision="tf32_bf16")
    b = ops.forward(arch, Xd, bank[20:].contiguous(), want_partials=True, precision="tf32_bf16")
    pm = torch.stack([a["part_mean"].reshape(-1), b["part_mean"].reshape(-1)])
    p2 = torch.stack([a["part_m2"].reshape(-1), b["part_m2"].reshape(-1)])
    mean_m, var_m = ops.moments_merge(pm, p2, [20, 28])
    torch.cuda.synchronize()
    assert max_rel(mean_m.cpu(), full["mean"].cpu().reshape(-1)) <= 1e-6
    assert max_rel(var_m.cpu(), full["var"].cpu().reshape(-1)) <= 1e-5
    perm = torch.randperm(S, generator=g)
    shuf = ops.forward(arch, Xd, bank[perm.cuda()].contiguous(), precision="tf32_bf16")
    assert max_rel(shuf["mean"].cpu(), full["mean"].cpu()) <= 1e-6
    assert max_rel(shuf["var"].cpu(), full["var"].cpu()) <= 1e-5
    idx = torch.randperm(N, generator=g)[:1500]
    ref = O.sample_predict(nets, X[idx], act, nout=1)
    mean_ref, var_ref = O.predict_mv(ref, len(idx))
    assert max_rel(full["mean"].cpu()[idx], mean_ref) <= 1e-5
    dv = (full["var"].cpu()[idx].double() - var_ref.double()).abs()
    assert bool((dv <= var_bound(ref, c=256.0, depth=3)).all())
    assert bool(torch.isfinite(full["mean"]).all()) and bool(torch.isfinite(full["var"]).all())


@pytest.mark.parametrize("prec", ["tf32_bf16", "fp32"])
def test_forward_streamed_equals_forward(prec):
    """Host-resident inputs, bank uploaded chunk by chunk behind the evaluation: same moments as one device call."""
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    dims, S, N = [16, 200, 200, 1], 23, 3000
    arch = Arch(dims, "relu")
    bank = arch.pack(random_nets(dims, S, seed=21))
    X = torch.randn(N, 16, generator=torch.Generator().manual_seed(22))
    ref = ops.forward(arch, X.cuda(), bank.cuda(), want_partials=True, precision=prec)
    out = ops.forward_streamed(arch, X.pin_memory(), bank.pin_memory(), chunk=5, precision=prec, want_partials=True)
    torch.cuda.synchronize()
    assert max_rel(out["mean"].cpu(), ref["mean"].cpu()) <= 1e-6
    assert max_rel(out["var"].cpu(), ref["var"].cpu()) <= 1e-5
    assert max_rel(out["part_mean"].cpu(), ref["part_mean"].cpu()) <= 1e-9
    assert max_rel(out["part_m2"].cpu(), ref["part_m2"].cpu()) <= 1e-8


def test_random_shape_sweep_tc_vs_simt():
    """Seeded sweep over the supported shape family (D <= 64, H <= 256, O <= 4, any N / S, all activations, with and
    without in-kernel masks): the tensor-core path against the FP32 SIMT path of the same library (which the other tests
    pin to the oracle).  Shapes the tensor-core path refuses are skipped through bnn_forward_tc_supported."""
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    rng = torch.Generator().manual_seed(2024)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=rng))
    done = 0
    for trial in range(40):
        dims = [ri(1, 64), ri(2, 256), ri(2, 256), ri(1, 4)]
        act = ["relu", "tanh", "sigmoid", "identity"][ri(0, 3)]
        S, N = ri(1, 9), ri(1, 700)
        masked = trial % 3 == 0
        arch = Arch(dims, act)
        precs = [pr for pr in ("tf32_bf16", "tf32x3") if ops.tc_supported(arch, S=S, masked=masked, precision=pr)]
        if not precs:
            continue
        bank = ((0.4 / (max(dims) ** 0.5) * 6) * torch.randn(S, arch.stride, generator=rng) * arch.valid_mask()).contiguous().cuda()
        X = torch.randn(N, dims[0], generator=rng).cuda()
        kw = {}
        if masked:
            kw["mask"] = ops.MaskConfig("bernoulli", [dims[0] > 1, True, True], p_drop=[0.1, 0.3, 0.2], seed=trial, sample_offset=7)
        a = ops.forward(arch, X, bank, want_pred=True, precision="fp32", **kw)
        for prec in precs:
            b = ops.forward(arch, X, bank, want_pred=True, precision=prec, **kw)
            assert max_rel(b["pred"].cpu(), a["pred"].cpu()) <= 1e-5, (dims, act, S, N, masked, prec)
            assert max_rel(b["mean"].cpu(), a["mean"].cpu()) <= 1e-5, (dims, act, S, N, masked, prec)
        done += 1
    assert done >= 15


@pytest.mark.parametrize("dims,act,S,N", [
    ([90, 512, 512, 512, 1], "tanh", 5, 1000),       # cfg 5's network: three hidden layers of 512
    ([90, 512, 512, 512, 1], "relu", 3, 300),
    ([7, 300, 1], "relu", 9, 777),                   # one hidden layer, width not a multiple of 32: N-tiles 256 + 48
    ([20, 64, 96, 40, 3], "sigmoid", 4, 520),        # three hidden layers, uneven widths, nout = 3 (sigmoid(0) != 0 in the padding)
    ([1000, 1024, 2], "tanh", 2, 260),               # widest supported: K = 1024 input, four N-tiles
    ([16, 320, 320, 1], "relu", 1, 200),             # H > 256: outside the resident-weights family; S = 1 -> var NaN
    ([33, 40, 40, 40, 40, 40, 2], "identity", 3, 130),   # five hidden layers
])
def test_deep_wide_streaming_kernel_matches_oracle(dims, act, S, N):
    """fwd_tcl_kernel (forward_tcl.cu): any depth, widths up to 1024, FP32-grade TF32 + BF16 products; 1e-5 vs the oracle."""
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    arch = Arch(dims, act)
    assert ops.tc_supported(arch, S=S)                      # ... through the streaming kernel
    nets = random_nets(dims, S, seed=sum(dims) + S)
    X = torch.randn(N, dims[0], generator=torch.Generator().manual_seed(N))
    ref = O.sample_predict(nets, X, act, nout=dims[-1])
    out = run(dims, act, nets, X, "tf32_bf16")
    assert max_rel(out["pred"], ref) <= 1e-5, max_rel(out["pred"], ref)
    mean_ref, var_ref = O.predict_mv(ref, N)
    assert max_rel(out["mean"], mean_ref) <= 1e-5
    if S > 1:
        dv = (out["var"].double() - var_ref.double()).abs()
        # the forward's own rounding grows with the length of the dot products (K up to 1024 here)
        vb = var_bound(ref, c=256.0 * max(1.0, max(dims) / 128.0), depth=len(dims) - 1)
        assert bool((dv <= vb).all()), float((dv / vb).max())
    else:
        assert bool(torch.isnan(out["var"]).all())


def test_deep_wide_moments_only_and_sample_groups():
    """Few points, many samples: sample groups (G > 1) + Chan merge of the groups; moments without materialising pred."""
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    dims, act, S, N = [12, 288, 288, 1], "tanh", 37, 300
    arch = Arch(dims, act)
    nets = random_nets(dims, S, seed=5)
    X = torch.randn(N, 12, generator=torch.Generator().manual_seed(6))
    out = ops.forward(arch, X.cuda(), arch.pack(nets).cuda(), want_pred=False, want_moments=True, want_partials=True, precision="tf32_bf16")
    ref = O.sample_predict(nets, X, act, nout=1)
    mean_ref, var_ref = O.predict_mv(ref, N)
    assert max_rel(out["mean"].cpu(), mean_ref) <= 1e-5
    dv = (out["var"].cpu().double() - var_ref.double()).abs()
    assert bool((dv <= var_bound(ref, c=256.0, depth=3)).all())
    assert max_rel(out["part_mean"].cpu(), ref.double().mean(0).reshape(N, 1)) <= 1e-5
