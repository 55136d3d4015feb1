"""Tensor-level wrappers over the C ABI.  torch is plumbing here: it owns device
memory and the current stream; every op below is one call into libbnn_b200.so.
No op has a CPU implementation -- CPU tensors are rejected."""
import ctypes as C

import torch

from . import _lib
from .layout import Arch

_PREC = {"fp32": _lib.PREC_FP32, "tf32x3": _lib.PREC_TF32X3, "tf32": _lib.PREC_TF32, "tf32_bf16": _lib.PREC_TF32_BF16}


def _dev(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: bnn_b200 has no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def _stream(t):
    _lib.check(_lib.lib().bnn_set_device(t.device.index), "bnn_set_device")
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class MaskConfig:
    """Per-sample input-unit scaling (dropout family).

    kind: "external" (tensor [S, mask_len]), "bernoulli" or "concrete" (Philox in-kernel).
    layer_masked: one bool per Linear; p_drop: drop probability per Linear.
    """

    def __init__(self, kind, layer_masked, mask=None, p_drop=None, temperature=0.1, seed=0, sample_offset=0):
        self.kind, self.layer_masked, self.mask = kind, [bool(b) for b in layer_masked], mask
        self.p_drop = list(p_drop) if p_drop is not None else [0.0] * len(layer_masked)
        self.temperature, self.seed, self.sample_offset = float(temperature), int(seed), int(sample_offset)

    def spec(self, arch):
        s = _lib.MaskSpec()
        s.mode = {"external": _lib.MASK_EXTERNAL, "bernoulli": _lib.MASK_BERNOULLI, "concrete": _lib.MASK_CONCRETE}[self.kind]
        for l in range(arch.n_linear):
            s.layer_masked[l] = 1 if self.layer_masked[l] else 0
            s.p_drop[l] = float(self.p_drop[l])
        _, mlen = arch.mask_layout(self.layer_masked)
        if self.kind == "external":
            m = _dev(self.mask, "mask")
            if m.dim() != 2 or m.shape[1] < mlen:
                raise ValueError(f"mask must be [S, >= {mlen}], got {tuple(m.shape)}")
            s.mask = m.data_ptr()
            s.mask_stride = m.shape[1]
        s.temperature = self.temperature
        s.seed = self.seed & 0xFFFFFFFFFFFFFFFF
        s.sample_offset = self.sample_offset
        return s


_AUTO_TC_MIN_FLOP = 5e7   # below this much work per call, precision="auto" stays on the single-launch FP32 kernel


def forward(arch: Arch, X, W, S=None, mask: MaskConfig = None, want_pred=False, want_moments=True,
            want_partials=False, precision="fp32"):
    """Fused S x N batched MLP forward (+ predictive moments).

    X: [N, dim]; W: bank [S, stride] or one shared network [stride] (then pass S).
    Returns dict with any of pred [S,N,nout], mean/var [N,nout], part_mean/part_m2 (float64 [N,nout]).
    """
    X = _dev(X, "X")
    W = _dev(W, "W")
    if X.dim() != 2 or X.shape[1] != arch.dims[0]:
        raise ValueError(f"X must be [N, {arch.dims[0]}], got {tuple(X.shape)}")
    a = _lib.ForwardArgs()
    a.desc = arch.desc()
    a.X, a.N = X.data_ptr(), X.shape[0]
    if W.dim() == 1:
        if S is None:
            raise ValueError("S is required with a single shared network")
        if W.shape[0] < arch.stride:
            raise ValueError("W shorter than the packed network")
        a.w_stride, a.S = 0, int(S)
    else:
        if W.shape[1] < arch.stride:
            raise ValueError(f"W must be [S, >= {arch.stride}], got {tuple(W.shape)}")
        a.w_stride, a.S = W.shape[1], W.shape[0]
        if S is not None and int(S) != W.shape[0]:
            raise ValueError("S does not match the bank")
    a.W = W.data_ptr()
    if mask is not None:
        a.mask = mask.spec(arch)
        if mask.kind == "external" and mask.mask.shape[0] != a.S:
            raise ValueError("mask rows != S")
    if precision == "auto":     # tensor cores (FP32-grade TF32 + BF16 corrections) when the shape is supported, else the FP32 SIMT kernel
        a.precision = _lib.PREC_TF32_BF16
        # tiny problems (cfg 1: 100 samples x 51 points of a 13-50-1 net) are launch-bound: the SIMT kernel is ONE launch + the
        # finalisation, the tensor-core paths need their operand-image passes first
        tiny = float(a.S) * float(X.shape[0]) * arch.flops_per_eval < _AUTO_TC_MIN_FLOP
        if tiny or not _lib.lib().bnn_forward_tc_supported(C.byref(a)):
            a.precision = _lib.PREC_FP32
    else:
        a.precision = _PREC[precision]
    N, O, Sn = X.shape[0], arch.nout, int(a.S)
    if N < 1 or Sn < 1:
        raise ValueError("need at least one point and one sample")
    out = {}
    dev = X.device
    if want_pred:
        out["pred"] = torch.empty(Sn, N, O, dtype=torch.float32, device=dev)
        a.pred = out["pred"].data_ptr()
    if want_moments:
        out["mean"] = torch.empty(N, O, dtype=torch.float32, device=dev)
        out["var"] = torch.empty(N, O, dtype=torch.float32, device=dev)
        a.mean, a.var = out["mean"].data_ptr(), out["var"].data_ptr()
    if want_partials:
        out["part_mean"] = torch.empty(N, O, dtype=torch.float64, device=dev)
        out["part_m2"] = torch.empty(N, O, dtype=torch.float64, device=dev)
        a.part_mean, a.part_m2 = out["part_mean"].data_ptr(), out["part_m2"].data_ptr()
    L = _lib.lib()
    # scratch: Welford partials and, on the tensor-core path, the operand images (needed even for pred-only calls)
    nbytes = L.bnn_forward_workspace_bytes(C.byref(a))
    if nbytes < 0:
        _lib.check(-1, "bnn_forward_workspace_bytes")
    if nbytes > 0:
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        out["_ws"] = ws   # keep alive until the caller drops the dict (stream-ordered use)
    _lib.check(L.bnn_forward(C.byref(a), _stream(X)), "bnn_forward")
    return out


def forward_streamed(arch: Arch, X_host, W_host, chunk=250, precision="auto", want_partials=False, device=None):
    """predict_mv for inputs that live in (pinned) HOST memory: the bank is uploaded in chunks of `chunk` samples on a copy
    stream while the previous chunk is being evaluated, each chunk is reduced to float64 per-point (mean, M2), and the
    chunks are Chan-merged -- the same arithmetic as the multi-GPU shard merge, so the result equals one `forward` call
    over the whole bank to FP64 rounding.  Returns {"mean", "var"} (float32) and, with want_partials, the merged float64
    {"part_mean", "part_m2"} of the whole bank (for `distributed.merge_shards`)."""
    if X_host.is_cuda or W_host.is_cuda:
        raise TypeError("forward_streamed takes host tensors (use forward for device tensors)")
    if W_host.dim() != 2 or W_host.shape[1] < arch.stride:
        raise ValueError(f"W must be [S, >= {arch.stride}]")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    S, stride = W_host.shape
    chunk = max(1, min(int(chunk), S))
    comp = torch.cuda.current_stream(dev)
    copy = torch.cuda.Stream(dev)
    Xd = X_host.to(dev, non_blocking=True)
    bufs = [torch.empty(chunk, stride, dtype=torch.float32, device=dev) for _ in range(2 if S > chunk else 1)]
    freed = [None, None]
    n_tot, mean, m2 = 0, None, None
    keep = []
    for i, lo in enumerate(range(0, S, chunk)):
        n = min(chunk, S - lo)
        buf = bufs[i % len(bufs)]
        with torch.cuda.stream(copy):
            if freed[i % 2] is not None:
                copy.wait_event(freed[i % 2])              # the evaluation that last read this buffer is done
            elif i == 0:
                copy.wait_stream(comp)
            buf[:n].copy_(W_host[lo:lo + n], non_blocking=True)
            up = torch.cuda.Event()
            up.record(copy)
        comp.wait_event(up)
        out = forward(arch, Xd, buf[:n], want_moments=False, want_partials=True, precision=precision)
        ev = torch.cuda.Event()
        ev.record(comp)
        freed[i % 2] = ev
        keep.append(out)
        pm, p2 = out["part_mean"], out["part_m2"]
        if mean is None:
            n_tot, mean, m2 = n, pm, p2
        else:                                               # Chan et al. pairwise merge, float64 on the device
            d = pm - mean
            nn = n_tot + n
            mean = mean + d * (n / nn)
            m2 = m2 + p2 + d * d * (n_tot * n / nn)
            n_tot = nn
    res = {}
    res["mean"], res["var"] = moments_merge(mean.unsqueeze(0), m2.unsqueeze(0), [n_tot])
    if want_partials:
        res["part_mean"], res["part_m2"] = mean, m2
    for b in bufs:
        b.record_stream(comp)
    return res


def tc_supported(arch: Arch, S=2, masked=False, precision="tf32_bf16"):
    """True if the tensor-core forward takes this architecture in this precision mode (see bnn_forward_tc_supported; the
    modes pad K differently, so a shape at the shared-memory limit may fit one and not the other)."""
    a = _lib.ForwardArgs()
    a.desc = arch.desc()
    a.N, a.S, a.w_stride = 1, S, arch.stride
    a.mask.mode = _lib.MASK_EXTERNAL if masked else _lib.MASK_NONE
    a.precision = _PREC[precision]
    return bool(_lib.lib().bnn_forward_tc_supported(C.byref(a)))


def forward_host(arch: Arch, X, W, S=None, mask: MaskConfig = None, want_pred=False, want_moments=True, want_partials=False,
                 precision="fp32", out=None):
    """End-to-end variant through the C ABI's host-buffer entry point (bnn_forward_host): X / W / mask / outputs are HOST
    tensors (pin them); the library does the H2D copies (large banks in chunks on a copy stream behind the evaluation of the
    previous chunk), the launches, the D2H copies and synchronises.  `out` (optional dict of pinned tensors from a previous
    call) is reused instead of allocating new pinned outputs."""
    for name, t in (("X", X), ("W", W)):
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous float32 host tensor")
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: bnn_b200 has no CPU fallback")
    a = _lib.ForwardArgs()
    a.desc = arch.desc()
    a.X, a.N = X.data_ptr(), X.shape[0]
    if W.dim() == 1:
        a.w_stride, a.S = 0, int(S)
    else:
        a.w_stride, a.S = W.shape[1], W.shape[0]
    a.W = W.data_ptr()
    if mask is not None:
        if mask.kind == "external":
            spec = MaskConfig("bernoulli", mask.layer_masked).spec(arch)   # fill the common fields
            spec.mode = _lib.MASK_EXTERNAL
            spec.mask = mask.mask.data_ptr()
            spec.mask_stride = mask.mask.shape[1]
            a.mask = spec
        else:
            a.mask = mask.spec(arch)
    if precision == "auto":
        a.N = X.shape[0]
        a.precision = _lib.PREC_TF32_BF16
        if not _lib.lib().bnn_forward_tc_supported(C.byref(a)):
            a.precision = _lib.PREC_FP32
    else:
        a.precision = _PREC[precision]
    N, O, Sn = X.shape[0], arch.nout, int(a.S)
    out = dict(out) if out is not None else {}

    def pinned(name, shape, dtype):
        t = out.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype).pin_memory()
            out[name] = t
        return t

    if want_pred:
        a.pred = pinned("pred", (Sn, N, O), torch.float32).data_ptr()
    if want_moments:
        a.mean, a.var = pinned("mean", (N, O), torch.float32).data_ptr(), pinned("var", (N, O), torch.float32).data_ptr()
    if want_partials:
        a.part_mean = pinned("part_mean", (N, O), torch.float64).data_ptr()
        a.part_m2 = pinned("part_m2", (N, O), torch.float64).data_ptr()
    L = _lib.lib()
    _lib.check(L.bnn_set_device(torch.cuda.current_device()), "bnn_set_device")
    _lib.check(L.bnn_forward_host(C.byref(a)), "bnn_forward_host")
    return out


def moments(pred, want_partials=False):
    """mean and unbiased var over dim 0 of pred [S, ...] (BNN.py:37)."""
    pred = _dev(pred, "pred")
    S = pred.shape[0]
    tail = pred.shape[1:]
    M = pred[0].numel()
    L = _lib.lib()
    mean = torch.empty(tail, dtype=torch.float32, device=pred.device)
    var = torch.empty(tail, dtype=torch.float32, device=pred.device)
    pm = pm2 = None
    if want_partials:
        pm = torch.empty(tail, dtype=torch.float64, device=pred.device)
        pm2 = torch.empty(tail, dtype=torch.float64, device=pred.device)
    ws = torch.empty(max(int(L.bnn_moments_workspace_bytes(S, M)), 16), dtype=torch.uint8, device=pred.device)
    _lib.check(L.bnn_moments(_ptr(pred), S, M, _ptr(mean), _ptr(var), _ptr(pm), _ptr(pm2), _ptr(ws), ws.numel(),
                             _stream(pred)), "bnn_moments")
    return (mean, var, pm, pm2) if want_partials else (mean, var)


def moments_merge(part_mean, part_m2, counts):
    """Chan merge of R shards: part_mean/part_m2 float64 [R, ...]; counts: R ints."""
    if part_mean.dtype != torch.float64 or part_m2.dtype != torch.float64 or not part_mean.is_cuda:
        raise TypeError("shard moments must be float64 CUDA tensors")
    R = part_mean.shape[0]
    tail = part_mean.shape[1:]
    M = part_mean[0].numel()
    mean = torch.empty(tail, dtype=torch.float32, device=part_mean.device)
    var = torch.empty(tail, dtype=torch.float32, device=part_mean.device)
    cnt = (C.c_int64 * R)(*[int(c) for c in counts])
    _lib.check(_lib.lib().bnn_moments_merge(_ptr(part_mean.contiguous()), _ptr(part_m2.contiguous()), cnt, R, M,
                                            _ptr(mean), _ptr(var), _stream(mean)), "bnn_moments_merge")
    return mean, var


def metrics(mean, var, y, noise_var=None):
    """rmse[nout], nll[nout] of BNN.validate (BNN.py:30-31); noise_var (per output) gives the
    experiments/SGDMC.py:83-93 variant.  Raises ValueError where the reference's Normal() would."""
    mean, var = _dev(mean, "mean"), _dev(var, "var")
    N = mean.shape[0]
    nout = mean[0].numel()
    y = _dev(y.reshape(N, -1).contiguous(), "y")
    if y.shape[1] != nout:
        raise ValueError(f"y has {y.shape[1]} outputs, predictions have {nout}")
    nv = None
    if noise_var is not None:
        nv = _dev(torch.as_tensor(noise_var, dtype=torch.float32, device=mean.device).reshape(-1).contiguous(), "noise_var")
        if nv.numel() != nout:
            raise ValueError("noise_var needs one value per output")
    out = torch.empty(2 * nout + 1, dtype=torch.float32, device=mean.device)
    acc = torch.empty(2 * nout + 1, dtype=torch.float64, device=mean.device)     # the kernel's FP64 accumulators
    _lib.check(_lib.lib().bnn_metrics(_ptr(mean), _ptr(var), _ptr(y), _ptr(nv), N, nout, _ptr(out), _ptr(acc),
                                      _stream(mean)), "bnn_metrics")
    host = out.cpu()
    if host[2 * nout] > 0:
        raise ValueError("predictive variance is not positive (the reference's Normal(scale) raises here too)")
    return host[:nout].clone(), host[nout: 2 * nout].clone()


def reparam_kl(arch: Arch, mu, rho, S, mode="bbb", eps=None, seed=0, sample_offset=0, weight_std=None):
    """W[s] = mu + sigma(rho) * eps_s for s < S (BNN_BBB.py:24-27 / BNN_SVI.py:42-46) and, when
    weight_std is given, KL(q || N(0, weight_std)) (BNN_BBB.py:107-110) as a float64 device scalar."""
    mu, rho = _dev(mu, "mu"), _dev(rho, "rho")
    if mu.numel() != arch.stride or rho.numel() != arch.stride:
        raise ValueError("mu/rho must be packed vectors")
    W = torch.empty(S, arch.stride, dtype=torch.float32, device=mu.device) if S > 0 else None
    if eps is not None:
        eps = _dev(eps, "eps")
        if tuple(eps.shape) != (S, arch.stride):
            raise ValueError("eps must be [S, stride]")
    kl = torch.zeros((), dtype=torch.float64, device=mu.device) if weight_std is not None else None
    d = arch.desc()
    _lib.check(_lib.lib().bnn_reparam_kl(C.byref(d), _lib.SCALE_BBB if mode == "bbb" else _lib.SCALE_SVI, _ptr(mu), _ptr(rho),
                                         _ptr(eps), seed & 0xFFFFFFFFFFFFFFFF, sample_offset, S, _ptr(W),
                                         float(weight_std) if weight_std is not None else 0.0, _ptr(kl), _stream(mu)),
               "bnn_reparam_kl")
    return W, kl


def dropout_masks(arch: Arch, cfg: MaskConfig, S, device, uniforms=None):
    """[S, mask_len] keep masks (BNN_Dropout.py:70-75 / BNN_CDropout.py:59-65)."""
    _, mlen = arch.mask_layout(cfg.layer_masked)
    out = torch.empty(S, mlen, dtype=torch.float32, device=device)
    if uniforms is not None:
        uniforms = _dev(uniforms, "uniforms")
    d = arch.desc()
    spec = cfg.spec(arch)
    _lib.check(_lib.lib().bnn_dropout_masks(C.byref(d), C.byref(spec), _ptr(uniforms), S, _ptr(out), _stream(out)),
               "bnn_dropout_masks")
    return out


def apply_masks(arch: Arch, layer_masked, base, masks):
    """Bank [S, stride] with W[s][:, j] = base[:, j] * mask[s][j] on masked layers (the reference's order)."""
    base, masks = _dev(base, "base"), _dev(masks, "masks")
    S = masks.shape[0]
    W = torch.empty(S, arch.stride, dtype=torch.float32, device=base.device)
    lm = (C.c_int32 * _lib.MAX_LINEAR)(*([1 if b else 0 for b in layer_masked] + [0] * (_lib.MAX_LINEAR - len(layer_masked))))
    d = arch.desc()
    _lib.check(_lib.lib().bnn_apply_masks(C.byref(d), lm, _ptr(base), _ptr(masks), masks.shape[1], S, _ptr(W), _stream(W)),
               "bnn_apply_masks")
    return W


def psgld_step(theta, grad, V, lr_head, lr_tail=None, n_head=None, alpha=0.99, lam=1e-5, noise=None, seed=0, step=0):
    """In-place preconditioned SGLD update of theta and V (see include/bnn_b200.h)."""
    theta, grad, V = _dev(theta, "theta"), _dev(grad, "grad"), _dev(V, "V")
    n = theta.numel()
    if noise is not None:
        noise = _dev(noise, "noise")
    _lib.check(_lib.lib().bnn_psgld_step(_ptr(theta), _ptr(grad), _ptr(V), n, n if n_head is None else n_head, lr_head,
                                         lr_head if lr_tail is None else lr_tail, alpha, lam, _ptr(noise),
                                         seed & 0xFFFFFFFFFFFFFFFF, step, _stream(theta)), "bnn_psgld_step")


def sgld_perm(seed, epoch, N, count, device, offset=0):
    """Entries [offset, offset + count) of the epoch's shuffled order (bnn_sgld_perm): int64 device tensor."""
    out = torch.empty(int(count), dtype=torch.int64, device=device)
    _lib.check(_lib.lib().bnn_sgld_perm(seed & 0xFFFFFFFFFFFFFFFF, int(epoch) & 0xFFFFFFFF, int(N), int(offset), int(count),
                                        _ptr(out), _stream(out)), "bnn_sgld_perm")
    return out


def sgld_set_grid_cap(ctas):
    """Cap the CTAs of every following fused-step launch of this process (0 = no cap): chains that run on different streams then
    share the GPU (bnn_sgld_set_grid_cap).  One chain alone leaves about a third of the SMs idle (dependency chain of the step)."""
    _lib.check(_lib.lib().bnn_sgld_set_grid_cap(int(ctas)), "bnn_sgld_set_grid_cap")


class SgldChain:
    """Device-resident state of one pSGLD chain + the fused step kernel (bnn_sgld_run, csrc/sgld_step.cu).

    theta = [packed network | log_precs | pad] (double-buffered), V (preconditioner, ones at start), t (steps taken: LR
    schedule + Philox subsequence).  Mirrors what BNN_SGDMC.train sets up at BNN_SGDMC.py:96-101."""

    def __init__(self, arch: Arch, batch, device, lr_weight, lr_noise, alpha_n, beta_n, seed=0, precond_decay=0.99,
                 diagonal_bias=1e-5, gain=5. / 3, precond_mode=0):
        L = _lib.lib()
        self.arch, self.batch, self.device = arch, int(batch), torch.device(device)
        d = arch.desc()
        self.n_theta = int(L.bnn_sgld_theta_len(C.byref(d)))
        self.n_thetaT = int(L.bnn_sgld_theta_t_len(C.byref(d)))
        ws = int(L.bnn_sgld_workspace_bytes(C.byref(d), self.batch))
        if min(self.n_theta, self.n_thetaT, ws) < 0:
            _lib.check(-1, "bnn_sgld_workspace_bytes")
        self.theta = torch.zeros(2, self.n_theta, device=self.device)
        self.thetaT = torch.zeros(2, self.n_thetaT, device=self.device)
        self.V = torch.ones(self.n_theta, device=self.device)
        self.ws = torch.zeros(ws, dtype=torch.uint8, device=self.device)
        self.loss = torch.zeros(1, device=self.device)
        self.cur, self.t = 0, 0
        c = _lib.SgldChain()
        c.desc, c.batch = d, self.batch
        c.theta, c.thetaT, c.V = self.theta.data_ptr(), self.thetaT.data_ptr(), self.V.data_ptr()
        c.workspace, c.workspace_bytes = self.ws.data_ptr(), ws
        c.lr_weight, c.lr_noise = float(lr_weight), float(lr_noise)
        c.precond_decay, c.diagonal_bias = float(precond_decay), float(diagonal_bias)
        c.alpha_n, c.beta_n, c.gain = float(alpha_n), float(beta_n), float(gain)
        c.precond_mode, c.seed = int(precond_mode), int(seed) & 0xFFFFFFFFFFFFFFFF
        self._c = c
        self._keep = None

    def set_state(self, packed_net, log_precs):
        """Start (or restart) the chain at this network / these log-precisions: V = 1, t = 0 (a fresh optimizer and scheduler,
        as every BNN_SGDMC.train call builds, BNN_SGDMC.py:100-101)."""
        a = self.arch
        self.cur, self.t = 0, 0
        self.theta.zero_()
        self.theta[0, : a.stride] = packed_net.to(self.device, torch.float32) * a.valid_mask().to(self.device)
        self.theta[0, a.stride: a.stride + a.nout] = log_precs.to(self.device, torch.float32)
        self.V.fill_(1.)
        _lib.check(_lib.lib().bnn_sgld_init(C.byref(self._c), self.cur, _stream(self.theta)), "bnn_sgld_init")

    def bind(self, X, y, num_train=None):
        """Training set of the chain.  num_train (default: rows of X) is the N of the N/|b| likelihood scale and the
        length of an epoch; a caller may bind a buffer holding only the rows its permutations will ever index."""
        X, y = _dev(X, "X"), _dev(y, "y")
        if X.dim() != 2 or X.shape[1] != self.arch.dims[0] or y.numel() != X.shape[0] * self.arch.nout:
            raise ValueError("X must be [N, dim] and y [N, nout]")
        self._keep = (X, y)
        self._c.X, self._c.y = X.data_ptr(), y.data_ptr()
        self._c.num_train = int(num_train) if num_train is not None else X.shape[0]

    def _perm(self, perm):
        if perm.dtype != torch.int64 or not perm.is_cuda or not perm.is_contiguous():
            raise TypeError("perm must be a contiguous int64 CUDA tensor")
        self._c.perm = perm.data_ptr()

    def run(self, perm, perm_off, n_steps, want_loss=True):
        """n_steps fused SGLD steps on the batches perm[perm_off + i * batch ...]; returns the device scalar holding the
        loss of the last step (valid after the launch completes)."""
        self._perm(perm)
        if perm_off + (n_steps - 1) * self.batch >= min(perm.numel(), self._c.num_train):
            raise ValueError("the launch would run past the permutation")
        _lib.check(_lib.lib().bnn_sgld_run(C.byref(self._c), self.cur, self.t, int(perm_off), int(n_steps),
                                           _ptr(self.loss) if want_loss else None, None, _stream(self.theta)), "bnn_sgld_run")
        self.t += n_steps
        self.cur ^= n_steps & 1
        return self.loss

    def grad(self, perm, perm_off=0):
        """(dU/dtheta [n_theta], loss U) at the current state for the batch at perm_off -- no update (tests)."""
        self._perm(perm)
        g = torch.zeros(self.n_theta, device=self.device)
        _lib.check(_lib.lib().bnn_sgld_run(C.byref(self._c), self.cur, self.t, int(perm_off), 1, _ptr(self.loss), _ptr(g),
                                           _stream(self.theta)), "bnn_sgld_run")
        return g, self.loss.clone()

    def check(self):
        st = C.c_int32(0)
        _lib.check(_lib.lib().bnn_sgld_status(C.byref(self._c), C.byref(st), _stream(self.theta)), "bnn_sgld_status")
        if st.value != 0:
            raise RuntimeError("fused SGLD step: a dependency wait timed out; the chain state is undefined")

    @property
    def state(self):
        return self.theta[self.cur]


def philox_words(seed, subsequence, tag, n, device):
    out = torch.empty(n, dtype=torch.int32, device=device)
    _lib.check(_lib.lib().bnn_philox_words(seed, subsequence, tag, n, _ptr(out), _stream(out)), "bnn_philox_words")
    return out


def philox_normals(seed, subsequence, tag, n, device):
    out = torch.empty(n, dtype=torch.float32, device=device)
    _lib.check(_lib.lib().bnn_philox_normals(seed, subsequence, tag, n, _ptr(out), _stream(out)), "bnn_philox_normals")
    return out


# ---- fused Bayes-by-Backprop training step -------------------------------------------------------------------
def bbb_train_supported(arch: Arch, batch):
    """True if the single-SM fused training kernel takes this network at this batch size (bnn_bbb_train_supported)."""
    d = arch.desc()
    return bool(_lib.lib().bnn_bbb_train_supported(C.byref(d), int(batch)))


class BbbTrainer:
    """Device-resident variational parameters (packed bias-last mu, rho), their Adam moments and the fused training kernel
    (bnn_bbb_train_run, csrc/bbb_train.cu): what BNN_BBB.train keeps in nn.Parameters + torch.optim.Adam (BNN_BBB.py:112-126)."""

    def __init__(self, arch: Arch, batch, device, lr=1e-2, kl_factor=1e-3, weight_std=0.1, betas=(0.9, 0.999), eps=1e-8, seed=0):
        if not bbb_train_supported(arch, batch):
            _lib.check(-1, "bnn_bbb_train_supported")
        self.arch, self.batch, self.device = arch, int(batch), torch.device(device)
        self.mu = torch.zeros(arch.stride, device=self.device)
        self.rho = torch.zeros(arch.stride, device=self.device)
        self.adam = torch.zeros(4, arch.stride, device=self.device)
        self.t = 0
        d = arch.desc()
        self.eps_per_step = int(_lib.lib().bnn_bbb_train_eps_per_step(C.byref(d), self.batch))
        c = _lib.BbbTrain()
        c.desc, c.batch = d, self.batch
        c.mu, c.rho, c.adam = self.mu.data_ptr(), self.rho.data_ptr(), self.adam.data_ptr()
        c.lr, c.beta1, c.beta2, c.eps = float(lr), float(betas[0]), float(betas[1]), float(eps)
        c.kl_factor, c.weight_std, c.seed = float(kl_factor), float(weight_std), int(seed) & 0xFFFFFFFFFFFFFFFF
        self._c = c
        self._keep = None

    def set_state(self, mu_packed, rho_packed):
        """Start at these variational parameters with a fresh optimizer (zero moments, step 0)."""
        self.mu.copy_(mu_packed.to(self.device, torch.float32))
        self.rho.copy_(rho_packed.to(self.device, torch.float32))
        self.adam.zero_()
        self.t = 0

    def bind(self, X, y):
        X, y = _dev(X, "X"), _dev(y, "y")
        if X.dim() != 2 or X.shape[1] != self.arch.dims[0] or y.numel() != X.shape[0] * self.arch.nout:
            raise ValueError("X must be [N, dim] and y [N, nout]")
        self._keep = (X, y)
        self._c.X, self._c.y, self._c.num_train = X.data_ptr(), y.data_ptr(), X.shape[0]

    def run(self, perm, perm_off, n_steps, want_stats=False, eps=None):
        """n_steps fused training steps on the batches perm[perm_off + i * batch ...].  Returns (mse[n_steps], kl[n_steps]) device
        tensors when want_stats, else None.  eps (tests): injected noise [n_steps, eps_per_step]."""
        if perm.dtype != torch.int64 or not perm.is_cuda or not perm.is_contiguous():
            raise TypeError("perm must be a contiguous int64 CUDA tensor")
        self._c.perm = perm.data_ptr()
        mse = kl = None
        if want_stats:
            mse = torch.empty(n_steps, device=self.device)
            kl = torch.empty(n_steps, device=self.device)
        if eps is not None:
            eps = _dev(eps, "eps").to(torch.float32).contiguous()
            if eps.numel() != n_steps * self.eps_per_step:
                raise ValueError(f"eps must hold {n_steps} x {self.eps_per_step} values")
        _lib.check(_lib.lib().bnn_bbb_train_run(C.byref(self._c), int(self.t), int(perm_off), int(n_steps),
                                                 mse.data_ptr() if want_stats else None, kl.data_ptr() if want_stats else None,
                                                 eps.data_ptr() if eps is not None else None, _stream(self.mu)), "bnn_bbb_train_run")
        self.t += int(n_steps)
        self._last = (perm, eps)      # stream-ordered use: keep alive until the next call
        return (mse, kl) if want_stats else None


# ---- fused MC-dropout training step --------------------------------------------------------------------------
def dropout_train_supported(arch: Arch, batch):
    """True if the single-SM fused training kernel takes this network at this batch size (bnn_dropout_train_supported)."""
    d = arch.desc()
    return bool(_lib.lib().bnn_dropout_train_supported(C.byref(d), int(batch)))


class DropoutTrainer:
    """The device-resident packed network, its Adam moments and the fused training kernel (bnn_dropout_train_run,
    csrc/dropout_train.cu): what BNN_Dropout.train keeps in nn.Parameters + torch.optim.Adam (BNN_Dropout.py:39-65)."""

    def __init__(self, arch: Arch, batch, device, lr=1e-2, l2_reg=1e-6, p_drop=0.05, betas=(0.9, 0.999), eps=1e-8, seed=0):
        if not dropout_train_supported(arch, batch):
            _lib.check(-1, "bnn_dropout_train_supported")
        self.arch, self.batch, self.device = arch, int(batch), torch.device(device)
        self.W = torch.zeros(arch.stride, device=self.device)
        self.adam = torch.zeros(2, arch.stride, device=self.device)
        self.t = 0
        d = arch.desc()
        self.mask_per_step = int(_lib.lib().bnn_dropout_train_mask_per_step(C.byref(d), self.batch))
        c = _lib.DropoutTrain()
        c.desc, c.batch = d, self.batch
        c.W, c.adam = self.W.data_ptr(), self.adam.data_ptr()
        c.lr, c.beta1, c.beta2, c.eps = float(lr), float(betas[0]), float(betas[1]), float(eps)
        c.l2_reg, c.p_drop, c.seed = float(l2_reg), float(p_drop), int(seed) & 0xFFFFFFFFFFFFFFFF
        self._c = c
        self._keep = None

    def set_state(self, packed):
        """Start at this packed network with a fresh optimizer (zero moments, step 0)."""
        self.W.copy_(packed.reshape(-1).to(self.device, torch.float32))
        self.adam.zero_()
        self.t = 0

    def bind(self, X, y):
        X, y = _dev(X, "X"), _dev(y, "y")
        if X.dim() != 2 or X.shape[1] != self.arch.dims[0] or y.numel() != X.shape[0] * self.arch.nout:
            raise ValueError("X must be [N, dim] and y [N, nout]")
        self._keep = (X, y)
        self._c.X, self._c.y, self._c.num_train = X.data_ptr(), y.data_ptr(), X.shape[0]

    def run(self, perm, perm_off, n_steps, want_loss=False, masks=None):
        """n_steps fused training steps on the batches perm[perm_off + i * batch ...].  Returns the per-step sums of squared
        errors (device tensor [n_steps]) when want_loss.  masks (tests): injected keep factors [n_steps, mask_per_step]."""
        if perm.dtype != torch.int64 or not perm.is_cuda or not perm.is_contiguous():
            raise TypeError("perm must be a contiguous int64 CUDA tensor")
        self._c.perm = perm.data_ptr()
        loss = torch.empty(n_steps, device=self.device) if want_loss else None
        if masks is not None:
            masks = _dev(masks, "masks").to(torch.float32).contiguous()
            if masks.numel() != n_steps * self.mask_per_step:
                raise ValueError(f"masks must hold {n_steps} x {self.mask_per_step} values")
        _lib.check(_lib.lib().bnn_dropout_train_run(C.byref(self._c), int(self.t), int(perm_off), int(n_steps),
                                                     loss.data_ptr() if want_loss else None,
                                                     masks.data_ptr() if masks is not None else None, _stream(self.W)),
                   "bnn_dropout_train_run")
        self.t += int(n_steps)
        self._last = (perm, masks)    # stream-ordered use: keep alive until the next call
        return loss


# ---- fused concrete-dropout training step --------------------------------------------------------------------
def cdropout_train_supported(arch: Arch, batch):
    """True if the single-SM fused training kernel takes this network at this batch size (bnn_cdropout_train_supported)."""
    d = arch.desc()
    return bool(_lib.lib().bnn_cdropout_train_supported(C.byref(d), int(batch)))


class CDropoutTrainer:
    """The device-resident packed network, the per-layer p_logits, their Adam moments, the noise level and the fused training
    kernel (bnn_cdropout_train_run, csrc/dropout_train.cu): the state of BNN_CDropout.train (BNN_CDropout.py:114-139)."""

    def __init__(self, arch: Arch, batch, device, lr=1e-2, lscale=1e-1, dr=2., min_noise=0., noise_level=1., betas=(0.9, 0.999),
                 eps=1e-8, seed=0):
        if not cdropout_train_supported(arch, batch):
            _lib.check(-1, "bnn_cdropout_train_supported")
        self.arch, self.batch, self.device = arch, int(batch), torch.device(device)
        self.W = torch.zeros(arch.stride, device=self.device)
        self.adam = torch.zeros(2, arch.stride, device=self.device)
        self.p_logit = torch.zeros(arch.n_linear, device=self.device)
        self.p_adam = torch.zeros(2, arch.n_linear, device=self.device)
        self.noise_level = torch.full((1,), float(noise_level), device=self.device)
        self.t = 0
        d = arch.desc()
        self.u_per_step = int(_lib.lib().bnn_cdropout_train_u_per_step(C.byref(d), self.batch))
        c = _lib.CDropoutTrain()
        c.desc, c.batch = d, self.batch
        c.W, c.adam, c.p_logit, c.p_adam = self.W.data_ptr(), self.adam.data_ptr(), self.p_logit.data_ptr(), self.p_adam.data_ptr()
        c.noise_level = self.noise_level.data_ptr()
        c.lr, c.beta1, c.beta2, c.eps = float(lr), float(betas[0]), float(betas[1]), float(eps)
        c.lscale, c.dr, c.min_noise, c.seed = float(lscale), float(dr), float(min_noise), int(seed) & 0xFFFFFFFFFFFFFFFF
        self._c = c
        self._keep = None

    def set_state(self, packed, p_logits, noise_level=None):
        """Start at this packed network / these p_logits with a fresh optimizer (zero moments, step 0)."""
        self.W.copy_(packed.reshape(-1).to(self.device, torch.float32))
        self.p_logit.copy_(torch.as_tensor(p_logits, dtype=torch.float32).reshape(-1))
        self.adam.zero_()
        self.p_adam.zero_()
        if noise_level is not None:
            self.noise_level.fill_(float(noise_level))
        self.t = 0

    def bind(self, X, y):
        X, y = _dev(X, "X"), _dev(y, "y")
        if X.dim() != 2 or X.shape[1] != self.arch.dims[0] or y.numel() != X.shape[0] * self.arch.nout:
            raise ValueError("X must be [N, dim] and y [N, nout]")
        self._keep = (X, y)
        self._c.X, self._c.y, self._c.num_train = X.data_ptr(), y.data_ptr(), X.shape[0]

    def run(self, perm, perm_off, n_steps, epoch_end=False, want_stats=False, uniforms=None):
        """n_steps fused training steps on the batches perm[perm_off + i * batch ...]; epoch_end: the launch updates noise_level
        from its squared errors (BNN_CDropout.py:134).  Returns per-step (squared errors, wreg, ent) [n_steps, 3] when want_stats.
        uniforms (tests): injected torch.rand_like draws [n_steps, u_per_step]."""
        if perm.dtype != torch.int64 or not perm.is_cuda or not perm.is_contiguous():
            raise TypeError("perm must be a contiguous int64 CUDA tensor")
        self._c.perm = perm.data_ptr()
        stats = torch.empty(n_steps, 3, device=self.device) if want_stats else None
        if uniforms is not None:
            uniforms = _dev(uniforms, "uniforms").to(torch.float32).contiguous()
            if uniforms.numel() != n_steps * self.u_per_step:
                raise ValueError(f"uniforms must hold {n_steps} x {self.u_per_step} values")
        _lib.check(_lib.lib().bnn_cdropout_train_run(C.byref(self._c), int(self.t), int(perm_off), int(n_steps), int(bool(epoch_end)),
                                                      stats.data_ptr() if want_stats else None,
                                                      uniforms.data_ptr() if uniforms is not None else None, _stream(self.W)),
                   "bnn_cdropout_train_run")
        self.t += int(n_steps)
        self._last = (perm, uniforms)  # stream-ordered use: keep alive until the next call
        return stats
