"""Drop-in for the reference's BNN_SGDMC (BNN_SGDMC.py:19-140): a pSGLD chain over an MLP with a
learned log-precision per output.

What runs where:
  * sample / sample_predict / predict_mv / validate: the fused CUDA forward over the [S, stride]
    bank of thinned chain states (no deepcopy'd nn.Modules);
  * train: the chain state theta = [packed network | log_precs] lives in ONE device vector; each
    SGLD step is a hand-derived minibatch gradient of U (BNN_SGDMC.py:61-83: a few GEMMs on the packed
    blocks; a fused step kernel is the row SURVEY.md 8f ranks next) followed by the fused bnn_psgld_step kernel (update + RMSprop
    preconditioner + in-kernel Philox noise + both LR groups, BNN_SGDMC.py:86-87,96-101); thinning
    snapshots are row copies into the bank (BNN_SGDMC.py:124).
pybnn is not needed.  The update rule is restated from Li et al. 2016 (UNPINNED).
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .BNN import BNN
from .bank import SampleBank
from .layout import Arch
from .util import NN


class BNN_SGDMC(nn.Module, BNN):
    def __init__(self, dim, act=nn.ReLU(), num_hiddens=[50], nout=1, conf=dict()):
        nn.Module.__init__(self)
        BNN.__init__(self)
        self.dim, self.act, self.num_hiddens, self.nout = dim, act, num_hiddens, nout
        self.steps_burnin = conf.get('steps_burnin', 2500)
        self.steps = conf.get('steps', 2500)
        self.keep_every = conf.get('keep_every', 50)
        self.batch_size = conf.get('batch_size', 32)
        self.warm_start = conf.get('warm_start', False)
        self.lr_weight = conf.get('lr_weight', 2e-2)
        self.lr_noise = conf.get('lr_noise', 1e-1)
        self.alpha_n = torch.as_tensor(1. * conf.get('alpha_n', 1e-2))
        self.beta_n = torch.as_tensor(1. * conf.get('beta_n', 1e-1))
        self.noise_level = conf.get('noise_level', None)
        if self.noise_level is not None:                       # BNN_SGDMC.py:39-45
            prec = 1 / self.noise_level ** 2
            prec_var = (prec * 0.25) ** 2
            self.beta_n = torch.as_tensor(prec / prec_var)
            self.alpha_n = torch.as_tensor(prec * self.beta_n)
            print("Reset alpha_n = %g, beta_n = %g" % (self.alpha_n, self.beta_n))
        # ours
        self.seed = int(conf.get('seed', 0))
        self.precision = conf.get('precision', 'auto')   # tensor cores (3xTF32, FP32-grade) when the shape allows, else FP32 SIMT
        self.psgld_alpha = conf.get('precondition_decay_rate', 0.99)
        self.psgld_lambda = conf.get('diagonal_bias', 1e-5)
        self.device = torch.device(conf['device']) if 'device' in conf else None
        self.use_graph = bool(conf.get('cuda_graph', True))

        self.arch = Arch([dim] + list(num_hiddens) + [nout], act)
        self.log_precs = nn.Parameter(torch.zeros(self.nout))
        self.nn = NN(dim, self.act, self.num_hiddens, self.nout)
        self.gain = 5. / 3
        self.nns = []
        self.init_nn()

    def init_nn(self):
        self.log_precs.data = (self.alpha_n / self.beta_n).log() * torch.ones(self.nout)
        for l in self.nn.nn:
            if type(l) == nn.Linear:
                nn.init.xavier_uniform_(l.weight, gain=self.gain)

    # ---- the reference's two energy terms on the module state (BNN_SGDMC.py:61-74), for callers that use them directly;
    #      the chain itself evaluates them fused on the packed vector (_neg_log_post / _grad_into)
    def log_prior(self):
        al, be = float(self.alpha_n), float(self.beta_n)
        lp = self.log_precs
        # Gamma(al, be).log_prob(exp(l)) + l: TransformedDistribution with ExpTransform().inv (:47, :62)
        log_p = (al * math.log(be) + (al - 1.) * lp - be * lp.exp() - math.lgamma(al) + lp).sum()
        for n, p in self.nn.nn.named_parameters():
            if "weight" in n:
                std = float(self.gain * np.sqrt(2. / (p.shape[0] + p.shape[1])))                 # :65
                log_p = log_p - (p * p).sum() / (2. * std * std) - p.numel() * (math.log(std) + 0.5 * math.log(2. * math.pi))
        return log_p

    def log_lik(self, X, y):
        y = y.view(-1, self.nout)
        out = self.nn(X).view(-1, self.nout)
        precs = self.log_precs.exp()
        return (-0.5 * precs * (y - out) ** 2 + 0.5 * self.log_precs - 0.5 * math.log(2. * math.pi)).sum()   # :72-74

    # ---- energy, evaluated on a packed parameter vector (BNN_SGDMC.py:61-74,83) ----------
    def _neg_log_post(self, theta, bx, by, num_train):
        """U = -(loglik * N/|b| + logprior) on a packed parameter vector.  Written with Python-float constants only
        (no host tensors are created), so that the whole forward + backward can be captured in a CUDA graph."""
        a = self.arch
        h = bx
        log_p = 0.
        for l in range(a.n_linear):
            blk = theta[a.offsets[l]: a.offsets[l] + a.outs[l] * a.lds[l]].view(a.outs[l], a.lds[l])
            W, b = blk[:, : a.ins[l]], blk[:, a.ins[l]]
            std = float(self.gain * np.sqrt(2. / (a.outs[l] + a.ins[l])))        # BNN_SGDMC.py:65
            # sum of Normal(0, std).log_prob(W)                                    (:66)
            log_p = log_p - (W * W).sum() / (2. * std * std) - W.numel() * (math.log(std) + 0.5 * math.log(2. * math.pi))
            h = torch.addmm(b, h, W.t())
            if l != a.n_linear - 1:
                h = self._act_fn(h)
        log_precs = theta[a.stride: a.stride + self.nout]
        precs = log_precs.exp()
        log_lik = (-0.5 * precs * (by - h) ** 2 + 0.5 * log_precs - 0.5 * math.log(2. * math.pi)).sum()   # :72-74
        al, be = float(self.alpha_n), float(self.beta_n)
        # Gamma(al, be).log_prob(exp(l)) + l  (TransformedDistribution with ExpTransform().inv, :47,62)
        log_p = log_p + (al * math.log(be) + (al - 1.) * log_precs - be * precs - math.lgamma(al) + log_precs).sum()
        return -1 * (log_lik * (num_train / bx.shape[0]) + log_p)

    @property
    def _act_fn(self):
        return {"relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "identity": lambda t: t}[self.arch.act]

    # ---- the chain ----------------------------------------------------------------------------
    def _chain_init(self, dev):
        a = self.arch
        n = (a.stride + self.nout + 3) // 4 * 4
        self._theta = torch.zeros(n, device=dev)
        self._theta[: a.stride] = a.pack([self.nn.nn])[0].to(dev)
        self._theta[a.stride: a.stride + self.nout] = self.log_precs.data.to(dev)
        self._V = torch.ones(n, device=dev)            # pybnn initialises the preconditioner state to ones
        self._t = 0
        self._valid = torch.zeros(n, device=dev)
        self._valid[: a.stride] = a.valid_mask().to(dev).float()
        self._valid[a.stride: a.stride + self.nout] = 1.
        self._graph = None

    def _grad_autograd(self, X, y, idx, num_train):
        """loss and masked gradient of U by torch autograd (reference implementation of the gradient; ~150 tiny kernels
        per step at 3x512 -- kept for the test that pins _grad_into against it and for activations without a closed-form
        derivative here)."""
        th = self._theta.detach().requires_grad_(True)
        loss = self._neg_log_post(th, X[idx], y[idx], num_train)
        grad, = torch.autograd.grad(loss, th)
        return loss.detach(), grad * self._valid

    def _grad_buffers(self, B, dev):
        """Scratch of the hand-derived gradient for batch size B: augmented layer inputs [B, ld] = (h | 1 | 0-pad), so that
        one GEMM against a packed block [out, ld] = (W | b | pad) gives W h + b, and dz^T @ (h | 1 | 0) gives (dW | db | 0)."""
        # one set per batch size, never freed while the chain lives: a captured CUDA graph keeps pointing at its set
        cache = self.__dict__.setdefault("_gb_cache", {})
        key = (int(B), str(dev), int(self._theta.data_ptr()))
        if key not in cache:
            a = self.arch
            hs = []
            for l in range(a.n_linear):
                h = torch.zeros(B, a.lds[l], device=dev)
                h[:, a.ins[l]] = 1.
                hs.append(h)
            grad = torch.zeros_like(self._theta)
            # d(-log prior)/dW = W / std_l^2 on weight entries only (BNN_SGDMC.py:65-66), 0 on biases / padding
            ps = torch.zeros_like(self._theta)
            logc = 0.
            for l in range(a.n_linear):
                std = float(self.gain * np.sqrt(2. / (a.outs[l] + a.ins[l])))
                blk = ps[a.offsets[l]: a.offsets[l] + a.outs[l] * a.lds[l]].view(a.outs[l], a.lds[l])
                blk[:, : a.ins[l]] = 1. / (std * std)
                logc += a.outs[l] * a.ins[l] * (math.log(std) + 0.5 * math.log(2. * math.pi))
            cache[key] = (hs, grad, ps, logc)
        return cache[key]

    def _grad_into(self, X, y, idx, num_train):
        """loss U (BNN_SGDMC.py:61-74,83) and its gradient at the current chain state for the minibatch idx, derived by hand
        on the packed parameter vector: L forward GEMMs, L + (L - 1) backward GEMMs and a handful of elementwise kernels
        (~40 launches instead of autograd's ~150), graph-capturable (no host tensors, no allocation-dependent control flow).
        tanh / relu / sigmoid / identity activations; pinned against autograd by tests/test_sampling_gpu.py."""
        with torch.no_grad():
            return self._grad_manual(X, y, idx, num_train)

    def _grad_manual(self, X, y, idx, num_train, gathered=None):
        """gathered = (B, y_batch): the minibatch is already in the augmented first-layer input of the B-row buffer set
        (bnn_sgld_gather, CUDA-graph path); otherwise it is gathered here from idx."""
        a = self.arch
        th = self._theta
        if gathered is None:
            bx, by = X[idx], y[idx]
            B = bx.shape[0]
        else:
            B, by = gathered
        hs, grad, ps, logc = self._grad_buffers(B, th.device)
        L = a.n_linear
        blks = [th[a.offsets[l]: a.offsets[l] + a.outs[l] * a.lds[l]].view(a.outs[l], a.lds[l]) for l in range(L)]
        if gathered is None:
            hs[0][:, : a.ins[0]] = bx
        for l in range(L):
            z = torch.mm(hs[l], blks[l].t())                       # W h + b (bias column meets the ones column)
            if l != L - 1:
                dst = hs[l + 1][:, : a.outs[l]]
                if a.act == "tanh":
                    torch.tanh(z, out=dst)
                elif a.act == "relu":
                    torch.clamp(z, min=0., out=dst)
                elif a.act == "sigmoid":
                    torch.sigmoid(z, out=dst)
                else:
                    dst.copy_(z)
        out = z                                                      # [B, nout]
        lp = th[a.stride: a.stride + self.nout]
        scale = float(num_train) / B
        al, be = float(self.alpha_n), float(self.beta_n)
        if th.is_cuda:
            # prior gradient + energy (bnn_sgld_prior), then residuals, loss U, dU/d(log_precs), dU/d(out) (bnn_sgld_head)
            scr = self.__dict__.setdefault("_head_scr", {})
            k = (B, self.nout, str(th.device), int(th.data_ptr()))
            if k not in scr:
                scr[k] = (torch.empty(B, self.nout, device=th.device), torch.zeros((), device=th.device),
                          torch.empty_like(th), torch.zeros(65, device=th.device), torch.zeros((), device=th.device))
            dz, loss, pw, pscr, quad = scr[k]
            ops.sgld_prior(ps, th, pw, pscr, quad)
            ops.sgld_head(out.contiguous(), by.contiguous(), lp, scale, al, be, quad, logc, dz,
                          grad[a.stride: a.stride + self.nout], loss)
        else:                                                        # same arithmetic in torch (CPU tensors: tests only)
            pw = ps * th                                             # W / std^2 on weight entries
            precs = lp.exp()
            r = out - by
            rr = (r * r).sum(0)
            # U = -(scale * loglik + logprior)
            log_lik = (-0.5 * precs * rr + B * (0.5 * lp - 0.5 * math.log(2. * math.pi))).sum()
            log_p = -0.5 * torch.dot(pw, th) - logc + \
                (al * math.log(be) + (al - 1.) * lp - be * precs - math.lgamma(al) + lp).sum()
            loss = -1 * (log_lik * scale + log_p)
            grad[a.stride: a.stride + self.nout] = -scale * (-0.5 * precs * rr + 0.5 * B) - (al - be * precs)
            dz = (scale * precs) * r                                 # dU/dout
        for l in range(L - 1, -1, -1):
            gb = grad[a.offsets[l]: a.offsets[l] + a.outs[l] * a.lds[l]].view(a.outs[l], a.lds[l])
            pb = pw[a.offsets[l]: a.offsets[l] + a.outs[l] * a.lds[l]].view(a.outs[l], a.lds[l])
            torch.addmm(pb, dz.t(), hs[l], out=gb)                   # (dW + W / std^2 | db | 0)
            if l > 0:
                dh = torch.mm(dz, blks[l])[:, : a.ins[l]]           # dU/dh_{l-1}
                h = hs[l][:, : a.ins[l]]
                if a.act == "tanh":
                    dz = torch.ops.aten.tanh_backward(dh, h)
                elif a.act == "relu":
                    dz = dh * (h > 0)
                elif a.act == "sigmoid":
                    dz = torch.ops.aten.sigmoid_backward(dh, h)
                else:
                    dz = dh.contiguous()
        return loss, grad

    def _build_graph(self, X, y, num_train):
        """Capture ONE WHOLE SGLD STEP in a CUDA graph: minibatch gather from the epoch permutation at a device-side
        offset, forward + backward (hand-derived, a few GEMMs; SURVEY.md 8f ranks a fused step kernel next), the fused pSGLD update
        with the LR-schedule factor computed in-kernel from a device step counter, and the counter increments.  At
        1x50 .. 3x512 the step is launch-bound: one replay instead of ~60 eager launches + Python is the difference
        between ~2k and >10k steps/s."""
        dev = self._theta.device
        a = self.arch
        self._g_off = torch.zeros((), dtype=torch.long, device=dev)           # position inside the epoch permutation
        self._g_step = torch.full((), self._t, dtype=torch.long, device=dev)  # t of the LR schedule / Philox stream
        self._g_loss = torch.zeros((), device=dev)
        self._g_perm = torch.zeros(num_train, dtype=torch.long, device=dev)

        B = self.batch_size
        self._g_yb = torch.empty(B, self.nout, device=dev)
        h0 = self._grad_buffers(B, dev)[0][0]                               # augmented first-layer input of the B-row set

        def one_step():
            ops.sgld_gather(X, y, self._g_perm, self._g_off, h0, self._g_yb)      # DataLoader batch at the device-side offset
            with torch.no_grad():
                loss, grad = self._grad_manual(X, y, None, num_train, gathered=(B, self._g_yb))
            ops.psgld_step_dev(self._theta, grad, self._V, self._g_step, self.lr_weight, lr_tail=self.lr_noise,
                               n_head=a.stride, alpha=self.psgld_alpha, lam=self.psgld_lambda, seed=self.seed)
            ops.sgld_advance(self._g_step, self._g_off, B)
            self._g_loss = loss                                             # the head kernel's own output buffer (no copy)

        # warm-up on a side stream (autograd/cuBLAS workspaces), on scratch copies of the chain state
        keep = (self._theta.clone(), self._V.clone())
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                one_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        self._theta.copy_(keep[0]); self._V.copy_(keep[1])
        self._g_step.fill_(self._t); self._g_off.zero_()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            one_step()
        # the capture itself does not execute: state is untouched, counters still at (t, 0)
        self._graph = graph

    def _sgld_step(self, X, y, idx, num_train):
        """One eager SGLD step (BNN_SGDMC.py:81-87) -- used without CUDA graphs and for the short last batch of an epoch."""
        a = self.arch
        loss, grad = self._grad_into(X, y, idx, num_train)
        f = float(np.float32((1 + self._t) ** -0.33))                    # LambdaLR, BNN_SGDMC.py:101
        ops.psgld_step(self._theta, grad.contiguous(), self._V, self.lr_weight * f, lr_tail=self.lr_noise * f, n_head=a.stride,
                       alpha=self.psgld_alpha, lam=self.psgld_lambda, seed=self.seed, step=self._t)
        self._t += 1
        return loss

    def sgld_steps(self, num_steps, X, y, gen):
        """BNN_SGDMC.py:76-91: epochs of shuffled minibatches (last one may be short) until num_steps are done."""
        num_train = X.shape[0]
        step_cnt, loss = 0, torch.zeros(())
        if self.use_graph and self._graph is None and num_train >= self.batch_size:
            self._build_graph(X, y, num_train)
        while step_cnt < num_steps:
            if self._perm is None or self._perm_pos >= num_train:
                self._perm = torch.randperm(num_train, generator=gen, device=X.device)   # DataLoader(shuffle=True), :110
                self._perm_pos = 0
                if self._graph is not None:
                    self._g_perm.copy_(self._perm)
                    self._g_off.zero_()
            full = self._perm_pos + self.batch_size <= num_train
            if self._graph is not None and full:
                self._graph.replay()                                     # gather + fwd + bwd + update + counters
                self._t += 1
                loss = self._g_loss
            else:
                idx = self._perm[self._perm_pos: self._perm_pos + self.batch_size]
                loss = self._sgld_step(X, y, idx, num_train)
                if self._graph is not None:
                    self._g_step.fill_(self._t)
            self._perm_pos += self.batch_size
            step_cnt += 1
        return loss

    def train(self, X, y):
        dev = self._dev()
        a = self.arch
        X = self._to_dev(X)
        y = self._to_dev(y).view(-1, self.nout)
        if not self.warm_start or not hasattr(self, "_theta"):
            if not self.warm_start:
                self.init_nn()
            self._chain_init(dev)
        self._graph = None
        self._perm, self._perm_pos = None, 0
        gen = torch.Generator(device=dev).manual_seed(self.seed)
        self.sgld_steps(self.steps_burnin, X, y, gen)                     # burn-in
        rows, step_cnt = [], 0
        while step_cnt < self.steps:
            loss = self.sgld_steps(self.keep_every, X, y, gen)
            step_cnt += self.keep_every
            self._theta.mul_(self._valid)                                 # Langevin noise also lands on layout padding: re-zero it
            rows.append(self._theta[: a.stride].clone())                  # thinning snapshot = one bank row (:124)
            prec = self._theta[a.stride: a.stride + self.nout].exp().mean()
            print('Step %4d, loss = %8.2f, precision = %g' % (step_cnt, float(loss), float(prec)), flush=True)
        self.nns = SampleBank(a, torch.stack(rows).contiguous())
        self.log_precs.data = self._theta[a.stride: a.stride + self.nout].detach().cpu()
        self.nn.nn.load_state_dict(a.to_module(self._theta[: a.stride]).state_dict())
        print('Number of samples: %d' % len(self.nns))

    # ---- the hot path (BNN_SGDMC.py:127-137) --------------------------------------------------
    def sample(self, num_samples=1):
        assert (num_samples <= len(self.nns))
        idx = np.random.permutation(len(self.nns))[:num_samples]
        return self._as_bank(self.nns)[idx]

    def sample_predict(self, nns, input):
        return self._sample_predict(nns, input, squeeze=False)        # [S, N, nout]

    def report(self):
        print(self.nn.nn)
