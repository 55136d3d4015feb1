"""Drop-in for the reference's BNN_SGDMC (BNN_SGDMC.py:19-140): a pSGLD chain over an MLP with a
learned log-precision per output.

What runs where:
  * sample / sample_predict / predict_mv / validate: the fused CUDA forward over the [S, stride]
    bank of thinned chain states (no deepcopy'd nn.Modules);
  * train: the chain state theta = [packed network | log_precs] lives in ONE device vector; each
    SGLD step is torch autograd for the minibatch gradient of U (BNN_SGDMC.py:61-83, the stop-gap
    SURVEY.md 8f ranks next) followed by the fused bnn_psgld_step kernel (update + RMSprop
    preconditioner + in-kernel Philox noise + both LR groups, BNN_SGDMC.py:86-87,96-101); thinning
    snapshots are row copies into the bank (BNN_SGDMC.py:124).
pybnn is not needed.  The update rule is restated from Li et al. 2016 (UNPINNED).
"""
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
        self.precision = conf.get('precision', 'fp32')
        self.psgld_alpha = conf.get('precondition_decay_rate', 0.99)
        self.psgld_lambda = conf.get('diagonal_bias', 1e-5)
        self.device = torch.device(conf['device']) if 'device' in conf else None

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

    # ---- energy, evaluated on a packed parameter vector (BNN_SGDMC.py:61-74,83) ----------
    def _neg_log_post(self, theta, bx, by, num_train):
        a = self.arch
        h = bx
        log_p = 0.
        for l in range(a.n_linear):
            blk = theta[a.offsets[l]: a.offsets[l] + a.outs[l] * a.lds[l]].view(a.outs[l], a.lds[l])
            W, b = blk[:, : a.ins[l]], blk[:, a.ins[l]]
            std = self.gain * np.sqrt(2. / (a.outs[l] + a.ins[l]))
            log_p = log_p + torch.distributions.Normal(0., std).log_prob(W).sum()
            h = torch.addmm(b, h, W.t())
            if l != a.n_linear - 1:
                h = self._act_fn(h)
        log_precs = theta[a.stride: a.stride + self.nout]
        precs = log_precs.exp()
        log_lik = (-0.5 * precs * (by - h) ** 2 + 0.5 * log_precs - 0.5 * np.log(2 * np.pi)).sum()
        al, be = self.alpha_n.to(theta.device), self.beta_n.to(theta.device)
        log_p = log_p + (torch.distributions.Gamma(al, be).log_prob(precs) + log_precs).sum()
        return -1 * (log_lik * (num_train / bx.shape[0]) + log_p)

    @property
    def _act_fn(self):
        return {"relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "identity": lambda t: t}[self.arch.act]

    def train(self, X, y):
        dev = self._dev()
        a = self.arch
        X = self._to_dev(X)
        y = self._to_dev(y).view(-1, self.nout)
        num_train = X.shape[0]
        if not self.warm_start or not hasattr(self, "_theta"):
            if not self.warm_start:
                self.init_nn()
            n = (a.stride + self.nout + 3) // 4 * 4
            self._theta = torch.zeros(n, device=dev)
            self._theta[: a.stride] = a.pack([self.nn.nn])[0].to(dev)
            self._theta[a.stride: a.stride + self.nout] = self.log_precs.data.to(dev)
            self._V = torch.ones(n, device=dev)            # pybnn initialises the preconditioner state to ones
            self._t = 0
        theta, V = self._theta, self._V
        valid = torch.zeros_like(theta)
        valid[: a.stride] = a.valid_mask().to(dev).float()
        valid[a.stride: a.stride + self.nout] = 1.
        gen = torch.Generator(device="cpu").manual_seed(self.seed)
        bank_rows = []
        step_cnt, total = 0, self.steps_burnin + self.steps
        loss = torch.zeros(())
        while step_cnt < total:
            perm = torch.randperm(num_train, generator=gen).to(dev)      # DataLoader(shuffle=True), BNN_SGDMC.py:110
            for i in range(0, num_train, self.batch_size):
                idx = perm[i: i + self.batch_size]
                th = theta.detach().requires_grad_(True)
                loss = self._neg_log_post(th, X[idx], y[idx], num_train)
                grad, = torch.autograd.grad(loss, th)
                grad = (grad * valid).contiguous()
                f = float(np.float32((1 + self._t) ** -0.33))                 # LambdaLR, BNN_SGDMC.py:101
                ops.psgld_step(theta, grad, V, self.lr_weight * f, lr_tail=self.lr_noise * f, n_head=a.stride,
                               alpha=self.psgld_alpha, lam=self.psgld_lambda, seed=self.seed, step=self._t)
                theta.mul_(valid)                                            # padding stays zero
                self._t += 1
                step_cnt += 1
                if step_cnt > self.steps_burnin and (step_cnt - self.steps_burnin) % self.keep_every == 0:
                    bank_rows.append(theta[: a.stride].clone())
                    prec = theta[a.stride: a.stride + self.nout].exp().mean()
                    print('Step %4d, loss = %8.2f, precision = %g' % (step_cnt - self.steps_burnin, float(loss), float(prec)),
                          flush=True)
                if step_cnt >= total:
                    break
        self.nns = SampleBank(a, torch.stack(bank_rows).contiguous())
        self.log_precs.data = theta[a.stride: a.stride + self.nout].detach().cpu()
        last = a.to_module(theta[: a.stride])
        self.nn.nn.load_state_dict(last.state_dict())
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
