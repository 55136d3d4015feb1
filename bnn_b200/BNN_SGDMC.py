"""Drop-in for the reference's BNN_SGDMC (BNN_SGDMC.py:19-140): a pSGLD chain over an MLP with a
learned log-precision per output.

What runs where:
  * sample / sample_predict / predict_mv / validate: the fused CUDA forward over the [S, stride]
    bank of thinned chain states (no deepcopy'd nn.Modules);
  * train: the chain state theta = [packed network | log_precs] lives in ONE device vector; every run of
    SGLD steps (BNN_SGDMC.py:76-91: shuffled minibatch, log_prior / log_lik, backward, pSGLD update with
    in-kernel Philox noise, both LR groups, LambdaLR factor) is ONE launch of the persistent fused step
    kernel (bnn_sgld_run, csrc/sgld_step.cu); thinning snapshots are row copies into the bank (:124).
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
        self.num_chains = int(conf.get('num_chains', 1))      # ours: chains trained side by side on one GPU, samples pooled
        self._epochs = [0] * max(1, self.num_chains)
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
        self.psgld_mode = int(conf.get('precond_mode', 0))     # 0: G = 1/(lam + sqrt(V)) (Li et al. 2016); 1: G = 1/sqrt(V + lam)
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

    # ---- the reference's two energy terms on the module state (BNN_SGDMC.py:61-74), for callers that use them directly;
    #      the chain itself evaluates them (and their gradient) inside the fused step kernel
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

    # ---- the chain ----------------------------------------------------------------------------
    def _new_chain(self, dev, seed=None):
        """What the reference's train() builds at BNN_SGDMC.py:96-101: a FRESH optimizer (V = 1) and scheduler (t = 0) over
        the CURRENT self.nn / self.log_precs -- also on a warm start, which only skips init_nn() (:114-115)."""
        a = self.arch
        chain = ops.SgldChain(a, self.batch_size, dev, self.lr_weight, self.lr_noise, float(self.alpha_n), float(self.beta_n),
                              seed=self.seed if seed is None else seed, precond_decay=self.psgld_alpha, diagonal_bias=self.psgld_lambda, gain=self.gain,
                              precond_mode=self.psgld_mode)
        chain.set_state(a.pack([self.nn.nn])[0], self.log_precs.data)
        return chain

    def sgld_steps(self, num_steps, num_train, chain=None, c=0):
        """BNN_SGDMC.py:76-91: iterate the shuffled loader (a NEW permutation every time the loader is re-entered, i.e. at
        every call and after every exhausted epoch; the last batch of an epoch may be short) until num_steps are done.
        Each run of consecutive steps inside one epoch is ONE launch of the fused step kernel."""
        chain = self._chain if chain is None else chain
        B = self.batch_size
        per_epoch = (num_train + B - 1) // B
        step_cnt, loss = 0, None
        while step_cnt < num_steps:
            n = min(num_steps - step_cnt, per_epoch)
            perm = ops.sgld_perm(self.seed + c, self._epochs[c], num_train, min(num_train, n * B), chain.device)
            self._epochs[c] += 1
            loss = chain.run(perm, 0, n, want_loss=(step_cnt + n >= num_steps))
            step_cnt += n
        return loss

    @property
    def _epoch(self):
        return self._epochs[0]

    @_epoch.setter
    def _epoch(self, v):
        self._epochs = [int(v)] * max(1, self.num_chains)

    def train(self, X, y):
        """BNN_SGDMC.train (BNN_SGDMC.py:93-125).  conf['num_chains'] = C > 1 (ours) runs C independent chains (seeds seed + c, the
        same initial network) SIDE BY SIDE on this GPU -- each chain's launches on its own stream, the CTAs of a launch capped at
        (#SMs / C) -- and pools their thinned samples into one bank of C x (steps / keep_every) networks; one chain alone leaves
        about a third of the SMs idle.  self.nn / self.log_precs end at chain 0's last state."""
        dev = self._dev()
        a = self.arch
        X = self._to_dev(X)
        y = self._to_dev(y).view(-1, self.nout)
        num_train = X.shape[0]
        if not self.warm_start:
            self.init_nn()
        C = max(1, self.num_chains)
        chains = []
        for c in range(C):
            ch = self._new_chain(dev, seed=self.seed + c)
            ch.bind(X, y)
            chains.append(ch)
        self._chain = chains[0]
        self._epoch = 0
        n_keep = (self.steps + self.keep_every - 1) // self.keep_every
        bank = torch.empty(C, n_keep, a.stride, device=dev)
        main = torch.cuda.current_stream(dev)
        streams = [main] if C == 1 else [torch.cuda.Stream(dev) for _ in range(C)]
        if C > 1:
            ops.sgld_set_grid_cap(max(1, torch.cuda.get_device_properties(dev).multi_processor_count // C))
            start = torch.cuda.Event()
            start.record(main)
            for st in streams:
                st.wait_event(start)
        try:
            for c in range(C):
                with torch.cuda.stream(streams[c]):
                    self.sgld_steps(self.steps_burnin, num_train, chains[c], c)       # burn-in
            step_cnt, k = 0, 0
            while step_cnt < self.steps:
                losses = []
                for c in range(C):
                    with torch.cuda.stream(streams[c]):
                        losses.append(self.sgld_steps(self.keep_every, num_train, chains[c], c))
                        bank[c, k].copy_(chains[c].state[: a.stride])      # thinning snapshot = one bank row (:124)
                step_cnt += self.keep_every
                k += 1
                if C > 1:
                    for st in streams:
                        st.synchronize()
                prec = chains[0].state[a.stride: a.stride + self.nout].exp().mean()
                print('Step %4d, loss = %8.2f, precision = %g' % (step_cnt, float(losses[0]), float(prec)), flush=True)
            for st in streams:
                st.synchronize()
        finally:
            if C > 1:
                ops.sgld_set_grid_cap(0)
        for ch in chains:
            ch.check()
        self.nns = SampleBank(a, bank[:, :k].reshape(C * k, a.stride).contiguous())
        self.log_precs.data = self._chain.state[a.stride: a.stride + self.nout].detach().cpu()
        self.nn.nn.load_state_dict(a.to_module(self._chain.state[: a.stride]).state_dict())
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
