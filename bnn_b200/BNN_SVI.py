"""Drop-in for the reference's BNN_SVI (BNN_SVI.py:11-79): mean-field SVI.

The reference needs pyro (absent offline; UNPINNED).  The guide is Normal(mu_n, softplus(sigma_n)) per
parameter tensor, weights AND biases (BNN_SVI.py:40-48); sample() draws all S networks with the fused
reparam kernel in SVI scale mode (std = softplus(sigma), NOT BBB's sqrt(softplus)).  train() restates
the model/guide pair (Normal(0, weight_std) prior, Normal(pred, noise_level) likelihood, minibatch
subsampling; BNN_SVI.py:25-38,50-64) as a single-sample reparameterised ELBO in plain torch.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .BNN import BNN
from .bank import SampleBank
from .layout import Arch
from .util import NN


class BNN_SVI(BNN):
    def __init__(self, dim, act=nn.ReLU(), num_hiddens=[50], conf=dict()):
        super().__init__()
        self.dim, self.act, self.num_hiddens = dim, act, num_hiddens
        self.num_iters = conf.get('num_iters', 4000)
        self.batch_size = conf.get('batch_size', 32)
        self.print_every = conf.get('print_every', 100)
        self.lr = conf.get('lr', 1e-2)
        self.weight_std = conf.get('weight_std', 1.0)
        self.noise_level = conf.get('noise_level', None)
        self.seed = int(conf.get('seed', 0))
        self.precision = conf.get('precision', 'auto')   # tensor cores (3xTF32, FP32-grade) when the shape allows, else FP32 SIMT
        self.device = torch.device(conf['device']) if 'device' in conf else None
        self.arch = Arch([dim] + list(num_hiddens) + [1], act)
        self.nn = NN(dim, self.act, self.num_hiddens, nout=1)
        # guide parameters in the packed layout (pyro.param "mu_<n>", "sigma_<n>"; init BNN_SVI.py:44-45)
        g = torch.Generator().manual_seed(self.seed)
        valid = self.arch.valid_mask().float()
        self.mu = (self.weight_std * torch.randn(self.arch.stride, generator=g)) * valid
        self.sigma = torch.randn(self.arch.stride, generator=g) * valid
        self._draws = 0

    def _net(self, theta, X):
        a, h = self.arch, X
        f = {"relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "identity": lambda t: t}[a.act]
        for l in range(a.n_linear):
            blk = theta[a.offsets[l]: a.offsets[l] + a.outs[l] * a.lds[l]].view(a.outs[l], a.lds[l])
            h = torch.addmm(blk[:, a.ins[l]], h, blk[:, : a.ins[l]].t())
            if l != a.n_linear - 1:
                h = f(h)
        return h.squeeze(-1)

    def train(self, X, y):
        if self.noise_level is None:
            print("No noise level provided, use noise_level = 0.05 * y.std()")
            self.noise_level = 0.05 * y.std()
        dev = self._dev()
        X, y = self._to_dev(X), self._to_dev(y).reshape(-1)
        num_train = X.shape[0]
        valid = self.arch.valid_mask().to(dev)
        mu = self.mu.to(dev).requires_grad_(True)
        sigma = self.sigma.to(dev).requires_grad_(True)
        opt = torch.optim.Adam([mu, sigma], lr=self.lr)
        gen = torch.Generator(device=dev).manual_seed(self.seed)
        noise = float(self.noise_level)
        self.rec = []
        for i in range(self.num_iters):
            opt.zero_grad()
            idx = torch.randperm(num_train, generator=gen, device=dev)[: min(num_train, self.batch_size)]
            std = F.softplus(sigma)
            theta = (mu + std * torch.randn(mu.shape, generator=gen, device=dev)) * valid
            log_lik = torch.distributions.Normal(self._net(theta, X[idx]), noise).log_prob(y[idx]).sum() * (num_train / idx.shape[0])
            log_prior = torch.distributions.Normal(0., self.weight_std).log_prob(theta)[valid].sum()
            log_q = torch.distributions.Normal(mu, std).log_prob(theta.detach() * 0 + theta)[valid].sum()
            loss = -(log_lik + log_prior - log_q)
            loss.backward()
            opt.step()
            if (i + 1) % self.print_every == 0:
                self.rec.append(float(loss) / num_train)
                print("[Iteration %05d/%05d] loss: %-4.3f" % (i + 1, self.num_iters, float(loss) / num_train))
        self.mu, self.sigma = (mu.detach() * valid).cpu(), (sigma.detach() * valid).cpu()

    def sample(self, num_samples=1, eps=None):
        dev = self._dev()
        W, _ = ops.reparam_kl(self.arch, self.mu.to(dev).contiguous(), self.sigma.to(dev).contiguous(), num_samples,
                              mode="svi", eps=eps, seed=self.seed, sample_offset=self._draws)
        self._draws += num_samples
        return SampleBank(self.arch, W)

    def sample_predict(self, nns, X):
        return self._sample_predict(nns, X, squeeze=True)

    def report(self):
        print(self.nn)
