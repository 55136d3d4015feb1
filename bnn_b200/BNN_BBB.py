"""Drop-in for the reference's BNN_BBB (BNN_BBB.py:85-152): Bayes by Backprop.

sample(): all S x P reparameterised draws w = mu + sqrt(softplus(max(rho,-35))) * eps in ONE launch
of the fused reparam kernel with in-kernel Philox (replaces GaussianLinear.rsample/sample_linear and
S x L fresh nn.Linear constructions, BNN_BBB.py:24-27,44-51,65-69,139-141); sample_predict /
predict_mv / validate: the fused forward.  train(): the reference's recipe (local reparameterisation,
analytic KL, Adam; BNN_BBB.py:29-37,101-137) runs in the fused single-SM training kernel (bnn_bbb_train_run,
SURVEY.md 8 row f2) when the network fits one SM's shared memory -- the reference's 1-50-50-1 does --,
else in plain torch on the device.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.parameter import Parameter

from . import ops
from .BNN import BNN
from .bank import SampleBank
from .layout import Arch
from .util import stable_noise_var


class GaussianLinear(nn.Module):
    """mu, rho of shape [out, in + 1], bias in the last column (BNN_BBB.py:14-22)."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.mu = Parameter(torch.zeros(out_features, 1 + in_features))
        self.rho = Parameter(torch.zeros(out_features, 1 + in_features))
        nn.init.xavier_uniform_(self.mu[:, :-1])
        nn.init.constant_(self.rho, val=-5.)

    def forward(self, input):                                  # local reparameterisation, BNN_BBB.py:29-37
        out_mu = F.linear(input, self.mu[:, :-1], self.mu[:, -1])
        out_std = F.linear(input ** 2, stable_noise_var(self.rho[:, :-1]), stable_noise_var(self.rho[:, -1])).sqrt()
        return out_mu + torch.randn_like(out_mu) * out_std


class BayesianNN(nn.Module):
    def __init__(self, dim, act=nn.ReLU(), num_hiddens=[50]):
        super().__init__()
        self.dim, self.act, self.num_hiddens = dim, act, num_hiddens
        layers, pre = [], dim
        for h in num_hiddens:
            layers += [GaussianLinear(pre, h), act]
            pre = h
        layers.append(GaussianLinear(pre, 1))
        self.nn = nn.Sequential(*layers)

    def gaussian_layers(self):
        return [l for l in self.nn if isinstance(l, GaussianLinear)]

    def forward(self, input):
        return self.nn(input)


class BNN_BBB(BNN):
    def __init__(self, dim, act=nn.ReLU(), num_hiddens=[50], conf=dict()):
        super().__init__()
        self.dim, self.act, self.num_hiddens = dim, act, num_hiddens
        self.num_epochs = conf.get('num_epochs', 2000)
        self.batch_size = conf.get('batch_size', 32)
        self.print_every = conf.get('print_every', 100)
        self.lr = conf.get('lr', 1e-2)
        self.weight_std = conf.get('weight_std', 0.1)
        self.kl_factor = conf.get('kl_factor', 1e-3)
        self.seed = int(conf.get('seed', 0))
        self.fused_train = bool(conf.get('fused_train', True))   # the single-SM fused training kernel when the network fits it
        self.trained_with = None
        self.precision = conf.get('precision', 'auto')   # tensor cores (3xTF32, FP32-grade) when the shape allows, else FP32 SIMT
        self.device = torch.device(conf['device']) if 'device' in conf else None
        self.arch = Arch([dim] + list(num_hiddens) + [1], act)
        self.nn = BayesianNN(dim, self.act, self.num_hiddens)
        self._draws = 0                                        # Philox subsequence counter: every draw is fresh

    def _packed_mu_rho(self):
        """(mu, rho) packed bias-last on the device; cached until a parameter is modified in place or replaced (tensor
        version counters), so that repeated sample() calls do not re-upload and re-pack the variational parameters."""
        gls = self.nn.gaussian_layers()
        dev = self._dev()
        return self._cached_pack("mu_rho", [p for l in gls for p in (l.mu, l.rho)],
                                 lambda: (self.arch.pack_bias_last([l.mu.detach().to(dev) for l in gls]),
                                          self.arch.pack_bias_last([l.rho.detach().to(dev) for l in gls])))

    def kl(self):
        """KL(q || N(0, weight_std)) of BNN_BBB.loss (BNN_BBB.py:107-110), by the fused kernel."""
        mu, rho = self._packed_mu_rho()
        _, kl = ops.reparam_kl(self.arch, mu, rho, 0, mode="bbb", weight_std=self.weight_std)
        return kl.float().cpu()

    def loss(self, X, y):
        num_x = X.shape[0]
        pred = self.nn(X.reshape(num_x, self.dim))
        mse = nn.MSELoss(reduction='mean')(pred, y.reshape(num_x, 1))
        kld = 0.
        for layer in self.nn.gaussian_layers():
            kld = kld + torch.distributions.kl_divergence(
                torch.distributions.Normal(layer.mu, stable_noise_var(layer.rho).sqrt()),
                torch.distributions.Normal(0., self.weight_std)).sum()
        return mse, kld

    def _train_fused(self, X, y):
        """BNN_BBB.train (BNN_BBB.py:112-137) through the fused single-SM training kernel (bnn_bbb_train_run): one launch per
        epoch, variational parameters and Adam moments device-resident, statistics read back only when an epoch is printed."""
        dev = self._dev()
        num_x = X.shape[0]
        gls = self.nn.gaussian_layers()
        tr = ops.BbbTrainer(self.arch, self.batch_size, dev, lr=self.lr, kl_factor=self.kl_factor, weight_std=self.weight_std,
                            seed=self.seed)
        tr.set_state(self.arch.pack_bias_last([l.mu.detach() for l in gls]), self.arch.pack_bias_last([l.rho.detach() for l in gls]))
        tr.bind(X, y.reshape(num_x, 1))
        gen = torch.Generator(device="cpu").manual_seed(self.seed)
        n_steps = (num_x + self.batch_size - 1) // self.batch_size
        for epoch in range(self.num_epochs):
            perm = torch.randperm(num_x, generator=gen).to(dev)           # DataLoader(shuffle=True), BNN_BBB.py:119
            show = (epoch + 1) % self.print_every == 0
            stats = tr.run(perm, 0, n_steps, want_stats=show)
            if show:
                mse, kl = stats
                rows = torch.full((n_steps,), float(self.batch_size), device=dev)
                rows[-1] = num_x - (n_steps - 1) * self.batch_size
                epoch_mse = float((mse * rows).sum() / num_x)
                epoch_kl = self.kl_factor * float(kl[-1])
                print("[Epoch %5d, loss = %8.2f (KL = %8.2f, mse = %8.2f), kl_factor = %g]" %
                      (epoch + 1, epoch_mse + epoch_kl, epoch_kl, epoch_mse, self.kl_factor))
        with torch.no_grad():
            for l, layer in enumerate(gls):
                layer.mu.copy_(self.arch.layer_view(tr.mu, l)[:, : self.arch.ins[l] + 1])
                layer.rho.copy_(self.arch.layer_view(tr.rho, l)[:, : self.arch.ins[l] + 1])
        self.trained_with = "fused"

    def train(self, X, y):
        dev = self._dev()
        num_x = X.shape[0]
        X = self._to_dev(X).reshape(num_x, self.dim)
        y = self._to_dev(y).reshape(num_x)
        self.nn.to(dev)
        if self.fused_train and dev.type == "cuda" and ops.bbb_train_supported(self.arch, self.batch_size):
            self.nn.train()
            self._train_fused(X.contiguous(), y.contiguous())
            self.nn.eval()
            return
        self.trained_with = "torch"
        self.nn.train()
        opt = torch.optim.Adam(self.nn.parameters(), self.lr)
        gen = torch.Generator(device="cpu").manual_seed(self.seed)
        for epoch in range(self.num_epochs):
            epoch_kl, epoch_mse = 0., 0.
            perm = torch.randperm(num_x, generator=gen).to(dev)
            for i in range(0, num_x, self.batch_size):
                idx = perm[i: i + self.batch_size]
                opt.zero_grad()
                mse, kl_term = self.loss(X[idx], y[idx])
                kl_loss = self.kl_factor * kl_term
                (mse + kl_loss).backward()
                opt.step()
                epoch_mse += float(mse) * idx.shape[0]
                epoch_kl = float(kl_loss)
            epoch_mse /= num_x
            if (epoch + 1) % self.print_every == 0:
                print("[Epoch %5d, loss = %8.2f (KL = %8.2f, mse = %8.2f), kl_factor = %g]" %
                      (epoch + 1, epoch_mse + epoch_kl, epoch_kl, epoch_mse, self.kl_factor))
        self.nn.eval()

    def sample(self, num_samples=1, eps=None):
        """eps (optional, [S, stride] packed) injects the reference's own draws for exact parity."""
        mu, rho = self._packed_mu_rho()
        W, _ = ops.reparam_kl(self.arch, mu, rho, num_samples, mode="bbb", eps=eps, seed=self.seed,
                              sample_offset=self._draws)
        self._draws += num_samples
        return SampleBank(self.arch, W)

    def sample_predict(self, nns, X):
        return self._sample_predict(nns, X, squeeze=True)              # [S, N]

    def report(self):
        print(self.nn)
