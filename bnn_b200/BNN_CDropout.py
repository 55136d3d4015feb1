"""Drop-in for the reference's BNN_CDropout (BNN_CDropout.py:95-167): concrete dropout.

sample(): the relaxed-Bernoulli keep masks z in (0,1) (temperature 0.1, every layer including
in == 1, BNN_CDropout.py:59-65 + util.py:45-50) for all S samples come from one launch of the mask
kernel; the forward scales each layer's input units by them over the one base network.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.distributions.utils import clamp_probs, probs_to_logits

from . import ops
from .BNN import BNN
from .bank import SampleBank
from .layout import Arch


class CDropout(nn.Module):
    def __init__(self, p=0.5):
        super().__init__()
        self.p_logit = nn.Parameter(probs_to_logits(torch.as_tensor(p), is_binary=True))

    def dropout_rate(self):
        return clamp_probs(torch.sigmoid(self.p_logit))                # BNN_CDropout.py:18-19

    def forward(self, input):                                          # training-time relaxation, :21-34
        p = self.dropout_rate()
        eps, temp = 1e-7, 0.1
        u = torch.rand_like(input)
        drop = torch.log(p + eps) - torch.log(1 - p + eps) + torch.log(u + eps) - torch.log(1 - u + eps)
        return input * (1 - torch.sigmoid(drop / temp))


class CDropoutLinear(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        self.layer = nn.Sequential(CDropout(), nn.Linear(in_features, out_features))

    def forward(self, input):
        return self.layer(input)

    def dropout_rate(self):
        return self.layer[0].dropout_rate()

    def reg(self):                                                     # BNN_CDropout.py:53-57
        p = self.layer[0].dropout_rate()
        entropy = -1 * (p * p.log() + (1 - p) * (1 - p).log())
        return torch.sum(self.layer[1].weight ** 2), 1. * self.layer[1].in_features * entropy


class NN_CDropout(nn.Module):
    def __init__(self, dim, act=nn.ReLU(), num_hiddens=[50]):
        super().__init__()
        layers, pre = [], dim
        for h in num_hiddens:
            layers += [CDropoutLinear(pre, h), act]
            pre = h
        layers.append(CDropoutLinear(pre, 1))
        self.nn = nn.Sequential(*layers)

    def cd_layers(self):
        return [l for l in self.nn if isinstance(l, CDropoutLinear)]

    def forward(self, input):
        return self.nn(input)


class BNN_CDropout(BNN):
    def __init__(self, dim, act=nn.ReLU(), num_hiddens=[50], conf=dict()):
        super().__init__()
        self.dim, self.act, self.num_hiddens = dim, act, num_hiddens
        self.num_epochs = conf.get('num_epochs', 400)
        self.batch_size = conf.get('batch_size', 32)
        self.print_every = conf.get('print_every', 100)
        self.normalize = conf.get('normalize', True)
        self.min_noise = conf.get('min_noise', 0.)
        self.lr = conf.get('lr', 1e-2)
        self.lscale = conf.get('lscale', 1e-1)
        self.dr = conf.get('dr', 2.)
        self.seed = int(conf.get('seed', 0))
        self.precision = conf.get('precision', 'auto')   # tensor cores (3xTF32, FP32-grade) when the shape allows, else FP32 SIMT
        self.device = torch.device(conf['device']) if 'device' in conf else None
        self.arch = Arch([dim] + list(num_hiddens) + [1], act)
        self.layer_masked = [True] * self.arch.n_linear               # every layer, BNN_CDropout.py:59-65
        self.nn = NN_CDropout(dim, self.act, self.num_hiddens)
        self.noise_level = 1.
        self._draws = 0
        self.fused_train = bool(conf.get('fused_train', True))   # the single-SM fused training kernel when the network fits it
        self.trained_with = None

    def _train_fused(self, X, y):
        """BNN_CDropout.train (BNN_CDropout.py:114-139) through the fused single-SM training kernel (bnn_cdropout_train_run): one
        launch per epoch; the network, the p_logits, the Adam moments AND noise_level stay on the device (the launch that ends an
        epoch updates noise_level itself), statistics are read back only when an epoch is printed."""
        dev = self._dev()
        num_train = X.shape[0]
        cls = self.nn.cd_layers()
        tr = ops.CDropoutTrainer(self.arch, self.batch_size, dev, lr=self.lr, lscale=self.lscale, dr=self.dr,
                                 min_noise=self.min_noise, noise_level=self.noise_level, seed=self.seed)
        tr.set_state(self.arch.pack([[(l.layer[1].weight.detach(), l.layer[1].bias.detach()) for l in cls]])[0],
                     torch.stack([l.layer[0].p_logit.detach().reshape(()) for l in cls]))
        tr.bind(X, y.reshape(num_train, 1))
        gen = torch.Generator(device="cpu").manual_seed(self.seed)
        n_steps = (num_train + self.batch_size - 1) // self.batch_size
        for epoch in range(self.num_epochs):
            perm = torch.randperm(num_train, generator=gen).to(dev)       # DataLoader(shuffle=True), BNN_CDropout.py:118
            show = (epoch + 1) % self.print_every == 0
            stats = tr.run(perm, 0, n_steps, epoch_end=True, want_stats=show)
            if show:
                mse, wreg, ent = [float(v) for v in stats.sum(dim=0)]
                print("Epoch %4d, mse = %g, noise = %g, wreg = %g, -entropy = %g" %
                      (epoch + 1, mse / num_train, float(tr.noise_level), wreg / num_train, -1 * ent / num_train))
        self.noise_level = float(tr.noise_level)
        with torch.no_grad():
            for l, (W, b), pl in zip(cls, self.arch.unpack(tr.W), tr.p_logit):
                l.layer[1].weight.copy_(W)
                l.layer[1].bias.copy_(b)
                l.layer[0].p_logit.copy_(pl)
        self.trained_with = "fused"

    def reg(self):                                                     # BNN_CDropout.py:141-151
        entropy, weight_reg = 0., 0.
        noise_var = self.noise_level ** 2
        for layer in self.nn.cd_layers():
            prob = 1 - layer.dropout_rate()
            w2, ent = layer.reg()
            weight_reg = weight_reg + self.lscale ** 2 * prob * noise_var * w2
            entropy = entropy + self.dr * 2 * noise_var * ent
        return weight_reg, entropy

    def train(self, X, y):
        dev = self._dev()
        X, y = self._to_dev(X), self._to_dev(y)
        num_train = X.shape[0]
        if self.fused_train and dev.type == "cuda" and ops.cdropout_train_supported(self.arch, self.batch_size):
            self.nn.to(dev)
            self._train_fused(X.contiguous().float(), y.contiguous().float())
            self.nn.eval()
            return
        self.trained_with = "torch"
        self.nn.to(dev).train()
        opt = torch.optim.Adam(self.nn.parameters(), lr=self.lr)
        crit = nn.MSELoss(reduction='sum')
        gen = torch.Generator(device="cpu").manual_seed(self.seed)
        for epoch in range(self.num_epochs):
            epoch_mse = 0.
            perm = torch.randperm(num_train, generator=gen).to(dev)
            for i in range(0, num_train, self.batch_size):
                idx = perm[i: i + self.batch_size]
                opt.zero_grad()
                mse_loss = crit(self.nn(X[idx]).squeeze(-1), y[idx])
                wreg, ent = self.reg()
                frac = idx.shape[0] / num_train
                (mse_loss + (wreg * frac - ent * frac)).backward()
                opt.step()
                epoch_mse += float(mse_loss)
            self.noise_level = max(self.min_noise, np.sqrt(epoch_mse / num_train))
            if (epoch + 1) % self.print_every == 0:
                print("Epoch %4d, mse = %g, noise = %g" % (epoch + 1, epoch_mse / num_train, self.noise_level))
        self.nn.eval()

    def drop_probs(self):
        return [float(l.dropout_rate().detach()) for l in self.nn.cd_layers()]

    def sample(self, num_samples=1, uniforms=None):
        """uniforms (optional, [S, mask_len]) injects the reference's torch.rand draws for exact parity."""
        dev = self._dev()
        lin = [(l.layer[1].weight, l.layer[1].bias) for l in self.nn.cd_layers()]
        base = self._cached_pack("base", [t for wb in lin for t in wb], lambda: self.arch.pack([lin])[0].to(dev))
        cfg = ops.MaskConfig("concrete", self.layer_masked, p_drop=self.drop_probs(), temperature=0.1, seed=self.seed,
                             sample_offset=self._draws)
        if uniforms is not None:
            uniforms = self._to_dev(uniforms)
        else:
            self._draws += num_samples
        masks = ops.dropout_masks(self.arch, cfg, num_samples, dev, uniforms=uniforms)
        return SampleBank(self.arch, base, masks=masks, layer_masked=self.layer_masked)

    def sample_predict(self, nns, X):
        return self._sample_predict(nns, X, squeeze=True)

    def report(self):
        print(self.nn.nn)
