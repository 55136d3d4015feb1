"""Drop-in for the reference's BNN_Dropout (BNN_Dropout.py:23-89): MC dropout.

sample(): one launch of the mask kernel draws the S x sum(in_l) Bernoulli(1-p) keep masks (Philox);
the S "networks" are the ONE trained base network plus those masks -- no deepcopy per sample
(BNN_Dropout.py:67-76).  The forward applies the mask to the layer's input units, which equals the
reference's zeroing of weight columns.  Masks are 0/1 with no 1/(1-p) rescale, one value per
(sample, layer, input unit) and only on layers with in > 1, exactly as the reference.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .BNN import BNN
from .bank import SampleBank
from .layout import Arch
from .util import NN


class NN_Dropout(NN):
    def __init__(self, dim, act=nn.ReLU(), num_hiddens=[50], dropout_rate=0.05):
        super().__init__(dim, act, num_hiddens)
        self.dropout_rate = dropout_rate

    def forward(self, x):                                      # BNN_Dropout.py:15-21
        scale_factor = 1 - self.dropout_rate
        for l in self.nn:
            if type(l) == nn.Linear and x.shape[1] > 1:
                x = F.dropout(x, p=self.dropout_rate, training=self.training) * scale_factor
            x = l(x)
        return x


class BNN_Dropout(BNN):
    def __init__(self, dim, act=nn.ReLU(), num_hiddens=[50], conf=dict()):
        super().__init__()
        self.dim, self.act, self.num_hiddens = dim, act, num_hiddens
        self.num_epochs = conf.get('num_epochs', 400)
        self.batch_size = conf.get('batch_size', 32)
        self.print_every = conf.get('print_every', 100)
        self.dropout_rate = conf.get('dropout_rate', 0.05)
        self.min_noise = conf.get('min_noise', 0.)
        self.l2_reg = conf.get('l2_reg', 1e-6)
        self.lr = conf.get('lr', 1e-2)
        self.seed = int(conf.get('seed', 0))
        self.precision = conf.get('precision', 'auto')   # tensor cores (3xTF32, FP32-grade) when the shape allows, else FP32 SIMT
        self.device = torch.device(conf['device']) if 'device' in conf else None
        self.arch = Arch([dim] + list(num_hiddens) + [1], act)
        self.layer_masked = [i > 1 for i in self.arch.ins]             # BNN_Dropout.py:73
        self.nn = NN_Dropout(dim, self.act, self.num_hiddens, self.dropout_rate)
        self.noise_level = 1.
        self._draws = 0
        self.fused_train = bool(conf.get('fused_train', True))   # the single-SM fused training kernel when the network fits it
        self.trained_with = None

    def _train_fused(self, X, y):
        """BNN_Dropout.train (BNN_Dropout.py:39-65) through the fused single-SM training kernel (bnn_dropout_train_run): one
        launch per epoch, the network and its Adam moments device-resident, one read-back per epoch for noise_level (:63)."""
        dev = self._dev()
        num_train = X.shape[0]
        tr = ops.DropoutTrainer(self.arch, self.batch_size, dev, lr=self.lr, l2_reg=self.l2_reg, p_drop=self.dropout_rate,
                                seed=self.seed)
        tr.set_state(self.arch.pack([self.nn.nn])[0])
        tr.bind(X, y.reshape(num_train, 1))
        gen = torch.Generator(device="cpu").manual_seed(self.seed)
        n_steps = (num_train + self.batch_size - 1) // self.batch_size
        losses = []
        for epoch in range(self.num_epochs):
            perm = torch.randperm(num_train, generator=gen).to(dev)       # DataLoader(shuffle=True), BNN_Dropout.py:53
            losses.append(tr.run(perm, 0, n_steps, want_loss=True).sum())
            if (epoch + 1) % self.print_every == 0:
                epoch_loss = float(losses[-1])
                print("[Epoch %5d, loss = %g, noise_level = %g]" %
                      (epoch + 1, epoch_loss / num_train, max(self.min_noise, np.sqrt(epoch_loss / num_train))))
        if losses:
            self.noise_level = max(self.min_noise, float(np.sqrt(float(losses[-1]) / num_train)))
        with torch.no_grad():
            lins = [m for m in self.nn.nn if isinstance(m, nn.Linear)]
            for lin, (W, b) in zip(lins, self.arch.unpack(tr.W)):
                lin.weight.copy_(W)
                lin.bias.copy_(b)
        self.trained_with = "fused"

    def train(self, X, y):
        assert (X.dim() == 2)
        assert (y.dim() == 1)
        dev = self._dev()
        X, y = self._to_dev(X), self._to_dev(y)
        num_train = X.shape[0]
        if self.fused_train and dev.type == "cuda" and ops.dropout_train_supported(self.arch, self.batch_size):
            self.nn.to(dev)
            self._train_fused(X.contiguous().float(), y.contiguous().float())
            self.nn.eval()
            return
        self.trained_with = "torch"
        self.nn.to(dev).train()
        decay = [p for n, p in self.nn.named_parameters() if "bias" not in n]
        no_decay = [p for n, p in self.nn.named_parameters() if "bias" in n]
        opt = torch.optim.Adam([{'params': decay, 'weight_decay': self.l2_reg},
                                {'params': no_decay, 'weight_decay': 0.}], lr=self.lr)
        crit = nn.MSELoss(reduction='sum')
        gen = torch.Generator(device="cpu").manual_seed(self.seed)
        for epoch in range(self.num_epochs):
            epoch_loss = 0.
            perm = torch.randperm(num_train, generator=gen).to(dev)
            for i in range(0, num_train, self.batch_size):
                idx = perm[i: i + self.batch_size]
                opt.zero_grad()
                loss = crit(self.nn(X[idx]).squeeze(-1), y[idx])
                loss.backward()
                opt.step()
                epoch_loss += float(loss)
            self.noise_level = max(self.min_noise, np.sqrt(epoch_loss / num_train))
            if (epoch + 1) % self.print_every == 0:
                print("[Epoch %5d, loss = %g, noise_level = %g]" % (epoch + 1, epoch_loss / num_train, self.noise_level))
        self.nn.eval()

    def mask_config(self, kind="bernoulli"):
        return ops.MaskConfig(kind, self.layer_masked, p_drop=[self.dropout_rate] * self.arch.n_linear, seed=self.seed,
                              sample_offset=self._draws)

    def sample(self, num_samples=1, masks=None):
        """masks (optional, [S, mask_len]) injects the reference's own draws for exact parity."""
        dev = self._dev()
        base = self._cached_pack("base", list(self.nn.nn.parameters()), lambda: self.arch.pack([self.nn.nn])[0].to(dev))
        if masks is None:
            masks = ops.dropout_masks(self.arch, self.mask_config(), num_samples, dev)
            self._draws += num_samples
        return SampleBank(self.arch, base, masks=self._to_dev(masks), layer_masked=self.layer_masked)

    def sample_predict(self, nns, X):
        return self._sample_predict(nns, X, squeeze=True)

    def report(self):
        print(self.nn)
        print('Dropout rate: %g' % self.dropout_rate)
        print('Noise level : %g' % self.noise_level)
