"""Drop-in for the helpers of the reference's util.py that sit on or next to the hot path."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class NN(nn.Module):
    """MLP builder with the reference's structure and init (util.py:8-33): Linear/act pairs, a final
    Linear, xavier-uniform weights, torch-default biases."""

    def __init__(self, dim, act=nn.ReLU(), num_hiddens=[50], nout=1):
        super().__init__()
        self.dim, self.nout, self.act = dim, nout, act
        self.num_hiddens, self.num_layers = num_hiddens, len(num_hiddens)
        layers, pre = [], dim
        for h in num_hiddens:
            layers += [nn.Linear(pre, h, bias=True), act]
            pre = h
        layers.append(nn.Linear(pre, nout, bias=True))
        self.nn = nn.Sequential(*layers)
        for l in self.nn:
            if type(l) == nn.Linear:
                nn.init.xavier_uniform_(l.weight)

    def forward(self, x):
        return self.nn(x)


def stable_noise_var(input):
    """util.py:52-53."""
    return F.softplus(torch.clamp(input, min=-35.))


def normalize(X, y):
    """util.py:63-73: per-feature mean/std of X, mean/std of y, zero stds replaced by 1."""
    assert X.dim() == 2
    assert y.dim() == 1
    x_mean, x_std = X.mean(dim=0), X.std(dim=0)
    y_mean, y_std = y.mean(), y.std()
    x_std[x_std == 0] = torch.tensor(1.)
    if y_std == 0:
        y_std = torch.tensor(1.)
    return x_mean, x_std, y_mean, y_std
