"""Packed parameter layout of include/bnn_b200.h and converters to/from the
reference's representation (an ``nn.Sequential`` of ``nn.Linear``/activation,
util.py:21-29).

One network = one flat float32 vector of ``stride`` floats; Linear l occupies
``out_l`` rows of ``ld_l = roundup4(in_l + 1)`` floats: weight row, bias, zero pad
(bias-last, like BNN_BBB.py:19 stores mu/rho).  A bank of S networks is ``[S, stride]``.
"""
import torch
import torch.nn as nn

from . import _lib

_ACT_IDS = {"relu": _lib.ACT_RELU, "tanh": _lib.ACT_TANH, "sigmoid": _lib.ACT_SIGMOID, "identity": _lib.ACT_IDENTITY}
_ACT_MODULES = {nn.ReLU: "relu", nn.Tanh: "tanh", nn.Sigmoid: "sigmoid", nn.Identity: "identity"}


def act_name(act):
    """Name of an activation given as a string or as the nn.Module the reference passes (util.py:9)."""
    if isinstance(act, str):
        if act not in _ACT_IDS:
            raise NotImplementedError(f"activation {act!r} has no CUDA implementation (have {sorted(_ACT_IDS)})")
        return act
    for cls, name in _ACT_MODULES.items():
        if type(act) is cls:
            return name
    raise NotImplementedError(f"activation {act!r} has no CUDA implementation (have {sorted(_ACT_IDS)})")


def act_module(name):
    return {"relu": nn.ReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid, "identity": nn.Identity}[name]()


class Arch:
    """Shape of one MLP: ``dims = [dim, h1, ..., hL, nout]`` plus the activation."""

    def __init__(self, dims, act="relu"):
        dims = [int(d) for d in dims]
        if not (2 <= len(dims) <= _lib.MAX_LINEAR + 1):
            raise ValueError(f"need between 1 and {_lib.MAX_LINEAR} Linear layers, got dims={dims}")
        if min(dims) < 1 or max(dims) > _lib.MAX_WIDTH:
            raise ValueError(f"layer widths must be in [1, {_lib.MAX_WIDTH}], got dims={dims}")
        self.dims = dims
        self.act = act_name(act)
        self.n_linear = len(dims) - 1
        self.ins = dims[:-1]
        self.outs = dims[1:]
        self.lds = [(i + 1 + 3) // 4 * 4 for i in self.ins]
        self.offsets = []
        o = 0
        for i in range(self.n_linear):
            self.offsets.append(o)
            o += self.outs[i] * self.lds[i]
        self.stride = o
        self.nout = dims[-1]
        self.n_params = sum((i + 1) * o for i, o in zip(self.ins, self.outs))   # the reference's P
        self.flops_per_eval = 2 * sum(i * o for i, o in zip(self.ins, self.outs))

    # ---- C structs -------------------------------------------------------
    def desc(self):
        d = _lib.MlpDesc()
        d.n_linear = self.n_linear
        for i, v in enumerate(self.dims):
            d.dims[i] = v
        d.act = _ACT_IDS[self.act]
        return d

    # ---- masks -------------------------------------------------------------
    def mask_layout(self, layer_masked):
        """(offsets per layer or -1, total length) of the concatenated input masks."""
        offs, o = [], 0
        for l in range(self.n_linear):
            if layer_masked[l]:
                offs.append(o)
                o += self.ins[l]
            else:
                offs.append(-1)
        return offs, o

    # ---- pack / unpack -------------------------------------------------------
    def layer_view(self, bank, l):
        """[..., out_l, ld_l] strided view of Linear l inside a packed vector or bank."""
        lead = bank.shape[:-1]
        return bank[..., self.offsets[l]: self.offsets[l] + self.outs[l] * self.lds[l]].reshape(*lead, self.outs[l], self.lds[l])

    def pack(self, nets, device=None, out=None):
        """nets: sequence of networks, each an nn.Module containing nn.Linear layers in
        order, or a list of (W[out,in], b[out]) pairs.  Returns the bank [S, stride]."""
        nets = list(nets)
        S = len(nets)
        bank = out if out is not None else torch.zeros(S, self.stride, dtype=torch.float32)
        for s, net in enumerate(nets):
            pairs = _linear_pairs(net)
            if len(pairs) != self.n_linear:
                raise ValueError(f"network {s} has {len(pairs)} Linear layers, expected {self.n_linear}")
            for l, (W, b) in enumerate(pairs):
                if tuple(W.shape) != (self.outs[l], self.ins[l]):
                    raise ValueError(f"network {s} layer {l}: weight {tuple(W.shape)} != {(self.outs[l], self.ins[l])}")
                v = self.layer_view(bank[s], l)
                v[:, : self.ins[l]] = W.detach().to(torch.float32)
                v[:, self.ins[l]] = b.detach().to(torch.float32)
        return bank.to(device) if device is not None else bank

    def pack_flat(self, flat):
        """[S, n_params] in the reference's parameter order (W0, b0, W1, b1, ...) -> bank."""
        flat = torch.as_tensor(flat, dtype=torch.float32)
        S = flat.shape[0]
        bank = torch.zeros(S, self.stride, dtype=torch.float32)
        o = 0
        for l in range(self.n_linear):
            nw = self.outs[l] * self.ins[l]
            v = self.layer_view(bank, l)
            v[:, :, : self.ins[l]] = flat[:, o: o + nw].reshape(S, self.outs[l], self.ins[l])
            o += nw
            v[:, :, self.ins[l]] = flat[:, o: o + self.outs[l]]
            o += self.outs[l]
        return bank

    def pack_bias_last(self, mats):
        """mats[l]: [..., out_l, in_l + 1] (BNN_BBB's mu/rho layout) -> packed [..., stride]."""
        lead = mats[0].shape[:-2]
        out = torch.zeros(*lead, self.stride, dtype=torch.float32, device=mats[0].device)
        for l, m in enumerate(mats):
            self.layer_view(out, l)[..., : self.ins[l] + 1] = m.detach().to(torch.float32)
        return out

    def unpack(self, vec):
        """One packed vector -> [(W, b), ...] (copies, CPU or device as given)."""
        pairs = []
        for l in range(self.n_linear):
            v = self.layer_view(vec, l)
            pairs.append((v[:, : self.ins[l]].clone(), v[:, self.ins[l]].clone()))
        return pairs

    def to_module(self, vec):
        """One packed vector -> nn.Sequential with the reference's structure (util.py:21-29)."""
        layers = []
        for l, (W, b) in enumerate(self.unpack(vec.detach().cpu())):
            lin = nn.Linear(self.ins[l], self.outs[l], bias=True)
            lin.weight.data = W
            lin.bias.data = b
            layers.append(lin)
            if l != self.n_linear - 1:
                layers.append(act_module(self.act))
        return nn.Sequential(*layers)

    def valid_mask(self):
        """bool [stride]: True where a packed float is a real parameter (not padding)."""
        m = torch.zeros(self.stride, dtype=torch.bool)
        for l in range(self.n_linear):
            self.layer_view(m, l)[:, : self.ins[l] + 1] = True
        return m


def _linear_pairs(net):
    if isinstance(net, nn.Module):
        return [(m.weight, m.bias) for m in net.modules() if isinstance(m, nn.Linear)]
    return list(net)
