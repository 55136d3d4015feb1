"""SampleBank: what ``sample()`` returns instead of a Python list of S ``nn.Module`` copies.

The reference materialises S separate networks (``deepcopy`` / fresh ``nn.Linear``:
BNN_SGDMC.py:124,127-129, BNN_BBB.py:44-51,65-69, BNN_Dropout.py:67-76, BNN_CDropout.py:59-65),
which costs as much as the forward itself.  Here the S networks stay one device tensor
``[S, stride]`` (or one base network + ``[S, mask_len]`` keep masks for the dropout family) and
this class is a lazy sequence view over it: ``len()``, indexing, slicing, iteration and calling an
element (``nns[i](x)``, as BO.py:32,82 does) all work; an ``nn.Sequential`` with the reference's
structure is only built when an element is actually asked for.
"""
import numpy as np
import torch


class SampleBank:
    def __init__(self, arch, weights, masks=None, layer_masked=None):
        """weights: [S, stride] bank, or [stride] base network when ``masks`` [S, mask_len] is given."""
        self.arch = arch
        self.weights = weights
        self.masks = masks
        self.layer_masked = layer_masked
        if masks is not None and weights.dim() != 1:
            raise ValueError("masked banks share one base network")

    # ---- sequence protocol -------------------------------------------------
    def __len__(self):
        return int(self.masks.shape[0] if self.masks is not None else self.weights.shape[0])

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            if idx < 0:
                idx += len(self)
            if not 0 <= idx < len(self):
                raise IndexError(idx)
            return self.module(int(idx))
        if isinstance(idx, slice):
            idx = list(range(*idx.indices(len(self))))
        idx = torch.as_tensor(np.asarray(idx), dtype=torch.long, device=self.weights.device)
        if self.masks is not None:
            return SampleBank(self.arch, self.weights, self.masks[idx].contiguous(), self.layer_masked)
        return SampleBank(self.arch, self.weights[idx].contiguous())

    def __iter__(self):
        return (self.module(i) for i in range(len(self)))

    # ---- materialisation ---------------------------------------------------
    def packed(self, i):
        """Packed parameter vector of sample i with the mask folded into the weight columns
        (W[:, j] *= m_j -- exactly what the reference's sample() stores)."""
        if self.masks is None:
            return self.weights[i]
        vec = self.weights.clone()
        offs, _ = self.arch.mask_layout(self.layer_masked)
        for l, o in enumerate(offs):
            if o >= 0:
                v = self.arch.layer_view(vec, l)
                v[:, : self.arch.ins[l]] *= self.masks[i, o: o + self.arch.ins[l]]
        return vec

    def module(self, i):
        """nn.Sequential(Linear, act, ..., Linear) for sample i, on the CPU like the reference's."""
        net = self.arch.to_module(self.packed(i))
        for p in net.parameters():
            p.requires_grad_(False)
        return net

    def to_modules(self):
        return [self.module(i) for i in range(len(self))]

    @classmethod
    def from_modules(cls, arch, modules, device):
        return cls(arch, arch.pack(modules).to(device))
