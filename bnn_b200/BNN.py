"""Drop-in for the reference's abstract base class (BNN.py:6-37).

Same methods, same argument meaning, same return shapes; ``predict_mv`` and ``validate`` run as
one fused CUDA launch (forward + Welford reduction) plus the metrics kernel instead of the
reference's Python loop over S networks.  Inputs may live on the CPU or on the GPU; results come
back on the device of the input, float32, like the reference's.
"""
from abc import ABC, abstractmethod

import torch

from . import ops
from .bank import SampleBank


class BNN(ABC):
    def __init__(self):
        pass

    @abstractmethod
    def train(self, X, y):
        """X: num_train x dim matrix, y: num_train vector (BNN.py:10-16)."""

    @abstractmethod
    def sample(self, num_samples=1):
        """num_samples posterior draws as a SampleBank (list-like of callables, BNN.py:18-23)."""

    # ---- plumbing shared by the subclasses --------------------------------------
    device = None          # set by subclasses (torch.device("cuda", k))
    precision = "auto"     # "auto" | "fp32" | "tf32_bf16" | "tf32x3" | "tf32" (see bnn_b200.ops.forward)

    def _cached_pack(self, name, params, build):
        """build() -> device tensor(s) derived from `params`; reused until one of them is modified in place, re-pointed
        (.data = ...) or replaced.  Keeps repeated sample() calls from re-packing / re-uploading unchanged networks."""
        key = (str(self._dev()),) + tuple((id(p), p._version, p.data_ptr()) for p in params)
        cache = self.__dict__.setdefault("_pack_cache", {})
        if name not in cache or cache[name][0] != key:
            cache[name] = (key, build())
        return cache[name][1]

    def _dev(self):
        if self.device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("bnn_b200 needs a CUDA device: there is no CPU fallback")
            self.device = torch.device("cuda", torch.cuda.current_device())
        return self.device

    def _to_dev(self, t):
        return torch.as_tensor(t, dtype=torch.float32).to(self._dev()).contiguous()

    def _as_bank(self, nns):
        """Accept what the reference accepts -- any sequence of sampled networks -- and our SampleBank."""
        if isinstance(nns, SampleBank):
            return nns
        return SampleBank.from_modules(self.arch, list(nns), self._dev())

    def _forward(self, nns, X, want_pred, want_moments, want_partials=False):
        bank = self._as_bank(nns)
        Xd = self._to_dev(X).reshape(-1, self.arch.dims[0])
        if bank.masks is not None:
            cfg = ops.MaskConfig("external", bank.layer_masked, mask=bank.masks)
            return ops.forward(self.arch, Xd, bank.weights, S=len(bank), mask=cfg, want_pred=want_pred,
                               want_moments=want_moments, want_partials=want_partials, precision=self.precision)
        return ops.forward(self.arch, Xd, bank.weights, want_pred=want_pred, want_moments=want_moments,
                           want_partials=want_partials, precision=self.precision)

    # ---- BNN.py:25-37 ----------------------------------------------------------------
    def validate(self, X, y, num_samples=20):
        nns = self.sample(num_samples)
        out = self._forward(nns, X, want_pred=False, want_moments=True)
        yd = self._to_dev(y).reshape(out["mean"].shape[0], -1)
        rmse, nll = ops.metrics(out["mean"], out["var"], yd)       # raises ValueError where Normal(scale<=0) would
        return rmse, nll

    def predict_mv(self, input, nn_samples):
        out = self._forward(nn_samples, input, want_pred=False, want_moments=True)
        back = input.device if torch.is_tensor(input) else torch.device("cpu")
        return out["mean"].to(back), out["var"].to(back)

    def _sample_predict(self, nns, X, squeeze):
        out = self._forward(nns, X, want_pred=True, want_moments=False)
        pred = out["pred"]
        if squeeze:                               # the non-SGDMC subclasses return [S, N] (nout == 1)
            pred = pred.reshape(pred.shape[0], pred.shape[1])
        back = X.device if torch.is_tensor(X) else torch.device("cpu")
        return pred.to(back)
