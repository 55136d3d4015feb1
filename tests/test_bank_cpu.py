"""SampleBank (what sample() returns) behaves like the reference's Python list of S networks: len, indexing, slicing,
iteration, callable elements (BO.py:32,82 calls nns[i](x)).  Pure host logic, no GPU."""
import numpy as np
import pytest
import torch

from helpers import random_nets
from oracle import bnn_oracle as O
from bnn_b200.bank import SampleBank
from bnn_b200.layout import Arch


def test_plain_bank_is_a_sequence_of_callables():
    dims, S = [5, 7, 3, 2], 6
    arch = Arch(dims, "tanh")
    nets = random_nets(dims, S, seed=3)
    bank = SampleBank(arch, arch.pack(nets))
    x = torch.randn(4, 5, generator=torch.Generator().manual_seed(1))
    assert len(bank) == S
    for i in (0, 3, -1):
        assert torch.allclose(bank[i](x), O.mlp_forward(nets[i], x, "tanh"), atol=1e-6)
    assert len(list(bank)) == S
    sub = bank[1:5:2]
    assert isinstance(sub, SampleBank) and len(sub) == 2
    assert torch.allclose(sub[1](x), O.mlp_forward(nets[3], x, "tanh"), atol=1e-6)
    pick = bank[np.array([4, 0])]
    assert len(pick) == 2 and torch.allclose(pick[0](x), O.mlp_forward(nets[4], x, "tanh"), atol=1e-6)
    with pytest.raises(IndexError):
        bank[S]
    mods = bank.to_modules()
    assert len(mods) == S and isinstance(mods[0], torch.nn.Sequential)
    again = SampleBank.from_modules(arch, mods, "cpu")
    assert torch.equal(again.weights, bank.weights)


def test_masked_bank_folds_masks_like_the_reference():
    """Dropout family: one base network + [S, sum(in)] keep masks; element i must be the reference's deep copy with its
    weight COLUMNS scaled (BNN_Dropout.py:70-75 / BNN_CDropout.py:59-65)."""
    dims, S = [4, 6, 5, 1], 5
    arch = Arch(dims, "relu")
    base = random_nets(dims, 1, seed=9)[0]
    g = torch.Generator().manual_seed(2)
    layer_masked = [True, True, True]
    masks = (torch.rand(S, sum(dims[:-1]), generator=g) > 0.3).float()
    bank = SampleBank(arch, arch.pack([base])[0], masks=masks, layer_masked=layer_masked)
    x = torch.randn(3, 4, generator=g)
    assert len(bank) == S
    for i in range(S):
        ms = [masks[i, 0:4], masks[i, 4:10], masks[i, 10:15]]
        ref = O.mlp_forward(O.scale_columns(base, ms, skip_unit_inputs=False), x, "relu")
        assert torch.allclose(bank[i](x), ref, atol=1e-6)
    assert len(bank[2:]) == S - 2
    with pytest.raises(ValueError):
        SampleBank(arch, arch.pack([base, base]), masks=masks, layer_masked=layer_masked)
