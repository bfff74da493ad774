"""The padded-word closed form (oracle/padded_closed_form.py, groundwork for skipping padded words in the pair kernels)
reproduces the oracle's full computation -- losses, masked score matrix and both gradients -- to fp64 round-off."""
import numpy as np
import pytest

from oracle import damsm_oracle as O
from oracle.padded_closed_form import words_loss_split


def rel(a, b):
    return float(np.abs(np.asarray(a) - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("B,T,R,cls,seed", [(5, 9, 16, True, 1), (6, 18, 49, False, 2), (4, 30, 9, True, 3)])
def test_split_equals_full(B, T, R, cls, seed):
    x = O.make_inputs(B, T, R, D=64, seed=seed, class_ids=cls, n_classes=3)
    assert (x["mask"] == 0).any() and (x["mask"].sum(1) >= 1).all()
    full = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], 4.0, 5.0, 10.0)
    split = words_loss_split(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], 4.0, 5.0, 10.0)
    assert abs(split["loss0"] - full["loss0"]) < 1e-10 and abs(split["loss1"] - full["loss1"]) < 1e-10
    fin = np.isfinite(full["sim"])
    assert np.array_equal(fin, np.isfinite(split["sim"])) and np.abs(split["sim"][fin] - full["sim"][fin]).max() < 1e-10
    assert rel(split["dwords"], full["dwords"]) < 1e-9 and rel(split["dregions"], full["dregions"]) < 1e-9


def test_split_with_general_mask():
    """Not only prefix masks: any word with mask 0 is 'padded' in the reference's sense."""
    x = O.make_inputs(4, 11, 16, D=32, seed=9, class_ids=False)
    m = (np.random.default_rng(4).random((4, 11)) > 0.45).astype(np.int64)
    m[:, 0] = 1
    full = O.words_loss(x["words"], x["regions"], m, x["labels"], None, 4.0, 5.0, 10.0)
    split = words_loss_split(x["words"], x["regions"], m, x["labels"], None, 4.0, 5.0, 10.0)
    assert abs(split["loss0"] - full["loss0"]) < 1e-10
    assert rel(split["dwords"], full["dwords"]) < 1e-9 and rel(split["dregions"], full["dregions"]) < 1e-9
