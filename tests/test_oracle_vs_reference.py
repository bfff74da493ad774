"""Live comparison against the reference source (only where /root/reference is mounted)."""
import numpy as np
import pytest

from oracle import damsm_oracle as O
from oracle import ref_shim as RS

pytestmark = pytest.mark.skipif(not RS.available(), reason="/root/reference not present (GPU box)")


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("B,T,R,cls,seed", [(5, 7, 9, True, 101), (8, 18, 49, False, 102), (4, 30, 16, True, 103)])
def test_live_reference(B, T, R, cls, seed):
    x = O.make_inputs(B, T, R, seed=seed, class_ids=cls, n_classes=3)
    ref = RS.ref_words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], (4, 5, 10))
    o = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], 4.0, 5.0, 10.0)
    assert abs(o["loss0"] - ref["loss0"]) < 2e-6 * max(1, abs(ref["loss0"]))
    assert abs(o["loss1"] - ref["loss1"]) < 2e-6 * max(1, abs(ref["loss1"]))
    assert rel(o["dwords"], ref["dwords"]) < 5e-6
    assert rel(o["dregions"], ref["dregions"]) < 5e-6
    rs = RS.ref_sent_loss(x["img"], x["sent"], x["labels"], x["class_ids"], 10.0)
    os_ = O.sent_loss(x["img"], x["sent"], x["labels"], x["class_ids"], 10.0)
    assert abs(os_["loss0"] - rs["loss0"]) < 2e-6 * max(1, abs(rs["loss0"]))
    assert rel(os_["dimg"], rs["dimg"]) < 5e-6 and rel(os_["dtxt"], rs["dtxt"]) < 5e-6


def test_general_mask_not_prefix():
    """words_mask is a general 0/1 mask, not only a prefix mask."""
    x = O.make_inputs(5, 9, 16, seed=7, class_ids=False)
    rng = np.random.default_rng(3)
    m = (rng.random((5, 9)) > 0.4).astype(np.int64)
    m[:, 0] = 1
    ref = RS.ref_words_loss(x["words"], x["regions"], m, x["labels"], None, (4, 5, 10))
    o = O.words_loss(x["words"], x["regions"], m, x["labels"], None, 4.0, 5.0, 10.0)
    assert abs(o["loss0"] - ref["loss0"]) < 2e-6 and rel(o["dwords"], ref["dwords"]) < 5e-6


@pytest.mark.parametrize("B,temp,seed", [(2, 0.5, 1), (7, 0.5, 2), (24, 0.1, 3)])
def test_live_nt_xent(B, temp, seed):
    rng = np.random.default_rng(seed)
    s = rng.standard_normal((B, 1, 64))
    zi = (0.5 * s[:, 0] + rng.standard_normal((B, 64))).astype(np.float32)
    zj = (0.5 * s[:, 0] + rng.standard_normal((B, 64))).astype(np.float32)
    ref = RS.ref_nt_xent(zi, zj, temp)
    o = O.nt_xent(zi, zj, temp)
    assert abs(o["loss"] - ref["loss"]) < 2e-6 * max(1, abs(ref["loss"]))
    assert rel(o["dz_i"], ref["dz_i"]) < 5e-6 and rel(o["dz_j"], ref["dz_j"]) < 5e-6
