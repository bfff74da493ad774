"""Sentence loss as one launch each way (csrc/sent_fused.cu; losses.py:51-91) against the fp64 oracle and the unfused path.
Parity bar 1e-5 (exact fp32)."""
import importlib

import numpy as np
import pytest
import torch

from oracle import damsm_oracle as O

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("t2i_clip-gan_b200")
ops = importlib.import_module("t2i_clip-gan_b200.ops")
TOL = 1e-5


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def inputs(b, d, seed, n_classes):
    rng = np.random.default_rng(seed)
    img = rng.standard_normal((b, d)).astype(np.float32)
    txt = (0.5 * img + rng.standard_normal((b, d))).astype(np.float32)
    cls = rng.integers(0, n_classes, size=b).astype(np.int64) if n_classes else None
    return img, txt, cls


@pytest.mark.parametrize("b,d,n_classes,seed", [(48, 512, 200, 1), (10, 512, 0, 2), (130, 100, 5, 3), (65, 33, 3, 4),
                                                (1, 16, 0, 5), (300, 256, 40, 6)])
def test_fused_sentence_loss_vs_oracle(b, d, n_classes, seed, monkeypatch):
    monkeypatch.setattr(ops, "SENT_FUSED_MAX_PAIRS", 1 << 40)      # the one-launch kernels at every size of this test
    img, txt, cls = inputs(b, d, seed, n_classes)
    labels = np.arange(b)
    o = O.sent_loss(img, txt, labels, cls, 10.0, g0=1.0, g1=0.7)
    ti = torch.tensor(img, device="cuda", requires_grad=True)
    tt = torch.tensor(txt, device="cuda", requires_grad=True)
    eng = pkg.get_engine("fp32")
    assert eng.sent_fused_ok(10.0) and not eng.sent_fused_ok(61.0)
    pkg._lib.reset_launch_count()
    l0, l1 = pkg.sent_loss(ti, tt, torch.arange(b, device="cuda"), cls, b, gamma3=10.0)
    (l0 + 0.7 * l1).backward()
    assert pkg._lib.launch_count() == 3               # fused forward, the two CE scalars, fused backward
    assert abs(l0.item() - o["loss0"]) <= TOL * max(1.0, abs(o["loss0"]))
    assert abs(l1.item() - o["loss1"]) <= TOL * max(1.0, abs(o["loss1"]))
    assert rel(ti.grad.cpu().numpy(), o["dimg"]) <= TOL
    assert rel(tt.grad.cpu().numpy(), o["dtxt"]) <= TOL


def test_fused_matches_unfused_block_and_zero_norm():
    """Engine level: a (rows x all columns) block with a row offset (the multi-GPU shard form), a zero-norm row."""
    b, d, br, off = 96, 64, 32, 32
    img, txt, cls = inputs(b, d, 7, 6)
    img[off + 3] = 0.0
    eng = pkg.get_engine("fp32")
    a = torch.tensor(img[off:off + br], device="cuda")
    t = torch.tensor(txt, device="cuda")
    c = torch.tensor(cls, device="cuda")
    lo, na, nb, rl, cm, cs = eng.sent_fwd(a, t, c[off:off + br], c, off, 10.0, 1e-8)
    lo2, na2, nb2 = eng.cos_logits(a, t, 10.0, 1e-8)
    rl2, cm2, cs2 = eng.ce_stats(lo2, c[off:off + br], c, off)
    assert torch.equal(torch.isinf(lo), torch.isinf(lo2))
    fin = ~torch.isinf(lo)
    assert (lo[fin] - lo2[fin]).abs().max().item() <= 1e-5
    assert (rl - rl2).abs().max().item() <= 1e-5
    col = torch.log(cs) + cm
    col2 = torch.log(cs2) + cm2
    assert (col - col2).abs().max().item() <= 1e-5
    gs = torch.tensor([1.0, 0.5], device="cuda")
    lab = torch.arange(b, device="cuda")
    da, db = eng.sent_bwd(a, t, na, nb, lo, rl, col, lab, gs, off, b, 10.0, 1e-8)
    da2, db2 = eng.cos_logits_bwd(a, t, na2, nb2, lo2, rl2, col2, lab, gs, off, b, 10.0, 1e-8)
    assert rel(da.cpu().numpy(), da2.cpu().numpy()) <= 1e-5
    assert rel(db.cpu().numpy(), db2.cpu().numpy()) <= 1e-5
    assert torch.isfinite(da).all() and torch.isfinite(db).all()


def test_size_dispatch():
    """Up to SENT_FUSED_MAX_PAIRS logits the loss is 3 launches; larger batches take the tiled unfused kernels."""
    for b, fused in ((48, True), (128, True), (160, False)):
        img, txt, _ = inputs(b, 64, 9, 0)
        ti = torch.tensor(img, device="cuda", requires_grad=True)
        tt = torch.tensor(txt, device="cuda", requires_grad=True)
        pkg._lib.reset_launch_count()
        l0, l1 = pkg.sent_loss(ti, tt, torch.arange(b, device="cuda"), None, b)
        (l0 + l1).backward()
        assert (pkg._lib.launch_count() == 3) == fused
        o = O.sent_loss(img, txt, np.arange(b), None, 10.0)
        assert abs(l0.item() - o["loss0"]) <= TOL * max(1.0, abs(o["loss0"]))
        assert rel(ti.grad.cpu().numpy(), o["dimg"]) <= TOL


def test_large_gamma3_uses_the_unfused_path():
    img, txt, cls = inputs(20, 32, 8, 0)
    o = O.sent_loss(img, txt, np.arange(20), None, 80.0)
    l0, l1 = pkg.sent_loss(torch.tensor(img, device="cuda"), torch.tensor(txt, device="cuda"),
                           torch.arange(20, device="cuda"), None, 20, gamma3=80.0)
    assert abs(l0.item() - o["loss0"]) <= 1e-4 * max(1.0, abs(o["loss0"]))
    assert abs(l1.item() - o["loss1"]) <= 1e-4 * max(1.0, abs(o["loss1"]))
