"""Skip-padded-words (tensor-core path): the pair kernels compute only the words t < nw[i]; every later word is padding
and enters through the closed form of csrc/pad_terms.cu.  The reference SUMS padded words into the score and gives them
gradient (losses.py:127,173-174,198-203), so parity is unchanged: losses and gradients -- including the gradient rows
of the padded words -- against the fp64 oracle, tolerance 2e-3 (bf16-input path)."""
import importlib

import numpy as np
import pytest
import torch

from oracle import damsm_oracle as O

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("t2i_clip-gan_b200")
TOL = 2e-3
GAM = (4.0, 5.0, 10.0)


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run(x, mask, cls=None, weights=(1.0, 1.0)):
    B = x["words"].shape[0]
    wv = torch.tensor(x["words"]).bfloat16().float()
    rv = torch.tensor(x["regions"]).bfloat16().float()
    o = O.words_loss(wv.numpy(), rv.numpy(), mask, np.arange(B), cls, *GAM, g0=weights[0], g1=weights[1])
    w = wv.cuda().requires_grad_(True)
    r = rv.cuda().requires_grad_(True)
    l0, l1, _ = pkg.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), torch.arange(B, device="cuda"), None, cls, B,
                               torch.tensor(mask), *GAM, precision="bf16")
    (weights[0] * l0 + weights[1] * l1).backward()
    return o, l0.item(), l1.item(), w.grad.cpu().numpy(), r.grad.cpu().numpy()


def check(o, l0, l1, gw, gr, mask=None):
    assert abs(l0 - o["loss0"]) <= TOL * max(1, abs(o["loss0"])), (l0, o["loss0"])
    assert abs(l1 - o["loss1"]) <= TOL * max(1, abs(o["loss1"])), (l1, o["loss1"])
    assert rel(gw, o["dwords"]) <= TOL, rel(gw, o["dwords"])
    assert rel(gr, o["dregions"]) <= TOL, rel(gr, o["dregions"])
    if mask is not None and (mask == 0).any():
        # the padded words' own gradient rows (what the closed form produces), relative to THEIR magnitude
        pad = mask == 0
        ref = o["dwords"][pad]
        assert np.abs(ref).max() > 0
        assert rel(gw[pad], ref) <= 5 * TOL, rel(gw[pad], ref)


@pytest.mark.parametrize("B,T,R,lens,seed", [
    (12, 77, 196, "short", 1),        # captions of 3..20 words: nw = 16 or 32 of 80 columns, 45+ skipped words each
    (10, 77, 49, "mixed", 2),         # lengths all over [2, 77]: every nw bucket, some captions with nothing skipped
    (8, 77, 196, "full", 3),          # no padding at all: the closed form contributes nothing
    (9, 40, 120, "mixed", 4),         # NT = 64
    (6, 100, 49, "mixed", 5),         # NT = 128
    (16, 28, 49, "short", 6),         # the reference's real T (words_num = 30 minus SOS/EOS), NT = 32
    (5, 18, 49, "mixed", 7),          # T = 18: nw in {16, 32}; words 16, 17 of a short caption are skipped
])
def test_skip_padded_losses_and_gradients(B, T, R, lens, seed):
    x = O.make_inputs(B, T, R, seed=seed, class_ids=True, n_classes=3)
    g = np.random.default_rng(seed)
    if lens == "short":
        L = g.integers(3, min(T, 20) + 1, size=B)
    elif lens == "full":
        L = np.full(B, T)
    else:
        L = g.integers(2, T + 1, size=B)
        L[0], L[-1] = T, 2
    mask = (np.arange(T)[None, :] < L[:, None]).astype(np.int64)
    out = run(x, mask, x["class_ids"], weights=(1.0, 0.5))
    check(*out, mask=mask)


def test_skip_padded_general_mask_with_holes():
    """Not a prefix mask: masked words BEFORE the last unmasked one stay inside the pair kernel (nw covers the last
    unmasked word), the ones after it go through the closed form."""
    B, T, R = 8, 77, 49
    x = O.make_inputs(B, T, R, seed=11, class_ids=False)
    g = np.random.default_rng(11)
    mask = (g.random((B, T)) < 0.4).astype(np.int64)
    mask[:, 0] = 1
    mask[0, :] = 0
    mask[0, :5] = 1                    # ends early: 72 skipped words
    mask[1, -1] = 1                    # last word valid: nothing skipped, many holes
    out = run(x, mask)
    check(*out, mask=mask)


def test_skip_padded_multi_chunk_backward_and_one_sided():
    """Ragged scratch columns across several workspace chunks; captions sorted by length; image-side-only gradient."""
    B, T, R = 40, 77, 196
    x = O.make_inputs(B, T, R, seed=21, class_ids=True, n_classes=5)
    g = np.random.default_rng(21)
    L = g.integers(2, T + 1, size=B)
    mask = (np.arange(T)[None, :] < L[:, None]).astype(np.int64)
    eng = pkg.get_engine("bf16")
    lib = pkg._lib.load()
    old = eng.tc_workspace_bytes
    eng.tc_workspace_bytes = lib.damsm_words_bwd_tc_fixed_bytes() + 400 * lib.damsm_words_bwd_tc_col_bytes(B, R)
    try:
        out = run(x, mask, x["class_ids"])
        check(*out, mask=mask)
        # image side only (DM-GAN generator step): same dregions, no dwords
        wv = torch.tensor(x["words"]).bfloat16().float().cuda()
        rv = torch.tensor(x["regions"]).bfloat16().float().cuda().requires_grad_(True)
        l0, l1, _ = pkg.words_loss(rv.permute(0, 2, 1), wv.permute(0, 2, 1), torch.arange(B, device="cuda"), None,
                                   x["class_ids"], B, torch.tensor(mask), *GAM, precision="bf16")
        (l0 + l1).backward()
        assert rel(rv.grad.cpu().numpy(), out[0]["dregions"]) <= TOL
    finally:
        eng.tc_workspace_bytes = old


def test_plan_counts_and_order():
    eng = pkg.get_engine("bf16")
    T = 77
    L = torch.tensor([1, 16, 17, 32, 33, 64, 65, 77, 5, 48])
    mask = (torch.arange(T)[None, :] < L[:, None]).to(torch.uint8).cuda()
    plan = eng.words_plan(mask, T)
    torch.cuda.synchronize()
    nw = plan["nw"].cpu().numpy()
    assert nw.tolist() == [16, 16, 32, 32, 48, 64, 80, 80, 16, 48]
    assert plan["nw_host"].numpy().tolist() == nw.tolist()
    order = plan["order"].cpu().numpy()
    assert sorted(order.tolist()) == list(range(10))
    assert (np.diff(nw[order]) <= 0).all()                  # longest first
