"""GPU parity: the CUDA path (through the reference-shaped Python API -> ctypes -> C ABI) against golden
vectors recorded from the unmodified reference and against the fp64 oracle on the same seeded inputs.

Tolerance (BASELINE.json north_star): fp32 path rel <= 1e-5 on losses and gradients, where rel for a
tensor is max|x - ref| / max|ref|."""
import importlib

import numpy as np
import pytest
import torch

from oracle import damsm_oracle as O
from golden_util import Golden, case_names

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("t2i_clip-gan_b200")

TOL = 1e-5


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def pretrain_layout(x, dev="cuda", dtype=torch.float32):
    """Tensors laid out exactly like pretrain_DAMSM.py:125,130: regions are a permuted view of the
    CLS-sliced (B, R+1, D) ViT output, words a permuted view of contiguous (B, T, D)."""
    B, R, D = x["regions"].shape
    full = torch.zeros(B, R + 1, D, device=dev, dtype=dtype)
    full[:, 1:, :] = torch.tensor(x["regions"], device=dev).to(dtype)
    full.requires_grad_(True)
    regions = full[:, 1:, :].permute(0, 2, 1)
    words_s = torch.tensor(x["words"], device=dev).to(dtype).requires_grad_(True)
    words = words_s.permute(0, 2, 1)
    return full, regions, words_s, words


def run_words(x, gammas, **kw):
    full, regions, words_s, words = pretrain_layout(x)
    B = words.shape[0]
    labels = torch.arange(B, device="cuda")
    mask = torch.tensor(x["mask"])                      # CPU int64, as at pretrain_DAMSM.py:110
    l0, l1, attn = pkg.words_loss(regions, words, labels, torch.tensor(x["cap_len"]), x["class_ids"], B, mask,
                                  *gammas, **kw)
    (l0 + l1).backward()
    return l0.item(), l1.item(), words_s.grad.cpu().numpy(), full.grad[:, 1:, :].cpu().numpy(), attn, full.grad


@pytest.mark.parametrize("name", case_names())
def test_words_loss_vs_golden(name):
    g = Golden(name)
    l0, l1, dw, dr, _, fg = run_words(g.x, g.gammas)
    assert abs(l0 - g.scalar("w_loss0")) <= TOL * max(1.0, abs(g.scalar("w_loss0")))
    assert abs(l1 - g.scalar("w_loss1")) <= TOL * max(1.0, abs(g.scalar("w_loss1")))
    assert g.rel_err("dwords", dw) <= TOL
    assert g.rel_err("dregions", dr) <= TOL
    assert float(fg[:, 0, :].abs().max()) == 0.0        # the CLS row is outside the view: no gradient


@pytest.mark.parametrize("name", case_names())
def test_words_loss_vs_oracle(name):
    g = Golden(name)
    x = g.x
    o = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], *g.gammas)
    l0, l1, dw, dr, _, _ = run_words(x, g.gammas)
    assert abs(l0 - o["loss0"]) <= TOL * max(1.0, abs(o["loss0"]))
    assert abs(l1 - o["loss1"]) <= TOL * max(1.0, abs(o["loss1"]))
    assert rel(dw, o["dwords"]) <= TOL
    assert rel(dr, o["dregions"]) <= TOL


@pytest.mark.parametrize("name", case_names())
def test_sent_loss_vs_golden(name):
    g = Golden(name)
    x = g.x
    img = torch.tensor(x["img"], device="cuda", requires_grad=True)
    txt = torch.tensor(x["sent"], device="cuda", requires_grad=True)
    l0, l1 = pkg.sent_loss(img, txt, torch.arange(g.B, device="cuda"), x["class_ids"], g.B, gamma3=g.gammas[2])
    (l0 + l1).backward()
    assert abs(l0.item() - g.scalar("s_loss0")) <= TOL * max(1.0, abs(g.scalar("s_loss0")))
    assert abs(l1.item() - g.scalar("s_loss1")) <= TOL * max(1.0, abs(g.scalar("s_loss1")))
    assert g.rel_err("dimg", img.grad.cpu().numpy()) <= TOL
    assert g.rel_err("dtxt", txt.grad.cpu().numpy()) <= TOL


@pytest.mark.parametrize("name", [n for n in case_names() if Golden(n).has("fa_wc")])
def test_func_attention_vs_golden(name):
    g = Golden(name)
    x = g.x
    q = torch.tensor(x["words"], device="cuda").permute(0, 2, 1).contiguous().requires_grad_(True)     # (B, D, T)
    c = torch.tensor(x["regions"], device="cuda").permute(0, 2, 1).contiguous().requires_grad_(True)   # (B, D, R)
    m = torch.tensor(x["mask"], device="cuda").unsqueeze(1)
    wc, attn = pkg.func_attention(q, c, g.gammas[0], m)
    rng = np.random.default_rng(int(g.z["meta"][4]) + 1000)
    dwc = torch.tensor(rng.standard_normal((g.B, g.T, g.D)).astype(np.float32), device="cuda")
    wc.backward(dwc)
    h = int(np.sqrt(g.R))
    assert tuple(wc.shape) == (g.B, g.T, g.D) and tuple(attn.shape) == (g.B, g.T, h, h)
    assert g.rel_err("fa_wc", wc.detach().cpu().numpy()) <= TOL
    assert g.rel_err("fa_attn", attn.detach().cpu().numpy()) <= TOL
    assert g.rel_err("fa_dquery", q.grad.permute(0, 2, 1).cpu().numpy()) <= TOL
    assert g.rel_err("fa_dcontext", c.grad.permute(0, 2, 1).cpu().numpy()) <= TOL


def test_func_attention_attn_gradient_vs_oracle():
    """Gradient flowing into the second output (attn), which the reference's autograd also supports."""
    x = O.make_inputs(5, 7, 16, seed=77, class_ids=False)
    q = torch.tensor(x["words"], device="cuda").permute(0, 2, 1).contiguous().requires_grad_(True)
    c = torch.tensor(x["regions"], device="cuda").permute(0, 2, 1).contiguous().requires_grad_(True)
    wc, attn = pkg.func_attention(q, c, 4.0, torch.tensor(x["mask"], device="cuda").unsqueeze(1))
    rng = np.random.default_rng(5)
    dat = torch.tensor(rng.standard_normal((5, 7, 4, 4)).astype(np.float32), device="cuda")
    attn.backward(dat)
    # plain torch fp64 reference of the same op on the CPU
    qq = torch.tensor(x["words"], dtype=torch.float64, requires_grad=True)
    cc = torch.tensor(x["regions"], dtype=torch.float64, requires_grad=True)
    qn = qq / (qq.norm(dim=2, keepdim=True) + 1e-8)
    cn = cc / (cc.norm(dim=2, keepdim=True) + 1e-8)
    s = torch.einsum("btd,brd->btr", qn, cn).masked_fill(torch.tensor(x["mask"]).unsqueeze(2) == 0, float("-inf"))
    p = torch.softmax(s, dim=1)
    p.backward(dat.cpu().double().reshape(5, 7, 16))
    assert rel(q.grad.permute(0, 2, 1).cpu().numpy(), qq.grad.numpy()) <= TOL
    assert rel(c.grad.permute(0, 2, 1).cpu().numpy(), cc.grad.numpy()) <= TOL


def test_attn_maps_entry_matches_reference():
    g = Golden("tiny_b6_t5_r9_cls")
    _, _, _, _, attn, _ = run_words(g.x, g.gammas)
    assert len(attn) == g.B
    a0 = attn[0]
    assert tuple(a0.shape) == (g.B, g.R, g.T)           # losses.py:143-144,249
    assert g.rel_err("attn0", a0.cpu().numpy()) <= TOL


def test_labels_none_returns_none_losses():
    g = Golden("tiny_b6_t5_r9_cls")
    _, regions, _, words = pretrain_layout(g.x)
    l0, l1, attn = pkg.words_loss(regions, words, None, None, None, g.B, torch.tensor(g.x["mask"]), *g.gammas)
    assert l0 is None and l1 is None and len(attn) == g.B


def test_dmgan_call_shape_image_side_gradients_only():
    """generator_loss call site (losses.py:350-354, trainer.py:338): 6 positional args, 4-D contiguous
    (B, 512, 7, 7) regions, words detached; gammas come from cfg (4/5/10)."""
    g = Golden("c3_dmgan_b10_t77_r49")
    x = g.x
    B = g.B
    reg = torch.tensor(x["regions"], device="cuda").permute(0, 2, 1).reshape(B, 512, 7, 7).contiguous().requires_grad_(True)
    words = torch.tensor(x["words"], device="cuda").permute(0, 2, 1).contiguous()      # no grad (detached)
    cap_lens = torch.tensor(x["cap_len"], device="cuda")
    l0, l1, _ = pkg.words_loss(reg, words, torch.arange(B, device="cuda"), cap_lens, x["class_ids"], B)
    ((l0 + l1) * 10.0).backward()                       # * cfg.TRAIN.SMOOTH.LAMBDA (losses.py:355)
    assert abs(l0.item() - g.scalar("w_loss0")) <= TOL * max(1.0, abs(g.scalar("w_loss0")))
    dr = reg.grad.reshape(B, 512, 49).permute(0, 2, 1).cpu().numpy() / 10.0
    assert g.rel_err("dregions", dr) <= TOL


def test_general_labels_and_loss_weights():
    """Arbitrary match_labels and different upstream gradients for loss0 / loss1."""
    x = O.make_inputs(7, 6, 9, seed=31, class_ids=False)
    perm = np.array([2, 0, 1, 4, 3, 6, 5])
    o = O.words_loss(x["words"], x["regions"], x["mask"], perm, None, 4.0, 5.0, 10.0, g0=0.3, g1=1.7)
    _, regions, words_s, words = pretrain_layout(x)
    l0, l1, _ = pkg.words_loss(regions, words, torch.tensor(perm, device="cuda"), None, None, 7,
                               torch.tensor(x["mask"]), 4.0, 5.0, 10.0)
    (0.3 * l0 + 1.7 * l1).backward()
    assert abs(l0.item() - o["loss0"]) <= TOL * max(1, abs(o["loss0"]))
    assert abs(l1.item() - o["loss1"]) <= TOL * max(1, abs(o["loss1"]))
    assert rel(words_s.grad.cpu().numpy(), o["dwords"]) <= TOL


def test_non_prefix_mask():
    x = O.make_inputs(6, 9, 16, seed=5, class_ids=False)
    rng = np.random.default_rng(3)
    m = (rng.random((6, 9)) > 0.4).astype(np.int64)
    m[:, 0] = 1
    x["mask"] = m
    o = O.words_loss(x["words"], x["regions"], m, x["labels"], None, 4.0, 5.0, 10.0)
    l0, l1, dw, dr, _, _ = run_words(x, (4.0, 5.0, 10.0))
    assert abs(l0 - o["loss0"]) <= TOL and rel(dw, o["dwords"]) <= TOL and rel(dr, o["dregions"]) <= TOL


def test_bf16_inputs_through_exact_path():
    """bf16 tensors are accepted in place (dtype-templated prologue); compare with the oracle fed the same
    rounded values.  Gradients come back in bf16, hence the looser bound (bf16 has 8 mantissa bits)."""
    x = O.make_inputs(8, 12, 16, seed=9, class_ids=False)
    xb = dict(x)
    for k in ("words", "regions"):
        xb[k] = torch.tensor(x[k]).bfloat16().float().numpy()
    o = O.words_loss(xb["words"], xb["regions"], x["mask"], x["labels"], None, 4.0, 5.0, 10.0)
    words_s = torch.tensor(x["words"], device="cuda").bfloat16().requires_grad_(True)
    reg_s = torch.tensor(x["regions"], device="cuda").bfloat16().requires_grad_(True)
    l0, l1, _ = pkg.words_loss(reg_s.permute(0, 2, 1), words_s.permute(0, 2, 1), torch.arange(8, device="cuda"),
                               None, None, 8, torch.tensor(x["mask"]), 4.0, 5.0, 10.0)
    (l0 + l1).backward()
    assert abs(l0.item() - o["loss0"]) <= TOL * max(1, abs(o["loss0"]))
    assert words_s.grad.dtype == torch.bfloat16
    assert rel(words_s.grad.float().cpu().numpy(), o["dwords"]) <= 1e-2
    assert rel(reg_s.grad.float().cpu().numpy(), o["dregions"]) <= 1e-2


def test_full_size_properties_c2():
    """BASELINE configs[1] at full size (B48 T18 R49): properties that need no oracle --
    (i) permuting the batch permutes nothing in the loss; (ii) the loss gradient sums: each image's
    region gradient is orthogonal to its own (normalised) direction because l2norm is scale invariant."""
    x = O.make_inputs(48, 18, 49, seed=2027, class_ids=False)
    l0, l1, dw, dr, _, _ = run_words(x, (4.0, 5.0, 10.0))
    perm = np.random.default_rng(0).permutation(48)
    xp = dict(x)
    for k in ("words", "regions", "mask", "cap_len"):
        xp[k] = x[k][perm]
    p0, p1, dwp, drp, _, _ = run_words(xp, (4.0, 5.0, 10.0))
    assert abs(p0 - l0) <= 1e-5 and abs(p1 - l1) <= 1e-5
    assert rel(dwp, dw[perm]) <= 1e-4
    radial = (dr * x["regions"]).sum(-1)                # d/ds of loss(s * v) at s=1 must vanish
    assert np.abs(radial).max() <= 1e-5 * np.abs(dr).max() * np.abs(x["regions"]).max() * 512 ** 0.5


# ----------------------------------------------------------------------------------------------- NT-Xent
@pytest.mark.parametrize("name", case_names())
def test_nt_xent_vs_golden(name):
    g = Golden(name)
    x = g.x
    zi = torch.tensor(x["sent"], device="cuda", requires_grad=True)
    zj = torch.tensor(x["img"], device="cuda", requires_grad=True)
    loss = pkg.nt_xent(zi, zj, g.scalar("ntx_temperature"))
    loss.backward()
    assert loss.dim() == 0
    assert abs(loss.item() - g.scalar("ntx_loss")) <= TOL * max(1.0, abs(g.scalar("ntx_loss")))
    assert g.rel_err("ntx_dzi", zi.grad.cpu().numpy()) <= TOL
    assert g.rel_err("ntx_dzj", zj.grad.cpu().numpy()) <= TOL


@pytest.mark.parametrize("B,D,temp", [(2, 8, 0.5), (3, 33, 0.07), (48, 512, 0.5), (300, 512, 0.5)])
def test_nt_xent_vs_oracle_shapes(B, D, temp):
    """Edge sizes: two pairs, odd D, the pretrain batch, a batch wider than one CTA pass."""
    rng = np.random.default_rng(B * 1000 + D)
    s = rng.standard_normal((B, D))
    zi = (0.5 * s + rng.standard_normal((B, D))).astype(np.float32)
    zj = (0.5 * s + rng.standard_normal((B, D))).astype(np.float32)
    o = O.nt_xent(zi, zj, temp, g=0.2)                      # trainer.py:426 scales the term by 0.2
    a = torch.tensor(zi, device="cuda", requires_grad=True)
    b = torch.tensor(zj, device="cuda", requires_grad=True)
    loss = pkg.nt_xent(a, b, temp)
    (loss * 0.2).backward()
    assert abs(loss.item() - o["loss"]) <= TOL * max(1.0, abs(o["loss"]))
    assert rel(a.grad.cpu().numpy(), o["dz_i"]) <= TOL and rel(b.grad.cpu().numpy(), o["dz_j"]) <= TOL


def test_nt_xent_single_pair_is_zero():
    """B = 1: the only candidate of each row is its positive, so the loss and its gradient vanish."""
    a = torch.randn(1, 16, device="cuda", requires_grad=True)
    b = torch.randn(1, 16, device="cuda", requires_grad=True)
    loss = pkg.nt_xent(a, b, 0.5)
    loss.backward()
    assert abs(loss.item()) <= 1e-6 and a.grad.abs().max().item() <= 1e-6 and b.grad.abs().max().item() <= 1e-6


def test_nt_xent_module_drop_in():
    """``from nt_xent import NT_Xent`` / ``from masks import mask_correlated_samples_2`` with the reference's call
    shape (pretrain_DAMSM.py:445-449, :173); a strided view and a one-sided gradient."""
    import os
    import sys
    pdir = os.path.dirname(os.path.abspath(pkg.__file__))
    sys.path.insert(0, pdir)
    try:
        for m in ("nt_xent", "masks"):
            sys.modules.pop(m, None)
        from masks import mask_correlated_samples, mask_correlated_samples_2
        from nt_xent import NT_Xent
    finally:
        sys.path.remove(pdir)
    B, D = 12, 64
    mask = mask_correlated_samples_2(B)
    ref_mask = torch.ones((2 * B, 2 * B), dtype=torch.bool).fill_diagonal_(False)
    for i in range(B):
        ref_mask[i, B + i] = False
        ref_mask[B + i, i] = False
    assert torch.equal(mask, ref_mask)

    class Args:
        batch_size = B
    assert torch.equal(mask_correlated_samples(Args()), ref_mask)
    crit = NT_Xent(B, 0.5, mask, torch.device("cuda"))
    rng = np.random.default_rng(5)
    wide = torch.tensor(rng.standard_normal((B, 2 * D)).astype(np.float32), device="cuda")
    zi = wide[:, :D].requires_grad_(True)                  # row stride 2*D
    zj = torch.tensor(rng.standard_normal((B, D)).astype(np.float32), device="cuda")
    loss = crit(zi, zj)
    loss.backward()
    o = O.nt_xent(wide[:, :D].detach().cpu().numpy(), zj.cpu().numpy(), 0.5)
    assert abs(loss.item() - o["loss"]) <= TOL * max(1.0, abs(o["loss"]))
    assert rel(zi.grad.cpu().numpy(), o["dz_i"]) <= TOL and zj.grad is None
    bad = mask.clone()
    bad[0, 1] = False
    with pytest.raises(ValueError):
        NT_Xent(B, 0.5, bad, torch.device("cuda"))
    with pytest.raises(ValueError):
        crit(zi[:5], zj[:5])
