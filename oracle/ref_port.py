"""Torch-CPU port of the reference's *procedure* for the hot path.  TEST INFRASTRUCTURE ONLY.

``/root/reference`` does not travel to the GPU box, and the reference is pure
Python (nothing to compile into ``oracle/_ref``), so this port is what
``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg time on the
box's host cores (``cpu_baseline.kind == "port"``).  It keeps the reference's
execution structure -- one Python iteration per caption, each scoring that
caption against every image with batched fp32 torch ops, including the
reference's inline sanity checks (which are part of what the reference pays
for) -- so its timing is representative of the reference's own CPU path:

  words_loss loop            /root/reference/DMGAN+CLIP/code/miscc/losses.py:228-251
  per-caption scoring        .../miscc/losses.py:95-216
  class mask + two CEs       .../miscc/losses.py:224-232, 254-269
  sent_loss                  .../miscc/losses.py:51-91

``tests/test_oracle_vs_reference.py`` checks it against the live reference
(loss and autograd gradients) whenever ``/root/reference`` is present, and
``tests/test_oracle_golden.py`` against the recorded golden vectors everywhere.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _unit(x, dim):
    # losses.py:13-18: eps is ADDED to the norm
    return x / (x.pow(2).sum(dim=dim, keepdim=True).sqrt() + 1e-8)


def _same_class_mask(class_ids, n):
    # losses.py:55-66 / :224-232 -- host loop, one numpy compare per sample
    rows = []
    for i in range(n):
        row = (class_ids == class_ids[i]).astype(np.uint8)
        row[i] = 0
        rows.append(row[None, :])
    return torch.from_numpy(np.concatenate(rows, 0).astype(bool))


def score_caption(word_bdt, region_bdr, word_mask_bt1, gamma1, gamma2, checks=True):
    """One caption (already tiled to the batch) against every image; losses.py:95-216."""
    ctx = _unit(region_bdr.transpose(1, 2).contiguous(), 2)          # (B, R, D)
    qry = _unit(word_bdt.transpose(1, 2).contiguous(), 2)            # (B, T, D)
    s = torch.bmm(qry, ctx.transpose(1, 2))                          # (B, T, R)
    s = s.masked_fill_(word_mask_bt1 == 0, float("-inf"))
    if checks:                                                        # :133-137
        for mk, row in zip(word_mask_bt1[0], s[0]):
            if mk == 0:
                assert row[0] == float("-inf")
    p = F.softmax(s.transpose(1, 2), dim=-1)                         # (B, R, T): over words
    if checks:                                                        # :145, :155-159
        assert torch.isclose(p[0][0].sum(), torch.tensor(1.0), rtol=1e-5)
        for mk, val in zip(word_mask_bt1[0], p[0][0]):
            if mk == 0:
                assert val == 0.0
    a = F.softmax(gamma1 * p, dim=1)                                 # over regions
    if checks:
        assert torch.isnan(a).sum() == 0
    a = a.permute(0, 2, 1)                                           # (B, T, R)
    c = torch.bmm(a, ctx)                                            # (B, T, D)
    if checks:
        assert torch.isnan(c).sum() == 0
    rho = F.cosine_similarity(c, qry, dim=2, eps=1e-6)               # (B, T)
    e = (rho * gamma2).exp_().sum(dim=1)
    return p, torch.log(torch.pow(e, 1.0 / gamma2))


def words_loss(region_bdr, words_bdt, labels, class_ids, words_mask, gamma1, gamma2, gamma3, checks=True):
    """losses.py:219-272.  Inputs in the reference layout (B,D,R) / (B,D,T)."""
    n = words_bdt.shape[0]
    sims = []
    for i in range(n):
        w = words_bdt[i].unsqueeze(0).contiguous().repeat(n, 1, 1)
        m = words_mask[i].contiguous().repeat(n, 1).unsqueeze(-1)
        _, r_qd = score_caption(w, region_bdr, m, gamma1, gamma2, checks)
        sims.append(r_qd)
    sim = torch.stack(sims) * gamma3
    if class_ids is not None:
        sim.data.masked_fill_(_same_class_mask(class_ids, n), float("-inf"))
    return F.cross_entropy(sim, labels), F.cross_entropy(sim.transpose(0, 1), labels)


def sent_loss(img_bd, txt_bd, labels, class_ids, gamma3, eps=1e-8):
    """losses.py:51-91."""
    n = img_bd.shape[0]
    a, b = img_bd.unsqueeze(0), txt_bd.unsqueeze(0)
    na = torch.norm(a, 2, dim=2, keepdim=True)
    nb = torch.norm(b, 2, dim=2, keepdim=True)
    sc = torch.bmm(a, b.transpose(1, 2)) / torch.bmm(na, nb.transpose(1, 2)).clamp(min=eps) * gamma3
    sc = sc.squeeze(0)
    if class_ids is not None:
        sc.data.masked_fill_(_same_class_mask(class_ids, n), float("-inf"))
    return F.cross_entropy(sc, labels), F.cross_entropy(sc.transpose(0, 1), labels)


def step(x, gammas=(4.0, 5.0, 10.0), checks=True):
    """One words_loss + sent_loss forward + backward on a ``make_inputs`` dict (numpy).
    Returns dict(losses..., grads...) as numpy."""
    w = torch.tensor(x["words"], dtype=torch.float32, requires_grad=True)
    r = torch.tensor(x["regions"], dtype=torch.float32, requires_grad=True)
    si = torch.tensor(x["img"], dtype=torch.float32, requires_grad=True)
    st = torch.tensor(x["sent"], dtype=torch.float32, requires_grad=True)
    lab = torch.tensor(x["labels"], dtype=torch.int64)
    m = torch.tensor(x["mask"], dtype=torch.int64)
    w0, w1 = words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), lab, x["class_ids"], m, *gammas, checks=checks)
    s0, s1 = sent_loss(si, st, lab, x["class_ids"], gammas[2])
    (w0 + w1 + s0 + s1).backward()
    return dict(w_loss0=w0.item(), w_loss1=w1.item(), s_loss0=s0.item(), s_loss1=s1.item(),
                dwords=w.grad.numpy(), dregions=r.grad.numpy(), dimg=si.grad.numpy(), dtxt=st.grad.numpy())


def ntxent_step(z_i, z_j, temperature):
    """NT_Xent.forward + backward the way the reference runs it (nt_xent.py:16-35): the (2B, 2B, D) broadcast
    cosine (:24), positives from the +-B diagonals (:26-29), negatives through the boolean mask of masks.py (:30),
    CrossEntropy(sum) / 2B (:32-35).  Returns dict(loss, dz_i, dz_j)."""
    a = torch.tensor(np.asarray(z_i), dtype=torch.float32, requires_grad=True)
    b = torch.tensor(np.asarray(z_j), dtype=torch.float32, requires_grad=True)
    bsz = a.shape[0]
    n2 = 2 * bsz
    p = torch.cat((a, b), dim=0)
    sim = torch.nn.functional.cosine_similarity(p.unsqueeze(1), p.unsqueeze(0), dim=2) / temperature
    positives = torch.cat((sim.diagonal(bsz), sim.diagonal(-bsz))).reshape(n2, 1)
    idx = torch.arange(n2)
    keep = (idx[:, None] != idx[None, :]) & ((idx[:, None] - idx[None, :]).abs() != bsz)
    logits = torch.cat((positives, sim[keep].reshape(n2, -1)), dim=1)
    loss = torch.nn.functional.cross_entropy(logits, torch.zeros(n2, dtype=torch.long), reduction="sum") / n2
    loss.backward()
    return dict(loss=float(loss.detach()), dz_i=a.grad.numpy().copy(), dz_j=b.grad.numpy().copy())


def rm_special_token_step(mask, words_emb, dout):
    """rm_special_token forward + backward the way the reference runs it (pretrain_DAMSM.py:58-79): a Python loop over
    the batch, one boolean reduction / torch.where per caption, slices + cat, a final stack.  Returns (out, mask_new, dx)."""
    x = torch.tensor(np.asarray(words_emb), requires_grad=True)
    m = torch.tensor(np.asarray(mask))
    n = x.shape[1]
    embs, masks = [], []
    for i in range(x.shape[0]):
        if int(m[i].sum()) == n:                                    # no padding: <eos> is the last row
            sel = slice(1, n - 1)
            embs.append(x[i, sel, :])
            masks.append(m[i, sel])
        else:
            e = int(torch.where(m[i] == 0)[0].min())                # first padding position; <eos> sits at e - 1
            embs.append(torch.cat([x[i, 1:e - 1, :], x[i, e:, :]], dim=0))
            masks.append(torch.cat([m[i, 1:e - 1], m[i, e:]], dim=0))
    out = torch.stack(embs, dim=0)
    out.backward(torch.tensor(np.asarray(dout)))
    return out.detach().numpy(), torch.stack(masks, dim=0).numpy(), x.grad.numpy()


def project_regions_step(subr, weight, bias, dy):
    """linear_subr + CLS drop the way the reference runs it (model.py:46 / pretrain_DAMSM.py:359 then :125): nn.Linear on
    the flattened tokens (CLS included), view back, slice, + autograd backward.  Returns (y, dsubr, dweight, dbias)."""
    x = torch.tensor(np.asarray(subr), dtype=torch.float32, requires_grad=True)
    w = torch.tensor(np.asarray(weight), dtype=torch.float32, requires_grad=True)
    b = torch.tensor(np.asarray(bias), dtype=torch.float32, requires_grad=True)
    bsz, _, k = x.shape
    y = torch.nn.functional.linear(x.view(-1, k), w, b).view(bsz, -1, w.shape[0])[:, 1:, :]
    y.backward(torch.tensor(np.asarray(dy), dtype=torch.float32))
    return y.detach().numpy(), x.grad.numpy(), w.grad.numpy(), b.grad.numpy()


def r_precision_step(img_code, sent_codes):
    """R-precision scoring the way the reference runs it (trainer.py:587-603): a Python loop over the batch, per image a
    1 x C torch.mm, two torch.norm calls, a 1 x C product of norms clamped at 1e-8, argmax.  Returns (scores0, hit)."""
    img = torch.tensor(np.asarray(img_code), dtype=torch.float32)
    cands = torch.tensor(np.asarray(sent_codes), dtype=torch.float32)
    out, hits = [], []
    for i in range(img.shape[0]):
        one = img[i].unsqueeze(0)
        sc = torch.mm(one, cands[i].t())
        nrm = torch.mm(torch.norm(one, 2, dim=1, keepdim=True), torch.norm(cands[i], 2, dim=1, keepdim=True).t())
        s0 = sc / nrm.clamp(min=1e-8)
        out.append(s0[0])
        hits.append(bool(torch.argmax(s0) == 0))
    return torch.stack(out).numpy(), np.array(hits)
