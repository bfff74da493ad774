"""Closed form for padded words (round-2 groundwork; TEST INFRASTRUCTURE, CPU only).

The reference keeps every padded word in the loss (losses.py:127,173-174,198-203): its softmax-over-words probability is
exactly 0, so its attention over regions is uniform, its context vector is the image's mean region v_bar_j, and its
cosine is rho_bar = (q_t . v_bar_j) / (max(|v_bar_j|, eps) max(|q_t|, eps)) -- one small GEMM of the padded words
against the B mean vectors instead of a full pass of the pair kernel.  The valid words never see the padded ones (the
softmax over words already excludes them), so

    sim[i, j] = gamma3/gamma2 * log( sum_{t valid} exp(gamma2 rho_t) + sum_{t padded} exp(gamma2 rho_bar_t) )

with rho_t from the pair kernel run on the valid words only.  This module states that split in torch fp64 with
autograd, so that tests can hold it against the oracle's full computation and its analytic gradients.
"""
from __future__ import annotations

import numpy as np
import torch

from . import damsm_oracle as O


def words_loss_split(words, regions, mask, labels, class_ids, gamma1, gamma2, gamma3):
    """Same contract as ``damsm_oracle.words_loss`` (returns loss0, loss1, sim, dwords, dregions), computed as
    'pair computation on the valid words' + 'closed form for the padded words'."""
    w = torch.tensor(np.asarray(words), dtype=torch.float64, requires_grad=True)
    r = torch.tensor(np.asarray(regions), dtype=torch.float64, requires_grad=True)
    m = torch.tensor(np.asarray(mask) != 0)
    B = w.shape[0]
    q = w / (w.pow(2).sum(-1, keepdim=True).sqrt() + O.L2_EPS)          # losses.py:13-18
    v = r / (r.pow(2).sum(-1, keepdim=True).sqrt() + O.L2_EPS)
    u = q.pow(2).sum(-1).sqrt()
    vbar = v.mean(dim=1)                                                # (B, D): context of every padded word
    nbar = vbar.pow(2).sum(-1).sqrt().clamp_min(O.COS_EPS)
    rows = []
    for i in range(B):
        valid, padded = m[i], ~m[i]
        qv, uv = q[i][valid], u[i][valid]
        S = torch.einsum("td,jrd->jtr", qv, v)                          # valid words only
        P = torch.softmax(S, dim=1)                                     # over words
        A = torch.softmax(gamma1 * P, dim=2)                            # over regions
        c = torch.einsum("jtr,jrd->jtd", A, v)
        rho = (c * qv[None]).sum(-1) / (c.pow(2).sum(-1).sqrt().clamp_min(O.COS_EPS) * uv.clamp_min(O.COS_EPS)[None])
        terms = [gamma2 * rho]
        if bool(padded.any()):
            qp, up = q[i][padded], u[i][padded]
            rho_bar = (vbar @ qp.T) / (nbar[:, None] * up.clamp_min(O.COS_EPS)[None, :])   # (B images, padded words)
            terms.append(gamma2 * rho_bar)
        rows.append(torch.logsumexp(torch.cat(terms, dim=1), dim=1) / gamma2)
    sim = gamma3 * torch.stack(rows)                                    # (captions, images)
    lab = torch.tensor(np.asarray(labels), dtype=torch.int64)
    simm = sim
    if class_ids is not None:
        cmask = torch.tensor(O.class_mask(class_ids, class_ids))
        simm = sim.masked_fill(cmask, float("-inf"))
    loss0 = torch.nn.functional.cross_entropy(simm, lab)
    loss1 = torch.nn.functional.cross_entropy(simm.T, lab)
    (loss0 + loss1).backward()
    return dict(loss0=loss0.item(), loss1=loss1.item(), sim=simm.detach().numpy(), dwords=w.grad.numpy(),
                dregions=r.grad.numpy())
