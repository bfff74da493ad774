"""A CPU engine with the same block-level interface as ``t2i_clip-gan_b200.engine.CudaEngine``, built on
the oracle (numpy / torch fp64).  TEST INFRASTRUCTURE: it exists so that the multi-rank host logic of
``ops.py`` (all-gather of image-side tensors, column log-sum-exp combine, loss all-reduce, gradient
reduce-scatter) can be exercised with the gloo backend on a machine without GPUs.  The product never
imports it."""
from __future__ import annotations

import numpy as np
import torch

from oracle import damsm_oracle as O


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64)


def _t(a, dtype=torch.float32):
    return torch.tensor(np.asarray(a), dtype=dtype)


class CheckerEngine:
    name = "checker"
    precision = "fp32"

    def l2norm_fwd(self, x3, want_bf16=False, pad8=False):
        x = _np(x3)
        xhat, nrm = O.l2norm(x)
        un = np.sqrt((xhat * xhat).sum(-1))
        return _t(xhat, torch.float64), None, _t(nrm[..., 0], torch.float64), _t(un, torch.float64)

    def l2norm_bwd(self, x3, norm, dxhat, kq=None):
        x = _np(x3)
        xhat, nrm = O.l2norm(x)
        g = _np(dxhat)
        if kq is not None:
            u2 = (xhat * xhat).sum(-1, keepdims=True)
            g = g - _np(kq)[..., None] * xhat / u2
        return _t(O.l2norm_bwd(x, nrm, g), x3.dtype)

    def gram(self, vhat):
        v = _np(vhat)
        return _t(np.einsum("jrd,jsd->jrs", v, v), torch.float64)

    def pack_columns(self, gram, vhat, vhat16=None):
        return {"gram": _np(gram)}

    def words_prepare_columns(self, vhat, vhat16=None):
        return self.pack_columns(self.gram(vhat), vhat, vhat16)

    def gram_bwd(self, hmat, vhat, dvhat):
        return _t(_np(dvhat) - np.einsum("jrs,jsd->jrd", _np(hmat), _np(vhat)), torch.float64)

    def _blocks(self, qhat, vhat, col, unorm, mask_u8, gammas):
        q, v, u, m = _np(qhat), _np(vhat), _np(unorm), _np(mask_u8)
        return [O._pair_block(q[i], u[i], m[i], v, col["gram"], gammas[0], gammas[1]) for i in range(q.shape[0])]

    def words_fwd(self, qhat, qhat16, vhat, col, unorm, mask_u8, gammas, want_stats=True):
        blocks = self._blocks(qhat, vhat, col, unorm, mask_u8, gammas)
        return _t(np.stack([gammas[2] * b["Rqd"] for b in blocks]), torch.float64)

    def ce_stats(self, logits, cls_rows, cls_cols, row_offset):
        L = logits.numpy()
        mk = O.class_mask(None if cls_rows is None else cls_rows.numpy(),
                          None if cls_cols is None else cls_cols.numpy(), row_offset)
        if mk is not None:
            L[mk] = -np.inf            # in place, like the CUDA kernel
        _, rlse, cmax, csum = O.ce_block_stats(L, None)
        return _t(rlse, torch.float64), _t(cmax, torch.float64), _t(csum, torch.float64)

    def ce_losses(self, logits, row_lse, col_lse, labels, row_offset, b_total):
        L, rl, cl, lab = _np(logits), _np(row_lse), _np(col_lse), labels.numpy()
        br = L.shape[0]
        rows = np.arange(br)
        l0 = (rl - L[rows, lab[row_offset + rows]]).sum() / b_total
        tgt = lab - row_offset
        own = (tgt >= 0) & (tgt < br)
        cols = np.arange(L.shape[1])[own]
        l1 = (cl[cols] - L[tgt[own], cols]).sum() / b_total
        return _t([l0, l1], torch.float64)

    def _g(self, logits, row_lse, col_lse, labels, gscale, row_offset, b_total):
        L, rl, cl, lab, gs = _np(logits), _np(row_lse), _np(col_lse), labels.numpy(), _np(gscale)
        br, bc = L.shape
        gi = row_offset + np.arange(br)
        gr = np.exp(L - rl[:, None])
        gr[np.arange(br), lab[gi]] -= 1.0
        gc = np.exp(L - cl[None, :])
        gc -= (lab[None, :] == gi[:, None]).astype(np.float64)
        g = (gs[0] * gr + gs[1] * gc) / b_total
        g[~np.isfinite(L)] = 0.0
        return g

    def words_bwd(self, qhat, qhat16, vhat, col, unorm, mask_u8, sim, row_lse, col_lse, labels, gscale,
                  row_offset, b_total, gammas, need_dq=True, need_dv=True):
        g = self._g(sim, row_lse, col_lse, labels, gscale, row_offset, b_total)
        q, v, u = _np(qhat), _np(vhat), _np(unorm)
        blocks = self._blocks(qhat, vhat, col, unorm, mask_u8, gammas)
        g1, g2, g3 = gammas
        dq, dv, H = np.zeros_like(q), np.zeros_like(v), np.zeros_like(col["gram"])
        kq = np.zeros(q.shape[:2])
        for i, k in enumerate(blocks):
            S, P, A, M = k["S"], k["P"], k["A"], k["M"]
            beta = g[i][:, None] * g3 * k["omega"]
            a = beta / (k["nc"] * k["uc"])
            okn = k["n"] > O.COS_EPS
            b = np.where(okn, beta * k["rho"] / np.where(okn, k["n"] ** 2, 1.0), 0.0)
            dP = g1 * A * (a[:, :, None] * S - b[:, :, None] * M)
            W = (P * dP).sum(axis=1)
            dS = a[:, :, None] * A + P * (dP - W[:, None, :])
            dq[i] = np.einsum("jtr,jrd->td", dS, v)
            kq[i] = (beta * k["rho"]).sum(axis=0)
            dv += np.einsum("jtr,td->jrd", dS, q[i])
            H += np.einsum("jt,jtr,jts->jrs", b, A, A)
        return _t(dq, torch.float64), _t(dv, torch.float64), _t(H, torch.float64), _t(kq, torch.float64)

    def cos_logits(self, a, b, gamma3, eps):
        A, B = _np(a), _np(b)
        na, nb = np.sqrt((A * A).sum(-1)), np.sqrt((B * B).sum(-1))
        return (_t(A @ B.T / np.maximum(na[:, None] * nb[None, :], eps) * gamma3, torch.float64),
                _t(na, torch.float64), _t(nb, torch.float64))

    def cos_logits_bwd(self, a, b, na, nb, logits, row_lse, col_lse, labels, gscale, row_offset, b_total,
                       gamma3, eps):
        g = self._g(logits, row_lse, col_lse, labels, gscale, row_offset, b_total) * gamma3
        A, B, ni, nt = _np(a), _np(b), _np(na), _np(nb)
        L = _np(logits)
        nn_ = ni[:, None] * nt[None, :]
        den = np.maximum(nn_, eps)
        ratio = np.where(np.isfinite(L), L / gamma3, 0.0)
        gd = g / den
        gn = np.where(nn_ > eps, -g * ratio / den, 0.0)
        da = gd @ B + (gn * nt[None, :]).sum(1)[:, None] * A / ni[:, None]
        db = gd.T @ A + (gn * ni[:, None]).sum(0)[:, None] * B / nt[:, None]
        return _t(da, a.dtype), _t(db, b.dtype)
