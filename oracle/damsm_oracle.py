"""CPU oracle for the DAMSM matching-loss hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``t2i_clip-gan_b200``) never does: it calls the CUDA C-ABI library and fails
loudly when that is missing.

What this is
------------
A vectorised float64 numpy restatement of the reference algorithm

  * ``l2norm``                 /root/reference/DMGAN+CLIP/code/miscc/losses.py:13-18
  * ``similarity_text_image``  .../miscc/losses.py:95-216
  * ``words_loss``             .../miscc/losses.py:219-272
  * ``sent_loss``              .../miscc/losses.py:51-91
  * class_ids masking          .../miscc/losses.py:55-66,84 and :224-232,256-263
  * ``func_attention``         /root/reference/DMGAN+CLIP/code/GlobalAttention.py:38-160

plus hand-derived backward passes.  The backward is written with the *same
closed forms the CUDA kernels use* (Gram-matrix form of the context norm,
per-word a_t / b_t coefficients, column term W_r, per-image H_j) so that the
formulas themselves are validated on the CPU against autograd of the real
reference before any GPU time is spent.

Parity pinning
--------------
The reference ships no tests and no golden vectors, so
the oracle is pinned against outputs of the reference source itself, imported
unmodified in the build container by ``oracle/ref_shim.py`` and recorded by
``oracle/make_golden.py`` into ``tests/golden/*.npz``
(``tests/test_oracle_golden.py`` checks the oracle against them everywhere;
``tests/test_oracle_vs_reference.py`` re-runs the live reference when
``/root/reference`` exists).

Conventions: caption i, image j, word t, region r.
``words``   (B, T, D)  raw word embeddings        (reference passes (B, D, T))
``regions`` (Bc, R, D) raw region features         (reference passes (B, D, R))
``mask``    (B, T)     1 = real word, 0 = padding
Rows may be a shard of the batch: ``row_offset`` says which global caption the
first local row is (labels / class-mask diagonal use global indices).
"""
from __future__ import annotations

import numpy as np

L2_EPS = 1e-8      # losses.py:13  (added to the norm)
COS_EPS = 1e-6     # losses.py:197 (each norm clamped)
SENT_EPS = 1e-8    # losses.py:51,79 (product of norms clamped)


# --------------------------------------------------------------------------- l2norm
def l2norm(x: np.ndarray, axis: int = -1):
    """losses.py:13-18 -- x / (sqrt(sum x^2) + 1e-8).  Returns (xhat, raw_norm)."""
    x = np.asarray(x, dtype=np.float64)
    nrm = np.sqrt((x * x).sum(axis=axis, keepdims=True))
    return x / (nrm + L2_EPS), nrm


def l2norm_bwd(x: np.ndarray, nrm: np.ndarray, dxhat: np.ndarray):
    """VJP of l2norm: dx = (dxhat - (xhat.dxhat) * x/||x||) / (||x|| + eps)."""
    s = nrm + L2_EPS
    xhat = x / s
    dot = (xhat * dxhat).sum(axis=-1, keepdims=True)
    safe = np.where(nrm > 0, nrm, 1.0)
    return (dxhat - dot * x / safe) / s


# --------------------------------------------------------------------------- per-caption core
def _pair_block(q, u, m, v, G, gamma1, gamma2):
    """One caption (q: (T,D) normalised, u: (T,) = ||q_t||, m: (T,)) against all
    images (v: (Bc,R,D) normalised, G: (Bc,R,R) = v v^T).  losses.py:113-203.

    Returns a dict of everything the backward needs.
    """
    S = np.einsum("td,jrd->jtr", q, v)                      # losses.py:117 (raw, unmasked)
    e1 = np.exp(S) * m[None, :, None]                       # :127 mask -> exp(-inf) = 0
    Z = e1.sum(axis=1)                                      # softmax over words, :143-144
    P = e1 / Z[:, None, :]
    e2 = np.exp(gamma1 * P)                                 # :173-174 softmax over regions
    Y = e2.sum(axis=2)
    A = e2 / Y[:, :, None]
    M = np.einsum("jtr,jrs->jts", A, G)                     # (A G): c_t . v_r
    N = (A * S).sum(axis=2)                                 # c_t . q_t   (identity, App. A)
    n2 = np.maximum((A * M).sum(axis=2), 0.0)               # ||c_t||^2 = a^T G a
    n = np.sqrt(n2)
    nc = np.maximum(n, COS_EPS)                             # :197 CosineSimilarity eps
    uc = np.maximum(u, COS_EPS)[None, :]
    rho = N / (nc * uc)                                     # :198
    x = gamma2 * rho
    xm = x.max(axis=1, keepdims=True)
    ex = np.exp(x - xm)
    lse = np.log(ex.sum(axis=1)) + xm[:, 0]
    Rqd = lse / gamma2                                      # :199-203
    omega = ex / ex.sum(axis=1, keepdims=True)
    return dict(S=S, P=P, A=A, M=M, N=N, n=n, nc=nc, uc=uc, rho=rho, Rqd=Rqd, omega=omega)


def words_sim(words, regions, mask, gamma1, gamma2):
    """R(Q,D)[i, j] for every caption i and image j (before gamma3).  fp64."""
    words = np.asarray(words, np.float64)
    regions = np.asarray(regions, np.float64)
    mask = (np.asarray(mask) != 0).astype(np.float64)
    q, _ = l2norm(words)
    v, _ = l2norm(regions)
    u = np.sqrt((q * q).sum(-1))
    G = np.einsum("jrd,jsd->jrs", v, v)
    out = np.empty((words.shape[0], regions.shape[0]))
    for i in range(words.shape[0]):
        out[i] = _pair_block(q[i], u[i], mask[i], v, G, gamma1, gamma2)["Rqd"]
    return out


# --------------------------------------------------------------------------- masked bidirectional CE
def class_mask(class_ids_rows, class_ids_cols, row_offset=0):
    """losses.py:55-66 / :224-232 -- True where the logit is forced to -inf:
    same class and not the caption's own image (global index)."""
    if class_ids_rows is None:
        return None
    cr = np.asarray(class_ids_rows).reshape(-1, 1)
    cc = np.asarray(class_ids_cols).reshape(1, -1)
    mk = cr == cc
    rows = np.arange(cr.shape[0]) + row_offset
    mk[np.arange(cr.shape[0]), rows] = False
    return mk


def ce_block_stats(logits, mask):
    """Row LSE (complete) and column (max, sum-exp) partials of a block."""
    L = np.array(logits, np.float64, copy=True)
    if mask is not None:
        L[mask] = -np.inf
    rmax = L.max(axis=1)
    rlse = np.log(np.exp(L - rmax[:, None]).sum(axis=1)) + rmax
    cmax = L.max(axis=0)
    csum = np.exp(L - cmax[None, :]).sum(axis=0)
    return L, rlse, cmax, csum


def ce_bidir_full(logits, labels, mask=None):
    """Unsharded: returns (loss_rows, loss_cols, dloss_rows/dlogits, dloss_cols/dlogits).

    loss_rows = CE(logits, labels) ; loss_cols = CE(logits.T, labels)."""
    L, rlse, cmax, csum = ce_block_stats(logits, mask)
    B = L.shape[0]
    labels = np.asarray(labels).astype(np.int64)
    idx = np.arange(B)
    clse = np.log(csum) + cmax
    loss_r = (rlse - L[idx, labels]).mean()
    loss_c = (clse - L[labels, idx]).mean()
    gr = np.exp(L - rlse[:, None])
    gr[idx, labels] -= 1.0
    gc = np.exp(L - clse[None, :])
    gc[labels, idx] -= 1.0
    return loss_r, loss_c, gr / B, gc / B


# --------------------------------------------------------------------------- words_loss
def words_loss(words, regions, mask, labels, class_ids, gamma1, gamma2, gamma3,
               g0=1.0, g1=1.0, want_grads=True):
    """losses.py:219-272 in fp64 with analytic gradients.

    Returns dict(loss0, loss1, sim (gamma3-scaled, masked), dwords, dregions)
    where d* are gradients of ``g0*loss0 + g1*loss1`` w.r.t. the raw inputs.
    loss0 is text->image (CE over images for each caption), loss1 image->text.
    """
    words = np.asarray(words, np.float64)
    regions = np.asarray(regions, np.float64)
    mk = (np.asarray(mask) != 0).astype(np.float64)
    B, T, D = words.shape
    Bc, R, _ = regions.shape
    q, qn = l2norm(words)
    v, vn = l2norm(regions)
    u = np.sqrt((q * q).sum(-1))
    G = np.einsum("jrd,jsd->jrs", v, v)
    blocks = []
    sim = np.empty((B, Bc))
    for i in range(B):
        blk = _pair_block(q[i], u[i], mk[i], v, G, gamma1, gamma2)
        sim[i] = gamma3 * blk["Rqd"]                        # losses.py:254
        blocks.append(blk if want_grads else None)
    cmask = class_mask(class_ids, class_ids) if class_ids is not None else None
    loss0, loss1, gr, gc = ce_bidir_full(sim, labels, cmask)
    simm = sim.copy()
    if cmask is not None:
        simm[cmask] = -np.inf
    out = dict(loss0=loss0, loss1=loss1, sim=simm)
    if not want_grads:
        return out
    g = g0 * gr + g1 * gc                                   # dL/dsim, exactly 0 at masked entries
    dq = np.zeros_like(q)
    dv = np.zeros_like(v)
    H = np.zeros_like(G)
    for i in range(B):
        k = blocks[i]
        S, P, A, M = k["S"], k["P"], k["A"], k["M"]
        beta = g[i][:, None] * gamma3 * k["omega"]          # dL/drho  (Bc, T)
        a = beta / (k["nc"] * k["uc"])
        okn = k["n"] > COS_EPS
        b = np.where(okn, beta * k["rho"] / np.where(okn, k["n"] ** 2, 1.0), 0.0)
        dP = gamma1 * A * (a[:, :, None] * S - b[:, :, None] * M)
        W = (P * dP).sum(axis=1)                            # (Bc, R) column term of softmax-over-words
        dS = a[:, :, None] * A + P * (dP - W[:, None, :])
        dq[i] += np.einsum("jtr,jrd->td", dS, v)
        oku = u[i] > COS_EPS
        kq = np.where(oku, (beta * k["rho"]).sum(axis=0) / np.where(oku, u[i] ** 2, 1.0), 0.0)
        dq[i] -= kq[:, None] * q[i]
        dv += np.einsum("jtr,td->jrd", dS, q[i])
        H += np.einsum("jt,jtr,jts->jrs", b, A, A)
    dv -= np.einsum("jrs,jsd->jrd", H, v)
    out["dwords"] = l2norm_bwd(words, qn, dq)
    out["dregions"] = l2norm_bwd(regions, vn, dv)
    out["dqhat"], out["dvhat"] = dq, dv
    return out


# --------------------------------------------------------------------------- sent_loss
def sent_loss(img, txt, labels, class_ids, gamma3, g0=1.0, g1=1.0):
    """losses.py:51-91.  scores0[i, j] = gamma3 * img_i.txt_j / max(|img_i||txt_j|, 1e-8);
    loss0 = CE(scores0) (image->text), loss1 = CE(scores0.T)."""
    img = np.asarray(img, np.float64)
    txt = np.asarray(txt, np.float64)
    ni = np.sqrt((img * img).sum(-1))
    nt = np.sqrt((txt * txt).sum(-1))
    dots = img @ txt.T
    nn_ = ni[:, None] * nt[None, :]
    den = np.maximum(nn_, SENT_EPS)
    sc = dots / den * gamma3
    cmask = class_mask(class_ids, class_ids) if class_ids is not None else None
    loss0, loss1, gr, gc = ce_bidir_full(sc, labels, cmask)
    scm = sc.copy()
    if cmask is not None:
        scm[cmask] = -np.inf
    g = (g0 * gr + g1 * gc) * gamma3                        # d/d(dots/den)
    live = nn_ > SENT_EPS                                   # clamp passes gradient only when not clamped
    gd = g / den
    gn = np.where(live, -g * dots / den ** 2, 0.0)          # d/d(ni*nt)
    dimg = gd @ txt + (gn * nt[None, :]).sum(1)[:, None] * img / np.where(ni > 0, ni, 1.0)[:, None]
    dtxt = gd.T @ img + (gn * ni[:, None]).sum(0)[:, None] * txt / np.where(nt > 0, nt, 1.0)[:, None]
    return dict(loss0=loss0, loss1=loss1, scores=scm, dimg=dimg, dtxt=dtxt)


# --------------------------------------------------------------------------- NT-Xent
NTX_EPS = 1e-8   # nn.CosineSimilarity default eps (nt_xent.py:14), clamps each norm


def nt_xent(z_i, z_j, temperature, g=1.0):
    """nt_xent.py:16-35 with the mask of masks.py:3-17.

    z = cat(z_i, z_j); sim = cos(z_a, z_b) / temperature (:23-24); positives on the +-B diagonals (:26-29);
    negatives = everything but the diagonal and the positive (:30, masks.py); logits = [positive, negatives],
    label 0, CrossEntropy(sum) / 2B (:32-35)  ==  mean_a( LSE_{b != a} sim[a, b] - sim[a, pos(a)] ).
    Returns dict(loss, sim (diag = -inf), dz_i, dz_j) with the analytic gradient of g * loss."""
    z = np.concatenate([np.asarray(z_i, np.float64), np.asarray(z_j, np.float64)], axis=0)
    n2 = z.shape[0]
    b = n2 // 2
    nr = np.sqrt((z * z).sum(-1))
    cn = np.maximum(nr, NTX_EPS)
    sim = (z @ z.T) / (cn[:, None] * cn[None, :]) / temperature
    np.fill_diagonal(sim, -np.inf)
    pos = (np.arange(n2) + b) % n2
    mx = sim.max(axis=1, keepdims=True)
    lse = np.log(np.exp(sim - mx).sum(axis=1)) + mx[:, 0]
    loss = float((lse - sim[np.arange(n2), pos]).mean())
    gm = np.exp(sim - lse[:, None])                           # softmax over b != a (exp(-inf) = 0 on the diagonal)
    gm[np.arange(n2), pos] -= 1.0
    gm *= g / n2                                              # dL/dsim
    w = gm / temperature / (cn[:, None] * cn[None, :])        # dL/d(z_a . z_b)
    sfin = np.where(np.isfinite(sim), sim, 0.0)
    live = nr > NTX_EPS
    coef = np.where(live, -(gm * sfin).sum(1) / np.where(live, nr, 1.0), 0.0) \
        + np.where(live, -(gm * sfin).sum(0) / np.where(live, nr, 1.0), 0.0)      # dL/d|z_x| via row x and column x
    dz = (w + w.T) @ z + coef[:, None] * z / np.where(nr > 0, nr, 1.0)[:, None]
    return dict(loss=loss, sim=sim, dz_i=dz[:b], dz_j=dz[b:])


# --------------------------------------------------------------------------- R-precision
def r_precision_scores(img_code, sent_codes, eps=1e-8):
    """trainer.py:587-603 for every image of the batch: scores = img . sent^T (:596), norm = |img| |sent| (:597-599),
    scores0 = scores / clamp(norm, min=1e-8) (:600), hit = argmax(scores0) == 0 (:601).  sent_codes (B, C, D), true
    caption first.  Returns (scores0 (B, C), hit (B,) bool)."""
    a = np.asarray(img_code, np.float64)
    c = np.asarray(sent_codes, np.float64)
    dots = np.einsum("bd,bcd->bc", a, c)
    nrm = np.sqrt((a * a).sum(-1))[:, None] * np.sqrt((c * c).sum(-1))
    s0 = dots / np.maximum(nrm, eps)
    return s0, s0.argmax(axis=1) == 0


# --------------------------------------------------------------------------- region projection
def project_regions(subr, weight, bias=None, dy=None):
    """AddLinearOnCLIP.linear_subr (nn.Linear(768, 512), model.py:21,46,78 / pretrain_DAMSM.py:350,359) followed by the
    CLS drop of pretrain_DAMSM.py:125: y[b, r-1] = subr[b, r] . W^T + bias for r = 1..R.  Returns y (B, R, N); with
    ``dy`` (B, R, N) also (dsubr (B, R+1, K) with zero CLS rows, dweight (N, K), dbias (N))."""
    x = np.asarray(subr, np.float64)
    w = np.asarray(weight, np.float64)
    y = x[:, 1:, :] @ w.T
    if bias is not None:
        y = y + np.asarray(bias, np.float64)
    if dy is None:
        return y
    g = np.asarray(dy, np.float64)
    dx = np.zeros_like(x)
    dx[:, 1:, :] = g @ w
    dw = np.einsum("brn,brk->nk", g, x[:, 1:, :])
    return y, dx, dw, g.sum(axis=(0, 1))


# --------------------------------------------------------------------------- rm_special_token
def rm_special_token(mask, words_emb):
    """pretrain_DAMSM.py:58-79.  Per caption: no 0 in the mask -> rows 1..n-2 (:67-69); else with e = index of the
    first 0 -> rows 1..e-2 followed by rows e..n-1 (:71-75), i.e. <sos> (row 0) and <eos> (row e-1) are removed.
    Returns (words_emb_new (B, n-2, D), mask_new (B, n-2), src (B, n-2) source row of every output row)."""
    mask = np.asarray(mask)
    words_emb = np.asarray(words_emb)
    b, n = mask.shape
    src = np.empty((b, n - 2), np.int64)
    for i in range(b):
        zeros = np.nonzero(mask[i] == 0)[0]
        e = int(zeros.min()) if zeros.size else n
        keep = [k for k in range(n) if k != 0 and k != e - 1]
        assert len(keep) == n - 2, "caption with fewer than two leading 1s: undefined in the reference"
        src[i] = keep
    rows = np.arange(b)[:, None]
    return words_emb[rows, src], mask[rows, src], src


# --------------------------------------------------------------------------- func_attention
def func_attention(query, context, gamma1, query_mask, d_wc=None):
    """GlobalAttention.py:38-160.

    query   (B, T, D) raw words (reference passes (B, D, T));
    context (B, R, D) raw regions (reference passes (B, D, R));
    query_mask (B, T).
    Returns weightedContext (B, T, D)  [built from the RAW context, :153] and
    attn (B, T, R) = softmax over words (the reference reshapes it to (B,T,h,w)).
    With ``d_wc`` (B,T,D) also returns (dquery, dcontext), the VJP of
    weightedContext (attn is returned detached from that VJP: the reference's attn
    output has no consumer that back-propagates).
    """
    query = np.asarray(query, np.float64)
    context = np.asarray(context, np.float64)
    mk = (np.asarray(query_mask) != 0).astype(np.float64)
    q, qn = l2norm(query)
    v, vn = l2norm(context)
    S = np.einsum("btd,brd->btr", q, v)                     # GlobalAttention.py:90 (transposed)
    e1 = np.exp(S) * mk[:, :, None]                         # :103-104
    P = e1 / e1.sum(axis=1, keepdims=True)
    e2 = np.exp(gamma1 * P)                                 # :146-147
    A = e2 / e2.sum(axis=2, keepdims=True)
    wc = np.einsum("btr,brd->btd", A, context)              # :153 raw context
    if d_wc is None:
        return wc, P
    d_wc = np.asarray(d_wc, np.float64)
    dA = np.einsum("btd,brd->btr", d_wc, context)
    dcontext = np.einsum("btr,btd->brd", A, d_wc)
    dX = A * (dA - (A * dA).sum(axis=2, keepdims=True))
    dP = gamma1 * dX
    dS = P * (dP - (P * dP).sum(axis=1, keepdims=True))
    dq = np.einsum("btr,brd->btd", dS, v)
    dv = np.einsum("btr,btd->brd", dS, q)
    dquery = l2norm_bwd(query, qn, dq)
    dcontext = dcontext + l2norm_bwd(context, vn, dv)
    return wc, P, dquery, dcontext


# --------------------------------------------------------------------------- seeded inputs
def make_inputs(B, T, R, D=512, seed=2026, class_ids=True, n_classes=200, dtype=np.float32):
    """Seeded synthetic generator of SURVEY.md section 8(d): a shared latent makes
    matched pairs moderately similar so the loss and gradients are non-degenerate."""
    rng = np.random.default_rng(seed)
    s = rng.standard_normal((B, 1, D))
    words = 0.25 * s + rng.standard_normal((B, T, D))
    regions = 0.25 * s + rng.standard_normal((B, R, D))
    sent = 0.25 * s[:, 0] + rng.standard_normal((B, D))
    img = 0.25 * s[:, 0] + rng.standard_normal((B, D))
    lo = max(2, T // 3)
    cap_len = rng.integers(lo, T + 1, size=B)
    mask = (np.arange(T)[None, :] < cap_len[:, None]).astype(np.int64)
    cids = rng.integers(0, n_classes, size=B).astype(np.int64) if class_ids else None
    return dict(words=words.astype(dtype), regions=regions.astype(dtype), sent=sent.astype(dtype),
                img=img.astype(dtype), mask=mask, cap_len=cap_len, class_ids=cids,
                labels=np.arange(B, dtype=np.int64))
