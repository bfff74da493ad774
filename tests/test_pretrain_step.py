"""The whole pretrain step of INTEGRATION.md section 2b through the drop-ins (project_regions -> rm_special_token ->
words_loss + sent_loss + nt_xent -> backward) against the same chain of oracle functions."""
import importlib

import numpy as np
import pytest
import torch

from oracle import damsm_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("precision,tol", [("fp32", 3e-3), ("bf16", 5e-3)])    # the projection runs as TF32 either way
def test_pretrain_step_with_every_drop_in(precision, tol):
    pkg = importlib.import_module("t2i_clip-gan_b200")
    B, R, K, N, n = 8, 49, 768, 512, 30
    rng = np.random.default_rng(17)
    shared = rng.standard_normal((B, 1, N))
    subr = rng.standard_normal((B, R + 1, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = (0.1 * rng.standard_normal(N)).astype(np.float32)
    words_full = (0.5 * shared + rng.standard_normal((B, n, N))).astype(np.float32)        # <sos> w.. <eos> pad..
    lens = rng.integers(4, n + 1, B)
    lens[0] = n
    mask_full = (np.arange(n)[None, :] < lens[:, None]).astype(np.int64)
    sent = (0.5 * shared[:, 0] + rng.standard_normal((B, N))).astype(np.float32)
    img = (0.5 * shared[:, 0] + rng.standard_normal((B, N))).astype(np.float32)
    sent2 = (0.5 * shared[:, 0] + rng.standard_normal((B, N))).astype(np.float32)
    labels = np.arange(B)

    # ---- oracle chain (fp64) ----
    y = O.project_regions(subr, W, bias)
    w_new, m_new, src = O.rm_special_token(mask_full, words_full)
    ow = O.words_loss(w_new, y, m_new, labels, None, 4.0, 5.0, 10.0)
    os_ = O.sent_loss(img, sent, labels, None, 10.0)
    on = O.nt_xent(sent, sent2, 0.5)
    _, _, dW, dbias = O.project_regions(subr, W, bias, ow["dregions"])
    dwords_full = np.zeros_like(words_full, dtype=np.float64)
    dwords_full[np.arange(B)[:, None], src] = ow["dwords"]
    dsent = os_["dtxt"] + on["dz_i"]

    # ---- drop-ins ----
    t = lambda a, g=False: torch.tensor(a, device="cuda").requires_grad_(g)
    subr_t, W_t, b_t = t(subr), t(W, True), t(bias, True)
    words_t, sent_t, img_t, sent2_t = t(words_full, True), t(sent, True), t(img, True), t(sent2)
    feats = pkg.project_regions(subr_t, W_t, b_t)
    w_emb, w_mask = pkg.rm_special_token(torch.tensor(mask_full, device="cuda"), words_t)
    lab = torch.arange(B, device="cuda")
    w0, w1, _ = pkg.words_loss(feats, w_emb.permute(0, 2, 1), lab, None, None, B, w_mask, 4.0, 5.0, 10.0,
                               precision=precision)
    s0, s1 = pkg.sent_loss(img_t, sent_t, lab, None, B)
    c = pkg.nt_xent(sent_t, sent2_t, 0.5)
    (w0 + w1 + s0 + s1 + c).backward()

    assert abs(w0.item() - ow["loss0"]) <= tol * max(1, abs(ow["loss0"])) and abs(w1.item() - ow["loss1"]) <= tol * max(1, abs(ow["loss1"]))
    assert abs(s0.item() - os_["loss0"]) <= 1e-5 and abs(c.item() - on["loss"]) <= 1e-5
    assert rel(W_t.grad.cpu().numpy(), dW) <= 2 * tol and rel(b_t.grad.cpu().numpy(), dbias) <= 2 * tol
    assert rel(words_t.grad.cpu().numpy(), dwords_full) <= 2 * tol
    assert rel(sent_t.grad.cpu().numpy(), dsent) <= 1e-5 and rel(img_t.grad.cpu().numpy(), os_["dimg"]) <= 1e-5
    assert float(words_t.grad[:, 0].abs().max()) == 0.0            # <sos> rows receive no gradient
