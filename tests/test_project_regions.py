"""Region projection fused with the l2norm prologue (model.py:46,78 + pretrain_DAMSM.py:125).
The reference op is a stock nn.Linear followed by a slice, so the oracle is pinned against torch itself on the CPU;
the tcgen05 kernel is held to the tensor-core tolerance of the path (rel <= 2e-3)."""
import importlib

import numpy as np
import pytest
import torch

from oracle import damsm_oracle as O

TOL = 2e-3


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def inputs(B, R, K, N, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, R + 1, K)).astype(np.float32)
    w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = (0.1 * rng.standard_normal(N)).astype(np.float32)
    return x, w, b


def test_oracle_matches_torch_linear_and_autograd():
    x, w, b = inputs(3, 5, 24, 16, 0)
    dy = np.random.default_rng(1).standard_normal((3, 5, 16))
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    wt = torch.tensor(w, dtype=torch.float64, requires_grad=True)
    bt = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    # the reference's expression: linear over the flattened tokens, view back, drop CLS (model.py:46, pretrain_DAMSM.py:125)
    yt = torch.nn.functional.linear(xt.view(-1, 24), wt, bt).view(3, -1, 16)[:, 1:, :]
    yt.backward(torch.tensor(dy))
    y, dx, dw, db = O.project_regions(x, w, b, dy)
    assert rel(y, yt.detach().numpy()) < 1e-12 and rel(dx, xt.grad.numpy()) < 1e-12
    assert rel(dw, wt.grad.numpy()) < 1e-12 and rel(db, bt.grad.numpy()) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("B,R,K,N,dtype", [(4, 49, 768, 512, torch.float32), (3, 196, 768, 512, torch.bfloat16),
                                           (5, 9, 40, 64, torch.float32), (2, 16, 72, 272, torch.bfloat16),
                                           (1, 300, 768, 512, torch.float32)])
def test_cuda_forward_backward_vs_oracle(B, R, K, N, dtype):
    pkg = importlib.import_module("t2i_clip-gan_b200")
    x, w, b = inputs(B, R, K, N, B * 100 + R)
    xt = torch.tensor(x, device="cuda").to(dtype).requires_grad_(True)
    wt = torch.tensor(w, device="cuda").to(dtype).requires_grad_(True)
    bt = torch.tensor(b, device="cuda").requires_grad_(True)
    feats = pkg.project_regions(xt, wt, bt)
    assert tuple(feats.shape) == (B, N, R) and feats.dtype == torch.float32
    dy = np.random.default_rng(7).standard_normal((B, R, N)).astype(np.float32)
    feats.backward(torch.tensor(dy, device="cuda").permute(0, 2, 1))
    y, dx, dw, db = O.project_regions(xt.detach().float().cpu().numpy(), wt.detach().float().cpu().numpy(), b, dy)
    tol = TOL if dtype == torch.float32 else 1e-4           # bf16 products are exact in the fp32 accumulator
    assert rel(feats.detach().permute(0, 2, 1).cpu().numpy(), y) <= tol
    gtol = TOL if dtype == torch.float32 else 1e-2          # bf16 gradients are rounded to bf16 on return
    assert rel(xt.grad.float().cpu().numpy(), dx) <= gtol and float(xt.grad[:, 0].abs().max()) == 0.0
    assert rel(wt.grad.float().cpu().numpy(), dw) <= gtol
    assert rel(bt.grad.cpu().numpy(), db) <= TOL


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_words_loss_reuses_the_epilogue_outputs(precision):
    """project_regions -> words_loss: the normalised copies come from the GEMM epilogue (cache hit); the result equals
    words_loss on a detached copy of the same features (cache miss -> own l2norm pass), and gradients reach W."""
    pkg = importlib.import_module("t2i_clip-gan_b200")
    B, T, R = 6, 18, 49
    xin = O.make_inputs(B, T, R, seed=5, class_ids=False)
    x, w, b = inputs(B, R, 768, 512, 3)
    xt = torch.tensor(x, device="cuda")
    wt = torch.tensor(w, device="cuda").requires_grad_(True)
    bt = torch.tensor(b, device="cuda")
    words = torch.tensor(xin["words"], device="cuda").permute(0, 2, 1)
    mask = torch.tensor(xin["mask"])
    lab = torch.arange(B, device="cuda")
    feats = pkg.project_regions(xt, wt, bt)
    assert pkg.ops._prologue_get(feats.permute(0, 2, 1)) is not None
    l0, l1, _ = pkg.words_loss(feats, words, lab, None, None, B, mask, 4.0, 5.0, 10.0, precision=precision)
    (l0 + l1).backward()
    assert wt.grad is not None and float(wt.grad.abs().max()) > 0
    copy = feats.detach().clone()
    assert pkg.ops._prologue_get(copy.permute(0, 2, 1)) is None
    m0, m1, _ = pkg.words_loss(copy, words, lab, None, None, B, mask, 4.0, 5.0, 10.0, precision=precision)
    assert abs(l0.item() - m0.item()) <= 2e-5 and abs(l1.item() - m1.item()) <= 2e-5
    # the loss agrees with the oracle fed the oracle's own projection, at tensor-core tolerance
    y = O.project_regions(x, w, b)
    o = O.words_loss(xin["words"], y, xin["mask"], xin["labels"], None, 4.0, 5.0, 10.0)
    assert abs(l0.item() - o["loss0"]) <= 5e-3 * max(1, abs(o["loss0"]))
    # in-place modification invalidates the cached copies
    feats2 = pkg.project_regions(xt, wt, bt)
    base = feats2.permute(0, 2, 1)
    with torch.no_grad():
        base.mul_(2.0)
    assert pkg.ops._prologue_get(base) is None
