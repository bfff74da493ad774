"""bf16 tensor-core (tcgen05) path against the fp64 oracle fed the same bf16-rounded inputs.
Tolerance (BASELINE.json north_star): rel <= 2e-3 on losses and gradients."""
import importlib

import numpy as np
import pytest
import torch

from oracle import damsm_oracle as O

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("t2i_clip-gan_b200")
TOL = 2e-3


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rounded(x):
    xb = dict(x)
    for k in ("words", "regions"):
        xb[k] = torch.tensor(x[k]).bfloat16().float().numpy()
    return xb


@pytest.mark.parametrize("B,T,R,seed", [(6, 5, 9, 1), (12, 18, 49, 2), (9, 28, 49, 3), (5, 77, 49, 4), (4, 77, 196, 5),
                                        (7, 40, 120, 6), (3, 64, 127, 7), (20, 18, 196, 8), (4, 30, 250, 9),
                                        (3, 128, 64, 10)])
def test_tc_forward_scores(B, T, R, seed):
    """sim matrix (before masking / CE) of the tcgen05 kernel vs the exact fp32 kernel and the oracle."""
    x = rounded(O.make_inputs(B, T, R, seed=seed, class_ids=False))
    ref = 10.0 * O.words_sim(x["words"], x["regions"], x["mask"], 4.0, 5.0)
    outs = {}
    for prec in ("fp32", "bf16"):
        eng = pkg.get_engine(prec)
        w = torch.tensor(x["words"], device="cuda")
        r = torch.tensor(x["regions"], device="cuda")
        qhat, qhat16, _, qun = eng.l2norm_fwd(w, want_bf16=prec == "bf16", pad8=True)
        vhat, vhat16, _, _ = eng.l2norm_fwd(r, want_bf16=prec == "bf16")
        col = eng.words_prepare_columns(vhat, vhat16)
        m = torch.tensor(x["mask"], device="cuda").to(torch.uint8)
        outs[prec] = eng.words_fwd(qhat, qhat16, vhat, col, qun, m, (4.0, 5.0, 10.0)).cpu().numpy()
    assert np.abs(outs["fp32"] - ref).max() <= 1e-4
    err = np.abs(outs["bf16"] - ref).max() / np.abs(ref).max()
    assert err <= TOL, err


@pytest.mark.parametrize("B,T,R,cls,seed", [(8, 18, 49, True, 11), (6, 77, 196, False, 12), (16, 28, 49, True, 13),
                                            (5, 18, 250, False, 14),      # R+1 > 224: 18-warp kernel variant
                                            (4, 100, 49, True, 15),       # T > 80: NT = 128
                                            (3, 7, 16, False, 16)])
def test_tc_words_loss_and_grads(B, T, R, cls, seed):
    x = rounded(O.make_inputs(B, T, R, seed=seed, class_ids=cls, n_classes=4))
    o = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], 4.0, 5.0, 10.0)
    w = torch.tensor(x["words"], device="cuda").requires_grad_(True)
    r = torch.tensor(x["regions"], device="cuda").requires_grad_(True)
    l0, l1, _ = pkg.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), torch.arange(B, device="cuda"), None,
                               x["class_ids"], B, torch.tensor(x["mask"]), 4.0, 5.0, 10.0, precision="bf16")
    (l0 + l1).backward()
    assert abs(l0.item() - o["loss0"]) <= TOL * max(1, abs(o["loss0"]))
    assert abs(l1.item() - o["loss1"]) <= TOL * max(1, abs(o["loss1"]))
    assert rel(w.grad.cpu().numpy(), o["dwords"]) <= TOL
    assert rel(r.grad.cpu().numpy(), o["dregions"]) <= TOL


def test_tc_multi_chunk_backward_matches_exact_path():
    """Larger batch, backward forced through several workspace chunks and several image splits per caption:
    tensor-core gradients against the exact fp32 CUDA path (itself pinned to the oracle) on the same inputs."""
    B, T, R = 96, 77, 196
    x = rounded(O.make_inputs(B, T, R, seed=21, class_ids=True, n_classes=7))
    eng = pkg.get_engine("bf16")
    lib = pkg._lib.load()
    old = eng.tc_workspace_bytes
    # room for 20 full-length captions (80 scratch columns each) per chunk -> several chunks for the 96 captions
    eng.tc_workspace_bytes = lib.damsm_words_bwd_tc_fixed_bytes() + 20 * 80 * lib.damsm_words_bwd_tc_col_bytes(B, R)
    try:
        res = {}
        for prec in ("fp32", "bf16"):
            w = torch.tensor(x["words"], device="cuda").requires_grad_(True)
            r = torch.tensor(x["regions"], device="cuda").requires_grad_(True)
            l0, l1, _ = pkg.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), torch.arange(B, device="cuda"), None,
                                       x["class_ids"], B, torch.tensor(x["mask"]), 4.0, 5.0, 10.0, precision=prec)
            (l0 + 0.5 * l1).backward()
            res[prec] = (l0.item(), l1.item(), w.grad.cpu().numpy(), r.grad.cpu().numpy())
    finally:
        eng.tc_workspace_bytes = old
    a, b = res["bf16"], res["fp32"]
    assert abs(a[0] - b[0]) <= TOL * max(1, abs(b[0])) and abs(a[1] - b[1]) <= TOL * max(1, abs(b[1]))
    assert rel(a[2], b[2]) <= TOL
    assert rel(a[3], b[3]) <= TOL


@pytest.mark.parametrize("side", ["regions", "words"])
def test_tc_one_sided_gradients(side):
    """Only one input requires grad (image side only = the DM-GAN generator step, trainer.py:338 / SURVEY 8f-4): the
    other side's GEMMs are skipped and the remaining gradient is unchanged."""
    B, T, R = 6, 77, 49
    x = rounded(O.make_inputs(B, T, R, seed=31, class_ids=True, n_classes=3))
    o = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], 4.0, 5.0, 10.0)
    w = torch.tensor(x["words"], device="cuda").requires_grad_(side == "words")
    r = torch.tensor(x["regions"], device="cuda").requires_grad_(side == "regions")
    before = pkg._lib.launch_count()
    l0, l1, _ = pkg.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), torch.arange(B, device="cuda"), None,
                               x["class_ids"], B, torch.tensor(x["mask"]), 4.0, 5.0, 10.0, precision="bf16")
    (l0 + l1).backward()
    assert pkg._lib.launch_count() > before
    if side == "regions":
        assert w.grad is None and rel(r.grad.cpu().numpy(), o["dregions"]) <= TOL
    else:
        assert r.grad is None and rel(w.grad.cpu().numpy(), o["dwords"]) <= TOL


def test_tc_unsupported_shape_falls_back_to_exact_cuda_path():
    """T=100 with R=100 does not fit the tensor-core kernel's shared memory (128-word tile resident): precision='bf16'
    then runs the exact fp32 CUDA kernels (a warning says so) -- still on the GPU, never on the CPU."""
    B, T, R = 3, 100, 100
    assert not pkg.ops.tc_shape_supported(T, R, 512) and pkg.ops.tc_shape_supported(77, R, 512)
    x = rounded(O.make_inputs(B, T, R, seed=41, class_ids=False))
    o = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], None, 4.0, 5.0, 10.0)
    w = torch.tensor(x["words"], device="cuda").requires_grad_(True)
    r = torch.tensor(x["regions"], device="cuda").requires_grad_(True)
    with pytest.warns(RuntimeWarning, match="exact fp32 CUDA path"):
        l0, l1, _ = pkg.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), torch.arange(B, device="cuda"), None, None, B,
                                   torch.tensor(x["mask"]), 4.0, 5.0, 10.0, precision="bf16")
    (l0 + l1).backward()
    assert abs(l0.item() - o["loss0"]) <= 1e-5 * max(1, abs(o["loss0"]))
    assert rel(w.grad.cpu().numpy(), o["dwords"]) <= 1e-5 and rel(r.grad.cpu().numpy(), o["dregions"]) <= 1e-5


@pytest.mark.parametrize("mode", ["fwd_clusters", "bwd_clusters_too", "no_clusters"])
def test_tc_cluster_modes_agree(mode, monkeypatch):
    """2-CTA clusters sharing the image stream by TMA multicast (default: forward only; DAMSM_TC_CLUSTER_BWD=1 adds the
    backward; DAMSM_TC_NO_CLUSTER=1 disables them): every mode meets the parity bar on the same inputs."""
    if mode == "bwd_clusters_too":
        monkeypatch.setenv("DAMSM_TC_CLUSTER_BWD", "1")
    if mode == "no_clusters":
        monkeypatch.setenv("DAMSM_TC_NO_CLUSTER", "1")
    B, T, R = 10, 77, 196
    x = rounded(O.make_inputs(B, T, R, seed=51, class_ids=True, n_classes=4))
    o = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], 4.0, 5.0, 10.0)
    w = torch.tensor(x["words"], device="cuda").requires_grad_(True)
    r = torch.tensor(x["regions"], device="cuda").requires_grad_(True)
    l0, l1, _ = pkg.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), torch.arange(B, device="cuda"), None,
                               x["class_ids"], B, torch.tensor(x["mask"]), 4.0, 5.0, 10.0, precision="bf16")
    (l0 + l1).backward()
    assert abs(l0.item() - o["loss0"]) <= TOL * max(1, abs(o["loss0"]))
    assert abs(l1.item() - o["loss1"]) <= TOL * max(1, abs(o["loss1"]))
    assert rel(w.grad.cpu().numpy(), o["dwords"]) <= TOL and rel(r.grad.cpu().numpy(), o["dregions"]) <= TOL


def test_full_size_properties_c4():
    """BASELINE configs[3] at full size (B1024 T77 R196, bf16 inputs, tensor-core path) -- far beyond what the oracle can
    run, so: (i) a sample of caption rows of the score matrix against the exact fp32 CUDA kernel (itself pinned to the
    oracle) at the full column width; (ii) permuting the batch leaves the losses unchanged and permutes the gradients;
    (iii) l2norm is scale invariant, so every region's gradient is orthogonal to the region vector."""
    B, T, R, D = 1024, 77, 196, 512
    g = torch.Generator().manual_seed(2029)
    s = torch.randn(B, 1, D, generator=g)
    # bf16-rounded values held as fp32 tensors, so that the returned gradients are fp32 (bf16 gradients would carry
    # a 2^-8 rounding of their own, above the bar checked in (ii))
    words = (0.25 * s + torch.randn(B, T, D, generator=g)).bfloat16().float()
    regions = (0.25 * s + torch.randn(B, R, D, generator=g)).bfloat16().float()
    cap_len = torch.randint(T // 3, T + 1, (B,), generator=g)
    mask = (torch.arange(T).reshape(1, T) < cap_len.reshape(B, 1)).to(torch.int64)
    labels = torch.arange(B, device="cuda")

    def run(w_h, r_h, m_h):
        w = w_h.cuda().requires_grad_(True)
        r = r_h.cuda().requires_grad_(True)
        l0, l1, _ = pkg.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), labels, None, None, B, m_h, 4.0, 5.0, 10.0,
                                   precision="bf16")
        (l0 + l1).backward()
        return l0.item(), l1.item(), w.grad.float(), r.grad.float()

    l0, l1, dw, dr = run(words, regions, mask)
    assert np.isfinite(l0) and np.isfinite(l1) and 0.0 < l0 < np.log(B) + 1e-3 and 0.0 < l1 < np.log(B) + 1e-3
    # (i) sampled rows of sim: tensor-core kernel vs exact fp32 kernel, all 1024 columns
    rows = torch.tensor([0, 1, 17, 511, 512, 1000, 1022, 1023])
    m8 = mask.cuda().to(torch.uint8)
    sims = {}
    for prec in ("fp32", "bf16"):
        eng = pkg.get_engine(prec)
        qhat, qhat16, _, qun = eng.l2norm_fwd(words[rows].cuda(), want_bf16=prec == "bf16", pad8=True)
        vhat, vhat16, _, _ = eng.l2norm_fwd(regions.cuda(), want_bf16=prec == "bf16")
        col = eng.words_prepare_columns(vhat, vhat16)
        sims[prec] = eng.words_fwd(qhat, qhat16, vhat, col, qun, m8[rows.cuda()], (4.0, 5.0, 10.0)).float().cpu().numpy()
    assert np.abs(sims["bf16"] - sims["fp32"]).max() <= TOL * np.abs(sims["fp32"]).max()
    # (ii) permutation of the batch
    perm = torch.randperm(B, generator=g)
    p0, p1, dwp, drp = run(words[perm], regions[perm], mask[perm])
    assert abs(p0 - l0) <= 1e-4 * max(1.0, abs(l0)) and abs(p1 - l1) <= 1e-4 * max(1.0, abs(l1))
    pc = perm.cuda()
    assert float((dwp - dw[pc]).abs().max()) <= TOL * float(dw.abs().max())
    assert float((drp - dr[pc]).abs().max()) <= TOL * float(dr.abs().max())
    # (iii) radial component of the region gradient
    radial = (dr * regions.cuda().float()).sum(-1)
    assert float(radial.abs().max()) <= 2e-2 * float(dr.abs().max()) * float(regions.float().abs().max()) * D ** 0.5
