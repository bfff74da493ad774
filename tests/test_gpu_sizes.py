"""Parity of the tensor-core (tcgen05) path AT THE SIZES THAT ARE BENCHMARKED.

The oracle (numpy fp64) is O(B^2 T R D) in Python and the reference cannot run B >= 256 at R = 196 (O(B^2 R D) retained
memory), so at these sizes the checker is the exact fp32 CUDA path, which tests/test_gpu_parity.py pins to the
oracle and the goldens at <= 1e-5.  Tolerance of the bf16-input path: rel <= 2e-3 (BASELINE.json north_star)."""
import importlib

import numpy as np
import pytest
import torch

from oracle import damsm_oracle as O

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("t2i_clip-gan_b200")
TOL = 2e-3
GAM = (4.0, 5.0, 10.0)


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def synth(B, T, R, seed, dtype=torch.float32, D=512):
    """bench.py's generator: latent-mixed words / regions, prefix masks U[T/3, T]; values bf16-rounded."""
    g = torch.Generator().manual_seed(seed)
    s = torch.randn(B, 1, D, generator=g)
    words = (0.25 * s + torch.randn(B, T, D, generator=g)).bfloat16()
    regions = (0.25 * s + torch.randn(B, R, D, generator=g)).bfloat16()
    cap = torch.randint(max(2, T // 3), T + 1, (B,), generator=g)
    mask = (torch.arange(T).reshape(1, T) < cap.reshape(B, 1)).to(torch.int64)
    return words.to(dtype), regions.to(dtype), mask


def run_words_loss(words, regions, mask, prec, weights=(1.0, 1.0), cls=None):
    w = words.cuda().requires_grad_(True)
    r = regions.cuda().requires_grad_(True)
    B = w.shape[0]
    l0, l1, _ = pkg.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), torch.arange(B, device="cuda"), None, cls, B,
                               mask, *GAM, precision=prec)
    (weights[0] * l0 + weights[1] * l1).backward()
    return l0.item(), l1.item(), w.grad.float(), r.grad.float()


def test_c4_full_size_losses_and_gradients_vs_exact_kernel():
    """BASELINE configs[3] at full size (B=1024, T=77, R=196): losses, dwords and dregions of the tcgen05 path against
    the exact fp32 CUDA path on the same bf16-rounded inputs."""
    words, regions, mask = synth(1024, 77, 196, seed=2029)
    ref = run_words_loss(words, regions, mask, "fp32")
    got = run_words_loss(words, regions, mask, "bf16")
    assert abs(got[0] - ref[0]) <= TOL * max(1.0, abs(ref[0])) and abs(got[1] - ref[1]) <= TOL * max(1.0, abs(ref[1]))
    ew, er = relmax(got[2], ref[2]), relmax(got[3], ref[3])
    print(f"C4 full size: loss {got[0]:.6f}/{ref[0]:.6f} {got[1]:.6f}/{ref[1]:.6f}  dwords rel {ew:.2e}  dregions rel {er:.2e}")
    assert ew <= TOL and er <= TOL
    # no systematic scale error either (a saturated / flushed fp16 scratch row would show here first)
    for g, h in ((got[2], ref[2]), (got[3], ref[3])):
        ratio = float((g.double() * h.double()).sum() / (h.double() * h.double()).sum())
        assert abs(ratio - 1.0) <= 1e-3, ratio


def test_c5_regime_row_block_vs_exact_engine():
    """The headline configuration's regime (BASELINE configs[4]: b_total = bc = 4096, T=77, R=196) on a block of 64
    caption rows: dqhat, dvhat - H vhat and kq of the tensor-core backward against the exact fp32 engine with the same
    sim / row_lse / col_lse -- the power-of-two fp16 scales of the scratch rows depend on b_total and are exercised
    here at their benchmarked value."""
    BR, BC, T, R = 64, 4096, 77, 196
    words, regions, mask = synth(BC, T, R, seed=2030)
    words, mask = words[:BR].cuda(), mask[:BR].cuda()
    regions = regions.cuda()
    mask_u8 = (mask != 0).to(torch.uint8).contiguous()
    f32, tc = pkg.get_engine("fp32"), pkg.get_engine("bf16")
    qhat, qhat16, _, qun = tc.l2norm_fwd(words, want_bf16=True, pad8=True)
    vhat, vhat16, _, _ = tc.l2norm_fwd(regions, want_bf16=True)
    gram = f32.gram(vhat)
    col32 = f32.pack_columns(gram, vhat, None)
    coltc = tc.pack_columns(gram, vhat, vhat16)
    sim32 = f32.words_fwd(qhat, None, vhat, col32, qun, mask_u8, GAM)
    simtc = tc.words_fwd(qhat, qhat16, vhat, coltc, qun, mask_u8, GAM)
    assert relmax(simtc, sim32) <= TOL
    labels = torch.arange(BC, device="cuda")
    outs = {}
    for name, eng, col, sim in (("fp32", f32, col32, sim32), ("bf16", tc, coltc, simtc)):
        s = sim.clone()
        row_lse, cmax, csum = eng.ce_stats(s, None, None, 0)
        # column statistics of the full 4096 x 4096 matrix are not available from 64 rows: extend the partial sums as
        # if every one of the 4096 rows contributed like the average of these 64 (same col_lse for both engines)
        col_lse = (torch.log(csum * (BC / BR)) + cmax) if name == "fp32" else outs["col_lse"]
        outs.setdefault("col_lse", col_lse)
        gscale = torch.tensor([1.0, 1.0], device="cuda")
        outs[name] = eng.words_bwd(qhat, qhat16, vhat, col, qun, mask_u8, s, row_lse, outs["col_lse"], labels, gscale,
                                   0, BC, GAM)
    # dvhat and hmat are compared through what they are for, dvhat - H vhat (ops.DamsmWordsLoss.backward): the
    # tensor-core path folds the skipped (padded) words' share of H vhat into dvhat instead of into H (pad_terms.cu)
    res = {}
    for name, eng in (("fp32", f32), ("bf16", tc)):
        dq, dv, hm, kq = outs[name]
        res[name] = (dq, eng.gram_bwd(hm, vhat, dv.clone()), kq)
    for k, what in enumerate(("dqhat", "dvhat - H vhat", "kq")):
        e = relmax(res["bf16"][k], res["fp32"][k])
        print(f"C5 regime row block: {what} rel {e:.2e}")
        assert e <= TOL, (what, e)


@pytest.mark.parametrize("scale", [1e4, 1e-4, 65536.0])
def test_tc_gradients_do_not_depend_on_the_loss_weight(scale):
    """the fp16 scratch rows are scaled relative to the upstream gradient, so a large loss weight
    (LAMBDA = 50, an AMP loss scale of 2^16) cannot saturate them and a tiny one cannot flush them to zero."""
    B, T, R = 24, 77, 196
    words, regions, mask = synth(B, T, R, seed=77)
    ref = run_words_loss(words, regions, mask, "fp32", weights=(1.0, 0.5))
    got = run_words_loss(words, regions, mask, "bf16", weights=(scale, 0.5 * scale))
    assert relmax(got[2] / scale, ref[2]) <= TOL
    assert relmax(got[3] / scale, ref[3]) <= TOL


def test_real_bf16_tensors_through_the_tensor_core_path():
    """torch.bfloat16 inputs (what bench.py times at C5) straight through precision='bf16': the gradients come back in
    bf16 (the inputs' dtype); compared with the oracle on the same values, tolerance 2e-3 + one bf16 rounding."""
    B, T, R = 12, 77, 196
    x = O.make_inputs(B, T, R, seed=91, class_ids=True, n_classes=4)
    wb = torch.tensor(x["words"]).bfloat16()
    rb = torch.tensor(x["regions"]).bfloat16()
    o = O.words_loss(wb.float().numpy(), rb.float().numpy(), x["mask"], x["labels"], x["class_ids"], *GAM)
    w = wb.cuda().requires_grad_(True)
    r = rb.cuda().requires_grad_(True)
    l0, l1, _ = pkg.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), torch.arange(B, device="cuda"), None,
                               x["class_ids"], B, torch.tensor(x["mask"]), *GAM, precision="bf16")
    (l0 + l1).backward()
    assert w.grad.dtype == torch.bfloat16 and r.grad.dtype == torch.bfloat16
    assert abs(l0.item() - o["loss0"]) <= TOL * max(1, abs(o["loss0"]))
    assert abs(l1.item() - o["loss1"]) <= TOL * max(1, abs(o["loss1"]))
    tol = TOL + 2.0 ** -8            # the returned gradient is rounded to bf16 once
    assert relmax(w.grad.float().cpu(), torch.tensor(o["dwords"])) <= tol
    assert relmax(r.grad.float().cpu(), torch.tensor(o["dregions"])) <= tol


@pytest.mark.parametrize("D", [20, 24, 36])
def test_exact_path_embedding_size_not_a_multiple_of_16(D):
    """the exact fp32 kernels stage K in chunks of 16; the tail chunk of D % 16 != 0 is zero-filled.
    (The reference works for any D.)"""
    B, T, R = 5, 9, 16
    x = O.make_inputs(B, T, R, D=D, seed=5 + D, class_ids=True, n_classes=3)
    o = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], *GAM)
    w = torch.tensor(x["words"], device="cuda").requires_grad_(True)
    r = torch.tensor(x["regions"], device="cuda").requires_grad_(True)
    l0, l1, _ = pkg.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), torch.arange(B, device="cuda"), None,
                               x["class_ids"], B, torch.tensor(x["mask"]), *GAM)
    (l0 + l1).backward()
    assert abs(l0.item() - o["loss0"]) <= 1e-5 * max(1, abs(o["loss0"]))
    assert abs(l1.item() - o["loss1"]) <= 1e-5 * max(1, abs(o["loss1"]))
    assert relmax(w.grad.cpu(), torch.tensor(o["dwords"])) <= 1e-5
    assert relmax(r.grad.cpu(), torch.tensor(o["dregions"])) <= 1e-5
    d_wc = np.random.default_rng(D).standard_normal((B, T, D)).astype(np.float32)
    fo_wc, _, fo_dq, fo_dc = O.func_attention(x["words"], x["regions"], 4.0, x["mask"], d_wc=d_wc)
    q = torch.tensor(x["words"], device="cuda").requires_grad_(True)
    c = torch.tensor(x["regions"], device="cuda").requires_grad_(True)
    wc, attn = pkg.func_attention(q.permute(0, 2, 1), c.permute(0, 2, 1), 4.0, torch.tensor(x["mask"]).unsqueeze(1))
    wc.backward(torch.tensor(d_wc, device="cuda"))
    assert relmax(wc.detach().cpu(), torch.tensor(fo_wc)) <= 1e-5
    assert relmax(q.grad.cpu(), torch.tensor(fo_dq)) <= 1e-5
    assert relmax(c.grad.cpu(), torch.tensor(fo_dc)) <= 1e-5
