"""Development probe: which part of the tensor-core backward is off in the C5-regime row block (br=64, bc=4096)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("t2i_clip-gan_b200")
eng = pkg.get_engine("bf16")

def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

# (a) the GEMM alone at the row count of the dvhat contraction
for (m, n, k) in ((802816, 512, 3840), (524288 + 256, 512, 256), (262144 + 256, 512, 256)):
    g = torch.Generator(device="cuda").manual_seed(1)
    a = (torch.randn(m, k, device="cuda", generator=g) * 0.1).half()
    b = (torch.randn(k, n, device="cuda", generator=g) * 0.1).half()
    c = torch.zeros(m, n, device="cuda")
    eng.gemm_tc(a, b, b_mn=True, out=c, accumulate=True)
    ref = (a.float() @ b.float())
    err = (c - ref).abs().amax(dim=1)
    bad = (err > 1e-3 * ref.abs().max()).nonzero().flatten()
    print(f"gemm M={m} N={n} K={k}: rel {relmax(c, ref):.2e}; bad rows {bad.numel()}", bad[:5].tolist(), bad[-5:].tolist() if bad.numel() else [])
    del a, b, c, ref

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_sizes as TS
BR, BC, T, R = 64, 4096, 77, 196
GAM = (4.0, 5.0, 10.0)
for lens in ("uniform", "full"):
    words, regions, mask = TS.synth(BC, T, R, seed=2030)
    if lens == "full":
        mask = torch.ones_like(mask)
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
    labels = torch.arange(BC, device="cuda")
    s = sim32.clone()
    row_lse, cmax, csum = f32.ce_stats(s, None, None, 0)
    col_lse = torch.log(csum * (BC / BR)) + cmax
    gscale = torch.tensor([1.0, 1.0], device="cuda")
    o32 = f32.words_bwd(qhat, qhat16, vhat, col32, qun, mask_u8, s, row_lse, col_lse, labels, gscale, 0, BC, GAM)
    s2 = simtc.clone()
    row_lse2, _, _ = tc.ce_stats(s2, None, None, 0)
    otc = tc.words_bwd(qhat, qhat16, vhat, coltc, qun, mask_u8, s2, row_lse2, col_lse, labels, gscale, 0, BC, GAM)
    for k, what in enumerate(("dqhat", "dvhat", "hmat", "kq")):
        print(lens, what, f"{relmax(otc[k], o32[k]):.2e}")
    e = (otc[1] - o32[1]).abs().amax(dim=(1, 2))
    ref = o32[1].abs().max()
    bad = (e > 2e-3 * ref).nonzero().flatten()
    print(lens, "images with a wrong dvhat:", bad.numel(), bad[:8].tolist(), bad[-8:].tolist() if bad.numel() else [])
    if bad.numel():
        j = int(bad[0])
        er = (otc[1][j] - o32[1][j]).abs().amax(dim=1)
        print("   rows of image", j, "max err per region (first 8):", (er[:8] / ref).tolist(), " ratio tc/f32 row0:", float(otc[1][j, 0].norm() / o32[1][j, 0].norm()))
