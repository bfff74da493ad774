"""Development probe: forward time per pair as a function of how many SMs are busy (shared-resource test)."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("t2i_clip-gan_b200")
eng = pkg.get_engine("bf16")
D, T, R, BC = 512, 77, 196, 1024
g = torch.Generator(device="cuda").manual_seed(0)
r = torch.randn(BC, R, D, device="cuda", generator=g)
vhat, vhat16, _, _ = eng.l2norm_fwd(r, want_bf16=True)
col = eng.words_prepare_columns(vhat, vhat16)
for br in (18, 37, 74, 111, 148, 296):
    w = torch.randn(br, T, D, device="cuda", generator=g)
    m = torch.ones(br, T, dtype=torch.uint8, device="cuda")
    qhat, qhat16, _, qun = eng.l2norm_fwd(w, want_bf16=True, pad8=True)
    fn = lambda: eng.words_fwd(qhat, qhat16, vhat, col, qun, m, (4.0, 5.0, 10.0))
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    waves = (br + 147) // 148
    print(f"rows {br:4d}: fwd {ms:.3f} ms  -> {ms * 1e-3 * 1.9e9 / (BC * waves):.0f} clk/pair per busy SM")
