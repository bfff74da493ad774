"""Development probe: sentence loss forward+backward, fused (one launch each way) vs unfused path, at several batch sizes."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("t2i_clip-gan_b200")
eng = pkg.get_engine("fp32")
for B in (48, 512, 1024, 4096):
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(B, 512, device="cuda", generator=g)
    t = torch.randn(B, 512, device="cuda", generator=g)
    lab = torch.arange(B, device="cuda")
    gs = torch.ones(2, device="cuda")
    def fused():
        lo, na, nb, rl, cm, cs = eng.sent_fwd(a, t, None, None, 0, 10.0, 1e-8)
        col = torch.log(cs) + cm
        eng.ce_losses(lo, rl, col, lab, 0, B)
        return eng.sent_bwd(a, t, na, nb, lo, rl, col, lab, gs, 0, B, 10.0, 1e-8)
    def unfused():
        lo, na, nb = eng.cos_logits(a, t, 10.0, 1e-8)
        rl, cm, cs = eng.ce_stats(lo, None, None, 0)
        col = torch.log(cs) + cm
        eng.ce_losses(lo, rl, col, lab, 0, B)
        return eng.cos_logits_bwd(a, t, na, nb, lo, rl, col, lab, gs, 0, B, 10.0, 1e-8)
    for name, fn in (("fused", fused), ("unfused", unfused)):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"B={B} {name}: {e0.elapsed_time(e1) / reps:.3f} ms", flush=True)
