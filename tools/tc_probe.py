"""Development probe: time the tcgen05 forward (and backward when present) at large batch."""
import importlib, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("t2i_clip-gan_b200")
eng = pkg.get_engine("bf16")
D = 512
for (B, T, R) in [(256, 77, 196), (1024, 77, 196), (1024, 18, 49), (4096, 77, 196)]:
    g = torch.Generator(device="cuda").manual_seed(0)
    w = torch.randn(B, T, D, device="cuda", generator=g)
    r = torch.randn(B, R, D, device="cuda", generator=g)
    m = torch.ones(B, T, dtype=torch.uint8, device="cuda")
    qhat, qhat16, _, qun = eng.l2norm_fwd(w, want_bf16=True)
    vhat, vhat16, _, _ = eng.l2norm_fwd(r, want_bf16=True)
    col = eng.words_prepare_columns(vhat, vhat16)
    torch.cuda.synchronize()
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sim = eng.words_fwd(qhat, qhat16, vhat, col, qun, m, (4.0, 5.0, 10.0))
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    fl = 4.0 * B * B * T * R * D
    print(f"B={B} T={T} R={R}: fwd {ms:.2f} ms, algorithmic {fl/ms/1e9:.1f} TFLOP/s ({fl/ms/1e9/1654.2*100:.1f}% of 1654), "
          f"per pair-tile {ms*1e-3*1.9e9*148/(B*B):.0f} clk@1.9GHz  simsum={sim.sum().item():.3f}")
