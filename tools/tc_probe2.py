"""Development probe: time forward and backward (fused kernel + GEMMs) of the tensor-core path."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("t2i_clip-gan_b200")
eng = pkg.get_engine("bf16")
D = 512
B, T, R = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 77, 196
g = torch.Generator(device="cuda").manual_seed(0)
w = torch.randn(B, T, D, device="cuda", generator=g)
r = torch.randn(B, R, D, device="cuda", generator=g)
m = torch.ones(B, T, dtype=torch.uint8, device="cuda")
qhat, qhat16, _, qun = eng.l2norm_fwd(w, want_bf16=True, pad8=True)
vhat, vhat16, _, _ = eng.l2norm_fwd(r, want_bf16=True)
col = eng.words_prepare_columns(vhat, vhat16)
sim = eng.words_fwd(qhat, qhat16, vhat, col, qun, m, (4.0, 5.0, 10.0))
row_lse, cmax, csum = eng.ce_stats(sim, None, None, 0)
col_lse = torch.log(csum) + cmax
gs = torch.ones(2, device="cuda")
def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
tf = timed(lambda: eng.words_fwd(qhat, qhat16, vhat, col, qun, m, (4.0, 5.0, 10.0)))
tb = timed(lambda: eng.words_bwd(qhat, qhat16, vhat, col, qun, m, sim, row_lse, col_lse, None, gs, 0, B, (4.0, 5.0, 10.0)))
k = 1e-3 * 1.9e9 * 148 / (B * B)
print(f"DBG={os.environ.get('DAMSM_DBG','0')} B={B}: fwd {tf:.2f} ms ({tf*k:.0f} clk/pair)  bwd {tb:.2f} ms ({tb*k:.0f} clk/pair)")
