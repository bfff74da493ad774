"""Development probe: SM clock and board power while the forward / fused-backward kernels run for seconds
(debug build: DAMSM_DBG ablations), to see which part of the per-pair work drives the power cap."""
import importlib, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("t2i_clip-gan_b200")
eng = pkg.get_engine("bf16")
D, T, R = 512, 77, 196
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
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
fwd = lambda: eng.words_fwd(qhat, qhat16, vhat, col, qun, m, (4.0, 5.0, 10.0))
bwd = lambda: eng.words_bwd(qhat, qhat16, vhat, col, qun, m, sim, row_lse, col_lse, None, gs, 0, B, (4.0, 5.0, 10.0))


def run(name, fn, dbg, secs=4.0):
    os.environ["DAMSM_DBG"] = str(dbg)
    fn(); torch.cuda.synchronize()
    f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                          "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); n = 0
    e0.record()
    while time.perf_counter() - t0 < secs:
        fn(); n += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    p.terminate(); p.wait()
    f.flush(); f.seek(0)
    rows = [l.split(",") for l in f.read().splitlines() if "," in l]
    rows = rows[len(rows) // 2:]                      # second half: after the cap has settled
    clk = sorted(float(a) for a, _ in rows)[len(rows) // 2]
    pw = sorted(float(b) for _, b in rows)[len(rows) // 2]
    ms = e0.elapsed_time(e1) / n
    print(f"{name:9s} DBG={dbg:3d}: {ms:8.2f} ms/launch  clk {clk:6.0f} MHz  power {pw:6.0f} W  -> {ms * 1e-3 * clk * 1e6 * 148 / (B * B):7.0f} clk/pair", flush=True)
    os.unlink(f.name)


for dbg in (0, 64, 32, 24):
    run("forward", fwd, dbg)
os.environ["DAMSM_BWD_FUSED_ONLY"] = "1"
for dbg in (0, 1, 4, 64):
    run("bwd-fused", bwd, dbg)
