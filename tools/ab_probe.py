"""Development probe: same-box A/B timing of library variants (make OUT=tools/_ab/libX.so OBJDIR=build/X EXTRA=-D...).

usage: python tools/ab_probe.py B reps lib1.so [lib2.so ...]   (each library runs in its own subprocess, round-robin
       `rounds` times; caption lengths U[T/3, T] like bench.py's c5 workload)"""
import importlib, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(b, reps, lib):
    sys.path.insert(0, ROOT)
    import torch
    pkg = importlib.import_module("t2i_clip-gan_b200")
    libmod = importlib.import_module("t2i_clip-gan_b200._lib")
    libmod.LIB_PATH = os.path.abspath(lib)
    eng = pkg.get_engine("bf16")
    T, R, D = 77, 196, 512
    g = torch.Generator(device="cuda").manual_seed(0)
    w = torch.randn(b, T, D, device="cuda", generator=g)
    r = torch.randn(b, R, D, device="cuda", generator=g)
    lens = torch.randint(T // 3, T + 1, (b,), device="cuda", generator=g)
    if os.environ.get("AB_LEN"):                              # every caption has this many words
        lens = torch.full((b,), int(os.environ["AB_LEN"]), device="cuda")
    m = (torch.arange(T, device="cuda")[None, :] < lens[:, None]).to(torch.uint8).contiguous()
    qhat, qhat16, _, qun = eng.l2norm_fwd(w, want_bf16=True, pad8=True)
    vhat, vhat16, _, _ = eng.l2norm_fwd(r, want_bf16=True)
    col = eng.words_prepare_columns(vhat, vhat16)
    gam = (4.0, 5.0, 10.0)
    sim = eng.words_fwd(qhat, qhat16, vhat, col, qun, m, gam)
    row_lse, cmax, csum = eng.ce_stats(sim, None, None, 0)
    col_lse = torch.log(csum) + cmax
    gs = torch.ones(2, device="cuda")

    def timed(fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, out
    tf, _ = timed(lambda: eng.words_fwd(qhat, qhat16, vhat, col, qun, m, gam))
    tb, out = timed(lambda: eng.words_bwd(qhat, qhat16, vhat, col, qun, m, sim, row_lse, col_lse, None, gs, 0, b, gam))
    chk = [float(x.double().abs().sum()) for x in out if x is not None]
    k = 1e-3 * 1.9e9 * 148 / (b * b)
    print(f"[{tf * k:6.0f} / {tb * k:6.0f} clk per pair @1.9GHz] ", end="")
    print(f"{os.path.basename(lib):28s} B={b}: fwd {tf:8.3f} ms  bwd {tb:8.3f} ms  sum {tf + tb:8.3f}  "
          f"sim {float(sim.double().sum()):.6f} chk {' '.join(f'{c:.6e}' for c in chk)}", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
    else:
        b, reps, libs = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3:]
        rounds = int(os.environ.get("AB_ROUNDS", "2"))
        for _ in range(rounds):
            for lib in libs:
                subprocess.run([sys.executable, __file__, "--child", str(b), str(reps), lib], check=False)
