import importlib, sys, os, faulthandler
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(30, exit=True)
import numpy as np, torch
from oracle import damsm_oracle as O
pkg = importlib.import_module("t2i_clip-gan_b200")
eng = pkg.get_engine("bf16")
B, T, R = 8, 18, 49
x = O.make_inputs(B, T, R, seed=11, class_ids=False)
w = torch.tensor(x["words"], device="cuda"); r = torch.tensor(x["regions"], device="cuda")
qhat, qhat16, _, qun = eng.l2norm_fwd(w, want_bf16=True, pad8=True)
vhat, vhat16, _, _ = eng.l2norm_fwd(r, want_bf16=True)
col = eng.words_prepare_columns(vhat, vhat16)
m = torch.tensor(x["mask"], device="cuda").to(torch.uint8)
sim = eng.words_fwd(qhat, qhat16, vhat, col, qun, m, (4.0, 5.0, 10.0))
row_lse, cmax, csum = eng.ce_stats(sim, None, None, 0)
col_lse = torch.log(csum) + cmax
torch.cuda.synchronize(); print("fwd done", flush=True)
gs = torch.ones(2, device="cuda")
out = eng.words_bwd(qhat, qhat16, vhat, col, qun, m, sim, row_lse, col_lse, None, gs, 0, B, (4.0, 5.0, 10.0))
torch.cuda.synchronize(); print("bwd done", [float(o.abs().sum()) for o in out], flush=True)
