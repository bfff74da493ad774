import importlib, sys, os, faulthandler
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(40, exit=True)
import numpy as np, torch
from oracle import damsm_oracle as O
pkg = importlib.import_module("t2i_clip-gan_b200")
B, T, R = 8, 18, 49
x = O.make_inputs(B, T, R, seed=11, class_ids=True, n_classes=4)
for k in ("words", "regions"):
    x[k] = torch.tensor(x[k]).bfloat16().float().numpy()
o = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], 4.0, 5.0, 10.0)
w = torch.tensor(x["words"], device="cuda").requires_grad_(True)
r = torch.tensor(x["regions"], device="cuda").requires_grad_(True)
print("fwd...", flush=True)
l0, l1, _ = pkg.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), torch.arange(B, device="cuda"), None,
                           x["class_ids"], B, torch.tensor(x["mask"]), 4.0, 5.0, 10.0, precision="bf16")
torch.cuda.synchronize()
print("fwd ok", l0.item(), l1.item(), o["loss0"], o["loss1"], flush=True)
(l0 + l1).backward()
torch.cuda.synchronize()
print("bwd ok", flush=True)
def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
print("dwords", rel(w.grad.cpu().numpy(), o["dwords"]), "dregions", rel(r.grad.cpu().numpy(), o["dregions"]))
