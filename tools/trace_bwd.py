"""Development helper (debug build, make EXTRA=-DDAMSM_TC_DEBUG): per-pair clock trace of the backward pair kernel."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("t2i_clip-gan_b200")
if os.environ.get("DAMSM_AB_LIB"):
    importlib.import_module("t2i_clip-gan_b200._lib").LIB_PATH = os.path.abspath(os.environ["DAMSM_AB_LIB"])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
L = int(sys.argv[2]) if len(sys.argv) > 2 else 77          # every caption has this many words
T, R, D = 77, 196, 512
g = torch.Generator(device="cuda").manual_seed(0)
w = torch.randn(B, T, D, device="cuda", generator=g).bfloat16().requires_grad_(True)
r = torch.randn(B, R, D, device="cuda", generator=g).bfloat16().requires_grad_(True)
m = (torch.arange(T, device="cuda")[None, :] < L).to(torch.int64).expand(B, T).contiguous()
l0, l1, _ = pkg.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), torch.arange(B, device="cuda"), None, None, B, m,
                           4.0, 5.0, 10.0, precision="bf16")
torch.cuda.synchronize()
os.environ["DAMSM_TRACE_BWD"] = "1"
(l0 + l1).backward()
torch.cuda.synchronize()
