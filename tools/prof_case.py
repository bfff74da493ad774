"""Development helper: one forward+backward of the tensor-core path at a given batch (for ncu)."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("t2i_clip-gan_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T, R, D = 77, 196, 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.Generator(device="cuda").manual_seed(0)
w = torch.randn(B, T, D, device="cuda", generator=g).bfloat16().requires_grad_(True)
r = torch.randn(B, R, D, device="cuda", generator=g).bfloat16().requires_grad_(True)
m = torch.ones(B, T, dtype=torch.int64, device="cuda")
if len(sys.argv) > 3 and sys.argv[3] == "uniform":      # bench.py's caption lengths, U[T/3, T]
    cap = torch.randint(T // 3, T + 1, (B,), device="cuda", generator=g)
    m = (torch.arange(T, device="cuda")[None, :] < cap[:, None]).to(torch.int64)
for _ in range(reps):
    l0, l1, _ = pkg.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), torch.arange(B, device="cuda"), None, None, B, m,
                               4.0, 5.0, 10.0, precision="bf16")
    (l0 + l1).backward()
torch.cuda.synchronize()
print("ok", l0.item(), l1.item())
