"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL): the row-sharded losses and gradients must
equal the single-GPU ones computed on the full batch by the same kernels."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from oracle import damsm_oracle as O

pkg = importlib.import_module("t2i_clip-gan_b200")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = dist.group.WORLD
ok = True
for prec, (B, T, R), tol in (("fp32", (16, 18, 49), 2e-5), ("bf16", (32, 77, 196), 2e-3)):
    x = O.make_inputs(B, T, R, seed=5, class_ids=True, n_classes=5)
    bl = B // world
    lo, hi = rank * bl, (rank + 1) * bl

    def run(sl, group, labels_n):
        w = torch.tensor(x["words"][sl], device="cuda").requires_grad_(True)
        r = torch.tensor(x["regions"][sl], device="cuda").requires_grad_(True)
        a = torch.tensor(x["img"][sl], device="cuda").requires_grad_(True)
        t = torch.tensor(x["sent"][sl], device="cuda").requires_grad_(True)
        n = w.shape[0]
        l0, l1, _ = pkg.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), torch.arange(labels_n, device="cuda"), None,
                                   x["class_ids"][sl], n, torch.tensor(x["mask"][sl]), 4.0, 5.0, 10.0, precision=prec,
                                   group=group)
        s0, s1 = pkg.sent_loss(a, t, torch.arange(labels_n, device="cuda"), x["class_ids"][sl], n, group=group)
        (l0 + l1 + s0 + s1).backward()
        return [v.item() for v in (l0, l1, s0, s1)], [v.grad.cpu().numpy() for v in (w, r, a, t)]

    ls, gs = run(slice(lo, hi), g, B)
    lf, gf = run(slice(0, B), None, B)
    el = max(abs(a - b) / max(1, abs(b)) for a, b in zip(ls, lf))
    eg = max(float(np.abs(a - b[lo:hi]).max() / np.abs(b).max()) for a, b in zip(gs, gf))
    print(f"rank {rank} {prec} B={B}: loss err {el:.2e} grad err {eg:.2e}", flush=True)
    ok = ok and el <= tol and eg <= tol
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
