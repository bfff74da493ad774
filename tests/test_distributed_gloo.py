"""Multi-rank host logic on CPU: world_size-2 gloo, caption rows sharded, every rank's images
serve as negatives.  The CUDA engine is replaced by the oracle-backed checker engine, so what is tested
is exactly ops.py's exchange steps: all-gather of vhat / sentence codes / class ids, column log-sum-exp
combine, loss all-reduce, reduce-scatter of the image-side gradient partials."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import damsm_oracle as O

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, port, cls, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from checker_engine import CheckerEngine
        pkg = importlib.import_module("t2i_clip-gan_b200")
        B, T, R = 8, 7, 9
        x = O.make_inputs(B, T, R, seed=42, class_ids=cls, n_classes=3)
        bl = B // WORLD
        lo, hi = rank * bl, (rank + 1) * bl
        w = torch.tensor(x["words"][lo:hi], dtype=torch.float64, requires_grad=True)
        r = torch.tensor(x["regions"][lo:hi], dtype=torch.float64, requires_grad=True)
        img = torch.tensor(x["img"][lo:hi], dtype=torch.float64, requires_grad=True)
        txt = torch.tensor(x["sent"][lo:hi], dtype=torch.float64, requires_grad=True)
        cid = None if x["class_ids"] is None else x["class_ids"][lo:hi]
        eng = CheckerEngine()
        l0, l1, _ = pkg.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), torch.arange(B), None, cid, bl,
                                   torch.tensor(x["mask"][lo:hi]), 4.0, 5.0, 10.0, group=dist.group.WORLD, engine=eng)
        s0, s1 = pkg.ops.DamsmSentLoss.apply(img, txt, torch.arange(B),
                                             None if cid is None else torch.tensor(cid), 10.0, 1e-8, eng,
                                             dist.group.WORLD)
        (0.7 * l0 + 1.3 * l1 + s0 + 2.0 * s1).backward()
        out[rank] = dict(l0=float(l0.detach()), l1=float(l1.detach()), s0=float(s0.detach()), s1=float(s1.detach()), dw=w.grad.numpy(), dr=r.grad.numpy(),
                         di=img.grad.numpy(), dt=txt.grad.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cls", [False, True])
def test_row_sharded_losses_match_single_process(cls):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(_free_port(), cls, out), nprocs=WORLD, join=True)
    B, T, R = 8, 7, 9
    x = O.make_inputs(B, T, R, seed=42, class_ids=cls, n_classes=3)
    ow = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], 4.0, 5.0, 10.0, g0=0.7, g1=1.3)
    os_ = O.sent_loss(x["img"], x["sent"], x["labels"], x["class_ids"], 10.0, g0=1.0, g1=2.0)
    bl = B // WORLD
    for rank in range(WORLD):
        o = out[rank]
        lo, hi = rank * bl, (rank + 1) * bl
        assert abs(o["l0"] - ow["loss0"]) < 1e-9 and abs(o["l1"] - ow["loss1"]) < 1e-9     # identical on every rank
        assert abs(o["s0"] - os_["loss0"]) < 1e-9 and abs(o["s1"] - os_["loss1"]) < 1e-9
        np.testing.assert_allclose(o["dw"], ow["dwords"][lo:hi], rtol=1e-6, atol=1e-10)
        np.testing.assert_allclose(o["dr"], ow["dregions"][lo:hi], rtol=1e-6, atol=1e-10)
        np.testing.assert_allclose(o["di"], os_["dimg"][lo:hi], rtol=1e-6, atol=1e-10)
        np.testing.assert_allclose(o["dt"], os_["dtxt"][lo:hi], rtol=1e-6, atol=1e-10)


def test_checker_engine_single_process_matches_oracle():
    """Sanity of the checker engine itself (no process group)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from checker_engine import CheckerEngine
    pkg = importlib.import_module("t2i_clip-gan_b200")
    x = O.make_inputs(5, 6, 9, seed=3, class_ids=True, n_classes=2)
    o = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], 4.0, 5.0, 10.0)
    w = torch.tensor(x["words"], dtype=torch.float64, requires_grad=True)
    r = torch.tensor(x["regions"], dtype=torch.float64, requires_grad=True)
    l0, l1, _ = pkg.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), torch.arange(5), None, x["class_ids"], 5,
                               torch.tensor(x["mask"]), 4.0, 5.0, 10.0, engine=CheckerEngine())
    (l0 + l1).backward()
    assert abs(float(l0) - o["loss0"]) < 1e-9 and abs(float(l1) - o["loss1"]) < 1e-9
    np.testing.assert_allclose(w.grad.numpy(), o["dwords"], rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(r.grad.numpy(), o["dregions"], rtol=1e-6, atol=1e-10)
