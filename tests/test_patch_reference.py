"""patch_reference(): the documented way to switch a reference checkout to the B200 kernels.
The reference's own ``miscc`` package and ``GlobalAttention`` module stay importable -- only ``words_loss``,
``sent_loss`` and ``func_attention`` are replaced, in the reference's modules and in every module that imported them
by name -- and the gammas come from the REFERENCE's cfg object at call time."""
import importlib
import sys
import types

import numpy as np
import pytest
import torch

from oracle import damsm_oracle as O
from oracle import ref_shim

pkg = importlib.import_module("t2i_clip-gan_b200")
needs_ref = pytest.mark.skipif(not ref_shim.available(), reason="reference modules not importable here")


def _consumer(L):
    """A module that binds the names like pretrain_DAMSM.py:5 / trainer.py:23-24 do."""
    m = types.ModuleType("_fake_pretrain")
    m.words_loss, m.sent_loss = L.words_loss, L.sent_loss
    sys.modules[m.__name__] = m
    return m


@needs_ref
def test_patch_replaces_only_the_hot_path_and_reads_the_reference_cfg():
    from checker_engine import CheckerEngine
    L, G, cfg = ref_shim.load()
    orig = (L.words_loss, L.sent_loss, G.func_attention, L.generator_loss, L.discriminator_loss, L.KL_loss)
    user = _consumer(L)
    try:
        pkg.patch_reference(L, G, engine=CheckerEngine())
        assert L.words_loss is not orig[0] and L.sent_loss is not orig[1] and G.func_attention is not orig[2]
        assert L.func_attention is G.func_attention                      # losses.py:8 alias follows
        assert user.words_loss is L.words_loss and user.sent_loss is L.sent_loss
        # everything else of the package is the reference's own
        assert (L.generator_loss, L.discriminator_loss, L.KL_loss) == orig[3:]
        x = O.make_inputs(6, 5, 9, seed=3, class_ids=True, n_classes=3)
        for g3 in (10.0, 7.0):                                            # a yml-loaded GAMMA3 reaches sent_loss
            cfg.TRAIN.SMOOTH.GAMMA3 = g3
            o = O.sent_loss(x["img"], x["sent"], x["labels"], x["class_ids"], g3)
            s0, s1 = user.sent_loss(torch.tensor(x["img"]), torch.tensor(x["sent"]), torch.tensor(x["labels"]),
                                    x["class_ids"], 6)
            assert abs(float(s0) - o["loss0"]) < 1e-6 and abs(float(s1) - o["loss1"]) < 1e-6
        cfg.TRAIN.SMOOTH.GAMMA3 = 10.0
        # stale 6-argument call (losses.py:352-354): gammas from cfg.TRAIN.SMOOTH
        cfg.TRAIN.SMOOTH.GAMMA1, cfg.TRAIN.SMOOTH.GAMMA2 = 4.0, 5.0
        o = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], 4.0, 5.0, 10.0)
        w = torch.tensor(x["words"]).permute(0, 2, 1)
        r = torch.tensor(x["regions"]).permute(0, 2, 1)
        l0, l1, _ = user.words_loss(r, w, torch.tensor(x["labels"]), torch.tensor(x["cap_len"]), x["class_ids"], 6)
        assert abs(float(l0) - o["loss0"]) < 1e-5 and abs(float(l1) - o["loss1"]) < 1e-5
    finally:
        pkg.unpatch_reference()
        sys.modules.pop("_fake_pretrain", None)
    assert (L.words_loss, L.sent_loss, G.func_attention) == orig[:3] and user.words_loss is orig[0]


@needs_ref
def test_the_reference_import_blocks_still_work_after_patching():
    """The import lines of pretrain_DAMSM.py:3-6 / trainer.py:16-24 that touch the patched packages keep working:
    nothing of ``miscc`` is shadowed."""
    L, G, cfg = ref_shim.load()
    try:
        pkg.patch_reference(L, G)
        for name in ("words_loss", "sent_loss", "discriminator_loss", "generator_loss", "KL_loss", "func_attention"):
            assert callable(getattr(L, name))
        assert hasattr(G, "GlobalAttentionGeneral") and hasattr(G, "GlobalAttention_text")     # model.py:12-13
        assert hasattr(cfg, "TRAIN") and cfg.TRAIN.SMOOTH.GAMMA3 == 10.0
    finally:
        pkg.unpatch_reference()


@needs_ref
@pytest.mark.gpu
def test_patched_reference_runs_on_the_gpu_kernels():
    L, G, cfg = ref_shim.load()
    x = O.make_inputs(8, 18, 49, seed=12, class_ids=True, n_classes=3)
    o = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], 4.0, 5.0, 10.0)
    before = pkg._lib.launch_count()
    try:
        pkg.patch_reference(L, G)
        w = torch.tensor(x["words"], device="cuda").requires_grad_(True)
        r = torch.tensor(x["regions"], device="cuda").requires_grad_(True)
        l0, l1, _ = L.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), torch.arange(8, device="cuda"),
                                 torch.tensor(x["cap_len"]), x["class_ids"], 8, torch.tensor(x["mask"]), 4.0, 5.0, 10.0)
        (l0 + l1).backward()
    finally:
        pkg.unpatch_reference()
    assert pkg._lib.launch_count() > before
    assert abs(l0.item() - o["loss0"]) <= 1e-5 and abs(l1.item() - o["loss1"]) <= 1e-5
    assert np.abs(w.grad.cpu().numpy() - o["dwords"]).max() <= 1e-5 * np.abs(o["dwords"]).max()
