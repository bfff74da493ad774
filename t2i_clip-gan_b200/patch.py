"""Switch an imported reference code base to the B200 kernels WITHOUT shadowing any of its packages.

The reference's scripts import far more from ``miscc`` / ``GlobalAttention`` than the hot path
(``miscc.utils``, ``cfg_from_file``, ``discriminator_loss`` / ``generator_loss`` / ``KL_loss``,
``GlobalAttentionGeneral`` ... -- pretrain_DAMSM.py:3-6, trainer.py:16-24, model.py:12-13), so putting stub modules
of the same names first on ``sys.path`` breaks them at their first import.  Instead, leave the reference's own
packages importable and replace only the three hot-path functions in place:

    import importlib
    damsm = importlib.import_module("t2i_clip-gan_b200")
    damsm.patch_reference(precision="bf16")        # after the reference's modules are importable, before training

What is replaced (every module that already holds a reference to the original function object is updated too, e.g.
``pretrain_DAMSM.words_loss`` bound by ``from miscc.losses import sent_loss, words_loss``, pretrain_DAMSM.py:5):

    miscc.losses.words_loss   (losses.py:219)      miscc.losses.sent_loss (losses.py:51)
    GlobalAttention.func_attention (GlobalAttention.py:38) and its alias miscc.losses.func_attention (losses.py:8)

The gammas the reference reads from its global ``cfg`` (losses.py:79; the stale 6-argument ``words_loss`` calls at
losses.py:352, trainer.py:235) are read from the REFERENCE's own ``miscc.config.cfg`` object at call time, so a yml
loaded with ``cfg_from_file`` takes effect.  ``unpatch_reference()`` restores the originals.
"""
from __future__ import annotations

import importlib
import sys

from . import ops

_saved = []


def _replace_everywhere(old, new, also=()):
    """Rebind every module-level name that currently refers to ``old`` (in sys.modules and in ``also``)."""
    seen = set()
    for mod in list(sys.modules.values()) + list(also):
        if id(mod) in seen:
            continue
        seen.add(id(mod))
        d = getattr(mod, "__dict__", None)
        if not isinstance(d, dict):
            continue
        for k, v in list(d.items()):
            if v is old:
                _saved.append((mod, k, old))
                d[k] = new


def patch_reference(losses_module=None, attention_module=None, *, precision=None, group=None, engine=None):
    """Replace the reference's hot-path functions by the drop-ins.  ``losses_module`` / ``attention_module`` default
    to the imported (or importable) ``miscc.losses`` / ``GlobalAttention`` of the reference.  ``precision``, ``group``
    and ``engine`` are forwarded to every ``words_loss`` / ``sent_loss`` call.  Returns the patched losses module."""
    L = losses_module or sys.modules.get("miscc.losses") or importlib.import_module("miscc.losses")
    G = attention_module or sys.modules.get("GlobalAttention") or importlib.import_module("GlobalAttention")
    cfg = getattr(L, "cfg", None)
    if cfg is None:
        raise RuntimeError("patch_reference: the losses module has no `cfg` (expected the reference's miscc/losses.py)")
    if getattr(L.words_loss, "_damsm_b200", False):
        return L
    extra = {}
    if precision is not None:
        extra["precision"] = precision
    if group is not None:
        extra["group"] = group
    if engine is not None:
        extra["engine"] = engine

    def words_loss(region_features, words_embs, match_labels, cap_lens, class_ids, batch_size,
                   words_mask=None, gamma1=None, gamma2=None, gamma3=None, **kw):
        s = cfg.TRAIN.SMOOTH
        return ops.words_loss(region_features, words_embs, match_labels, cap_lens, class_ids, batch_size, words_mask,
                              s.GAMMA1 if gamma1 is None else gamma1, s.GAMMA2 if gamma2 is None else gamma2,
                              s.GAMMA3 if gamma3 is None else gamma3, **dict(extra, **kw))

    def sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, eps=1e-8, **kw):
        kw.setdefault("gamma3", cfg.TRAIN.SMOOTH.GAMMA3)          # losses.py:79 reads it at call time
        fw = {k: v for k, v in extra.items() if k != "precision"}
        return ops.sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, eps, **dict(fw, **kw))

    def func_attention(query, context, gamma1, query_mask, **kw):
        return ops.func_attention(query, context, gamma1, query_mask, **kw)

    for f in (words_loss, sent_loss, func_attention):
        f._damsm_b200 = True
    words_loss.__doc__ = "B200 drop-in for miscc.losses.words_loss (losses.py:219); see t2i_clip-gan_b200.ops.words_loss"
    sent_loss.__doc__ = "B200 drop-in for miscc.losses.sent_loss (losses.py:51); see t2i_clip-gan_b200.ops.sent_loss"
    func_attention.__doc__ = "B200 drop-in for GlobalAttention.func_attention (GlobalAttention.py:38)"
    _replace_everywhere(L.words_loss, words_loss, (L, G))
    _replace_everywhere(L.sent_loss, sent_loss, (L, G))
    _replace_everywhere(G.func_attention, func_attention, (L, G))
    return L


def unpatch_reference():
    """Undo ``patch_reference``."""
    while _saved:
        mod, k, old = _saved.pop()
        mod.__dict__[k] = old
