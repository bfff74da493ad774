"""Drop-in for the hot-path functions of the reference's ``miscc/losses.py``:

    from miscc.losses import sent_loss, words_loss        # pretrain_DAMSM.py:5, trainer.py:23-24

Same names, positional signatures and return values as DMGAN+CLIP/code/miscc/losses.py:51 and :219; the
work is done by the sm_100a kernels of libdamsm_b200.so.  Gammas that the reference reads from its global
``cfg`` (losses.py:79; the stale 6-argument ``words_loss`` calls at losses.py:352 and trainer.py:235) are read
from ``miscc.config.cfg`` here as well.
"""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("t2i_clip-gan_b200")

from .config import cfg  # noqa: E402


def l2norm(X, dim, eps=1e-8):
    """losses.py:13-18 (kept for callers that import it; the fused kernels normalise internally)."""
    return X / (X.pow(2).sum(dim=dim, keepdim=True).sqrt() + eps)


def words_loss(region_features, words_embs, match_labels, cap_lens, class_ids, batch_size,
               words_mask=None, gamma1=None, gamma2=None, gamma3=None, **kw):
    s = cfg.TRAIN.SMOOTH
    return _pkg.words_loss(region_features, words_embs, match_labels, cap_lens, class_ids, batch_size, words_mask,
                           s.GAMMA1 if gamma1 is None else gamma1, s.GAMMA2 if gamma2 is None else gamma2,
                           s.GAMMA3 if gamma3 is None else gamma3, **kw)


def sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, eps=1e-8, **kw):
    kw.setdefault("gamma3", cfg.TRAIN.SMOOTH.GAMMA3)
    return _pkg.sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, eps, **kw)


func_attention = _pkg.func_attention
