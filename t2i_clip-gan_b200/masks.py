"""Drop-in for ``from masks import mask_correlated_samples(_2)`` (DMGAN+CLIP/code/trainer.py:31,
pretrain_DAMSM.py:34; definitions at masks.py:3-17)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("t2i_clip-gan_b200")


def mask_correlated_samples(args):
    """masks.py:3-9: ``args`` is anything with a ``batch_size`` attribute (the trainer passes itself)."""
    return _pkg.standard_ntxent_mask(args.batch_size)


def mask_correlated_samples_2(batch_size):
    """masks.py:11-17."""
    return _pkg.standard_ntxent_mask(batch_size)
