"""Drop-in for ``from GlobalAttention import func_attention`` (DMGAN+CLIP/code/miscc/losses.py:8;
definition at GlobalAttention.py:38-160)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("t2i_clip-gan_b200")

func_attention = _pkg.func_attention


def l2norm(X, dim, eps=1e-8):
    """GlobalAttention.py:25-30."""
    return X / (X.pow(2).sum(dim=dim, keepdim=True).sqrt() + eps)
