"""Identifier-friendly alias of the ``t2i_clip-gan_b200`` package (whose directory name has a hyphen)."""
import importlib as _il
import sys as _sys

_pkg = _il.import_module("t2i_clip-gan_b200")
_sys.modules[__name__] = _pkg
