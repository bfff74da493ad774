"""Build recipe for ``oracle/_ref``: the reference's own hot-path modules as importable byte code.  TEST INFRASTRUCTURE ONLY.

The reference (``/root/reference``) is pure Python and does not exist on the GPU box.  To let ``bench.py``'s
``cpu_baseline`` leg and ``bench.py --impl reference`` time the UNMODIFIED reference there (``cpu_baseline.kind ==
"reference"``) instead of a port, this script byte-compiles the reference files the hot path needs, from where they
lie, into ``oracle/_ref/`` (byte-code files ``<module>.refbc``: build outputs only -- ``oracle/_ref/`` is git-ignored and no
reference source is copied into the repository; the directory travels to the GPU box like the built ``.so``).

    python oracle/build_ref.py        # needs /root/reference; a no-op (exit 0) where it is absent

Files (all under ``DMGAN+CLIP/code/``): ``miscc/__init__.py``, ``miscc/config.py``, ``miscc/losses.py``
(words_loss :219-272, sent_loss :51-91, similarity_text_image :95-216), ``GlobalAttention.py`` (func_attention
:38-160), ``nt_xent.py``, ``masks.py``.  ``oracle/ref_shim.py`` imports them from ``/root/reference`` when that exists
and from ``oracle/_ref`` otherwise.
"""
from __future__ import annotations

import os
import py_compile
import sys
import warnings

SRC_ROOT = "/root/reference/DMGAN+CLIP/code"
OUT_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ("miscc/__init__.py", "miscc/config.py", "miscc/losses.py", "GlobalAttention.py", "nt_xent.py", "masks.py")


def build(verbose: bool = True) -> bool:
    if not os.path.isfile(os.path.join(SRC_ROOT, "miscc", "losses.py")):
        if verbose:
            print("oracle/build_ref: reference sources not present at", SRC_ROOT, "- nothing to do")
        return False
    for rel in FILES:
        # byte code only; not named *.pyc because tools that snapshot the tree (gpurun) drop those
        dst = os.path.join(OUT_ROOT, rel[:-3] + ".refbc")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                          # GlobalAttention.py:1 has an invalid escape sequence
            py_compile.compile(os.path.join(SRC_ROOT, rel), cfile=dst, dfile=os.path.join("<reference>", rel),
                               doraise=True)
    with open(os.path.join(OUT_ROOT, "PYTHON_TAG"), "w") as f:       # byte code is interpreter-specific
        f.write(sys.implementation.cache_tag + "\n")
    if verbose:
        print("oracle/build_ref: compiled", len(FILES), "reference modules into", OUT_ROOT)
    return True


if __name__ == "__main__":
    build()
