"""ctypes binding of libdamsm_b200.so (the C ABI declared in include/damsm_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, an exception is
raised.  PyTorch is used by the callers only to own device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libdamsm_b200.so")

_p, _i, _l, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argument ctypes, in header order (return type is int unless listed in _RESTYPE)
SIGNATURES = {
    "damsm_version": [],
    "damsm_last_error": [],
    "damsm_device_info": [_p, _p, _p, _p],
    "damsm_l2norm_fwd": [_p, _i, _l, _l, _l, _l, _l, _l, _p, _p, _l, _p, _p, _p],
    "damsm_l2norm_bwd": [_p, _i, _l, _l, _l, _l, _l, _l, _p, _p, _l, _l, _p, _p, _l, _l, _l, _p],
    "damsm_gram_f32": [_p, _l, _l, _l, _p, _p],
    "damsm_gram_bwd_f32": [_p, _p, _l, _l, _l, _p, _p],
    "damsm_words_fwd_f32": [_p, _p, _p, _p, _p, _l, _l, _l, _l, _l, _f, _f, _f, _p, _p],
    "damsm_words_bwd_f32": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _l, _l, _l, _l, _l, _l, _l, _f, _f, _f,
                            _p, _p, _p, _p, _p],
    "damsm_words_f32_smem_bytes": [_l, _l],
    "damsm_resize_nearest_fwd": [_p, _l, _l, _l, _l, _l, _l, _p, _p],
    "damsm_resize_nearest_bwd": [_p, _l, _l, _l, _l, _l, _p, _p],
    "damsm_gemm_tc": [_p, _l, _i, _p, _l, _i, _i, _l, _l, _l, _f, _p, _i, _p, _l, _p],
    "damsm_words_tc_gx_cols": [_l],
    "damsm_gram_pack_tc": [_p, _l, _l, _p, _p],
    "damsm_words_tc_smem_bytes": [_l, _l, _l],
    "damsm_words_tc_plan": [_p, _l, _l, _p, _p, _p],
    "damsm_words_fwd_tc": [_p, _l, _p, _p, _p, _p, _p, _p, _p, _l, _l, _l, _l, _l, _f, _f, _f, _p, _p, _p],
    "damsm_words_bwd_tc_col_bytes": [_l, _l],
    "damsm_words_bwd_tc_fixed_bytes": [],
    "damsm_words_bwd_tc": [_p, _l, _p, _p, _p, _p, _p, _p, _p, _p, _p, _l, _p, _p, _p, _p, _p, _p, _l, _l, _l, _l, _l,
                             _l, _l, _f, _f, _f, _p, _l, _p, _p, _p, _p, _p, _p],
    "damsm_pad_terms_fwd": [_p, _p, _p, _p, _l, _l, _l, _l, _l, _l, _f, _p, _p, _p, _p, _p, _p],
    "damsm_pad_terms_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _l, _l, _l, _l, _l, _l, _l, _l,
                              _f, _f, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "damsm_ce_stats_f32": [_p, _p, _p, _l, _l, _l, _p, _p, _p, _p],
    "damsm_ce_losses_f32": [_p, _p, _p, _p, _l, _l, _l, _l, _p, _p],
    "damsm_cos_logits_f32": [_p, _l, _p, _l, _l, _l, _l, _f, _f, _p, _p, _p, _p],
    "damsm_cos_logits_bwd_f32": [_p, _l, _p, _l, _p, _p, _p, _p, _p, _p, _p, _l, _l, _l, _l, _l, _f, _f,
                                 _p, _p, _p, _p],
    "damsm_sent_fused_ok": [_f],
    "damsm_sent_fwd_fused_f32": [_p, _l, _p, _l, _p, _p, _l, _l, _l, _l, _f, _f, _p, _p, _p, _p, _p, _p, _p],
    "damsm_sent_bwd_fused_f32": [_p, _l, _p, _l, _p, _p, _p, _p, _p, _p, _p, _l, _l, _l, _l, _l, _f, _f, _p, _p, _p],
    "damsm_ntxent_fwd_f32": [_p, _l, _l, _l, _f, _f, _p, _p, _p, _p, _p],
    "damsm_ntxent_bwd_f32": [_p, _l, _l, _l, _f, _f, _p, _p, _p, _p, _p, _p, _p],
    "damsm_rprecision_f32": [_p, _l, _p, _l, _l, _l, _l, _l, _f, _p, _p, _p],
    "damsm_project_regions_fwd": [_p, _i, _l, _l, _l, _p, _p, _l, _p, _p, _p, _p, _p, _p],
    "damsm_project_regions_bwd": [_p, _l, _l, _l, _p, _l, _p, _p, _p, _p, _p, _p],
    "damsm_rm_special_token_fwd": [_p, _l, _l, _l, _l, _l, _l, _p, _l, _l, _p, _p, _p],
    "damsm_rm_special_token_bwd": [_p, _l, _l, _l, _l, _p, _l, _l, _p, _p],
    "damsm_func_attention_fwd_f32": [_p, _p, _p, _l, _l, _l, _p, _l, _l, _l, _l, _f, _p, _p, _p, _p],
    "damsm_func_attention_bwd_f32": [_p, _p, _p, _l, _l, _l, _p, _p, _p, _p, _l, _l, _l, _l, _f, _p, _p, _p, _p],
}
_RESTYPE = {"damsm_last_error": C.c_char_p, "damsm_words_f32_smem_bytes": C.c_int64,
            "damsm_words_tc_gx_cols": C.c_int64, "damsm_words_tc_smem_bytes": C.c_int64,
            "damsm_words_bwd_tc_col_bytes": C.c_int64, "damsm_words_bwd_tc_fixed_bytes": C.c_int64}

# kernels launched per successful call of each entry point (bench.py reports the total as gpu_launches)
LAUNCHES = {
    "damsm_l2norm_fwd": 1, "damsm_l2norm_bwd": 1, "damsm_gram_f32": 1, "damsm_gram_bwd_f32": 1,
    "damsm_words_fwd_f32": 1, "damsm_words_bwd_f32": 1,
    "damsm_resize_nearest_fwd": 1, "damsm_resize_nearest_bwd": 1, "damsm_gemm_tc": 1, "damsm_gram_pack_tc": 1, "damsm_words_tc_plan": 1, "damsm_words_fwd_tc": 1, "damsm_words_bwd_tc": 2,
    "damsm_pad_terms_fwd": 2, "damsm_ce_stats_f32": 2, "damsm_ce_losses_f32": 1,
    "damsm_cos_logits_f32": 4, "damsm_cos_logits_bwd_f32": 5,
    "damsm_sent_fwd_fused_f32": 1, "damsm_sent_bwd_fused_f32": 1,
    "damsm_ntxent_fwd_f32": 3, "damsm_ntxent_bwd_f32": 4,
    "damsm_rm_special_token_fwd": 1, "damsm_rm_special_token_bwd": 1,
    "damsm_project_regions_fwd": 1, "damsm_project_regions_bwd": 1, "damsm_rprecision_f32": 1,
    "damsm_func_attention_fwd_f32": 1, "damsm_func_attention_bwd_f32": 1,
}
_launches = 0

_lib = None


class DamsmError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources in-tree for sm_100a (``make`` at the repo root)."""
    cmd = ["make", "-C", _ROOT, "all"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise DamsmError("building libdamsm_b200.so failed")
    return LIB_PATH


def load():
    """Load (once) and type the library.  Raises DamsmError when the .so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise DamsmError(
            f"{LIB_PATH} not found: the DAMSM kernels are CUDA-only (no CPU fallback). "
            "Build it with `make` at the repository root or `python -c 'import __graft_entry__ as g; g.build()'`.")
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, C.c_int)
    if lib.damsm_version() != 1:
        raise DamsmError("libdamsm_b200.so ABI version mismatch")
    _lib = lib
    return lib


def call(name: str, *args):
    """Call an int-returning entry point and raise on a non-zero status."""
    global _launches
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise DamsmError(f"{name} failed ({rc}): {lib.damsm_last_error().decode(errors='replace')}")
    _launches += LAUNCHES.get(name, 0)


def add_launches(n: int) -> None:
    """Entry points that launch a data-dependent number of kernels report the extra ones here."""
    global _launches
    _launches += int(n)


def launch_count() -> int:
    return _launches


def reset_launch_count() -> None:
    global _launches
    _launches = 0


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
