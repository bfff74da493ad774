"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/damsm_b200.h declares, and the product path refuses to run without CUDA (no fallback)."""
import ctypes
import importlib
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("t2i_clip-gan_b200")


def header_symbols():
    src = open(os.path.join(ROOT, "include", "damsm_b200.h")).read()
    return sorted(set(re.findall(r"DAMSM_API[^;(]*?\b(damsm_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(pkg._lib.LIB_PATH):
        pkg._lib.build()
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/damsm_b200.h but not exported"
    assert lib.damsm_version() == 1


def test_ctypes_table_matches_header():
    assert sorted(pkg._lib.SIGNATURES) == header_symbols()
    src = open(os.path.join(ROOT, "include", "damsm_b200.h")).read()
    # argument counts agree with the header prototypes
    for name, args in pkg._lib.SIGNATURES.items():
        m = re.search(r"DAMSM_API[^;(]*?\b" + name + r"\s*\(([^;]*?)\)\s*;", src, flags=re.S)
        assert m, name
        body = m.group(1).strip()
        n = 0 if body in ("", "void") else body.count(",") + 1
        assert n == len(args), (name, n, len(args))


def test_host_only_entry_points():
    lib = pkg._lib.load()
    assert lib.damsm_words_f32_smem_bytes(18, 49) > 0
    assert lib.damsm_words_f32_smem_bytes(77, 196) <= 227 * 1024
    assert lib.damsm_words_f32_smem_bytes(129, 49) < 0
    assert lib.damsm_words_f32_smem_bytes(18, 257) < 0


def test_no_cpu_fallback():
    w = torch.randn(4, 512, 6)
    r = torch.randn(4, 512, 9)
    with pytest.raises((pkg.DamsmError, RuntimeError)):
        pkg.words_loss(r, w, torch.arange(4), torch.full((4,), 6), None, 4, torch.ones(4, 6, dtype=torch.int64),
                       4.0, 5.0, 10.0)
    with pytest.raises((pkg.DamsmError, RuntimeError)):
        pkg.sent_loss(torch.randn(4, 512), torch.randn(4, 512), torch.arange(4), None, 4)


def test_argument_validation():
    w = torch.randn(4, 512, 6)
    r = torch.randn(4, 512, 9)
    with pytest.raises(ValueError):
        pkg.words_loss(r, w, torch.arange(4), None, None, 5, torch.ones(4, 6), 4.0, 5.0, 10.0)   # wrong batch_size
    with pytest.raises(ValueError):
        pkg.func_attention(w, torch.randn(4, 512, 10), 4.0, torch.ones(4, 1, 6))                   # R not a square


def test_reference_import_lines_work():
    """`from miscc.losses import sent_loss, words_loss` / `from GlobalAttention import func_attention`
    (pretrain_DAMSM.py:5, trainer.py:23-24, losses.py:8) resolve to the drop-ins when the package
    directory is on sys.path."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); from miscc.losses import sent_loss, words_loss; "
            "from GlobalAttention import func_attention; from miscc.config import cfg; "
            "print(cfg.TRAIN.SMOOTH.GAMMA3, words_loss.__module__)") % os.path.join(ROOT, "t2i_clip-gan_b200")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split()[0] == "10.0"


def test_tensor_core_shape_query_is_a_host_function():
    """damsm_words_tc_smem_bytes needs no GPU: the wrapper uses it to route unsupported shapes to the exact path."""
    import importlib
    pkg = importlib.import_module("t2i_clip-gan_b200")
    assert pkg.ops.tc_shape_supported(77, 196, 512) and pkg.ops.tc_shape_supported(18, 49, 512)
    for bad in ((129, 49, 512), (77, 256, 512), (18, 49, 500), (100, 196, 512)):
        assert not pkg.ops.tc_shape_supported(*bad), bad
    with __import__("pytest").raises(ValueError):
        pkg.ops._pick_precision("fp16", 18, 49, 512)


def test_no_cpu_fallback_for_the_widened_operators():
    """CPU tensors never reach a CPU implementation: the C layer (or the wrapper) refuses them."""
    with pytest.raises((pkg.DamsmError, RuntimeError)):
        pkg.nt_xent(torch.randn(4, 16), torch.randn(4, 16), 0.5)
    with pytest.raises((pkg.DamsmError, RuntimeError)):
        pkg.rm_special_token(torch.ones(3, 6, dtype=torch.int64), torch.randn(3, 6, 8))
    with pytest.raises((pkg.DamsmError, RuntimeError)):
        pkg.project_regions(torch.randn(2, 5, 32), torch.randn(16, 32), torch.randn(16))
    with pytest.raises((pkg.DamsmError, RuntimeError)):
        pkg.r_precision_scores(torch.randn(2, 8), torch.randn(2, 3, 8))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No silent fallback when the CUDA library is absent: loading raises with the path and the reason."""
    monkeypatch.setattr(pkg._lib, "_lib", None)
    monkeypatch.setattr(pkg._lib, "LIB_PATH", str(tmp_path / "libdamsm_b200.so"))
    with pytest.raises(pkg.DamsmError, match="no CPU fallback"):
        pkg._lib.load()


def test_reference_import_lines_for_nt_xent_and_masks():
    """`from nt_xent import NT_Xent`, `from masks import mask_correlated_samples_2` (pretrain_DAMSM.py:34-35)."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); from nt_xent import NT_Xent; "
            "from masks import mask_correlated_samples, mask_correlated_samples_2; "
            "m = mask_correlated_samples_2(3); c = NT_Xent(3, 0.5, m, 'cpu'); print(int(m.sum()), c.batch_size)"
            ) % os.path.join(ROOT, "t2i_clip-gan_b200")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["24", "3"]                 # 6 x 6 minus the diagonal (6) minus the positives (6)
