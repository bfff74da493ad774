"""Host-side logic of the tensor-core path that needs no GPU: the chunk table of the backward, the shared-memory budget of
every kernel instance (``damsm_words_tc_smem_bytes`` = ``tc_layout``: operand ring as deep as shared memory allows) and the
workspace sizing helpers."""
import importlib

import numpy as np
import pytest

pkg = importlib.import_module("t2i_clip-gan_b200")
engine = importlib.import_module("t2i_clip-gan_b200.engine")
SMEM_LIMIT = 232448


def test_tc_chunks_cover_the_sorted_captions_in_equal_column_counts():
    rng = np.random.default_rng(0)
    for n, kc_max in ((1, 128), (7, 128), (300, 1000), (4096, 16400), (513, 80)):
        nw = np.sort(16 * rng.integers(1, 6, size=n))[::-1].astype(np.int64)       # sorted, longest first
        koff, pos = engine.CudaEngine._tc_chunks(nw, kc_max)
        assert koff[0] == 0 and koff[-1] == nw.sum() and len(koff) == n + 1
        assert pos[0] == 0 and pos[-1] == n and np.all(np.diff(pos) > 0)
        sizes = koff[pos[1:]] - koff[pos[:-1]]
        assert sizes.max() <= kc_max and sizes.min() > 0
        assert sizes.max() - sizes.min() <= 2 * nw.max()                           # about equal column counts


def test_tc_chunks_reject_a_workspace_smaller_than_one_caption():
    with pytest.raises(pkg.DamsmError):
        engine.CudaEngine._tc_chunks(np.array([80, 64], dtype=np.int64), 48)


@pytest.mark.parametrize("d", [64, 256, 512])
def test_every_supported_shape_fits_shared_memory(d):
    """The ring depth adapts to the instance: whatever (T, R) the library accepts must fit the 227 KB of a CTA, and a
    shape that fits with the 3-slot minimum must not be rejected."""
    lib = pkg._lib.load()
    supported = 0
    for t in (1, 7, 16, 17, 32, 33, 48, 64, 65, 77, 80, 81, 100, 128):
        for r in (1, 9, 16, 49, 64, 100, 127, 128, 144, 196, 224, 250, 255):
            need = lib.damsm_words_tc_smem_bytes(t, r, d)
            assert need > 0
            if need <= SMEM_LIMIT:
                supported += 1
            assert pkg.ops.tc_shape_supported(t, r, d) == (need <= SMEM_LIMIT)
    assert supported > 100
    # the benchmarked shape and the caption-length group instances at R = 196
    for t in (77, 64, 32):
        assert lib.damsm_words_tc_smem_bytes(t, 196, 512) <= SMEM_LIMIT
    # smaller caption tiles leave room for more ring slots: the footprint does not grow with T within one instance
    assert lib.damsm_words_tc_smem_bytes(30, 196, 512) == lib.damsm_words_tc_smem_bytes(32, 196, 512)
    assert lib.damsm_words_tc_smem_bytes(128, 255, 512) > SMEM_LIMIT            # routed to the exact path
    for bad in ((0, 49), (129, 49), (77, 0), (77, 256)):
        assert lib.damsm_words_tc_smem_bytes(bad[0], bad[1], d) < 0


def test_backward_workspace_sizing_helpers():
    lib = pkg._lib.load()
    # per scratch column: fp16 dS rows (bc*R), fp16 e2 rows (bc * R padded to a multiple of 64), fp32 scale (bc)
    assert lib.damsm_words_bwd_tc_col_bytes(4096, 196) == 4096 * (196 * 2 + 256 * 2 + 4)
    assert lib.damsm_words_bwd_tc_col_bytes(10, 49) == 10 * (49 * 2 + 64 * 2 + 4)
    assert lib.damsm_words_bwd_tc_fixed_bytes() == 256
    assert lib.damsm_words_tc_gx_cols(196) == 256 and lib.damsm_words_tc_gx_cols(49) == 64
    assert lib.damsm_sent_fused_ok(10.0) == 1 and lib.damsm_sent_fused_ok(61.0) == 0 and lib.damsm_sent_fused_ok(-1.0) == 0
