"""R-precision scoring (trainer.py:587-603).  PINNED: the oracle is checked against golden vectors recorded
from the reference's OWN statements (trainer.py:596-601, cut out of the training method with ``ast`` by
``oracle/ref_shim.ref_r_precision``; ``oracle/make_golden.py r_precision``), live against them where /root/reference
exists, and against the procedure port; the CUDA kernel is checked against the oracle."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import damsm_oracle as O
from oracle import ref_port


def inputs(B, C, D, seed):
    rng = np.random.default_rng(seed)
    img = rng.standard_normal((B, D)).astype(np.float32)
    cand = rng.standard_normal((B, C, D)).astype(np.float32)
    cand[::2, 0] = img[::2] + 0.3 * rng.standard_normal((len(img[::2]), D)).astype(np.float32)   # half the images: true caption close
    return img, cand


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aux")


@pytest.mark.parametrize("name", ["rprec_b9_c100_d64", "rprec_b4_c100_d512", "rprec_b5_c7_d33"])
def test_oracle_matches_golden_from_the_reference_statements(name):
    from oracle import make_golden
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    seed, B, C, D = (int(v) for v in g["meta"])
    img, cand = make_golden.rprec_inputs(seed, B, C, D)
    s, h = O.r_precision_scores(img, cand)
    assert np.abs(s - g["scores0"]).max() < 1e-6 and np.array_equal(h, g["hit"])


def test_oracle_matches_live_reference_statements():
    from oracle import ref_shim
    if not ref_shim.sources_available():
        pytest.skip("/root/reference not present")
    img, cand = inputs(7, 100, 128, 3)
    s_ref, h_ref = ref_shim.ref_r_precision(img, cand)
    s, h = O.r_precision_scores(img, cand)
    assert np.abs(s - s_ref).max() < 1e-6 and np.array_equal(h, h_ref)


def test_oracle_matches_procedure_port():
    img, cand = inputs(9, 100, 64, 0)
    s_ref, h_ref = ref_port.r_precision_step(img, cand)
    s, h = O.r_precision_scores(img, cand)
    assert np.abs(s - s_ref).max() < 1e-6 and np.array_equal(h, h_ref) and 0 < h.sum() < 9


@pytest.mark.gpu
@pytest.mark.parametrize("B,C,D", [(10, 100, 512), (3, 1, 16), (5, 37, 33), (64, 100, 512)])
def test_cuda_matches_oracle(B, C, D):
    pkg = importlib.import_module("t2i_clip-gan_b200")
    img, cand = inputs(B, C, D, B + C)
    s, h = O.r_precision_scores(img, cand)
    sc, hit = pkg.r_precision_scores(torch.tensor(img, device="cuda"), torch.tensor(cand, device="cuda"))
    assert hit.dtype == torch.bool and tuple(sc.shape) == (B, C)
    assert np.abs(sc.cpu().numpy() - s).max() <= 1e-5
    assert np.array_equal(hit.cpu().numpy(), h)


@pytest.mark.gpu
def test_cuda_ties_and_degenerate_norms():
    """Exact ties resolve to the first maximum (torch.argmax); a zero image code gives scores 0 everywhere -> hit."""
    pkg = importlib.import_module("t2i_clip-gan_b200")
    img = torch.zeros(2, 8, device="cuda")
    img[1, 0] = 1.0
    cand = torch.zeros(2, 4, 8, device="cuda")
    cand[1, :, 0] = torch.tensor([0.5, 2.0, 2.0, -1.0])          # cosines 1, 1, 1, -1: tie -> index 0
    sc, hit = pkg.r_precision_scores(img, cand)
    assert hit.tolist() == [True, True] and float(sc[0].abs().max()) == 0.0
    cand[1, 0, 0] = -0.5                                            # now the first maximum is index 1
    _, hit = pkg.r_precision_scores(img, cand)
    assert hit.tolist() == [True, False]
