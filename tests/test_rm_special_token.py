"""rm_special_token (pretrain_DAMSM.py:58-79): oracle vs the reference's own function (goldens and,
where /root/reference is mounted, live) and the CUDA gather/scatter kernels vs the oracle (bit-exact: it is a copy)."""
import glob
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import damsm_oracle as O
from oracle import ref_shim as RS
from oracle.make_golden import SAMPLE_STRIDE, rm_special_inputs

AUX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aux")
CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(AUX, "rm_*.npz")))


def load(name):
    z = np.load(os.path.join(AUX, name + ".npz"))
    seed, B, n, D = (int(v) for v in z["meta"])
    mask, emb = rm_special_inputs(seed, B, n, D)
    return z, mask, emb


def check_against_golden(z, e, m):
    assert np.array_equal(np.asarray(m, np.int64), z["mask_new"])
    assert np.array_equal(np.asarray(e, np.float32).reshape(-1)[::SAMPLE_STRIDE], z["emb_sample"])
    assert float(np.asarray(e, np.float64).sum()) == float(z["emb_sum"])


def test_cases_present():
    assert len(CASES) == 3


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_golden(name):
    z, mask, emb = load(name)
    e, m, src = O.rm_special_token(mask, emb)
    check_against_golden(z, e, m)
    assert (src[:, 0] >= 1).all() and (src[:, 0] <= 2).all() and (np.diff(src, axis=1) >= 1).all()


def test_procedure_port_matches_oracle():
    """oracle/ref_port.rm_special_token_step (bench.py's CPU baseline for the rmtok workloads)."""
    from oracle import ref_port
    mask, emb = rm_special_inputs(11, 9, 14, 8)
    dout = np.random.default_rng(2).standard_normal((9, 12, 8)).astype(np.float32)
    out, m_new, dx = ref_port.rm_special_token_step(mask, emb, dout)
    e, m, src = O.rm_special_token(mask, emb)
    want = np.zeros_like(emb)
    want[np.arange(9)[:, None], src] = dout
    assert np.array_equal(out, e) and np.array_equal(m_new, m) and np.array_equal(dx, want)


@pytest.mark.skipif(not RS.available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference_non_prefix_mask():
    """The reference keys on the FIRST zero of the mask; later ones do not matter."""
    rng = np.random.default_rng(1)
    mask = (rng.random((6, 11)) > 0.3).astype(np.int64)
    mask[:, :2] = 1
    emb = rng.standard_normal((6, 11, 4)).astype(np.float32)
    e, m, _ = O.rm_special_token(mask, emb)
    re, rm = RS.ref_rm_special_token(mask, emb)
    assert np.array_equal(e, re) and np.array_equal(m, rm)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cuda_matches_golden_and_oracle(name, dtype):
    pkg = importlib.import_module("t2i_clip-gan_b200")
    z, mask, emb = load(name)
    x = torch.tensor(emb, device="cuda").to(dtype).requires_grad_(True)
    m_in = torch.tensor(mask, device="cuda")
    out, m_new = pkg.rm_special_token(m_in, x)
    e_ref, m_ref, src = O.rm_special_token(mask, x.detach().float().cpu().numpy())
    assert out.dtype == dtype and m_new.dtype == torch.int64
    assert np.array_equal(out.detach().float().cpu().numpy(), e_ref) and np.array_equal(m_new.cpu().numpy(), m_ref)
    if dtype == torch.float32:
        check_against_golden(z, out.detach().cpu().numpy(), m_new.cpu().numpy())
    # backward: the gradient of sum(out * g) is g scattered back, zero rows at <sos> and <eos>
    g = torch.randn_like(out)
    out.backward(g)
    want = np.zeros(emb.shape, np.float32)
    want[np.arange(emb.shape[0])[:, None], src] = g.float().cpu().numpy()
    assert np.array_equal(x.grad.float().cpu().numpy(), want)


@pytest.mark.gpu
def test_cuda_strided_input_bool_mask_and_errors():
    pkg = importlib.import_module("t2i_clip-gan_b200")
    mask, emb = rm_special_inputs(3, 6, 12, 24)
    wide = torch.tensor(np.concatenate([emb, emb], axis=2), device="cuda")          # (B, n, 2D)
    x = wide[:, :, 3:3 + 21]                                                        # odd offset and width: 4-byte path
    out, m_new = pkg.rm_special_token(torch.tensor(mask.astype(bool), device="cuda"), x)
    e_ref, m_ref, _ = O.rm_special_token(mask, x.cpu().numpy())
    assert m_new.dtype == torch.bool
    assert np.array_equal(out.cpu().numpy(), e_ref) and np.array_equal(m_new.cpu().numpy(), m_ref.astype(bool))
    xt = torch.tensor(emb, device="cuda").permute(0, 2, 1)                          # innermost dim strided
    out2, _ = pkg.rm_special_token(torch.ones(6, 24, dtype=torch.int64, device="cuda"), xt)
    e2, _, _ = O.rm_special_token(np.ones((6, 24), np.int64), xt.cpu().numpy())
    assert np.array_equal(out2.cpu().numpy(), e2)
    with pytest.raises(ValueError):
        pkg.rm_special_token(torch.ones(6, 2, device="cuda"), torch.zeros(6, 2, 8, device="cuda"))
    with pytest.raises(ValueError):
        pkg.rm_special_token(torch.ones(5, 12, device="cuda"), torch.tensor(emb, device="cuda"))
