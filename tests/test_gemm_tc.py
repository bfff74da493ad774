"""The tensor-core GEMM (csrc/gemm_tc.cu) that carries every dense contraction of the backward passes, against torch
fp64 on the same (already rounded) operands: all four storage-order combinations, the three operand formats, ragged
extents, accumulate, device-side alpha and the split-K epilogue."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("t2i_clip-gan_b200")
DT = {"fp16": torch.float16, "bf16": torch.bfloat16, "tf32": torch.float32}


def operands(m, n, k, dtype, a_mn, b_mn, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(m, k, device="cuda", generator=g).to(dtype)
    b = torch.randn(k, n, device="cuda", generator=g).to(dtype)
    a_store = a.t().contiguous() if a_mn else a.contiguous()           # (K, M) or (M, K)
    b_store = b.contiguous() if b_mn else b.t().contiguous()           # (K, N) or (N, K)
    return a, b, a_store, b_store


@pytest.mark.parametrize("fmt", ["fp16", "bf16", "tf32"])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("m,n,k", [(256, 256, 64), (512, 512, 1024), (304, 520, 200), (1000, 72, 136), (40, 768, 4104)])
def test_gemm_tc_against_torch(fmt, a_mn, b_mn, m, n, k):
    eng = pkg.get_engine("bf16")
    a, b, a_s, b_s = operands(m, n, k, DT[fmt], a_mn, b_mn, seed=m + 3 * n + 7 * k)
    ref = a.double() @ b.double()
    c = eng.gemm_tc(a_s, b_s, a_mn=a_mn, b_mn=b_mn)
    tol = 2e-3 if fmt == "tf32" else 2e-5
    err = float((c.double() - ref).abs().max() / ref.abs().max())
    assert err <= tol, (fmt, a_mn, b_mn, m, n, k, err)


def test_gemm_tc_accumulate_alpha_and_device_scalar():
    eng = pkg.get_engine("bf16")
    m, n, k = 520, 512, 328
    a, b, a_s, b_s = operands(m, n, k, torch.float16, False, True, seed=5)
    c0 = torch.randn(m, n, device="cuda")
    c = c0.clone()
    s = torch.tensor([0.25], device="cuda")
    eng.gemm_tc(a_s, b_s, b_mn=True, out=c, accumulate=True, alpha=3.0, alpha_dev=s)
    ref = c0.double() + 0.75 * (a.double() @ b.double())
    assert float((c.double() - ref).abs().max() / ref.abs().max()) <= 2e-5
    # a strided output (rows of a wider matrix)
    wide = torch.zeros(m, n + 64, device="cuda")
    eng.gemm_tc(a_s, b_s, b_mn=True, out=wide[:, 32:32 + n])
    assert float((wide[:, 32:32 + n].double() - a.double() @ b.double()).abs().max()) <= 2e-5 * float(ref.abs().max())
    assert float(wide[:, :32].abs().max()) == 0.0 and float(wide[:, 32 + n:].abs().max()) == 0.0


def test_gemm_tc_split_k_small_output_long_k():
    """dW of the region projection: 512 x 768 output, K = B (R+1) rows -> K is split over the SMs (red.add epilogue)."""
    eng = pkg.get_engine("bf16")
    m, n, k = 512, 768, 48 * 50 * 8
    a, b, a_s, b_s = operands(m, n, k, torch.float32, True, True, seed=9)
    c = eng.gemm_tc(a_s, b_s, a_mn=True, b_mn=True)
    ref = a.double() @ b.double()
    assert float((c.double() - ref).abs().max() / ref.abs().max()) <= 2e-3
    c2 = torch.ones(m, n, device="cuda")
    eng.gemm_tc(a_s.half(), b_s.half(), a_mn=True, b_mn=True, out=c2, accumulate=True)
    ref2 = 1.0 + a.half().double() @ b.half().double()
    assert float((c2.double() - ref2).abs().max() / ref2.abs().max()) <= 2e-5
