"""The oracle (numpy fp64 restatement and the torch procedure port) against golden vectors recorded
from the unmodified reference (the reference has no tests of its own)."""
import numpy as np
import pytest

from oracle import damsm_oracle as O
from oracle import ref_port
from golden_util import Golden, case_names

# the goldens are fp32 outputs of the reference; its own fp32 round-off vs fp64 is ~1e-6
TOL_LOSS = 2e-6
TOL_GRAD = 5e-6


@pytest.mark.parametrize("name", case_names())
def test_words_loss_oracle_matches_reference(name):
    g = Golden(name)
    x = g.x
    o = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], *g.gammas)
    assert abs(o["loss0"] - g.scalar("w_loss0")) <= TOL_LOSS * max(1.0, abs(g.scalar("w_loss0")))
    assert abs(o["loss1"] - g.scalar("w_loss1")) <= TOL_LOSS * max(1.0, abs(g.scalar("w_loss1")))
    assert g.rel_err("dwords", o["dwords"]) <= TOL_GRAD
    assert g.rel_err("dregions", o["dregions"]) <= TOL_GRAD


@pytest.mark.parametrize("name", case_names())
def test_sent_loss_oracle_matches_reference(name):
    g = Golden(name)
    x = g.x
    o = O.sent_loss(x["img"], x["sent"], x["labels"], x["class_ids"], g.gammas[2])
    assert abs(o["loss0"] - g.scalar("s_loss0")) <= TOL_LOSS * max(1.0, abs(g.scalar("s_loss0")))
    assert abs(o["loss1"] - g.scalar("s_loss1")) <= TOL_LOSS * max(1.0, abs(g.scalar("s_loss1")))
    assert g.rel_err("dimg", o["dimg"]) <= TOL_GRAD
    assert g.rel_err("dtxt", o["dtxt"]) <= TOL_GRAD


@pytest.mark.parametrize("name", case_names())
def test_nt_xent_oracle_matches_reference(name):
    g = Golden(name)
    x = g.x
    o = O.nt_xent(x["sent"], x["img"], g.scalar("ntx_temperature"))
    assert abs(o["loss"] - g.scalar("ntx_loss")) <= TOL_LOSS * max(1.0, abs(g.scalar("ntx_loss")))
    assert g.rel_err("ntx_dzi", o["dz_i"]) <= TOL_GRAD
    assert g.rel_err("ntx_dzj", o["dz_j"]) <= TOL_GRAD


@pytest.mark.parametrize("name", [n for n in case_names() if Golden(n).has("fa_wc")])
def test_func_attention_oracle_matches_reference(name):
    g = Golden(name)
    x = g.x
    rng = np.random.default_rng(int(g.z["meta"][4]) + 1000)
    dwc = rng.standard_normal((g.B, g.T, g.D)).astype(np.float32)
    wc, P, dq, dc = O.func_attention(x["words"], x["regions"], g.gammas[0], x["mask"], dwc)
    h = int(np.sqrt(g.R))
    assert g.rel_err("fa_wc", wc) <= TOL_GRAD
    assert g.rel_err("fa_attn", P.reshape(g.B, g.T, h, h)) <= TOL_GRAD
    assert g.rel_err("fa_dquery", dq) <= TOL_GRAD
    assert g.rel_err("fa_dcontext", dc) <= TOL_GRAD


@pytest.mark.parametrize("name", ["tiny_b6_t5_r9_cls", "ragged_b7_t13_r16", "c3_dmgan_b10_t77_r49"])
def test_procedure_port_matches_reference(name):
    """oracle/ref_port.py (what bench.py times as the CPU baseline) reproduces the reference's numbers."""
    g = Golden(name)
    r = ref_port.step(g.x, g.gammas)
    for k in ("w_loss0", "w_loss1", "s_loss0", "s_loss1"):
        assert abs(r[k] - g.scalar(k)) <= 1e-5 * max(1.0, abs(g.scalar(k))), k
    for k in ("dwords", "dregions", "dimg", "dtxt"):
        assert g.rel_err(k, r[k]) <= 2e-5, k


def test_padded_words_get_gradient():
    """padding is masked only in the first softmax; padded words still
    contribute to R(Q,D) and receive gradient (losses.py:127,173-174,198-203)."""
    g = Golden("tiny_b6_t5_r9_cls")
    x = g.x
    o = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], *g.gammas)
    pad = x["mask"] == 0
    assert pad.any()
    assert np.abs(o["dwords"][pad]).max() > 1e-6


def test_attention_map_golden():
    g = Golden("tiny_b6_t5_r9_cls")
    x = g.x
    q, _ = O.l2norm(x["words"])
    v, _ = O.l2norm(x["regions"])
    u = np.sqrt((q * q).sum(-1))
    G = np.einsum("jrd,jsd->jrs", v, v)
    blk = O._pair_block(q[0], u[0], (x["mask"][0] != 0).astype(float), v, G, g.gammas[0], g.gammas[1])
    # reference attn_maps[0] is (B, R, T): softmax over words of caption 0 vs every image
    assert g.rel_err("attn0", blk["P"].transpose(0, 2, 1)) <= TOL_GRAD


@pytest.mark.parametrize("name", ["tiny_b6_t5_r9_cls", "c2_coco_b48_t18_r49"])
def test_ntxent_procedure_port_matches_reference(name):
    """oracle/ref_port.ntxent_step (bench.py's CPU baseline for the NT-Xent workloads)."""
    g = Golden(name)
    r = ref_port.ntxent_step(g.x["sent"], g.x["img"], g.scalar("ntx_temperature"))
    assert abs(r["loss"] - g.scalar("ntx_loss")) <= 1e-5 * max(1.0, abs(g.scalar("ntx_loss")))
    assert g.rel_err("ntx_dzi", r["dz_i"]) <= 2e-5 and g.rel_err("ntx_dzj", r["dz_j"]) <= 2e-5
