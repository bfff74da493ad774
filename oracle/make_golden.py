"""Record golden vectors from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (where /root/reference is mounted):

    python oracle/make_golden.py

Imports the reference's own ``miscc/losses.py`` and ``GlobalAttention.py`` through
``oracle/ref_shim.py`` (no reference file is modified or copied), feeds them the
seeded inputs of ``oracle.damsm_oracle.make_inputs`` and stores inputs seeds +
outputs (losses, autograd gradients, attention maps) in ``tests/golden/*.npz``.
The fixtures are small (inputs are regenerated from the seed, only outputs and a
checksum of the inputs are stored) so they travel to the GPU box, where
/root/reference does not exist.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import damsm_oracle as O      # noqa: E402
from oracle import ref_shim as RS         # noqa: E402

GAMMAS = (4.0, 5.0, 10.0)                 # cfg/DAMSM/bird.yml:27-29, coco.yml:27-29

# name -> (B, T, R, seed, class_ids?, n_classes, grad-slice)   D = 512 everywhere
CASES = {
    "tiny_b6_t5_r9_cls":      (6, 5, 9, 11, True, 3),
    "c1_bird_b48_t18_r49":    (48, 18, 49, 2026 + 0, True, 200),    # BASELINE configs[0]
    "c2_coco_b48_t18_r49":    (48, 18, 49, 2026 + 1, False, 0),     # BASELINE configs[1]
    "c3_dmgan_b10_t77_r49":   (10, 77, 49, 2026 + 2, True, 200),    # BASELINE configs[2]
    "t28_b16_r49_cls":        (16, 28, 49, 7, True, 5),             # reference's actual T = 30-2
    "vitb16_b6_t77_r196":     (6, 77, 196, 9, False, 0),            # configs[3]/[4] tile shape, small batch
    "ragged_b7_t13_r16":      (7, 13, 16, 13, True, 2),             # odd sizes, heavy class collisions
}


NTX_TEMPERATURE = 0.5                     # pretrain_DAMSM.py:447, trainer.py:288

SAMPLE_STRIDE = 61   # gradients / maps are stored as a strided sample + full-tensor norms


def pack(rec, key, arr, full=False):
    """Store ``arr`` whole when small, else every SAMPLE_STRIDE-th element plus |.|_2, sum, max|.|."""
    arr = np.asarray(arr, np.float32)
    if full or arr.size <= 40000:
        rec[key] = arr
    else:
        rec[key + "__sample"] = arr.reshape(-1)[::SAMPLE_STRIDE].copy()
        rec[key + "__stats"] = np.array([np.sqrt((arr.astype(np.float64) ** 2).sum()), arr.astype(np.float64).sum(),
                                         np.abs(arr).max()], np.float64)
        rec[key + "__shape"] = np.array(arr.shape, np.int64)


def digest(x):
    h = hashlib.sha256()
    for k in ("words", "regions", "sent", "img", "mask"):
        h.update(np.ascontiguousarray(x[k]).tobytes())
    if x["class_ids"] is not None:
        h.update(np.ascontiguousarray(x["class_ids"]).tobytes())
    return h.hexdigest()


def rm_special_inputs(seed, B, n, D):
    """Seeded (mask, words_emb) the way the tokenizer pads them: <sos> w.. <eos> pad.. (prefix masks), including a
    full row (no padding) and the shortest caption (<sos><eos> only)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(2, n + 1, B)
    lens[0], lens[1 % B] = n, 2
    mask = (np.arange(n)[None, :] < lens[:, None]).astype(np.int64)
    return mask, rng.standard_normal((B, n, D)).astype(np.float32)


RM_CASES = {"rm_b7_n30_d16": (5, 7, 30, 16), "rm_b48_n30_d512": (6, 48, 30, 512), "rm_b5_n79_d512": (7, 5, 79, 512)}


def main_rm(out_dir):
    """rm_special_token (pretrain_DAMSM.py:58-79) outputs of the reference's own function -> tests/golden/aux/."""
    aux = os.path.join(out_dir, "aux")
    os.makedirs(aux, exist_ok=True)
    for name, (seed, B, n, D) in RM_CASES.items():
        mask, emb = rm_special_inputs(seed, B, n, D)
        e, m = RS.ref_rm_special_token(mask, emb)
        rec = dict(meta=np.array([seed, B, n, D], np.int64), mask_new=m.astype(np.int64),
                   emb_sum=np.float64(e.astype(np.float64).sum()), emb_l2=np.float64(np.sqrt((e.astype(np.float64) ** 2).sum())),
                   emb_sample=e.reshape(-1)[::SAMPLE_STRIDE].copy())
        np.savez_compressed(os.path.join(aux, name + ".npz"), **rec)
        print(f"{name}: out {e.shape}, kept-mask sum {int(m.sum())}")


RPREC_CASES = {"rprec_b9_c100_d64": (0, 9, 100, 64), "rprec_b4_c100_d512": (1, 4, 100, 512), "rprec_b5_c7_d33": (2, 5, 7, 33)}


def rprec_inputs(seed, B, C, D):
    rng = np.random.default_rng(seed)
    img = rng.standard_normal((B, D)).astype(np.float32)
    cand = rng.standard_normal((B, C, D)).astype(np.float32)
    cand[::2, 0] = img[::2] + 0.3 * rng.standard_normal((len(img[::2]), D)).astype(np.float32)
    return img, cand


def main_rprec(out_dir):
    """R-precision scores (trainer.py:596-601) of the reference's own statements -> tests/golden/aux/."""
    aux = os.path.join(out_dir, "aux")
    os.makedirs(aux, exist_ok=True)
    for name, (seed, B, C, D) in RPREC_CASES.items():
        img, cand = rprec_inputs(seed, B, C, D)
        s, h = RS.ref_r_precision(img, cand)
        np.savez_compressed(os.path.join(aux, name + ".npz"), meta=np.array([seed, B, C, D], np.int64),
                            scores0=s.astype(np.float32), hit=h)
        print(f"{name}: hits {int(h.sum())}/{B}")


def main():
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "rm_special_token":      # only the token-gather fixtures
        return main_rm(out_dir)
    if len(sys.argv) > 1 and sys.argv[1] == "r_precision":           # only the R-precision fixtures
        return main_rprec(out_dir)
    main_rm(out_dir)
    main_rprec(out_dir)
    for name, (B, T, R, seed, cls, ncls) in CASES.items():
        x = O.make_inputs(B, T, R, seed=seed, class_ids=cls, n_classes=max(ncls, 1))
        w = RS.ref_words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], GAMMAS)
        s = RS.ref_sent_loss(x["img"], x["sent"], x["labels"], x["class_ids"], GAMMAS[2])
        rec = dict(meta=np.array([B, T, R, 512, seed, int(cls), max(ncls, 1)], np.int64),
                   gammas=np.array(GAMMAS), digest=np.array(digest(x)),
                   w_loss0=np.float32(w["loss0"]), w_loss1=np.float32(w["loss1"]),
                   s_loss0=np.float32(s["loss0"]), s_loss1=np.float32(s["loss1"]))
        for key, arr in (("dwords", w["dwords"]), ("dregions", w["dregions"]), ("attn0", w["attn0"]),
                         ("dimg", s["dimg"]), ("dtxt", s["dtxt"])):
            pack(rec, key, arr)
        nx = RS.ref_nt_xent(x["sent"], x["img"], NTX_TEMPERATURE)       # as pretrain_DAMSM.py:170-174: two (B, D) codes
        rec["ntx_loss"] = np.float32(nx["loss"])
        rec["ntx_temperature"] = np.float64(NTX_TEMPERATURE)
        pack(rec, "ntx_dzi", nx["dz_i"])
        pack(rec, "ntx_dzj", nx["dz_j"])
        if int(np.sqrt(R)) ** 2 == R:
            rng = np.random.default_rng(seed + 1000)
            dwc = rng.standard_normal((B, T, 512)).astype(np.float32)
            fa = RS.ref_func_attention(x["words"], x["regions"], GAMMAS[0], x["mask"], dwc)
            for key in ("wc", "attn", "dquery", "dcontext"):
                pack(rec, "fa_" + key, fa[key])
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
        print(f"{name}: w=({w['loss0']:.6f},{w['loss1']:.6f}) s=({s['loss0']:.6f},{s['loss1']:.6f})")


if __name__ == "__main__":
    main()
