"""Load the golden vectors recorded from the unmodified reference (oracle/make_golden.py)."""
from __future__ import annotations

import glob
import hashlib
import os

import numpy as np

from oracle import damsm_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SAMPLE_STRIDE = 61


def case_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def _digest(x):
    h = hashlib.sha256()
    for k in ("words", "regions", "sent", "img", "mask"):
        h.update(np.ascontiguousarray(x[k]).tobytes())
    if x["class_ids"] is not None:
        h.update(np.ascontiguousarray(x["class_ids"]).tobytes())
    return h.hexdigest()


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        B, T, R, D, seed, cls, ncls = (int(v) for v in self.z["meta"])
        self.B, self.T, self.R, self.D = B, T, R, D
        self.gammas = tuple(float(g) for g in self.z["gammas"])
        self.x = O.make_inputs(B, T, R, D=D, seed=seed, class_ids=bool(cls), n_classes=ncls)
        assert _digest(self.x) == str(self.z["digest"]), "seeded inputs no longer reproduce the recorded ones"

    def has(self, key):
        return key in self.z.files or key + "__sample" in self.z.files

    def scalar(self, key):
        return float(self.z[key])

    def rel_err(self, key, arr):
        """max |arr - golden| / max |golden| over the stored elements (+ norm check for sampled tensors)."""
        arr = np.asarray(arr, np.float64)
        if key in self.z.files:
            ref = self.z[key].astype(np.float64)
            assert ref.shape == arr.shape, (key, ref.shape, arr.shape)
            return float(np.abs(arr - ref).max() / max(np.abs(ref).max(), 1e-30))
        ref = self.z[key + "__sample"].astype(np.float64)
        l2, sm, mx = self.z[key + "__stats"]
        assert tuple(self.z[key + "__shape"]) == arr.shape, (key, arr.shape)
        e_sample = np.abs(arr.reshape(-1)[::SAMPLE_STRIDE] - ref).max() / max(mx, 1e-30)
        e_norm = abs(np.sqrt((arr ** 2).sum()) - l2) / max(l2, 1e-30)
        return float(max(e_sample, e_norm))
