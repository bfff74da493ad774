"""bench.py output contract on the CPU-runnable arm: exactly one JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout[:500]
    return json.loads(lines[0])


@pytest.mark.parametrize("workload,metric", [("c2", "damsm_fwd_bwd_matched_pairs_per_s"),
                                             ("ntx48", "nt_xent_fwd_bwd_rows_per_s"),
                                             ("rmtok48", "rm_special_token_fwd_bwd_captions_per_s"),
                                             ("proj48", "project_regions_fwd_bwd_images_per_s")])
def test_reference_arm_prints_one_json_line(workload, metric):
    d = run("--impl", "reference", "--workload", workload, "--steps", "2", "--warmup", "1")
    assert d["impl"] == "reference" and d["metric"] == metric and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["config"]["workload"] == workload and d["vs_baseline"] is None and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    # the word/sentence loss arm runs the unmodified reference when it is importable (oracle/ref_shim.available())
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    if workload == "c2":
        from oracle import ref_shim
        assert cb["kind"] == ("reference" if ref_shim.available() else "port")
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_reference_arm_defaults_to_c5_and_maps_no_repo_library():
    """Both arms measure c5 by default, print the same ``config`` keys, and the reference arm never loads the product's
    shared library (it is imported before the package would be: bench.py main())."""
    env = dict(os.environ, DAMSM_BENCH_LIST_MAPS="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    d = json.loads([l for l in res.stdout.splitlines() if l.strip()][0])
    c = d["config"]
    assert c["workload"] == "c5" and c["global_batch"] == 4096 and "extrapolated" in d["cpu_baseline"]["sample"]
    assert set(c) >= {"workload", "description", "global_batch", "local_batch", "T", "R", "D", "class_mask", "gammas",
                      "precision", "parallelism", "scaling", "caption_lengths", "step"}
    assert "libdamsm_b200" not in res.stderr, "the reference arm mapped the product library"
    assert "mapped-libraries:" in res.stderr
