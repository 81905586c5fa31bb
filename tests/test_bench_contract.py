"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the required keys,
and the algorithmic denominators match SURVEY.md 8(d) for the cheap configs."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1s",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and d["config"]["workload"] == "cfg1s"


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--workload", "cfg1s"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_denominators_match_survey():
    import bench
    from sparse_matrix_mult_b200 import synthetic
    w = synthetic.workload("cfg1")
    info, flops = bench.describe(w, "cfg1")
    assert info["products"] == 1_000_408 and flops == 2_000_816            # SURVEY.md 8(d), cfg 1
    assert bench.csr_bytes(w["a"]) == 12 * 100_000 + 4 * 10_001
    w = synthetic.workload("cfg3s")
    flops_full, p1, p2_full = bench.triple_flops(w["a"], w["b"], upper=False)
    flops_up, p1b, p2_up = bench.triple_flops(w["a"], w["b"], upper=True)
    assert p1 == p1b == bench.count_products(w["a"], w["b"]) and 0 < p2_up < p2_full
    # brute force on the merged structure of T = H Q
    import numpy as np
    t = (abs(w["a"]) @ abs(w["b"])).tocsr()
    h = w["a"].tocsc()
    col_rows = [h.indices[h.indptr[c]:h.indptr[c + 1]] for c in range(h.shape[1])]
    brute = sum(int((col_rows[c] >= i).sum()) for i in range(t.shape[0]) for c in t.indices[t.indptr[i]:t.indptr[i + 1]])
    assert brute == p2_up
