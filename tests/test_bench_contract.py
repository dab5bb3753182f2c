"""CPU: the reference arm of bench.py prints ONE JSON line with the contract's keys, and the presets / committed
evaluation counts it relies on are consistent.  (The B200 arm needs a GPU: the driver runs it at round end.)"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--preset", "small",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert set(d["config"]) == {"workload"}  # identical `config` in both arms
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["n_lml_grad_evals_timed"] == 3 and cb["extrapolated"] is True
    assert cb["value"] == d["value"] == d["e2e"]["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["value"] > 0 and np.isfinite(d["ms_per_step"])


def test_presets_and_committed_counts():
    sys.path.insert(0, ROOT)
    import bench

    ev = json.load(open(bench.EVALS_FILE))
    for p in ("cfg2", "cfg3", "cfg4", "cfg5"):
        assert p in bench.PRESETS and ev[p]["evals"] > 0 and bench.committed_evals(p) == ev[p]["evals"]
    X, y, lab, Q = bench.make_workload("cfg2", M=100)
    assert X.shape == (2000, 2) and lab.shape == (2000,) and Q.shape == (100, 2) and lab.max() == 3
    X, y, lab, Q = bench.make_workload("small", M=64)
    assert X.shape == (4096, 6) and set(np.unique(lab)) == {0, 1, 2, 3}
    X, y, lab, Q = bench.make_workload("cfg5", M=8)
    assert X.shape[1] == 8 and lab.max() == 63 and 480 <= np.bincount(lab).min() and np.bincount(lab).max() <= 544
