"""bench.py contract checks that run without a GPU: the reference arm prints ONE JSON line with the
required keys (on a shrunken workload), and the workload/roofline bookkeeping is self-consistent."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--workload", "config5block", "--scale", "0.25"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert abs(d["value"] - d["config"]["n_obs"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]


def test_algorithmic_counts_follow_the_survey_formulas():
    sys.path.insert(0, ROOT)
    import bench
    import feba_b200 as fb
    prob = fb.synth.make_network(10, 300, 6, 5, mode="free")
    c = bench.algorithmic_counts(prob)
    u_c, NC = prob.u_c, 10
    assert c["B_asm"] == 24 * prob.n_obs + 8 * (6 * prob.numImg + NC + 3 * prob.numPts) + 8 * (u_c * (u_c + 1) // 2 + u_c) \
        + 96 * prob.numtie
    assert abs(c["F_chol"] - u_c ** 3 / 3) < 1
    assert c["B_rsd"] == 64 * prob.n_obs
