"""Host-side problem build: native packer vs the interpreted mirror of main.m:196-384
on a BASELINE.json workload written in the reference's text formats.

    python scripts/pack_bench.py [--workload config4] [--scale 1.0] [--python]

Prints one JSON line: file sizes, seconds of feba_pack_read (stages + total incl. the copies into
numpy), observations/s, and -- with --python -- the interpreted build on the same files.
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import bench  # noqa: E402
import feba_b200 as fb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config4")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--python", action="store_true", help="also time problem.load_problem (minutes at 10M observations)")
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    prob, desc = bench.make_workload(args.workload, args.scale)
    with tempfile.TemporaryDirectory() as d:
        t0 = time.perf_counter()
        fb.save_problem(prob, d, "net")
        t_write = time.perf_counter() - t0
        sizes = {e: os.path.getsize(os.path.join(d, "net" + e)) for e in (".pho", ".ext", ".cnt", ".int")}
        best, nat = None, None
        for _ in range(3):
            t0 = time.perf_counter()
            nat = fb.load_problem_native(d, threads=args.threads)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        assert nat is not None and nat.n_obs == prob.n_obs
        assert np.array_equal(nat.obs_pt, prob.obs_pt) and np.array_equal(nat.obs_img, prob.obs_img)
        assert np.array_equal(nat.obs_x, prob.obs_x)            # repr() round-trips doubles exactly
        line = {"workload": args.workload, "scale": args.scale, "n_obs": prob.n_obs, "n_img": prob.numImg,
                "n_pts": prob.numPts, "bytes": sizes, "write_s": t_write, "native_total_s": best,
                "native_stages_s": {"small_tables": nat.pack_seconds[0], "pho": nat.pack_seconds[1],
                                    "tie": nat.pack_seconds[2]},
                "native_obs_per_s": prob.n_obs / best, "threads": args.threads or os.cpu_count()}
        if args.python:
            t0 = time.perf_counter()
            py = fb.load_problem(d)
            line["python_total_s"] = time.perf_counter() - t0
            assert np.array_equal(py.obs_pt, nat.obs_pt) and np.array_equal(py.pt_tie, nat.pt_tie)
            line["speedup_vs_python"] = line["python_total_s"] / best
    print(json.dumps(line))


if __name__ == "__main__":
    main()
