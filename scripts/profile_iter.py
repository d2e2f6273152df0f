"""A few Gauss-Newton iterations of one workload and nothing else (the command ncu wraps)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feba_b200 as fb                                   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", type=int, default=3, help="BASELINE.json configs index 1..4")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--residuals", action="store_true")
ap.add_argument("--cov", action="store_true")
a = ap.parse_args()
prob = fb.synth.baseline_config(a.workload, scale=a.scale)
err, x0, _ = fb.Buildxhat(prob)
import time as _t
_t0 = _t.perf_counter()
_h = fb.Handle(prob)
print(f"feba_create (sort by point, uploads, pair schedule): {1e3 * (_t.perf_counter() - _t0):.1f} ms", flush=True)
with _h as h:
    h.set_xhat(x0)
    for i in range(a.iters):
        try:
            ds = h.iterate()
        except fb.FebaError as exc:          # timing experiments (FEBA_CHOL_SKIP) produce garbage numerics
            if exc.code != fb.lib.FEBA_ERR_NUMERIC:
                raise
            ds = float("nan")
        print(f"iteration {i + 1}: deltasum {ds:.6e}  timing {h.last_timing()}", flush=True)
    if a.residuals:
        r = h.residuals()
        print("sigma02", r["sigma02"])
    if a.cov:
        import time
        t0 = time.perf_counter()
        q = h.cov_diag()
        t1 = time.perf_counter()
        blk = h.cov_block(list(range(h.u_c - 10, h.u_c)))
        print(f"covariance stage: diag of all {h.u} unknowns {1e3 * (t1 - t0):.1f} ms (incl. inverse of the reduced "
              f"system), IOP block {1e3 * (time.perf_counter() - t1):.2f} ms; min/max cofactor {q.min():.3e} {q.max():.3e}")
    print("launches", h.launch_count())
