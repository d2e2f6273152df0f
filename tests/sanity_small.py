"""Small end-to-end diagnostic: every stage of the CUDA path against the oracle on tiny problems.
Used for the first GPU contact and under compute-sanitizer (keeps sizes minimal)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feba_b200 as fb                                   # noqa: E402
from feba_b200 import synth                              # noqa: E402
from oracle import dense, model, sparse                  # noqa: E402
from tests import golden                                 # noqa: E402


def stage_report(tag, prob):
    err, x0, _ = fb.Buildxhat(prob)
    nb = sparse.normal_blocks(prob, x0)
    d_ref, S_ref, g_ref = sparse.reduce_and_solve(prob, nb)
    d_ref = sparse.unscale(prob, nb["L"], nb["q"], d_ref)
    with fb.Handle(prob) as h:
        h.set_xhat(x0)
        assert np.array_equal(h.get_xhat(), x0)
        h.iterate_assemble()
        S, g = h.debug_reduced()
        dc, dp = h.iterate_solve()
        d = h.get_delta()
        sc = np.sqrt(np.abs(np.diag(S_ref)))
        eS = np.max(np.abs(S - S_ref) / np.outer(sc, sc))
        eg = np.max(np.abs(g - g_ref) / sc) / np.max(np.abs(g_ref) / sc)
        ed = np.linalg.norm(d - d_ref) / np.linalg.norm(d_ref)
        print(f"[{tag}] n_obs={prob.n_obs} u={prob.u} u_c={prob.u_c}: S err {eS:.2e}  g err {eg:.2e}  "
              f"delta err {ed:.2e}  deltasum {dc + dp:.9e} (ref {np.sum(np.abs(d_ref)):.9e})", flush=True)
        h.set_xhat(x0)
        it, trace = h.solve()
        res = h.residuals()
        xh = h.get_xhat()
    ref = sparse.gauss_newton(prob, x0)
    print(f"[{tag}] iterations {it} (ref {ref['iterations']})  trace {np.array2string(trace, precision=6)}")
    print(f"[{tag}] max|v-v_ref| {np.max(np.abs(res['v'] - ref['v'])):.2e}  sigma02 {res['sigma02']:.12f} "
          f"(ref {ref['sigma02']:.12f})  max|RSD-ref| {np.max(np.abs(res['RSD'] - ref['RSD'])):.2e}  "
          f"xhat rel {np.linalg.norm(xh - ref['xhat']) / np.linalg.norm(ref['xhat']):.2e}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cam0", "free", "eop", "mixed", "big"]
    if "cam0" in which:
        stage_report("cam0 pinhole inner", golden.load_cam0())
        stage_report("cam0 fisheye free", golden.load_cam0(type="fisheye", inner=0))
    if "free" in which:
        stage_report("synthetic free", synth.make_network(12, 400, 8, 31, mode="free"))
    if "eop" in which:
        stage_report("synthetic eop", synth.make_network(12, 400, 8, 32, mode="eop"))
    if "mixed" in which:
        stage_report("synthetic mixed", synth.make_network(12, 400, 8, 33, mode="mixed", n_control=40))
    if "big" in which:
        stage_report("synthetic >32 obs/pt", synth.make_network(64, 120, 40, 5, mode="mixed", n_control=20))
