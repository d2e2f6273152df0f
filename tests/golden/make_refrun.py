"""Freeze outputs of the REFERENCE'S OWN SOURCE, executed by the MATLAB-subset interpreter
(oracle/mlab.py, oracle/refrun.py), into fixtures that travel without the reference tree.
Needs /root/reference; run in the build container:  python tests/golden/make_refrun.py

* cam0_refrun_pinhole.npz -- bundled cam0 data, shipped config.cfg (Type 'pinhole', inner constraints,
                             IOP + 5 radial + decentering terms estimated):
                             Buildxhat.m (xhat0, names); BuildAwG.m at xhat0 (non-zeros of A, w, G,
                             dist_scaling); main.m:396-494 (deltasum trace, xhat of every iteration, last delta,
                             diag(Cx)); main.m:446-456 Correlation (IOP sub-matrix and the EOP+IOP blocks of
                             images 1 and 17); main.m:569 v; BuildRSD.m; main.m:592-602 RMSx RMSy RMS sigma02
* cam0_refrun_fisheye.npz -- the same with Type 'fisheye' (no Correlation: it is a u^2 interpreted loop)
* syn_refrun_mixed.npz    -- synthetic 8-image network (synth.make_network(8, 150, 6, 4242, mode='mixed',
                             n_control=30, NK=3)), Estimate_Yc = 0, Estimate_Phi = 0, no decentering terms, no inner
                             constraints, control + tie points: exercises the flag compaction of the xhat layout
* syn_refrun_2cam.npz     -- synthetic free network (inner constraints, Estimate_AllGCP, IOP + 5 radial + decentering),
                             8 images dealt round-robin to TWO cameras: per-camera IOP blocks (cam_num, main.m:322;
                             BuildAwG.m:110-155,448)
* syn_refrun_eop.npz      -- synthetic EOP-only adjustment with fixed control points (the shape of BASELINE configs[1])
These are outputs of the reference's statements run on NumPy / libm, not of MATLAB (see oracle/mlab.py).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import feba_b200 as fb                      # noqa: E402
from oracle import refrun                   # noqa: E402
from tests import golden                    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def synthetic_mixed():
    prob = fb.synth.make_network(8, 150, 6, 4242, mode="mixed", n_control=30, NK=3)
    s = prob.settings
    s.Estimate_Yc, s.Estimate_p, s.Estimate_decent, s.Inner_Constraints = 0, 0, 0, 0
    return prob


def synthetic_two_cameras():
    import copy
    base = fb.synth.make_network(8, 150, 6, 777, mode="free")
    prob = copy.copy(base)
    prob.settings = copy.copy(base.settings)
    prob.img_cam = (np.arange(base.numImg) % 2).astype(np.int32)
    prob.iop0 = np.repeat(base.iop0, 2, axis=0)
    prob.iop0[1, :3] += np.array([0.6, -0.4, 1.1])
    prob.cam_box = np.repeat(base.cam_box, 2, axis=0)
    prob.camera_ids = ["0", "1"]
    return prob


def synthetic_eop_only():
    return fb.synth.make_network(8, 150, 6, 778, mode="eop")


def freeze(R, prob, name, correlation=False, blocks=()):
    t0 = time.time()
    err, xhat0, names = R.buildxhat(prob)
    assert err == 0
    awg = R.buildawg(prob, xhat0)
    assert awg["error"] == 0
    r, c, val = refrun.sparse_rows(awg["A"])
    run = R.gauss_newton(prob, xhat0, correlation=correlation)
    extra = {}
    if correlation:
        s = prob.settings
        ui, uc, off = s.u_perimage, s.u_percam, s.u_perimage * prob.numImg
        C = run["Correlation"]
        extra["corr_iop"] = C[off:off + uc, off:off + uc]
        for j in blocks:
            idx = np.concatenate([ui * j + np.arange(ui), off + np.arange(uc)])
            extra[f"corr_img{j}"] = C[np.ix_(idx, idx)]
    np.savez_compressed(
        os.path.join(OUT, name), xhat0=xhat0, xhatnames=np.array(names), A_shape=np.array(awg["A"].shape),
        A_rows=r, A_cols=c, A_vals=val, w0=awg["w"], G0=(awg["G"] if awg["G"] is not None else np.zeros((0, 0))),
        dist_scaling=awg["dist_scaling"], iterations=run["iterations"], deltasum=np.array(run["deltasum"]),
        xhat_arr=run["xhat_arr"], xhat=run["xhat"], delta=run["delta"], Cx_diag=run["Cx_diag"], v=run["v"],
        RSD=run["RSD"], RMSx=run["RMSx"], RMSy=run["RMSy"], RMS=run["RMS"], sigma02=run["sigma02"], **extra)
    print(f"{name}: {run['iterations']} iterations, deltasum {run['deltasum']}, sigma02 {run['sigma02']!r} "
          f"({time.time() - t0:.1f} s)")


def main():
    R = refrun.Reference()
    freeze(R, golden.load_cam0(), "cam0_refrun_pinhole.npz", correlation=True, blocks=(0, 16))
    freeze(R, golden.load_cam0(type="fisheye"), "cam0_refrun_fisheye.npz")
    freeze(R, synthetic_mixed(), "syn_refrun_mixed.npz")
    freeze(R, synthetic_two_cameras(), "syn_refrun_2cam.npz")
    freeze(R, synthetic_eop_only(), "syn_refrun_eop.npz")


if __name__ == "__main__" and "--report" not in sys.argv:
    main()


def report_case():
    """Small mixed network (control + tie points, IOP + 2 radial + decentering terms, no inner constraints)."""
    return fb.synth.make_network(8, 60, 6, 4242, mode="mixed", n_control=15, NK=2)


def freeze_report():
    """syn_refrun_report.out -- the text the reference's OWN report writer (main.m:631-950, printCell.m,
    printEOP/Dist/TIE, countImagePoints/TargetImages) produces for report_case() after its own loop
    (main.m:396-602), with fopen/fprintf redirected to a text sink; version / date / time fixed."""
    R = refrun.Reference()
    prob = report_case()
    txt = refrun.report_text(R, prob, fb.Buildxhat(prob)[1])
    with open(os.path.join(OUT, "syn_refrun_report.out"), "w") as fh:
        fh.write(txt)
    print("syn_refrun_report.out:", len(txt), "bytes")


if __name__ == "__main__" and "--report" in sys.argv:
    freeze_report()
