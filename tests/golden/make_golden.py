"""Regenerate the golden fixtures.  Needs the reference tree (/root/reference); run in the build
container:  python tests/golden/make_golden.py

* cam0_problem.npz        -- the bundled cam0.* + config.cfg after main.m:60-384 (numeric SoA)
* cam0_refsrc_t<k>.npz    -- outputs of the reference's OWN BuildAwG.m statements (executed from
                             the reference source text by oracle/refexpr.py; MATLAB is not
                             available) at the initial xhat, for typeint k = 0..4: fx fy w Je Jc Jt G
* cam0_gn_<type>.npz      -- one full run of the literal dense restatement (oracle/dense.py) of
                             main.m:412-494 + 569-602 for the shipped config (pinhole) and for
                             Type 'fisheye': xhat, deltasum trace, v, RSD, sigma02, RMSx, RMSy
                             (oracle output; the reference's own loop, executed, is frozen by make_refrun.py)
* cam0_gn_fisheye_exact.npz -- the same loop with every step solved to extended precision
                             (oracle/exact.py).  cond(N) ~ 2e13 on this case: the explicit-inverse
                             run above deviates from this one by 4.3e-7 px in v (its own round-off),
                             the Cholesky/Schur paths by 4e-11.
"""
import dataclasses
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import feba_b200 as fb                      # noqa: E402
from oracle import dense, exact, model, refexpr    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    prob = fb.load_problem(refexpr.REFERENCE_ROOT)
    sd = dataclasses.asdict(prob.settings)
    sd["Meas_std_y"] = np.nan if sd["Meas_std_y"] is None else sd["Meas_std_y"]
    np.savez_compressed(
        os.path.join(OUT, "cam0_problem.npz"), obs_x=prob.obs_x, obs_y=prob.obs_y, obs_img=prob.obs_img,
        obs_pt=prob.obs_pt, img_cam=prob.img_cam, eop0=prob.eop0, iop0=prob.iop0, cam_box=prob.cam_box,
        xyz0=prob.xyz0, pt_tie=prob.pt_tie, tie_pt=prob.tie_pt, point_ids=np.array(prob.point_ids),
        image_ids=np.array(prob.image_ids), camera_ids=np.array(prob.camera_ids),
        **{"s_" + k: np.array(v) for k, v in sd.items()})
    err, xhat0, _ = fb.Buildxhat(prob)
    assert err == 0
    eop, iop, xyz = model.gather_params(prob, xhat0)
    ref = refexpr.ReferenceBuildAwG()
    for t, name in enumerate(refexpr.TYPE_NAMES):
        prob.settings.type = name
        q = refexpr.reference_observation_equations(prob, eop, iop, xyz, ref)
        np.savez_compressed(os.path.join(OUT, f"cam0_refsrc_t{t}.npz"),
                            **{k: q[k] for k in ("fx", "fy", "w", "Je", "Jc", "Jt", "G", "scale")})
    for name in ("pinhole", "fisheye"):
        prob.settings.type = name
        out = dense.gauss_newton(prob, xhat0)
        np.savez_compressed(os.path.join(OUT, f"cam0_gn_{name}.npz"), xhat0=xhat0, xhat=out["xhat"],
                            deltasum=np.array(out["deltasum"]), v=out["v"], RSD=out["RSD"],
                            sigma02=out["sigma02"], RMSx=out["RMSx"], RMSy=out["RMSy"], RMS=out["RMS"],
                            delta=out["delta"], iterations=out["iterations"])
        print(name, out["iterations"], out["deltasum"], out["sigma02"])
    prob.settings.type = "fisheye"
    x, trace = xhat0.copy(), []
    while True:
        d = exact.exact_step(prob, x)
        xprev, x = x, x + d
        trace.append(float(np.sum(np.abs(d))))
        if trace[-1] <= prob.settings.threshold or len(trace) >= prob.settings.Iteration_Cap:
            break
    _, A, w, _, _ = model.BuildAwG(prob, xprev)
    v = A @ d + w
    s02 = float(v @ (dense.weights(prob) * v)) / (A.shape[0] - A.shape[1])
    np.savez_compressed(os.path.join(OUT, "cam0_gn_fisheye_exact.npz"), xhat0=xhat0, xhat=x,
                        deltasum=np.array(trace), v=v, sigma02=s02, iterations=len(trace), delta=d,
                        RSD=dense.BuildRSD(prob, v, x))
    print("fisheye exact", len(trace), trace, s02)


if __name__ == "__main__":
    main()
