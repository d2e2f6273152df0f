"""ORACLE (test infrastructure, not product code) -- literal dense Gauss-Newton loop.

Statement-by-statement NumPy restatement of ``main.m:396-494`` (weights, loop, bordered
inverse, un-scaling, update, stop test), ``main.m:569`` (residuals), ``functions/BuildRSD.m``
and ``main.m:592-602`` (RMS, variance factor), including the quirks listed in SURVEY.md
section 8a-Q.  Dense n x u design matrix and explicit inverse, exactly like the reference, so
it is for the bundled cam0 data and small synthetic networks only.  P is kept as its diagonal
(the reference stores the same diagonal densely, main.m:398-405).

Parity status: see ``oracle/model.py``.  Loop level: checked against the reference's own loop
statements executed by ``oracle/refrun.py`` (xhat 7e-12, v 3e-12 px, sigma02 5e-14 on cam0;
``tests/test_reference_source_run.py``); no MATLAB process is available.
"""
from __future__ import annotations

import numpy as np

from .model import BuildAwG, layout


def weights(prob):
    """diag(P), main.m:396-405: 1/sigma_x^2 on odd rows, 1/sigma_y^2 on even rows."""
    s = prob.settings
    cl = np.tile(np.array([s.sigma_x ** 2, s.sigma_y ** 2]), prob.n_obs)
    return 1.0 / cl


def sumabs(vect):
    """functions/sumabs.m:2-15 -- sequential sum of |v_i|."""
    total = 0.0
    for v in np.asarray(vect).reshape(-1):
        total += abs(float(v))
    return total


def solve_step(prob, A, w, G, dist_scaling, Pd):
    """One pass of main.m:424-482: returns (delta un-scaled, Cx with un-scaled diagonal)."""
    s = prob.settings
    u = A.T @ (Pd * w)                                                    # main.m:424
    N = A.T @ (Pd[:, None] * A)                                           # main.m:425
    nu = N.shape[0]
    if s.Inner_Constraints:                                               # main.m:428-440
        NG = np.block([[N, G], [G.T, np.zeros((G.shape[1], G.shape[1]))]])
        uG = np.concatenate([u, np.zeros(G.shape[1])])
        Cx = np.linalg.inv(NG)
        delta = -(Cx @ uG)[:nu]
        Cx = Cx[:nu, :nu]
    else:                                                                 # main.m:441-444
        Cx = np.linalg.inv(N)
        delta = -Cx @ u
    Cx = Cx.copy()
    dg = np.sqrt(np.diag(Cx).astype(complex))                             # main.m:446-456
    Corr = (Cx / (dg[:, None] * dg[None, :])).real                        # (before un-scaling)
    NK = s.Num_Radial_Distortions
    for i in range(dist_scaling.shape[0]):                                # main.m:460-482
        if s.Estimate_radial:
            ri = int(dist_scaling[i, 0]) - 1
            for j in range(NK):
                delta[ri + j] /= dist_scaling[i, j + 2]
                Cx[ri + j, ri + j] /= dist_scaling[i, j + 2] ** 2
        if s.Estimate_decent:
            di = int(dist_scaling[i, 1]) - 1
            for j in range(2):
                delta[di + j] /= dist_scaling[i, 2]
                Cx[di + j, di + j] /= dist_scaling[i, 2] ** 2
    return delta, Cx, N, u, Corr


def BuildRSD(prob, v, xhat):
    """functions/BuildRSD.m:9-42 -> columns r, vx, vy, vr, vt (n_obs x 5).

    xp, yp come from the POST-update xhat (BuildRSD.m:14-27) or the file values.
    """
    s = prob.settings
    L = layout(prob)
    cam = prob.img_cam[prob.obs_img]
    xp = prob.iop0[cam, 0].copy()
    yp = prob.iop0[cam, 1].copy()
    base = L["off_cam"] + L["u_cam"] * cam
    if s.Estimate_xp:
        xp = xhat[base + L["ccols"][0]]
    if s.Estimate_yp:
        yp = xhat[base + L["ccols"][1]]
    vx, vy = v[0::2], v[1::2]
    xbar, ybar = prob.obs_x - xp, prob.obs_y - yp
    theta = np.arctan2(ybar, xbar)
    Phi = np.arctan2(vy, vx)
    v_dist = np.sqrt(vx ** 2 + vy ** 2)
    vr = v_dist * np.cos(theta - Phi)
    vt = v_dist * np.sin(theta - Phi)
    r = np.sqrt(xbar ** 2 + ybar ** 2)
    return np.stack([r, vx, vy, vr, vt], axis=-1)


def gauss_newton(prob, xhat0, max_iter=None, want_cov=True):
    """main.m:407-494 + 567-602.  Returns a dict of everything the report stage consumes."""
    s = prob.settings
    Pd = weights(prob)
    xhat = np.array(xhat0, dtype=np.float64).copy()
    deltasum, count, trace = 100.0, 0, []
    cap = s.Iteration_Cap if max_iter is None else max_iter
    A = w = delta = Cx = Corr = None
    while deltasum > s.threshold:                                         # main.m:412
        count += 1
        err, A, w, G, dist_scaling = BuildAwG(prob, xhat)                 # main.m:416
        if err:
            raise RuntimeError("Error building A and w")
        delta, Cx, _, _, Corr = solve_step(prob, A, w, G, dist_scaling, Pd)
        xhat = xhat + delta                                               # main.m:484
        deltasum = sumabs(delta)                                          # main.m:487
        trace.append(deltasum)
        if count >= cap:                                                  # main.m:490-493
            break
    v = A @ delta + w                                                     # main.m:569
    RSD = BuildRSD(prob, v, xhat)                                         # main.m:571
    RMSx = np.sqrt(np.mean(v[0::2] ** 2))                                 # main.m:594-598
    RMSy = np.sqrt(np.mean(v[1::2] ** 2))
    RMS = np.sqrt(RMSx ** 2 + RMSy ** 2)
    sigma02 = float(v @ (Pd * v)) / (A.shape[0] - A.shape[1])             # main.m:601
    out = dict(xhat=xhat, iterations=count, deltasum=trace, v=v, RSD=RSD,
               RMSx=RMSx, RMSy=RMSy, RMS=RMS, sigma02=sigma02, delta=delta)
    if want_cov:
        out["Cx"] = sigma02 * Cx                                          # main.m:602
        out["Correlation"] = Corr
    return out
