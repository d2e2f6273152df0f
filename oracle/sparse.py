"""ORACLE (test infrastructure, not product code) -- block-sparse / Schur restatement.

Same Gauss-Newton step as ``oracle/dense.py`` (``main.m:412-494``, ``main.m:569-602``) but with
the weight matrix as two scalars and the normal equations in block form, so it scales to the
synthetic configurations the literal dense path cannot hold (an 8 TB ``P`` at config 2).
Independent of the dense path in how ``N`` is formed and solved; the two must agree
(``tests/test_oracle.py``).  Algebra:

    N = [N_cc  W ; W' V],  V = blkdiag(V_p) (3x3 per tie point, exactly block diagonal)
    S = N_cc - W V^-1 W',  g = u_c - W V^-1 u_p                      (point elimination)
    inner constraints: G is non-zero only in EOP rows (BuildAwG.m:514-527) so the border stays
    in the camera block:  M = S + Gc Gc',  (Gc' M^-1 Gc) k = -Gc' M^-1 g,  d_c = -M^-1 (g + Gc k)
    d_p = -V_p^-1 (u_p + W_p' d_c)

Parity status: see ``oracle/model.py``.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

from .model import G_rows, gather_params, layout, observation_equations


def compact_jacobians(prob, L, q):
    """Keep only the estimated columns, in xhat order (BuildAwG.m:217-365, :367-451)."""
    esel = [c for c in range(6) if L["ecols"][c] >= 0]
    csel = [c for c in range(len(L["ccols"])) if L["ccols"][c] >= 0]
    return q["Je"][:, :, esel], q["Jc"][:, :, csel]


def normal_blocks(prob, xhat):
    """Block normal equations at ``xhat``.  Returns dict with N_cc (dense u_c x u_c), u_c vector,
    V (nTie,3,3), u_p (nTie,3), W (scipy CSR u_c x 3 nTie), Gc (u_c x 7 or None), plus the
    per-observation Jacobians for the residual stage."""
    s = prob.settings
    L = layout(prob)
    eop, iop, xyz = gather_params(prob, xhat)
    q = observation_equations(prob, eop, iop, xyz)
    Je, Jc = compact_jacobians(prob, L, q)
    Jt = q["Jt"]
    n_obs, ui, uc = prob.n_obs, L["u_img"], L["u_cam"]
    pw = np.array([1.0 / s.sigma_x ** 2, 1.0 / s.sigma_y ** 2])           # main.m:396-405
    img, cam = prob.obs_img, q["cam"]
    tie = prob.pt_tie[prob.obs_pt]
    u_c = L["off_tie"]
    Jcc = np.concatenate([Je, Jc], axis=2)                                # (n,2,ui+uc)
    PJ = Jcc * pw[None, :, None]
    blk = np.einsum("nra,nrb->nab", PJ, Jcc)                              # per-obs (ui+uc)^2
    rhs = np.einsum("nra,nr->na", PJ, q["w"])
    cols = np.concatenate([ui * img[:, None] + np.arange(ui)[None, :],
                           L["off_cam"] + uc * cam[:, None] + np.arange(uc)[None, :]], axis=1)
    N_cc = np.zeros((u_c, u_c))
    np.add.at(N_cc, (cols[:, :, None], cols[:, None, :]), blk)
    g_c = np.zeros(u_c)
    np.add.at(g_c, cols, rhs)
    nT = prob.numtie
    out = dict(L=L, q=q, Je=Je, Jc=Jc, Jt=Jt, N_cc=N_cc, u_c=g_c, pw=pw, cols=cols, tie=tie, eop=eop)
    if nT:
        ts = np.nonzero(tie >= 0)[0]
        PJt = Jt[ts] * pw[None, :, None]
        V = np.zeros((nT, 3, 3))
        np.add.at(V, tie[ts], np.einsum("nra,nrb->nab", PJt, Jt[ts]))
        u_p = np.zeros((nT, 3))
        np.add.at(u_p, tie[ts], np.einsum("nra,nr->na", PJt, q["w"][ts]))
        Wb = np.einsum("nra,nrb->nab", PJ[ts], Jt[ts])                    # (nts, ui+uc, 3)
        rr = np.repeat(cols[ts][:, :, None], 3, axis=2)
        cc = (3 * tie[ts])[:, None, None] + np.arange(3)[None, None, :] + 0 * rr
        W = sp.coo_matrix((Wb.reshape(-1), (rr.reshape(-1), cc.reshape(-1))),
                          shape=(u_c, 3 * nT)).tocsr()                    # duplicates are summed
        out.update(V=V, u_p=u_p, W=W)
    if s.Inner_Constraints:
        Gi = G_rows(eop)
        Gc = np.zeros((u_c, 7))
        for j in np.unique(img):
            Gc[6 * j:6 * j + 6] = Gi[j]
        out["Gc"] = Gc
    return out


def reduce(prob, nb):
    """Point elimination: S = N_cc - W V^-1 W', g = u_c - W V^-1 u_p (symmetrised).  Returns S, g, Vinv."""
    S = nb["N_cc"].copy()
    g = nb["u_c"].copy()
    nT = prob.numtie
    Vinv = None
    if nT:
        Vinv = np.linalg.inv(nb["V"])
        Vi = sp.bsr_matrix((Vinv, np.arange(nT), np.arange(nT + 1)), shape=(3 * nT, 3 * nT)).tocsr()
        WV = nb["W"] @ Vi
        S -= (WV @ nb["W"].T).toarray()
        g -= WV @ nb["u_p"].reshape(-1)
    S = 0.5 * (S + S.T)
    return S, g, Vinv


def solve_reduced(prob, S, g, Gc=None):
    """(Bordered) solve of the reduced system: returns the scaled camera-part increment d_c."""
    if prob.settings.Inner_Constraints:
        M = S + Gc @ Gc.T
        cf = sla.cho_factor(M, lower=True)
        Y = sla.cho_solve(cf, np.column_stack([g, Gc]))
        k = np.linalg.solve(Gc.T @ Y[:, 1:], -(Gc.T @ Y[:, 0]))
        return -(Y[:, 0] + Y[:, 1:] @ k)
    cf = sla.cho_factor(S, lower=True)
    return -sla.cho_solve(cf, g)


def back_substitute(prob, nb, Vinv, d_c):
    """d_p = -V_p^-1 (u_p + W_p' d_c); returns the full scaled delta."""
    L = nb["L"]
    u_c = L["off_tie"]
    nT = prob.numtie
    delta = np.zeros(L["u"])
    delta[:u_c] = d_c
    if nT:
        t = nb["u_p"] + (nb["W"].T @ d_c).reshape(nT, 3)
        delta[u_c:] = -np.einsum("tab,tb->ta", Vinv, t).reshape(-1)
    return delta


def reduce_and_solve(prob, nb):
    """Schur complement + (bordered) solve.  Returns scaled delta (length u), S, g."""
    S, g, Vinv = reduce(prob, nb)
    d_c = solve_reduced(prob, S, g, nb.get("Gc"))
    return back_substitute(prob, nb, Vinv, d_c), S, g


def unscale(prob, L, q, delta):
    """main.m:458-482: divide the distortion increments by r_max^(2j) / r_max^2."""
    s = prob.settings
    d = delta.copy()
    NK = L["NK"]
    for c in range(prob.numCam):
        base = L["off_cam"] + L["u_cam"] * c
        if s.Estimate_radial:
            for j in range(s.Num_Radial_Distortions):
                d[base + L["ccols"][3] + j] /= q["scale"][c, j]
        if s.Estimate_decent:
            for j in range(2):
                d[base + L["ccols"][3 + NK] + j] /= q["scale"][c, 0]
    return d


def iterate(prob, xhat):
    """One Gauss-Newton step (main.m:416-488).  Returns xhat_new, deltasum, state for residuals."""
    nb = normal_blocks(prob, xhat)
    delta_s, S, g = reduce_and_solve(prob, nb)
    delta = unscale(prob, nb["L"], nb["q"], delta_s)
    deltasum = float(np.sum(np.abs(delta)))                               # main.m:487
    return xhat + delta, deltasum, dict(nb=nb, delta=delta, S=S, g=g, delta_scaled=delta_s)


def residuals(prob, nb, delta, xhat_new):
    """main.m:569 (v = A*delta + w with the last A, w and the UN-scaled delta), BuildRSD,
    main.m:592-602."""
    from .dense import BuildRSD
    L = nb["L"]
    u_c = L["off_tie"]
    dcols = delta[nb["cols"]]                                             # (n, ui+uc)
    Jcc = np.concatenate([nb["Je"], nb["Jc"]], axis=2)
    v = np.einsum("nra,na->nr", Jcc, dcols) + nb["q"]["w"]
    tie = nb["tie"]
    ts = np.nonzero(tie >= 0)[0]
    if ts.size:
        dp = delta[u_c:].reshape(-1, 3)[tie[ts]]
        v[ts] += np.einsum("nra,na->nr", nb["Jt"][ts], dp)
    vv = v.reshape(-1)
    RSD = BuildRSD(prob, vv, xhat_new)
    RMSx = np.sqrt(np.mean(v[:, 0] ** 2)); RMSy = np.sqrt(np.mean(v[:, 1] ** 2))
    vPv = float(np.sum(v[:, 0] ** 2) * nb["pw"][0] + np.sum(v[:, 1] ** 2) * nb["pw"][1])
    sigma02 = vPv / (prob.n - L["u"])                                     # main.m:601
    return dict(v=vv, RSD=RSD, RMSx=RMSx, RMSy=RMSy, RMS=np.sqrt(RMSx ** 2 + RMSy ** 2),
                sigma02=sigma02)


def gauss_newton(prob, xhat0, max_iter=None):
    s = prob.settings
    xhat = np.array(xhat0, dtype=np.float64).copy()
    deltasum, count, trace, st = 100.0, 0, [], None
    cap = s.Iteration_Cap if max_iter is None else max_iter
    while deltasum > s.threshold:
        count += 1
        xhat, deltasum, st = iterate(prob, xhat)
        trace.append(deltasum)
        if count >= cap:
            break
    out = residuals(prob, st["nb"], st["delta"], xhat)
    out.update(xhat=xhat, iterations=count, deltasum=trace, delta=st["delta"])
    return out
