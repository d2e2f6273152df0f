"""ORACLE (test infrastructure, not product code) -- observation equations.

CPU restatement (NumPy, IEEE double) of the per-observation part of the reference's
``functions/BuildAwG.m``: parameter gather, object->camera transform, distortion,
projection for the five models, analytic Jacobian blocks, misclosure and the
inner-constraint rows.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
CPU-baseline / ``--impl reference`` legs may import this package; the product path
(``fish-eye_bundle_adjustment_b200``) never does.

Parity status: the reference is MATLAB and neither MATLAB nor Octave exists in the build
container, and the reference ships no tests or golden outputs.  What this oracle is pinned to:
(1) the reference's own source EXECUTED -- ``oracle/mlab.py`` / ``oracle/refrun.py`` run the whole
``BuildAwG.m`` (and ``Buildxhat.m``, ``BuildRSD.m``, ``sumabs.m``, the loop / residual / statistics
statements of ``main.m``) from ``/root/reference`` with a MATLAB-subset interpreter; this module's
``A``, ``w``, ``G``, ``dist_scaling`` match to 3e-16 / 5e-13 px / 9e-16, the loops built on it to 1e-11
(``tests/test_reference_source_run.py``, frozen in ``tests/golden/*_refrun_*.npz``); (2) expression
level, ``oracle/refexpr.py`` (``tests/test_oracle.py``, ``tests/golden/cam0_refsrc_t*.npz``); (3) the one
shipped output of the reference, ``cam0.int:2`` (``tests/test_oracle.py::test_known_answer_...``).
Not pinned: a MATLAB process (libm / LAPACK differences of a few ulp, amplified by cond(N) after the
explicit inverse).

The Jacobian here is *not* a transcription of the generated expressions: it is the chain
rule through (U,V,W) of the same function, which is what the
generated code differentiates.
"""
from __future__ import annotations

import numpy as np


def layout(prob):
    """Column/row offsets of the unknown vector (Buildxhat.m:22-135, BuildAwG.m:24-25,52,97,110).

    Returns dict with u_img (u_perimage), u_cam (u_percam), ecols (6 entries: slot inside the
    image block or -1), ccols (3+NK+2 entries, order xp yp c k1.. p1 p2: slot inside the camera
    block or -1), off_cam, off_tie, u.
    """
    s = prob.settings
    NK = s.NK
    ecols, k = [], 0
    for f in s.eop_flags:
        ecols.append(k if f else -1)
        k += 1 if f else 0
    u_img = k
    ccols, k = [], 0
    for f in (s.Estimate_xp, s.Estimate_yp, s.Estimate_c):
        ccols.append(k if f else -1)
        k += 1 if f else 0
    for _ in range(NK):
        ccols.append(k if s.Estimate_radial else -1)
        k += 1 if s.Estimate_radial else 0
    for _ in range(2):
        ccols.append(k if s.Estimate_decent else -1)
        k += 1 if s.Estimate_decent else 0
    u_cam = k
    off_cam = u_img * prob.numImg
    off_tie = off_cam + u_cam * prob.numCam
    return dict(u_img=u_img, u_cam=u_cam, ecols=np.array(ecols), ccols=np.array(ccols),
                off_cam=off_cam, off_tie=off_tie, u=off_tie + 3 * prob.numtie, NK=NK)


def gather_params(prob, xhat):
    """Current EOP / IOP / XYZ tables: from ``xhat`` where estimated, else the file values.

    BuildAwG.m:52-93 (EOPs), :96-107 (tie XYZ), :110-155 (IOPs, K, P).
    """
    L = layout(prob)
    eop = prob.eop0.copy()
    iop = prob.iop0.copy()
    xyz = prob.xyz0.copy()
    xe = xhat[:L["off_cam"]].reshape(prob.numImg, L["u_img"]) if L["u_img"] else None
    for q in range(6):
        if L["ecols"][q] >= 0:
            eop[:, q] = xe[:, L["ecols"][q]]
    xc = xhat[L["off_cam"]:L["off_tie"]].reshape(prob.numCam, L["u_cam"]) if L["u_cam"] else None
    for q in range(iop.shape[1]):
        if L["ccols"][q] >= 0:
            iop[:, q] = xc[:, L["ccols"][q]]
    if prob.numtie:
        xt = xhat[L["off_tie"]:].reshape(prob.numtie, 3)
        sel = prob.tie_pt >= 0
        xyz[prob.tie_pt[sel]] = xt[sel]
    return eop, iop, xyz


def rotation(w, p, k):
    """M = R3(k) R2(p) R1(w) as written out in BuildAwG.m:163-165 (rows U, V, W)."""
    cw, sw, cp, sp, ck, sk = np.cos(w), np.sin(w), np.cos(p), np.sin(p), np.cos(k), np.sin(k)
    M = np.empty(np.shape(w) + (3, 3))
    M[..., 0, 0] = ck * cp
    M[..., 0, 1] = cw * sk + ck * sp * sw
    M[..., 0, 2] = sk * sw - ck * cw * sp
    M[..., 1, 0] = -cp * sk
    M[..., 1, 1] = ck * cw - sk * sp * sw
    M[..., 1, 2] = ck * sw + cw * sk * sp
    M[..., 2, 0] = sp
    M[..., 2, 1] = -cp * sw
    M[..., 2, 2] = cp * cw
    return M


def _g_and_dg(typeint, theta, R, W):
    """g(theta), g'(theta) for typeint 0..4 (BuildAwG.m:184-208)."""
    if typeint == 0:
        return theta, np.ones_like(theta)
    if typeint == 1:                      # -c*U/W == -c*(U/R)*tan(atan(R/W))
        t = R / W
        return t, 1.0 + t * t
    if typeint == 2:
        return 2.0 * np.sin(0.5 * theta), np.cos(0.5 * theta)
    if typeint == 3:
        return np.sin(theta), np.cos(theta)
    if typeint == 4:
        t = np.tan(0.5 * theta)
        return 2.0 * t, 1.0 + t * t
    raise ValueError("BuildAwG, invalid type in data.settings.type")


def observation_equations(prob, eop, iop, xyz, idx=None):
    """Per observation: f, w, Jacobian blocks w.r.t. ALL 6 EOPs, ALL 3+NK+2 camera
    parameters (distortion columns pre-scaled as the reference does) and X,Y,Z.

    Returns dict: fx, fy, w (n,2), Je (n,2,6), Jc (n,2,3+NK+2), Jt (n,2,3), scale (nCam,NK)
    = r_max^(2j) (BuildAwG.m:422-426).  ``idx`` restricts to a subset of observations.
    """
    s = prob.settings
    NK = s.NK
    typeint = s.typeint
    sl = slice(None) if idx is None else idx
    x, y = prob.obs_x[sl], prob.obs_y[sl]
    im, pt = prob.obs_img[sl], prob.obs_pt[sl]
    cam = prob.img_cam[im]
    Xc, Yc, Zc, w, p, k = (eop[im, q] for q in range(6))
    X, Y, Z = xyz[pt, 0], xyz[pt, 1], xyz[pt, 2]
    xp, yp, c = iop[cam, 0], iop[cam, 1], iop[cam, 2]
    K = iop[cam, 3:3 + NK]
    P1, P2 = iop[cam, 3 + NK], iop[cam, 4 + NK]
    y_dir = prob.cam_box[cam, 0]

    # BuildAwG.m:163-166
    M = rotation(w, p, k)
    d = np.stack([X - Xc, Y - Yc, Z - Zc], axis=-1)
    UVW = np.einsum("nij,nj->ni", M, d)
    U, V, W = UVW[:, 0], UVW[:, 1], UVW[:, 2]
    R = np.sqrt(U * U + V * V)
    # BuildAwG.m:168-181  (distortion at the OBSERVED coordinates)
    xb, yb = x - xp, y - yp
    r2 = xb * xb + yb * yb
    r = np.sqrt(r2)
    rpow = np.stack([r ** (2 * (j + 1)) for j in range(NK)], axis=-1)      # r^(2j)
    delta_r = np.sum(K * rpow, axis=-1)
    dec_x = P1 * (yb * yb + 3 * xb * xb) + 2 * P2 * xb * yb
    dec_y = P2 * (xb * xb + 3 * yb * yb) + 2 * P1 * xb * yb
    # BuildAwG.m:184-208
    theta = np.arctan(R / W)
    g, dg = _g_and_dg(typeint, theta, R, W)
    sfac = g / R
    fx = -c * U * sfac + xp + delta_r * xb + dec_x
    fy = -c * y_dir * V * sfac + yp + delta_r * yb + dec_y

    # chain rule through (U,V,W)  (checked against BuildAwG.m:223-495)
    D = R * R + W * W
    dth = np.stack([U * W / (R * D), V * W / (R * D), -R / D], axis=-1)
    ds = (dg / R)[:, None] * dth - (g / (R * R))[:, None] * np.stack([U / R, V / R, np.zeros_like(U)], -1)
    Juvw = np.empty((x.shape[0], 2, 3))
    Juvw[:, 0, :] = -c[:, None] * (U[:, None] * ds)
    Juvw[:, 0, 0] += -c * sfac
    Juvw[:, 1, :] = -(c * y_dir)[:, None] * (V[:, None] * ds)
    Juvw[:, 1, 1] += -(c * y_dir) * sfac
    Jt = np.einsum("nab,nbc->nac", Juvw, M)                                 # d f / d(X,Y,Z)
    Je = np.empty((x.shape[0], 2, 6))
    Je[:, :, 0:3] = -Jt                                                     # d f / d(Xc,Yc,Zc)
    cw, sw, cp, sp, ck, sk = np.cos(w), np.sin(w), np.cos(p), np.sin(p), np.cos(k), np.sin(k)
    # d(U,V,W)/d omega = M * (0, dz, -dy)
    d_om = M[:, :, 1] * d[:, 2:3] - M[:, :, 2] * d[:, 1:2]
    # d(U,V,W)/d phi = R3 * dR2/dphi * (R1 d)
    ex = d[:, 0]
    ez = -sw * d[:, 1] + cw * d[:, 2]
    t1 = -sp * ex - cp * ez
    t3 = cp * ex - sp * ez
    d_ph = np.stack([ck * t1, -sk * t1, t3], axis=-1)
    # d(U,V,W)/d kappa = (V, -U, 0)
    d_ka = np.stack([V, -U, np.zeros_like(U)], axis=-1)
    Je[:, :, 3] = np.einsum("nab,nb->na", Juvw, d_om)
    Je[:, :, 4] = np.einsum("nab,nb->na", Juvw, d_ph)
    Je[:, :, 5] = np.einsum("nab,nb->na", Juvw, d_ka)

    # IOP block, BuildAwG.m:367-451 (order xp yp c k1..kNK p1 p2)
    Jc = np.zeros((x.shape[0], 2, 3 + NK + 2))
    j1 = np.arange(1, NK + 1, dtype=np.float64)
    rpm = np.stack([r ** (2 * j) for j in range(NK)], axis=-1)             # r^((j-1)*2)
    sum_K = np.sum(K * rpow, axis=-1)
    sum_2jK = np.sum(2 * j1 * K * rpm, axis=-1)
    Jc[:, 0, 0] = 1 - sum_K - sum_2jK * xb * xb - 6 * P1 * xb - 2 * P2 * yb   # :375-383
    Jc[:, 1, 0] = -sum_2jK * xb * yb - 2 * P1 * yb - 2 * P2 * xb
    Jc[:, 0, 1] = -sum_2jK * xb * yb - 2 * P2 * xb - 2 * P1 * yb              # :388-396
    Jc[:, 1, 1] = 1 - sum_K - sum_2jK * yb * yb - 6 * P2 * yb - 2 * P1 * xb
    Jc[:, 0, 2] = -U * sfac                                                   # :399-418
    Jc[:, 1, 2] = -y_dir * V * sfac
    box = prob.cam_box
    rmax = np.sqrt(((box[:, 3] - box[:, 1]) * 0.5) ** 2 + ((box[:, 4] - box[:, 2]) * 0.5) ** 2)  # :422
    scale = np.stack([rmax ** (2 * (j + 1)) for j in range(NK)], axis=-1)     # :424-426
    sc = scale[cam]
    Jc[:, 0, 3:3 + NK] = rpow * xb[:, None] / sc                              # :428-438
    Jc[:, 1, 3:3 + NK] = rpow * yb[:, None] / sc
    s1 = sc[:, 0]
    Jc[:, 0, 3 + NK] = (yb * yb + 3 * xb * xb) / s1                            # :439-445
    Jc[:, 0, 4 + NK] = 2 * xb * yb / s1
    Jc[:, 1, 3 + NK] = 2 * xb * yb / s1
    Jc[:, 1, 4 + NK] = (xb * xb + 3 * yb * yb) / s1
    wv = np.stack([fx - x, fy - y], axis=-1)                                  # :505-512
    return dict(fx=fx, fy=fy, w=wv, Je=Je, Jc=Jc, Jt=Jt, scale=scale, r=r, cam=cam)


def G_rows(eop):
    """Inner-constraint rows per image from the CURRENT EOPs (BuildAwG.m:514-527): (nImg,6,7)."""
    Xc, Yc, Zc, w, p, _ = (eop[:, q] for q in range(6))
    n = eop.shape[0]
    G = np.zeros((n, 6, 7))
    G[:, 0, 0] = 1; G[:, 0, 4] = -Zc; G[:, 0, 5] = Yc; G[:, 0, 6] = Xc
    G[:, 1, 1] = 1; G[:, 1, 3] = Zc; G[:, 1, 5] = -Xc; G[:, 1, 6] = Yc
    G[:, 2, 2] = 1; G[:, 2, 3] = -Yc; G[:, 2, 4] = Xc; G[:, 2, 6] = Zc
    G[:, 3, 3] = -1; G[:, 3, 4] = -np.sin(w) * np.tan(p); G[:, 3, 5] = np.cos(w) * np.tan(p)
    G[:, 4, 4] = -np.cos(w); G[:, 4, 5] = -np.sin(w)
    G[:, 5, 4] = np.sin(w) / np.cos(p); G[:, 5, 5] = -np.cos(w) / np.cos(p)
    return G


def BuildAwG(prob, xhat):
    """``[error, A, misclosure, G, dist_scaling] = BuildAwG(data, xhat)`` -- DENSE, literal.

    A is n x u (BuildAwG.m:41); rows 2i-1,2i belong to observation i (:356); column blocks at
    u_img*(ext_index-1)+1 (:358), off_cam + u_cam*(cam_num-1)+1 (:448), off_tie + 3*(tieIndex-1)+1
    (:97,501).  dist_scaling: nCam x (2+NK) = [radial index (1-based), decentering index, r_max^2j]
    (:29-32, :138, :150, :424-426).  Small problems only.
    """
    s = prob.settings
    if s.typeint < 0:
        return 1, None, None, None, None
    L = layout(prob)
    eop, iop, xyz = gather_params(prob, xhat)
    q = observation_equations(prob, eop, iop, xyz)
    n_obs, u = prob.n_obs, L["u"]
    A = np.zeros((2 * n_obs, u))
    rows = np.arange(n_obs) * 2
    for col in range(6):
        if L["ecols"][col] >= 0:
            cidx = L["u_img"] * prob.obs_img + L["ecols"][col]
            A[rows, cidx] = q["Je"][:, 0, col]
            A[rows + 1, cidx] = q["Je"][:, 1, col]
    for col in range(3 + L["NK"] + 2):
        if L["ccols"][col] >= 0:
            cidx = L["off_cam"] + L["u_cam"] * q["cam"] + L["ccols"][col]
            A[rows, cidx] = q["Jc"][:, 0, col]
            A[rows + 1, cidx] = q["Jc"][:, 1, col]
    tie = prob.pt_tie[prob.obs_pt]
    sel = np.nonzero(tie >= 0)[0]
    for col in range(3):
        cidx = L["off_tie"] + 3 * tie[sel] + col
        A[rows[sel], cidx] = q["Jt"][sel, 0, col]
        A[rows[sel] + 1, cidx] = q["Jt"][sel, 1, col]
    wv = q["w"].reshape(-1)
    if s.Inner_Constraints:
        G = np.zeros((u, 7))
        Gi = G_rows(eop)
        used = np.unique(prob.obs_img)                  # only images that have observations (:515)
        for j in used:
            G[6 * j:6 * j + 6, :] = Gi[j]
    else:
        G = 0
    dist_scaling = np.zeros((prob.numCam, 2 + L["NK"]))
    cams_used = np.unique(q["cam"])
    if s.Estimate_radial:
        dist_scaling[cams_used, 0] = L["off_cam"] + L["u_cam"] * cams_used + L["ccols"][3] + 1
    if s.Estimate_decent:
        dist_scaling[cams_used, 1] = L["off_cam"] + L["u_cam"] * cams_used + L["ccols"][3 + L["NK"]] + 1
    dist_scaling[cams_used, 2:] = q["scale"][cams_used]
    return 0, A, wv, G, dist_scaling
