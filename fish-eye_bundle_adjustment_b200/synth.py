"""Deterministic synthetic fish-eye networks (BASELINE.json configs 2-5).

The reference ships one data set (cam0, 1,029 observations) and no generator; the larger
configurations named in BASELINE.json are produced here.  Recipe:

* camera = the bundled one: sensor box 0 0 2448 2048, y_dir -1, xp yp c of ``cam0.int:2``,
  equidistant model; K/P of fish-eye magnitude (NOT the pinhole-fitted ones of cam0.int);
* cameras on a regular grid above a point field with relief, looking down (W < 0 for every
  kept ray), small random tilts, random kappa; object box scaled like cam0 (0..6,500 mm);
* each point is observed by its ``m`` nearest images in which it projects inside the sensor
  box with incidence <= 67 deg (cam0's maximum) and R > 0;
* observations = exact forward model solved for the observed (x, y) (the distortion is a
  function of the observed coordinates, BuildAwG.m:168-181, so a fixed-point iteration) plus
  N(0, sigma^2) noise; initial values = truth + small perturbation so the first step is small.

RNG: ``numpy.random.default_rng(seed)``.  Output is a :class:`Problem` (and, through
``save_problem``, the five text files + .cfg that ``main.m`` reads).
"""
from __future__ import annotations

import math

import numpy as np

from .problem import Problem, Settings

CAM_BOX = (-1.0, 0.0, 0.0, 2448.0, 2048.0)                 # cam0.int:1
XP, YP, C = 1207.903, 1013.724, 1234.758                    # cam0.int:2
K_TRUE = (-4.5e-9, 1.3e-15, 3.6e-22, -9.6e-29, 1.8e-35)     # fish-eye magnitude
P_TRUE = (1.2e-7, 3.7e-7)
MAX_INCIDENCE = math.radians(67.0)

# seeds for BASELINE.json configs 2..5
SEEDS = {2: 20200502, 3: 20200503, 4: 20200504, 5: 20200505}


def _rotation(w, p, k):
    cw, sw, cp, sp, ck, sk = np.cos(w), np.sin(w), np.cos(p), np.sin(p), np.cos(k), np.sin(k)
    M = np.empty(w.shape + (3, 3))
    M[..., 0, 0] = ck * cp; M[..., 0, 1] = cw * sk + ck * sp * sw; M[..., 0, 2] = sk * sw - ck * cw * sp
    M[..., 1, 0] = -cp * sk; M[..., 1, 1] = ck * cw - sk * sp * sw; M[..., 1, 2] = ck * sw + cw * sk * sp
    M[..., 2, 0] = sp; M[..., 2, 1] = -cp * sw; M[..., 2, 2] = cp * cw
    return M


def _project_truth(eop, M, xyz, iop, y_dir, iters=6):
    """Forward model of BuildAwG.m:163-187 solved for the observed (x, y)."""
    d = xyz - eop[:, 0:3]
    U = np.einsum("nj,nj->n", M[:, 0, :], d)
    V = np.einsum("nj,nj->n", M[:, 1, :], d)
    W = np.einsum("nj,nj->n", M[:, 2, :], d)
    R = np.sqrt(U * U + V * V)
    with np.errstate(divide="ignore", invalid="ignore"):
        theta = np.arctan(R / W)
        bx = -iop[2] * U / R * theta
        by = -iop[2] * y_dir * V / R * theta
    NK = len(iop) - 5
    K = iop[3:3 + NK]; P1, P2 = iop[3 + NK], iop[4 + NK]
    x = bx + iop[0]; y = by + iop[1]
    for _ in range(iters):
        xb, yb = x - iop[0], y - iop[1]
        r2 = xb * xb + yb * yb
        dr = np.zeros_like(x); rp = np.ones_like(x)
        for j in range(NK):
            rp = rp * r2
            dr = dr + K[j] * rp
        x = bx + iop[0] + dr * xb + P1 * (yb * yb + 3 * xb * xb) + 2 * P2 * xb * yb
        y = by + iop[1] + dr * yb + P2 * (xb * xb + 3 * yb * yb) + 2 * P1 * xb * yb
    return x, y, theta, R, W


def make_network(n_img: int, n_pts: int, m: int, seed: int, *, mode: str = "free",
                 NK: int = 5, sigma: float = 0.3, extent: float = 6500.0,
                 n_control: int = 0, order: str = "image", type: str = "fisheye",
                 threshold: float = 1e-6, perturb: float = 1.0) -> Problem:
    """Build a synthetic network.

    ``mode``: ``"eop"``  -- EOP-only, every point a fixed control point (config 2);
              ``"free"`` -- inner constraints, IOP + radial + decentering estimated, every point
                            a tie point through ``Estimate_AllGCP`` (configs 3, 4, 5);
              ``"mixed"``-- IOPs estimated, ``n_control`` fixed control points, the rest tie
                            points listed in TIE, no inner constraints.
    ``order``: observation order of the PHO table, ``"image"`` (image-major like cam0.pho) or
              ``"point"``.
    """
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(seed)
    gx = int(math.ceil(math.sqrt(n_img)))
    gy = int(math.ceil(n_img / gx))
    d = extent / gx
    H = d * math.sqrt(1.4 * m / math.pi) / 0.75
    ii = np.arange(n_img)
    cx = (ii % gx + 0.5) * d + rng.normal(0, 0.05 * d, n_img)
    cy = (ii // gx + 0.5) * d + rng.normal(0, 0.05 * d, n_img)
    cz = H * (1.0 + rng.uniform(-0.05, 0.05, n_img))
    eop_t = np.stack([cx, cy, cz, rng.normal(0, 0.08, n_img), rng.normal(0, 0.08, n_img),
                      rng.uniform(-math.pi, math.pi, n_img)], axis=-1)
    iop_t = np.array([XP, YP, C, *K_TRUE[:NK], *([0.0] * max(0, NK - 5)), *P_TRUE])
    M_all = _rotation(eop_t[:, 3], eop_t[:, 4], eop_t[:, 5])
    tree = cKDTree(eop_t[:, 0:2])
    kq = min(n_img, 2 * m + 12)
    y_dir = CAM_BOX[0]

    xyz_list, obs = [], []            # obs: (pt_local, img, x, y)
    need, got = n_pts, 0
    while got < n_pts:
        nb = int((need - got) * 1.15) + 64
        pts = np.stack([rng.uniform(0, gx * d, nb), rng.uniform(0, gy * d, nb),
                        rng.uniform(0, 0.3 * H, nb)], axis=-1)
        _, nbr = tree.query(pts[:, 0:2], k=kq)
        nbr = nbr.reshape(nb, kq)
        pi = np.repeat(np.arange(nb), kq)
        im = nbr.reshape(-1)
        x, y, theta, R, W = _project_truth(eop_t[im], M_all[im], pts[pi], iop_t, y_dir)
        ok = ((W < 0) & (np.abs(theta) <= MAX_INCIDENCE) & (R > 1e-6)
              & (x > CAM_BOX[1] + 2) & (x < CAM_BOX[3] - 2) & (y > CAM_BOX[2] + 2) & (y < CAM_BOX[4] - 2))
        ok = ok.reshape(nb, kq)
        rank = np.cumsum(ok, axis=1)
        keep = ok & (rank <= m)                    # the m nearest in-view images
        cnt = keep.sum(axis=1)
        good = np.nonzero(cnt >= min(3, m))[0][: n_pts - got]
        remap = -np.ones(nb, dtype=np.int64); remap[good] = got + np.arange(good.size)
        kf = keep.reshape(-1) & (remap[pi] >= 0)
        obs.append((remap[pi][kf], im[kf], x[kf], y[kf]))
        xyz_list.append(pts[good])
        got += good.size
    xyz_t = np.concatenate(xyz_list)
    o_pt = np.concatenate([o[0] for o in obs]).astype(np.int32)
    o_im = np.concatenate([o[1] for o in obs]).astype(np.int32)
    o_x = np.concatenate([o[2] for o in obs]) + rng.normal(0, sigma, o_pt.size)
    o_y = np.concatenate([o[3] for o in obs]) + rng.normal(0, sigma, o_pt.size)
    if order == "image":
        perm = np.argsort(o_im, kind="stable")
    else:
        perm = np.argsort(o_pt, kind="stable")
    o_pt, o_im, o_x, o_y = o_pt[perm], o_im[perm], o_x[perm], o_y[perm]
    used = np.zeros(n_img, dtype=bool); used[o_im] = True
    if not used.all():
        raise RuntimeError("an image has no observations; increase n_pts or decrease n_img")

    s = Settings(Iteration_Cap=100, threshold=threshold, Meas_std=sigma, Meas_std_y=None, type=type,
                 Num_Radial_Distortions=NK)
    eop0 = eop_t.copy()
    eop0[:, 0:3] += rng.normal(0, 1.0 * perturb, (n_img, 3))
    eop0[:, 3:6] += rng.normal(0, 1e-4 * perturb, (n_img, 3))
    iop0 = iop_t.copy()[None, :]
    xyz0 = xyz_t.copy()
    pt_tie = np.full(n_pts, -1, dtype=np.int32)
    tie_pt = np.zeros(0, dtype=np.int32)
    point_ids = [f"T{p:07d}" for p in range(n_pts)]
    if mode == "eop":
        s.Inner_Constraints = 0
    else:
        s.Estimate_xp = s.Estimate_yp = s.Estimate_c = 1
        s.Estimate_radial = s.Estimate_decent = 1
        s.Estimate_tie = 1
        iop0[0, 0:3] += rng.normal(0, 0.1 * perturb, 3)
        if mode == "free":
            s.Inner_Constraints = 1
            s.Estimate_AllGCP = 1
            # TIE = unique(PHO(:,1)) -> sorted ids; ids are zero padded so that is slot order
            tie_pt = np.arange(n_pts, dtype=np.int32)
            pt_tie = np.arange(n_pts, dtype=np.int32)
        elif mode == "mixed":
            s.Inner_Constraints = 0
            ctrl = rng.choice(n_pts, size=n_control, replace=False)
            is_tie = np.ones(n_pts, dtype=bool); is_tie[ctrl] = False
            tie_pt = np.nonzero(is_tie)[0].astype(np.int32)
            pt_tie[tie_pt] = np.arange(tie_pt.size, dtype=np.int32)
        else:
            raise ValueError(mode)
        tsel = pt_tie >= 0
        xyz0[tsel] += rng.normal(0, 1.0 * perturb, (int(tsel.sum()), 3))
    prob = Problem(settings=s, obs_x=o_x, obs_y=o_y, obs_img=o_im, obs_pt=o_pt,
                   img_cam=np.zeros(n_img, dtype=np.int32), eop0=eop0, iop0=iop0,
                   cam_box=np.array([CAM_BOX]), xyz0=xyz0, pt_tie=pt_tie, tie_pt=tie_pt,
                   point_ids=point_ids, image_ids=[str(1000 + j) for j in range(n_img)],
                   camera_ids=["0"])
    prob.truth = dict(eop=eop_t, iop=iop_t, xyz=xyz_t)
    return prob


def baseline_config(idx: int, scale: float = 1.0, block: int = 0) -> Problem:
    """BASELINE.json ``configs[idx]`` (idx 1..4; idx 0 is the bundled cam0 data on disk).

    ``scale`` shrinks images and points together for tests (1.0 = the named size).
    """
    if idx == 1:    # EOP-only, fixed control points: 50 images, 20k points, ~500k observations
        return make_network(max(8, int(50 * scale)), max(200, int(20000 * scale)), 25,
                            SEEDS[2], mode="eop")
    if idx == 2:    # free network, IOP + distortion: 500 images, 200k points, ~5M observations
        return make_network(max(8, int(500 * scale)), max(200, int(200000 * scale)), 25,
                            SEEDS[3], mode="free")
    if idx == 3:    # large network: 2,000 images, 1M points, ~10M observations
        return make_network(max(8, int(2000 * scale)), max(200, int(1000000 * scale)), 10,
                            SEEDS[4], mode="free")
    if idx == 4:    # one block of the BatchRun sweep: 200 images, 20k points, ~200k observations
        return make_network(max(8, int(200 * scale)), max(200, int(20000 * scale)), 10,
                            SEEDS[5] + block, mode="free")
    raise ValueError("idx must be 1..4")
