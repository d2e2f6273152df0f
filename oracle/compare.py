"""ORACLE (test infrastructure, not product code) -- how two solutions are compared.

``group_rel`` is the scaled, group-normalised error of SURVEY.md 7.2-1 used by the parity tests and by the
``parity`` object of bench.py (north-star tolerance: 1e-9 relative on xhat)."""
from __future__ import annotations

import numpy as np

from . import model


def group_rel(prob, a, b):
    """max over parameter groups of ||a_g - b_g|| / ||b_g|| -- the scaled, group-normalised error of
    SURVEY.md 7.2-1.  Groups: EOP positions, EOP angles, (xp, yp, c), radial terms in their scaled
    units K_j r_max^(2j), decentering terms scaled by r_max^2, tie-point coordinates."""
    L = model.layout(prob)
    ui, uc, NK = L["u_img"], L["u_cam"], L["NK"]
    box = prob.cam_box
    rmax2 = ((box[:, 3] - box[:, 1]) * 0.5) ** 2 + ((box[:, 4] - box[:, 2]) * 0.5) ** 2
    a, b = a.copy(), b.copy()
    groups = []
    if ui:
        e = np.arange(L["off_cam"]).reshape(prob.numImg, ui)
        pos = [L["ecols"][q] for q in range(3) if L["ecols"][q] >= 0]
        ang = [L["ecols"][q] for q in range(3, 6) if L["ecols"][q] >= 0]
        if pos:
            groups.append(e[:, pos].ravel())
        if ang:
            groups.append(e[:, ang].ravel())
    cam0 = L["off_cam"] + uc * np.arange(prob.numCam)
    lin = [cam0 + L["ccols"][q] for q in range(3) if L["ccols"][q] >= 0]
    if lin:
        groups.append(np.concatenate(lin))
    if L["ccols"][3] >= 0:
        rad = []
        for j in range(NK):
            idx = cam0 + L["ccols"][3 + j]
            a[idx] *= rmax2 ** (j + 1); b[idx] *= rmax2 ** (j + 1)
            rad.append(idx)
        groups.append(np.concatenate(rad))
    if L["ccols"][3 + NK] >= 0:
        dec = []
        for j in range(2):
            idx = cam0 + L["ccols"][3 + NK + j]
            a[idx] *= rmax2; b[idx] *= rmax2
            dec.append(idx)
        groups.append(np.concatenate(dec))
    if prob.numtie:
        groups.append(np.arange(L["off_tie"], L["u"]))
    worst = 0.0
    for g in groups:
        den = np.linalg.norm(b[g])
        if den > 0:
            worst = max(worst, np.linalg.norm(a[g] - b[g]) / den)
    return worst
