"""Problem build: the reference's ``data`` struct as numeric structure-of-arrays.

Mirrors, on the host, what ``main.m`` does between reading the files and entering the
Gauss-Newton loop:

* settings extraction                      main.m:112-177  -> :class:`Settings`
* string -> numeric, degrees -> radians     main.m:196-258
* ``Estimate_AllGCP`` => TIE = unique(PHO)  main.m:260-264
* ``data.points(i)`` index resolution       main.m:277-384  -> :class:`Problem`
* ``Buildxhat``                             functions/Buildxhat.m:2-136

The reference stores one ~30-field struct per observation (AoS, parameters copied into
every record, O(n*m) ``strcmp`` scans).  The boundary type here is the SoA the C-ABI takes
(SURVEY.md section 8b): per-observation ``x, y, image slot, point slot`` plus per-image /
per-camera / per-point tables.  IDs are resolved with hash maps (first match wins, as the
linear scans of main.m:286-291, :310-315, :346-351, :364-369 do).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field, asdict
from typing import List, Optional, Sequence

import numpy as np

from . import formats

TYPE_NAMES = ("fisheye", "pinhole", "equisolid", "orthographic", "stereographic")  # BuildAwG.m:184-208


@dataclass
class Settings:
    """``data.settings`` (main.m:112-177).  Names follow the reference's fields."""

    Iteration_Cap: int = 100
    threshold: float = 1e-6
    Meas_std: float = 1.0
    Meas_std_y: Optional[float] = None      # None <=> data.settings.no_std_y (main.m:129)
    Inner_Constraints: int = 0
    Estimate_Xc: int = 1
    Estimate_Yc: int = 1
    Estimate_Zc: int = 1
    Estimate_w: int = 1
    Estimate_p: int = 1
    Estimate_k: int = 1
    Estimate_xp: int = 0
    Estimate_yp: int = 0
    Estimate_c: int = 0
    Estimate_radial: int = 0
    Num_Radial_Distortions: int = 1
    Estimate_decent: int = 0
    Estimate_tie: int = 0
    Estimate_AllGCP: int = 0
    type: str = "fisheye"
    Check_Points: int = 0
    Output_Filename: str = "output.out"

    @property
    def typeint(self) -> int:
        """``typeint`` of BuildAwG.m:184-214; -1 for an invalid name (error=1 there)."""
        return TYPE_NAMES.index(self.type) if self.type in TYPE_NAMES else -1

    @property
    def eop_flags(self):
        return (self.Estimate_Xc, self.Estimate_Yc, self.Estimate_Zc,
                self.Estimate_w, self.Estimate_p, self.Estimate_k)

    @property
    def NK(self) -> int:
        """BuildAwG.m:18-20 clamps Num_Radial_Distortions to >= 1."""
        return max(int(self.Num_Radial_Distortions), 1)

    @property
    def u_perimage(self) -> int:            # BuildAwG.m:24
        return int(sum(self.eop_flags))

    @property
    def u_percam(self) -> int:              # BuildAwG.m:25
        return int(self.Estimate_c + self.Estimate_xp + self.Estimate_yp
                   + self.Estimate_radial * self.NK + self.Estimate_decent * 2)

    @property
    def sigma_x(self) -> float:
        return float(self.Meas_std)

    @property
    def sigma_y(self) -> float:             # main.m:397-402
        return float(self.Meas_std if self.Meas_std_y is None else self.Meas_std_y)

    def cfg_dict(self) -> dict:
        """The ``name value`` rows ``main.m`` expects in a .cfg (config.cfg:6-43)."""
        d = {
            "Iteration_Cap": self.Iteration_Cap, "Threshold_Value": self.threshold,
            "Meas_std": float(self.Meas_std),
        }
        if self.Meas_std_y is not None:
            d["Meas_std_y"] = float(self.Meas_std_y)
        d.update({
            "Inner_Constraints": self.Inner_Constraints,
            "Estimate_Xc": self.Estimate_Xc, "Estimate_Yc": self.Estimate_Yc,
            "Estimate_Zc": self.Estimate_Zc, "Estimate_Omega": self.Estimate_w,
            "Estimate_Phi": self.Estimate_p, "Estimate_Kappa": self.Estimate_k,
            "Estimate_xp": self.Estimate_xp, "Estimate_yp": self.Estimate_yp,
            "Estimate_c": self.Estimate_c,
            "Estimate_Radial_Distortions": self.Estimate_radial,
            "Num_Radial_Distortions": self.Num_Radial_Distortions,
            "Estimate_Decentering_Distortions": self.Estimate_decent,
            "Estimate_tie": self.Estimate_tie, "Estimate_AllGCP": self.Estimate_AllGCP,
            "Type": self.type, "Check_Points": self.Check_Points,
        })
        return d


def settings_from_cfg(CFG, folder_name: str = "output") -> Optional[Settings]:
    """main.m:112-177.  Returns None where the reference prints 'Error getting settings'."""
    s = Settings()
    v, e = formats.findSetting(CFG, "Output_Filename", 0)
    s.Output_Filename = v if e == 0 else folder_name + ".out"          # main.m:116-120
    v, e = formats.findSetting(CFG, "Meas_std", 0)
    if e > 0:                                                           # main.m:126-128
        s.Meas_std, s.Meas_std_y = 1.0, None
    else:
        s.Meas_std = float(v)
        vy, ey = formats.findSetting(CFG, "Meas_std_y", 0)              # main.m:129
        s.Meas_std_y = float(vy) if ey == 0 else None
    v, e = formats.findSetting(CFG, "Type", 0)
    s.type = v if e == 0 else "fisheye"                                 # main.m:134-137
    v, e = formats.findSetting(CFG, "Check_Points", 0, True)
    s.Check_Points = int(v) if e == 0 else 0                            # main.m:142-145
    err = 0
    s.Iteration_Cap, err = formats.findSetting(CFG, "Iteration_Cap", err)
    s.threshold, err = formats.findSetting(CFG, "Threshold_Value", err)
    for attr, name in (("Inner_Constraints", "Inner_Constraints"),
                       ("Estimate_Xc", "Estimate_Xc"), ("Estimate_Yc", "Estimate_Yc"),
                       ("Estimate_Zc", "Estimate_Zc"), ("Estimate_w", "Estimate_Omega"),
                       ("Estimate_p", "Estimate_Phi"), ("Estimate_k", "Estimate_Kappa"),
                       ("Estimate_c", "Estimate_c"), ("Estimate_xp", "Estimate_xp"),
                       ("Estimate_yp", "Estimate_yp"),
                       ("Estimate_radial", "Estimate_Radial_Distortions")):
        v, err = formats.findSetting(CFG, name, err, True)
        setattr(s, attr, int(v) if v in (0, 1) else v)
    s.Num_Radial_Distortions, err = formats.findSetting(CFG, "Num_Radial_Distortions", err)
    v, err = formats.findSetting(CFG, "Estimate_Decentering_Distortions", err, True)
    s.Estimate_decent = int(v) if v in (0, 1) else v
    v, err = formats.findSetting(CFG, "Estimate_tie", err, True)
    s.Estimate_tie = int(v) if v in (0, 1) else v
    v, err = formats.findSetting(CFG, "Estimate_AllGCP", err, True)
    s.Estimate_AllGCP = int(v) if v in (0, 1) else v
    if err > 0:
        print("Error getting settings")                                 # main.m:173-177
        return None
    s.Iteration_Cap = int(s.Iteration_Cap)
    s.Num_Radial_Distortions = int(s.Num_Radial_Distortions)
    s.threshold = float(s.threshold)
    return s


@dataclass
class Problem:
    """Numeric SoA form of the reference's ``data`` struct (0-based slots).

    ``obs_img[i]``  = ``data.points(i).ext_index - 1``   (main.m:298)
    ``obs_pt[i]``   = ``data.points(i).cnt_index - 1``   (main.m:358)
    ``img_cam[j]``  = ``cam_num - 1`` of EXT row j        (main.m:322)
    ``pt_tie[p]``   = ``tieIndex - 1`` or -1              (main.m:362-375)
    ``tie_pt[t]``   = CNT row of TIE entry t              (Buildxhat.m:110-122)
    ``eop0``  EXT cols 3..8, angles in radians            (main.m:209-218)
    ``iop0``  INT row 2: xp yp c k1..kNK p1 p2            (main.m:325-330)
    ``cam_box`` INT row 1: y_dir xmin ymin xmax ymax      (main.m:331-343)
    """

    settings: Settings
    obs_x: np.ndarray
    obs_y: np.ndarray
    obs_img: np.ndarray
    obs_pt: np.ndarray
    img_cam: np.ndarray
    eop0: np.ndarray
    iop0: np.ndarray
    cam_box: np.ndarray
    xyz0: np.ndarray
    pt_tie: np.ndarray
    tie_pt: np.ndarray
    point_ids: Optional[Sequence[str]] = None
    image_ids: Optional[Sequence[str]] = None
    camera_ids: Optional[Sequence[str]] = None

    # ---- counts (main.m:379-383)
    @property
    def n_obs(self) -> int:
        return int(self.obs_x.shape[0])

    @property
    def n(self) -> int:
        return 2 * self.n_obs

    @property
    def numImg(self) -> int:
        return int(self.eop0.shape[0])

    @property
    def numCam(self) -> int:
        return int(self.iop0.shape[0])

    @property
    def numPts(self) -> int:
        return int(self.xyz0.shape[0])

    @property
    def numtie(self) -> int:
        return int(self.tie_pt.shape[0])

    # ---- xhat layout (Buildxhat.m:22-135)
    @property
    def u_c(self) -> int:
        s = self.settings
        return s.u_perimage * self.numImg + s.u_percam * self.numCam

    @property
    def u(self) -> int:
        return self.u_c + 3 * self.numtie

    def point_name(self, p: int) -> str:
        return str(self.point_ids[p]) if self.point_ids is not None else f"P{p}"

    def image_name(self, j: int) -> str:
        return str(self.image_ids[j]) if self.image_ids is not None else f"{j}"

    def camera_name(self, c: int) -> str:
        return str(self.camera_ids[c]) if self.camera_ids is not None else f"{c}"

    def validate(self) -> None:
        """Conditions the reference needs but does not check."""
        s = self.settings
        if s.typeint < 0:
            raise ValueError("BuildAwG, invalid type in data.settings.type")   # BuildAwG.m:209-213
        if s.Inner_Constraints and s.u_perimage != 6:
            raise ValueError("Inner_Constraints needs all six EOPs estimated "
                             "(Gblock is always 6 rows, BuildAwG.m:516-525)")
        if self.iop0.shape[1] != 3 + s.NK + 2:
            raise ValueError("iop0 must have 3+NK+2 columns")
        if self.n_obs and (self.obs_img.min() < 0 or self.obs_img.max() >= self.numImg):
            raise ValueError("obs_img out of range")
        if self.n_obs and (self.obs_pt.min() < 0 or self.obs_pt.max() >= self.numPts):
            raise ValueError("obs_pt out of range")
        if not np.all(np.abs(self.cam_box[:, 0]) == 1.0):
            raise ValueError("y_dir should be +-1 only")                       # main.m:334-337


def Buildxhat(prob: Problem):
    """``[error, xhat, xhatnames] = Buildxhat(data, EXT, INT, TIE, CNT)`` (Buildxhat.m:2-136).

    Order: per image (EXT row order) the estimated ones of Xc Yc Zc w p k; per camera (INT
    order) the estimated ones of xp yp c k1..kNK p1 p2; per TIE entry X Y Z.
    """
    s = prob.settings
    xhat: List[float] = []
    names: List[str] = []
    enames = ("Xc", "Yc", "Zc", "w", "p", "k")
    for j in range(prob.numImg):                                           # Buildxhat.m:22-62
        img, cam = prob.image_name(j), prob.camera_name(int(prob.img_cam[j]))
        for q in range(6):
            if s.eop_flags[q]:
                xhat.append(float(prob.eop0[j, q]))
                names.append(f"{enames[q]}_{img}_{cam}")
    NK = s.Num_Radial_Distortions
    for c in range(prob.numCam):                                           # Buildxhat.m:65-105
        cam = prob.camera_name(c)
        row = prob.iop0[c]
        if s.Estimate_xp:
            xhat.append(float(row[0])); names.append(f"xp_{cam}")
        if s.Estimate_yp:
            xhat.append(float(row[1])); names.append(f"yp_{cam}")
        if s.Estimate_c:
            xhat.append(float(row[2])); names.append(f"c_{cam}")
        if s.Estimate_radial:
            for j in range(NK):
                xhat.append(float(row[3 + j])); names.append(f"k{j + 1}_{cam}")
        if s.Estimate_decent:
            for j in range(2):
                xhat.append(float(row[3 + NK + j])); names.append(f"p{j + 1}_{cam}")
    for t in range(prob.numtie):                                           # Buildxhat.m:108-135
        p = int(prob.tie_pt[t])
        if p < 0:
            print(f"Error Buildxhat(): can't find tie {t} from .tie in .cnt")
            return 1, None, None
        xhat.extend(float(v) for v in prob.xyz0[p])
        pid = prob.point_name(p)
        names.extend((f"X_{pid}", f"Y_{pid}", f"Z_{pid}"))
    return 0, np.asarray(xhat, dtype=np.float64), names


# ------------------------------------------------------------------- file loading


def _first_index(ids: Sequence[str]) -> dict:
    m = {}
    for k, v in enumerate(ids):
        m.setdefault(v, k)          # linear scans in main.m stop at the first match
    return m


def load_problem(folder: str, cfg_folder: Optional[str] = None) -> Optional[Problem]:
    """main.m:60-384 for a data folder.  ``cfg_folder``: where the .cfg lives when the data
    folder has none (main.m:66-85 falls back to the project directory)."""
    cfg_dir = folder
    if not any(f.endswith(".cfg") for f in os.listdir(folder)):
        cfg_dir = cfg_folder if cfg_folder is not None else folder
    term, files = formats.ReadFiles([".cfg"], cfg_dir)
    if term:
        print("Error reading files"); return None
    CFG = files[0]
    term, files = formats.ReadFiles([".pho", ".ext", ".cnt", ".int"], folder)
    if term:
        print("Error reading files"); return None
    PHO, EXT, CNT, INT = files
    s = settings_from_cfg(CFG, os.path.basename(os.path.abspath(folder)))
    if s is None:
        return None
    TIE: List[str] = []
    if s.Estimate_tie == 1 and s.Estimate_AllGCP == 0:                    # main.m:180-188
        term, files = formats.ReadFiles([".tie"], folder)
        if term:
            print("Error reading files"); return None
        TIE = [r[0] for r in files[0]]
    d2 = formats.str2double
    pho_pt = [r[0] for r in PHO]; pho_img = [r[1] for r in PHO]
    obs_x = np.array([d2(r[2]) for r in PHO], dtype=np.float64)
    obs_y = np.array([d2(r[3]) for r in PHO], dtype=np.float64)
    image_ids = [r[0] for r in EXT]; ext_cam = [r[1] for r in EXT]
    eop0 = np.array([[d2(r[2]), d2(r[3]), d2(r[4]),
                      d2(r[5]) * math.pi / 180, d2(r[6]) * math.pi / 180,
                      d2(r[7]) * math.pi / 180] for r in EXT], dtype=np.float64)   # main.m:209-218
    point_ids = [r[0] for r in CNT]
    xyz0 = np.array([[d2(r[1]), d2(r[2]), d2(r[3])] for r in CNT], dtype=np.float64)
    NK = s.Num_Radial_Distortions
    camera_ids, box, iop = [], [], []
    for i in range(0, len(INT), 2):                                        # main.m:231-256
        r1, r2 = INT[i], INT[i + 1]
        camera_ids.append(r1[0])
        box.append([d2(r1[k]) for k in range(1, 6)])
        row = [d2(r2[k]) for k in range(3)]
        for k in range(3, 5 + NK):
            row.append(d2(r2[k]) if k < len(r2) and r2[k] is not None else 0.0)
        iop.append(row)
    cam_box = np.array(box, dtype=np.float64); iop0 = np.array(iop, dtype=np.float64)
    if s.Estimate_AllGCP == 1:                                             # main.m:260-264
        TIE = sorted(set(pho_pt)); s.Estimate_tie = 1
    ext_map, cam_map, cnt_map, tie_map = (_first_index(image_ids), _first_index(camera_ids),
                                          _first_index(point_ids), _first_index(TIE))
    for nm, ids, mp in (("image", pho_img, ext_map), ("target", pho_pt, cnt_map)):
        missing = [i for i in ids if i not in mp]
        if missing:
            print(f"Could not find {nm} {missing[0]} from .pho")            # main.m:293-297,352-356
            return None
    obs_img = np.array([ext_map[i] for i in pho_img], dtype=np.int32)
    obs_pt = np.array([cnt_map[i] for i in pho_pt], dtype=np.int32)
    if any(c not in cam_map for c in ext_cam):
        print("Could not find camera from .ext in .int"); return None       # main.m:316-320
    img_cam_all = np.array([cam_map[c] for c in ext_cam], dtype=np.int32)
    numImg = len(set(pho_img))                                             # main.m:379
    numCam = len(set(ext_cam[j] for j in obs_img))                        # main.m:380
    if obs_img.max() >= numImg or img_cam_all[:numImg].max() >= numCam:
        # Buildxhat.m:22-30 takes EXT rows 1..numImg and INT cameras 1..numCam as the slots
        print("EXT/INT must list exactly the images/cameras used in PHO, first"); return None
    pt_tie = np.full(len(point_ids), -1, dtype=np.int32)
    used = set(pho_pt)
    for t, name in enumerate(TIE):
        # isTie is decided per observation by the target ID (main.m:362-375)
        if name in cnt_map and name in used and tie_map[name] == t:
            pt_tie[cnt_map[name]] = t
    tie_pt = np.array([cnt_map.get(name, -1) for name in TIE], dtype=np.int32)
    prob = Problem(settings=s, obs_x=obs_x, obs_y=obs_y, obs_img=obs_img, obs_pt=obs_pt,
                   img_cam=img_cam_all[:numImg].copy(), eop0=eop0[:numImg].copy(),
                   iop0=iop0[:numCam].copy(), cam_box=cam_box[:numCam].copy(), xyz0=xyz0,
                   pt_tie=pt_tie, tie_pt=tie_pt.reshape(-1), point_ids=point_ids,
                   image_ids=image_ids[:numImg], camera_ids=camera_ids[:numCam])
    return prob


def save_problem(prob: Problem, folder: str, stem: str = "net") -> None:
    """Write a problem in the reference's five text formats + .cfg (so main.m can read it)."""
    os.makedirs(folder, exist_ok=True)
    pid = [prob.point_name(p) for p in range(prob.numPts)]
    iid = [prob.image_name(j) for j in range(prob.numImg)]
    cid = [prob.camera_name(c) for c in range(prob.numCam)]
    formats.write_pho(os.path.join(folder, stem + ".pho"), [pid[p] for p in prob.obs_pt],
                      [iid[j] for j in prob.obs_img], prob.obs_x, prob.obs_y)
    formats.write_ext(os.path.join(folder, stem + ".ext"), iid,
                      [cid[c] for c in prob.img_cam], prob.eop0)
    formats.write_cnt(os.path.join(folder, stem + ".cnt"), pid, prob.xyz0)
    formats.write_int(os.path.join(folder, stem + ".int"), cid, prob.cam_box, prob.iop0)
    if prob.numtie and not prob.settings.Estimate_AllGCP:
        formats.write_tie(os.path.join(folder, stem + ".tie"), [pid[p] for p in prob.tie_pt])
    formats.write_cfg(os.path.join(folder, stem + ".cfg"), prob.settings.cfg_dict())
