"""The reference's report stage: the ``.out`` file of main.m:604-950 and the check-point
differences of main.m:604-627, written from the outputs of the CUDA path.

Host code that runs once after the loop.  Everything numeric comes from the device results: ``xhat``,
``sqrt(diag(Cx))`` for all unknowns (``feba_cov_diag`` x sigma02, main.m:602), the IOP correlation
sub-matrix per camera and the EOP+IOP correlation block per image (``feba_cov_block``, normalised as
main.m:446-456), ``RSD`` (corrected measurements x + vx, y + vy, main.m:587-590), RMS and sigma02.
Formats follow the reference's ``fprintf`` patterns (``printEOP`` / ``printDist`` / ``printTIE`` at
main.m:972-980, ``printCell.m``) so that a file written here can be diffed against a MATLAB run.

Known quirks of the reference that are kept (they shape the file): the column width looks at image IDs
only when they are longer than the longest target ID (main.m:711-717); the labels of the
EOP/IOP mean-correlation table carry an empty entry between the EOP and IOP names
(main.m:841-850: ``names`` already starts with the padding cell when it is appended).
``num2str(x)`` is ``%d`` for integers and ``%.Ng`` with N = max(floor(log10|x|)+5, 5) otherwise (MATLAB's rule
for scalars: four digits after the leading ones).
"""
from __future__ import annotations

import datetime
import math
from typing import List, Optional, Sequence

import numpy as np

from . import formats
from .problem import Problem

LINE = "*" * 109                                                            # main.m:632
EOP_LABELS = ("Xc", "Yc", "Zc", "Omega", "Phi", "Kappa")                    # main.m:744-776


def num2str(v, digits: Optional[int] = None) -> str:
    """MATLAB ``num2str(x)`` / ``num2str(x, n)`` for scalars (approximation, see module docstring)."""
    if isinstance(v, str):
        return v
    v = float(v)
    if digits is not None:
        return "%.*g" % (digits, v)
    if math.isfinite(v) and v == round(v):
        return "%d" % int(v)
    if not math.isfinite(v) or v == 0:
        return "%g" % v
    return "%.*g" % (min(max(int(math.floor(math.log10(abs(v)))) + 5, 5), 16), v)


def print_cell(rows: Sequence[Sequence], prefix: str = "", padding: int = 4) -> str:
    """``printCell(fileID, Mcell, prefix, padding)`` (functions/printCell.m:1-40)."""
    width = max((len(r[0]) for r in rows), default=0)
    out = []
    for name, val in rows:
        if name == "\\line":
            out.append("-" * (width + padding + 6))
        elif name == "\\n":
            out.append("")
        else:
            out.append(prefix + name + " " + "." * (width + padding - len(name)) + " " + num2str(val))
    return "\n".join(out) + "\n"


def column_width(prob: Problem) -> int:
    """main.m:700-720 as coded: ``img_width`` only moves for image IDs longer than ``target_width``."""
    used_pts = sorted({prob.point_name(int(p)) for p in np.unique(prob.obs_pt)})
    used_img = sorted({prob.image_name(int(j)) for j in np.unique(prob.obs_img)})
    target_width = max((len(s) for s in used_pts), default=0)
    img_width = 0
    for s in used_img:
        if len(s) > target_width:
            img_width = len(s)
    return max(target_width, img_width, 12) + 2


def settings_rows(s) -> List[list]:
    """``[fieldnames(data.settings) struct2cell(data.settings)]`` in assignment order (main.m:116-171)."""
    rows = [["Output_Filename", s.Output_Filename], ["Meas_std", s.Meas_std]]
    if s.Meas_std_y is not None:
        rows += [["Meas_std_y", s.Meas_std_y], ["no_std_y", 0]]
    else:
        rows += [["no_std_y", 1]]                     # main.m:400 removes the Meas_std_y field (rmfield)
    rows += [["type", s.type], ["Check_Points", s.Check_Points], ["Iteration_Cap", s.Iteration_Cap],
             ["threshold", s.threshold], ["Inner_Constraints", s.Inner_Constraints]]
    for k in ("Estimate_Xc", "Estimate_Yc", "Estimate_Zc", "Estimate_w", "Estimate_p", "Estimate_k", "Estimate_c",
              "Estimate_xp", "Estimate_yp", "Estimate_radial", "Num_Radial_Distortions", "Estimate_decent",
              "Estimate_tie", "Estimate_AllGCP"):
        rows.append([k, getattr(s, k)])
    return rows


def iop_names(s) -> List[str]:
    names = []
    if s.Estimate_xp:
        names.append("xp")
    if s.Estimate_yp:
        names.append("yp")
    if s.Estimate_c:
        names.append("c")
    if s.Estimate_radial:
        names += [f"k{j + 1}" for j in range(s.Num_Radial_Distortions)]
    if s.Estimate_decent:
        names += ["p1", "p2"]
    return names


def _lower_triangle(labels: Sequence[str], header: Sequence[str], M: np.ndarray) -> str:
    """Name row then the lower triangle, ``%-6.2s`` labels and ``%-+6.2f`` entries (main.m:833-842, :925-931)."""
    txt = "".join("%-6.2s" % n for n in header) + "\n"
    for j in range(M.shape[0]):
        txt += "%-6.2s" % labels[j] + "".join("%-+6.2f" % M[j, k] for k in range(j + 1)) + "\n"
    return txt + "\n"


def check_point_differences(prob: Problem, xhat: np.ndarray, CZE: formats.StringTable) -> dict:
    """main.m:604-627: estimated minus given coordinates of every check point that is an estimated
    target, their mean and RMS per axis.  Points not in xhat are reported and left out (the
    reference prints a warning and leaves the row empty)."""
    tie_of = {prob.point_name(int(p)): t for t, p in enumerate(prob.tie_pt) if p >= 0}
    off = prob.u_c
    names, diff = [], []
    for row in CZE:
        t = tie_of.get(row[0])
        if t is None:
            print(f"Warning: Check point not found in xhat -> {row[0]}")           # main.m:614
            continue
        given = np.array([formats.str2double(c) for c in row[1:4]])
        names.append(row[0])
        diff.append(xhat[off + 3 * t: off + 3 * t + 3] - given)                   # main.m:621
    d = np.array(diff).reshape(-1, 3)
    return dict(names=names, diff=d, mean=d.mean(axis=0) if len(d) else np.full(3, np.nan),
                rms=np.sqrt((d ** 2).mean(axis=0)) if len(d) else np.full(3, np.nan))


def write_out(path: str, prob: Problem, out: dict, version: str = "feba_b200", cp: Optional[dict] = None,
              when: Optional[str] = None) -> None:
    """main.m:629-950.  ``out``: result of ``adjust(..., cov=True)`` (``xhat``, ``Cx_diag``,
    ``Correlation_IOP``, ``Correlation_image``, ``RSD``, ``RMSx/RMSy/RMS``, ``sigma02``, ``iterations``,
    ``elapsed``).  ``cp``: ``check_point_differences`` when ``Check_Points`` is set."""
    s = prob.settings
    xhat, std = out["xhat"], np.sqrt(out["Cx_diag"])
    W = column_width(prob)
    dec = 5                                                                  # main.m:698
    fs = f"%-{W}.{dec}s%-{W}.{dec}f%-{W}.{dec}f\n"                           # printEOP   main.m:972-974
    fe = f"%-{W}.{dec}s%-{W}.{dec}e%-{W}.{dec}e\n"                           # printDist  main.m:975-977
    ft = f"%-{W}s%-{W}.0d" + f"%-{W}.{dec}f" * 6 + "\n"                      # printTIE   main.m:978-980
    when = when or datetime.datetime.now().strftime("%d-%b-%Y %H:%M:%S")
    ui, uc = s.u_perimage, s.u_percam
    n = prob.n
    numGCP = int(np.unique(prob.obs_pt).size)                                # main.m:382
    parts: List[str] = []
    w = parts.append
    w(f"Version: {version}\n")                                                # main.m:639
    w("Fish-eye model Bundle Adjustment\nWynand Tredoux -- University of Calgary -- 2020\n\n")
    w(LINE)
    w(f"\n\nExecution date:\t{when}\nTime Taken:\t\t{num2str(out['elapsed'])} seconds\n"
      f"Iterations:\t\t{out['iterations']}\nModel Used:\t\t{s.type}")
    w("\n\nSettings used:\n")
    w(print_cell(settings_rows(s), "\t\t", 4))
    w("\n" + LINE + "\n")
    w("\nObservations/Unknowns Summary\n\n")                                  # main.m:655-684
    nic = 7 * s.Inner_Constraints
    w(print_cell([
        ["Number of Photos", prob.numImg], ["Total EOP unknowns", ui * prob.numImg],
        ["Number of Cameras", prob.numCam],
        ["Total IOP unknowns", (s.Estimate_c + s.Estimate_xp + s.Estimate_yp) * prob.numCam],
        ["Total distortion unknowns",
         (s.Estimate_radial * s.Num_Radial_Distortions + s.Estimate_decent * 2) * prob.numCam],
        ["Number of tie/control points", numGCP], ["Number of tie/control points to be estimated", prob.numtie],
        ["Number of control/tie point unknowns", prob.numtie * 3], ["\\line", ""], ["Total Unknowns", len(xhat)],
        ["\\n", ""], ["Number of image points", n // 2], ["Total number of observations", n],
        ["Number of Inner Constraints", nic], ["\\line", ""], ["Total Number of Observations", n + nic], ["\\n", ""],
        ["Total Degrees of Freedom", n + nic - len(xhat)], ["\\n", ""],
        ["A-Posteriori", num2str(out["sigma02"], 10)], ["RMSx", num2str(out["RMSx"], 10)],
        ["RMSy", num2str(out["RMSy"], 10)], ["RMS", num2str(out["RMS"], 10)], ["\\n", ""]], "", 4))
    w(LINE + "\n\n")

    # ---- EOPs per image (main.m:723-777); angles in degrees
    w("Estimated EOPs\nEOP Name\tValue\tStandard Deviation\n")
    per_image = np.bincount(prob.obs_img, minlength=prob.numImg)             # countImagePoints, main.m:981-988
    k = 0
    for j in range(prob.numImg):
        w("\n")
        w(print_cell([["Image", prob.image_name(j)], ["Camera", prob.camera_name(int(prob.img_cam[j]))],
                      ["Number of image points", int(per_image[j])], ["\\line", ""]], "", 4))
        for q in range(6):
            if s.eop_flags[q]:
                f = 180.0 / math.pi if q >= 3 else 1.0
                w(fs % (EOP_LABELS[q], xhat[k] * f, std[k] * f))
                k += 1

    # ---- IOPs and distortions per camera, IOP correlation sub-matrix (main.m:779-842)
    w("\n" + LINE + "\n\nEstimated IOPs and Distortions for each Camera\nIOP Name\tValue\tStandard Deviation\n\n")
    inames = iop_names(s)
    for c in range(prob.numCam):
        box = prob.cam_box[c]
        w(print_cell([["Camera", prob.camera_name(c)], ["y axis dir", num2str(box[0])], ["x min", num2str(box[1])],
                      ["y min", num2str(box[2])], ["x max", num2str(box[3])], ["y max", num2str(box[4])],
                      ["\\line", ""]], "", 4))
        for nm in inames:
            w((fe if nm[0] in "kp" else fs) % (nm, xhat[k], std[k]))     # printDist for k*, p*
            k += 1
        w("\nIOP Correlation sub-matrix\n-------------------------------\n")
        if uc:
            w(_lower_triangle(inames, [""] + inames, np.asarray(out["Correlation_IOP"][c])))
        else:
            w("\n\n")

    # ---- ground coordinates (main.m:866-888)
    if s.Estimate_tie:
        w("\n" + LINE + "\n\nEstimated Ground Coordinates of targets\nTargetID\tnumImages\tX\tY\tZ\tstdX\tstdY\tstdZ\n\n")
        per_point = np.bincount(prob.obs_pt, minlength=prob.numPts)          # countTargetImages, main.m:989-996
        var = np.zeros((prob.numtie, 3))
        for t in range(prob.numtie):
            p = int(prob.tie_pt[t])
            var[t] = out["Cx_diag"][k:k + 3]
            w(ft % ((prob.point_name(p), int(per_point[p])) + tuple(xhat[k:k + 3]) + tuple(std[k:k + 3])))
            k += 3
        avg = np.sqrt(var.mean(axis=0)) if prob.numtie else np.full(3, np.nan)
        w("\n\t\tMeanStd X\tMeanStd Y\tMeanStd Z\n")
        w(("\t\t" + f"%-{W}.{dec}f" * 3 + "\n") % tuple(avg))

    # ---- corrected image measurements (main.m:587-590, :891-895)
    w("\n" + LINE + "\n\nCorrected Image Measurements\nPointID\tImageID\tCorrected x\tCorrected y\n\n")
    fm = f"%-{W}s%-{W}s%-{W}.{dec}f%-{W}.{dec}f\n"
    RSD = out["RSD"]
    xc, yc = prob.obs_x + RSD[:, 1], prob.obs_y + RSD[:, 2]
    w("".join(fm % (prob.point_name(int(prob.obs_pt[i])), prob.image_name(int(prob.obs_img[i])), xc[i], yc[i])
              for i in range(prob.n_obs)))
    if k != len(xhat):
        print("warning: xhat_count didn't end on it's expected value (unknowns + 1)")   # main.m:897-899

    # ---- mean |correlation| between EOPs and IOPs per camera (main.m:901-934)
    w("\n" + LINE + "\n\nAbsolute (positive) mean correlation coefficients between EOPs and IOPs\n\n")
    enames = [EOP_LABELS[q] for q in range(6) if s.eop_flags[q]]
    labels = enames + [""] + inames                                          # quirk, see module docstring
    order = sorted(range(prob.numImg), key=lambda j: prob.camera_name(int(prob.img_cam[j])))   # sortrows(...,2)
    by_cam: dict = {}
    for j in order:
        by_cam.setdefault(prob.camera_name(int(prob.img_cam[j])), []).append(j)
    for cam, imgs in by_cam.items():
        w(f"Camera {cam}\n")
        mean = sum(np.abs(np.tril(np.asarray(out["Correlation_image"][j]))) for j in imgs) / len(imgs)
        w(_lower_triangle(labels, [""] + labels, mean))

    # ---- check points (main.m:937-946)
    if s.Check_Points and cp is not None:
        w("\n" + LINE + "\n\nCheck point differences\n")
        w((f"%-{W}s" * 4 + "\n\n") % ("TargetID", "diff X", "diff Y", "diff Z"))
        fc = f"%-{W}s" + f"%-{W}.{dec}f" * 3 + "\n"
        for nm, d in zip(cp["names"], cp["diff"]):
            w(fc % ((nm,) + tuple(d)))
        w("\n" + fc % (("Mean",) + tuple(cp["mean"])))
        w(fc % (("RMS",) + tuple(cp["rms"])))
    with open(path, "w") as fh:
        fh.write("".join(parts))
