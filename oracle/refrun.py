"""ORACLE TOOLING (test infrastructure) -- run the reference's own functions and loop on a Problem.

Builds the MATLAB workspace ``main.m`` has at line 384 (``data.points(i)`` with its ~30 fields,
``data.settings``, ``data.num*``, the numeric cells ``EXT`` / ``INT`` / ``CNT`` and the string column
``TIE``; main.m:196-384) from a :class:`Problem`, then executes with ``oracle/mlab.py``, straight from the
files under ``/root/reference``:

* ``functions/Buildxhat.m``   (whole function)            -> ``buildxhat``
* ``functions/BuildAwG.m``    (whole function)            -> ``buildawg``
* ``main.m:396-405``          weights ``P``               |
* ``main.m:407-444, 458-494`` the Gauss-Newton loop       |-> ``gauss_newton``
* ``main.m:446-456``          ``Correlation`` (optional: a u^2 interpreted loop, minutes at u = 580)
* ``main.m:569``              ``v = A*delta + w``         |
* ``functions/BuildRSD.m``    (whole function)            |
* ``main.m:592-602`` + local ``rms`` (main.m:998-1002)    |
* ``functions/sumabs.m``      (whole function)
* ``main.m:105-384`` + ``functions/findSetting.m``: settings and the problem build -> ``ReferenceProblemBuild``

The only statements left out are the ones that talk to the user (dropped by the interpreter) and
``tic`` / plotting.  Nothing of the reference is stored in this repository; see ``mlab.py`` for what the
interpreter is and is not.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from . import mlab
from .mlab import Cell, Char, Mat, StrCol, StrMat, Struct, StructArray


def _settings(prob) -> Struct:
    s = prob.settings
    st = Struct(Output_Filename=s.Output_Filename, Meas_std=float(s.Meas_std))
    if s.Meas_std_y is None:                                   # main.m:126-130
        st.Meas_std_y, st.no_std_y = -1.0, 1.0
    else:
        st.Meas_std_y, st.no_std_y = float(s.Meas_std_y), 0.0
    st.type = s.type
    st.Check_Points = float(s.Check_Points)
    st.Iteration_Cap = float(s.Iteration_Cap)
    st.threshold = float(s.threshold)
    for k in ("Inner_Constraints", "Estimate_Xc", "Estimate_Yc", "Estimate_Zc", "Estimate_w", "Estimate_p",
              "Estimate_k", "Estimate_c", "Estimate_xp", "Estimate_yp", "Estimate_radial", "Num_Radial_Distortions",
              "Estimate_decent", "Estimate_tie", "Estimate_AllGCP"):
        setattr(st, k, float(getattr(s, k)))
    return st


def workspace(prob) -> dict:
    """``data``, ``EXT``, ``INT``, ``CNT``, ``TIE`` as main.m:196-384 leaves them."""
    s = prob.settings
    NK = int(s.Num_Radial_Distortions)
    pts = []
    for i in range(prob.n_obs):
        j, p = int(prob.obs_img[i]), int(prob.obs_pt[i])
        c = int(prob.img_cam[j])
        e, io, box = prob.eop0[j], prob.iop0[c], prob.cam_box[c]
        t = int(prob.pt_tie[p])
        pts.append(Struct(
            x=float(prob.obs_x[i]), y=float(prob.obs_y[i]), targetID=prob.point_name(p), imageID=prob.image_name(j),
            ext_index=float(j + 1), cameraID=prob.camera_name(c), Xc=float(e[0]), Yc=float(e[1]), Zc=float(e[2]),
            w=float(e[3]), p=float(e[4]), k=float(e[5]), int_index=float(2 * c + 1), cam_num=float(c + 1),
            xp=float(io[0]), yp=float(io[1]), c=float(io[2]), K=Mat(np.array(io[3:3 + NK]).reshape(-1, 1)),
            P=Mat(np.array(io[3 + NK:5 + NK]).reshape(-1, 1)), xmin=float(box[1]), ymin=float(box[2]),
            xmax=float(box[3]), ymax=float(box[4]), y_dir=float(box[0]), cnt_index=float(p + 1),
            X=float(prob.xyz0[p, 0]), Y=float(prob.xyz0[p, 1]), Z=float(prob.xyz0[p, 2]),
            tieIndex=float(t + 1) if t >= 0 else -1.0, isTie=1.0 if t >= 0 else 0.0))
    data = Struct(points=StructArray(pts), settings=_settings(prob), numImg=float(prob.numImg),
                  numCam=float(prob.numCam), n=float(2 * prob.n_obs),
                  numGCP=float(np.unique(prob.obs_pt).size), numtie=float(prob.numtie))
    EXT = Cell.of([[prob.image_name(j), prob.camera_name(int(prob.img_cam[j]))] + [float(v) for v in prob.eop0[j]]
                   for j in range(prob.numImg)])
    rows = []
    for c in range(prob.numCam):
        width = max(6, 5 + NK)
        r1 = [prob.camera_name(c)] + [float(v) for v in prob.cam_box[c]]
        r2 = [float(v) for v in prob.iop0[c]]
        rows.append(r1 + [Mat(np.zeros((0, 0)))] * (width - len(r1)))
        rows.append(r2 + [Mat(np.zeros((0, 0)))] * (width - len(r2)))
    INT = Cell.of(rows)
    CNT = Cell.of([[prob.point_name(p)] + [float(v) for v in prob.xyz0[p]] for p in range(prob.numPts)])
    TIE = StrCol(prob.point_name(int(p)) for p in prob.tie_pt)
    return dict(data=data, EXT=EXT, INT=INT, CNT=CNT, TIE=TIE)


class Reference:
    """The reference's functions, transpiled once."""

    def __init__(self):
        if not mlab.available():
            raise RuntimeError("reference tree not mounted")
        self.prog = mlab.Program()
        for f in ("BuildAwG", "Buildxhat", "BuildRSD", "sumabs"):
            names = self.prog.add_functions(mlab.read_file(f"functions/{f}.m"))
            assert f in names, (f, names)
        assert self.prog.add_functions(mlab.read_lines("main.m", 997, 1002)) == ["rms"]   # local function of main.m
        self.src_weights = mlab.read_lines("main.m", 396, 405)
        self.src_init = mlab.read_lines("main.m", 407, 410)
        self.src_loop_head = mlab.read_lines("main.m", 412, 444)        # while ... solve (no 'end' yet)
        self.src_corr = mlab.read_lines("main.m", 446, 456)
        self.src_loop_tail = mlab.read_lines("main.m", 458, 494)        # un-scaling ... end of while
        self.src_resid = mlab.read_lines("main.m", 569, 571)
        self.src_stats = mlab.read_lines("main.m", 592, 602)
        self.src_xcorr = mlab.read_lines("main.m", 587, 590)
        self.src_report = mlab.read_lines("main.m", 631, 950)           # the .out writer (check points off)
        assert self.prog.add_functions(mlab.read_file("functions/printCell.m")) == ["printCell"]
        names = self.prog.add_functions(mlab.read_lines("main.m", 972, 996))
        assert names == ["printEOP", "printDist", "printTIE", "countImagePoints", "countTargetImages"], names

    # ---- single functions
    def buildxhat(self, prob):
        ws = workspace(prob)
        err, xhat, names = self.prog.env["Buildxhat"](ws["data"], ws["EXT"], ws["INT"], ws["TIE"], ws["CNT"])
        return int(err), xhat.a.ravel().copy(), [str(v) for v in names.a.ravel(order="F")]

    def buildawg(self, prob, xhat, ws: Optional[dict] = None):
        ws = ws or workspace(prob)
        err, A, w, G, ds = self.prog.env["BuildAwG"](ws["data"], Mat(np.asarray(xhat, float).reshape(-1, 1)))
        return dict(error=int(err), A=A.a, w=w.a.ravel(), G=(G.a if isinstance(G, Mat) else None), dist_scaling=ds.a)

    # ---- main.m:396-602
    def gauss_newton(self, prob, xhat0, correlation: bool = False, max_iter: Optional[int] = None) -> dict:
        ws = workspace(prob)
        if max_iter is not None:
            ws["data"].settings.Iteration_Cap = float(max_iter)
        ws["xhat"] = Mat(np.asarray(xhat0, float).reshape(-1, 1))
        ws = self.prog.run(self.src_weights, ws, "weights")
        src = self.src_init + self.src_loop_head + (self.src_corr if correlation else "") + self.src_loop_tail
        ws = self.prog.run(src, ws, "loop")
        ws = self.prog.run(self.src_resid + self.src_stats, ws, "residuals")
        RSD = ws["RSD"]
        rsd = np.array([[float(RSD.a[i, j]) for j in range(4, 9)] for i in range(RSD.a.shape[0])])
        out = dict(xhat=ws["xhat"].a.ravel().copy(), iterations=int(ws["count"]),
                   deltasum=[float(v) for v in ws["deltasumarr"].a.ravel()], delta=ws["delta"].a.ravel().copy(),
                   xhat_arr=ws["xhat_arr"].a.copy(), v=ws["v"].a.ravel().copy(), RSD=rsd, RMSx=float(ws["RMSx"]),
                   RMSy=float(ws["RMSy"]), RMS=float(ws["RMS"]), sigma02=float(ws["sigma02"]),
                   Cx_diag=np.diag(ws["Cx"].a).copy(), w=ws["w"].a.ravel().copy())
        if correlation:
            out["Correlation"] = ws["Correlation"].a.copy()
        return out


class ReferenceProblemBuild:
    """main.m:105-384 executed: files{1..4} -> PHO/EXT/CNT/INT (:105-110), settings through the
    reference's ``findSetting.m`` (:112-177), the .tie read (:179-188), string -> numeric cells
    (:196-258), ``Estimate_AllGCP`` (:260-264) and the ``data.points`` build with its linear ``strcmp``
    scans (:277-384).  ``ReadFiles.m`` itself cannot run (``dir``, ``readmatrix``, dialogs): its stand-in
    hands the interpreter the string tables ``formats.read_string_table`` produces, i.e. the statement of
    ``readmatrix``'s options at ReadFiles.m:49 -- that one built-in stays unexecuted."""

    def __init__(self):
        if not mlab.available():
            raise RuntimeError("reference tree not mounted")
        self.prog = mlab.Program()
        assert self.prog.add_functions(mlab.read_file("functions/findSetting.m")) == ["findSetting"]
        self.src = (mlab.read_lines("main.m", 105, 110) + mlab.read_lines("main.m", 112, 177)
                    + mlab.read_lines("main.m", 179, 188) + mlab.read_lines("main.m", 196, 264)
                    + mlab.read_lines("main.m", 277, 384))

    def run(self, folder: str, cfg_folder: Optional[str] = None) -> Optional[dict]:
        import glob
        import os
        from feba_b200 import formats

        def table(ext, where=folder):
            hits = sorted(glob.glob(os.path.join(where, "*" + ext)))
            if len(hits) != 1:
                return None
            return StrMat(formats.read_string_table(hits[0]))

        def ReadFiles(exts):                                     # stand-in, see the class docstring
            out = Cell(exts.a.size, 1)
            for k, e in enumerate(exts.a.ravel(order="F")):
                t = table(str(e))
                if t is None:
                    return 1.0, out
                out.a[k, 0] = t
            return 0.0, out

        cfg_dir = folder if glob.glob(os.path.join(folder, "*.cfg")) else (cfg_folder or folder)
        files = Cell(4, 1)
        for k, e in enumerate((".pho", ".ext", ".cnt", ".int")):
            files.a[k, 0] = table(e)
        self.prog.env["ReadFiles"] = ReadFiles
        ws = dict(files=files, CFG=table(".cfg", cfg_dir), data=Struct(settings=Struct(), points=StructArray([])),
                  pwd=Char(os.path.abspath(folder)), main_error=0.0)
        ws = self.prog.run(self.src, ws, "problem_build")
        if "EXT" not in ws or not hasattr(ws["data"], "numImg"):   # a `return` ended the script early
            return None
        return ws


def report_text(R: "Reference", prob, xhat0, version: str = "v-test\n", date: str = "01-Jan-2020 00:00:00",
                time: str = "1.25", want_workspace: bool = False):
    """main.m:396-602 and then the .out writer main.m:631-950 executed (fopen/fprintf go to a text sink):
    the text of the reference's report for this problem.  ``Check_Points`` must be 0 (main.m:604-627 needs the
    .cze table and ``find``)."""
    assert not prob.settings.Check_Points
    ws = workspace(prob)
    ws["xhat"] = Mat(np.asarray(xhat0, float).reshape(-1, 1))
    ws["xhatnames"] = None
    ws = R.prog.run(R.src_weights, ws, "weights")
    ws = R.prog.run(R.src_init + R.src_loop_head + R.src_corr + R.src_loop_tail, ws, "loop")
    ws = R.prog.run(R.src_resid + R.src_xcorr + R.src_stats, ws, "residuals")
    ws.update(version=Char(version), date=Char(date), time=Char(time), mfiles=Char(""))
    ws = R.prog.run(R.src_report, ws, "report")
    return (ws["fileID"].text(), ws) if want_workspace else ws["fileID"].text()


def sparse_rows(A: np.ndarray):
    """Non-zeros of a design matrix (for compact fixtures)."""
    r, c = np.nonzero(A)
    return r.astype(np.int32), c.astype(np.int32), A[r, c]


__all__ = ["Reference", "ReferenceProblemBuild", "workspace", "report_text", "sparse_rows", "math"]
