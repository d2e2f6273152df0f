"""Host-side mirror of the reference's entry points over the C ABI.

``main(folder, plot)``  -- main.m:10-32 (same arguments, returns ``main_error`` 0/1);
``BatchRun(folders)``    -- BatchRun.m:42-65 with the GUI folder pick replaced by an argument and
                            ``findfiles`` (BatchRun.m:68-150) kept as the recursive discovery;
``adjust(prob)``         -- the part between "files are read" and "report is written":
                            Buildxhat (main.m:388), the Gauss-Newton loop (main.m:412-494) and the
                            residual stage (main.m:569-602), all on the GPU through ``lib.Handle``.

The loop body is one ``feba_iterate`` per pass -- the MEX gateway a MATLAB user would call makes
exactly these calls.  Console lines follow main.m (``Iteration k:``, the echoed
``deltasum =``, ``Elapsed time is ...``, ``sigma02 =``).  Output files: .out (``report.write_out``, main.m:629-950),
.rsd (main.m:957, BuildRSD.m:6,40) and .par (main.m:773-823, :958).  Files are read and IDs resolved
by the native packer (``pack.load_problem_native``, include/feba_pack.h); ``native=False`` keeps the
interpreted mirror of main.m:196-384.
"""
from __future__ import annotations

import os
import time
from typing import List, Optional, Sequence

import numpy as np

from . import formats, report
from .lib import FebaError, Handle
from .pack import load_problem_native
from .problem import Buildxhat, Problem, load_problem


def covariance_outputs(h: Handle, prob: Problem, sigma02: float, images: Optional[Sequence[int]] = None) -> dict:
    """What the report stage reads from ``Cx`` / ``Correlation`` for the EOP/IOP unknowns
    (main.m:602, :728-881): ``Cx_diag`` (variances, distortion entries un-scaled as main.m:468-480),
    the IOP correlation sub-matrix per camera (main.m:828) and the EOP+IOP correlation block of the
    requested images (main.m:846-863).  Correlations are taken before un-scaling (main.m:446-456)."""
    s = prob.settings
    ui, uc = s.u_perimage, s.u_percam
    off_cam = ui * prob.numImg
    out = {"Cx_diag": sigma02 * h.cov_diag()}

    def corr(idx):
        B = h.cov_block(idx)
        d = np.sqrt(np.diag(B))
        return B / np.outer(d, d)

    if uc:
        out["Correlation_IOP"] = [corr(off_cam + uc * c + np.arange(uc)) for c in range(prob.numCam)]
    out["Correlation_image"] = {}
    for j in (images if images is not None else []):
        cam = int(prob.img_cam[j])
        idx = np.concatenate([ui * j + np.arange(ui), off_cam + uc * cam + np.arange(uc)])
        out["Correlation_image"][int(j)] = corr(idx)
    return out


def adjust(prob: Problem, xhat0: Optional[np.ndarray] = None, verbose: bool = True,
           handle: Optional[Handle] = None, cov: bool = False, plan: int = 0) -> dict:
    """main.m:386-602 for an already built ``data``.  Returns xhat, iterations, deltasum trace,
    v, RSD (n_obs x 5: r vx vy vr vt), RMSx, RMSy, RMS, sigma02, elapsed seconds; with ``cov`` also
    the covariance outputs of the EOP/IOP unknowns (``covariance_outputs``).  ``plan``: row order of the
    reduced system (feba_settings.plan: 0 automatic, -1 Buildxhat order / dense, 1 nested dissection)."""
    prob.validate()
    t0 = time.perf_counter()                                              # main.m:386 tic
    if xhat0 is None:
        err, xhat0, _ = Buildxhat(prob)                                   # main.m:388
        if err:
            raise FebaError(1, "Error building xhat")                     # main.m:389-393
    own = handle is None
    h = Handle(prob, plan=plan) if own else handle
    try:
        h.set_xhat(xhat0)
        s = prob.settings
        deltasum, count, trace = 100.0, 0, []                             # main.m:407-408
        while deltasum > s.threshold:                                     # main.m:412
            count += 1
            if verbose:
                print(f"Iteration {count}:")                              # main.m:414
            deltasum = h.iterate()                                        # main.m:416-487
            trace.append(deltasum)
            if verbose:
                print(f"deltasum = {deltasum:.6g}")                       # main.m:487 (echo)
            if count >= s.Iteration_Cap:                                  # main.m:490-493
                if verbose:
                    print("Iteration Cap reached. This can be changed in the .cfg file")
                break
        xhat = h.get_xhat()
        elapsed = time.perf_counter() - t0                                # main.m:496 toc
        if verbose:
            print(f"Elapsed time is {elapsed:.6f} seconds.")              # main.m:497
        res = h.residuals()                                               # main.m:569-601
        if verbose:
            print(f"sigma02 = {res['sigma02']:.6g}")                      # main.m:601 (echo)
        res.update(xhat=xhat, iterations=count, deltasum=trace, elapsed=elapsed,
                   delta=h.get_delta(), timing=h.last_timing(), launches=h.launch_count())
        if cov:
            if h.sparse_info()["datum_images"] > 0:
                # Cx is the top-left block of the inverse of the bordered DENSE normal matrix (main.m:432): the
                # nested-dissection plan of a free network factorises S + E E' instead.  Repeat the LAST iteration
                # (same linearisation point: xhat - delta) in the dense order and take the covariances there.
                with Handle(prob, plan=-1) as hd:
                    hd.set_xhat(xhat - res["delta"])
                    hd.iterate()
                    res.update(covariance_outputs(hd, prob, res["sigma02"], images=range(prob.numImg)))
            else:
                res.update(covariance_outputs(h, prob, res["sigma02"], images=range(prob.numImg)))
        return res
    finally:
        if own:
            h.close()


def write_rsd(path: str, prob: Problem, RSD: np.ndarray) -> None:
    """``writecell(RSD, name.rsd, 'Delimiter','tab')`` (main.m:957): targetID imageID x y r vx vy vr vt."""
    with open(path, "w") as fh:
        for i in range(prob.n_obs):
            row = [prob.point_name(int(prob.obs_pt[i])), prob.image_name(int(prob.obs_img[i])),
                   repr(float(prob.obs_x[i])), repr(float(prob.obs_y[i]))]
            row += [repr(float(val)) for val in RSD[i]]
            fh.write("\t".join(row) + "\n")


def write_par(path: str, prob: Problem, xhat: np.ndarray, Cx_diag: np.ndarray) -> None:
    """``writecell(PAR, name.par, 'Delimiter','tab')`` (main.m:958).  PAR = three header rows
    (main.m:773-775), then per camera ``Camera <id>`` and one row ``name value std`` per estimated
    IOP / distortion term in xhat order (main.m:793-823); std = sqrt(Cx(k,k))."""
    import datetime
    s = prob.settings
    ui, uc = s.u_perimage, s.u_percam
    rows = [["Created with Fish-eye model Bundle Adjustment version:", "feba_b200", ""],
            ["Execution date", datetime.date.today().strftime("%d-%b-%Y"), ""], ["", "", ""]]
    k = ui * prob.numImg
    for c in range(prob.numCam):
        rows.append(["Camera", prob.camera_name(c), ""])
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
        assert len(names) == uc
        for nm in names:
            rows.append([nm, repr(float(xhat[k])), repr(float(np.sqrt(Cx_diag[k])))])
            k += 1
    with open(path, "w") as fh:
        for r in rows:
            fh.write("\t".join(r) + "\n")


def main(folder: Optional[str] = None, plot: bool = True, cfg_folder: Optional[str] = None,
         write_files: bool = True, verbose: bool = True, native: bool = True):
    """``main_error = main(folder, plot)`` (main.m:10).  ``folder`` None = current directory
    (non-batch mode, main.m:24-31).  Returns 0/1; the results of the last run are in
    ``main.last`` (MATLAB keeps them in the workspace / output files)."""
    main.last = None
    data_dir = os.getcwd() if folder is None else folder
    try:
        prob = (load_problem_native if native else load_problem)(data_dir, cfg_folder=cfg_folder)   # main.m:60-384
    except (OSError, IndexError, ValueError) as exc:
        print(f"Error reading files ({exc})")
        return 1
    if prob is None:
        return 1
    CZE = None
    if prob.settings.Check_Points:                                         # main.m:266-275
        term, files = formats.ReadFiles([".cze"], data_dir)
        if term:
            print("Error reading files")
            return 1
        CZE = files[0]
    try:
        out = adjust(prob, verbose=verbose, cov=write_files)
    except (FebaError, ValueError) as exc:
        print("Error building A and w")                                    # main.m:417-421
        print(str(exc))
        return 1
    if write_files:
        name = os.path.splitext(os.path.basename(prob.settings.Output_Filename))[0]
        write_rsd(os.path.join(data_dir, name + ".rsd"), prob, out["RSD"])      # main.m:957
        write_par(os.path.join(data_dir, name + ".par"), prob, out["xhat"], out["Cx_diag"])   # main.m:958
        cp = report.check_point_differences(prob, out["xhat"], CZE) if CZE is not None else None  # main.m:604-627
        if verbose:
            print("Writing output file...")                                # main.m:631
        report.write_out(os.path.join(data_dir, prob.settings.Output_Filename), prob, out, cp=cp)
    if verbose:
        print("Done!")
    out["problem"] = prob
    main.last = out
    return 0


main.last = None


def findfiles(root: str, exts: Sequence[str] = (".pho", ".ext", ".cnt", ".int")) -> List[str]:
    """BatchRun.m:68-150: folders (recursively) that hold exactly one file of each extension."""
    found = []
    for cur, dirs, files in os.walk(root):
        dirs.sort()
        if all(sum(f.endswith(e) for f in files) == 1 for e in exts):
            found.append(cur)
    return found


def BatchRun(selpath: Sequence[str], cfg_folder: Optional[str] = None, verbose: bool = False,
             concurrent: bool = False) -> int:
    """BatchRun.m:42-65: run ``main(folder, false)`` over every data folder; stop at the first
    error.  The reference runs them one at a time on the host.  ``concurrent=True`` loads every
    folder first and advances all adjustments together on the current CUDA device
    (``batch.adjust_batch``; same results, .rsd files written at the end); the multi-GPU sweep
    shards folders over ranks (``bench.py --workload config5``)."""
    allfolders: List[str] = []
    for p in selpath:
        allfolders += findfiles(p)
    if not concurrent:
        for folder in allfolders:
            if main(folder, False, cfg_folder=cfg_folder, verbose=verbose) == 1:
                return 1
        return 0
    from .batch import adjust_batch
    probs = []
    for folder in allfolders:
        try:
            prob = load_problem_native(folder, cfg_folder=cfg_folder)
            if prob is not None:
                prob.validate()
        except (OSError, IndexError, ValueError) as exc:
            print(f"Error reading files ({exc})")
            prob = None
        if prob is None:
            return 1
        probs.append(prob)
    try:
        outs = adjust_batch(probs)
    except (FebaError, ValueError) as exc:
        print("Error building A and w")
        print(str(exc))
        return 1
    for folder, prob, out in zip(allfolders, probs, outs):
        name = os.path.splitext(os.path.basename(prob.settings.Output_Filename))[0]
        write_rsd(os.path.join(folder, name + ".rsd"), prob, out["RSD"])
    BatchRun.last = outs
    return 0


BatchRun.last = None
