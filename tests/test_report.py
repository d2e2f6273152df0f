"""Report stage: .out writer and check-point differences, fed here by the dense
oracle (literal main.m) so the test runs without a GPU; on the GPU the same writer is fed by the
covariance outputs of the CUDA path (tests/test_gpu_parity.py::test_main_and_batchrun_over_text_files)."""
import re

import pytest

import numpy as np

import feba_b200 as fb
from feba_b200 import report
from oracle import dense
from tests import golden


def oracle_outputs(prob):
    """What ``adjust(..., cov=True)`` returns, taken from the literal restatement of main.m."""
    err, x0, _ = fb.Buildxhat(prob)
    o = dense.gauss_newton(prob, x0)
    s = prob.settings
    ui, uc, off = s.u_perimage, s.u_percam, s.u_perimage * prob.numImg
    C = o["Correlation"]
    o["Cx_diag"] = np.diag(o["Cx"]).copy()
    o["Correlation_IOP"] = [C[off + uc * c: off + uc * (c + 1), off + uc * c: off + uc * (c + 1)]
                            for c in range(prob.numCam)]
    o["Correlation_image"] = {}
    for j in range(prob.numImg):
        cam = int(prob.img_cam[j])
        idx = np.concatenate([ui * j + np.arange(ui), off + uc * cam + np.arange(uc)])
        o["Correlation_image"][j] = C[np.ix_(idx, idx)]
    o["elapsed"] = 1.25
    return o


def test_num2str_and_print_cell():
    assert report.num2str(42) == "42" and report.num2str(0.3) == "0.3" and report.num2str(1e-6) == "1e-06"
    assert report.num2str(0.618691873512, 10) == "0.6186918735"
    txt = report.print_cell([["ab", 1], ["\\line", ""], ["abcd", "x"], ["\\n", ""]], ">", 2)
    # the longest first-column entry counts '\\line' (5 characters) too, as printCell.m:3-8 does
    assert txt == ">ab ..... 1\n" + "-" * 13 + "\n>abcd ... x\n\n"


def test_out_file_of_cam0(tmp_path):
    prob = golden.load_cam0()                                   # shipped config: pinhole, IOP + distortions
    out = oracle_outputs(prob)
    path = tmp_path / "cam0.out"
    report.write_out(str(path), prob, out, version="v-test", when="01-Jan-2020 00:00:00")
    text = path.read_text()
    lines = text.split("\n")
    s = prob.settings
    assert lines[0] == "Version: v-test" and report.LINE in text and len(report.LINE) == 109
    assert f"Iterations:\t\t{out['iterations']}" in text and "Model Used:\t\tpinhole" in text
    # summary block (main.m:655-684)
    def field(name):
        m = re.search(r"^" + re.escape(name) + r" \.+ (\S+)$", text, flags=re.M)
        assert m, name
        return m.group(1)
    assert int(field("Number of Photos")) == 42 and int(field("Number of image points")) == 1029
    assert int(field("Total Unknowns")) == prob.u == 580
    assert int(field("Total Degrees of Freedom")) == 2 * 1029 + 7 * s.Inner_Constraints - 580
    assert field("A-Posteriori") == "%.10g" % out["sigma02"]
    assert field("RMS") == "%.10g" % out["RMS"]
    # EOP table: first image, widths and degrees (main.m:723-777, printEOP)
    W = report.column_width(prob)
    assert W == 14
    k0 = lines.index("Estimated EOPs")
    blk = lines[k0:k0 + 14]
    assert blk[3].startswith("Image ") and blk[3].endswith(" " + prob.image_name(0))
    xc = [ln for ln in blk if ln.startswith("Xc")][0]
    assert xc == "%-14.5s%-14.5f%-14.5f" % ("Xc", out["xhat"][0], np.sqrt(out["Cx"][0, 0]))
    om = [ln for ln in blk if ln.startswith("Omega")][0]
    assert om == "%-14.5s%-14.5f%-14.5f" % ("Omega", np.degrees(out["xhat"][3]), np.degrees(np.sqrt(out["Cx"][3, 3])))
    # IOP block (printEOP for xp yp c, printDist for k*, p*), IOP correlation lower triangle
    off = s.u_perimage * prob.numImg
    names = report.iop_names(s)
    for q, nm in enumerate(names):
        fmt = "%-14.5s%-14.5e%-14.5e" if nm[0] in "kp" else "%-14.5s%-14.5f%-14.5f"
        assert fmt % (nm, out["xhat"][off + q], np.sqrt(out["Cx"][off + q, off + q])) in lines, nm
    k1 = lines.index("IOP Correlation sub-matrix")
    assert lines[k1 + 2] == "".join("%-6.2s" % n for n in [""] + names)
    assert lines[k1 + 3] == "%-6.2s%-+6.2f" % (names[0], 1.0)
    assert len(lines[k1 + 2 + len(names)]) == 6 * (len(names) + 1)
    # ground coordinates: one row per TIE entry + mean std (main.m:866-888)
    k2 = lines.index("Estimated Ground Coordinates of targets")
    rows = lines[k2 + 3:k2 + 3 + prob.numtie]
    t = 5
    p = int(prob.tie_pt[t])
    u0 = prob.u_c + 3 * t
    expect = ("%-14s%-14.0d" + "%-14.5f" * 6) % ((prob.point_name(p), int(np.sum(prob.obs_pt == p)))
                                                  + tuple(out["xhat"][u0:u0 + 3])
                                                  + tuple(np.sqrt(np.diag(out["Cx"])[u0:u0 + 3])))
    assert rows[t] == expect
    var = np.diag(out["Cx"])[prob.u_c:].reshape(-1, 3)
    assert ("\t\t" + "%-14.5f" * 3) % tuple(np.sqrt(var.mean(axis=0))) in lines
    # corrected measurements (main.m:587-590, :891-895): one row per observation, PHO order
    k3 = lines.index("Corrected Image Measurements")
    rows = lines[k3 + 3:k3 + 3 + prob.n_obs]
    i = 777
    assert rows[i] == "%-14s%-14s%-14.5f%-14.5f" % (prob.point_name(int(prob.obs_pt[i])),
                                                    prob.image_name(int(prob.obs_img[i])),
                                                    prob.obs_x[i] + out["RSD"][i, 1], prob.obs_y[i] + out["RSD"][i, 2])
    # mean |correlation| table: one camera, (6 + gap + IOP) labels, entries are means over 42 images
    k4 = lines.index("Absolute (positive) mean correlation coefficients between EOPs and IOPs")
    assert lines[k4 + 2] == "Camera " + prob.camera_name(0)
    mean = sum(np.abs(out["Correlation_image"][j]) for j in range(42)) / 42
    assert lines[k4 + 4] == "%-6.2s%-+6.2f" % ("Xc", 1.0)
    assert lines[k4 + 5] == "%-6.2s%-+6.2f%-+6.2f" % ("Yc", mean[1, 0], 1.0)
    assert lines[k4 + 4 + 6].startswith("      ")               # the reference's empty label before the IOP names


def test_check_point_differences(capsys):
    prob = golden.load_cam0()
    err, x0, _ = fb.Buildxhat(prob)
    t = 3
    p = int(prob.tie_pt[t])
    xyz = x0[prob.u_c + 3 * t: prob.u_c + 3 * t + 3]
    CZE = [[prob.point_name(p), repr(float(xyz[0]) - 1.0), repr(float(xyz[1]) + 2.0), repr(float(xyz[2]))], ["nope", "0", "0", "0"]]
    cp = report.check_point_differences(prob, x0, CZE)
    assert "Check point not found in xhat -> nope" in capsys.readouterr().out
    assert cp["names"] == [prob.point_name(p)]
    assert np.allclose(cp["diff"], [[1.0, -2.0, 0.0]], atol=1e-9)
    assert np.allclose(cp["rms"], [1.0, 2.0, 0.0], atol=1e-9)


def test_main_writes_out_rsd_par_from_text_files(tmp_path, monkeypatch):
    """File-level flow of main.m on the CPU: native problem build -> (adjust replaced by the literal
    oracle, the GPU is absent here) -> .out/.rsd/.par + check points from a .cze file."""
    import sys
    prob = fb.synth.make_network(8, 150, 6, 77, mode="mixed", n_control=20)
    prob.settings.Check_Points = 1
    d = tmp_path / "blockA"
    fb.save_problem(prob, str(d), "net")
    with open(d / "net.cfg", "a") as fh:
        fh.write("Check_Points\t1\n")
    t = 2
    p = int(prob.tie_pt[t])
    fb.formats.write_cnt(str(d / "net.cze"), [prob.point_name(p)], prob.xyz0[p:p + 1] + 0.5)
    mod = sys.modules["feba_b200.main"]
    monkeypatch.setattr(mod, "adjust", lambda pr, verbose=True, cov=False: oracle_outputs(pr))
    assert fb.main(str(d), False, verbose=False) == 0
    out = fb.main.last
    text = (d / "blockA.out").read_text()                        # Output_Filename defaults to <folder>.out (main.m:116-120)
    assert "Check point differences" in text and "Estimated Ground Coordinates of targets" in text
    u0 = prob.u_c + 3 * t
    diff = out["xhat"][u0:u0 + 3] - (prob.xyz0[p] + 0.5)
    W = report.column_width(out["problem"])
    assert (f"%-{W}s" + f"%-{W}.5f" * 3) % ((prob.point_name(p),) + tuple(diff)) in text.split("\n")
    assert (d / "blockA.rsd").exists() and (d / "blockA.par").exists()
    # a missing .cze with Check_Points = 1 is a read error (main.m:266-275)
    (d / "net.cze").unlink()
    assert fb.main(str(d), False, verbose=False) == 1


# ------------------------------------------------------------------ formatting primitives vs the reference's source

@pytest.mark.skipif(not __import__("oracle.mlab", fromlist=["x"]).available(), reason="reference tree not mounted")
def test_formatting_primitives_against_executed_reference():
    """functions/printCell.m and the local printEOP / printDist / printTIE of main.m (:972-980), executed by
    the MATLAB-subset interpreter (fprintf with positional conversions into a text sink), produce the same
    text as report.print_cell and the format strings of report.write_out."""
    from oracle import mlab
    from oracle.mlab import Cell, Char, FileSink
    P = mlab.Program()
    assert P.add_functions(mlab.read_file("functions/printCell.m")) == ["printCell"]
    assert P.add_functions(mlab.read_lines("main.m", 972, 980)) == ["printEOP", "printDist", "printTIE"]
    prob = golden.load_cam0()
    tables = [
        (report.settings_rows(prob.settings), "\t\t", 4),
        ([["Number of Photos", "42"], ["Total EOP unknowns", "252"], ["\\line", ""], ["Total Unknowns", "580"],
          ["\\n", ""], ["A-Posteriori", report.num2str(0.618691873512, 10)]], "", 4),
        ([["Image", "101"], ["Camera", "0"], ["Number of image points", "27"], ["\\line", ""]], "", 4),
    ]
    for rows, prefix, padding in tables:
        sink = FileSink()
        cell = Cell.of([[Char(r[0]), Char(r[1]) if isinstance(r[1], str) else float(r[1])] for r in rows])
        P.env["printCell"](sink, cell, Char(prefix.replace("\t", "\\t")), float(padding))
        assert sink.text() == report.print_cell(rows, prefix, padding)
    W, dec = 14, 5
    for name, val, std in (("Xc", 4264.5531234, 0.4123456), ("Omega", -90.04123, 0.0101), ("c", 1234.75756, 0.25)):
        sink = FileSink()
        P.env["printEOP"](sink, Char(name), val, std, Char(str(W)), Char(str(dec)))
        assert sink.text() == f"%-{W}.{dec}s%-{W}.{dec}f%-{W}.{dec}f\n" % (name, val, std)
    for name, val, std in (("k1", -2.2407864e-07, 1.3e-09), ("p2", 5.96155801e-07, 2.2e-08)):
        sink = FileSink()
        P.env["printDist"](sink, Char(name), val, std, Char(str(W)), Char(str(dec)))
        assert sink.text() == f"%-{W}.{dec}s%-{W}.{dec}e%-{W}.{dec}e\n" % (name, val, std)
    sink = FileSink()
    XYZ, sd = mlab.Mat([[2.018], [2574.346], [3519.11]]), mlab.Mat([[0.11], [0.22], [0.33]])
    P.env["printTIE"](sink, Char("AL01"), 17.0, XYZ, sd, Char(str(W)), Char(str(dec)))
    assert sink.text() == (f"%-{W}s%-{W}.0d" + f"%-{W}.{dec}f" * 6 + "\n") % ("AL01", 17, 2.018, 2574.346, 3519.11,
                                                                          0.11, 0.22, 0.33)


# ------------------------------------------------------------------ the whole .out against the reference's own writer

def _our_report(prob, tmp_path):
    out = oracle_outputs(prob)
    out["elapsed"] = "1.25"
    path = tmp_path / "ours.out"
    report.write_out(str(path), prob, out, version="v-test", when="01-Jan-2020 00:00:00")
    return path.read_text()


def test_out_file_equals_frozen_report_of_the_reference_writer(tmp_path):
    """tests/golden/syn_refrun_report.out is the text the reference's OWN report writer (main.m:631-950 with
    printCell.m and its local print functions, executed by the MATLAB-subset interpreter after the reference's
    own loop) produced for a small mixed network: report.write_out, fed by the oracle, must give the same file,
    character for character (625 lines: settings, summary, EOP / IOP / correlation / ground-coordinate / corrected-
    measurement / mean-correlation tables)."""
    from tests.golden.make_refrun import report_case
    ours = _our_report(report_case(), tmp_path)
    ref = open(golden.path("syn_refrun_report.out")).read()
    assert ours.split("\n") == ref.split("\n")


@pytest.mark.skipif(not __import__("oracle.mlab", fromlist=["x"]).available(), reason="reference tree not mounted")
def test_out_file_equals_live_report_two_cameras_inner_constraints(tmp_path):
    """Live: free network (inner constraints, every point a tie point) with two cameras -> per-camera IOP
    tables, IOP correlation sub-matrices and mean EOP x IOP correlations, through the reference's writer."""
    from oracle import refrun
    from tests.golden.make_refrun import synthetic_two_cameras
    import copy
    base = synthetic_two_cameras()
    # smaller than the loop fixture: the reference's Correlation is a u^2 interpreted loop per iteration
    prob = fb.synth.make_network(8, 40, 6, 779, mode="free", NK=2)
    prob.img_cam = base.img_cam.copy()
    prob.iop0 = np.repeat(prob.iop0, 2, axis=0)
    prob.iop0[1, :3] += np.array([0.6, -0.4, 1.1])
    prob.cam_box = np.repeat(prob.cam_box, 2, axis=0)
    prob.camera_ids = ["0", "1"]
    x0 = fb.Buildxhat(prob)[1]
    ref, ws = refrun.report_text(refrun.Reference(), prob, x0, want_workspace=True)
    ours = _our_report(prob, tmp_path)
    assert ours.split("\n") == ref.split("\n")
    # the cells the reference hands to writecell (main.m:957-958): PAR rows and RSD rows vs write_par / write_rsd
    out = oracle_outputs(prob)
    fb.write_par(str(tmp_path / "o.par"), prob, out["xhat"], out["Cx_diag"])
    fb.write_rsd(str(tmp_path / "o.rsd"), prob, out["RSD"])
    par = [ln.split("\t") for ln in (tmp_path / "o.par").read_text().splitlines()]
    PAR = ws["PAR"].a
    assert PAR.shape[0] == len(par) and PAR.shape[1] == 3
    for i in range(3, len(par)):                                     # rows 1-3 are the version / date header
        assert str(PAR[i, 0]) == par[i][0]
        if par[i][0] == "Camera":
            assert str(PAR[i, 1]) == par[i][1]
        else:
            assert abs(float(PAR[i, 1]) - float(par[i][1])) <= 1e-9 * abs(float(par[i][1]))
            assert abs(float(PAR[i, 2]) - float(par[i][2])) <= 1e-6 * abs(float(par[i][2]))
    rsd = [ln.split("\t") for ln in (tmp_path / "o.rsd").read_text().splitlines()]
    RSD = ws["RSD"].a
    assert RSD.shape == (len(rsd), 9)
    for i in (0, 7, len(rsd) - 1):
        assert [str(RSD[i, 0]), str(RSD[i, 1])] == rsd[i][:2]
        assert [float(RSD[i, k]) for k in (2, 3)] == [float(rsd[i][k]) for k in (2, 3)]
        assert np.allclose([float(RSD[i, k]) for k in range(4, 9)], [float(v) for v in rsd[i][4:]], rtol=0, atol=1e-8)
