"""Native problem build (include/feba_pack.h) against the interpreted mirror of
main.m:196-384 (``problem.load_problem``): same arrays, same IDs, same error behaviour.  Host code
only -- runs without a GPU.  Comparisons are exact (strings -> doubles must round identically)."""
import os
import re

import numpy as np
import pytest

import feba_b200 as fb
from tests import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARRAYS = ("obs_x", "obs_y", "obs_img", "obs_pt", "img_cam", "eop0", "iop0", "cam_box", "xyz0", "pt_tie", "tie_pt")


def same_problem(a, b):
    for k in ARRAYS:
        x, y = getattr(a, k), getattr(b, k)
        assert x.shape == y.shape and x.dtype == y.dtype, k
        assert np.array_equal(x, y, equal_nan=True), k                 # bit-exact
    assert list(a.point_ids) == list(b.point_ids)
    assert list(a.image_ids) == list(b.image_ids)
    assert list(a.camera_ids) == list(b.camera_ids)
    assert a.settings == b.settings


def both(folder, **kw):
    return fb.load_problem(folder, **kw), fb.load_problem_native(folder, **kw)


def test_header_symbols_are_exported():
    text = open(os.path.join(ROOT, "include", "feba_pack.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(feba_pack_[a-z_0-9]+)\s*\(", text)))
    lib = fb.lib.load()
    for n in names:
        assert hasattr(lib, n), n
    assert names == sorted(fb.pack.PACK_EXPORTS)


def test_cam0_round_trip_through_the_text_formats(tmp_path):
    prob = golden.load_cam0()
    fb.save_problem(prob, str(tmp_path), "cam0")
    a, b = both(str(tmp_path))
    assert a is not None and b is not None
    same_problem(a, b)
    assert b.n_obs == 1029 and b.numImg == 42 and b.numtie == 106     # SURVEY.md section 8 size table
    assert fb.Buildxhat(a)[1].tobytes() == fb.Buildxhat(b)[1].tobytes()


@pytest.mark.skipif(not os.path.exists("/root/reference/cam0.pho"), reason="reference tree not mounted")
def test_bundled_dataset_from_the_reference_tree():
    a, b = both("/root/reference")
    assert a is not None and b is not None
    same_problem(a, b)
    same_problem(golden.load_cam0(), b)                                # the committed fixture is this data set


@pytest.mark.parametrize("mode", ["free", "eop", "mixed"])
def test_synthetic_networks(tmp_path, mode):
    kw = dict(n_control=25) if mode == "mixed" else {}
    prob = fb.synth.make_network(9, 400, 7, 11, mode=mode, **kw)
    fb.save_problem(prob, str(tmp_path), "net")
    a, b = both(str(tmp_path))
    assert a is not None and b is not None
    same_problem(a, b)
    assert b.numtie == prob.numtie
    if mode == "free":                                                 # Estimate_AllGCP: TIE = unique(PHO(:,1))
        assert b.numtie == len(set(prob.obs_pt.tolist())) and b.settings.Estimate_tie == 1


def test_threads_give_the_same_arrays(tmp_path):
    # > 1 MiB of .pho so the file is really cut into pieces
    prob = fb.synth.make_network(40, 6000, 8, 5, mode="free")
    fb.save_problem(prob, str(tmp_path), "net")
    assert os.path.getsize(tmp_path / "net.pho") > (1 << 20)
    one = fb.load_problem_native(str(tmp_path), threads=1)
    many = fb.load_problem_native(str(tmp_path), threads=7)
    ref = fb.load_problem(str(tmp_path))
    same_problem(one, many)
    same_problem(ref, many)


def _write(folder, **files):
    base = {
        "n.cfg": "\n".join(f"{k}\t{v}" for k, v in {
            "Iteration_Cap": 10, "Threshold_Value": 1e-6, "Meas_std": 0.3, "Inner_Constraints": 0, "Estimate_Xc": 1,
            "Estimate_Yc": 1, "Estimate_Zc": 1, "Estimate_Omega": 1, "Estimate_Phi": 1, "Estimate_Kappa": 1,
            "Estimate_xp": 0, "Estimate_yp": 0, "Estimate_c": 0, "Estimate_Radial_Distortions": 0,
            "Num_Radial_Distortions": 3, "Estimate_Decentering_Distortions": 0, "Estimate_tie": 1,
            "Estimate_AllGCP": 0, "Type": "'fisheye'"}.items()) + "\n",
        # ragged rows, comments, blank lines, CRLF, leading blanks, repeated delimiters (ReadFiles.m:49)
        "n.pho": "# point image x y\r\n  A\timg1   10.5\t-2e1\r\n\r\nB img1 +3 .5 # trailing\nA img2 1e400 junk\nC\timg2\t7\n",
        "n.ext": "img1 cam 0 0 0 90 180 -45\n\nimg2 cam 1 2 3 1 2 3\nimg1 cam 9 9 9 9 9 9\n",
        "n.cnt": "A 1 2 3\nB 4 5 6\nC 7 8 nan\nA 0 0 0\nD 1 1 1\n",
        "n.int": "cam -1 0 0 100 200\n50.5 60.5 70 1e-9\n",
        "n.tie": "B\nA\nZ\nB\nD\n",
    }
    base.update(files)
    for name, text in base.items():
        if text is not None:
            with open(os.path.join(folder, name), "w", newline="") as fh:
                fh.write(text)


def test_ragged_commented_and_duplicate_rows(tmp_path):
    _write(str(tmp_path))
    a, b = both(str(tmp_path))
    assert a is not None and b is not None
    same_problem(a, b)
    assert b.n_obs == 4
    assert b.obs_x.tolist()[:2] == [10.5, 3.0] and np.isinf(b.obs_x[2]) and np.isnan(b.obs_y[2])
    assert np.isnan(b.obs_y[3])                                        # short row: <missing> -> NaN
    assert b.obs_img.tolist() == [0, 0, 1, 1] and b.obs_pt.tolist() == [0, 1, 0, 2]   # first match wins
    assert b.iop0.tolist() == [[50.5, 60.5, 70.0, 1e-9, 0.0, 0.0, 0.0, 0.0]]         # main.m:243-253
    assert b.eop0[0, 3] == 90 * np.pi / 180
    # TIE rows B A Z B D: Z is not in .cnt (-1), the second B never wins, D is not observed
    assert b.tie_pt.tolist() == [1, 0, -1, 1, 4] and b.pt_tie.tolist() == [1, 0, -1, -1, -1]


@pytest.mark.parametrize("case,needle", [
    (dict(**{"n.pho": "A imgX 1 2\n"}), "Could not find image imgX"),
    (dict(**{"n.pho": "Q img1 1 2\n"}), "Could not find target Q"),
    (dict(**{"n.ext": "img1 other 0 0 0 0 0 0\nimg2 cam 0 0 0 0 0 0\n"}), "Could not find camera"),
    (dict(**{"n.ext": "img0 cam 0 0 0 0 0 0\nimg1 cam 0 0 0 0 0 0\nimg2 cam 1 2 3 1 2 3\n"}), "EXT/INT must list"),
    (dict(**{"n.int": "cam -1 0 0 100 200\n"}), "two rows per camera"),
    (dict(**{"n.tie": None}), "Error"),
])
def test_errors_follow_the_reference(tmp_path, capsys, case, needle):
    _write(str(tmp_path), **case)
    assert fb.load_problem_native(str(tmp_path)) is None
    assert needle in capsys.readouterr().out


def test_pack_problem_fills_the_abi_struct(tmp_path):
    import ctypes as C
    prob = fb.synth.make_network(6, 120, 5, 3, mode="free")
    fb.save_problem(prob, str(tmp_path), "net")
    lib = fb.pack._lib()
    s = fb.lib.settings_struct(prob.settings)
    pk = C.c_void_p()
    enc = lambda n: os.fsencode(str(tmp_path / n))
    assert lib.feba_pack_read(enc("net.pho"), enc("net.ext"), enc("net.cnt"), enc("net.int"), None,
                              s.num_radial, 1, 2, C.byref(pk)) == 0          # free mode: Estimate_AllGCP, no .tie
    pr = fb.lib.FebaProblem()
    assert lib.feba_pack_problem(pk, C.byref(s), C.byref(pr)) == 0
    assert (pr.n_obs, pr.n_img, pr.n_cam, pr.n_pts, pr.n_tie) == (prob.n_obs, prob.numImg, 1, prob.numPts, prob.numtie)
    assert np.array_equal(np.ctypeslib.as_array(pr.obs_pt, shape=(prob.n_obs,)), prob.obs_pt)
    assert pr.settings.num_radial == s.num_radial and pr.settings.sigma_x == s.sigma_x
    s.num_radial += 1
    assert lib.feba_pack_problem(pk, C.byref(s), C.byref(pr)) != 0       # read with another NK
    lib.feba_pack_free(pk)
