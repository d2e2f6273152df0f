"""Host mirror of ReadFiles / findSetting / Buildxhat / the five text formats."""
import os

import numpy as np

import feba_b200 as fb
from feba_b200 import formats, synth


def test_findsetting_semantics(tmp_path):
    p = tmp_path / "a.cfg"
    p.write_text("# comment\nIteration_Cap\t100\nType 'fisheye'   # trailing\nBad abc\nFlag 2\n\n")
    CFG = formats.read_string_table(str(p))
    assert formats.findSetting(CFG, "Iteration_Cap", 0) == (100.0, 0)
    assert formats.findSetting(CFG, "Type", 0) == ("fisheye", 0)
    v, e = formats.findSetting(CFG, "Missing", 3)
    assert e == 4                                   # findSetting.m:33-37
    v, e = formats.findSetting(CFG, "Bad", 0)
    assert e == 1 and np.isnan(v)                   # findSetting.m:40-44
    v, e = formats.findSetting(CFG, "Flag", 0, True)
    assert e == 1                                   # findSetting.m:47-53


def test_readfiles_needs_exactly_one_file(tmp_path):
    term, _ = fb.ReadFiles([".pho"], str(tmp_path))
    assert term == 1
    (tmp_path / "a.pho").write_text("P1 1 1.0 2.0\n")
    (tmp_path / "b.pho").write_text("P1 1 1.0 2.0\n")
    term, _ = fb.ReadFiles([".pho"], str(tmp_path))
    assert term == 1
    os.remove(tmp_path / "b.pho")
    term, files = fb.ReadFiles([".pho"], str(tmp_path))
    assert term == 0 and files[0] == [["P1", "1", "1.0", "2.0"]]


def test_text_format_round_trip(tmp_path):
    for mode, kw in (("free", {}), ("mixed", dict(n_control=10)), ("eop", {})):
        prob = synth.make_network(9, 120, 5, 5, mode=mode, **kw)
        d = tmp_path / mode
        fb.save_problem(prob, str(d))
        back = fb.load_problem(str(d))
        assert back is not None
        assert back.settings.u_perimage == prob.settings.u_perimage
        assert back.u == prob.u and back.numtie == prob.numtie
        for k in ("obs_x", "obs_y", "obs_img", "obs_pt", "img_cam", "iop0", "cam_box", "xyz0", "pt_tie"):
            assert np.array_equal(getattr(back, k), getattr(prob, k)), (mode, k)
        assert np.allclose(back.eop0, prob.eop0, rtol=0, atol=1e-12)   # degrees round trip
        e1, x1, n1 = fb.Buildxhat(prob)
        e2, x2, n2 = fb.Buildxhat(back)
        assert e1 == e2 == 0 and n1 == n2 and np.allclose(x1, x2, rtol=0, atol=1e-9)


def test_buildxhat_layout_and_names():
    from tests import golden
    prob = golden.load_cam0()
    err, xhat, names = fb.Buildxhat(prob)
    assert err == 0 and xhat.size == 580 == prob.u                      # u = 580
    assert names[0].startswith("Xc_101_0") and names[5].startswith("k_101_0")
    assert names[252:262] == ["xp_0", "yp_0", "c_0", "k1_0", "k2_0", "k3_0", "k4_0", "k5_0", "p1_0", "p2_0"]
    assert names[262].startswith("X_")
    prob.settings.Estimate_Zc = 0
    prob.settings.Inner_Constraints = 0
    err, xhat2, names2 = fb.Buildxhat(prob)
    assert xhat2.size == 580 - 42 and not any(n.startswith("Zc_") for n in names2)
