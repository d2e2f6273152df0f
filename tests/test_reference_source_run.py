"""The oracle against the reference's own source EXECUTED function by function and loop by loop.

MATLAB is not available, but the reference's numerical core is a narrow subset of the language;
``oracle/mlab.py`` transpiles it at run time and ``oracle/refrun.py`` runs ``Buildxhat.m``, ``BuildAwG.m``,
``BuildRSD.m``, ``sumabs.m`` and the loop / residual / statistics statements of ``main.m`` as written
(read from /root/reference; nothing is copied).  Two kinds of tests:

* live  -- need the reference tree (skipped on the GPU box): oracle == executed reference;
* frozen -- ``tests/golden/*_refrun_*.npz`` (made by ``tests/golden/make_refrun.py``): the same
  comparisons against committed outputs of those runs, so they travel; the GPU suite checks the CUDA
  path against the same files (``tests/test_gpu_parity.py::test_cuda_path_against_executed_reference``).

Tolerances: a design-matrix entry differs by rounding only (1e-13 relative to the largest entry of its
column group); everything after the explicit inverse of main.m:432/442 carries cond(N)*eps of it
(cond 1e11..2e13 on cam0), so loop-level quantities use the north-star tolerances (xhat 1e-9 grouped,
v 1e-8*max|v|, sigma02 1e-8) -- the observed differences are 1e-11..1e-14.
"""
import numpy as np
import pytest

import feba_b200 as fb
from oracle import dense, mlab, model, sparse
from tests import golden
from tests.golden.make_refrun import synthetic_eop_only, synthetic_mixed, synthetic_two_cameras

live = pytest.mark.skipif(not mlab.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    from oracle import refrun
    return refrun.Reference()


# ------------------------------------------------------------------ the interpreter itself

def test_interpreter_semantics():
    P = mlab.Program()
    P.add_functions("""
function [s, M, c] = f(x, n)
s = 0;                      % comment with a 'quote
M = zeros(2,3);
c = cell(1,2);
for i = 1:n
    if i == 2 && x > 0
        s = s - x^2;
    elseif i == 3 || x < 0
        s = s + [1 -x; 2 - x 3]*[1; 1];
    else
        s = s + x^-1;
    end
end
M(1:2,2) = [4; 5;];
M(2,:) = M(2,:)./2;
v = M(:,2)';
K = [v(1) v(end)];
s = s + K(2)*x^(2*1) + length(K) - -x^2;
c(2) = {strcat('a_',num2str(3))};
c{1} = sum(v.^2);
end
""")
    s, M, c = P.env["f"](2.0, 3.0)
    # i=1: 1/2; i=2: -4; i=3: + [1 -2; 0 3]*[1;1] = [-1; 3]; then + 2.5*4 + 2 + 4
    assert np.array_equal(s.a, [[-3.5 - 1 + 16.0], [-3.5 + 3 + 16.0]])
    assert np.array_equal(M.a, [[0, 4, 0], [0, 2.5, 0]])
    assert c.a[0, 1] == "a_3" and c.a[0, 0] == 16 + 6.25
    with pytest.raises(mlab.MlabError):
        P.add_functions("function y = g(A)\ny = A^2;\nend\n")
        P.env["g"](mlab.Mat(np.eye(2)))                    # matrix power is outside the subset: refuse


# ------------------------------------------------------------------ live: executed reference vs oracle

@live
def test_buildxhat_executed_reference(ref):
    for prob in (golden.load_cam0(), synthetic_mixed()):
        err, xhat, names = ref.buildxhat(prob)
        e2, x2, n2 = fb.Buildxhat(prob)
        assert err == e2 == 0 and np.array_equal(xhat, x2) and names == n2


def _cmp_awg(prob, got, x0):
    err, A, w, G, ds = dense.BuildAwG(prob, x0)
    assert got["error"] == err == 0 and got["A"].shape == A.shape
    assert np.array_equal(got["A"] != 0, A != 0)                        # same sparsity pattern
    L = model.layout(prob)
    for lo, hi in ((0, L["off_cam"]), (L["off_cam"], L["off_tie"]), (L["off_tie"], A.shape[1])):
        if hi > lo:
            scale = np.max(np.abs(A[:, lo:hi]))
            assert np.max(np.abs(got["A"][:, lo:hi] - A[:, lo:hi])) < 1e-13 * scale
    assert np.max(np.abs(got["w"] - w)) < 1e-11                         # pixels; |w| up to a few px
    if prob.settings.Inner_Constraints:
        assert np.max(np.abs(got["G"] - G)) < 1e-12 * np.max(np.abs(G))
    assert np.allclose(got["dist_scaling"], ds, rtol=1e-15, atol=0)


@live
@pytest.mark.parametrize("typ", ["fisheye", "pinhole", "equisolid", "orthographic", "stereographic"])
def test_buildawg_executed_reference_cam0(ref, typ):
    prob = golden.load_cam0(type=typ)
    x0 = fb.Buildxhat(prob)[1]
    _cmp_awg(prob, ref.buildawg(prob, x0), x0)


@live
def test_buildawg_executed_reference_flag_compaction(ref):
    prob = synthetic_mixed()
    x0 = fb.Buildxhat(prob)[1]
    assert prob.settings.u_perimage == 4 and prob.settings.u_percam == 6
    _cmp_awg(prob, ref.buildawg(prob, x0), x0)


def _cmp_run(prob, run, out, tol_x=1e-9, tol_v=1e-8, tol_d=1e-6):
    assert out["iterations"] == int(run["iterations"])
    # the tail of the trace (values below the 1e-6 threshold) is round-off of the explicit inverse; the
    # stop decision itself is covered by the iteration count above
    assert np.allclose(out["deltasum"], run["deltasum"], rtol=1e-3, atol=5e-8)
    assert np.allclose(out["deltasum"][:2], run["deltasum"][:2], rtol=tol_d)
    vmax = np.max(np.abs(run["v"]))
    assert np.max(np.abs(out["v"] - run["v"])) < tol_v * vmax
    assert np.max(np.abs(out["RSD"] - run["RSD"])) < tol_v * max(1.0, vmax)
    for k in ("RMSx", "RMSy", "RMS", "sigma02"):
        assert abs(out[k] - float(run[k])) < tol_v * float(run[k]), k
    # xhat: the scaled, group-normalised error the GPU tests use (positions, angles, xp yp c, radial and
    # decentering terms in their scaled units, tie points: SURVEY.md 7.2-1)
    from tests.test_gpu_parity import group_rel
    assert group_rel(prob, out["xhat"], np.asarray(run["xhat"])) < tol_x


@live
@pytest.mark.parametrize("case", ["cam0_pinhole", "synthetic_mixed", "synthetic_eop"])
def test_loop_executed_reference(ref, case):
    prob = {"cam0_pinhole": golden.load_cam0, "synthetic_mixed": synthetic_mixed, "synthetic_eop": synthetic_eop_only}[case]()
    x0 = fb.Buildxhat(prob)[1]
    run = ref.gauss_newton(prob, x0)
    _cmp_run(prob, run, dense.gauss_newton(prob, x0))
    _cmp_run(prob, run, sparse.gauss_newton(prob, x0))
    lit = dense.gauss_newton(prob, x0)
    assert np.max(np.abs(run["Cx_diag"] - np.diag(lit["Cx"])) / np.diag(lit["Cx"])) < 1e-6


# ------------------------------------------------------------------ frozen: committed outputs of those runs

CASES = {"cam0_refrun_pinhole": lambda: golden.load_cam0(), "cam0_refrun_fisheye": lambda: golden.load_cam0(type="fisheye"),
         "syn_refrun_mixed": synthetic_mixed, "syn_refrun_2cam": synthetic_two_cameras, "syn_refrun_eop": synthetic_eop_only}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_against_frozen_reference_run(name):
    z = np.load(golden.path(name + ".npz"))
    prob = CASES[name]()
    err, x0, names = fb.Buildxhat(prob)
    assert np.array_equal(x0, z["xhat0"]) and names == [str(v) for v in z["xhatnames"]]
    A = np.zeros(tuple(z["A_shape"]))
    A[z["A_rows"], z["A_cols"]] = z["A_vals"]
    got = dict(error=0, A=A, w=z["w0"], G=z["G0"], dist_scaling=z["dist_scaling"])
    _cmp_awg(prob, got, x0)
    # cam0 + Type 'fisheye' starts far away (first sum|delta| = 509) at cond 2e13: two runs of the SAME
    # explicit inverse on inputs that differ by one ulp already differ by 4e-7 in the first step
    # (509.28580 executed reference vs 509.28601 literal oracle).  That case is
    # compared at what the reference's algorithm can reproduce; the others meet the north-star tolerances
    loose = name == "cam0_refrun_fisheye"
    out = dense.gauss_newton(prob, x0)
    _cmp_run(prob, z, out, tol_x=2e-5 if loose else 1e-9, tol_v=1e-5 if loose else 1e-8, tol_d=1e-5 if loose else 1e-6)
    if not loose:
        _cmp_run(prob, z, sparse.gauss_newton(prob, x0))
    if "corr_iop" in z.files:                                             # main.m:446-456
        s = prob.settings
        ui, uc, off = s.u_perimage, s.u_percam, s.u_perimage * prob.numImg
        C = out["Correlation"]
        assert np.max(np.abs(C[off:off + uc, off:off + uc] - z["corr_iop"])) < 1e-6
        for j in (0, 16):
            idx = np.concatenate([ui * j + np.arange(ui), off + np.arange(uc)])
            assert np.max(np.abs(C[np.ix_(idx, idx)] - z[f"corr_img{j}"])) < 1e-6


# ------------------------------------------------------------------ live: the problem build (main.m:105-384)

def _cmp_build(ws, prob):
    """Workspace left by the executed main.m:105-384 vs the Problem of load_problem / the native packer."""
    d, s = ws["data"], prob.settings
    st = d.settings
    assert str(st.type) == s.type and str(st.Output_Filename) == s.Output_Filename
    assert st.Meas_std == s.Meas_std and bool(st.no_std_y) == (s.Meas_std_y is None)
    for k in ("Iteration_Cap", "threshold", "Inner_Constraints", "Estimate_Xc", "Estimate_Yc", "Estimate_Zc",
              "Estimate_w", "Estimate_p", "Estimate_k", "Estimate_c", "Estimate_xp", "Estimate_yp", "Estimate_radial",
              "Num_Radial_Distortions", "Estimate_decent", "Estimate_tie", "Estimate_AllGCP", "Check_Points"):
        assert float(getattr(st, k)) == float(getattr(s, k)), k
    assert (d.numImg, d.numCam, d.n, d.numtie) == (prob.numImg, prob.numCam, 2 * prob.n_obs, prob.numtie)
    assert d.numGCP == np.unique(prob.obs_pt).size
    NK = s.Num_Radial_Distortions
    same = lambda a, b: (a == b) or (np.isnan(a) and np.isnan(b))
    for i in range(prob.n_obs):
        p = d.points(i + 1)
        j, q = int(prob.obs_img[i]), int(prob.obs_pt[i])
        c = int(prob.img_cam[j])
        assert same(p.x, prob.obs_x[i]) and same(p.y, prob.obs_y[i])
        assert (p.ext_index - 1, p.cnt_index - 1, p.cam_num - 1) == (j, q, c)
        assert str(p.targetID) == prob.point_name(q) and str(p.imageID) == prob.image_name(j)
        assert [p.Xc, p.Yc, p.Zc, p.w, p.p, p.k] == list(prob.eop0[j])           # bit-exact, degrees -> radians
        assert [p.xp, p.yp, p.c] == list(prob.iop0[c, :3])
        assert list(p.K.a.ravel()) == list(prob.iop0[c, 3:3 + NK]) and list(p.P.a.ravel()) == list(prob.iop0[c, 3 + NK:])
        assert [p.y_dir, p.xmin, p.ymin, p.xmax, p.ymax] == list(prob.cam_box[c])
        assert all(same(a, b) for a, b in zip([p.X, p.Y, p.Z], prob.xyz0[q]))
        assert p.tieIndex - 1 == prob.pt_tie[q] if p.isTie else (p.tieIndex == -1 and prob.pt_tie[q] == -1)
    TIE = ws["TIE"]
    ids = [str(v) for v in TIE.a.ravel(order="F")] if hasattr(TIE, "a") else []
    assert ids == [prob.point_name(int(q)) if q >= 0 else ids[t] for t, q in enumerate(prob.tie_pt)]


@live
@pytest.mark.parametrize("case", ["cam0", "free_allgcp", "mixed_tie_file", "ragged"])
def test_problem_build_executed_reference(tmp_path, case):
    """main.m:105-384 (findSetting.m, str2double conversion, Estimate_AllGCP, the strcmp scans that build
    data.points) executed on the same files vs the interpreted mirror and the native packer."""
    from oracle import refrun
    from tests.test_pack import _write, same_problem
    if case == "cam0":
        folder = mlab.REFERENCE_ROOT
    else:
        folder = str(tmp_path / "data9")
        if case == "ragged":
            import os
            os.makedirs(folder)
            _write(folder)
        else:
            prob = fb.synth.make_network(9, 120, 5, 31, mode="free" if case == "free_allgcp" else "mixed",
                                         **({} if case == "free_allgcp" else dict(n_control=15)))
            fb.save_problem(prob, folder, "net")
    ws = refrun.ReferenceProblemBuild().run(folder)
    assert ws is not None
    py, nat = fb.load_problem(folder), fb.load_problem_native(folder)
    same_problem(py, nat)
    _cmp_build(ws, nat)


@live
@pytest.mark.parametrize("what", ["image", "target", "camera"])
def test_problem_build_errors_executed_reference(tmp_path, what):
    """Unknown image / target ID: the executed main.m stops (errordlg + return, main.m:293-297, :352-356)
    and so do the mirror and the native packer.  Unknown camera ID: the reference's scan runs
    ``for j = 1:2:length(INT)`` over a cell that is wider than it is tall, so with one camera MATLAB
    indexes past the last row before it can reach its own error dialog (main.m:310-320; ``length`` is the
    LARGEST dimension) -- the interpreter reproduces that as an IndexError; ours reports the missing ID."""
    from oracle import refrun
    prob = fb.synth.make_network(9, 120, 5, 31, mode="mixed", n_control=15)
    folder = str(tmp_path / "data9")
    fb.save_problem(prob, folder, "net")
    name = {"image": "net.pho", "target": "net.pho", "camera": "net.ext"}[what]
    lines = open(f"{folder}/{name}").read().split("\n")
    cols = lines[5].split("\t")
    cols[{"image": 1, "target": 0, "camera": 1}[what]] = "NOSUCH"
    lines[5] = "\t".join(cols)
    open(f"{folder}/{name}", "w").write("\n".join(lines))
    B = refrun.ReferenceProblemBuild()
    if what == "camera":
        with pytest.raises(IndexError):
            B.run(folder)
    else:
        assert B.run(folder) is None
    assert fb.load_problem(folder) is None and fb.load_problem_native(folder) is None
