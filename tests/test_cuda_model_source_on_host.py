"""The CUDA path's per-observation SOURCE on the CPU.

``csrc/feba_model.cuh`` (image / camera table rows, projection for the five models, distortion, the
chain-rule Jacobian blocks, misclosure) is plain scalar C++ marked ``__host__ __device__``.  The kernels
inline it on the GPU; here the same file is compiled with g++ (``tests/host_model/model_host.cpp``) and
every observation of a data set is pushed through it and compared with

* the reference's own ``BuildAwG.m`` EXECUTED by the MATLAB-subset interpreter (frozen non-zeros of A and
  w in ``tests/golden/*_refrun_*.npz``; live for all five projection models when the reference tree is
  mounted), and
* the NumPy oracle.

This checks the arithmetic the kernels are made of without a GPU; it is test infrastructure, not a CPU
path of the product (libfeba.so never calls the host instantiation), and the GPU suite remains the check
of what the device actually computes (FMA contraction, the reductions around these statements).
Tolerance: 1e-12 of the largest entry of each column group, 1e-10 px on the misclosure.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import feba_b200 as fb
from oracle import mlab, model
from tests import golden
from tests.test_reference_source_run import CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_model", "model_host.cpp")
HDR = os.path.join(ROOT, "fish-eye_bundle_adjustment_b200", "csrc", "feba_model.cuh")
LIB = os.path.join(ROOT, "tests", "_build", "libfeba_model_host.so")
_pd = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        # -ffp-contract=off: plain IEEE operations, no fused multiply-add on the host
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall",
                        "-Wno-unknown-pragmas", SRC, "-o", LIB], check=True)
    lib = C.CDLL(LIB)
    lib.feba_host_observation.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double] + [_pd] * 8
    _pi = C.POINTER(C.c_int)
    lib.feba_host_residual.argtypes = ([C.c_int, C.c_int, C.c_int, C.c_double, C.c_double] + [_pd] * 4 + [_pi, _pi]
                                       + [_pd] * 3 + [C.c_double, C.c_double, _pd, _pd])
    return lib


def design_matrix_from_cuda_source(lib, prob, xhat):
    """A and w assembled (dense, like the reference) from the per-observation blocks of the CUDA source."""
    s = prob.settings
    L = model.layout(prob)
    eop, iop, xyz = model.gather_params(prob, xhat)
    NK = s.NK
    A = np.zeros((2 * prob.n_obs, prob.u))
    w = np.zeros(2 * prob.n_obs)
    Je, Jc, Jt, ww = np.zeros(12), np.zeros(2 * (NK + 5)), np.zeros(6), np.zeros(2)
    p = lambda a: a.ctypes.data_as(_pd)
    for i in range(prob.n_obs):
        j, q = int(prob.obs_img[i]), int(prob.obs_pt[i])
        c = int(prob.img_cam[j])
        e, io, bx, X = (np.ascontiguousarray(v, dtype=np.float64) for v in (eop[j], iop[c], prob.cam_box[c], xyz[q]))
        assert lib.feba_host_observation(s.typeint, NK, float(prob.obs_x[i]), float(prob.obs_y[i]), p(e), p(io), p(bx),
                                         p(X), p(Je), p(Jc), p(Jt), p(ww)) == 0
        for r in range(2):
            for k in range(6):
                if L["ecols"][k] >= 0:
                    A[2 * i + r, L["u_img"] * j + L["ecols"][k]] = Je[6 * r + k]
            for k in range(NK + 5):
                if L["ccols"][k] >= 0:
                    A[2 * i + r, L["off_cam"] + L["u_cam"] * c + L["ccols"][k]] = Jc[(NK + 5) * r + k]
            t = int(prob.pt_tie[q])
            if t >= 0:
                A[2 * i + r, L["off_tie"] + 3 * t: L["off_tie"] + 3 * t + 3] = Jt[3 * r: 3 * r + 3]
        w[2 * i: 2 * i + 2] = ww
    return A, w


def compare(prob, A, w, A_ref, w_ref):
    L = model.layout(prob)
    assert np.array_equal(A != 0, A_ref != 0)
    for lo, hi in ((0, L["off_cam"]), (L["off_cam"], L["off_tie"]), (L["off_tie"], A.shape[1])):
        if hi > lo:
            assert np.max(np.abs(A[:, lo:hi] - A_ref[:, lo:hi])) < 1e-12 * np.max(np.abs(A_ref[:, lo:hi]))
    assert np.max(np.abs(w - w_ref)) < 1e-10


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_model_source_against_frozen_executed_reference(host, name):
    z = np.load(golden.path(name + ".npz"))
    prob = CASES[name]()
    A_ref = np.zeros(tuple(z["A_shape"]))
    A_ref[z["A_rows"], z["A_cols"]] = z["A_vals"]
    A, w = design_matrix_from_cuda_source(host, prob, z["xhat0"])
    compare(prob, A, w, A_ref, z["w0"])


@pytest.mark.skipif(not mlab.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("typ", ["equisolid", "orthographic", "stereographic"])
def test_cuda_model_source_against_live_executed_reference(host, typ):
    from oracle import refrun
    prob = golden.load_cam0(type=typ)
    x0 = fb.Buildxhat(prob)[1]
    ref = refrun.Reference().buildawg(prob, x0)
    A, w = design_matrix_from_cuda_source(host, prob, x0)
    compare(prob, A, w, ref["A"], ref["w"])


def test_cuda_model_source_against_numpy_oracle_after_an_update(host):
    """Away from the initial values (after one Gauss-Newton step), all NK from 1 to 8."""
    from oracle import dense
    for NK in (1, 2, 8):
        prob = fb.synth.make_network(8, 120, 6, 900 + NK, mode="free", NK=NK)
        x0 = fb.Buildxhat(prob)[1]
        x1 = dense.gauss_newton(prob, x0, max_iter=1)["xhat"]
        err, A_ref, w_ref, G, ds = dense.BuildAwG(prob, x1)
        A, w = design_matrix_from_cuda_source(host, prob, x1)
        compare(prob, A, w, A_ref, w_ref)


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_residual_source_against_frozen_executed_reference(host, name):
    """What k_residuals computes per observation (csrc/feba_model.cuh: residual_of, rsd_row) against the
    reference's own v = A*delta + w (main.m:569, scaled distortion columns x un-scaled increment) and BuildRSD.m,
    as executed and frozen: linearisation point = xhat before the last update, increment = the last delta,
    xp / yp of the RSD columns from the final xhat."""
    z = np.load(golden.path(name + ".npz"))
    prob = CASES[name]()
    s = prob.settings
    L = model.layout(prob)
    NK = s.NK
    x_prev, x_fin, delta = z["xhat_arr"][-2], z["xhat"], z["delta"]
    assert np.allclose(x_prev + delta, x_fin, rtol=0, atol=1e-9 * np.max(np.abs(x_fin)))
    eop, iop, xyz = model.gather_params(prob, x_prev)
    _, iop_fin, _ = model.gather_params(prob, x_fin)
    ecol = np.ascontiguousarray(L["ecols"], dtype=np.int32)
    ccol = np.ascontiguousarray(L["ccols"][:NK + 5], dtype=np.int32)
    p = lambda a: a.ctypes.data_as(_pd)
    pi = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    v, rsd = np.zeros(2), np.zeros(5)
    V, RSD = np.zeros(2 * prob.n_obs), np.zeros((prob.n_obs, 5))
    for i in range(prob.n_obs):
        j, q = int(prob.obs_img[i]), int(prob.obs_pt[i])
        c = int(prob.img_cam[j])
        t = int(prob.pt_tie[q])
        e, io, bx, X = (np.ascontiguousarray(a, dtype=np.float64) for a in (eop[j], iop[c], prob.cam_box[c], xyz[q]))
        d_img = np.ascontiguousarray(delta[L["u_img"] * j: L["u_img"] * (j + 1)])
        d_cam = np.ascontiguousarray(delta[L["off_cam"] + L["u_cam"] * c: L["off_cam"] + L["u_cam"] * (c + 1)])
        d_pt = np.ascontiguousarray(delta[L["off_tie"] + 3 * t: L["off_tie"] + 3 * t + 3]) if t >= 0 else None
        assert host.feba_host_residual(s.typeint, NK, int(L["u_cam"] > 0), float(prob.obs_x[i]), float(prob.obs_y[i]),
                                       p(e), p(io), p(bx), p(X), pi(ecol), pi(ccol), p(d_img), p(d_cam),
                                       p(d_pt) if d_pt is not None else None, float(iop_fin[c, 0]),
                                       float(iop_fin[c, 1]), p(v), p(rsd)) == 0
        V[2 * i: 2 * i + 2] = v
        RSD[i] = rsd
    vmax = np.max(np.abs(z["v"]))
    assert np.max(np.abs(V - z["v"])) < 1e-10 * vmax
    assert np.max(np.abs(RSD - z["RSD"])) < 1e-10 * max(1.0, vmax)


@pytest.mark.parametrize("name", ["cam0_refrun_pinhole", "cam0_refrun_fisheye", "syn_refrun_2cam"])
def test_cuda_inner_constraint_rows_against_frozen_executed_reference(host, name):
    """k_G_rows' per-image block (inner_constraint_rows) against the G the executed BuildAwG.m:514-527 built."""
    z = np.load(golden.path(name + ".npz"))
    prob = CASES[name]()
    assert prob.settings.Inner_Constraints and prob.settings.u_perimage == 6
    eop, _, _ = model.gather_params(prob, z["xhat0"])
    host.feba_host_inner_constraint_rows.argtypes = [_pd, _pd]
    G = np.zeros((prob.u, 7))
    blk = np.zeros(42)
    for j in range(prob.numImg):
        e = np.ascontiguousarray(eop[j], dtype=np.float64)
        host.feba_host_inner_constraint_rows(e.ctypes.data_as(_pd), blk.ctypes.data_as(_pd))
        G[6 * j: 6 * j + 6] = blk.reshape(6, 7)
    assert np.array_equal(G != 0, z["G0"] != 0)
    assert np.max(np.abs(G - z["G0"])) < 1e-13 * np.max(np.abs(z["G0"]))


def test_pinhole_point_on_the_optical_axis_is_finite(host):
    """Type 'pinhole' is fx = -c*U/W in the reference (BuildAwG.m:190-193): a point on the optical axis
    (U = V = 0, R = 0) has finite fx and finite typeint==1 Jacobians there.  The CUDA source must not go
    through g(theta)/R for that model.  Checked against central differences of its own misclosure."""
    NK = 2
    eop = np.array([100.0, 200.0, 1500.0, 0.03, -0.02, 0.7])
    iop = np.array([1207.9, 1013.7, 1234.7, -4.5e-9, 1.3e-15, 1.2e-7, 3.7e-7])
    box = np.array([-1.0, 0.0, 0.0, 2448.0, 2048.0])
    # the object point straight down the optical axis: (X - Xc) = -d * M(3,:)'  =>  U = V = 0, W = -d
    w_, p_, k_ = eop[3:6]
    M = fb.synth._rotation(np.array([w_]), np.array([p_]), np.array([k_]))[0]
    xyz = eop[0:3] - 900.0 * M[2]
    p = lambda a: a.ctypes.data_as(_pd)

    def run(e, io, X):
        Je, Jc, Jt, ww = np.zeros(12), np.zeros(2 * (NK + 5)), np.zeros(6), np.zeros(2)
        assert host.feba_host_observation(1, NK, 1300.0, 900.0, p(np.ascontiguousarray(e)), p(np.ascontiguousarray(io)),
                                          p(box), p(np.ascontiguousarray(X)), p(Je), p(Jc), p(Jt), p(ww)) == 0
        return Je.reshape(2, 6), Jc.reshape(2, NK + 5), Jt.reshape(2, 3), ww.copy()

    Je, Jc, Jt, w0 = run(eop, iop, xyz)
    for blk in (Je, Jc, Jt, w0):
        assert np.all(np.isfinite(blk))
    # on the axis the projection term vanishes: fx = xp + distortion(x, y)
    for k in range(3):                                   # d/dX, d/dY, d/dZ by central differences
        h = 1e-3
        dp, dm = xyz.copy(), xyz.copy()
        dp[k] += h; dm[k] -= h
        num = (run(eop, iop, dp)[3] - run(eop, iop, dm)[3]) / (2 * h)
        assert np.allclose(Jt[:, k], num, rtol=1e-6, atol=1e-9)
    for k in range(6):
        h = 1e-3 if k < 3 else 1e-7
        ep, em = eop.copy(), eop.copy()
        ep[k] += h; em[k] -= h
        num = (run(ep, iop, xyz)[3] - run(em, iop, xyz)[3]) / (2 * h)
        assert np.allclose(Je[:, k], num, rtol=1e-5, atol=1e-6)
    # U = V = 0 exactly (nadir image, point straight below): everything finite, Jt(:, Z) = 0
    nadir = np.array([100.0, 200.0, 1500.0, 0.0, 0.0, 0.0])
    below = np.array([100.0, 200.0, 600.0])
    Je, Jc, Jt, w0 = run(nadir, iop, below)
    for blk in (Je, Jc, Jt, w0):
        assert np.all(np.isfinite(blk))
    assert Jt[0, 2] == 0.0 and Jt[1, 2] == 0.0
    assert np.isclose(Jt[0, 0], -iop[2] / -900.0) and np.isclose(Jt[1, 1], iop[2] / -900.0)   # -c/W, -c*y_dir/W
