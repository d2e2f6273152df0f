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
