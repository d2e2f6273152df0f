"""Algebra behind k_backsub_rec (csrc/feba_kernels.cu), emulated in numpy on random blocks.

The reference back-substitutes the tie points as  d_p = -V_p^-1 (u_p + W_p' d_c)  (main.m:455-456 through the explicit
inverse of the full normal matrix).  The CUDA path keeps, per point, L^-1 (V_p = L L'), ut = L^-1 u_p and
Fc = Wc L^-T, and per observation Je and Z = P Jt L^-T (the records of the point pass), and forms

    d_p = -L^-T ( ut + Fc' d_cam + sum_a Z_a' Je_a d_e(i_a) )

without evaluating a Jacobian a second time.  This checks that identity, including not-estimated parameters
(ecol / ccol = -1 columns are skipped on both sides)."""
import numpy as np
import pytest


@pytest.mark.parametrize("seed,n_obs,est_e,est_c", [(1, 3, 6, 10), (2, 11, 6, 10), (3, 40, 4, 7), (4, 2, 6, 0)])
def test_record_form_equals_normal_equation_form(seed, n_obs, est_e, est_c):
    rng = np.random.default_rng(seed)
    NC = 10
    P = np.diag(rng.uniform(0.5, 2.0, 2))                     # weights of x and y (main.m:398-405)
    Je = rng.normal(size=(n_obs, 2, 6))
    Jc = rng.normal(size=(n_obs, 2, NC))
    Jt = rng.normal(size=(n_obs, 2, 3))
    w = rng.normal(size=(n_obs, 2))
    ecol = np.array([1] * est_e + [0] * (6 - est_e), bool)   # estimated EOPs / camera parameters
    ccol = np.array([1] * est_c + [0] * (NC - est_c), bool)
    rng.shuffle(ecol)
    rng.shuffle(ccol)
    d_e = rng.normal(size=(n_obs, 6)) * ecol                  # increment of each observation's image
    d_c = rng.normal(size=NC) * ccol

    # reference form
    V = sum(Jt[a].T @ P @ Jt[a] for a in range(n_obs)) + 1e-3 * np.eye(3)
    u = sum(Jt[a].T @ P @ w[a] for a in range(n_obs))
    Wt_dc = sum(Jt[a].T @ P @ (Je[a] @ d_e[a] + Jc[a] @ d_c) for a in range(n_obs))
    ref = -np.linalg.solve(V, u + Wt_dc)

    # record form
    L = np.linalg.cholesky(V)
    Linv = np.linalg.inv(L)
    ut = Linv @ u
    Wc = sum(Jc[a].T @ P @ Jt[a] for a in range(n_obs))      # NC x 3
    Fc = Wc @ Linv.T
    s = ut + Fc.T @ d_c
    for a in range(n_obs):
        Z = P @ Jt[a] @ Linv.T                                # 2 x 3, rec1[12..17]
        s = s + Z.T @ (Je[a] @ d_e[a])
    got = -Linv.T @ s
    assert np.max(np.abs(got - ref)) < 1e-11 * max(1.0, np.max(np.abs(ref)))
