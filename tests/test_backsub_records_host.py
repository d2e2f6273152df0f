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


@pytest.mark.parametrize("seed,n_obs,tie", [(5, 4, True), (6, 17, True), (7, 9, False)])
def test_camera_block_from_records_equals_schur_complement(seed, n_obs, tie):
    """k_cam_rec (csrc/feba_assemble.cu): with H_a = P Jc_a - Z_a Fc' and r_a = P w_a - Z_a L^-1 u_p (rec2 of the point
    pass; Z = 0, Fc = 0 for control points),
        sum_a H_a' P^-1 H_a = sum_a Jc_a' P Jc_a - Fc Fc'      and      sum_a H_a' P^-1 r_a = sum_a Jc_a' P w_a - Fc L^-1 u_p,
    i.e. the point's contribution to the Schur-complemented camera block and right-hand side
    (N_cc - W_c V^-1 W_c', u_c - W_c V^-1 u_p of main.m:424-437 after eliminating the point)."""
    rng = np.random.default_rng(seed)
    NC = 10
    P = np.diag(rng.uniform(0.5, 2.0, 2))
    Pinv = np.linalg.inv(P)
    Jc = rng.normal(size=(n_obs, 2, NC))
    Jt = rng.normal(size=(n_obs, 2, 3))
    w = rng.normal(size=(n_obs, 2))
    Ncc = sum(Jc[a].T @ P @ Jc[a] for a in range(n_obs))
    uc = sum(Jc[a].T @ P @ w[a] for a in range(n_obs))
    if tie:
        V = sum(Jt[a].T @ P @ Jt[a] for a in range(n_obs))       # exactly the sum: the identity needs sum Z'P^-1 Z = I
        up = sum(Jt[a].T @ P @ w[a] for a in range(n_obs))
        Wc = sum(Jc[a].T @ P @ Jt[a] for a in range(n_obs))
        ref_block = Ncc - Wc @ np.linalg.solve(V, Wc.T)
        ref_rhs = uc - Wc @ np.linalg.solve(V, up)
        Linv = np.linalg.inv(np.linalg.cholesky(V))
        Fc, ut = Wc @ Linv.T, Linv @ up
    else:
        ref_block, ref_rhs = Ncc, uc
        Linv, Fc, ut = np.zeros((3, 3)), np.zeros((NC, 3)), np.zeros(3)
    blk, rhs = np.zeros((NC, NC)), np.zeros(NC)
    for a in range(n_obs):
        Z = P @ Jt[a] @ Linv.T
        H = P @ Jc[a] - Z @ Fc.T            # 2 x NC
        r = P @ w[a] - Z @ ut
        blk += H.T @ Pinv @ H
        rhs += H.T @ Pinv @ r
    scale = np.max(np.abs(Ncc))
    assert np.max(np.abs(blk - ref_block)) < 1e-12 * scale
    assert np.max(np.abs(rhs - ref_rhs)) < 1e-12 * max(1.0, np.max(np.abs(uc)))
