"""ORACLE TOOLING (test infrastructure) -- extended-precision solution of one Gauss-Newton step.

The normal matrix of this problem has cond 1e11..2e13 on the bundled data, so
two *correct* double-precision solvers (explicit inverse as in main.m:432/442, LU, Cholesky with a
border) differ from each other by far more than 1e-9 in a single step.  To judge a solver, this
module computes the step to ~1e-17: the normal equations are accumulated in ``numpy.longdouble``
(x87 80-bit) from the double-precision A, w, G of ``oracle.model.BuildAwG``, the (bordered) system
is solved by double LU + iterative refinement with longdouble residuals.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from .dense import weights
from .model import BuildAwG


def refined_solve(K: np.ndarray, b: np.ndarray, iters: int = 12) -> np.ndarray:
    """Solve K x = b with K, b longdouble; result correct to ~eps_longdouble * small factor."""
    Kd = np.asarray(K, dtype=np.float64)
    lu = sla.lu_factor(Kd)
    x = np.zeros(b.shape, dtype=np.longdouble)
    for _ in range(iters):
        r = b - K @ x
        dx = sla.lu_solve(lu, np.asarray(r, dtype=np.float64))
        x = x + dx.astype(np.longdouble)
        if np.max(np.abs(dx)) <= 1e-18 * np.max(np.abs(x)):
            break
    return x


def exact_step(prob, xhat):
    """delta of main.m:424-482 (un-scaled) at ``xhat`` to extended precision (returned as float64)."""
    s = prob.settings
    err, A, w, G, ds = BuildAwG(prob, xhat)
    assert err == 0
    Pd = weights(prob).astype(np.longdouble)
    Al = A.astype(np.longdouble)
    N = Al.T @ (Pd[:, None] * Al)
    u = Al.T @ (Pd * w.astype(np.longdouble))
    nu = N.shape[0]
    if s.Inner_Constraints:
        Gl = G.astype(np.longdouble)
        K = np.zeros((nu + 7, nu + 7), dtype=np.longdouble)
        K[:nu, :nu] = N
        K[:nu, nu:] = Gl
        K[nu:, :nu] = Gl.T
        rhs = np.concatenate([u, np.zeros(7, dtype=np.longdouble)])
        delta = -refined_solve(K, rhs)[:nu]
    else:
        delta = -refined_solve(N, u)
    delta = delta.copy()
    NK = s.Num_Radial_Distortions
    for i in range(ds.shape[0]):                      # main.m:460-482
        if s.Estimate_radial:
            ri = int(ds[i, 0]) - 1
            for j in range(NK):
                delta[ri + j] /= ds[i, j + 2]
        if s.Estimate_decent:
            di = int(ds[i, 1]) - 1
            delta[di] /= ds[i, 2]
            delta[di + 1] /= ds[i, 2]
    return np.asarray(delta, dtype=np.float64)
