"""RESEARCH NOTE (test infrastructure, not product code; uses the oracle as the checker).

Question for the next round: the reduced camera system S of a photogrammetric block is block-banded
(config 4: 2,680 of 16,653 lower 64x64 tiles are non-zero after symbolic fill, 1.0e10 instead of 5.3e11
flop), but the free-network datum is applied as M = S + Gc Gc' (oracle/sparse.py::solve_reduced, the CUDA
path's k_G_rows / chol), which is dense.  Can the bordered system K = [S Gc; Gc' 0] be solved to the same
accuracy from a factor that keeps S's pattern?

Scheme tried here:  M_s = S + E E',  E = Gc restricted to the rows of a few "datum" images (ordered last,
next to the dense camera rows, so E E' adds no fill).  M_s is positive definite (a similarity transform that
leaves >= 2 images in place is the identity) but conditioned like a minimal-constraints datum.  The bordered
system is solved by block elimination over Y = M_s^-1 [g Gc E] (the augmented block row that already rides
through the factorisation, 15 instead of 8 columns) and a 14x14 border, followed by steps of iterative
refinement on the TRUE K (one sparse product with S and one pair of substitutions per step).

    python tests/research/sparse_reduced_prototype.py [n_img] [n_pts]
"""
import os
import sys

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import feba_b200 as fb                                   # noqa: E402
from oracle import exact, sparse                         # noqa: E402


def bordered_truth(S, g, Gc):
    n, q = Gc.shape
    K = np.zeros((n + q, n + q), dtype=np.longdouble)
    K[:n, :n] = S
    K[:n, n:] = Gc
    K[n:, :n] = Gc.T
    b = np.concatenate([-g, np.zeros(q)]).astype(np.longdouble)
    return np.asarray(exact.refined_solve(K, b)[:n], dtype=np.float64)


def datum_rows(prob, L, n_sel, where="spread"):
    """EOP rows of n_sel images: spread over the block (they would be ordered last in the factorisation, next
    to the dense camera rows) or simply the LAST n_sel images of the image order (no permutation at all)."""
    ui = prob.settings.u_perimage
    if where == "spread":
        imgs = np.unique(np.linspace(0, prob.numImg - 1, n_sel).round().astype(int))
    elif where == "ends":
        # half of them spread over the FIRST 7 % of the image order, half over the LAST 7 % (one supertile each
        # at 13 supertiles): the last supertile row is dense anyway (camera rows), so E E' adds no fill
        w = max(n_sel // 2, int(0.07 * prob.numImg))
        a = np.linspace(0, w - 1, n_sel // 2).round().astype(int)
        imgs = np.unique(np.concatenate([a, prob.numImg - 1 - a]))
    else:
        imgs = np.arange(prob.numImg - n_sel, prob.numImg)
    return np.concatenate([np.arange(ui * i, ui * i + ui) for i in imgs]), imgs


def solve_sparse_datum(S, g, Gc, rows, refine):
    """rows: one index array (datum images) or a list of them (segments: M_s = S + sum_s E_s E_s')."""
    n, q = Gc.shape
    d = 1.0 / np.sqrt(np.diag(S))                         # Jacobi scaling, as the CUDA path does
    Ss, gs, Gs = S * np.outer(d, d), g * d, Gc * d[:, None]
    segs = rows if isinstance(rows, list) else [rows]
    Es = []
    for r in segs:
        E = np.zeros_like(Gs)
        E[r] = Gs[r]
        Es.append(E)
    E = np.column_stack(Es)
    Ms = Ss + E @ E.T
    cf = sla.cho_factor(Ms, lower=True)
    B = np.column_stack([Gs, E])                          # n x (7 + 7 segments)

    def apply_Kinv(r1, r2):
        # K = [Ms - E E'   Gs; Gs' 0]; unknowns delta, k, t = E' delta:
        #   Ms delta + Gs k - E t = r1 ; Gs' delta = r2 ; E' delta - t = 0
        Y = sla.cho_solve(cf, np.column_stack([r1, B]))
        y0, YB = Y[:, 0], Y[:, 1:]
        sgn = np.concatenate([np.ones(q), -np.ones(E.shape[1])])   # delta = y0 - YB (sgn * [k; t])
        A = B.T @ YB * sgn[None, :]
        A[q:, q:] += np.eye(E.shape[1])                   # E'delta - t = 0  ->  E'y0 = (E'YB sgn + I) [.. t]
        rhs = B.T @ y0
        rhs[:q] -= r2
        kt = np.linalg.solve(A, rhs)
        return y0 - YB @ (sgn * kt), kt[:q]

    delta, k = apply_Kinv(-gs, np.zeros(q))
    hist = [delta * d]
    for _ in range(refine):
        r1 = -gs - (Ss @ delta + Gs @ k)                   # residual on the TRUE bordered system
        r2 = -(Gs.T @ delta)
        dd, dk = apply_Kinv(r1, r2)
        delta, k = delta + dd, k + dk
        hist.append(delta * d)
    return hist, np.linalg.cond(Ms)


def main():
    n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    n_pts = int(sys.argv[2]) if len(sys.argv) > 2 else 100 * n_img
    prob = fb.synth.make_network(n_img, n_pts, 10, 4242, mode="free")
    err, x0, _ = fb.Buildxhat(prob)
    nb = sparse.normal_blocks(prob, x0)
    S, g, _ = sparse.reduce(prob, nb)
    Gc = nb["Gc"]
    truth = bordered_truth(S, g, Gc)
    nrm = np.linalg.norm(truth)
    dense = sparse.solve_reduced(prob, S, g, Gc)
    d = 1.0 / np.sqrt(np.diag(S + Gc @ Gc.T))
    condM = np.linalg.cond((S + Gc @ Gc.T) * np.outer(d, d))
    print(f"n_img {n_img}  n_obs {prob.n_obs}  u_c {S.shape[0]}  cond(scaled M) {condM:.2e}")
    print(f"dense M = S + Gc Gc' (today's path):           rel err {np.linalg.norm(dense - truth) / nrm:.2e}")
    for where, n_sel in (("spread", 2), ("spread", 4), ("spread", 8), ("last", 8), ("last", 32),
                         ("ends", 4), ("ends", 8), ("ends", 16)):
        rows, imgs = datum_rows(prob, nb["L"], n_sel, where)
        hist, condMs = solve_sparse_datum(S, g, Gc, rows, 3)
        errs = "  ".join(f"{np.linalg.norm(h - truth) / nrm:.2e}" for h in hist)
        print(f"M_s = S + E E', {len(imgs):2d} datum images ({where:6s}), cond(M_s) {condMs:.2e}: rel err after 0..3 "
              f"refinements  {errs}")
    ui = prob.settings.u_perimage
    for n_seg in (2, 4, 7, 8):
        cuts = np.linspace(0, prob.numImg, n_seg + 1).round().astype(int)
        segs = [np.arange(ui * a, ui * b) for a, b in zip(cuts[:-1], cuts[1:])]
        hist, condMs = solve_sparse_datum(S, g, Gc, segs, 3)
        errs = "  ".join(f"{np.linalg.norm(h - truth) / nrm:.2e}" for h in hist)
        print(f"M_s = S + sum of {n_seg} contiguous segments' E_s E_s' (block diagonal, no fill), cond(M_s) {condMs:.2e}: "
              f"rel err after 0..3 refinements  {errs}")
    # --- the same with the conditioning the CUDA path applies first (k_G_condition: G -> G C with
    # C C' = (G'G)^-1 (G' diag(S) G) (G'G)^-1; the bordered solution depends only on the column space of G)
    cn = 1.0 / np.linalg.norm(Gc, axis=0)
    Gn = Gc * cn
    A1, A2 = Gn.T @ Gn, Gn.T @ (np.diag(S)[:, None] * Gn)
    Gt = Gc @ (np.diag(cn) @ np.linalg.solve(A1, np.linalg.cholesky(A2)))
    M = S + Gt @ Gt.T
    d = 1.0 / np.sqrt(np.diag(M))
    cf = sla.cho_factor(M * np.outer(d, d), lower=True)
    Gs = Gt * d[:, None]
    Y = sla.cho_solve(cf, np.column_stack([g * d, Gs]))
    k = np.linalg.solve(Gs.T @ Y[:, 1:], -(Gs.T @ Y[:, 0]))
    dl = -(Y[:, 0] + Y[:, 1:] @ k) * d
    print(f"with k_G_condition + Jacobi scaling: dense M (the CUDA path's form) cond {np.linalg.cond(M * np.outer(d, d)):.2e} "
          f"rel err {np.linalg.norm(dl - truth) / nrm:.2e}")
    for where, n_sel in (("ends", 4), ("spread", 4)):
        rows, imgs = datum_rows(prob, nb["L"], n_sel, where)
        hist, condMs = solve_sparse_datum(S, g, Gt, rows, 1)
        errs = "  ".join(f"{np.linalg.norm(h - truth) / nrm:.2e}" for h in hist)
        print(f"with k_G_condition: M_s, {len(imgs)} datum images ({where}), cond(M_s) {condMs:.2e}: rel err after 0..1 "
              f"refinements  {errs}")


if __name__ == "__main__":
    main()
