/*
 * ORACLE (test infrastructure, not product code) -- C / OpenMP restatement of the reference's
 * per-iteration arithmetic, in block form, for sizes the NumPy oracle cannot hold.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs may load
 * this library (oracle/cport.py); the product path never does.
 *
 * What it restates (reference file:line):
 *   functions/BuildAwG.m:46-512   parameter gather, (U,V,W) = M d, distortion at the observed x,y,
 *                                 projection for typeint 0..4, Jacobian blocks, misclosure w
 *   main.m:396-405                weights 1/sigma_x^2, 1/sigma_y^2 (two scalars instead of a dense P)
 *   main.m:424-425                u = A'Pw, N = A'PA -- formed directly in block form and reduced to
 *                                 the camera system S = N_cc - W V^-1 W', g = u_c - W V^-1 u_p
 *                                 (the explicit inverse of main.m:432/442 eliminates the points
 *                                 implicitly; algebraically identical)
 *   main.m:443                    d_p = -V_p^-1 (u_p + W_p' d_c)
 *   main.m:569                    v = A delta + w with the scaled distortion columns of A
 * The Jacobian is the chain rule through (U,V,W), written independently of
 * oracle/model.py and of the CUDA kernels; tests/test_oracle.py checks it against the frozen outputs
 * of the reference's own generated expressions.  Parity status: see oracle/model.py (pinned against
 * the reference's source executed by oracle/mlab.py and at expression level; no MATLAB process).
 *
 * Build: oracle/Makefile -> oracle/_build/libfeba_oracle.so   (gcc -O2 -fopenmp)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXNK 8
#define MAXNC (MAXNK + 5)
#define MAXCAMPT 8 /* distinct cameras observing one point */

typedef struct {
    double Je[2][6];
    double Jc[2][MAXNC];
    double Jt[2][3];
    double w[2];
} obs_jac;

/* M = R3(k) R2(p) R1(w) as written in BuildAwG.m:163-165 (rows give U, V, W). */
static void rotation(const double* e, double M[3][3]) {
    const double cw = cos(e[3]), sw = sin(e[3]), cp = cos(e[4]), sp = sin(e[4]), ck = cos(e[5]), sk = sin(e[5]);
    M[0][0] = ck * cp;  M[0][1] = cw * sk + ck * sp * sw;  M[0][2] = sk * sw - ck * cw * sp;
    M[1][0] = -cp * sk; M[1][1] = ck * cw - sk * sp * sw;  M[1][2] = ck * sw + cw * sk * sp;
    M[2][0] = sp;       M[2][1] = -cp * sw;                M[2][2] = cp * cw;
}

/* One observation: BuildAwG.m:163-512.  iop: xp yp c k1..kNK p1 p2; box: y_dir xmin ymin xmax ymax. */
static void obs_eq(int type, int NK, double x, double y, const double* e, const double* iop, const double* box,
                   const double* X, obs_jac* o) {
    double M[3][3];
    rotation(e, M);
    const double d[3] = {X[0] - e[0], X[1] - e[1], X[2] - e[2]};
    const double U = M[0][0] * d[0] + M[0][1] * d[1] + M[0][2] * d[2];
    const double V = M[1][0] * d[0] + M[1][1] * d[1] + M[1][2] * d[2];
    const double W = M[2][0] * d[0] + M[2][1] * d[1] + M[2][2] * d[2];
    const double R = sqrt(U * U + V * V);
    const double xp = iop[0], yp = iop[1], c = iop[2], yd = box[0];
    const double* K = iop + 3;
    const double P1 = iop[3 + NK], P2 = iop[4 + NK];
    const double xb = x - xp, yb = y - yp;
    const double r2 = xb * xb + yb * yb;
    double rpow[MAXNK], dr = 0.0, s2jK = 0.0, rp = 1.0;
    for (int j = 0; j < NK; ++j) {
        s2jK += 2.0 * (j + 1) * K[j] * rp;            /* 2 j K_j r^(2(j-1)) */
        rp *= r2;
        rpow[j] = rp;                                 /* r^(2j) */
        dr += K[j] * rp;                              /* BuildAwG.m:172-176 */
    }
    /* g(theta), g'(theta), theta = atan(R/W)  (BuildAwG.m:184-208) */
    const double t = R / W, th = atan(t);
    double g, dg;
    switch (type) {
        case 0: g = th; dg = 1.0; break;
        case 1: g = t; dg = 1.0 + t * t; break;
        case 2: g = 2.0 * sin(0.5 * th); dg = cos(0.5 * th); break;
        case 3: g = sin(th); dg = cos(th); break;
        default: { const double h = tan(0.5 * th); g = 2.0 * h; dg = 1.0 + h * h; } break;
    }
    const double s = g / R;
    const double fx = -c * U * s + xp + dr * xb + P1 * (yb * yb + 3.0 * xb * xb) + 2.0 * P2 * xb * yb;
    const double fy = -c * yd * V * s + yp + dr * yb + P2 * (xb * xb + 3.0 * yb * yb) + 2.0 * P1 * xb * yb;
    o->w[0] = fx - x;                                  /* BuildAwG.m:507-512 */
    o->w[1] = fy - y;
    /* d s / d(U,V,W) */
    const double D = R * R + W * W;
    const double thU = U * W / (R * D), thV = V * W / (R * D), thW = -R / D;
    const double a = dg / R, b = g / (R * R * R);
    const double dsU = a * thU - b * U, dsV = a * thV - b * V, dsW = a * thW;
    const double Jx[3] = {-c * (s + U * dsU), -c * U * dsV, -c * U * dsW};
    const double Jy[3] = {-c * yd * V * dsU, -c * yd * (s + V * dsV), -c * yd * V * dsW};
    for (int k = 0; k < 3; ++k) {
        o->Jt[0][k] = Jx[0] * M[0][k] + Jx[1] * M[1][k] + Jx[2] * M[2][k];   /* BuildAwG.m:454-503 */
        o->Jt[1][k] = Jy[0] * M[0][k] + Jy[1] * M[1][k] + Jy[2] * M[2][k];
        o->Je[0][k] = -o->Jt[0][k];                                          /* BuildAwG.m:290-336 */
        o->Je[1][k] = -o->Jt[1][k];
    }
    /* d(U,V,W)/d omega, d phi, d kappa */
    const double ck = cos(e[5]), sk = sin(e[5]);
    const double om[3] = {M[0][1] * d[2] - M[0][2] * d[1], M[1][1] * d[2] - M[1][2] * d[1],
                          M[2][1] * d[2] - M[2][2] * d[1]};
    const double ph[3] = {-ck * W, sk * W, ck * U - sk * V};
    const double ka[3] = {V, -U, 0.0};
    for (int r = 0; r < 2; ++r) {
        const double* Jr = r ? Jy : Jx;
        o->Je[r][3] = Jr[0] * om[0] + Jr[1] * om[1] + Jr[2] * om[2];         /* BuildAwG.m:223-270 */
        o->Je[r][4] = Jr[0] * ph[0] + Jr[1] * ph[1] + Jr[2] * ph[2];
        o->Je[r][5] = Jr[0] * ka[0] + Jr[1] * ka[1] + Jr[2] * ka[2];
    }
    /* IOP block, BuildAwG.m:367-451; distortion columns pre-divided by r_max^(2j) (:422-445) */
    const double hx = (box[3] - box[1]) * 0.5, hy = (box[4] - box[2]) * 0.5, rmax2 = hx * hx + hy * hy;
    o->Jc[0][0] = 1.0 - dr - s2jK * xb * xb - 6.0 * P1 * xb - 2.0 * P2 * yb;
    o->Jc[1][0] = -s2jK * xb * yb - 2.0 * P1 * yb - 2.0 * P2 * xb;
    o->Jc[0][1] = -s2jK * xb * yb - 2.0 * P2 * xb - 2.0 * P1 * yb;
    o->Jc[1][1] = 1.0 - dr - s2jK * yb * yb - 6.0 * P2 * yb - 2.0 * P1 * xb;
    o->Jc[0][2] = -U * s;
    o->Jc[1][2] = -yd * V * s;
    double sc = 1.0;
    for (int j = 0; j < NK; ++j) {
        sc *= rmax2;
        o->Jc[0][3 + j] = rpow[j] * xb / sc;
        o->Jc[1][3 + j] = rpow[j] * yb / sc;
    }
    o->Jc[0][3 + NK] = (yb * yb + 3.0 * xb * xb) / rmax2;
    o->Jc[1][3 + NK] = 2.0 * xb * yb / rmax2;
    o->Jc[0][4 + NK] = 2.0 * xb * yb / rmax2;
    o->Jc[1][4 + NK] = (xb * xb + 3.0 * yb * yb) / rmax2;
}

typedef struct {
    long n_obs;
    int n_img, n_cam, n_pts, n_tie;
    const double *ox, *oy;
    const int *oimg, *opt, *img_cam, *pt_tie;
    const double *eop, *iop, *cam_box, *xyz;     /* CURRENT parameter tables */
    const int *ecol, *ccol;                      /* slots of the estimated parameters or -1 */
    int NK, type, ui, uc;
    double px, py;
    const int *pt_start, *pt_obs;                /* observations grouped by point (PHO indices) */
} oracle_problem;

static void sym3_inv(const double* v, double* inv) { /* v, inv: [00 10 11 20 21 22] */
    const double a = v[0], b = v[1], c = v[2], d = v[3], e = v[4], f = v[5];
    const double c00 = c * f - e * e, c10 = d * e - b * f, c20 = b * e - c * d;
    const double det = a * c00 + b * c10 + d * c20;
    inv[0] = c00 / det;
    inv[1] = c10 / det;
    inv[2] = (a * f - d * d) / det;
    inv[3] = c20 / det;
    inv[4] = (b * d - a * e) / det;
    inv[5] = (a * c - b * b) / det;
}
#define SYM3(m, i, j) ((m)[((i) > (j) ? (i) * ((i) + 1) / 2 + (j) : (j) * ((j) + 1) / 2 + (i))])

static inline void atomic_add(double* p, double v) {
#pragma omp atomic
    *p += v;
}

/* Reduced camera system (lower triangle of S, column-major with leading dimension ldS; g) and the
 * per-point V^-1 (packed 6) and u_p (3).  S and g must be zeroed by the caller. */
int feba_oracle_assemble(const oracle_problem* P, double* S, long ldS, double* g, double* Vinv_out, double* up_out) {
    const int NC = P->NK + 5, ui = P->ui, uc = P->uc;
    const long off_cam = (long)ui * P->n_img;
    const double pw[2] = {P->px, P->py};
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 64)
    for (int pt = 0; pt < P->n_pts; ++pt) {
        const int beg = P->pt_start[pt], end = P->pt_start[pt + 1], m = end - beg;
        if (m == 0) continue;
        const int tie = P->pt_tie[pt];
        obs_jac* J = (obs_jac*)malloc(sizeof(obs_jac) * (size_t)m);
        /* cameras this point is seen by (usually one) and their W_c = sum Jc' P Jt */
        int cams[MAXCAMPT], ncam = 0;
        double V[6] = {0}, up[3] = {0}, Wc[MAXCAMPT][MAXNC][3];
        memset(Wc, 0, sizeof(Wc));
        for (int a = 0; a < m; ++a) {
            const int o = P->pt_obs[beg + a], img = P->oimg[o], cam = P->img_cam[img];
            obs_eq(P->type, P->NK, P->ox[o], P->oy[o], P->eop + 6 * img, P->iop + NC * cam, P->cam_box + 5 * cam,
                   P->xyz + 3 * pt, &J[a]);
            if (tie < 0) continue;
            int lc = 0;
            while (lc < ncam && cams[lc] != cam) ++lc;
            if (lc == ncam) {
                if (ncam == MAXCAMPT) { bad = 2; continue; }
                cams[ncam++] = cam;
            }
            int e = 0;
            for (int i = 0; i < 3; ++i)
                for (int k = 0; k <= i; ++k)
                    V[e++] += J[a].Jt[0][i] * pw[0] * J[a].Jt[0][k] + J[a].Jt[1][i] * pw[1] * J[a].Jt[1][k];
            for (int k = 0; k < 3; ++k)
                up[k] += J[a].Jt[0][k] * pw[0] * J[a].w[0] + J[a].Jt[1][k] * pw[1] * J[a].w[1];
            for (int j = 0; j < NC; ++j)
                for (int k = 0; k < 3; ++k)
                    Wc[lc][j][k] += J[a].Jc[0][j] * pw[0] * J[a].Jt[0][k] + J[a].Jc[1][j] * pw[1] * J[a].Jt[1][k];
        }
        double Vi[6] = {0};
        if (tie >= 0) {
            sym3_inv(V, Vi);
            if (!(Vi[0] == Vi[0])) bad = 1;
            memcpy(Vinv_out + 6 * (long)tie, Vi, sizeof(Vi));
            memcpy(up_out + 3 * (long)tie, up, sizeof(up));
        }
        /* ---- direct terms and image-keyed Schur terms, per observation */
        for (int a = 0; a < m; ++a) {
            const int oa = P->pt_obs[beg + a], ia = P->oimg[oa], cama = P->img_cam[ia];
            const long ra = (long)ui * ia, rc = off_cam + (long)uc * cama;
            double We[6][3], Ye[6][3];
            for (int i = 0; i < 6; ++i)
                for (int k = 0; k < 3; ++k)
                    We[i][k] = J[a].Je[0][i] * pw[0] * J[a].Jt[0][k] + J[a].Je[1][i] * pw[1] * J[a].Jt[1][k];
            for (int i = 0; i < 6; ++i)
                for (int k = 0; k < 3; ++k)
                    Ye[i][k] = tie >= 0 ? We[i][0] * SYM3(Vi, 0, k) + We[i][1] * SYM3(Vi, 1, k) + We[i][2] * SYM3(Vi, 2, k) : 0.0;
            for (int i = 0; i < 6; ++i) {
                if (P->ecol[i] < 0) continue;
                const double gv = J[a].Je[0][i] * pw[0] * J[a].w[0] + J[a].Je[1][i] * pw[1] * J[a].w[1] -
                                  (Ye[i][0] * up[0] + Ye[i][1] * up[1] + Ye[i][2] * up[2]);
                atomic_add(&g[ra + P->ecol[i]], gv);
                for (int j = 0; j <= i; ++j) {
                    if (P->ecol[j] < 0) continue;
                    atomic_add(&S[(ra + P->ecol[i]) + ldS * (ra + P->ecol[j])],
                               J[a].Je[0][i] * pw[0] * J[a].Je[0][j] + J[a].Je[1][i] * pw[1] * J[a].Je[1][j]);
                }
            }
            for (int j = 0; j < NC; ++j) {
                if (P->ccol[j] < 0) continue;
                atomic_add(&g[rc + P->ccol[j]], J[a].Jc[0][j] * pw[0] * J[a].w[0] + J[a].Jc[1][j] * pw[1] * J[a].w[1]);
                for (int i = 0; i < 6; ++i)
                    if (P->ecol[i] >= 0)
                        atomic_add(&S[(rc + P->ccol[j]) + ldS * (ra + P->ecol[i])],
                                   J[a].Jc[0][j] * pw[0] * J[a].Je[0][i] + J[a].Jc[1][j] * pw[1] * J[a].Je[1][i]);
                for (int l = 0; l <= j; ++l)
                    if (P->ccol[l] >= 0)
                        atomic_add(&S[(rc + P->ccol[j]) + ldS * (rc + P->ccol[l])],
                                   J[a].Jc[0][j] * pw[0] * J[a].Jc[0][l] + J[a].Jc[1][j] * pw[1] * J[a].Jc[1][l]);
            }
            if (tie < 0) continue;
            for (int b = 0; b < m; ++b) {                      /* image x image */
                const int ob = P->pt_obs[beg + b], ib = P->oimg[ob];
                if (ib > ia) continue;
                const long rb = (long)ui * ib;
                double Wb[6][3];
                for (int i = 0; i < 6; ++i)
                    for (int k = 0; k < 3; ++k)
                        Wb[i][k] = J[b].Je[0][i] * pw[0] * J[b].Jt[0][k] + J[b].Je[1][i] * pw[1] * J[b].Jt[1][k];
                for (int i = 0; i < 6; ++i) {
                    if (P->ecol[i] < 0) continue;
                    for (int j = 0; j < 6; ++j) {
                        if (P->ecol[j] < 0 || (ib == ia && j > i)) continue;
                        atomic_add(&S[(ra + P->ecol[i]) + ldS * (rb + P->ecol[j])],
                                   -(Ye[i][0] * Wb[j][0] + Ye[i][1] * Wb[j][1] + Ye[i][2] * Wb[j][2]));
                    }
                }
            }
            for (int lc = 0; lc < ncam; ++lc) {                /* camera x image */
                const long rcc = off_cam + (long)uc * cams[lc];
                for (int j = 0; j < NC; ++j) {
                    if (P->ccol[j] < 0) continue;
                    for (int i = 0; i < 6; ++i)
                        if (P->ecol[i] >= 0)
                            atomic_add(&S[(rcc + P->ccol[j]) + ldS * (ra + P->ecol[i])],
                                       -(Wc[lc][j][0] * Ye[i][0] + Wc[lc][j][1] * Ye[i][1] + Wc[lc][j][2] * Ye[i][2]));
                }
            }
        }
        /* ---- camera x camera Schur terms and camera right-hand side, per camera pair */
        if (tie >= 0) {
            for (int l1 = 0; l1 < ncam; ++l1) {
                double Yc[MAXNC][3];
                for (int j = 0; j < NC; ++j)
                    for (int k = 0; k < 3; ++k)
                        Yc[j][k] = Wc[l1][j][0] * SYM3(Vi, 0, k) + Wc[l1][j][1] * SYM3(Vi, 1, k) + Wc[l1][j][2] * SYM3(Vi, 2, k);
                const long r1 = off_cam + (long)uc * cams[l1];
                for (int j = 0; j < NC; ++j) {
                    if (P->ccol[j] < 0) continue;
                    atomic_add(&g[r1 + P->ccol[j]], -(Yc[j][0] * up[0] + Yc[j][1] * up[1] + Yc[j][2] * up[2]));
                    for (int l2 = 0; l2 < ncam; ++l2) {
                        const long r2 = off_cam + (long)uc * cams[l2];
                        for (int l = 0; l < NC; ++l) {
                            if (P->ccol[l] < 0) continue;
                            if (r1 + P->ccol[j] < r2 + P->ccol[l]) continue;       /* lower triangle only */
                            atomic_add(&S[(r1 + P->ccol[j]) + ldS * (r2 + P->ccol[l])],
                                       -(Yc[j][0] * Wc[l2][l][0] + Yc[j][1] * Wc[l2][l][1] + Yc[j][2] * Wc[l2][l][2]));
                        }
                    }
                }
            }
        }
        free(J);
    }
    return bad;
}

/* d_p = -V_p^-1 (u_p + W_p' d_c) with the SCALED camera increment d_c (main.m:443). */
int feba_oracle_backsub(const oracle_problem* P, const double* dcam, const double* Vinv, const double* up,
                        double* dpts) {
    const int NC = P->NK + 5, ui = P->ui, uc = P->uc;
    const long off_cam = (long)ui * P->n_img;
    const double pw[2] = {P->px, P->py};
#pragma omp parallel for schedule(dynamic, 256)
    for (int pt = 0; pt < P->n_pts; ++pt) {
        const int tie = P->pt_tie[pt];
        if (tie < 0) continue;
        double t[3] = {up[3 * (long)tie], up[3 * (long)tie + 1], up[3 * (long)tie + 2]};
        for (int q = P->pt_start[pt]; q < P->pt_start[pt + 1]; ++q) {
            const int o = P->pt_obs[q], img = P->oimg[o], cam = P->img_cam[img];
            obs_jac J;
            obs_eq(P->type, P->NK, P->ox[o], P->oy[o], P->eop + 6 * img, P->iop + NC * cam, P->cam_box + 5 * cam,
                   P->xyz + 3 * pt, &J);
            double s[2] = {0, 0};
            for (int i = 0; i < 6; ++i)
                if (P->ecol[i] >= 0) {
                    const double d = dcam[(long)ui * img + P->ecol[i]];
                    s[0] += J.Je[0][i] * d;
                    s[1] += J.Je[1][i] * d;
                }
            for (int j = 0; j < NC; ++j)
                if (P->ccol[j] >= 0) {
                    const double d = dcam[off_cam + (long)uc * cam + P->ccol[j]];
                    s[0] += J.Jc[0][j] * d;
                    s[1] += J.Jc[1][j] * d;
                }
            for (int k = 0; k < 3; ++k) t[k] += J.Jt[0][k] * pw[0] * s[0] + J.Jt[1][k] * pw[1] * s[1];
        }
        const double* Vi = Vinv + 6 * (long)tie;
        for (int k = 0; k < 3; ++k)
            dpts[3 * (long)tie + k] = -(SYM3(Vi, k, 0) * t[0] + SYM3(Vi, k, 1) * t[1] + SYM3(Vi, k, 2) * t[2]);
    }
    return 0;
}

/* v = A delta + w at the linearisation point (tables in P), UN-scaled delta against the SCALED
 * distortion columns (main.m:569 after main.m:458-482), PHO order, interleaved x y. */
int feba_oracle_residuals(const oracle_problem* P, const double* dcam_unscaled, const double* dpts, double* v) {
    const int NC = P->NK + 5, ui = P->ui, uc = P->uc;
    const long off_cam = (long)ui * P->n_img;
#pragma omp parallel for schedule(static)
    for (long o = 0; o < P->n_obs; ++o) {
        const int img = P->oimg[o], pt = P->opt[o], cam = P->img_cam[img], tie = P->pt_tie[pt];
        obs_jac J;
        obs_eq(P->type, P->NK, P->ox[o], P->oy[o], P->eop + 6 * img, P->iop + NC * cam, P->cam_box + 5 * cam,
               P->xyz + 3 * (long)pt, &J);
        double r[2] = {J.w[0], J.w[1]};
        for (int i = 0; i < 6; ++i)
            if (P->ecol[i] >= 0) {
                const double d = dcam_unscaled[(long)ui * img + P->ecol[i]];
                r[0] += J.Je[0][i] * d;
                r[1] += J.Je[1][i] * d;
            }
        for (int j = 0; j < NC; ++j)
            if (P->ccol[j] >= 0) {
                const double d = dcam_unscaled[off_cam + (long)uc * cam + P->ccol[j]];
                r[0] += J.Jc[0][j] * d;
                r[1] += J.Jc[1][j] * d;
            }
        if (tie >= 0)
            for (int k = 0; k < 3; ++k) {
                r[0] += J.Jt[0][k] * dpts[3 * (long)tie + k];
                r[1] += J.Jt[1][k] * dpts[3 * (long)tie + k];
            }
        v[2 * o] = r[0];
        v[2 * o + 1] = r[1];
    }
    return 0;
}

/* Jacobian rows of single observations, for the expression-level checks of the C restatement. */
int feba_oracle_obs(int type, int NK, long n, const double* x, const double* y, const double* eop /* n x 6 */,
                    const double* iop /* n x NC */, const double* box /* n x 5 */, const double* xyz /* n x 3 */,
                    double* Je /* n x 2 x 6 */, double* Jc /* n x 2 x NC */, double* Jt /* n x 2 x 3 */,
                    double* w /* n x 2 */) {
    const int NC = NK + 5;
    for (long i = 0; i < n; ++i) {
        obs_jac J;
        obs_eq(type, NK, x[i], y[i], eop + 6 * i, iop + NC * i, box + 5 * i, xyz + 3 * i, &J);
        for (int r = 0; r < 2; ++r) {
            for (int k = 0; k < 6; ++k) Je[(i * 2 + r) * 6 + k] = J.Je[r][k];
            for (int k = 0; k < NC; ++k) Jc[(i * 2 + r) * NC + k] = J.Jc[r][k];
            for (int k = 0; k < 3; ++k) Jt[(i * 2 + r) * 3 + k] = J.Jt[r][k];
            w[i * 2 + r] = J.w[r];
        }
    }
    return 0;
}

/* OpenMP threads of the next calls (bench.py: all host cores, also when a launcher exported OMP_NUM_THREADS=1) */
void feba_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int feba_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
