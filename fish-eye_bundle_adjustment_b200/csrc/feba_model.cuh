// Per-observation model: projection (5 models), distortion, analytic Jacobian, misclosure.
//
// Replaces the per-observation body of the reference's BuildAwG (functions/BuildAwG.m:46-512).
// The reference evaluates ~100 machine-generated closed forms that each recompute U,V,W,R,theta
// and every sin/cos from scratch (~5,300 trig calls per observation); here the same function is
// differentiated by the chain rule through (U,V,W) with one atan, one sqrt and no trig per
// observation (the rotation matrix comes from a per-image table, feba_image_table).
//
//   (U,V,W) = M (X-Xc, Y-Yc, Z-Zc),  R = sqrt(U^2+V^2),  theta = atan(R/W)      BuildAwG.m:163-166
//   fx = -c   (U/R) g(theta) + xp + dr*xb + P1 (yb^2+3xb^2) + 2 P2 xb yb         BuildAwG.m:168-208
//   fy = -c yd(V/R) g(theta) + yp + dr*yb + P2 (xb^2+3yb^2) + 2 P1 xb yb
//   xb = x-xp, yb = y-yp (OBSERVED coordinates), dr = sum_j K_j r^(2j)
//
// The functions here are plain scalar C++ (FEBA_HD = __host__ __device__): the kernels inline them, and
// tests/host_model compiles THIS FILE with g++ to check the same source on the CPU against the reference's
// own BuildAwG.m (tests/test_cuda_model_source_on_host.py) -- a check of the arithmetic, not a CPU path of
// the product (nothing in the library calls the host instantiation).
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define FEBA_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define FEBA_HD inline
#endif
#include <stdint.h>

namespace feba {

constexpr int kMaxNK = 8;
constexpr int kImgStride = 16;   // doubles per image-table row (128 B)
constexpr int kCamStride = 32;   // doubles per camera-table row

// image table row: [0..2] Xc Yc Zc, [3..11] M row-major, [12] cos(kappa), [13] sin(kappa)
// camera table row: [0] xp [1] yp [2] c [3] y_dir [4] P1 [5] P2 [6..6+NK) K_j
//                   [16..16+NK) 1/r_max^(2j)   (BuildAwG.m:422-426; columns are pre-divided)

// Image-table row from the EOPs of one image.  M = R3(kappa) R2(phi) R1(omega) as written in
// BuildAwG.m:163-165.
FEBA_HD void image_table_row(const double* __restrict__ e, double* __restrict__ t) {
    double sw, cw, sp, cp, sk, ck;
    sincos(e[3], &sw, &cw);
    sincos(e[4], &sp, &cp);
    sincos(e[5], &sk, &ck);
    t[0] = e[0]; t[1] = e[1]; t[2] = e[2];
    t[3] = ck * cp;  t[4] = cw * sk + ck * sp * sw;  t[5] = sk * sw - ck * cw * sp;
    t[6] = -cp * sk; t[7] = ck * cw - sk * sp * sw;  t[8] = ck * sw + cw * sk * sp;
    t[9] = sp;       t[10] = -cp * sw;               t[11] = cp * cw;
    t[12] = ck; t[13] = sk; t[14] = 0.0; t[15] = 0.0;
}

// Camera-table row from xp yp c k1..kNK p1 p2 (p) and y_dir xmin ymin xmax ymax (b).
FEBA_HD void camera_table_row(int NK, const double* __restrict__ p, const double* __restrict__ b,
                              double* __restrict__ t) {
    for (int k = 0; k < kCamStride; ++k) t[k] = 0.0;
    t[0] = p[0]; t[1] = p[1]; t[2] = p[2]; t[3] = b[0];
    t[4] = p[3 + NK]; t[5] = p[4 + NK];
    for (int j = 0; j < NK; ++j) t[6 + j] = p[3 + j];
    const double hx = (b[3] - b[1]) * 0.5, hy = (b[4] - b[2]) * 0.5;
    const double rmax2 = hx * hx + hy * hy;          // r_max^2  (BuildAwG.m:422)
    double s = 1.0;
    for (int j = 0; j < NK; ++j) {
        s *= rmax2;                                  // r_max^(2j) (BuildAwG.m:424-426)
        t[16 + j] = 1.0 / s;
        t[24 + j] = s;
    }
}

// Inner-constraint rows of one image from its CURRENT EOPs (Gblock of BuildAwG.m:514-527): rows = the six
// EOPs, columns = translation X Y Z, rotation omega phi kappa, scale.
FEBA_HD void inner_constraint_rows(const double* __restrict__ e, double G[6][7]) {
    const double Xc = e[0], Yc = e[1], Zc = e[2];
    double sw, cw;
    sincos(e[3], &sw, &cw);
    const double tp = tan(e[4]), secp = 1.0 / cos(e[4]);
    const double rows[6][7] = {{1, 0, 0, 0, -Zc, Yc, Xc},
                               {0, 1, 0, Zc, 0, -Xc, Yc},
                               {0, 0, 1, -Yc, Xc, 0, Zc},
                               {0, 0, 0, -1, -sw * tp, cw * tp, 0},
                               {0, 0, 0, 0, -cw, -sw, 0},
                               {0, 0, 0, 0, sw * secp, -cw * secp, 0}};
    for (int q = 0; q < 6; ++q)
        for (int c = 0; c < 7; ++c) G[q][c] = rows[q][c];
}

template <int NK>
struct ObsJac {
    double Je[2][6];        // d(fx,fy)/d(Xc,Yc,Zc,omega,phi,kappa)         BuildAwG.m:217-352
    double Jc[2][NK + 5];   // d/d(xp,yp,c,k1..kNK (scaled),p1,p2 (scaled))  BuildAwG.m:367-445
    double Jt[2][3];        // d/d(X,Y,Z)                                    BuildAwG.m:454-503
    double w[2];            // misclosure fx-x, fy-y                         BuildAwG.m:505-512
};

FEBA_HD void g_and_dg(int type, double R, double W, double& g, double& dg) {
    // g(theta), g'(theta), theta = atan(R/W)  (BuildAwG.m:184-208)
    const double t = R / W;
    if (type == 0) {            // equidistant fish-eye: r = c*theta
        g = atan(t);
        dg = 1.0;
    } else if (type == 1) {     // pinhole: written -c*U/W in the reference
        g = t;
        dg = 1.0 + t * t;
    } else if (type == 2) {     // equisolid: 2 sin(theta/2)
        const double h = 0.5 * atan(t);
        g = 2.0 * sin(h);
        dg = cos(h);
    } else if (type == 3) {     // orthographic: sin(theta)
        const double th = atan(t);
        g = sin(th);
        dg = cos(th);
    } else {                    // stereographic: 2 tan(theta/2)
        const double th2 = tan(0.5 * atan(t));
        g = 2.0 * th2;
        dg = 1.0 + th2 * th2;
    }
}

// WANT_J = false computes only the misclosure (fx-x, fy-y).
template <int NK, bool WANT_CAM>
FEBA_HD void observation(int type, double x, double y,
                         const double* __restrict__ it,   // image row
                         const double* __restrict__ ct,   // camera row
                         double X, double Y, double Z, ObsJac<NK>& o) {
    const double dX = X - it[0], dY = Y - it[1], dZ = Z - it[2];
    const double m00 = it[3], m01 = it[4], m02 = it[5];
    const double m10 = it[6], m11 = it[7], m12 = it[8];
    const double m20 = it[9], m21 = it[10], m22 = it[11];
    const double ck = it[12], sk = it[13];
    const double U = m00 * dX + m01 * dY + m02 * dZ;
    const double V = m10 * dX + m11 * dY + m12 * dZ;
    const double W = m20 * dX + m21 * dY + m22 * dZ;
    const double R2 = U * U + V * V;
    const double R = sqrt(R2);
    // s = g(theta)/R and its derivatives with respect to (U,V,W).  The pinhole model is s = 1/W exactly
    // (the reference writes it -c*U/W, BuildAwG.m:190-193): no 1/R, so a point on the optical axis
    // (U = V = 0) stays finite as it does in the reference; the other four models divide by R there too.
    double s, dsU, dsV, dsW;
    if (type == 1) {
        const double invW = 1.0 / W;
        s = invW;
        dsU = 0.0;
        dsV = 0.0;
        dsW = -invW * invW;
    } else {
        double g, dg;
        g_and_dg(type, R, W, g, dg);
        const double invR = 1.0 / R;
        const double invD = 1.0 / (R2 + W * W);
        s = g * invR;
        const double a = dg * invR;
        const double b = s * invR * invR;
        const double thU = U * W * invR * invD, thV = V * W * invR * invD, thW = -R * invD;
        dsU = a * thU - b * U;
        dsV = a * thV - b * V;
        dsW = a * thW;
    }
    const double xp = ct[0], yp = ct[1], c = ct[2], yd = ct[3], P1 = ct[4], P2 = ct[5];
    const double xb = x - xp, yb = y - yp;
    const double r2 = xb * xb + yb * yb;
    double dr = 0.0, s2jK = 0.0, rp = 1.0;     // rp = r^(2j)
    double rpow[NK];
#pragma unroll
    for (int j = 0; j < NK; ++j) {
        const double Kj = ct[6 + j];
        s2jK += (2.0 * (j + 1)) * Kj * rp;     // 2 j K_j r^(2(j-1))
        rp *= r2;
        rpow[j] = rp;
        dr += Kj * rp;
    }
    const double xx = xb * xb, yy = yb * yb, xy = xb * yb;
    const double fx = -c * U * s + xp + dr * xb + P1 * (yy + 3.0 * xx) + 2.0 * P2 * xy;
    const double fy = -c * yd * V * s + yp + dr * yb + P2 * (xx + 3.0 * yy) + 2.0 * P1 * xy;
    o.w[0] = fx - x;
    o.w[1] = fy - y;

    const double cy = c * yd;
    const double JxU = -c * (s + U * dsU), JxV = -c * U * dsV, JxW = -c * U * dsW;
    const double JyU = -cy * V * dsU, JyV = -cy * (s + V * dsV), JyW = -cy * V * dsW;
    // tie block = J_uvw * M ; EOP position block = -tie block
    o.Jt[0][0] = JxU * m00 + JxV * m10 + JxW * m20;
    o.Jt[0][1] = JxU * m01 + JxV * m11 + JxW * m21;
    o.Jt[0][2] = JxU * m02 + JxV * m12 + JxW * m22;
    o.Jt[1][0] = JyU * m00 + JyV * m10 + JyW * m20;
    o.Jt[1][1] = JyU * m01 + JyV * m11 + JyW * m21;
    o.Jt[1][2] = JyU * m02 + JyV * m12 + JyW * m22;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) o.Je[r][k] = -o.Jt[r][k];
    // d(U,V,W)/d omega = M(:,1) dZ - M(:,2) dY ; d/d phi = (-ck W, sk W, ck U - sk V) ;
    // d/d kappa = (V, -U, 0)
    const double oU = m01 * dZ - m02 * dY, oV = m11 * dZ - m12 * dY, oW = m21 * dZ - m22 * dY;
    const double pU = -ck * W, pV = sk * W, pW = ck * U - sk * V;
    o.Je[0][3] = JxU * oU + JxV * oV + JxW * oW;
    o.Je[1][3] = JyU * oU + JyV * oV + JyW * oW;
    o.Je[0][4] = JxU * pU + JxV * pV + JxW * pW;
    o.Je[1][4] = JyU * pU + JyV * pV + JyW * pW;
    o.Je[0][5] = JxU * V - JxV * U;
    o.Je[1][5] = JyU * V - JyV * U;
    if (WANT_CAM) {
        o.Jc[0][0] = 1.0 - dr - s2jK * xx - 6.0 * P1 * xb - 2.0 * P2 * yb;     // BuildAwG.m:375-383
        o.Jc[1][0] = -s2jK * xy - 2.0 * P1 * yb - 2.0 * P2 * xb;
        o.Jc[0][1] = -s2jK * xy - 2.0 * P2 * xb - 2.0 * P1 * yb;               // BuildAwG.m:388-396
        o.Jc[1][1] = 1.0 - dr - s2jK * yy - 6.0 * P2 * yb - 2.0 * P1 * xb;
        o.Jc[0][2] = -U * s;                                                   // BuildAwG.m:399-418
        o.Jc[1][2] = -yd * V * s;
#pragma unroll
        for (int j = 0; j < NK; ++j) {                                         // BuildAwG.m:428-438
            const double sc = ct[16 + j];
            o.Jc[0][3 + j] = rpow[j] * xb * sc;
            o.Jc[1][3 + j] = rpow[j] * yb * sc;
        }
        const double s1 = ct[16];                                              // BuildAwG.m:439-445
        o.Jc[0][3 + NK] = (yy + 3.0 * xx) * s1;
        o.Jc[1][3 + NK] = 2.0 * xy * s1;
        o.Jc[0][4 + NK] = 2.0 * xy * s1;
        o.Jc[1][4 + NK] = (xx + 3.0 * yy) * s1;
    }
}

// v = A*delta + w for one observation (main.m:569) from its Jacobian blocks at the LAST linearisation point:
// d_img / d_cam = this image's / this camera's slice of the UN-scaled increment (main.m:458-482) -- applied
// to the SCALED distortion columns of Jc, as the reference codes it -- d_pt = increment of the tie point or
// nullptr for a control point.  ecol[6] / ccol[NK+5]: slot of each parameter inside its block or -1.
template <int NK, bool HAS_CAM>
FEBA_HD void residual_of(const ObsJac<NK>& J, const int* __restrict__ ecol, const int* __restrict__ ccol,
                         const double* __restrict__ d_img, const double* __restrict__ d_cam,
                         const double* __restrict__ d_pt, double v[2]) {
    v[0] = J.w[0];
    v[1] = J.w[1];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        if (ecol[i] >= 0) {
            const double d = d_img[ecol[i]];
            v[0] += J.Je[0][i] * d;
            v[1] += J.Je[1][i] * d;
        }
    }
    if (HAS_CAM) {
#pragma unroll
        for (int j = 0; j < NK + 5; ++j) {
            if (ccol[j] >= 0) {
                const double d = d_cam[ccol[j]];
                v[0] += J.Jc[0][j] * d;
                v[1] += J.Jc[1][j] * d;
            }
        }
    }
    if (d_pt) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double d = d_pt[k];
            v[0] += J.Jt[0][k] * d;
            v[1] += J.Jt[1][k] * d;
        }
    }
}

// BuildRSD.m:29-40: r, vx, vy, vr, vt of one observation; xp, yp are the POST-update values (BuildRSD.m:14-27).
FEBA_HD void rsd_row(double x, double y, double xp, double yp, const double v[2], double* __restrict__ r) {
    const double xb = x - xp, yb = y - yp;
    const double theta = atan2(yb, xb), Phi = atan2(v[1], v[0]);
    const double vd = sqrt(v[0] * v[0] + v[1] * v[1]);
    r[0] = sqrt(xb * xb + yb * yb);
    r[1] = v[0];
    r[2] = v[1];
    r[3] = vd * cos(theta - Phi);
    r[4] = vd * sin(theta - Phi);
}

}  // namespace feba
