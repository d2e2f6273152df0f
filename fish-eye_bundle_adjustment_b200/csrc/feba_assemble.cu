// Fused BuildAwG + block normal equations + Schur elimination of the object points, without
// atomics on the reduced system: a point pass, an image pass and an image-pair pass over a
// schedule that is built once per problem (the sparsity of the network does not change between
// Gauss-Newton iterations).
//
// Reference path replaced: functions/BuildAwG.m:46-512 (A, w), main.m:424-425 (u = A'Pw, N = A'PA)
// and the elimination of the tie-point unknowns that the explicit inverse of main.m:432/442
// performs implicitly.  A and N are never formed.
//
// Algebra (per tie point p, observations a = 1..m in images i_a; P = diag(1/sx^2, 1/sy^2)):
//   V  = sum_a Jt_a' P Jt_a = L L'                       3x3, Cholesky in registers
//   Z_a = P Jt_a L^-T            (2x3)    ut = L^-1 u_p,  u_p = sum_a Jt_a' P w_a
//   Fc  = Wc L^-T                (NC x 3) Wc = sum_a Jc_a' P Jt_a
//   W V^-1 W' = (W L^-T)(W L^-T)'  so with  C_ab = Z_a Z_b' (2x2):
//     S[i_a, i_b]   -= Je_a' C_ab Je_b                   image-pair pass   (a != b)
//     S[i_a, i_a]   += Je_a' (P - C_aa) Je_a             image pass
//     g[i_a]        += Je_a' r_a,   r_a = P w_a - Z_a ut
//     S[cam, i_a]   += H_a Je_a,    H_a = Jc_a' P - Fc Z_a'
//     S[cam, cam]   += sum_a Jc_a' P Jc_a - Fc Fc',  g[cam] += sum_a Jc_a' P w_a - Fc ut
//                    = sum_a H_a P^-1 H_a'            = sum_a H_a P^-1 r_a      (sum_a Jc_a'Z_a = Fc, sum_a Z_a'P^-1 Z_a = I)
//                    single camera: k_cam_rec from the records; several cameras: in the multi-camera point pass
// Control points (not estimated) contribute the same terms with Z = 0, Fc = 0.
// The point pass stores, per observation, rec1 = {Je (2x6), Z (2x3)} and rec2 = {r (2), H (NC x 2)} in
// IMAGE-MAJOR order (an image's records are contiguous: the image pass streams them, the pair pass
// gathers monotonically inside two images' ranges); the other two passes only read those records.  Every entry of S has exactly one writer per
// kernel and partial sums are combined in a fixed order, so the result is run-to-run deterministic.
#include <cub/cub.cuh>

#include <cstdlib>
#include <vector>

#include "feba_dev.h"
#include "feba_kernels.h"
#include "feba_model.cuh"

namespace feba {

// ------------------------------------------------------------------------------------------
// point pass: one GROUP of G lanes (G = 32, or 16 when points have few observations: two points per
// warp, the halves run independently with their own __syncwarp masks) per object point, lanes over
// its observations (chunks of G)
template <int NK, int G>
struct alignas(16) PtSmem {
    static constexpr int NC = NK + 5;
    double rowJc[G][2][NC];
    double rowPJc[G][2][NC];      // P Jc: the weighted rows, so that the sums below are one product per term
    double rowJt[G][2][3];
    double rowPJt[G][2][3];
    double roww[G][2];
    double Wc[NC][3];
    double Fc[NC][3];
    double V[6];
    double up[3];
    int pos[G];
};

template <int NK, bool HAS_CAM, int G>
__global__ void __launch_bounds__(128) k_point_pass(DevProblem P) {
    constexpr int NC = NK + 5;
    constexpr int ND = NC * (NC + 1) / 2;          // packed camera-camera block
    constexpr int DPL = (ND + G - 1) / G;          // entries of D per lane
    constexpr int WPL = (NC * 3 + G - 1) / G;      // entries of Wc per lane
    constexpr int R2 = 2 + 2 * NC;                 // doubles per rec2 record
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PtSmem<NK, G>* sm_all = reinterpret_cast<PtSmem<NK, G>*>(smem_raw);
    const int lane = threadIdx.x & (G - 1);                  // lane inside the group
    const int wib = threadIdx.x / G;                         // group inside the CTA
    PtSmem<NK, G>& sm = sm_all[wib];
    const int nwarp = gridDim.x * (blockDim.x / G);          // groups in the grid
    const int gw = blockIdx.x * (blockDim.x / G) + wib;      // this group
    const unsigned gmask = G == 32 ? 0xffffffffu : (0xffffu << (16 * ((threadIdx.x & 31) >> 4)));
    const double pw[2] = {P.px, P.py};
    const int type = P.type;

    int dI[DPL], dJ[DPL];
    double dacc[DPL];
#pragma unroll
    for (int t = 0; t < DPL; ++t) {
        const int e = lane + G * t;
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;    // row of packed lower-triangular index e
        dI[t] = i;
        dJ[t] = e - i * (i + 1) / 2;
        dacc[t] = 0.0;
    }
    double gcacc = 0.0;

    for (int seg = gw; seg < P.n_seg; seg += nwarp) {
        const int beg = P.seg_start[seg], end = P.seg_start[seg + 1];
        const int pt = P.seg_pt[seg];
        const bool is_tie = P.pt_tie[pt] >= 0;
        const double X = P.xyz[3 * pt], Y = P.xyz[3 * pt + 1], Z = P.xyz[3 * pt + 2];
        const bool single = (end - beg) <= G;
        ObsJac<NK> J;
        bool have_J = false;

        // ---------------- pass 1: V, u_p, Wc and the direct camera-camera sums
        double vacc = 0.0;                          // lanes 0..5: V entries, 6..8: u_p
        double wcacc[WPL];
#pragma unroll
        for (int t = 0; t < WPL; ++t) wcacc[t] = 0.0;
        if (is_tie || HAS_CAM) {
            for (int c0 = beg; c0 < end; c0 += G) {
                const int o = c0 + lane;
                const bool act = o < end;
                if (act) {
                    const int img = P.oimg[o];
                    observation<NK, HAS_CAM>(type, P.ox[o], P.oy[o], P.img_tab + kImgStride * img,
                                             P.cam_tab + kCamStride * P.img_cam[img], X, Y, Z, J);
                }
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    if (HAS_CAM) {
#pragma unroll
                        for (int j = 0; j < NC; ++j) {
                            sm.rowJc[lane][r][j] = act ? J.Jc[r][j] : 0.0;
                            sm.rowPJc[lane][r][j] = act ? J.Jc[r][j] * pw[r] : 0.0;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        sm.rowJt[lane][r][k] = act ? J.Jt[r][k] : 0.0;
                        sm.rowPJt[lane][r][k] = act ? J.Jt[r][k] * pw[r] : 0.0;
                    }
                    sm.roww[lane][r] = act ? J.w[r] : 0.0;
                }
                __syncwarp(gmask);
                const int nrow = min(G, end - c0);
                // (x * p) * y with the product x * p taken once per row: the same roundings as x * p * y per term
                // (the direct camera sums  sum Jc'P Jc  and  sum Jc'P w  over ALL observations need no per-point
                    // structure: k_cam_direct forms them, one thread per observation)
                if (HAS_CAM) {
                    if (is_tie) {
#pragma unroll
                        for (int t = 0; t < WPL; ++t) {
                            const int e = lane + G * t;
                            if (e < NC * 3) {
                                const int i = e / 3, k = e - 3 * i;
                                double a = 0.0;
                                for (int l = 0; l < nrow; ++l)
                                    a += sm.rowPJc[l][0][i] * sm.rowJt[l][0][k] + sm.rowPJc[l][1][i] * sm.rowJt[l][1][k];
                                wcacc[t] += a;
                            }
                        }
                    }
                }
                if (is_tie && lane < 9) {
                    double a = 0.0;
                    if (lane < 6) {
                        const int i = lane < 1 ? 0 : (lane < 3 ? 1 : 2);
                        const int k = lane - i * (i + 1) / 2;
                        for (int l = 0; l < nrow; ++l)
                            a += sm.rowPJt[l][0][i] * sm.rowJt[l][0][k] + sm.rowPJt[l][1][i] * sm.rowJt[l][1][k];
                    } else {
                        const int k = lane - 6;
                        for (int l = 0; l < nrow; ++l)
                            a += sm.rowPJt[l][0][k] * sm.roww[l][0] + sm.rowPJt[l][1][k] * sm.roww[l][1];
                    }
                    vacc += a;
                }
                __syncwarp(gmask);
            }
            have_J = single;
        }
        // ---------------- V = L L', L^-1, ut = L^-1 u_p, Fc = Wc L^-T
        double i00 = 0, i10 = 0, i11 = 0, i20 = 0, i21 = 0, i22 = 0;
        double ut[3] = {0, 0, 0};
        if (is_tie) {
            if (lane < 6) sm.V[lane] = vacc;
            else if (lane < 9) sm.up[lane - 6] = vacc;
            if (HAS_CAM) {
#pragma unroll
                for (int t = 0; t < WPL; ++t) {
                    const int e = lane + G * t;
                    if (e < NC * 3) (&sm.Wc[0][0])[e] = wcacc[t];
                }
            }
            __syncwarp(gmask);
            const double v0 = sm.V[0], v1 = sm.V[1], v2 = sm.V[2], v3 = sm.V[3], v4 = sm.V[4], v5 = sm.V[5];
            const double l00 = sqrt(v0);
            const double l10 = v1 / l00, l20 = v3 / l00;
            const double l11 = sqrt(v2 - l10 * l10);
            const double l21 = (v4 - l20 * l10) / l11;
            const double l22 = sqrt(v5 - l20 * l20 - l21 * l21);
            i00 = 1.0 / l00; i11 = 1.0 / l11; i22 = 1.0 / l22;
            i10 = -l10 * i00 * i11;
            i21 = -l21 * i11 * i22;
            i20 = -(l20 * i00 + l21 * i10) * i22;
            const double u0 = sm.up[0], u1 = sm.up[1], u2 = sm.up[2];
            ut[0] = i00 * u0;
            ut[1] = i10 * u0 + i11 * u1;
            ut[2] = i20 * u0 + i21 * u1 + i22 * u2;
            if (HAS_CAM) {
                if (lane < NC) {
                    const double w0 = sm.Wc[lane][0], w1 = sm.Wc[lane][1], w2 = sm.Wc[lane][2];
                    sm.Fc[lane][0] = w0 * i00;
                    sm.Fc[lane][1] = w0 * i10 + w1 * i11;
                    sm.Fc[lane][2] = w0 * i20 + w1 * i21 + w2 * i22;
                }
                __syncwarp(gmask);
                if (!P.cam_rec) {          // else k_cam_rec forms the whole camera block from the records
#pragma unroll
                    for (int t = 0; t < DPL; ++t) {
                        if (lane + G * t < ND)
                            dacc[t] -= sm.Fc[dI[t]][0] * sm.Fc[dJ[t]][0] + sm.Fc[dI[t]][1] * sm.Fc[dJ[t]][1] +
                                       sm.Fc[dI[t]][2] * sm.Fc[dJ[t]][2];
                    }
                    if (lane < NC) gcacc -= sm.Fc[lane][0] * ut[0] + sm.Fc[lane][1] * ut[1] + sm.Fc[lane][2] * ut[2];
                }
            }
            if (P.pt_rec) {
                // what k_backsub_rec needs of this point besides the per-observation records
                constexpr int PR = 9 + 3 * NC;
                double* pr = P.pt_rec + (size_t)PR * P.pt_tie[pt];
                if (lane < 9) {
                    const double v = lane == 0 ? i00 : lane == 1 ? i10 : lane == 2 ? i11 : lane == 3 ? i20 : lane == 4 ? i21
                                   : lane == 5 ? i22 : lane == 6 ? ut[0] : lane == 7 ? ut[1] : ut[2];
                    pr[lane] = v;
                }
                if (HAS_CAM)
                    for (int e = lane; e < 3 * NC; e += G) pr[9 + e] = (&sm.Fc[0][0])[e];
            }
        }
        // ---------------- pass 2: per-observation records, staged through shared memory (the
        // pass-1 row buffers are dead by now) and written as 16-byte units, contiguous per record,
        // at the observation's position in the image-major record arrays
        double* buf = &sm.rowJc[0][0][0];              // rowJc .. roww: (4 NC + 14) * G doubles, >= G * max(18, R2)
        for (int a0 = beg; a0 < end; a0 += G) {
            const int o = a0 + lane;
            const bool act = o < end;
            const int nrow = min(G, end - a0);
            double Zm[2][3], ra[2];
            if (act) {
                if (!have_J) {
                    const int img = P.oimg[o];
                    observation<NK, HAS_CAM>(type, P.ox[o], P.oy[o], P.img_tab + kImgStride * img,
                                             P.cam_tab + kCamStride * P.img_cam[img], X, Y, Z, J);
                }
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const double t0 = pw[r] * J.Jt[r][0], t1 = pw[r] * J.Jt[r][1], t2 = pw[r] * J.Jt[r][2];
                    Zm[r][0] = t0 * i00;                     // zero for control points (i.. = 0)
                    Zm[r][1] = t0 * i10 + t1 * i11;
                    Zm[r][2] = t0 * i20 + t1 * i21 + t2 * i22;
                    ra[r] = pw[r] * J.w[r] - (Zm[r][0] * ut[0] + Zm[r][1] * ut[1] + Zm[r][2] * ut[2]);
                }
                sm.pos[lane] = P.ipos ? P.ipos[o] : o;      // image-major (round-1 form) or the observation's own position
                double* r1 = buf + kRec1 * lane;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) r1[6 * r + i] = J.Je[r][i];
#pragma unroll
                    for (int k = 0; k < 3; ++k) r1[12 + 3 * r + k] = Zm[r][k];
                }
            }
            __syncwarp(gmask);
            for (int u = lane; u < nrow * (kRec1 / 2); u += G) {
                const int k = u / (kRec1 / 2), part = u - k * (kRec1 / 2);
                reinterpret_cast<double2*>(P.rec1 + (size_t)kRec1 * sm.pos[k])[part] =
                    reinterpret_cast<const double2*>(buf + kRec1 * k)[part];
            }
            __syncwarp(gmask);
            if (act) {
                double* r2 = buf + R2 * lane;
                r2[0] = ra[0];
                r2[1] = ra[1];
                if (HAS_CAM) {
#pragma unroll
                    for (int j = 0; j < NC; ++j) {
                        double f0 = 0, f1 = 0, f2 = 0;
                        if (is_tie) { f0 = sm.Fc[j][0]; f1 = sm.Fc[j][1]; f2 = sm.Fc[j][2]; }
#pragma unroll
                        for (int r = 0; r < 2; ++r)
                            r2[2 + 2 * j + r] = J.Jc[r][j] * pw[r] - (f0 * Zm[r][0] + f1 * Zm[r][1] + f2 * Zm[r][2]);
                    }
                }
            }
            __syncwarp(gmask);
            {
                constexpr int U2 = HAS_CAM ? R2 / 2 : 1;     // without camera unknowns only r (2) is used
                for (int u = lane; u < nrow * U2; u += G) {
                    const int k = u / U2, part = u - k * U2;
                    reinterpret_cast<double2*>(P.rec2 + (size_t)R2 * sm.pos[k])[part] =
                        reinterpret_cast<const double2*>(buf + R2 * k)[part];
                }
            }
            __syncwarp(gmask);
        }
    }
    // camera-camera block and camera right-hand side of this warp -> partial buffer (fixed-order sum later)
    if (HAS_CAM) {
        double* part = P.cam_part + (size_t)kCamPart * gw;
#pragma unroll
        for (int t = 0; t < DPL; ++t)
            if (lane + G * t < ND) part[lane + G * t] = dacc[t];
        if (lane < NC) part[ND + lane] = gcacc;
    }
}

// ------------------------------------------------------------------------------------------
// Multi-camera variant of the point pass (n_cam > 1 with camera unknowns): a point may be seen by
// images of up to kMaxCamPt different cameras.  Per camera c of the point: Wc_c = sum_{a in c} Jc_a'PJt_a,
// Fc_c = Wc_c L^-T.  Own-camera terms go through the records exactly as in the single-camera
// kernel (H_a uses Fc of the observation's camera, the image pass adds H Je to that camera's rows);
// camera x camera blocks (all pairs), the camera right-hand sides and the cross terms
// -(Fc_c Z_a') Je_a for c != camera(a) are added to S with atomics (this path is not bit-reproducible
// from run to run; the single-camera path is).
constexpr int kMaxCamPt = 4;

template <int NK, int G>
struct alignas(16) PtSmemMC {
    static constexpr int NC = NK + 5;
    double rowJc[G][2][NC];
    double rowJt[G][2][3];
    double roww[G][2];
    double Wc[kMaxCamPt][NC][3];
    double Fc[kMaxCamPt][NC][3];
    double V[6];
    double up[3];
    int pos[G];
    int camrow[G];
};

template <int NK, int G>
__global__ void __launch_bounds__(128) k_point_pass_mc(DevProblem P, int* __restrict__ info) {
    constexpr int NC = NK + 5;
    constexpr int ND = NC * (NC + 1) / 2;
    constexpr int R2 = 2 + 2 * NC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PtSmemMC<NK, G>* sm_all = reinterpret_cast<PtSmemMC<NK, G>*>(smem_raw);
    const int lane = threadIdx.x & (G - 1);
    const int wib = threadIdx.x / G;
    PtSmemMC<NK, G>& sm = sm_all[wib];
    const int nwarp = gridDim.x * (blockDim.x / G);
    const int gw = blockIdx.x * (blockDim.x / G) + wib;
    const unsigned gmask = G == 32 ? 0xffffffffu : (0xffffu << (16 * ((threadIdx.x & 31) >> 4)));
    const double pw[2] = {P.px, P.py};
    const int type = P.type;
    const size_t aug = (size_t)P.n_pad;

    for (int seg = gw; seg < P.n_seg; seg += nwarp) {
        const int beg = P.seg_start[seg], end = P.seg_start[seg + 1];
        const int pt = P.seg_pt[seg];
        const bool is_tie = P.pt_tie[pt] >= 0;
        const double X = P.xyz[3 * pt], Y = P.xyz[3 * pt + 1], Z = P.xyz[3 * pt + 2];
        const bool single = (end - beg) <= G;
        ObsJac<NK> J;
        int mycam = 0;
        int cams[kMaxCamPt] = {-1, -1, -1, -1};
        int ncam = 0;
        // zero the per-camera accumulators
        for (int e = lane; e < kMaxCamPt * NC * 3; e += G) (&sm.Wc[0][0][0])[e] = 0.0;
        double vacc = 0.0;
        __syncwarp(gmask);
        // ---------------- pass 1
        for (int c0 = beg; c0 < end; c0 += G) {
            const int o = c0 + lane;
            const bool act = o < end;
            if (act) {
                const int img = P.oimg[o];
                mycam = P.img_cam[img];
                observation<NK, true>(type, P.ox[o], P.oy[o], P.img_tab + kImgStride * img,
                                      P.cam_tab + kCamStride * mycam, X, Y, Z, J);
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int j = 0; j < NC; ++j) sm.rowJc[lane][r][j] = act ? J.Jc[r][j] : 0.0;
#pragma unroll
                for (int k = 0; k < 3; ++k) sm.rowJt[lane][r][k] = act ? J.Jt[r][k] : 0.0;
                sm.roww[lane][r] = act ? J.w[r] : 0.0;
            }
            sm.camrow[lane] = act ? mycam : -1;
            __syncwarp(gmask);
            const int nrow = min(G, end - c0);
            // camera list of the point, first-occurrence order (every lane builds the same list)
            for (int l = 0; l < nrow; ++l) {
                const int c = sm.camrow[l];
                bool found = false;
#pragma unroll
                for (int k = 0; k < kMaxCamPt; ++k) found = found || (k < ncam && cams[k] == c);
                if (!found) {
                    if (ncam < kMaxCamPt) {
#pragma unroll
                        for (int k = 0; k < kMaxCamPt; ++k)
                            if (k == ncam) cams[k] = c;
                        ++ncam;
                    } else if (lane == 0) {
                        atomicExch(info, 3);
                    }
                }
            }
            // direct camera terms of this chunk: Jc'PJc (lower, own camera) and Jc'Pw
            for (int e = lane; e < ND + NC; e += G) {
                int i = 0, j = 0;
                if (e < ND) {
                    while ((i + 1) * (i + 2) / 2 <= e) ++i;
                    j = e - i * (i + 1) / 2;
                } else {
                    i = e - ND;
                }
                for (int k = 0; k < ncam; ++k) {
                    double a = 0.0;
                    for (int l = 0; l < nrow; ++l) {
                        if (sm.camrow[l] != cams[k]) continue;
                        if (e < ND)
                            a += sm.rowJc[l][0][i] * pw[0] * sm.rowJc[l][0][j] + sm.rowJc[l][1][i] * pw[1] * sm.rowJc[l][1][j];
                        else
                            a += sm.rowJc[l][0][i] * pw[0] * sm.roww[l][0] + sm.rowJc[l][1][i] * pw[1] * sm.roww[l][1];
                    }
                    if (a == 0.0) continue;
                    const size_t rc = (size_t)P.off_cam + (size_t)P.uc * cams[k];
                    if (e < ND) {
                        if (P.ccol[i] >= 0 && P.ccol[j] >= 0)
                            atomicAdd(&P.S[(rc + P.ccol[i]) + (size_t)P.ld * (rc + P.ccol[j])], a);
                    } else if (P.ccol[i] >= 0) {
                        atomicAdd(&P.S[aug + (size_t)P.ld * (rc + P.ccol[i])], a);
                    }
                }
            }
            if (is_tie) {
                for (int e = lane; e < NC * 3; e += G) {
                    const int i = e / 3, k3 = e - 3 * i;
                    for (int k = 0; k < ncam; ++k) {
                        double a = 0.0;
                        for (int l = 0; l < nrow; ++l)
                            if (sm.camrow[l] == cams[k])
                                a += sm.rowJc[l][0][i] * pw[0] * sm.rowJt[l][0][k3] + sm.rowJc[l][1][i] * pw[1] * sm.rowJt[l][1][k3];
                        sm.Wc[k][i][k3] += a;                  // entry owned by this lane
                    }
                }
                if (lane < 9) {
                    double a = 0.0;
                    if (lane < 6) {
                        const int i = lane < 1 ? 0 : (lane < 3 ? 1 : 2);
                        const int k = lane - i * (i + 1) / 2;
                        for (int l = 0; l < nrow; ++l)
                            a += sm.rowJt[l][0][i] * pw[0] * sm.rowJt[l][0][k] + sm.rowJt[l][1][i] * pw[1] * sm.rowJt[l][1][k];
                    } else {
                        const int k = lane - 6;
                        for (int l = 0; l < nrow; ++l)
                            a += sm.rowJt[l][0][k] * pw[0] * sm.roww[l][0] + sm.rowJt[l][1][k] * pw[1] * sm.roww[l][1];
                    }
                    vacc += a;
                }
            }
            __syncwarp(gmask);
        }
        const bool have_J = single;
        // ---------------- V = L L', ut, Fc per camera, camera x camera Schur terms
        double i00 = 0, i10 = 0, i11 = 0, i20 = 0, i21 = 0, i22 = 0;
        double ut[3] = {0, 0, 0};
        if (is_tie) {
            if (lane < 6) sm.V[lane] = vacc;
            else if (lane < 9) sm.up[lane - 6] = vacc;
            __syncwarp(gmask);
            const double v0 = sm.V[0], v1 = sm.V[1], v2 = sm.V[2], v3 = sm.V[3], v4 = sm.V[4], v5 = sm.V[5];
            const double l00 = sqrt(v0);
            const double l10 = v1 / l00, l20 = v3 / l00;
            const double l11 = sqrt(v2 - l10 * l10);
            const double l21 = (v4 - l20 * l10) / l11;
            const double l22 = sqrt(v5 - l20 * l20 - l21 * l21);
            i00 = 1.0 / l00; i11 = 1.0 / l11; i22 = 1.0 / l22;
            i10 = -l10 * i00 * i11;
            i21 = -l21 * i11 * i22;
            i20 = -(l20 * i00 + l21 * i10) * i22;
            const double u0 = sm.up[0], u1 = sm.up[1], u2 = sm.up[2];
            ut[0] = i00 * u0;
            ut[1] = i10 * u0 + i11 * u1;
            ut[2] = i20 * u0 + i21 * u1 + i22 * u2;
            for (int e = lane; e < ncam * NC; e += G) {
                const int k = e / NC, j = e - k * NC;
                const double w0 = sm.Wc[k][j][0], w1 = sm.Wc[k][j][1], w2 = sm.Wc[k][j][2];
                sm.Fc[k][j][0] = w0 * i00;
                sm.Fc[k][j][1] = w0 * i10 + w1 * i11;
                sm.Fc[k][j][2] = w0 * i20 + w1 * i21 + w2 * i22;
            }
            __syncwarp(gmask);
            for (int k1 = 0; k1 < ncam; ++k1) {
                const size_t r1 = (size_t)P.off_cam + (size_t)P.uc * cams[k1];
                for (int e = lane; e < NC; e += G)             // right-hand side: -Fc ut
                    if (P.ccol[e] >= 0)
                        atomicAdd(&P.S[aug + (size_t)P.ld * (r1 + P.ccol[e])],
                                  -(sm.Fc[k1][e][0] * ut[0] + sm.Fc[k1][e][1] * ut[1] + sm.Fc[k1][e][2] * ut[2]));
                for (int k2 = 0; k2 < ncam; ++k2) {
                    const size_t r2 = (size_t)P.off_cam + (size_t)P.uc * cams[k2];
                    if (r1 < r2) continue;                      // lower block triangle: (k1 rows, k2 columns)
                    for (int e = lane; e < NC * NC; e += G) {
                        const int i = e / NC, j = e - i * NC;
                        if (P.ccol[i] < 0 || P.ccol[j] < 0) continue;
                        if (k1 == k2 && j > i) continue;
                        atomicAdd(&P.S[(r1 + P.ccol[i]) + (size_t)P.ld * (r2 + P.ccol[j])],
                                  -(sm.Fc[k1][i][0] * sm.Fc[k2][j][0] + sm.Fc[k1][i][1] * sm.Fc[k2][j][1] +
                                    sm.Fc[k1][i][2] * sm.Fc[k2][j][2]));
                    }
                }
            }
        }
        // ---------------- pass 2: records + cross-camera terms
        double* buf = &sm.rowJc[0][0][0];
        for (int a0 = beg; a0 < end; a0 += G) {
            const int o = a0 + lane;
            const bool act = o < end;
            const int nrow = min(G, end - a0);
            double Zm[2][3], ra[2];
            int img = 0;
            if (act) {
                img = P.oimg[o];
                if (!have_J) {
                    mycam = P.img_cam[img];
                    observation<NK, true>(type, P.ox[o], P.oy[o], P.img_tab + kImgStride * img,
                                          P.cam_tab + kCamStride * mycam, X, Y, Z, J);
                }
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const double t0 = pw[r] * J.Jt[r][0], t1 = pw[r] * J.Jt[r][1], t2 = pw[r] * J.Jt[r][2];
                    Zm[r][0] = t0 * i00;
                    Zm[r][1] = t0 * i10 + t1 * i11;
                    Zm[r][2] = t0 * i20 + t1 * i21 + t2 * i22;
                    ra[r] = pw[r] * J.w[r] - (Zm[r][0] * ut[0] + Zm[r][1] * ut[1] + Zm[r][2] * ut[2]);
                }
                sm.pos[lane] = P.ipos ? P.ipos[o] : o;      // image-major (round-1 form) or the observation's own position
                double* r1 = buf + kRec1 * lane;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) r1[6 * r + i] = J.Je[r][i];
#pragma unroll
                    for (int k = 0; k < 3; ++k) r1[12 + 3 * r + k] = Zm[r][k];
                }
            }
            __syncwarp(gmask);
            for (int u = lane; u < nrow * (kRec1 / 2); u += G) {
                const int k = u / (kRec1 / 2), part = u - k * (kRec1 / 2);
                reinterpret_cast<double2*>(P.rec1 + (size_t)kRec1 * sm.pos[k])[part] =
                    reinterpret_cast<const double2*>(buf + kRec1 * k)[part];
            }
            __syncwarp(gmask);
            if (act) {
                int kown = 0;
#pragma unroll
                for (int k = 0; k < kMaxCamPt; ++k)
                    if (cams[k] == mycam) kown = k;
                double* r2 = buf + R2 * lane;
                r2[0] = ra[0];
                r2[1] = ra[1];
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    double f0 = 0, f1 = 0, f2 = 0;
                    if (is_tie) { f0 = sm.Fc[kown][j][0]; f1 = sm.Fc[kown][j][1]; f2 = sm.Fc[kown][j][2]; }
#pragma unroll
                    for (int r = 0; r < 2; ++r)
                        r2[2 + 2 * j + r] = J.Jc[r][j] * pw[r] - (f0 * Zm[r][0] + f1 * Zm[r][1] + f2 * Zm[r][2]);
                }
                // cross terms: rows of the OTHER cameras of this point x columns of this image
                if (is_tie) {
                    const size_t col0 = (size_t)P.img_row[img];
                    for (int k = 0; k < ncam; ++k) {
                        if (k == kown) continue;
                        const size_t rc = (size_t)P.off_cam + (size_t)P.uc * cams[k];
                        for (int j = 0; j < NC; ++j) {
                            if (P.ccol[j] < 0) continue;
                            const double h0 = -(sm.Fc[k][j][0] * Zm[0][0] + sm.Fc[k][j][1] * Zm[0][1] + sm.Fc[k][j][2] * Zm[0][2]);
                            const double h1 = -(sm.Fc[k][j][0] * Zm[1][0] + sm.Fc[k][j][1] * Zm[1][1] + sm.Fc[k][j][2] * Zm[1][2]);
#pragma unroll
                            for (int i = 0; i < 6; ++i)
                                if (P.ecol[i] >= 0)
                                    atomicAdd(&P.S[(rc + P.ccol[j]) + (size_t)P.ld * (col0 + P.ecol[i])],
                                              h0 * J.Je[0][i] + h1 * J.Je[1][i]);
                        }
                    }
                }
            }
            __syncwarp(gmask);
            for (int u = lane; u < nrow * (R2 / 2); u += G) {
                const int k = u / (R2 / 2), part = u - k * (R2 / 2);
                reinterpret_cast<double2*>(P.rec2 + (size_t)R2 * sm.pos[k])[part] =
                    reinterpret_cast<const double2*>(buf + R2 * k)[part];
            }
            __syncwarp(gmask);
        }
    }
}

// Direct camera terms of a single-camera problem:  sum_a Jc_a' P Jc_a  (packed lower, ND entries) and
// sum_a Jc_a' P w_a  (NC) over ALL observations -- one thread per observation, its ND + NC sums in registers, a
// fixed-order reduction per CTA, one row of the camera partial buffer per CTA (k_cam_reduce adds the rows of
// the point pass, which hold the Schur parts -Fc Fc', -Fc ut, and these).  In the point pass the same sums cost
// ND + NC shared-memory row loops per point on lanes that are mostly idle there.
template <int NK>
__global__ void __launch_bounds__(256) k_cam_direct(DevProblem P, const int* __restrict__ opt, int row0) {
    constexpr int NC = NK + 5;
    constexpr int ND = NC * (NC + 1) / 2;
    constexpr int NE = ND + NC;
    __shared__ double red[8][NE];
    double acc[NE];
#pragma unroll
    for (int k = 0; k < NE; ++k) acc[k] = 0.0;
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < P.n_obs; o += (int64_t)gridDim.x * blockDim.x) {
        const int img = P.oimg[o], pt = opt[o];
        ObsJac<NK> J;
        observation<NK, true>(P.type, P.ox[o], P.oy[o], P.img_tab + kImgStride * img, P.cam_tab + kCamStride * P.img_cam[img],
                              P.xyz[3 * pt], P.xyz[3 * pt + 1], P.xyz[3 * pt + 2], J);
        int e = 0;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const double t0 = J.Jc[0][i] * P.px, t1 = J.Jc[1][i] * P.py;
#pragma unroll
            for (int j = 0; j <= i; ++j) acc[e++] += t0 * J.Jc[0][j] + t1 * J.Jc[1][j];
            acc[ND + i] += t0 * J.w[0] + t1 * J.w[1];
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NE; ++k) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], s);
        if (lane == 0) red[warp][k] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < NE) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += red[w][threadIdx.x];
        P.cam_part[(size_t)kCamPart * (row0 + blockIdx.x) + threadIdx.x] = tot;
    }
}

// Camera block and camera right-hand side of a single-camera problem from the records alone.  With
// H_a = P Jc_a - Z_a Fc' and r_a = P w_a - Z_a ut (rec2), sum_a Jc_a' Z_a = Fc and sum_a Z_a' P^-1 Z_a = I per point give
//   sum_a H_a' P^-1 H_a = sum_a Jc_a' P Jc_a - sum_p Fc Fc'      (the Schur-complemented camera block)
//   sum_a H_a' P^-1 r_a = sum_a Jc_a' P w_a  - sum_p Fc ut       (its right-hand side)
// (control points: Z = 0).  One streaming read of rec2 (176 B per observation for NK = 5), no Jacobian evaluation,
// and a sum of positive semidefinite terms instead of a difference of two large ones.  Same partial-buffer rows
// and fixed-order reduction as k_cam_direct.
template <int NK>
__global__ void __launch_bounds__(256) k_cam_rec(DevProblem P, int row0) {
    constexpr int NC = NK + 5;
    constexpr int ND = NC * (NC + 1) / 2;
    constexpr int NE = ND + NC;
    constexpr int R2 = 2 + 2 * NC;
    __shared__ double red[8][NE];
    double acc[NE];
#pragma unroll
    for (int k = 0; k < NE; ++k) acc[k] = 0.0;
    const double ipx = 1.0 / P.px, ipy = 1.0 / P.py;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < P.n_obs; q += (int64_t)gridDim.x * blockDim.x) {
        const double2* p2 = reinterpret_cast<const double2*>(P.rec2 + (size_t)R2 * q);
        double r[R2];
#pragma unroll
        for (int k = 0; k < R2 / 2; ++k) { const double2 v = p2[k]; r[2 * k] = v.x; r[2 * k + 1] = v.y; }
        int e = 0;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const double t0 = r[2 + 2 * i] * ipx, t1 = r[3 + 2 * i] * ipy;
#pragma unroll
            for (int j = 0; j <= i; ++j) acc[e++] += t0 * r[2 + 2 * j] + t1 * r[3 + 2 * j];
            acc[ND + i] += t0 * r[0] + t1 * r[1];
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NE; ++k) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], s);
        if (lane == 0) red[warp][k] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < NE) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += red[w][threadIdx.x];
        P.cam_part[(size_t)kCamPart * (row0 + blockIdx.x) + threadIdx.x] = tot;
    }
}

// Fixed-order sum of the per-warp camera partials into S (camera-camera block, lower) and g.
// Thread (e, s): entry e, slice s of the warps; slices are combined in order.
__global__ void __launch_bounds__(1024) k_cam_reduce(DevProblem P, int n_warps) {
    __shared__ double part[8][128];
    const int NC = P.NC, ND = NC * (NC + 1) / 2, NE = ND + NC;
    const int e = threadIdx.x & 127, s = threadIdx.x >> 7;
    double acc = 0.0;
    if (e < NE)
        for (int w = s; w < n_warps; w += 8) acc += P.cam_part[(size_t)kCamPart * w + e];
    part[s][e] = acc;
    __syncthreads();
    if (s == 0 && e < NE) {
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += part[k][e];
        if (e < ND) {
            int i = 0;
            while ((i + 1) * (i + 2) / 2 <= e) ++i;
            const int j = e - i * (i + 1) / 2;
            if (P.ccol[i] >= 0 && P.ccol[j] >= 0)      // slots are monotone in parameter order: row >= col
                P.S[(size_t)(P.off_cam + P.ccol[i]) + (size_t)P.ld * (P.off_cam + P.ccol[j])] += tot;
        } else {
            const int j = e - ND;
            if (P.ccol[j] >= 0) P.S[(size_t)P.n_pad + (size_t)P.ld * (P.off_cam + P.ccol[j])] += tot;
        }
    }
}

// ------------------------------------------------------------------------------------------
// image pass: one CTA per image, threads over its observations; diagonal 6x6 block, right-hand
// side and camera x image block of the image.  Two sweeps keep the accumulators in registers.
template <int NK, bool HAS_CAM>
__global__ void __launch_bounds__(128) k_image_pass(DevProblem P) {
    constexpr int NC = NK + 5;
    constexpr int R2 = 2 + 2 * NC;
    constexpr int NA = HAS_CAM ? (NC * 6 > 27 ? NC * 6 : 27) : 27;
    __shared__ double part[4][NA];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int img = blockIdx.x; img < P.n_img; img += gridDim.x) {
        const int beg = P.img_start[img], end = P.img_start[img + 1];
        const size_t col0 = (size_t)P.img_row[img];
        // ---- sweep 1: Je' (P - Z Z') Je  (21 entries, lower) and Je' r (6)
        {
            double acc[27];
#pragma unroll
            for (int k = 0; k < 27; ++k) acc[k] = 0.0;
            for (int t = beg + tid; t < end; t += 128) {
                const double2* r1v = reinterpret_cast<const double2*>(P.rec1 + (size_t)kRec1 * t);
                double r1[kRec1];
#pragma unroll
                for (int k = 0; k < kRec1 / 2; ++k) { const double2 v = r1v[k]; r1[2 * k] = v.x; r1[2 * k + 1] = v.y; }
                const double2 rav = *reinterpret_cast<const double2*>(P.rec2 + (size_t)R2 * t);
                double Je[2][6], Zm[2][3];
#pragma unroll
                for (int k = 0; k < 6; ++k) { Je[0][k] = r1[k]; Je[1][k] = r1[6 + k]; }
#pragma unroll
                for (int k = 0; k < 3; ++k) { Zm[0][k] = r1[12 + k]; Zm[1][k] = r1[15 + k]; }
                const double ra0 = rav.x, ra1 = rav.y;
                const double p00 = P.px - (Zm[0][0] * Zm[0][0] + Zm[0][1] * Zm[0][1] + Zm[0][2] * Zm[0][2]);
                const double p01 = -(Zm[0][0] * Zm[1][0] + Zm[0][1] * Zm[1][1] + Zm[0][2] * Zm[1][2]);
                const double p11 = P.py - (Zm[1][0] * Zm[1][0] + Zm[1][1] * Zm[1][1] + Zm[1][2] * Zm[1][2]);
                int e = 0;
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const double t0 = p00 * Je[0][i] + p01 * Je[1][i], t1 = p01 * Je[0][i] + p11 * Je[1][i];
#pragma unroll
                    for (int j = 0; j <= i; ++j) acc[e++] += t0 * Je[0][j] + t1 * Je[1][j];
                }
#pragma unroll
                for (int i = 0; i < 6; ++i) acc[21 + i] += Je[0][i] * ra0 + Je[1][i] * ra1;
            }
#pragma unroll
            for (int k = 0; k < 27; ++k) {
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], s);
                if (lane == 0) part[warp][k] = acc[k];
            }
            __syncthreads();
            if (tid < 27) {
                const double tot = part[0][tid] + part[1][tid] + part[2][tid] + part[3][tid];
                if (tid < 21) {
                    int i = 0;
                    while ((i + 1) * (i + 2) / 2 <= tid) ++i;
                    const int j = tid - i * (i + 1) / 2;
                    if (P.ecol[i] >= 0 && P.ecol[j] >= 0)
                        P.S[(col0 + P.ecol[i]) + (size_t)P.ld * (col0 + P.ecol[j])] += tot;
                } else if (P.ecol[tid - 21] >= 0) {
                    P.S[(size_t)P.n_pad + (size_t)P.ld * (col0 + P.ecol[tid - 21])] += tot;
                }
            }
            __syncthreads();
        }
        // ---- sweep 2: camera x image block  sum_a H_a Je_a  (NC x 6)
        if (HAS_CAM) {
            double acc[NC * 6];
#pragma unroll
            for (int k = 0; k < NC * 6; ++k) acc[k] = 0.0;
            for (int t = beg + tid; t < end; t += 128) {
                const double2* r1v = reinterpret_cast<const double2*>(P.rec1 + (size_t)kRec1 * t);
                const double2* r2v = reinterpret_cast<const double2*>(P.rec2 + (size_t)R2 * t + 2);
                double Je[2][6];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double2 v0 = r1v[k], v1 = r1v[3 + k];
                    Je[0][2 * k] = v0.x; Je[0][2 * k + 1] = v0.y;
                    Je[1][2 * k] = v1.x; Je[1][2 * k + 1] = v1.y;
                }
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    const double2 hv = r2v[j];
                    const double h0 = hv.x, h1 = hv.y;
#pragma unroll
                    for (int i = 0; i < 6; ++i) acc[6 * j + i] += h0 * Je[0][i] + h1 * Je[1][i];
                }
            }
#pragma unroll
            for (int k = 0; k < NC * 6; ++k) {
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], s);
                if (lane == 0) part[warp][k] = acc[k];
            }
            __syncthreads();
            if (tid < NC * 6) {
                const int j = tid / 6, i = tid - 6 * j;
                if (P.ccol[j] >= 0 && P.ecol[i] >= 0) {
                    const double tot = part[0][tid] + part[1][tid] + part[2][tid] + part[3][tid];
                    const int cam = P.img_cam[img];
                    P.S[(size_t)(P.off_cam + P.uc * cam + P.ccol[j]) + (size_t)P.ld * (col0 + P.ecol[i])] += tot;
                }
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------
// image-pair pass: one warp per block (i_a >= i_b) of the schedule, lanes over the block's
// (a, b) observation pairs (the tie points the two images share): S[i_a, i_b] -= sum Je_a' Z_a Z_b' Je_b.
__global__ void __launch_bounds__(128) k_pair_pass(DevProblem P) {
    const int lane = threadIdx.x & 31;
    const int nwarp = gridDim.x * (blockDim.x >> 5);
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (int blk = gw; blk < P.n_blocks; blk += nwarp) {
        const int4 b = P.blocks[blk];                  // x: image a, y: image b, z: first pair, w: pairs
        double acc[36];
#pragma unroll
        for (int k = 0; k < 36; ++k) acc[k] = 0.0;
        for (int t = lane; t < b.w; t += 32) {
            const int2 pr = P.pairs[(size_t)b.z + t];
            // 144-byte records, 16-byte aligned: nine 16-byte loads each, all in flight together
            const double2* ra2 = reinterpret_cast<const double2*>(P.rec1 + (size_t)kRec1 * pr.x);
            const double2* rb2 = reinterpret_cast<const double2*>(P.rec1 + (size_t)kRec1 * pr.y);
            double ra[kRec1], rb[kRec1];
#pragma unroll
            for (int k = 0; k < kRec1 / 2; ++k) {
                const double2 va = ra2[k], vb = rb2[k];
                ra[2 * k] = va.x; ra[2 * k + 1] = va.y;
                rb[2 * k] = vb.x; rb[2 * k + 1] = vb.y;
            }
            const double c00 = ra[12] * rb[12] + ra[13] * rb[13] + ra[14] * rb[14];
            const double c01 = ra[12] * rb[15] + ra[13] * rb[16] + ra[14] * rb[17];
            const double c10 = ra[15] * rb[12] + ra[16] * rb[13] + ra[17] * rb[14];
            const double c11 = ra[15] * rb[15] + ra[16] * rb[16] + ra[17] * rb[17];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const double e0 = ra[i] * c00 + ra[6 + i] * c10, e1 = ra[i] * c01 + ra[6 + i] * c11;   // (Je_a' C)(i, :)
#pragma unroll
                for (int j = 0; j < 6; ++j) acc[6 * i + j] += e0 * rb[j] + e1 * rb[6 + j];
            }
        }
#pragma unroll
        for (int k = 0; k < 36; ++k) {
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], s);
        }
        // lane k (and k + 32) writes entry k; every lane holds all sums after the butterfly
        const size_t rowa = (size_t)P.img_row[b.x], colb = (size_t)P.img_row[b.y];   // row(a) >= row(b) by construction
        const bool diag = b.x == b.y;
#pragma unroll
        for (int k = 0; k < 36; ++k) {
            if ((k & 31) == lane) {
                const int i = k / 6, j = k - 6 * i;
                if (P.ecol[i] >= 0 && P.ecol[j] >= 0 && (!diag || j <= i))
                    P.S[(rowa + P.ecol[i]) + (size_t)P.ld * (colb + P.ecol[j])] -= acc[k];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Chunk form of the image / image-pair passes (feba_chunks.h): records are chunk-major (the point pass wrote them
// at the observations' own positions), one CTA per chunk of consecutive points, one LANE per item:
//   image slot  : the chunk's observations of one image -> diagonal block (21), right-hand side (6),
//                 camera x image block (NC x 6), the sums of k_image_pass;
//   block slot  : the chunk's observation pairs of one image pair -> 6 x 6 block, the sums of k_pair_pass;
// results go to PARTIAL blocks (one writer each).  The ~200 KB of records of a chunk are read from HBM once and
// re-read from L1/L2 (each record enters ~9 pairs).
constexpr int kImgPartDev = 108;      // = feba_chunks.h::kImgPart: 27 + 6 * 13, padded to a multiple of 4

// Shared-memory copy of a chunk's records: per observation Je (2x6), Y = Je' Z (6x3, formed once here instead of
// once per pair) and r (2), row stride 33 doubles (odd: the 32 lanes of a warp, each reading ITS OWN record with
// 8-byte loads, hit different banks).  Gathering whole 144-byte records per lane from global memory costs one L1
// wavefront per 16-byte load per lane -- the bound the image-pair pass of round 1 ran into (810 M wavefronts =
// 2.9 ms on BASELINE configs[3]); from shared memory the same gather is several times cheaper, and every record
// comes in from HBM once.
constexpr int kRecS = 33;
constexpr int kChunkObsDev = 416;      // = feba_chunks.h::kChunkObs; larger chunks (one huge point) gather from global
constexpr size_t kChunkSmem = (size_t)kChunkObsDev * kRecS * sizeof(double);    // 109,824 B: two CTAs per SM

// Sum v[0..31] over the lanes of a warp so that lane l ends with entry l (in v[0]): 31 shuffles instead of the
// 160 of a full butterfly per entry.
__device__ __forceinline__ double warp_reduce_scatter32(double (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const double send = upper ? v[i] : v[i + s];
            const double keep = upper ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

template <int NK, bool HAS_CAM>
__global__ void __launch_bounds__(256, 2) k_chunk_reduce(DevProblem P, ChunkDev C) {
    constexpr int NC = NK + 5;
    constexpr int R2 = 2 + 2 * NC;
    extern __shared__ __align__(16) double srec[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = blockIdx.x; c < C.n_chunks; c += gridDim.x) {
        const size_t o0 = (size_t)C.obs0[c];
        const int n_ob = C.obs0[c + 1] - C.obs0[c];
        const int s_lo = C.img0[c], n_is = C.img0[c + 1] - s_lo;
        const int b_lo = C.blk0[c], n_bs = C.blk0[c + 1] - b_lo;
        const bool staged = n_ob <= kChunkObsDev;
        __syncthreads();                                         // the previous chunk's readers are done
        if (staged) {
            // one thread per observation: its ten 16-byte loads are in flight together
            for (int ob = threadIdx.x; ob < n_ob; ob += blockDim.x) {
                const double2* q = reinterpret_cast<const double2*>(P.rec1 + (size_t)kRec1 * (o0 + ob));
                double r[kRec1];
#pragma unroll
                for (int k = 0; k < kRec1 / 2; ++k) { const double2 v = q[k]; r[2 * k] = v.x; r[2 * k + 1] = v.y; }
                const double2 rav = *reinterpret_cast<const double2*>(P.rec2 + (size_t)R2 * (o0 + ob));
                double* d = srec + ob * kRecS;
#pragma unroll
                for (int k = 0; k < 12; ++k) d[k] = r[k];
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int k = 0; k < 3; ++k) d[12 + 3 * i + k] = r[i] * r[12 + k] + r[6 + i] * r[15 + k];
                d[30] = rav.x;
                d[31] = rav.y;
            }
        }
        __syncthreads();
        // Je (12) and Y (18) of chunk-local observation t
        auto rec_of = [&](int t, double* je, double* y) {
            if (staged) {
                const double* q = srec + t * kRecS;
                if (je) {
#pragma unroll
                    for (int k = 0; k < 12; ++k) je[k] = q[k];
                }
                if (y) {
#pragma unroll
                    for (int k = 0; k < 18; ++k) y[k] = q[12 + k];
                }
            } else {
                const double2* q = reinterpret_cast<const double2*>(P.rec1 + (size_t)kRec1 * (o0 + t));
                double r[kRec1];
#pragma unroll
                for (int k = 0; k < kRec1 / 2; ++k) { const double2 v = q[k]; r[2 * k] = v.x; r[2 * k + 1] = v.y; }
                if (je) {
#pragma unroll
                    for (int k = 0; k < 12; ++k) je[k] = r[k];
                }
                if (y) {
#pragma unroll
                    for (int i = 0; i < 6; ++i)
#pragma unroll
                        for (int k = 0; k < 3; ++k) y[3 * i + k] = r[i] * r[12 + k] + r[6 + i] * r[15 + k];
                }
            }
        };
        // one WARP per item, lanes over the item's observation pairs / observations (every lane reads ITS OWN
        // records from shared memory), sums combined by a reduce-scatter; items of a chunk are dealt out over the
        // eight warps.  The index of a warp's NEXT item is loaded while the current one is worked on.
        const int n_items = n_bs + n_is;
        for (int item = warp; item < n_items; item += (int)(blockDim.x >> 5)) {
            if (item < n_bs) {
                const int bs = b_lo + item;
                const int q0 = C.bslot_pair0[bs], q1 = C.bslot_pair0[bs + 1];
                double* out = C.blk_part + (size_t)36 * bs;
                // rows 0..2 and rows 3..5 of the 6 x 6 block in two sweeps over the pairs (18 sums live at a time:
                // the whole block plus two records would not fit 128 registers)
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    double acc[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) acc[k] = 0.0;
                    for (int q = q0 + lane; q < q1; q += 32) {
                        const unsigned int pr = C.pairs[q];
                        double yb[18];
                        rec_of((int)(pr >> 16), nullptr, yb);
                        const int ta = (int)(pr & 0xffffu);
                        double ya[9];
                        if (staged) {
                            const double* qa = srec + ta * kRecS + 12 + 9 * half;
#pragma unroll
                            for (int k = 0; k < 9; ++k) ya[k] = qa[k];
                        } else {
                            double yfull[18];
                            rec_of(ta, nullptr, yfull);
#pragma unroll
                            for (int k = 0; k < 9; ++k) ya[k] = half ? yfull[9 + k] : yfull[k];
                        }
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int j = 0; j < 6; ++j)
                                acc[6 * i + j] += ya[3 * i] * yb[3 * j] + ya[3 * i + 1] * yb[3 * j + 1] + ya[3 * i + 2] * yb[3 * j + 2];
                    }
                    const double v0 = warp_reduce_scatter32(acc, lane);
                    if (lane < 18) out[18 * half + lane] = v0;
                }
            } else {
                const int slot = s_lo + (item - n_bs);
                const int q0 = C.slot_obs0[slot], q1 = C.slot_obs0[slot + 1];
                double* out = C.img_part + (size_t)kImgPartDev * slot;
                {   // diagonal block Je'P Je - Y Y' (= Je'(P - Z Z')Je) and right-hand side Je' r
                    double acc[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) acc[k] = 0.0;
                    for (int q = q0 + lane; q < q1; q += 32) {
                        const int t = C.slot_obs[q];
                        double je[12], y[18];
                        rec_of(t, je, y);
                        double ra0, ra1;
                        if (staged) {
                            ra0 = srec[t * kRecS + 30];
                            ra1 = srec[t * kRecS + 31];
                        } else {
                            const double2 rav = *reinterpret_cast<const double2*>(P.rec2 + (size_t)R2 * (o0 + t));
                            ra0 = rav.x;
                            ra1 = rav.y;
                        }
                        int e = 0;
#pragma unroll
                        for (int i = 0; i < 6; ++i) {
                            const double t0 = P.px * je[i], t1 = P.py * je[6 + i];
#pragma unroll
                            for (int j = 0; j <= i; ++j)
                                acc[e++] += t0 * je[j] + t1 * je[6 + j] -
                                            (y[3 * i] * y[3 * j] + y[3 * i + 1] * y[3 * j + 1] + y[3 * i + 2] * y[3 * j + 2]);
                        }
#pragma unroll
                        for (int i = 0; i < 6; ++i) acc[21 + i] += je[i] * ra0 + je[6 + i] * ra1;
                    }
                    const double v0 = warp_reduce_scatter32(acc, lane);
                    if (lane < 27) out[lane] = v0;
                }
                if (HAS_CAM) {   // camera x image block  sum_a H_a Je_a  (NC x 6), 32 entries at a time
#pragma unroll
                    for (int g0 = 0; g0 < NC * 6; g0 += 30) {          // five camera rows = 30 entries per pass
                        double acc[32];
#pragma unroll
                        for (int k = 0; k < 32; ++k) acc[k] = 0.0;
                        for (int q = q0 + lane; q < q1; q += 32) {
                            const int t = C.slot_obs[q];
                            const double2* r2v = reinterpret_cast<const double2*>(P.rec2 + (size_t)R2 * (o0 + t) + 2);
                            double je[12];
                            rec_of(t, je, nullptr);
#pragma unroll
                            for (int jj = 0; jj < 5; ++jj) {
                                if (g0 / 6 + jj < NC) {
                                    const double2 hv = r2v[g0 / 6 + jj];
#pragma unroll
                                    for (int i = 0; i < 6; ++i) acc[6 * jj + i] += hv.x * je[i] + hv.y * je[6 + i];
                                }
                            }
                        }
                        const double v0 = warp_reduce_scatter32(acc, lane);
                        if (lane < 30 && g0 + lane < NC * 6) out[27 + g0 + lane] = v0;
                    }
                }
            }
        }
    }
}

// Final sums of the image partials: one CTA of 128 threads per image, thread e owns entry e of the partial
// (diagonal block, right-hand side, camera x image block); slots are added in chunk order.
__global__ void __launch_bounds__(128) k_sum_img_parts(DevProblem P, ChunkDev C) {
    const int NE = 27 + (P.uc > 0 ? 6 * P.NC : 0);
    const int e = threadIdx.x;
    for (int img = blockIdx.x; img < P.n_img; img += gridDim.x) {
        if (e >= NE) continue;
        double tot = 0.0;
        for (int q = C.timg_ptr[img]; q < C.timg_ptr[img + 1]; ++q)
            tot += C.img_part[(size_t)kImgPartDev * C.timg_slots[q] + e];
        const size_t col0 = (size_t)P.img_row[img];
        if (e < 21) {
            int i = 0;
            while ((i + 1) * (i + 2) / 2 <= e) ++i;
            const int j = e - i * (i + 1) / 2;
            if (P.ecol[i] >= 0 && P.ecol[j] >= 0) P.S[(col0 + P.ecol[i]) + (size_t)P.ld * (col0 + P.ecol[j])] += tot;
        } else if (e < 27) {
            if (P.ecol[e - 21] >= 0) P.S[(size_t)P.n_pad + (size_t)P.ld * (col0 + P.ecol[e - 21])] += tot;
        } else {
            const int j = (e - 27) / 6, i = (e - 27) - 6 * j;
            if (P.ccol[j] >= 0 && P.ecol[i] >= 0) {
                const int cam = P.img_cam[img];
                P.S[(size_t)(P.off_cam + P.uc * cam + P.ccol[j]) + (size_t)P.ld * (col0 + P.ecol[i])] += tot;
            }
        }
    }
}

// Final sums of the image-pair partials: one warp per distinct image pair (row(a) >= row(b)), lanes over the 36
// entries, slots in chunk order:  S[i_a, i_b] -= sum.
__global__ void __launch_bounds__(128) k_sum_blk_parts(DevProblem P, ChunkDev C) {
    const int lane = threadIdx.x & 31;
    const int nwarp = gridDim.x * (blockDim.x >> 5);
    for (int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < C.n_tblk; t += nwarp) {
        double a0 = 0.0, a1 = 0.0;
        for (int q = C.tblk_ptr[t]; q < C.tblk_ptr[t + 1]; ++q) {
            const double* p = C.blk_part + (size_t)36 * C.tblk_slots[q];
            a0 += p[lane];
            if (lane < 4) a1 += p[32 + lane];
        }
        const int ia = C.tblk_a[t], ib = C.tblk_b[t];
        const size_t rowa = (size_t)P.img_row[ia], colb = (size_t)P.img_row[ib];
        const bool diag = ia == ib;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = lane + 32 * h;
            if (k < 36) {
                const int i = k / 6, j = k - 6 * i;
                if (P.ecol[i] >= 0 && P.ecol[j] >= 0 && (!diag || j <= i))
                    P.S[(rowa + P.ecol[i]) + (size_t)P.ld * (colb + P.ecol[j])] -= (h ? a1 : a0);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// schedule construction (once per problem, on the device)

// A pair (o, b) of observations of one tie point belongs to the block (image of o, image of b) when the image of b
// comes EARLIER in the row order of the reduced system (img_row; the lower triangle is stored) -- or is the same
// image (diagonal block, both orders).
__global__ void k_pair_count(int64_t n_obs, const int* __restrict__ oseg, const int* __restrict__ seg_start,
                             const int* __restrict__ seg_pt, const int* __restrict__ pt_tie,
                             const int* __restrict__ oimg, const int* __restrict__ img_row,
                             long long* __restrict__ cnt) {
    const int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (o >= n_obs) return;
    const int seg = oseg[o];
    long long c = 0;
    if (pt_tie[seg_pt[seg]] >= 0) {
        const int ia = oimg[o], ra = img_row[ia];
        for (int b = seg_start[seg]; b < seg_start[seg + 1]; ++b) {
            const int ib = oimg[b];
            if (img_row[ib] < ra || (ib == ia && b != o)) ++c;
        }
    }
    cnt[o] = c;
}

__global__ void k_pair_fill(int64_t n_obs, int n_img, const int* __restrict__ oseg, const int* __restrict__ seg_start,
                            const int* __restrict__ seg_pt, const int* __restrict__ pt_tie,
                            const int* __restrict__ oimg, const int* __restrict__ img_row,
                            const long long* __restrict__ off,
                            unsigned long long* __restrict__ keys, unsigned long long* __restrict__ vals) {
    const int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (o >= n_obs) return;
    const int seg = oseg[o];
    if (pt_tie[seg_pt[seg]] < 0) return;
    const int ia = oimg[o], ra = img_row[ia];
    long long q = off[o];
    for (int b = seg_start[seg]; b < seg_start[seg + 1]; ++b) {
        const int ib = oimg[b];
        if (img_row[ib] < ra || (ib == ia && b != o)) {
            keys[q] = (unsigned long long)ia * (unsigned long long)n_img + (unsigned long long)ib;
            vals[q] = ((unsigned long long)(unsigned)o << 32) | (unsigned)b;
            ++q;
        }
    }
}

__global__ void k_blocks_fill(int n_blocks, int n_img, const unsigned long long* __restrict__ ukeys,
                              const long long* __restrict__ starts, const long long* __restrict__ counts,
                              int4* __restrict__ blocks) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_blocks) return;
    const unsigned long long key = ukeys[k];
    blocks[k] = make_int4((int)(key / (unsigned long long)n_img), (int)(key % (unsigned long long)n_img),
                          (int)starts[k], (int)counts[k]);
}

__global__ void k_pairs_unpack(long long n, const unsigned long long* __restrict__ vals,
                               const int* __restrict__ ipos, int2* __restrict__ pairs) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) pairs[i] = make_int2(ipos[(int)(vals[i] >> 32)], ipos[(int)(vals[i] & 0xffffffffull)]);
}

#define SCHED_CU(call)                      \
    do {                                    \
        cudaError_t e_ = (call);            \
        if (e_ != cudaSuccess) {            \
            for (void* p_ : tmp) cudaFree(p_); \
            return e_;                      \
        }                                   \
    } while (0)

// Builds P.pairs / P.blocks / P.n_blocks.  oseg: segment of every (point-major) observation.
// Device allocations that must outlive the call are returned through keep[].
cudaError_t build_pair_schedule(DevProblem& P, const int* d_oseg, long long* n_pairs_out, void** keep_pairs,
                                void** keep_blocks, cudaStream_t st) {
    std::vector<void*> tmp;
    auto talloc = [&](void** p, size_t bytes) {
        cudaError_t e = cudaMalloc(p, bytes ? bytes : 8);
        if (e == cudaSuccess) tmp.push_back(*p);
        return e;
    };
    *n_pairs_out = 0;
    *keep_pairs = *keep_blocks = nullptr;
    P.pairs = nullptr;
    P.blocks = nullptr;
    P.n_blocks = 0;
    const int64_t n = P.n_obs;
    if (n == 0 || P.n_tie == 0) return cudaSuccess;
    long long *cnt, *off;
    SCHED_CU(talloc((void**)&cnt, (size_t)(n + 1) * sizeof(long long)));
    SCHED_CU(talloc((void**)&off, (size_t)(n + 1) * sizeof(long long)));
    SCHED_CU(cudaMemsetAsync(cnt, 0, (size_t)(n + 1) * sizeof(long long), st));
    const int grid = (int)((n + 255) / 256);
    k_pair_count<<<grid, 256, 0, st>>>(n, d_oseg, P.seg_start, P.seg_pt, P.pt_tie, P.oimg, P.img_row, cnt);
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    SCHED_CU(cub::DeviceScan::ExclusiveSum(nullptr, scratch_bytes, cnt, off, (int)(n + 1), st));
    SCHED_CU(talloc(&scratch, scratch_bytes));
    SCHED_CU(cub::DeviceScan::ExclusiveSum(scratch, scratch_bytes, cnt, off, (int)(n + 1), st));
    long long n_pairs = 0;
    SCHED_CU(cudaMemcpyAsync(&n_pairs, off + n, sizeof(long long), cudaMemcpyDeviceToHost, st));
    SCHED_CU(cudaStreamSynchronize(st));
    *n_pairs_out = n_pairs;
    if (n_pairs == 0) {
        for (void* p_ : tmp) cudaFree(p_);
        return cudaSuccess;
    }
    if (n_pairs > 2000000000LL) {
        for (void* p_ : tmp) cudaFree(p_);
        return cudaErrorInvalidValue;
    }
    unsigned long long *keys, *vals, *keys2, *vals2;
    SCHED_CU(talloc((void**)&keys, (size_t)n_pairs * 8));
    SCHED_CU(talloc((void**)&vals, (size_t)n_pairs * 8));
    SCHED_CU(talloc((void**)&keys2, (size_t)n_pairs * 8));
    SCHED_CU(talloc((void**)&vals2, (size_t)n_pairs * 8));
    k_pair_fill<<<grid, 256, 0, st>>>(n, P.n_img, d_oseg, P.seg_start, P.seg_pt, P.pt_tie, P.oimg, P.img_row, off, keys,
                                      vals);
    int bits = 1;
    while (bits < 64 && (1ull << bits) < (unsigned long long)P.n_img * (unsigned long long)P.n_img) ++bits;
    void* s2 = nullptr;
    size_t s2_bytes = 0;
    SCHED_CU(cub::DeviceRadixSort::SortPairs(nullptr, s2_bytes, keys, keys2, vals, vals2, (int)n_pairs, 0, bits, st));
    SCHED_CU(talloc(&s2, s2_bytes));
    SCHED_CU(cub::DeviceRadixSort::SortPairs(s2, s2_bytes, keys, keys2, vals, vals2, (int)n_pairs, 0, bits, st));
    // runs of equal keys = blocks
    unsigned long long* ukeys = keys;          // reuse: unique keys (at most n_pairs)
    long long* counts;
    int* d_runs;
    SCHED_CU(talloc((void**)&counts, (size_t)n_pairs * sizeof(long long)));
    SCHED_CU(talloc((void**)&d_runs, sizeof(int)));
    void* s3 = nullptr;
    size_t s3_bytes = 0;
    SCHED_CU(cub::DeviceRunLengthEncode::Encode(nullptr, s3_bytes, keys2, ukeys, counts, d_runs, (int)n_pairs, st));
    SCHED_CU(talloc(&s3, s3_bytes));
    SCHED_CU(cub::DeviceRunLengthEncode::Encode(s3, s3_bytes, keys2, ukeys, counts, d_runs, (int)n_pairs, st));
    int n_blocks = 0;
    SCHED_CU(cudaMemcpyAsync(&n_blocks, d_runs, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCHED_CU(cudaStreamSynchronize(st));
    long long* starts;
    SCHED_CU(talloc((void**)&starts, (size_t)n_blocks * sizeof(long long)));
    {
        void* s4 = nullptr;
        size_t s4_bytes = 0;
        SCHED_CU(cub::DeviceScan::ExclusiveSum(nullptr, s4_bytes, counts, starts, n_blocks, st));
        SCHED_CU(talloc(&s4, s4_bytes));
        SCHED_CU(cub::DeviceScan::ExclusiveSum(s4, s4_bytes, counts, starts, n_blocks, st));
    }
    int4* blocks = nullptr;
    int2* pairs = nullptr;
    cudaError_t e = cudaMalloc((void**)&blocks, (size_t)n_blocks * sizeof(int4));
    if (e == cudaSuccess) e = cudaMalloc((void**)&pairs, (size_t)n_pairs * sizeof(int2));
    if (e != cudaSuccess) {
        if (blocks) cudaFree(blocks);
        for (void* p_ : tmp) cudaFree(p_);
        return e;
    }
    k_blocks_fill<<<(n_blocks + 255) / 256, 256, 0, st>>>(n_blocks, P.n_img, ukeys, starts, counts, blocks);
    k_pairs_unpack<<<(int)((n_pairs + 255) / 256), 256, 0, st>>>(n_pairs, vals2, P.ipos, pairs);
    e = cudaStreamSynchronize(st);
    for (void* p_ : tmp) cudaFree(p_);
    if (e != cudaSuccess) {
        cudaFree(blocks);
        cudaFree(pairs);
        return e;
    }
    P.pairs = pairs;
    P.blocks = blocks;
    P.n_blocks = n_blocks;
    *keep_pairs = pairs;
    *keep_blocks = blocks;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// launcher

#define FEBA_NK_DISPATCH2(NKV, HASCAM, CALL)                                  \
    switch (NKV) {                                                            \
        case 1: { constexpr int NK_ = 1; if (HASCAM) { constexpr bool HC_ = true; CALL; } else { constexpr bool HC_ = false; CALL; } } break; \
        case 2: { constexpr int NK_ = 2; if (HASCAM) { constexpr bool HC_ = true; CALL; } else { constexpr bool HC_ = false; CALL; } } break; \
        case 3: { constexpr int NK_ = 3; if (HASCAM) { constexpr bool HC_ = true; CALL; } else { constexpr bool HC_ = false; CALL; } } break; \
        case 4: { constexpr int NK_ = 4; if (HASCAM) { constexpr bool HC_ = true; CALL; } else { constexpr bool HC_ = false; CALL; } } break; \
        case 5: { constexpr int NK_ = 5; if (HASCAM) { constexpr bool HC_ = true; CALL; } else { constexpr bool HC_ = false; CALL; } } break; \
        case 6: { constexpr int NK_ = 6; if (HASCAM) { constexpr bool HC_ = true; CALL; } else { constexpr bool HC_ = false; CALL; } } break; \
        case 7: { constexpr int NK_ = 7; if (HASCAM) { constexpr bool HC_ = true; CALL; } else { constexpr bool HC_ = false; CALL; } } break; \
        case 8: { constexpr int NK_ = 8; if (HASCAM) { constexpr bool HC_ = true; CALL; } else { constexpr bool HC_ = false; CALL; } } break; \
        default: return cudaErrorInvalidValue;                                \
    }

// lanes per object point: 16 when points have few observations on average (two points per warp)
static int group_lanes(const DevProblem& P) { return (P.n_seg > 0 && P.n_obs <= 12 * (int64_t)P.n_seg) ? 16 : 32; }

static int point_pass_grid(const DevProblem& P, int sm_count) {
    const int per_cta = 128 / group_lanes(P);
    int grid = (P.n_seg + per_cta - 1) / per_cta;
    const int cap = sm_count * 4;
    if (grid > cap) grid = cap;
    return grid < 1 ? 1 : grid;
}

// number of point groups in the grid = rows of the camera partial buffer
int assemble_warps(const DevProblem& P, int sm_count) { return point_pass_grid(P, sm_count) * (128 / group_lanes(P)); }
// rows of the camera partial buffer: one per point group of the point pass + one per CTA of k_cam_direct
int cam_part_rows(const DevProblem& P, int sm_count) { return assemble_warps(P, sm_count) + sm_count; }

template <int NK>
static cudaError_t launch_cam_direct_t(const DevProblem& P, const int* opt, int sm_count, cudaStream_t st) {
    k_cam_direct<NK><<<sm_count, 256, 0, st>>>(P, opt, assemble_warps(P, sm_count));
    return cudaGetLastError();
}
static cudaError_t launch_cam_direct(const DevProblem& P, const int* opt, int sm_count, cudaStream_t st) {
    if (P.cam_rec) {
        const int row0 = assemble_warps(P, sm_count);
        switch (P.NK) {
            case 1: k_cam_rec<1><<<sm_count, 256, 0, st>>>(P, row0); break;
            case 2: k_cam_rec<2><<<sm_count, 256, 0, st>>>(P, row0); break;
            case 3: k_cam_rec<3><<<sm_count, 256, 0, st>>>(P, row0); break;
            case 4: k_cam_rec<4><<<sm_count, 256, 0, st>>>(P, row0); break;
            case 5: k_cam_rec<5><<<sm_count, 256, 0, st>>>(P, row0); break;
            case 6: k_cam_rec<6><<<sm_count, 256, 0, st>>>(P, row0); break;
            case 7: k_cam_rec<7><<<sm_count, 256, 0, st>>>(P, row0); break;
            case 8: k_cam_rec<8><<<sm_count, 256, 0, st>>>(P, row0); break;
            default: return cudaErrorInvalidValue;
        }
        return cudaGetLastError();
    }
    switch (P.NK) {
        case 1: return launch_cam_direct_t<1>(P, opt, sm_count, st);
        case 2: return launch_cam_direct_t<2>(P, opt, sm_count, st);
        case 3: return launch_cam_direct_t<3>(P, opt, sm_count, st);
        case 4: return launch_cam_direct_t<4>(P, opt, sm_count, st);
        case 5: return launch_cam_direct_t<5>(P, opt, sm_count, st);
        case 6: return launch_cam_direct_t<6>(P, opt, sm_count, st);
        case 7: return launch_cam_direct_t<7>(P, opt, sm_count, st);
        case 8: return launch_cam_direct_t<8>(P, opt, sm_count, st);
        default: return cudaErrorInvalidValue;
    }
}

template <int NK, bool HC, int G>
static cudaError_t launch_point_pass_t(const DevProblem& P, int sm_count, cudaStream_t st) {
    const size_t smem = (128 / G) * sizeof(PtSmem<NK, G>);
    static SmemOptIn opt;
    cudaError_t e0 = opt.ensure(k_point_pass<NK, HC, G>, smem);
    if (e0 != cudaSuccess) return e0;
    k_point_pass<NK, HC, G><<<point_pass_grid(P, sm_count), 128, smem, st>>>(P);
    return cudaGetLastError();
}

template <int NK, int G>
static cudaError_t launch_point_pass_mc_t(const DevProblem& P, int sm_count, int* info, cudaStream_t st) {
    const size_t smem = (128 / G) * sizeof(PtSmemMC<NK, G>);
    static SmemOptIn opt;
    cudaError_t e0 = opt.ensure(k_point_pass_mc<NK, G>, smem);
    if (e0 != cudaSuccess) return e0;
    k_point_pass_mc<NK, G><<<point_pass_grid(P, sm_count), 128, smem, st>>>(P, info);
    return cudaGetLastError();
}

cudaError_t launch_assemble(const DevProblem& P, int sm_count, int* info, cudaStream_t st, int64_t* launches,
                            const int* opt, const ChunkDev* chunks) {
    const bool hc = P.uc > 0;
    const bool mc = hc && P.n_cam > 1;
    if (chunks && !mc && P.n_seg > 0) {
        // chunk form: point pass (records at the observations' own positions), one CTA per chunk for the image
        // and image-pair partials, fixed-order final sums
        cudaError_t e = cudaSuccess;
        if (group_lanes(P) == 16) {
            FEBA_NK_DISPATCH2(P.NK, hc, (e = launch_point_pass_t<NK_, HC_, 16>(P, sm_count, st)));
        } else {
            FEBA_NK_DISPATCH2(P.NK, hc, (e = launch_point_pass_t<NK_, HC_, 32>(P, sm_count, st)));
        }
        if (e != cudaSuccess) return e;
        int grid = chunks->n_chunks < sm_count * 2 ? chunks->n_chunks : sm_count * 2;
        {
            static SmemOptIn opt[2 * 9];
            FEBA_NK_DISPATCH2(P.NK, hc, (e = opt[2 * NK_ + (HC_ ? 1 : 0)].ensure(k_chunk_reduce<NK_, HC_>, kChunkSmem)));
            if (e != cudaSuccess) return e;
        }
        FEBA_NK_DISPATCH2(P.NK, hc, (k_chunk_reduce<NK_, HC_><<<grid, 256, kChunkSmem, st>>>(P, *chunks)));
        *launches += 2;
        if (hc) {
            e = launch_cam_direct(P, opt, sm_count, st);
            if (e != cudaSuccess) return e;
            k_cam_reduce<<<1, 1024, 0, st>>>(P, cam_part_rows(P, sm_count));
            *launches += 2;
        }
        k_sum_img_parts<<<P.n_img < sm_count * 8 ? P.n_img : sm_count * 8, 128, 0, st>>>(P, *chunks);
        ++*launches;
        if (chunks->n_tblk > 0) {
            int g2 = (chunks->n_tblk + 3) / 4;
            if (g2 > sm_count * 16) g2 = sm_count * 16;
            k_sum_blk_parts<<<g2, 128, 0, st>>>(P, *chunks);
            ++*launches;
        }
        return cudaGetLastError();
    }
    if (P.n_seg > 0) {
        cudaError_t e = cudaSuccess;
        if (mc) {
            const bool g16 = group_lanes(P) == 16;
            switch (P.NK) {
#define FEBA_MC_CASE(N) \
    case N: e = g16 ? launch_point_pass_mc_t<N, 16>(P, sm_count, info, st) : launch_point_pass_mc_t<N, 32>(P, sm_count, info, st); break;
                FEBA_MC_CASE(1) FEBA_MC_CASE(2) FEBA_MC_CASE(3) FEBA_MC_CASE(4)
                FEBA_MC_CASE(5) FEBA_MC_CASE(6) FEBA_MC_CASE(7) FEBA_MC_CASE(8)
#undef FEBA_MC_CASE
                default: return cudaErrorInvalidValue;
            }
        } else if (group_lanes(P) == 16) {
            FEBA_NK_DISPATCH2(P.NK, hc, (e = launch_point_pass_t<NK_, HC_, 16>(P, sm_count, st)));
        } else {
            FEBA_NK_DISPATCH2(P.NK, hc, (e = launch_point_pass_t<NK_, HC_, 32>(P, sm_count, st)));
        }
        if (e != cudaSuccess) return e;
        ++*launches;
        int grid = P.n_img < sm_count * 4 ? P.n_img : sm_count * 4;
        FEBA_NK_DISPATCH2(P.NK, hc, (k_image_pass<NK_, HC_><<<grid, 128, 0, st>>>(P)));
        ++*launches;
        if (hc && !mc) {
            e = launch_cam_direct(P, opt, sm_count, st);
            if (e != cudaSuccess) return e;
            k_cam_reduce<<<1, 1024, 0, st>>>(P, cam_part_rows(P, sm_count));
            *launches += 2;
        }
        if (P.n_blocks > 0) {
            // resident CTAs per SM (measured 1, 2, 4, 8 on config 4: 13.4, 11.3, 12.0, 11.1 ms assembly).
            // Also tried and rejected: listing the blocks in a locality-preserving (Morton) order of the
            // camera positions for L2 reuse of the partner records -- no change (8.30 vs 8.22 ms): the
            // kernel is bound by L1 wavefronts of the record gathers (l1tex 71 % busy), not by DRAM.
            // Tried and rejected: a cooperative gather (nine lanes per 144-byte record, staged through
            // shared memory) -- 19 ms instead of 10.6 ms: the exposed load latency per 32-pair step is
            // not hidden with 12 warps per SM, while the lane-per-record version has all 36 loads of a
            // pair in flight at once.
            static int per_sm = -1;
            if (per_sm < 0) {
                const char* e = std::getenv("FEBA_PAIR_CTAS_PER_SM");
                per_sm = e ? std::atoi(e) : 8;
                if (per_sm < 1) per_sm = 1;
            }
            int g2 = (P.n_blocks + 3) / 4;
            if (g2 > sm_count * per_sm) g2 = sm_count * per_sm;
            k_pair_pass<<<g2, 128, 0, st>>>(P);
            ++*launches;
        }
    }
    return cudaGetLastError();
}

}  // namespace feba
