// Gauss-Newton hot-path kernels (sm_100a): parameter tables, fused BuildAwG + normal-equation
// blocks + Schur elimination of the object points, camera-parameter update, point
// back-substitution, residual / RSD reduction.
//
// Reference path replaced: functions/BuildAwG.m:46-528 (A, w, G), main.m:424-425 (u=A'Pw,
// N=A'PA), the point part of main.m:428-444, main.m:458-488 (un-scaling, update, sumabs),
// main.m:569 (v=A*delta+w), functions/BuildRSD.m:9-42, main.m:594-601.
// The design matrix A and the full normal matrix N are never formed.
#include "feba_dev.h"
#include "feba_kernels.h"
#include "feba_model.cuh"

namespace feba {

// ------------------------------------------------------------------------------------------
// K0: per-image and per-camera tables from the current parameters.
// M = R3(kappa) R2(phi) R1(omega) as written in BuildAwG.m:163-165.
__global__ void k_tables(int n_img, int n_cam, int NK, const double* __restrict__ eop,
                         const double* __restrict__ iop, const double* __restrict__ cam_box,
                         double* __restrict__ img_tab, double* __restrict__ cam_tab) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_img) image_table_row(eop + 6 * i, img_tab + kImgStride * i);
    if (i < n_cam) camera_table_row(NK, iop + (NK + 5) * i, cam_box + 5 * i, cam_tab + kCamStride * i);
}

// Inner-constraint rows per image from the CURRENT EOPs (BuildAwG.m:514-527), written as the
// augmented rows 1..7 of S (row r of the augmented block holds column r-1 of G).
__global__ void k_G_rows(DevProblem P, const double* __restrict__ eop) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_img) return;
    double G[6][7];
    inner_constraint_rows(eop + 6 * i, G);
    const size_t row0 = (size_t)P.img_row[i];              // inner constraints: all six EOPs are unknowns
    for (int q = 0; q < 6; ++q) {
        for (int c = 0; c < 7; ++c) {
            P.S[(size_t)(P.n_pad + 1 + c) + (size_t)P.ld * (row0 + q)] = G[q][c];
            P.Gt[8 * (row0 + q) + c] = G[q][c];
        }
        P.Gt[8 * (row0 + q) + 7] = 0.0;
    }
}

// Conditioning of the inner-constraint border.  The bordered solution (main.m:428-437) depends only
// on the column SPACE of G (G' delta = 0), so G may be replaced by G C for any non-singular 7x7 C.
// As written (BuildAwG.m:516-525) G mixes unit entries with coordinates in millimetres:
// G G' reaches 4e7 on rows whose normal-matrix diagonal is 4e1, and M = S + G G' then loses the
// digits of S that matter (measured: step error 9e-6 instead of 1e-7 on the bundled data).
// C is chosen so that, in the Jacobi-scaled system, the datum directions get eigenvalue ~1:
//   C C' = (G'G)^-1 (G' diag(S) G) (G'G)^-1,   C = diag(1/c) A1^-1 chol(A2)
// with c the column norms of G, A1 = Gn'Gn, A2 = Gn' diag(S) Gn, Gn = G diag(1/c).
// Three small launches (the single-CTA version took 0.26 ms on BASELINE configs[3] -- on the critical path of
// every iteration at every number of GPUs): (1) Gram sums of the image rows, one partial per CTA, (2) fixed-order
// sum of the partials + the 7x7 algebra on one thread -> C, (3) rows rewritten in place (compact copy Gt and the
// augmented rows 1..7 of S).
constexpr int kGramCtas = 64;

__global__ void __launch_bounds__(256) k_G_gram(DevProblem P, const double* __restrict__ dg, double* __restrict__ part) {
    __shared__ double red[8][56];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double a[56];                                       // a1 (28, lower of G'G) then a2 (28, lower of G' diag(S) G)
#pragma unroll
    for (int e = 0; e < 56; ++e) a[e] = 0.0;
    for (int r = blockIdx.x * 256 + tid; r < P.n_pad; r += gridDim.x * 256) {
        if (!is_image_row(P, r)) continue;
        double g[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) g[k] = P.Gt[8 * (size_t)r + k];
        const double sd = dg[r];                        // diag(S), summed over the ranks of a group
        int e = 0;
#pragma unroll
        for (int i = 0; i < 7; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const double p = g[i] * g[j];
                a[e] += p;
                a[28 + e] += p * sd;
                ++e;
            }
    }
#pragma unroll
    for (int e = 0; e < 56; ++e) {
#pragma unroll
        for (int s2 = 16; s2 > 0; s2 >>= 1) a[e] += __shfl_xor_sync(0xffffffffu, a[e], s2);
        if (lane == 0) red[warp][e] = a[e];
    }
    __syncthreads();
    if (tid < 56) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][tid];
        part[56 * blockIdx.x + tid] = t;
    }
}

// Cm (7x7, row-major, 49 doubles) from the Gram partials
__global__ void k_G_matrix(const double* __restrict__ part, int n_part, double* __restrict__ Cout) {
    __shared__ double tot[56];
    if (threadIdx.x < 56) {
        double t = 0.0;
        for (int c = 0; c < n_part; ++c) t += part[56 * c + threadIdx.x];
        tot[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    double A1[7][7], A2[7][7], Cm[7][7], cn[7];
    {
        int e = 0;
        for (int i = 0; i < 7; ++i)
            for (int j = 0; j <= i; ++j, ++e) {
                A1[i][j] = A1[j][i] = tot[e];
                A2[i][j] = A2[j][i] = tot[28 + e];
            }
    }
    for (int k = 0; k < 7; ++k) cn[k] = A1[k][k] > 0.0 ? 1.0 / sqrt(A1[k][k]) : 1.0;
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) { A1[i][j] *= cn[i] * cn[j]; A2[i][j] *= cn[i] * cn[j]; }
    // Lw = chol(A2) (lower); on failure fall back to sqrt(mean diagonal) * I
    double Lw[7][7];
    bool ok = true;
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) Lw[i][j] = 0.0;
    for (int j = 0; j < 7 && ok; ++j) {
        double d = A2[j][j];
        for (int k = 0; k < j; ++k) d -= Lw[j][k] * Lw[j][k];
        if (!(d > 1e-14 * A2[j][j])) { ok = false; break; }
        Lw[j][j] = sqrt(d);
        for (int i = j + 1; i < 7; ++i) {
            double v = A2[i][j];
            for (int k = 0; k < j; ++k) v -= Lw[i][k] * Lw[j][k];
            Lw[i][j] = v / Lw[j][j];
        }
    }
    // X = A1^-1 Lw by Gaussian elimination with partial pivoting
    double Q[7][14];
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) { Q[i][j] = A1[i][j]; Q[i][7 + j] = Lw[i][j]; }
    for (int c = 0; c < 7 && ok; ++c) {
        int p = c;
        for (int r = c + 1; r < 7; ++r) if (fabs(Q[r][c]) > fabs(Q[p][c])) p = r;
        if (!(fabs(Q[p][c]) > 1e-13)) { ok = false; break; }
        if (p != c) for (int j = 0; j < 14; ++j) { const double t = Q[c][j]; Q[c][j] = Q[p][j]; Q[p][j] = t; }
        for (int r = 0; r < 7; ++r) {
            if (r == c) continue;
            const double f = Q[r][c] / Q[c][c];
            for (int j = c; j < 14; ++j) Q[r][j] -= f * Q[c][j];
        }
    }
    if (ok) {
        for (int i = 0; i < 7; ++i)
            for (int j = 0; j < 7; ++j) Cm[i][j] = cn[i] * Q[i][7 + j] / Q[i][i];
    } else {
        double tr = 0.0;
        for (int i = 0; i < 7; ++i) tr += A2[i][i];
        const double sc = tr > 0.0 ? sqrt(tr / 7.0) : 1.0;
        for (int i = 0; i < 7; ++i)
            for (int j = 0; j < 7; ++j) Cm[i][j] = (i == j) ? cn[i] * sc : 0.0;
    }
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) Cout[7 * i + j] = Cm[i][j];
}

__global__ void __launch_bounds__(256) k_G_apply(DevProblem P, const double* __restrict__ Cin) {
    __shared__ double Cm[49];
    if (threadIdx.x < 49) Cm[threadIdx.x] = Cin[threadIdx.x];
    __syncthreads();
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= P.n_pad || !is_image_row(P, r)) return;
    double g[7], o[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) g[k] = P.Gt[8 * (size_t)r + k];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 7; ++k) acc += g[k] * Cm[7 * k + j];
        o[j] = acc;
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        P.Gt[8 * (size_t)r + j] = o[j];
        P.S[(size_t)(P.n_pad + 1 + j) + (size_t)P.ld * r] = o[j];
    }
}

// Sparse-datum form (feba_sparse.h): after the conditioning G -> G C, the rows of the datum images are
// copied into the augmented rows 8..14 (E = G~ restricted to those images; the rows of all other images
// stay zero there) and the compact copy Gt keeps the datum rows only, so that k_diag_scale and
// k_border_scale form M_s = S + E E' instead of the dense S + G~ G~'.  Augmented rows 1..7 keep all of G~.
__global__ void k_datum_split(DevProblem P) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.n_pad || !is_image_row(P, r)) return;
    const bool is_datum = P.datum[P.row_ext[r] / 6] != 0;
    double* g = P.Gt + 8 * (size_t)r;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const double v = is_datum ? g[k] : 0.0;
        P.S[(size_t)(P.n_pad + 8 + k) + (size_t)P.ld * r] = v;
        g[k] = v;
    }
}

// Unit diagonal on the padding rows so the padded matrix stays positive definite.
__global__ void k_pad_diag(DevProblem P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.n_pad && P.row_ext[i] < 0 && row_init_here(P, i)) P.S[(size_t)i + (size_t)P.ld * i] = 1.0;
}

// Jacobi scaling of the system that is factorised:  M = S + Gc Gc' (Gc is non-zero
// only in EOP rows, so the border stays in the camera block), d_i = 1/sqrt(M_ii).  The unknowns mix
// millimetres, radians, pixels and scaled distortion terms (diag(M) spans 1e0..1e9 on the bundled
// data, cond(M) ~ 1e11): Cholesky itself is insensitive to this scaling, the inverted 64x64
// diagonal factors used by the DMMA triangular solves are not, so M is equilibrated first.
__global__ void k_diag_extract(DevProblem P, double* __restrict__ dg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.n_pad) dg[i] = P.S[(size_t)i + (size_t)P.ld * i];
}

__global__ void k_diag_scale(DevProblem P, const double* __restrict__ dg, double* __restrict__ dvec,
                             int* __restrict__ info) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_pad) return;
    double d = P.row_ext[i] < 0 ? 1.0 : dg[i];                  // padding rows: unit diagonal
    if (P.inner && is_image_row(P, i)) {
        const double* g = P.Gt + 8 * (size_t)i;
#pragma unroll
        for (int k = 0; k < 7; ++k) d += g[k] * g[k];
    }
    if (!(d > 0.0)) {
        atomicExch(info, 1);
        d = 1.0;
    }
    dvec[i] = rsqrt(d);
}

// One pass over the structurally non-zero 64x64 blocks of the lower triangle (block list of the plan; every
// lower block for a dense plan): M~_rc = (S_rc + G_r . G_c) d_r d_c, and the augmented rows (right-hand side
// g, columns of G / E) scaled by d_c.  One CTA of 32 x 8 threads per block, x over rows (coalesced).
__global__ void __launch_bounds__(256) k_border_scale(DevProblem P, const double* __restrict__ dvec,
                                                      const int2* __restrict__ blocks) {
    const int2 b = blocks[blockIdx.x];                    // x: block row, y: block column (x >= y)
    const bool aug = b.x * kBlk >= P.n_pad;
    for (int cc = threadIdx.y; cc < kBlk; cc += 8) {
        const int c = b.y * kBlk + cc;
        const double dc = dvec[c];
        const bool c_img = P.inner && !aug && is_image_row(P, c);
        for (int rr = threadIdx.x; rr < kBlk; rr += 32) {
            const int r = b.x * kBlk + rr;
            if (aug) {
                // row 0 (right-hand side) is a partial sum of this rank; rows 1.. (G~, E) are the same on every
                // rank of a group: kept by the rank that initialises column c, zero elsewhere
                if (rr < P.aug_rows)
                    P.S[(size_t)r + (size_t)P.ld * c] =
                        (rr == 0 || row_init_here(P, c)) ? P.S[(size_t)r + (size_t)P.ld * c] * dc : 0.0;
            } else if (r >= c) {
                double v = P.S[(size_t)r + (size_t)P.ld * c];
                if (c_img && is_image_row(P, r) && row_init_here(P, r)) {
                    const double4* gr = reinterpret_cast<const double4*>(P.Gt + 8 * (size_t)r);
                    const double4* gc = reinterpret_cast<const double4*>(P.Gt + 8 * (size_t)c);
                    const double4 r0 = gr[0], r1 = gr[1], c0 = gc[0], c1 = gc[1];
                    v += r0.x * c0.x + r0.y * c0.y + r0.z * c0.z + r0.w * c0.w + r1.x * c1.x + r1.y * c1.y + r1.z * c1.z;
                }
                P.S[(size_t)r + (size_t)P.ld * c] = v * dvec[r] * dc;
            }
        }
    }
}

// Clear the structurally non-zero blocks of S (the assembly adds into them, the factorisation fills them)
// and the augmented block row; everything else is never written and stays zero from the handle's creation.
__global__ void __launch_bounds__(256) k_clear_blocks(DevProblem P, const int2* __restrict__ blocks) {
    const int2 b = blocks[blockIdx.x];
    double* base = P.S + (size_t)b.x * kBlk + (size_t)P.ld * b.y * kBlk;
    for (int cc = threadIdx.y; cc < kBlk; cc += 8)
        for (int rr = threadIdx.x; rr < kBlk; rr += 32) base[rr + (size_t)P.ld * cc] = 0.0;
}

// ------------------------------------------------------------------------------------------
// xhat <-> parameter tables.  Layout of xhat: Buildxhat.m:22-135 (per image the estimated EOPs,
// per camera the estimated xp yp c k1..kNK p1 p2, per TIE entry X Y Z); the gather of
// BuildAwG.m:52-155 (estimated -> from xhat, else the file value) is done once per set, not per
// observation.  scatter: xhat -> tables (both coordinate buffers), gather: tables -> xhat.
__global__ void k_xhat_scatter(DevProblem P, const double* __restrict__ xhat, double* __restrict__ eop,
                               double* __restrict__ iop, const int* __restrict__ tie_pt) {
    const int64_t n_e = (int64_t)P.n_img * 6, n_c = (int64_t)P.n_cam * P.NC, n_t = (int64_t)P.n_tie * 3;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_e + n_c + n_t;
         i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n_e) {
            const int im = (int)(i / 6), q = (int)(i - 6 * (int64_t)im);
            if (P.ecol[q] >= 0) eop[i] = xhat[(int64_t)P.ui * im + P.ecol[q]];
        } else if (i < n_e + n_c) {
            const int64_t r = i - n_e;
            const int cam = (int)(r / P.NC), q = (int)(r - (int64_t)cam * P.NC);
            if (P.ccol[q] >= 0) iop[r] = xhat[(int64_t)P.ext_off_cam + (int64_t)P.uc * cam + P.ccol[q]];
        } else {
            const int64_t r = i - n_e - n_c;
            const int t = (int)(r / 3), k = (int)(r - 3 * (int64_t)t);
            const int pt = tie_pt[t];
            if (pt >= 0) {
                const double val = xhat[(int64_t)P.n_red + r];
                P.xyz[3 * (int64_t)pt + k] = val;
                P.xyz_prev[3 * (int64_t)pt + k] = val;
            }
        }
    }
}

__global__ void k_xhat_gather(DevProblem P, double* __restrict__ xhat, const double* __restrict__ eop,
                              const double* __restrict__ iop, const int* __restrict__ tie_pt) {
    const int64_t n_e = (int64_t)P.n_img * 6, n_c = (int64_t)P.n_cam * P.NC, n_t = (int64_t)P.n_tie * 3;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_e + n_c + n_t;
         i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n_e) {
            const int im = (int)(i / 6), q = (int)(i - 6 * (int64_t)im);
            if (P.ecol[q] >= 0) xhat[(int64_t)P.ui * im + P.ecol[q]] = eop[i];
        } else if (i < n_e + n_c) {
            const int64_t r = i - n_e;
            const int cam = (int)(r / P.NC), q = (int)(r - (int64_t)cam * P.NC);
            if (P.ccol[q] >= 0) xhat[(int64_t)P.ext_off_cam + (int64_t)P.uc * cam + P.ccol[q]] = iop[r];
        } else {
            const int64_t r = i - n_e - n_c;
            const int t = (int)(r / 3), k = (int)(r - 3 * (int64_t)t);
            const int pt = tie_pt[t];
            if (pt >= 0) xhat[(int64_t)P.n_red + r] = P.xyz[3 * (int64_t)pt + k];
        }
    }
}

// delta in xhat layout (un-scaled, main.m:458-482): camera part from dcam_unscaled, ties from dpts.
__global__ void k_delta_gather(DevProblem P, double* __restrict__ delta) {
    const int64_t n = (int64_t)P.n_red + 3 * (int64_t)P.n_tie;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        delta[i] = i < P.n_red ? P.dcam_unscaled[row_of_ext(P, (int)i)] : P.dpts[i - P.n_red];
}

// ------------------------------------------------------------------------------------------
// 3x3 symmetric inverse through a Cholesky factor.  v = [v00 v10 v11 v20 v21 v22].
__device__ __forceinline__ void sym3_inverse(const double* v, double* inv) {
    const double l00 = sqrt(v[0]);
    const double l10 = v[1] / l00;
    const double l20 = v[3] / l00;
    const double l11 = sqrt(v[2] - l10 * l10);
    const double l21 = (v[4] - l20 * l10) / l11;
    const double l22 = sqrt(v[5] - l20 * l20 - l21 * l21);
    // inverse of L (lower)
    const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
    const double i10 = -l10 * i00 * i11;
    const double i21 = -l21 * i11 * i22;
    const double i20 = -(l20 * i00 + l21 * i10) * i22;
    // V^-1 = L^-T L^-1
    inv[0] = i00 * i00 + i10 * i10 + i20 * i20;
    inv[1] = i10 * i11 + i20 * i21;
    inv[2] = i11 * i11 + i21 * i21;
    inv[3] = i20 * i22;
    inv[4] = i21 * i22;
    inv[5] = i22 * i22;
}

__device__ __forceinline__ double sym3(const double* m, int i, int j) {
    // element (i,j) of a packed symmetric 3x3 [00 10 11 20 21 22]
    const int a = i > j ? i : j, b = i > j ? j : i;
    return m[a * (a + 1) / 2 + b];
}

// ------------------------------------------------------------------------------------------
// Camera-part update (main.m:458-488 for the EOP/IOP unknowns): delta_c = -sol, un-scale the
// distortion increments by r_max^(2j) / r_max^2, add to the parameter tables, partial sumabs.
__global__ void k_update_cam(DevProblem P, const double* __restrict__ sol, const double* __restrict__ dvec,
                             double* __restrict__ dcam,
                             double* __restrict__ dcam_unscaled, double* __restrict__ eop,
                             double* __restrict__ iop, double* __restrict__ out_sumabs) {
    __shared__ double red[1024];
    double acc = 0.0;
    for (int i = threadIdx.x; i < P.n_pad; i += blockDim.x) {
        double d = 0.0, du = 0.0;
        const int e = P.row_ext[i];                // entry of xhat this row belongs to, -1: padding
        if (e >= 0) {
            d = -sol[i] * dvec[i];                 // undo the Jacobi scaling of the reduced system
            du = d;
            if (e < P.ext_off_cam) {
                const int im = e / P.ui, slot = e - im * P.ui;
                int q = 0;
                for (int k = 0; k < 6; ++k) if (P.ecol[k] == slot) q = k;
                eop[6 * im + q] += du;
            } else {
                const int r = e - P.ext_off_cam;
                const int cam = r / P.uc, slot = r - cam * P.uc;
                int q = 0;
                for (int k = 0; k < P.NC; ++k) if (P.ccol[k] == slot) q = k;
                const double* ct = P.cam_tab + kCamStride * cam;
                if (q >= 3 && q < 3 + P.NK) du = d / ct[24 + (q - 3)];      // main.m:467
                else if (q >= 3 + P.NK) du = d / ct[24];                     // main.m:476,479
                iop[P.NC * cam + q] += du;
            }
            acc += fabs(du);
        }
        dcam[i] = d;
        dcam_unscaled[i] = du;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out_sumabs[0] = red[0];
}

// ------------------------------------------------------------------------------------------
// K4: back-substitution of the tie points, d_p = -V_p^-1 (u_p + W_p' d_c), X += d_p, sumabs.
// Recomputes the Jacobians at the linearisation point (W_p is never stored):
//   W_p' d_c = sum_a Jt_a' P (Je_a d_e(i_a) + Jc_a d_cam).
// One group of G lanes per point (G = 16: two points per warp when points have few observations).
template <int NK, bool HAS_CAM, int G>
__global__ void __launch_bounds__(128) k_backsub(DevProblem P) {
    constexpr int NC = NK + 5;
    const int lane = threadIdx.x & (G - 1);
    const int nwarp = gridDim.x * (blockDim.x / G);
    const int gw = blockIdx.x * (blockDim.x / G) + threadIdx.x / G;
    const unsigned gmask = G == 32 ? 0xffffffffu : (0xffffu << (16 * ((threadIdx.x & 31) >> 4)));
    const double pw[2] = {P.px, P.py};
    double dsum = 0.0;
    for (int seg = gw; seg < P.n_seg; seg += nwarp) {
        const int pt = P.seg_pt[seg];
        const int tie = P.pt_tie[pt];
        if (tie < 0) continue;
        const int beg = P.seg_start[seg], end = P.seg_start[seg + 1];
        const double X = P.xyz[3 * pt], Y = P.xyz[3 * pt + 1], Z = P.xyz[3 * pt + 2];
        double acc[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) acc[k] = 0.0;
        for (int o = beg + lane; o < end; o += G) {
            ObsJac<NK> J;
            const int img = P.oimg[o];
            const int cam = P.img_cam[img];
            observation<NK, HAS_CAM>(P.type, P.ox[o], P.oy[o], P.img_tab + kImgStride * img,
                                     P.cam_tab + kCamStride * cam, X, Y, Z, J);
            double q[2] = {0.0, 0.0};
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                if (P.ecol[i] >= 0) {
                    const double d = P.dcam[P.img_row[img] + P.ecol[i]];
                    q[0] += J.Je[0][i] * d;
                    q[1] += J.Je[1][i] * d;
                }
            }
            if (HAS_CAM) {
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    if (P.ccol[j] >= 0) {
                        const double d = P.dcam[P.off_cam + P.uc * cam + P.ccol[j]];
                        q[0] += J.Jc[0][j] * d;
                        q[1] += J.Jc[1][j] * d;
                    }
                }
            }
            int e = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int k = 0; k <= i; ++k)
                    acc[e++] += J.Jt[0][i] * pw[0] * J.Jt[0][k] + J.Jt[1][i] * pw[1] * J.Jt[1][k];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                acc[6 + k] += J.Jt[0][k] * pw[0] * J.w[0] + J.Jt[1][k] * pw[1] * J.w[1];
                acc[9 + k] += J.Jt[0][k] * pw[0] * q[0] + J.Jt[1][k] * pw[1] * q[1];
            }
        }
#pragma unroll
        for (int k = 0; k < 12; ++k)
#pragma unroll
            for (int s = G / 2; s > 0; s >>= 1) acc[k] += __shfl_xor_sync(gmask, acc[k], s);
        double Vinv[6];
        sym3_inverse(acc, Vinv);
        double t[3] = {acc[6] + acc[9], acc[7] + acc[10], acc[8] + acc[11]};
        double d[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            d[k] = -(sym3(Vinv, k, 0) * t[0] + sym3(Vinv, k, 1) * t[1] + sym3(Vinv, k, 2) * t[2]);
        if (lane < 3) {
            const double dv = lane == 0 ? d[0] : (lane == 1 ? d[1] : d[2]);
            const double xv = lane == 0 ? X : (lane == 1 ? Y : Z);
            // each point belongs to exactly one warp and was read above: in-place update is safe
            P.xyz_prev[3 * pt + lane] = xv;
            P.xyz[3 * pt + lane] = xv + dv;                      // main.m:484
            P.dpts[3 * tie + lane] = dv;
        }
        dsum += fabs(d[0]) + fabs(d[1]) + fabs(d[2]);            // main.m:487 (sumabs)
    }
    if (lane == 0) P.partial[gw] = dsum;
}

// K4 from the records of the point pass of the same iteration (single camera or no camera unknowns): with
// V = L L', Z_a = P Jt_a L^-T (rec1), ut = L^-1 u_p and Fc = Wc L^-T (pt_rec),
//   d_p = -L^-T ( ut + Fc' d_cam + sum_a Z_a' Je_a d_e(i_a) ):
// no Jacobian is evaluated again; one 144-byte record gather per observation.
template <int NK, bool HAS_CAM, int G>
__global__ void __launch_bounds__(128) k_backsub_rec(DevProblem P) {
    constexpr int NC = NK + 5;
    constexpr int PR = 9 + 3 * NC;
    const int lane = threadIdx.x & (G - 1);
    const int nwarp = gridDim.x * (blockDim.x / G);
    const int gw = blockIdx.x * (blockDim.x / G) + threadIdx.x / G;
    const unsigned gmask = G == 32 ? 0xffffffffu : (0xffffu << (16 * ((threadIdx.x & 31) >> 4)));
    double dsum = 0.0;
    for (int seg = gw; seg < P.n_seg; seg += nwarp) {
        const int pt = P.seg_pt[seg];
        const int tie = P.pt_tie[pt];
        if (tie < 0) continue;
        const int beg = P.seg_start[seg], end = P.seg_start[seg + 1];
        double s[3] = {0.0, 0.0, 0.0};
        for (int o = beg + lane; o < end; o += G) {
            const size_t pos = P.ipos ? (size_t)P.ipos[o] : (size_t)o;
            const int row = P.img_row[P.oimg[o]];
            const double2* q2 = reinterpret_cast<const double2*>(P.rec1 + (size_t)kRec1 * pos);
            double r[kRec1];
#pragma unroll
            for (int k = 0; k < kRec1 / 2; ++k) { const double2 v = q2[k]; r[2 * k] = v.x; r[2 * k + 1] = v.y; }
            double q[2] = {0.0, 0.0};
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                if (P.ecol[i] >= 0) {
                    const double d = P.dcam[row + P.ecol[i]];
                    q[0] += r[i] * d;
                    q[1] += r[6 + i] * d;
                }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) s[k] += r[12 + k] * q[0] + r[15 + k] * q[1];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int w = G / 2; w > 0; w >>= 1) s[k] += __shfl_xor_sync(gmask, s[k], w);
        const double* pr = P.pt_rec + (size_t)PR * tie;
        if (HAS_CAM) {
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                if (P.ccol[j] >= 0) {
                    const double d = P.dcam[P.off_cam + P.ccol[j]];          // single camera
#pragma unroll
                    for (int k = 0; k < 3; ++k) s[k] += pr[9 + 3 * j + k] * d;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) s[k] += pr[6 + k];
        const double i00 = pr[0], i10 = pr[1], i11 = pr[2], i20 = pr[3], i21 = pr[4], i22 = pr[5];
        const double d[3] = {-(i00 * s[0] + i10 * s[1] + i20 * s[2]), -(i11 * s[1] + i21 * s[2]), -(i22 * s[2])};
        if (lane < 3) {
            const double dv = lane == 0 ? d[0] : (lane == 1 ? d[1] : d[2]);
            const double xv = P.xyz[3 * pt + lane];
            P.xyz_prev[3 * pt + lane] = xv;
            P.xyz[3 * pt + lane] = xv + dv;                      // main.m:484
            P.dpts[3 * tie + lane] = dv;
        }
        dsum += fabs(d[0]) + fabs(d[1]) + fabs(d[2]);            // main.m:487 (sumabs)
    }
    if (lane == 0) P.partial[gw] = dsum;
}

// Fixed-order sum of per-warp / per-block partials (deterministic).
__global__ void k_sum_partials(const double* __restrict__ partial, int n, int stride, int offset,
                               double* __restrict__ out) {
    __shared__ double red[1024];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[(size_t)i * stride + offset];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

// ------------------------------------------------------------------------------------------
// K5: residuals v = A*delta + w with A, w of the LAST iteration (linearisation point before the
// last update) and the UN-scaled delta against the SCALED distortion columns (main.m:569 after
// main.m:458-482 -- reproduced as coded), BuildRSD columns r vx vy vr vt with xp, yp from the
// post-update parameters (BuildRSD.m:14-27), and sum vx^2, sum vy^2 for main.m:594-601.
template <int NK, bool HAS_CAM>
__global__ void __launch_bounds__(256) k_residuals(DevProblem P, const int* __restrict__ opt,
                                                   const double* __restrict__ xyz_prev,
                                                   const double* __restrict__ iop_new,
                                                   double* __restrict__ v_out, double* __restrict__ rsd_out) {
    __shared__ double red[2][256];
    double sx = 0.0, sy = 0.0;
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < P.n_obs;
         o += (int64_t)gridDim.x * blockDim.x) {
        const int img = P.oimg[o], pt = opt[o];
        const int tie = P.pt_tie[pt];
        ObsJac<NK> J;
        const double x = P.ox[o], y = P.oy[o];
        const int cam = P.img_cam[img];
        observation<NK, HAS_CAM>(P.type, x, y, P.img_tab + kImgStride * img, P.cam_tab + kCamStride * cam,
                                 xyz_prev[3 * pt], xyz_prev[3 * pt + 1], xyz_prev[3 * pt + 2], J);
        double v[2];
        residual_of<NK, HAS_CAM>(J, P.ecol, P.ccol, P.dcam_unscaled + P.img_row[img],
                                 P.dcam_unscaled + P.off_cam + P.uc * cam, tie >= 0 ? P.dpts + 3 * tie : nullptr, v);
        sx += v[0] * v[0];
        sy += v[1] * v[1];
        const int row = P.operm[o];
        if (v_out) {
            v_out[2 * (size_t)row] = v[0];
            v_out[2 * (size_t)row + 1] = v[1];
        }
        if (rsd_out)
            rsd_row(x, y, iop_new[P.NC * cam], iop_new[P.NC * cam + 1], v, rsd_out + 5 * (size_t)row);
    }
    red[0][threadIdx.x] = sx;
    red[1][threadIdx.x] = sy;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            red[0][threadIdx.x] += red[0][threadIdx.x + s];
            red[1][threadIdx.x] += red[1][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        P.partial[2 * blockIdx.x] = red[0][0];
        P.partial[2 * blockIdx.x + 1] = red[1][0];
    }
}

// ------------------------------------------------------------------------------------------
// Covariance outputs of the EOP/IOP part.  Qxx_cc(i,j) = d_i d_j (Q_ij - Y_i T7inv Y_j')
// (Q = M~^-1 lower-stored, Y = M~^-1 G~; without inner constraints the second term is absent).
__device__ __forceinline__ double cov_entry(const DevProblem& P, const double* __restrict__ Q,
                                            const double* __restrict__ Y, const double* __restrict__ T7inv,
                                            const double* __restrict__ dvec, int i, int j) {
    const int a = i > j ? i : j, b = i > j ? j : i;
    double q = Q[(size_t)a + (size_t)P.n_pad * b];
    if (P.inner) {
        const double* yi = Y + 8 * (size_t)i;
        const double* yj = Y + 8 * (size_t)j;
        double corr = 0.0;
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            double t = 0.0;
#pragma unroll
            for (int c = 0; c < 7; ++c) t += T7inv[7 * r + c] * yj[c];
            corr += yi[r] * t;
        }
        q -= corr;
    }
    return q * dvec[i] * dvec[j];
}

// diag(Cx)/sigma02 for the EOP/IOP unknowns, distortion entries un-scaled: /r_max^(4j), /r_max^4
// (main.m:468,477,480 divide Cx(k,k) by dist_scaling^2).
__global__ void k_cov_diag_cam(DevProblem P, const double* __restrict__ Q, const double* __restrict__ Y,
                               const double* __restrict__ T7inv, const double* __restrict__ dvec,
                               double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // entry of the EOP/IOP part of xhat
    if (i >= P.n_red) return;
    const int row = row_of_ext(P, i);
    double q = cov_entry(P, Q, Y, T7inv, dvec, row, row);
    if (i >= P.ext_off_cam) {
        const int r = i - P.ext_off_cam;
        const int cam = r / P.uc, slot = r - cam * P.uc;
        int p = 0;
        for (int k = 0; k < P.NC; ++k) if (P.ccol[k] == slot) p = k;
        const double* ct = P.cam_tab + kCamStride * cam;
        if (p >= 3 && p < 3 + P.NK) q /= ct[24 + (p - 3)] * ct[24 + (p - 3)];
        else if (p >= 3 + P.NK) q /= ct[24] * ct[24];
    }
    out[i] = q;
}

// k x k block of Cx/sigma02 BEFORE un-scaling (what main.m:446-456 turns into Correlation): idx are
// unknown indices of the EOP/IOP part.
__global__ void k_cov_block(DevProblem P, const double* __restrict__ Q, const double* __restrict__ Y,
                            const double* __restrict__ T7inv, const double* __restrict__ dvec,
                            const long long* __restrict__ idx, int k, double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= k * k) return;
    const int a = e / k, b = e - a * k;
    out[e] = cov_entry(P, Q, Y, T7inv, dvec, row_of_ext(P, (int)idx[a]), row_of_ext(P, (int)idx[b]));
}

// Tie-point variances: Q_pp = V^-1 + V^-1 (W_p' Qxx_cc W_p) V^-1 per tie point, from the
// cofactors of the EOP/IOP part.  One warp per point, lanes over its observations.  W_p has one 6x3
// block We_a = Je_a'PJt_a per observation (rows of image i_a) and one NC x 3 block per camera seeing
// the point (Wc_k = sum_{a in k} Jc_a'PJt_a).  With w'_e = d_e w_e (d = Jacobi scaling) the bilinear
// form splits as  R = sum_{e,f} w'_e Q(e,f) w'_f'  -  E' T7inv E,  E = sum_e Y_e' w'_e  (7 x 3).
// Observations beyond one chunk of 32 are handled by re-evaluating the partner chunk (as the
// Jacobians are never stored).  Output: diag(Q_pp) (3 per tie point), cofactors (x sigma02 -> variances).
template <int NK, bool HAS_CAM>
__global__ void __launch_bounds__(128) k_cov_points(DevProblem P, const double* __restrict__ Q,
                                                    const double* __restrict__ Y, const double* __restrict__ T7inv,
                                                    const double* __restrict__ dvec, double* __restrict__ out) {
    constexpr int NC = NK + 5;
    constexpr int KC = 4;                                   // cameras per point (as in the point pass)
    __shared__ double sB[4][32][18];                        // partner chunk: w' rows of 32 observations
    __shared__ int sImg[4][32];
    __shared__ double sWc[4][KC][NC][3];
    __shared__ int sCam[4][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nwarp = gridDim.x * (blockDim.x >> 5);
    const int gw = blockIdx.x * (blockDim.x >> 5) + wib;
    const double pw[2] = {P.px, P.py};
    const int np = P.n_pad;
    auto Qs = [&](int i, int j) { return i >= j ? Q[(size_t)i + (size_t)np * j] : Q[(size_t)j + (size_t)np * i]; };

    for (int seg = gw; seg < P.n_seg; seg += nwarp) {
        const int pt = P.seg_pt[seg];
        const int tie = P.pt_tie[pt];
        if (tie < 0) continue;
        const int beg = P.seg_start[seg], end = P.seg_start[seg + 1];
        // the linearisation point of the last iteration (P.xyz has already been updated)
        const double X = P.xyz_prev[3 * pt], Yc_ = P.xyz_prev[3 * pt + 1], Z = P.xyz_prev[3 * pt + 2];
        int cams[KC] = {-1, -1, -1, -1};
        int ncam = 0;
        for (int e = lane; e < KC * NC * 3; e += 32) (&sWc[wib][0][0][0])[e] = 0.0;
        __syncwarp();
        // ---- pass 1: V, per-camera Wc
        double V[6] = {0, 0, 0, 0, 0, 0};
        for (int c0 = beg; c0 < end; c0 += 32) {
            const int o = c0 + lane;
            const bool act = o < end;
            ObsJac<NK> J;
            int cam = -1;
            if (act) {
                const int img = P.oimg[o];
                cam = P.img_cam[img];
                observation<NK, HAS_CAM>(P.type, P.ox[o], P.oy[o], P.img_tab + kImgStride * img,
                                         P.cam_tab + kCamStride * cam, X, Yc_, Z, J);
                int e = 0;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int k = 0; k <= i; ++k)
                        V[e++] += J.Jt[0][i] * pw[0] * J.Jt[0][k] + J.Jt[1][i] * pw[1] * J.Jt[1][k];
            }
            sCam[wib][lane] = cam;
            __syncwarp();
            const int nrow = min(32, end - c0);
            for (int l = 0; l < nrow; ++l) {
                const int c = sCam[wib][l];
                bool found = false;
#pragma unroll
                for (int k = 0; k < KC; ++k) found = found || (k < ncam && cams[k] == c);
                if (!found && ncam < KC) {
#pragma unroll
                    for (int k = 0; k < KC; ++k)
                        if (k == ncam) cams[k] = c;
                    ++ncam;
                }
            }
            if (HAS_CAM && act) {
                int kown = 0;
#pragma unroll
                for (int k = 0; k < KC; ++k)
                    if (cams[k] == cam) kown = k;
#pragma unroll
                for (int j = 0; j < NC; ++j)
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        atomicAdd(&sWc[wib][kown][j][k], J.Jc[0][j] * pw[0] * J.Jt[0][k] + J.Jc[1][j] * pw[1] * J.Jt[1][k]);
            }
            __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < 6; ++k)
#pragma unroll
            for (int s_ = 16; s_ > 0; s_ >>= 1) V[k] += __shfl_xor_sync(0xffffffffu, V[k], s_);
        double Vi[6];
        sym3_inverse(V, Vi);
        // ---- pass 2: R (3x3, full) and E (7x3)
        double R[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        double E[7][3];
#pragma unroll
        for (int r = 0; r < 7; ++r) E[r][0] = E[r][1] = E[r][2] = 0.0;
        for (int a0 = beg; a0 < end; a0 += 32) {
            const int oa = a0 + lane;
            const bool acta = oa < end;
            double A_[6][3];
            int rowa = 0;
            if (acta) {
                const int img = P.oimg[oa];
                rowa = P.img_row[img];
                ObsJac<NK> J;
                observation<NK, false>(P.type, P.ox[oa], P.oy[oa], P.img_tab + kImgStride * img,
                                       P.cam_tab + kCamStride * P.img_cam[img], X, Yc_, Z, J);
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const double d = P.ecol[i] >= 0 ? dvec[rowa + P.ecol[i]] : 0.0;
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        A_[i][k] = d * (J.Je[0][i] * pw[0] * J.Jt[0][k] + J.Je[1][i] * pw[1] * J.Jt[1][k]);
                }
                if (P.inner) {
#pragma unroll
                    for (int i = 0; i < 6; ++i)
                        if (P.ecol[i] >= 0) {
                            const double* y = Y + 8 * (size_t)(rowa + P.ecol[i]);
#pragma unroll
                            for (int r = 0; r < 7; ++r)
#pragma unroll
                                for (int k = 0; k < 3; ++k) E[r][k] += y[r] * A_[i][k];
                        }
                }
                // image x camera: A' Q[img rows, cam rows] Wc' + transpose
                if (HAS_CAM) {
                    for (int kc = 0; kc < ncam; ++kc) {
                        const int rc = P.off_cam + P.uc * cams[kc];
                        for (int j = 0; j < NC; ++j) {
                            if (P.ccol[j] < 0) continue;
                            const double dj = dvec[rc + P.ccol[j]];
                            const double w0 = dj * sWc[wib][kc][j][0], w1 = dj * sWc[wib][kc][j][1], w2 = dj * sWc[wib][kc][j][2];
                            double t0 = 0, t1 = 0, t2 = 0;           // sum_i A'[i][:] Q(i, j)
#pragma unroll
                            for (int i = 0; i < 6; ++i)
                                if (P.ecol[i] >= 0) {
                                    const double q = Qs(rc + P.ccol[j], rowa + P.ecol[i]);
                                    t0 += A_[i][0] * q; t1 += A_[i][1] * q; t2 += A_[i][2] * q;
                                }
                            const double t[3] = {t0, t1, t2}, w[3] = {w0, w1, w2};
#pragma unroll
                            for (int x = 0; x < 3; ++x)
#pragma unroll
                                for (int y2 = 0; y2 < 3; ++y2) R[x][y2] += t[x] * w[y2] + w[x] * t[y2];
                        }
                    }
                }
            }
            // image x image over all partner chunks
            for (int b0 = beg; b0 < end; b0 += 32) {
                __syncwarp();
                const int ob = b0 + lane;
                sImg[wib][lane] = -1;
                if (b0 == a0) {
                    if (acta) {
                        sImg[wib][lane] = rowa;
#pragma unroll
                        for (int i = 0; i < 6; ++i)
#pragma unroll
                            for (int k = 0; k < 3; ++k) sB[wib][lane][3 * i + k] = A_[i][k];
                    }
                } else if (ob < end) {
                    const int img = P.oimg[ob];
                    const int rowb = P.img_row[img];
                    ObsJac<NK> Jb;
                    observation<NK, false>(P.type, P.ox[ob], P.oy[ob], P.img_tab + kImgStride * img,
                                           P.cam_tab + kCamStride * P.img_cam[img], X, Yc_, Z, Jb);
                    sImg[wib][lane] = rowb;
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const double d = P.ecol[i] >= 0 ? dvec[rowb + P.ecol[i]] : 0.0;
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            sB[wib][lane][3 * i + k] =
                                d * (Jb.Je[0][i] * pw[0] * Jb.Jt[0][k] + Jb.Je[1][i] * pw[1] * Jb.Jt[1][k]);
                    }
                }
                __syncwarp();
                if (acta) {
                    const int nb = min(32, end - b0);
                    for (int b = 0; b < nb; ++b) {
                        const int rowb = sImg[wib][b];
                        // t = A'' Q[rows a, rows b] (3 x 6), then R += t B'  (full double sum over a and b)
#pragma unroll
                        for (int jb = 0; jb < 6; ++jb) {
                            if (P.ecol[jb] < 0) continue;
                            double t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
                            for (int i = 0; i < 6; ++i)
                                if (P.ecol[i] >= 0) {
                                    const double q = Qs(rowa + P.ecol[i], rowb + P.ecol[jb]);
                                    t0 += A_[i][0] * q; t1 += A_[i][1] * q; t2 += A_[i][2] * q;
                                }
                            const double b0_ = sB[wib][b][3 * jb], b1_ = sB[wib][b][3 * jb + 1], b2_ = sB[wib][b][3 * jb + 2];
                            R[0][0] += t0 * b0_; R[0][1] += t0 * b1_; R[0][2] += t0 * b2_;
                            R[1][0] += t1 * b0_; R[1][1] += t1 * b1_; R[1][2] += t1 * b2_;
                            R[2][0] += t2 * b0_; R[2][1] += t2 * b1_; R[2][2] += t2 * b2_;
                        }
                    }
                }
            }
            __syncwarp();
        }
        // camera x camera and the camera rows of E: spread over the lanes
        if (HAS_CAM) {
            for (int e = lane; e < ncam * NC * ncam * NC; e += 32) {
                const int e1 = e / (ncam * NC), e2 = e - e1 * (ncam * NC);
                const int k1 = e1 / NC, j1 = e1 - k1 * NC, k2 = e2 / NC, j2 = e2 - k2 * NC;
                if (P.ccol[j1] < 0 || P.ccol[j2] < 0) continue;
                const int i1 = P.off_cam + P.uc * cams[k1] + P.ccol[j1], i2 = P.off_cam + P.uc * cams[k2] + P.ccol[j2];
                const double q = Qs(i1, i2) * dvec[i1] * dvec[i2];
#pragma unroll
                for (int x = 0; x < 3; ++x)
#pragma unroll
                    for (int y2 = 0; y2 < 3; ++y2) R[x][y2] += sWc[wib][k1][j1][x] * q * sWc[wib][k2][j2][y2];
            }
            if (P.inner)
                for (int e = lane; e < ncam * NC; e += 32) {
                    const int k1 = e / NC, j1 = e - k1 * NC;
                    if (P.ccol[j1] < 0) continue;
                    const int i1 = P.off_cam + P.uc * cams[k1] + P.ccol[j1];
                    const double* y = Y + 8 * (size_t)i1;
#pragma unroll
                    for (int r = 0; r < 7; ++r)
#pragma unroll
                        for (int k = 0; k < 3; ++k) E[r][k] += y[r] * dvec[i1] * sWc[wib][k1][j1][k];
                }
        }
#pragma unroll
        for (int x = 0; x < 3; ++x)
#pragma unroll
            for (int y2 = 0; y2 < 3; ++y2)
#pragma unroll
                for (int s_ = 16; s_ > 0; s_ >>= 1) R[x][y2] += __shfl_xor_sync(0xffffffffu, R[x][y2], s_);
        if (P.inner) {
#pragma unroll
            for (int r = 0; r < 7; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k)
#pragma unroll
                    for (int s_ = 16; s_ > 0; s_ >>= 1) E[r][k] += __shfl_xor_sync(0xffffffffu, E[r][k], s_);
            // R -= E' T7inv E
#pragma unroll
            for (int x = 0; x < 3; ++x)
#pragma unroll
                for (int y2 = 0; y2 < 3; ++y2) {
                    double acc = 0.0;
                    for (int r = 0; r < 7; ++r)
                        for (int c = 0; c < 7; ++c) acc += E[r][x] * T7inv[7 * r + c] * E[c][y2];
                    R[x][y2] -= acc;
                }
        }
        if (lane < 3) {
            // (V^-1 + V^-1 R V^-1)(lane, lane)
            double vrow[3] = {sym3(Vi, lane, 0), sym3(Vi, lane, 1), sym3(Vi, lane, 2)};
            double acc = vrow[lane == 0 ? 0 : (lane == 1 ? 1 : 2)];
#pragma unroll
            for (int x = 0; x < 3; ++x)
#pragma unroll
                for (int y2 = 0; y2 < 3; ++y2) acc += vrow[x] * R[x][y2] * vrow[y2];
            out[3 * (size_t)tie + lane] = acc;
        }
    }
}

// ------------------------------------------------------------------------------------------
// host launchers (dispatch on NK and on whether any camera parameter is estimated)

#define FEBA_NK_DISPATCH(NKV, HASCAM, CALL)                                  \
    switch (NKV) {                                                           \
        case 1: { constexpr int NK_ = 1; if (HASCAM) { constexpr bool HC_ = true; CALL; } else { constexpr bool HC_ = false; CALL; } } break; \
        case 2: { constexpr int NK_ = 2; if (HASCAM) { constexpr bool HC_ = true; CALL; } else { constexpr bool HC_ = false; CALL; } } break; \
        case 3: { constexpr int NK_ = 3; if (HASCAM) { constexpr bool HC_ = true; CALL; } else { constexpr bool HC_ = false; CALL; } } break; \
        case 4: { constexpr int NK_ = 4; if (HASCAM) { constexpr bool HC_ = true; CALL; } else { constexpr bool HC_ = false; CALL; } } break; \
        case 5: { constexpr int NK_ = 5; if (HASCAM) { constexpr bool HC_ = true; CALL; } else { constexpr bool HC_ = false; CALL; } } break; \
        case 6: { constexpr int NK_ = 6; if (HASCAM) { constexpr bool HC_ = true; CALL; } else { constexpr bool HC_ = false; CALL; } } break; \
        case 7: { constexpr int NK_ = 7; if (HASCAM) { constexpr bool HC_ = true; CALL; } else { constexpr bool HC_ = false; CALL; } } break; \
        case 8: { constexpr int NK_ = 8; if (HASCAM) { constexpr bool HC_ = true; CALL; } else { constexpr bool HC_ = false; CALL; } } break; \
        default: return cudaErrorInvalidValue;                               \
    }

cudaError_t launch_tables(const DevProblem& P, const double* eop, const double* iop, const double* cam_box,
                          double* img_tab, double* cam_tab, cudaStream_t st) {
    const int n = P.n_img > P.n_cam ? P.n_img : P.n_cam;
    k_tables<<<(n + 127) / 128, 128, 0, st>>>(P.n_img, P.n_cam, P.NK, eop, iop, cam_box, img_tab, cam_tab);
    return cudaGetLastError();
}

cudaError_t launch_clear_blocks(const DevProblem& P, const int2* blocks, int n_blocks, cudaStream_t st) {
    if (n_blocks > 0) k_clear_blocks<<<n_blocks, dim3(32, 8), 0, st>>>(P, blocks);
    return cudaGetLastError();
}

// first half of the border stage: G rows of the current EOPs and dg = diag(S) of this rank's partial system
// (a group sums dg over its ranks before launch_border_scale: conditioning of G and Jacobi scaling must be
// the same everywhere)
cudaError_t launch_border_prepare(const DevProblem& P, const double* eop, double* dg, cudaStream_t st, int64_t* launches) {
    if (P.inner) {
        k_G_rows<<<(P.n_img + 127) / 128, 128, 0, st>>>(P, eop);
        ++*launches;
    }
    k_diag_extract<<<(P.n_pad + 255) / 256, 256, 0, st>>>(P, dg);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_border_scale(const DevProblem& P, const double* dg, double* dvec, int* info, const int2* blocks,
                                int n_blocks, double* gwork, cudaStream_t st, int64_t* launches) {
    if (P.inner) {
        // gwork: kGramCtas x 56 Gram partials, then the 7 x 7 matrix C
        k_G_gram<<<kGramCtas, 256, 0, st>>>(P, dg, gwork);
        k_G_matrix<<<1, 64, 0, st>>>(gwork, kGramCtas, gwork + 56 * kGramCtas);
        k_G_apply<<<(P.n_pad + 255) / 256, 256, 0, st>>>(P, gwork + 56 * kGramCtas);
        *launches += 3;
        if (P.datum) {
            k_datum_split<<<(P.n_pad + 255) / 256, 256, 0, st>>>(P);
            ++*launches;
        }
    }
    if (P.n_pad > P.n_red) {
        k_pad_diag<<<(P.n_pad + 255) / 256, 256, 0, st>>>(P);
        ++*launches;
    }
    k_diag_scale<<<(P.n_pad + 255) / 256, 256, 0, st>>>(P, dg, dvec, info);
    k_border_scale<<<n_blocks, dim3(32, 8), 0, st>>>(P, dvec, blocks);
    *launches += 2;
    return cudaGetLastError();
}

static int backsub_lanes(const DevProblem& P) { return (P.n_seg > 0 && P.n_obs <= 12 * (int64_t)P.n_seg) ? 16 : 32; }

// number of point groups of the back-substitution grid (= partial sums it writes)
int backsub_warps(const DevProblem& P, int sm_count) {
    const int per_cta = 128 / backsub_lanes(P);
    int grid = (P.n_seg + per_cta - 1) / per_cta;
    const int cap = sm_count * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    return grid * per_cta;
}

cudaError_t launch_backsub(const DevProblem& P, int sm_count, cudaStream_t st) {
    const bool hc = P.uc > 0;
    const int G = backsub_lanes(P);
    const int grid = backsub_warps(P, sm_count) / (128 / G);
    if (P.pt_rec) {
        if (G == 16) {
            FEBA_NK_DISPATCH(P.NK, hc, (k_backsub_rec<NK_, HC_, 16><<<grid, 128, 0, st>>>(P)));
        } else {
            FEBA_NK_DISPATCH(P.NK, hc, (k_backsub_rec<NK_, HC_, 32><<<grid, 128, 0, st>>>(P)));
        }
        return cudaGetLastError();
    }
    if (G == 16) {
        FEBA_NK_DISPATCH(P.NK, hc, (k_backsub<NK_, HC_, 16><<<grid, 128, 0, st>>>(P)));
    } else {
        FEBA_NK_DISPATCH(P.NK, hc, (k_backsub<NK_, HC_, 32><<<grid, 128, 0, st>>>(P)));
    }
    return cudaGetLastError();
}

// group runs: keep the rows this rank contributes (its own subtrees; shared rows on rank 0), zero the rest, so
// that the sum over the ranks gives every row exactly once
__global__ void k_keep_own_rows(DevProblem P, double* __restrict__ vec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.n_pad && !row_init_here(P, i)) vec[i] = 0.0;
}

cudaError_t launch_keep_own_rows(const DevProblem& P, double* vec, cudaStream_t st) {
    k_keep_own_rows<<<(P.n_pad + 255) / 256, 256, 0, st>>>(P, vec);
    return cudaGetLastError();
}

// group runs: zero the entries of tie points owned by other ranks (tie_mine[t] = 1: owned here)
__global__ void k_keep_own_ties(int64_t n_tie, const unsigned char* __restrict__ tie_mine, double* __restrict__ xyz3) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < 3 * n_tie && !tie_mine[i / 3]) xyz3[i] = 0.0;
}

cudaError_t launch_keep_own_ties(int64_t n_tie, const unsigned char* tie_mine, double* xyz3, cudaStream_t st) {
    if (n_tie > 0) k_keep_own_ties<<<(unsigned)((3 * n_tie + 255) / 256), 256, 0, st>>>(n_tie, tie_mine, xyz3);
    return cudaGetLastError();
}

// group runs, owned-only transfers: packed coordinates of the tie points this rank owns <-> xhat on the device
__global__ void k_ties_pack(int n_own, const int* __restrict__ own_ties, int64_t n_red, const double* __restrict__ xhat,
                            double* __restrict__ packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * n_own) packed[i] = xhat[n_red + 3 * (int64_t)own_ties[i / 3] + i % 3];
}
__global__ void k_ties_unpack(int n_own, const int* __restrict__ own_ties, int64_t n_red, double* __restrict__ xhat,
                              const double* __restrict__ packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * n_own) xhat[n_red + 3 * (int64_t)own_ties[i / 3] + i % 3] = packed[i];
}
// same layout on both sides (the tie part of xhat): only the owned entries move.  One side may be page-locked HOST
// memory addressed by the device (zero copy over PCIe): no staging buffer, no host-side gather
__global__ void k_ties_copy_owned(int n_own, const int* __restrict__ own_ties, const double* __restrict__ src,
                                  double* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * n_own) {
        const int64_t at = 3 * (int64_t)own_ties[i / 3] + i % 3;
        dst[at] = src[at];
    }
}
cudaError_t launch_ties_copy_owned(int n_own, const int* own_ties, const double* src, double* dst, cudaStream_t st) {
    if (n_own <= 0) return cudaSuccess;
    k_ties_copy_owned<<<(3 * n_own + 255) / 256, 256, 0, st>>>(n_own, own_ties, src, dst);
    return cudaGetLastError();
}
cudaError_t launch_ties_pack(int n_own, const int* own_ties, int64_t n_red, double* xhat, double* packed, bool unpack,
                             cudaStream_t st) {
    if (n_own <= 0) return cudaSuccess;
    const int grid = (3 * n_own + 255) / 256;
    if (unpack) k_ties_unpack<<<grid, 256, 0, st>>>(n_own, own_ties, n_red, xhat, packed);
    else k_ties_pack<<<grid, 256, 0, st>>>(n_own, own_ties, n_red, xhat, packed);
    return cudaGetLastError();
}

cudaError_t launch_update_cam(const DevProblem& P, const double* sol, const double* dvec, double* dcam,
                              double* dcam_unscaled, double* eop, double* iop, double* out_sumabs, cudaStream_t st) {
    k_update_cam<<<1, 1024, 0, st>>>(P, sol, dvec, dcam, dcam_unscaled, eop, iop, out_sumabs);
    return cudaGetLastError();
}

cudaError_t launch_sum_partials(const double* partial, int n, int stride, int offset, double* out,
                                cudaStream_t st) {
    k_sum_partials<<<1, 1024, 0, st>>>(partial, n, stride, offset, out);
    return cudaGetLastError();
}

static int stream_grid(int64_t n, int sm_count) {
    int64_t grid = (n + 255) / 256;
    if (grid > (int64_t)sm_count * 8) grid = (int64_t)sm_count * 8;
    return grid < 1 ? 1 : (int)grid;
}

cudaError_t launch_xhat_scatter(const DevProblem& P, int sm_count, const double* xhat, double* eop, double* iop,
                                const int* tie_pt, cudaStream_t st) {
    const int64_t n = (int64_t)P.n_img * 6 + (int64_t)P.n_cam * P.NC + (int64_t)P.n_tie * 3;
    k_xhat_scatter<<<stream_grid(n, sm_count), 256, 0, st>>>(P, xhat, eop, iop, tie_pt);
    return cudaGetLastError();
}

cudaError_t launch_xhat_gather(const DevProblem& P, int sm_count, double* xhat, const double* eop,
                               const double* iop, const int* tie_pt, cudaStream_t st) {
    const int64_t n = (int64_t)P.n_img * 6 + (int64_t)P.n_cam * P.NC + (int64_t)P.n_tie * 3;
    k_xhat_gather<<<stream_grid(n, sm_count), 256, 0, st>>>(P, xhat, eop, iop, tie_pt);
    return cudaGetLastError();
}

cudaError_t launch_delta_gather(const DevProblem& P, int sm_count, double* delta, cudaStream_t st) {
    k_delta_gather<<<stream_grid((int64_t)P.n_red + 3 * (int64_t)P.n_tie, sm_count), 256, 0, st>>>(P, delta);
    return cudaGetLastError();
}

cudaError_t launch_cov_points(const DevProblem& P, int sm_count, const double* Q, const double* Y, const double* T7inv,
                              const double* dvec, double* out, cudaStream_t st) {
    const bool hc = P.uc > 0;
    int grid = (P.n_seg + 3) / 4;
    if (grid > sm_count * 8) grid = sm_count * 8;
    if (grid < 1) grid = 1;
    FEBA_NK_DISPATCH(P.NK, hc, (k_cov_points<NK_, HC_><<<grid, 128, 0, st>>>(P, Q, Y, T7inv, dvec, out)));
    return cudaGetLastError();
}

cudaError_t launch_cov_diag_cam(const DevProblem& P, const double* Q, const double* Y, const double* T7inv,
                                const double* dvec, double* out, cudaStream_t st) {
    k_cov_diag_cam<<<(P.n_red + 127) / 128, 128, 0, st>>>(P, Q, Y, T7inv, dvec, out);
    return cudaGetLastError();
}

cudaError_t launch_cov_block(const DevProblem& P, const double* Q, const double* Y, const double* T7inv,
                             const double* dvec, const long long* idx, int k, double* out, cudaStream_t st) {
    k_cov_block<<<(k * k + 127) / 128, 128, 0, st>>>(P, Q, Y, T7inv, dvec, idx, k, out);
    return cudaGetLastError();
}

int residual_blocks(const DevProblem& P, int sm_count) {
    int64_t grid = (P.n_obs + 255) / 256;
    const int cap = sm_count * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    return (int)grid;
}

cudaError_t launch_residuals(const DevProblem& P, int sm_count, const int* opt, const double* xyz_prev,
                             const double* iop_new, double* v_out, double* rsd_out, cudaStream_t st) {
    const bool hc = P.uc > 0;
    const int grid = residual_blocks(P, sm_count);
    FEBA_NK_DISPATCH(P.NK, hc, (k_residuals<NK_, HC_><<<grid, 256, 0, st>>>(P, opt, xyz_prev, iop_new, v_out, rsd_out)));
    return cudaGetLastError();
}

}  // namespace feba
