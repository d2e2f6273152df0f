// Block-sparse form of the reduced camera system (opt-in, FEBA_SPARSE=1; single GPU, tile task graph).
//
// The reduced system S of a photogrammetric block couples two images only when they share a tie point
// (the pair schedule of feba_assemble.cu), so at supertile granularity S is banded plus the dense camera
// rows: BASELINE configs[3] has 2,680 of 16,653 non-zero lower 64x64 tiles after symbolic fill (1.0e10
// instead of 5.3e11 flop).  The free-network datum as the dense path applies it, M = S + G~ G~'
// (k_border_scale), fills every tile.  Here the factorised matrix is
//     M_s = S + E E',   E = G~ restricted to the rows of a few datum images
// (four at each end of the image order: their supertiles are coupled through the dense last block row
// anyway, so E E' adds no fill).  M_s is positive definite (a similarity transform that leaves two images
// in place is the identity) and, with the datum images at opposite ends of the block, as well conditioned
// as M (tests/research/sparse_reduced_prototype.py, step error against extended precision with the
// conditioned G~ of k_G_condition: dense form 1e-13 / 3e-12 at 400 / 900 images, this form 6e-13 / 2e-12).
// The bordered system
//     [S G~; G~' 0] [delta; k] = [-g; 0]         (main.m:428-437)
// is solved exactly by block elimination over M_s with t = E' delta as 7 further unknowns:
//     M_s delta + G~ k - E t = -g,   G~' delta = 0,   E' delta - t = 0.
// The augmented block row carries B = [g G~ E] (15 columns instead of 8) through the factorisation, which
// leaves T = -B' M_s^-1 B in the augmented diagonal block; sparse_border_solve() below turns T into the 14
// coefficients c with  L^-1 (g + G~ k - E t) = column 0 + sum_m c[m] column(1+m)  (k_combine).
//
// Plain C++ so that tests/host_model can compile it with g++ (tests/test_sparse_reduced_host.py).
#pragma once
#include <vector>

#include "feba_model.cuh"

namespace feba {

constexpr int kDatumCols = 7;                     // columns of G (BuildAwG.m:514-527)
constexpr int kSparseAugRows = 1 + 2 * kDatumCols;   // g, G~ (7), E (7)

// T: full symmetric (1 + 14) x (1 + 14) matrix -B' M_s^-1 B, index 0 = g, 1..7 = G~, 8..14 = E.
// coef[0..6] = k, coef[7..13] = -t.  Returns false when the 14x14 system is singular.
FEBA_HD bool sparse_border_solve(const double T[kSparseAugRows][kSparseAugRows], double coef[2 * kDatumCols]) {
    constexpr int q = kDatumCols, n = 2 * kDatumCols;
    double Q[n][n + 1];
    for (int i = 0; i < q; ++i) {
        for (int j = 0; j < q; ++j) {
            Q[i][j] = T[1 + i][1 + j];                    // G~' delta = 0
            Q[i][q + j] = -T[1 + i][1 + q + j];
            Q[q + i][j] = T[1 + q + i][1 + j];            // E' delta - t = 0
            Q[q + i][q + j] = -T[1 + q + i][1 + q + j] - (i == j ? 1.0 : 0.0);
        }
        Q[i][n] = -T[1 + i][0];
        Q[q + i][n] = -T[1 + q + i][0];
    }
    for (int c = 0; c < n; ++c) {                         // Gaussian elimination, partial pivoting
        int p = c;
        for (int r = c + 1; r < n; ++r)
            if (fabs(Q[r][c]) > fabs(Q[p][c])) p = r;
        if (Q[p][c] == 0.0) return false;
        if (p != c)
            for (int j = 0; j <= n; ++j) { const double t = Q[c][j]; Q[c][j] = Q[p][j]; Q[p][j] = t; }
        for (int r = c + 1; r < n; ++r) {
            const double f = Q[r][c] / Q[c][c];
            for (int j = c; j <= n; ++j) Q[r][j] -= f * Q[c][j];
        }
    }
    double z[n];
    for (int r = n - 1; r >= 0; --r) {
        double t = Q[r][n];
        for (int j = r + 1; j < n; ++j) t -= Q[r][j] * z[j];
        z[r] = t / Q[r][r];
    }
    for (int i = 0; i < q; ++i) {
        coef[i] = z[i];
        coef[q + i] = -z[q + i];
    }
    return true;
}

// ---- supertile pattern (host).  nz is (NT + 1) x (NT + 1) row-major, lower triangle in use; row NT is the
// augmented block row.  A supertile is T 64-blocks (the last one the remainder), as in chol_dag.

struct SparsePattern {
    int NT = 0, T = 0;
    std::vector<unsigned char> nz;
    int NR() const { return NT + 1; }
    bool at(int i, int j) const { return nz[(size_t)i * NR() + j] != 0; }
    void set(int i, int j) {
        if (i < j) { const int t = i; i = j; j = t; }
        nz[(size_t)i * NR() + j] = 1;
    }
    int supertile_of_row(int row) const {
        const int s = row / (64 * T);
        return s < NT ? s : NT - 1;
    }
    // envelope: first 64-block column that can be non-zero in 64-block row k (nb block rows)
    std::vector<int> row_first_block(int nb) const {
        std::vector<int> v((size_t)nb, 0);
        for (int k = 0; k < nb; ++k) {
            const int I = k / T < NT ? k / T : NT - 1;
            int J = 0;
            while (J < I && !at(I, J)) ++J;
            v[(size_t)k] = J * T;
        }
        return v;
    }
    // couple the unknowns [a0, a1] with [b0, b1] (inclusive row ranges of the reduced system)
    void couple(int a0, int a1, int b0, int b1) {
        for (int i = supertile_of_row(a0); i <= supertile_of_row(a1); ++i)
            for (int j = supertile_of_row(b0); j <= supertile_of_row(b1); ++j) set(i, j);
    }
};

// Datum images: four at each end of the image order (all of them when the block is tiny), spread from the
// first image to the last one entirely inside supertile 0, and from the first one entirely inside the last
// supertile to the last image.  (Step error against extended precision with the conditioned G~, 900 images:
// 4 images 1e-11, 8 images 2e-12, 16 images 5e-13; tests/research/sparse_reduced_prototype.py.)
constexpr int kDatumImages = 8;
inline std::vector<int> sparse_datum_images(int n_img, int ui, int nb, int T) {
    std::vector<int> v;
    if (n_img <= kDatumImages) {
        for (int i = 0; i < n_img; ++i) v.push_back(i);
        return v;
    }
    const int NT = (nb + T - 1) / T, h = kDatumImages / 2;
    int b = (64 * T) / ui - 1;                              // last image entirely inside supertile 0
    if (b < h - 1) b = h - 1;
    if (b > n_img / 2 - 1) b = n_img / 2 - 1;
    int c = ((NT - 1) * 64 * T + ui - 1) / ui;              // first image entirely inside the last supertile
    if (c > n_img - h) c = n_img - h;
    if (c <= b) c = b + 1;
    for (int i = 0; i < h; ++i) v.push_back((int)((long long)b * i / (h - 1)));
    for (int i = 0; i < h; ++i) v.push_back(c + (int)((long long)(n_img - 1 - c) * i / (h - 1)));
    return v;
}

// blocks_ab: n_blocks pairs (image a, image b) of the pair schedule.  ui unknowns per image, camera rows
// [off_cam, n_red), padding rows up to 64 nb.  Symbolic fill included.
inline SparsePattern sparse_supertile_pattern(int nb, int T, int ui, int n_img, int off_cam, int n_red, int n_blocks,
                                              const int* blocks_ab, const std::vector<int>& datum) {
    SparsePattern P;
    P.T = T;
    P.NT = (nb + T - 1) / T;
    const int NT = P.NT, NR = NT + 1;
    P.nz.assign((size_t)NR * NR, 0);
    auto img_rows = [&](int im, int& r0, int& r1) { r0 = ui * im; r1 = ui * im + ui - 1; };
    for (int s = 0; s < NT; ++s) P.set(s, s);
    for (int im = 0; im < n_img && ui > 0; ++im) {         // an image block may straddle two supertiles
        int r0, r1;
        img_rows(im, r0, r1);
        P.couple(r0, r1, r0, r1);
    }
    for (int e = 0; e < n_blocks && ui > 0; ++e) {
        int a0, a1, b0, b1;
        img_rows(blocks_ab[2 * e], a0, a1);
        img_rows(blocks_ab[2 * e + 1], b0, b1);
        P.couple(a0, a1, b0, b1);
    }
    for (size_t x = 0; x < datum.size() && ui > 0; ++x)
        for (size_t y = 0; y <= x; ++y) {
            int a0, a1, b0, b1;
            img_rows(datum[x], a0, a1);
            img_rows(datum[y], b0, b1);
            P.couple(a0, a1, b0, b1);
        }
    if (n_red > off_cam)                                    // camera rows: dense
        for (int i = P.supertile_of_row(off_cam); i <= P.supertile_of_row(n_red - 1); ++i)
            for (int j = 0; j <= i; ++j) P.set(i, j);
    for (int j = 0; j <= NT; ++j) P.set(NT, j);             // augmented block row
    for (int k = 0; k < NT; ++k)                            // symbolic factorisation
        for (int i = k + 1; i < NR; ++i) {
            if (!P.at(i, k)) continue;
            for (int j = k + 1; j <= i; ++j)
                if (P.at(j, k)) P.set(i, j);
        }
    return P;
}
}  // namespace feba
