// Dense factorisation and solve of the reduced camera system on the device (FP64).
//
// Replaces the explicit inverses of the reference, main.m:432 (Cx = NG^-1, bordered with the
// inner-constraint matrix G) and main.m:442 (Cx = N^-1), for the point-eliminated system:
//   M = S (+ Gc Gc' with inner constraints),  M = L L'   (recursive blocked Cholesky)
//   the right-hand side g and the 7 columns of Gc ride along as an augmented block row, so the
//   forward substitutions Y = L^-1 [g Gc] and the 8x8 Schur complement T = -[g Gc]' M^-1 [g Gc]
//   fall out of the factorisation's own TRSM / SYRK steps;
//   border: (Gc' M^-1 Gc) k = -Gc' M^-1 g ;  sol = L^-T (y_g + Y_G k) ;  delta_c = -sol.
// The one real dense contraction of the path (trailing SYRK/GEMM updates) runs on the FP64
// tensor pipe: mma.sync.m8n8k4.f64 (DMMA).  tcgen05 has no FP64 kind, so the warp-level DMMA is
// the tensor path for doubles on sm_100a as well.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include <vector>

#include "feba_dev.h"
#include "feba_kernels.h"
#include "feba_sparse.h"

namespace feba {

// ------------------------------------------------------------------------------------------
// C (M x N) -= A (M x K) * B (N x K)'   -- column-major, all dimensions multiples of 64.
// CTA tile 64x64, 4 warps of 32x32, K tile 16, 3-stage cp.async pipeline, DMMA m8n8k4 -- the only
// FP64 tensor shape sm_100a has in SASS (m16n8k{4,8,16} all lower to DMMA.8x8x4; measured issue peak
// 37.1 TFLOP/s, scripts/ubench/dmma_rate.cu).  Measured on the top-level trailing update of the
// u_c = 12,010 system (95x95 lower tiles, K = 6016): 33.9 TFLOP/s, tensor pipe 82 % active
// (profiles/).  A 128x128-tile variant (8 warps of 64x32, 164 registers, one CTA per SM) was tried
// and was slower at every size of this recursion (factorisation 36.8 -> 38.9..54.6 ms).
constexpr int GT = 64;      // CTA tile
constexpr int GK = 16;      // K tile
constexpr int GS = 68;      // smem row stride in doubles (== 4 mod 16: conflict-free fragments)
constexpr int GSTAGES = 3;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// MODE 0: C -= A B'   MODE 1: the same, tiles of the lower block triangle only (SYRK-shaped update)
// MODE 2: C = A B' (overwrite).  MODE 2 may run in place (C == A, one tile column, K == 64): every
// global read of the CTA's A rows has landed in shared memory before its first store.
// MODE 3: C = A B' (overwrite), lower tiles only, for UPPER-triangular A == B (rows bx of A are zero left of
// column block bx, so the K loop starts there): Q = U U' of the covariance stage.
template <int MODE>
__global__ void __launch_bounds__(128) k_gemm_nt(double* __restrict__ C, int ldc, const double* __restrict__ A,
                                                 int lda, const double* __restrict__ B, int ldb, int K) {
    const int bx = blockIdx.x, by = blockIdx.y;
    if ((MODE == 1 || MODE == 3) && by > bx) return;
    extern __shared__ __align__(16) double gsm[];
    double(*As)[GK][GS] = reinterpret_cast<double(*)[GK][GS]>(gsm);
    double(*Bs)[GK][GS] = reinterpret_cast<double(*)[GK][GS]>(gsm + GSTAGES * GK * GS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 1, wn = warp >> 1;
    const int k0 = MODE == 3 ? bx * GT : 0;                 // first non-zero column of A's row block
    const double* Ag = A + (size_t)bx * GT + (size_t)lda * k0;
    const double* Bg = B + (size_t)by * GT + (size_t)ldb * k0;
    const int nk = (K - k0) / GK;

    auto load_stage = [&](int s, int kt) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int q = tid + 128 * t;
            const int k = q >> 5, i2 = (q & 31) * 2;
            cp_async16(&As[s][k][i2], Ag + i2 + (size_t)lda * (kt * GK + k));
            cp_async16(&Bs[s][k][i2], Bg + i2 + (size_t)ldb * (kt * GK + k));
        }
    };
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < GSTAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    const int lr = lane >> 2, lk = lane & 3;
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();
        const int nxt = kt + GSTAGES - 1;
        if (nxt < nk) load_stage(nxt % GSTAGES, nxt);
        cp_async_commit();
        const int s = kt % GSTAGES;
#pragma unroll
        for (int kk = 0; kk < GK / 4; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                a[t] = As[s][kk * 4 + lk][wm * 32 + t * 8 + lr];
                b[t] = Bs[s][kk * 4 + lk][wn * 32 + t * 8 + lr];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    if (MODE == 2) __syncthreads();
    double* Cg = C + (size_t)bx * GT + (size_t)ldc * by * GT;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = wm * 32 + i * 8 + lr;
            const int c = wn * 32 + j * 8 + 2 * lk;
            if (MODE == 2 || MODE == 3) {
                Cg[r + (size_t)ldc * c] = acc[i][j][0];
                Cg[r + (size_t)ldc * (c + 1)] = acc[i][j][1];
            } else {
                Cg[r + (size_t)ldc * c] -= acc[i][j][0];
                Cg[r + (size_t)ldc * (c + 1)] -= acc[i][j][1];
            }
        }
}

#ifdef FEBA_POTRF_PROF
__device__ long long g_potrf_prof[32];
#define POTRF_T(i) do { if (threadIdx.x == 0) g_potrf_prof[i] = clock64(); } while (0)
#else
#define POTRF_T(i) do { } while (0)
#endif
// ------------------------------------------------------------------------------------------
// 64x64 Cholesky of a diagonal block (lower) and the inverse of its factor, one CTA of 256 threads,
// matrix in shared memory, blocked by 16:
//   for each 16-column panel: (1) 16x16 diagonal factor (row per lane, columns broadcast with
//   shuffles; every warp does it redundantly so that the shuffles stay in convergent code),
//   (2) rows below: X L_kk^-T by substitution, one thread per row, (3) rank-16 update of the trailing
//   lower triangle spread over the CTA.  Phase timings (cycles, scripts/ubench/potrf_bench.cu): load 1.1k,
//   4 x (diag16 4.9k, panel 0.8k, trailing 2.0/1.8/1.1k), inverse 0.9k + 6.8k, stores 1.5k: 37.8k = 20 us.
// Then Linv = L^-1: the four 16x16 diagonal blocks by substitution on the identity (one thread per
// column), the six off-diagonal blocks by block forward substitution
//   X_ij = -Linv_ii * sum_{k=j}^{i-1} L_ik X_kj      (X_jj = Linv_jj), by block distance 1, 2, 3.
// Linv (full 64x64, zero upper triangle) lets every later triangular solve with this block run as
// a DMMA GEMM.  (A thread-per-entry version with predicated register slots needed ~350 SASS
// instructions per elimination step: 78 us per block, ~15 ms per factorisation at u_c = 12,010.)
// 1/sqrt(x) for a positive, normally scaled x (the matrix is Jacobi-scaled): single-precision
// reciprocal square root + two Newton steps in double.  Sits on the pivot-to-pivot dependency chain
// of the factorisation, where the library rsqrt costs about twice as much.
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y = (double)rsqrtf((float)x);
    const double hx = 0.5 * x;
    y = y * (1.5 - hx * y * y);
    y = y * (1.5 - hx * y * y);
    return y;
}

// shared-memory views of one 64x64 factorisation (2 * 64 * 65 + 3 * 16 * 17 + 64 doubles)
struct PotrfSmem {
    double (*sA)[kBlk + 1];      // matrix (lower triangle, upper zero) -> L
    double (*sI)[kBlk + 1];      // zero -> L^-1
    double (*sT)[16][17];        // block products of the inverse
    double* rdiag;               // 1 / L_jj
};
constexpr size_t kPotrfSmemDoubles = 2 * kBlk * (kBlk + 1) + 3 * 16 * 17 + kBlk;

__device__ __forceinline__ PotrfSmem potrf_views(double* psm) {
    PotrfSmem S;
    S.sA = reinterpret_cast<double(*)[kBlk + 1]>(psm);
    S.sI = reinterpret_cast<double(*)[kBlk + 1]>(psm + kBlk * (kBlk + 1));
    S.sT = reinterpret_cast<double(*)[16][17]>(psm + 2 * kBlk * (kBlk + 1));
    S.rdiag = psm + 2 * kBlk * (kBlk + 1) + 3 * 16 * 17;
    return S;
}

// sA := chol(sA) (lower), 256 threads, all of them call; returns true when a pivot was not positive
__device__ __forceinline__ bool potrf64_factor(const PotrfSmem& S, int tid) {
    double(*sA)[kBlk + 1] = S.sA;
    double* rdiag = S.rdiag;
    const int lane = tid & 31, warp = tid >> 5;
    bool bad = false;
#pragma unroll 1
    for (int kb = 0; kb < 4; ++kb) {
        const int o = 16 * kb;
        // (1) diagonal block: every warp factorises it redundantly (the shuffles then sit in convergent
        // code: inside `if (warp == 0)` each one was wrapped in a WARPSYNC.COLLECTIVE sequence), warp 0 stores
        {
            const int r = lane & 15;
            double d[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) d[c] = sA[o + r][o + c];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const double piv = __shfl_sync(0xffffffffu, d[j], j);
                if (!(piv > 0.0)) bad = true;
                const double rs = fast_rsqrt(piv);
                const double l = d[j] * rs;                 // L(r, j) for r >= j
                d[j] = l;
#pragma unroll
                for (int c = j + 1; c < 16; ++c) {
                    const double lc = __shfl_sync(0xffffffffu, l, c);
                    d[c] -= l * lc;                         // rows r < c hold unused values
                }
            }
            __syncthreads();                                // every warp has read the block
            if (warp == 0 && lane < 16) {
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    if (c <= r) sA[o + r][o + c] = d[c];
                rdiag[o + r] = 1.0 / d[r];
            }
        }
        __syncthreads();
        
        // (2) panel below the diagonal block: one thread per row, column-oriented substitution (after
        // x_j is final the remaining right-hand sides are updated independently: short dependency chain)
        const int nrow = 48 - o;
        if (tid < nrow) {
            const int r = o + 16 + tid;
            double x[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) x[c] = sA[r][o + c];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                x[j] *= rdiag[o + j];
#pragma unroll
                for (int c = j + 1; c < 16; ++c) x[c] -= x[j] * sA[o + c][o + j];
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) sA[r][o + c] = x[c];
        }
        __syncthreads();
        
        // (3) trailing update, lower triangle of the remaining nrow x nrow block: thread (row i, group g of
        // 5) keeps its row of the panel in registers and walks the columns jj = g, g+5, ... <= i, two
        // independent accumulation chains at a time.  (A DMMA version of this update and of the inverse
        // blocks below was measured slower overall: 42.1k vs 37.8k cycles per block.)
        if (tid < 5 * nrow) {
            const int i = tid % nrow, g = tid / nrow;
            const int r = o + 16 + i;
            double a[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = sA[r][o + k];
            int jj = g;
            for (; jj + 5 <= i; jj += 10) {
                const int c0 = o + 16 + jj, c1 = c0 + 5;
                double acc0 = sA[r][c0], acc1 = sA[r][c1];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    acc0 -= a[k] * sA[c0][o + k];
                    acc1 -= a[k] * sA[c1][o + k];
                }
                sA[r][c0] = acc0;
                sA[r][c1] = acc1;
            }
            if (jj <= i) {
                const int c0 = o + 16 + jj;
                double acc0 = sA[r][c0];
#pragma unroll
                for (int k = 0; k < 16; ++k) acc0 -= a[k] * sA[c0][o + k];
                sA[r][c0] = acc0;
            }
        }
        __syncthreads();
        
    }
    return bad;
}

// sI := sA^-1 (lower), 256 threads; sI must be zero above and at entry
__device__ __forceinline__ void potrf64_inverse(const PotrfSmem& S, int tid) {
    double(*sA)[kBlk + 1] = S.sA;
    double(*sI)[kBlk + 1] = S.sI;
    double(*sT)[16][17] = S.sT;
    double* rdiag = S.rdiag;
    // ---- inverse, diagonal blocks: thread (b, c) solves L_bb x = e_c, column-oriented
    if (tid < 64) {
        const int o = tid & ~15, c = tid & 15;
        double x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = (i == c) ? 1.0 : 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            x[i] *= rdiag[o + i];
#pragma unroll
            for (int i2 = i + 1; i2 < 16; ++i2) x[i2] -= sA[o + i2][o + i] * x[i];
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) sI[o + i][o + c] = x[i];
    }
    __syncthreads();
    // ---- off-diagonal blocks by block distance dist = i - j (fully unrolled: every bound is static, so
    // the shared-memory loads of a dot product are issued together instead of one round trip per term)
    const int er = tid >> 4, ec = tid & 15;            // one entry of a 16x16 block per thread
#pragma unroll
    for (int dist = 1; dist <= 3; ++dist) {
        // T_ij = sum_{k=j}^{i-1} L_ik X_kj
#pragma unroll
        for (int bj = 0; bj < 4 - dist; ++bj) {
            const int bi = bj + dist;
            double acc = 0.0;
#pragma unroll
            for (int k = 16 * bj; k < 16 * bi; ++k) acc += sA[16 * bi + er][k] * sI[k][16 * bj + ec];
            sT[bj][er][ec] = acc;
        }
        __syncthreads();
        // X_ij = -Linv_ii T_ij
#pragma unroll
        for (int bj = 0; bj < 4 - dist; ++bj) {
            const int bi = bj + dist;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 16; ++k) acc += sI[16 * bi + er][16 * bi + k] * sT[bj][k][ec];
            sI[16 * bi + er][16 * bj + ec] = -acc;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_potrf64_inv(double* __restrict__ A, int ld, double* __restrict__ Linv,
                                                    int* __restrict__ info) {
    extern __shared__ __align__(16) double psm[];
    const PotrfSmem S = potrf_views(psm);
    double(*sA)[kBlk + 1] = S.sA;
    double(*sI)[kBlk + 1] = S.sI;
    const int tid = threadIdx.x, lane = tid & 31;
    POTRF_T(0);
    {
        // all 16 loads of a thread in flight together (a rolled load->store loop costs one L2 round
        // trip per iteration: 4.2k of the kernel's 44k cycles)
        double v[16];
        const int r = tid & 63, c0 = tid >> 6;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int c = c0 + 4 * t;
            v[t] = (r >= c) ? A[r + (size_t)ld * c] : 0.0;
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int c = c0 + 4 * t;
            sA[r][c] = v[t];
            sI[r][c] = 0.0;
        }
    }
    __syncthreads();
    POTRF_T(1);
    const bool bad = potrf64_factor(S, tid);
    POTRF_T(5);
    if (bad && lane == 0) atomicExch(info, 1);
    {
        const int r = tid & 63, c0 = tid >> 6;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int c = c0 + 4 * t;
            if (r >= c) A[r + (size_t)ld * c] = sA[r][c];
        }
    }
    POTRF_T(6);
    potrf64_inverse(S, tid);
    POTRF_T(8);
    {
        const int r = tid & 63, c0 = tid >> 6;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int c = c0 + 4 * t;
            Linv[r + (size_t)kBlk * c] = sI[r][c];
        }
    }
    POTRF_T(9);
}

// ------------------------------------------------------------------------------------------
// Fused triangular solve against up to four 64-column blocks: X (64 rows per CTA, nt*64 columns)
// := X * L^-T with L the nt x nt block triangle at (c0, c0):
//   for j = 0..nt-1:  X_j := (X_j - sum_{i<j} X_i L_ji') * Linv_j'
// The CTA keeps its 64 x (nt*64) row block in shared memory (139 KB), streams the L_ji / Linv_j
// tiles through a double buffer with cp.async and runs every product on DMMA.  One launch replaces
// the 2 nt - 1 launches (nt leaf multiplications + nt - 1 small GEMMs) of the recursion below four
// blocks: those small launches were latency-bound (profiles/: 685 leaf + ~370 small GEMM launches,
// ~13 ms of a 42 ms factorisation under ncu).
constexpr int TS = 68;                      // smem stride (doubles) of a 64-wide tile row: == 4 mod 16
constexpr int TF_MAX = 4;                  // 64-row strips (139 KB of shared memory)
constexpr int TF_MAX16 = 6;                // 16-row strips: a whole supertile column of the plan (tile_max = 6) per launch
constexpr size_t kTrsmFusedSmem = (size_t)(TF_MAX + 2) * 64 * TS * sizeof(double);
constexpr size_t kTrsmFusedSmem16 = (size_t)(TF_MAX16 * 64 * 20 + 2 * 64 * TS) * sizeof(double);

// dst[k][i] = src[i + ld * k] for ROWS rows x 64 columns, 16-byte chunks, 256 threads
template <int ROWS, int STRIDE>
__device__ __forceinline__ void load_tile(double (*dst)[STRIDE], const double* __restrict__ src, int ld, int tid) {
    constexpr int CH = ROWS / 2;                   // 16-byte chunks per column
#pragma unroll
    for (int t = 0; t < (64 * CH + 255) / 256; ++t) {
        const int q = tid + 256 * t;
        if (64 * CH % 256 == 0 || q < 64 * CH) {
            const int k = q / CH, i2 = (q % CH) * 2;
            cp_async16(&dst[k][i2], src + i2 + (size_t)ld * k);
        }
    }
}

// RM = rows of X per CTA: 64 (one 64-row block per CTA) or 16 (four CTAs per block: more CTAs for the
// deep levels of the recursion where only a few row blocks exist).
template <int RM>
__global__ void __launch_bounds__(256) k_trsm_fused(double* X, int ldx, const double* A, int ld,
                                                    const double* __restrict__ Linv, int r0, int c0, int nt) {
    extern __shared__ __align__(16) double fsm[];
    constexpr int XS = RM == 64 ? TS : 20;        // row stride of the X tiles (== 4 mod 16 either way)
    double(*Xs)[64][XS] = reinterpret_cast<double(*)[64][XS]>(fsm);                          // [tile][k][row]
    double(*Ls)[64][TS] = reinterpret_cast<double(*)[64][TS]>(fsm + (size_t)(RM == 64 ? TF_MAX : TF_MAX16) * 64 * XS); // [buf][k][n]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // RM = 64: 2 x 4 warps of 32 rows x 16 columns; RM = 16: 1 x 8 warps of 16 rows x 8 columns
    constexpr int MF = RM == 64 ? 4 : 2, NF = RM == 64 ? 2 : 1;
    const int wm = RM == 64 ? (warp & 1) : 0, wn = RM == 64 ? (warp >> 1) : warp;
    const int row0 = wm * 32, col0 = wn * (8 * NF);
    const int lr = lane >> 2, lk = lane & 3;
    // X block (r, c) lives at X + r*64 + ldx*c*64 (X == A, ldx == ld inside the factorisation)
    double* Xg = X + (size_t)r0 * kBlk + (size_t)blockIdx.x * RM + (size_t)ldx * c0 * kBlk;
    auto Btile = [&](int seq, const double*& src, int& bld) {
        // sequence: j = 0: Linv_0 | j = 1: L_10, Linv_1 | j = 2: L_20, L_21, Linv_2 | ...
        int j = 0;
        while ((j + 1) * (j + 2) / 2 <= seq) ++j;
        const int i = seq - j * (j + 1) / 2;
        if (i == j) {
            src = Linv + (size_t)(c0 + j) * kBlk * kBlk;
            bld = kBlk;
        } else {
            src = A + (size_t)(c0 + j) * kBlk + (size_t)ld * (c0 + i) * kBlk;
            bld = ld;
        }
    };
    const int nseq = nt * (nt + 1) / 2;
    for (int t = 0; t < nt; ++t) load_tile<RM, XS>(Xs[t], Xg + (size_t)ldx * t * kBlk, ldx, tid);
    {
        const double* src;
        int bld;
        Btile(0, src, bld);
        load_tile<64, TS>(Ls[0], src, bld, tid);
    }
    cp_async_commit();
    int seq = 0;
    double acc[MF][NF][2];
    auto mma_tile = [&](const double (*Xa)[XS], const double (*Lb)[TS]) {
#pragma unroll 4
        for (int kk = 0; kk < 16; ++kk) {
            double a[MF], b[NF];
#pragma unroll
            for (int t = 0; t < MF; ++t) a[t] = Xa[kk * 4 + lk][row0 + t * 8 + lr];
#pragma unroll
            for (int t = 0; t < NF; ++t) b[t] = Lb[kk * 4 + lk][col0 + t * 8 + lr];
#pragma unroll
            for (int i = 0; i < MF; ++i)
#pragma unroll
                for (int j = 0; j < NF; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    };
    auto next_tile = [&]() {        // prefetch tile seq+1, then make tile seq visible to the whole CTA
        if (seq + 1 < nseq) {
            const double* src;
            int bld;
            Btile(seq + 1, src, bld);
            load_tile<64, TS>(Ls[(seq + 1) & 1], src, bld, tid);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
    };
    for (int j = 0; j < nt; ++j) {
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
            for (int jj = 0; jj < NF; ++jj) acc[i][jj][0] = acc[i][jj][1] = 0.0;
        for (int i = 0; i < j; ++i) {
            next_tile();
            mma_tile(Xs[i], Ls[seq & 1]);
            __syncthreads();                                  // Ls[seq & 1] may be refilled two tiles later
            ++seq;
        }
        if (j > 0) {
#pragma unroll
            for (int i = 0; i < MF; ++i)
#pragma unroll
                for (int jj = 0; jj < NF; ++jj) {
                    const int r = row0 + i * 8 + lr, c = col0 + jj * 8 + 2 * lk;
                    Xs[j][c][r] -= acc[i][jj][0];
                    Xs[j][c + 1][r] -= acc[i][jj][1];
                    acc[i][jj][0] = acc[i][jj][1] = 0.0;
                }
        }
        next_tile();                                          // also orders the X_j update above
        mma_tile(Xs[j], Ls[seq & 1]);
        __syncthreads();                                      // everyone has read X_j
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
            for (int jj = 0; jj < NF; ++jj) {
                const int r = row0 + i * 8 + lr, c = col0 + jj * 8 + 2 * lk;
                Xs[j][c][r] = acc[i][jj][0];
                Xs[j][c + 1][r] = acc[i][jj][1];
            }
        __syncthreads();
        ++seq;
    }
    cp_async_wait<0>();
    // write the rows back (coalesced over rows)
    constexpr int CH = RM / 2;
    for (int t = 0; t < nt; ++t)
        for (int q = tid; q < 64 * CH; q += 256) {
            const int k = q / CH, i2 = (q % CH) * 2;
            *reinterpret_cast<double2*>(Xg + i2 + (size_t)ldx * (t * kBlk + k)) =
                *reinterpret_cast<const double2*>(&Xs[t][k][i2]);
        }
}

#define AT(A, ld, br, bc) ((A) + (size_t)(br) * kBlk + (size_t)(ld) * (bc) * kBlk)
#define LINV(W, b) ((W) + (size_t)(b) * kBlk * kBlk)

// ------------------------------------------------------------------------------------------
// Left-looking factorisation of ONE 64-block column j of a supertile (blocks [b0, b0 + n) on the diagonal) in ONE
// launch: CTA r handles block row i = j + r of the tile,
//     D = A[j,j] - sum_{k<j} A[j,k] A[j,k]'   and   L = chol(D), Linv = L^-1     -- by EVERY CTA, redundantly
//     r == 0 stores Linv (A[j,j] keeps D, see below);   r > 0:  A[i,j] = (A[i,j] - sum_{k<j} A[i,k] A[j,k]') Linv'.
// The recursive form issues potrf -> triangular solve -> symmetric update per block, three dependent launches on
// the critical path of the factorisation (~67 us per block on BASELINE configs[3]); here the 20 us diagonal
// factorisation is repeated by the few CTAs of the column instead of being waited for, and a block of the chain
// costs one launch.  Products on DMMA (m8n8k4), tiles through shared memory.
constexpr size_t kColumnSmem = (kPotrfSmemDoubles + 2 * 64 * TS) * sizeof(double);

// finish != 0 (one CTA): only D = A[j,j] - sum_{k<j} A[j,k] A[j,k]' is formed and stored (lower triangle) -- the
// Schur complement T of the augmented block row, which is never factorised.
__device__ __forceinline__ void chol_column_body(double* __restrict__ A, int ld, double* __restrict__ Linv, int b0, int j,
                                                 int* __restrict__ info, int finish) {
    extern __shared__ __align__(16) double csm[];
    const PotrfSmem S = potrf_views(csm);
    double(*Tj)[TS] = reinterpret_cast<double(*)[TS]>(csm + kPotrfSmemDoubles);            // A[j,k] as [k][row]
    double(*Ti)[TS] = reinterpret_cast<double(*)[TS]>(csm + kPotrfSmemDoubles + 64 * TS);  // A[i,k] as [k][row]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 1, wn = warp >> 1;
    const int row0 = wm * 32, col0 = wn * 16;
    const int lr = lane >> 2, lk = lane & 3;
    const int bj = b0 + j, bi = bj + (int)blockIdx.x;
    const bool diag = blockIdx.x == 0;
    double accD[4][2][2], accR[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) accD[i][jj][0] = accD[i][jj][1] = accR[i][jj][0] = accR[i][jj][1] = 0.0;
    for (int k = 0; k < j; ++k) {
        __syncthreads();                                     // the tiles of the previous step have been read
        load_tile<64, TS>(Tj, AT(A, ld, bj, b0 + k), ld, tid);
        if (!diag) load_tile<64, TS>(Ti, AT(A, ld, bi, b0 + k), ld, tid);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < 16; ++kk) {
            double aj[4], ai[4], b[2];
#pragma unroll
            for (int t = 0; t < 4; ++t) aj[t] = Tj[kk * 4 + lk][row0 + t * 8 + lr];
#pragma unroll
            for (int t = 0; t < 2; ++t) b[t] = Tj[kk * 4 + lk][col0 + t * 8 + lr];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) dmma884(accD[i][jj][0], accD[i][jj][1], aj[i], b[jj]);
            if (!diag) {
#pragma unroll
                for (int t = 0; t < 4; ++t) ai[t] = Ti[kk * 4 + lk][row0 + t * 8 + lr];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) dmma884(accR[i][jj][0], accR[i][jj][1], ai[i], b[jj]);
            }
        }
    }
    __syncthreads();
    // D -> sA (lower triangle, zero above), sI := 0; the row block's right-hand side R -> Ti as [column][row]
    {
        const double* Djj = AT(A, ld, bj, bj);
        const double* Cij = AT(A, ld, bi, bj);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = row0 + i * 8 + lr, c = col0 + jj * 8 + 2 * lk + h;
                    S.sA[r][c] = r >= c ? Djj[r + (size_t)ld * c] - accD[i][jj][h] : 0.0;
                    S.sI[r][c] = 0.0;
                    if (!diag) Ti[c][r] = Cij[r + (size_t)ld * c] - accR[i][jj][h];
                }
    }
    __syncthreads();
    if (finish) {
        const int r = tid & 63, c0 = tid >> 6;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int c = c0 + 4 * t;
            if (r >= c) A[(size_t)bj * kBlk + r + (size_t)ld * ((size_t)bj * kBlk + c)] = S.sA[r][c];
        }
        return;
    }
    const bool bad = potrf64_factor(S, tid);
    // A[j,j] is NOT overwritten with L: the other CTAs of this launch read it (D) at their own pace, and no later
    // kernel reads the diagonal block of a factor -- triangular solves use Linv (forward: k_trsm_fused, backward:
    // k_backstep, inverse: rtrsm_identity) and the off-diagonal blocks only
    if (diag && bad && lane == 0) atomicExch(info, 1);
    potrf64_inverse(S, tid);
    __syncthreads();
    if (diag) {
        double* Lout = LINV(Linv, bj);
        const int r = tid & 63, c0 = tid >> 6;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int c = c0 + 4 * t;
            Lout[r + (size_t)kBlk * c] = S.sI[r][c];
        }
        return;
    }
    // X = R Linv':  X(r, n) = sum_c R(r, c) Linv(n, c)
    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) acc[i][jj][0] = acc[i][jj][1] = 0.0;
#pragma unroll 4
    for (int kk = 0; kk < 16; ++kk) {
        double a[4], b[2];
#pragma unroll
        for (int t = 0; t < 4; ++t) a[t] = Ti[kk * 4 + lk][row0 + t * 8 + lr];
#pragma unroll
        for (int t = 0; t < 2; ++t) b[t] = S.sI[col0 + t * 8 + lr][kk * 4 + lk];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) dmma884(acc[i][jj][0], acc[i][jj][1], a[i], b[jj]);
    }
    double* Xg = AT(A, ld, bi, bj);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const int r = row0 + i * 8 + lr, c = col0 + jj * 8 + 2 * lk;
            Xg[r + (size_t)ld * c] = acc[i][jj][0];
            Xg[r + (size_t)ld * (c + 1)] = acc[i][jj][1];
        }
}

__global__ void __launch_bounds__(256) k_chol_column(double* __restrict__ A, int ld, double* __restrict__ Linv,
                                                     int b0, int j, int* __restrict__ info, int finish) {
    chol_column_body(A, ld, Linv, b0, j, info, finish);
}

// The same column of MANY independent systems of one shape (BatchRun sweep, feba_batch): blockIdx.y = system.
__global__ void __launch_bounds__(256) k_chol_column_batched(double* const* __restrict__ As, int ld,
                                                             double* const* __restrict__ Linvs, int b0, int j,
                                                             int* const* __restrict__ infos, int finish) {
    chol_column_body(As[blockIdx.y], ld, Linvs[blockIdx.y], b0, j, infos[blockIdx.y], finish);
}

// Timing experiments only (results are garbage): FEBA_CHOL_SKIP bitmask drops kernel classes from the
// factorisation -- 1: potrf64, 2: fused triangular leaves, 4: products with < 256 tiles, 8: the rest.
static int chol_skip() {
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("FEBA_CHOL_SKIP");
        v = e ? std::atoi(e) : 0;
    }
    return v;
}

static cudaError_t gemm_nt(double* C, int ldc, const double* A, int lda, const double* B, int ldb, int mb,
                           int nbk, int kb, int mode, cudaStream_t st, int64_t* launches) {
    const size_t smem = 2 * GSTAGES * GK * GS * sizeof(double);
    static SmemOptIn o0, o1, o2, o3;
    cudaError_t e = o0.ensure(k_gemm_nt<0>, smem);
    if (e == cudaSuccess) e = o1.ensure(k_gemm_nt<1>, smem);
    if (e == cudaSuccess) e = o2.ensure(k_gemm_nt<2>, smem);
    if (e == cudaSuccess) e = o3.ensure(k_gemm_nt<3>, smem);
    if (e != cudaSuccess) return e;
    {
        const long long tiles = (long long)mb * nbk / (mode == 1 || mode == 3 ? 2 : 1);
        if (chol_skip() & (tiles < 256 ? 4 : 8)) return cudaSuccess;
    }
    dim3 grid(mb, nbk);
    if (mode == 1) k_gemm_nt<1><<<grid, 128, smem, st>>>(C, ldc, A, lda, B, ldb, kb * kBlk);
    else if (mode == 2) k_gemm_nt<2><<<grid, 128, smem, st>>>(C, ldc, A, lda, B, ldb, kb * kBlk);
    else if (mode == 3) k_gemm_nt<3><<<grid, 128, smem, st>>>(C, ldc, A, lda, B, ldb, kb * kBlk);
    else k_gemm_nt<0><<<grid, 128, smem, st>>>(C, ldc, A, lda, B, ldb, kb * kBlk);
    ++*launches;
    return cudaGetLastError();
}


// X (mr x n blocks at block (r0, c0)) := X * L^-T with L the n x n block triangle at (c0, c0).
// Below five blocks the whole solve is one fused launch (k_trsm_fused).
static cudaError_t rtrsm(double* X, int ldx, double* A, int ld, const double* Linv, int r0, int mr, int c0, int n,
                         cudaStream_t st, int64_t* launches) {
    if (n <= TF_MAX || (n <= TF_MAX16 && mr <= 74)) {
        static SmemOptIn o64, o16;
        cudaError_t e0 = o64.ensure(k_trsm_fused<64>, kTrsmFusedSmem);
        if (e0 == cudaSuccess) e0 = o16.ensure(k_trsm_fused<16>, kTrsmFusedSmem16);
        if (e0 != cudaSuccess) return e0;
        if (chol_skip() & 2) return cudaSuccess;
        // few row blocks (deep recursion levels): 16-row strips give four times the CTAs
        if (mr <= 74) k_trsm_fused<16><<<mr * 4, 256, kTrsmFusedSmem16, st>>>(X, ldx, A, ld, Linv, r0, c0, n);
        else k_trsm_fused<64><<<mr, 256, kTrsmFusedSmem, st>>>(X, ldx, A, ld, Linv, r0, c0, n);
        ++*launches;
        return cudaGetLastError();
    }
    const int n1 = n / 2, n2 = n - n1;
    cudaError_t e = rtrsm(X, ldx, A, ld, Linv, r0, mr, c0, n1, st, launches);
    if (e != cudaSuccess) return e;
    // X2 -= X1 * L21'
    e = gemm_nt(AT(X, ldx, r0, c0 + n1), ldx, AT(X, ldx, r0, c0), ldx, AT(A, ld, c0 + n1, c0), ld, mr, n2, n1, 0, st,
                launches);
    if (e != cudaSuccess) return e;
    return rtrsm(X, ldx, A, ld, Linv, r0, mr, c0 + n1, n2, st, launches);
}

// Factor block range [b0, b0+n); the block with index aug_blk (if inside) is not factorised.
static cudaError_t rchol(double* A, int ld, double* Linv, int b0, int n, int aug_blk, int* info, cudaStream_t st,
                         int64_t* launches) {
    if (n == 1) {
        if (b0 == aug_blk) return cudaSuccess;
        constexpr size_t psmem = (2 * kBlk * (kBlk + 1) + 3 * 16 * 17 + kBlk) * sizeof(double);
        static SmemOptIn opt;
        cudaError_t e0 = opt.ensure(k_potrf64_inv, psmem);
        if (e0 != cudaSuccess) return e0;
        if (chol_skip() & 1) return cudaSuccess;
        k_potrf64_inv<<<1, 256, psmem, st>>>(AT(A, ld, b0, b0), ld, LINV(Linv, b0), info);
        ++*launches;
        return cudaGetLastError();
    }
    const int n1 = n / 2, n2 = n - n1;
    cudaError_t e = rchol(A, ld, Linv, b0, n1, aug_blk, info, st, launches);
    if (e != cudaSuccess) return e;
    e = rtrsm(A, ld, A, ld, Linv, b0 + n1, n2, b0, n1, st, launches);
    if (e != cudaSuccess) return e;
    e = gemm_nt(AT(A, ld, b0 + n1, b0 + n1), ld, AT(A, ld, b0 + n1, b0), ld, AT(A, ld, b0 + n1, b0), ld, n2, n2, n1,
                1, st, launches);
    if (e != cudaSuccess) return e;
    return rchol(A, ld, Linv, b0 + n1, n2, aug_blk, info, st, launches);
}

// DIAG task of the task graph: blocks [b0, b0 + n) on the diagonal.  Supertiles of a plan (<= kColumnMax blocks) go
// column by column through k_chol_column, one launch per block; larger ones keep the recursive form, whose big
// products use many CTAs.  FEBA_CHOL_COLUMNS=0: always recursive.
constexpr int kColumnMax = 8;
static cudaError_t chol_diag_tile(double* A, int ld, double* Linv, int b0, int n, int* info, cudaStream_t st,
                                  int64_t* launches) {
    static const bool use_columns = !(std::getenv("FEBA_CHOL_COLUMNS") && std::atoi(std::getenv("FEBA_CHOL_COLUMNS")) == 0);
    if (!use_columns || n > kColumnMax || (chol_skip() & 1)) return rchol(A, ld, Linv, b0, n, -1, info, st, launches);
    static SmemOptIn opt;
    cudaError_t e0 = opt.ensure(k_chol_column, kColumnSmem);
    if (e0 != cudaSuccess) return e0;
    for (int j = 0; j < n; ++j) {
        k_chol_column<<<n - j, 256, kColumnSmem, st>>>(A, ld, Linv, b0, j, info, 0);
        ++*launches;
    }
    return cudaGetLastError();
}

// Small systems (<= kAugColumnsMax blocks): nb + 1 launches instead of ~3 nb -- column j of the factor with the
// augmented block row as its last row block, then the Schur complement T of the augmented block.
constexpr int kAugColumnsMax = 32;
cudaError_t chol_augmented(double* A, int ld, int nb, double* Linv, int* info, cudaStream_t st, int64_t* launches) {
    static const bool use_columns = !(std::getenv("FEBA_CHOL_COLUMNS") && std::atoi(std::getenv("FEBA_CHOL_COLUMNS")) == 0);
    if (!use_columns || nb > kAugColumnsMax || (chol_skip() & 1)) return rchol(A, ld, Linv, 0, nb + 1, nb, info, st, launches);
    static SmemOptIn opt;
    cudaError_t e0 = opt.ensure(k_chol_column, kColumnSmem);
    if (e0 != cudaSuccess) return e0;
    for (int j = 0; j < nb; ++j) {
        k_chol_column<<<nb + 1 - j, 256, kColumnSmem, st>>>(A, ld, Linv, 0, j, info, 0);
        ++*launches;
    }
    k_chol_column<<<1, 256, kColumnSmem, st>>>(A, ld, Linv, 0, nb, info, 1);
    ++*launches;
    return cudaGetLastError();
}

// The same for a batch of systems of one shape: pointer arrays on the device, one launch per column for all.
cudaError_t chol_augmented_batched(double* const* As, int ld, int nb, double* const* Linvs, int* const* infos, int n_sys,
                                   cudaStream_t st, int64_t* launches) {
    static SmemOptIn opt;
    cudaError_t e0 = opt.ensure(k_chol_column_batched, kColumnSmem);
    if (e0 != cudaSuccess) return e0;
    for (int j = 0; j < nb; ++j) {
        k_chol_column_batched<<<dim3(nb + 1 - j, n_sys), 256, kColumnSmem, st>>>(As, ld, Linvs, 0, j, infos, 0);
        ++*launches;
    }
    k_chol_column_batched<<<dim3(1, n_sys), 256, kColumnSmem, st>>>(As, ld, Linvs, 0, nb, infos, 1);
    ++*launches;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Task-graph (right-looking, supertile) form of the same factorisation.  The recursive form above is
// a single chain: its ~190 diagonal 64x64 factorisations, leaves and small products (~12 ms at
// u_c = 12,010) run with most of the GPU idle.  Here the matrix is cut into supertiles of T 64-blocks;
//   DIAG(k)      : recursive factorisation of supertile (k,k)                    (the chain above, small)
//   TRSM(i,k)    : X(i,k) := X(i,k) L_kk^-T                     i > k (incl. the augmented block row)
//   UPDATE(i,j,k): A(i,j) -= X(i,k) X(j,k)'                     k < j <= i
// are enqueued on a pool of streams with event dependencies per tile (reads after the tile's last
// write; every write waits for the previous writer), so that DIAG / TRSM of step k+1 overlap the
// bulk updates of step k.  Stream 0 has high priority and carries the critical path
// (DIAG(k), TRSM(k+1,k), UPDATE(k+1,k+1,k)).  Captured into the iteration's CUDA graph like everything
// else, the events become graph edges.
// FEBA_UPD1_BULK=1: the first update of the next diagonal tile runs with the bulk updates (same CTA
// footprint, so stream priority is effective there) instead of on the chain stream.
static bool upd1_bulk() {
    static const bool v = std::getenv("FEBA_UPD1_BULK") != nullptr;
    return v;
}

// Task graph over the supertiles of a plan (feba_order.h): tile t covers the 64-blocks [b0[t], b0[t+1]); tile NT is
// the augmented block row.  nz (optional): (NT+1) x (NT+1) row-major pattern of the structurally non-zero tiles
// INCLUDING fill; TRSM / UPDATE tasks on zero tiles are not issued.  chain (optional): stream slot of every
// tile's panel chain -- tiles of independent subtrees of a nested-dissection order get different slots so that
// their DIAG / TRSM chains run side by side; slot 0 is the high-priority stream (root node).  Columns k in
// [k_begin, k_end) whose owner is this rank (or shared, owner < 0) are eliminated: a group of GPUs calls it once
// for its own subtrees, sums the shared trailing part over the ranks, and once more for the shared top.
// The same tasks while `main` is being CAPTURED into a CUDA graph: every task is issued on `main` itself after
// the stream's capture dependencies have been set to exactly the graph nodes that last wrote the tiles the task
// reads or writes (cudaStreamUpdateCaptureDependencies), and the nodes the task leaves behind become the new
// last writers of its output tile.  The captured graph then carries the TRUE dependencies of the factorisation
// and nothing else: with a pool of streams, two unrelated tasks that share a stream are ordered by it (measured on
// BASELINE configs[3], nested-dissection plan, 1,689 kernel nodes: 8 / 16 / 32 streams -> 9.9 / 9.1 / 7.0 ms).
// FEBA_CHAIN_PROF=1: time stamps (external event nodes of the captured graph) when the DIAG task of every supertile
// may start and when it ends; chain_prof_report() prints them after a replay.  Diagnostic only.
struct ChainProf {
    std::vector<cudaEvent_t> ev;       // 2 per recorded tile: start, end
    std::vector<int> tile, blocks, chain;
    cudaEvent_t origin = nullptr;
    bool printed = false;
};
static ChainProf g_chain;

void chain_prof_report() {
    if (g_chain.ev.empty() || g_chain.printed || !g_chain.origin) return;
    g_chain.printed = true;
    fprintf(stderr, "[feba chain prof] tile blocks slot | DIAG may start, DIAG ends (ms after the factorisation began)\n");
    for (size_t i = 0; i < g_chain.tile.size(); ++i) {
        float a = -1.f, b = -1.f;
        if (cudaEventElapsedTime(&a, g_chain.origin, g_chain.ev[2 * i]) != cudaSuccess) cudaGetLastError();
        if (cudaEventElapsedTime(&b, g_chain.origin, g_chain.ev[2 * i + 1]) != cudaSuccess) cudaGetLastError();
        fprintf(stderr, "[feba chain prof] %3d %2d %3d | %7.3f %7.3f\n", g_chain.tile[i], g_chain.blocks[i], g_chain.chain[i], a, b);
    }
}

static cudaError_t chol_tiles_captured(double* A, int ld, int nb, double* Linv, int* info, cudaStream_t main,
                                       int64_t* launches, const TileView& V, int k_begin, int k_end, int rank) {
    const int NT = V.NT;
    const int NR = NT + 1;
    auto on = [&](int i, int j) { return V.nz == nullptr || V.nz[(size_t)i * NR + j] != 0; };
    auto blk0 = [&](int t) { return t == NT ? nb : V.b0[t]; };
    auto nblk = [&](int t) { return t == NT ? 1 : V.b0[t + 1] - V.b0[t]; };
    auto mine = [&](int k) { return V.owner == nullptr || V.owner[k] < 0 || V.owner[k] == rank; };
    auto tid = [&](int i, int j) { return (size_t)i * NR + j; };
    typedef std::vector<cudaGraphNode_t> NodeSet;
    auto current = [&](NodeSet& out) -> cudaError_t {
        cudaStreamCaptureStatus st;
        const cudaGraphNode_t* deps = nullptr;
        size_t nd = 0;
        cudaError_t e = cudaStreamGetCaptureInfo(main, &st, nullptr, nullptr, &deps, &nd);
        if (e != cudaSuccess) return e;
        if (st != cudaStreamCaptureStatusActive) return cudaErrorInvalidValue;
        out.assign(deps, deps + nd);
        return cudaSuccess;
    };
    NodeSet base;
    cudaError_t e = current(base);
    if (e != cudaSuccess) return e;
    std::vector<NodeSet> last((size_t)NR * NR);       // empty: not written in this call -> whatever preceded it
    std::vector<char> written((size_t)NR * NR, 0);
    NodeSet scratch;
    auto begin_task = [&](std::initializer_list<size_t> tiles) -> cudaError_t {
        scratch.clear();
        for (size_t t : tiles) {
            const NodeSet& d = written[t] ? last[t] : base;
            scratch.insert(scratch.end(), d.begin(), d.end());
        }
        std::sort(scratch.begin(), scratch.end());
        scratch.erase(std::unique(scratch.begin(), scratch.end()), scratch.end());
        return cudaStreamUpdateCaptureDependencies(main, scratch.data(), scratch.size(), cudaStreamSetCaptureDependencies);
    };
    auto end_task = [&](size_t t) -> cudaError_t {
        written[t] = 1;
        return current(last[t]);
    };
#define DAG_CU(x)                 \
    do {                          \
        e = (x);                  \
        if (e != cudaSuccess) return e; \
    } while (0)
    static const bool prof_env = std::getenv("FEBA_CHAIN_PROF") != nullptr;
    const bool prof = prof_env && g_chain.ev.empty() == (k_begin == 0) && !g_chain.printed;
    auto stamp = [&](cudaEvent_t* out) -> cudaError_t {
        cudaError_t e2 = cudaEventCreate(out);
        if (e2 != cudaSuccess) return e2;
        return cudaEventRecordWithFlags(*out, main, cudaEventRecordExternal);
    };
    if (prof && k_begin == 0) DAG_CU(stamp(&g_chain.origin));
    for (int k = k_begin; k < k_end && k < NT; ++k) {
        if (!mine(k) || nblk(k) == 0) continue;
        DAG_CU(begin_task({tid(k, k)}));
        if (prof) {
            cudaEvent_t ev;
            DAG_CU(stamp(&ev));
            g_chain.ev.push_back(ev);
        }
        DAG_CU(chol_diag_tile(A, ld, Linv, blk0(k), nblk(k), info, main, launches));
        if (prof) {
            cudaEvent_t ev;
            DAG_CU(stamp(&ev));
            g_chain.ev.push_back(ev);
            g_chain.tile.push_back(k);
            g_chain.blocks.push_back(nblk(k));
            g_chain.chain.push_back(V.chain ? V.chain[k] : 0);
        }
        DAG_CU(end_task(tid(k, k)));
        for (int i = k + 1; i < NR; ++i) {
            if (!on(i, k) || nblk(i) == 0) continue;
            DAG_CU(begin_task({tid(k, k), tid(i, k)}));
            DAG_CU(rtrsm(A, ld, A, ld, Linv, blk0(i), nblk(i), blk0(k), nblk(k), main, launches));
            DAG_CU(end_task(tid(i, k)));
        }
        for (int i = k + 1; i < NR; ++i)
            for (int j = k + 1; j <= i; ++j) {
                if (!on(i, k) || !on(j, k) || nblk(i) == 0 || nblk(j) == 0) continue;
                if (j >= NT && !(j == NT && i == NT)) continue;     // augmented diagonal block: plain lower update
                DAG_CU(begin_task({tid(i, k), tid(j, k), tid(i, j)}));
                DAG_CU(gemm_nt(AT(A, ld, blk0(i), blk0(j)), ld, AT(A, ld, blk0(i), blk0(k)), ld,
                               AT(A, ld, blk0(j), blk0(k)), ld, nblk(i), nblk(j), nblk(k), i == j ? 1 : 0, main, launches));
                DAG_CU(end_task(tid(i, j)));
            }
    }
#undef DAG_CU
    // join: everything that follows depends on every task (and on what preceded the call)
    scratch = base;
    for (size_t t = 0; t < last.size(); ++t)
        if (written[t]) scratch.insert(scratch.end(), last[t].begin(), last[t].end());
    std::sort(scratch.begin(), scratch.end());
    scratch.erase(std::unique(scratch.begin(), scratch.end()), scratch.end());
    return cudaStreamUpdateCaptureDependencies(main, scratch.data(), scratch.size(), cudaStreamSetCaptureDependencies);
}

cudaError_t chol_tiles(double* A, int ld, int nb, double* Linv, int* info, const DagStreams& D, cudaStream_t main,
                       int64_t* launches, const TileView& V, int k_begin, int k_end, int rank) {
    {
        static const bool exact = !(std::getenv("FEBA_EXACT_DEPS") && std::atoi(std::getenv("FEBA_EXACT_DEPS")) == 0);
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (exact && cudaStreamIsCapturing(main, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusActive)
            return chol_tiles_captured(A, ld, nb, Linv, info, main, launches, V, k_begin, k_end, rank);
    }
    const int NT = V.NT;
    const int NR = NT + 1;
    auto on = [&](int i, int j) { return V.nz == nullptr || V.nz[(size_t)i * NR + j] != 0; };
    auto blk0 = [&](int t) { return t == NT ? nb : V.b0[t]; };
    auto nblk = [&](int t) { return t == NT ? 1 : V.b0[t + 1] - V.b0[t]; };
    auto mine = [&](int k) { return V.owner == nullptr || V.owner[k] < 0 || V.owner[k] == rank; };
    if (NR * NR > D.n_events || D.n_streams < 2) return cudaErrorInvalidValue;
    std::vector<int> last((size_t)NR * NR, -1);       // stream of the last writer of tile (i,j), -1: none yet
    auto tid = [&](int i, int j) { return i * NR + j; };
    const int NB = D.n_streams - 1;                   // streams 1..NB: bulk and the chains of the subtrees
    auto stream_of = [&](int i, int j) { return 1 + (i * 3 + j * 7) % NB; };
    auto chain_of = [&](int k) {
        if (V.chain == nullptr) return 0;
        return V.chain[k] <= 0 ? 0 : 1 + (V.chain[k] - 1) % NB;
    };
    cudaError_t e = cudaEventRecord(D.fork, main);
    if (e != cudaSuccess) return e;
    for (int s = 0; s < D.n_streams; ++s) {
        e = cudaStreamWaitEvent(D.streams[s], D.fork, 0);
        if (e != cudaSuccess) return e;
    }
    auto acquire = [&](int sid, int i, int j) -> cudaError_t {
        const int w = last[(size_t)tid(i, j)];
        if (w >= 0 && w != sid) return cudaStreamWaitEvent(D.streams[sid], D.events[tid(i, j)], 0);
        return cudaSuccess;
    };
    auto release = [&](int sid, int i, int j) -> cudaError_t {
        last[(size_t)tid(i, j)] = sid;
        return cudaEventRecord(D.events[tid(i, j)], D.streams[sid]);
    };
#define DAG_CU(x)                 \
    do {                          \
        e = (x);                  \
        if (e != cudaSuccess) return e; \
    } while (0)
    for (int k = k_begin; k < k_end && k < NT; ++k) {
        if (!mine(k) || nblk(k) == 0) continue;
        const int ck = chain_of(k);
        {   // DIAG(k)
            DAG_CU(acquire(ck, k, k));
            DAG_CU(chol_diag_tile(A, ld, Linv, blk0(k), nblk(k), info, D.streams[ck], launches));
            DAG_CU(release(ck, k, k));
        }
        int first_row = -1;                           // first coupled tile below k: the next link of the chain
        for (int i = k + 1; i < NR; ++i) {
            if (!on(i, k) || nblk(i) == 0) continue;
            if (first_row < 0) first_row = i;
            const int sid = (i == first_row) ? ck : stream_of(i, k);
            DAG_CU(acquire(sid, k, k));
            DAG_CU(acquire(sid, i, k));
            DAG_CU(rtrsm(A, ld, A, ld, Linv, blk0(i), nblk(i), blk0(k), nblk(k), D.streams[sid], launches));
            DAG_CU(release(sid, i, k));
        }
        for (int i = k + 1; i < NR; ++i)
            for (int j = k + 1; j <= i; ++j) {
                if (!on(i, k) || !on(j, k) || nblk(i) == 0 || nblk(j) == 0) continue;
                if (j == NT && i == NT) {
                    // augmented diagonal block T (never factorised): plain lower update
                } else if (j >= NT) {
                    continue;
                }
                const int sid = (i == first_row && j == first_row && !upd1_bulk()) ? ck : stream_of(i, j);
                DAG_CU(acquire(sid, i, k));
                DAG_CU(acquire(sid, j, k));
                DAG_CU(acquire(sid, i, j));
                DAG_CU(gemm_nt(AT(A, ld, blk0(i), blk0(j)), ld, AT(A, ld, blk0(i), blk0(k)), ld,
                               AT(A, ld, blk0(j), blk0(k)), ld, nblk(i), nblk(j), nblk(k), i == j ? 1 : 0,
                               D.streams[sid], launches));
                DAG_CU(release(sid, i, j));
            }
    }
#undef DAG_CU
    for (int s = 0; s < D.n_streams; ++s) {
        e = cudaEventRecord(D.join[s], D.streams[s]);
        if (e != cudaSuccess) return e;
        e = cudaStreamWaitEvent(main, D.join[s], 0);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------
// The same task graph over a group of GPUs (one process each).  Supertile column k
// belongs to rank k % world (1-D block-cyclic): its owner runs DIAG(k) and TRSM(.,k) and broadcasts
// each finished tile; every rank applies UPDATE(i,j,k) to the columns j it owns (the 64-row augmented
// diagonal block is updated by everybody -- it is tiny and all ranks need it).  Collectives need
// contiguous buffers and one issue order per communicator: all of them go through dist.stream in
// program order, a tile is packed into dist.staging by a strided device copy, broadcast, and unpacked
// on the receivers.  Streams 0..3 of the pool have high priority: 0 carries the critical path, 1..3
// the rest of the panel (TRSM and the look-ahead updates of the next panel); 4.. the bulk updates.
struct DistProf {
    static constexpr int kCols = 12;
    std::vector<cudaEvent_t> ev;
    std::vector<int> pm, join;
    int m_fork = -1, m_end = -1, NT = 0, T = 0, rank = 0, world = 1;
    bool printed = false;
};
static DistProf g_prof;

void dist_prof_report() {
    if (g_prof.ev.empty() || g_prof.printed) return;
    g_prof.printed = true;
    auto at = [&](int m) {
        float ms = -1.f;
        if (m >= 0 && cudaEventElapsedTime(&ms, g_prof.ev[g_prof.m_fork], g_prof.ev[m]) != cudaSuccess) {
            cudaGetLastError();
            ms = -2.f;
        }
        return ms;
    };
    const int r = g_prof.rank;
    fprintf(stderr, "[feba dist prof] rank %d/%d T=%d NT=%d total %.3f ms; stream ends:", r, g_prof.world, g_prof.T,
            g_prof.NT, at(g_prof.m_end));
    for (int m : g_prof.join) fprintf(stderr, " %.2f", at(m));
    fprintf(stderr, "\n[feba dist prof] rank k | diag_start diag_end trsm1_end trsm2_end trsmN_end | comm_in tile1_out "
                    "tile2_out col_out | upd1_end upd21_end last_upd_end (ms after fork)\n");
    for (int k = 0; k < g_prof.NT; ++k) {
        const int* m = &g_prof.pm[(size_t)k * DistProf::kCols];
        fprintf(stderr, "[feba dist prof] %d %2d | %7.3f %7.3f %7.3f %7.3f %7.3f | %7.3f %7.3f %7.3f %7.3f | %7.3f %7.3f %7.3f\n",
                r, k, at(m[0]), at(m[1]), at(m[2]), at(m[3]), at(m[4]), at(m[5]), at(m[6]), at(m[7]), at(m[8]), at(m[9]),
                at(m[10]), at(m[11]));
    }
}

cudaError_t chol_dag_dist(double* A, int ld, int nb, double* Linv, int* info, const DagStreams& D, DistCtx& dist,
                          cudaStream_t main, int64_t* launches) {
    const int T = D.tile_blocks;
    const int NT = (nb + T - 1) / T;
    const int NR = NT + 1;
    auto blk0 = [&](int t) { return t == NT ? nb : t * T; };
    auto nblk = [&](int t) { return t == NT ? 1 : (t == NT - 1 ? nb - T * (NT - 1) : T); };
    if (NR * NR > D.n_events || D.n_streams < 6 || !dist.comm) return cudaErrorInvalidValue;
    const int COMM = D.n_streams;                     // pseudo stream id of dist.stream
    auto S = [&](int sid) { return sid == COMM ? dist.stream : D.streams[sid]; };
    std::vector<int> last(NR * NR, -1);
    auto tid = [&](int i, int j) { return i * NR + j; };
    auto panel_stream = [&](int i) { return 1 + i % 3; };
    auto bulk_stream = [&](int i, int j) { return 4 + (i * 3 + j * 7) % (D.n_streams - 4); };
    auto owner = [&](int k) { return k % dist.world; };
    // FEBA_DIST_PROF=1: time stamps of the panel chain, the collectives and the updates, recorded as
    // external event nodes of the captured graph and printed by dist_prof_report() after a replay
    static const bool prof_env = std::getenv("FEBA_DIST_PROF") != nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(main, &cap);
    const bool prof = prof_env && cap == cudaStreamCaptureStatusActive && g_prof.ev.empty();
    auto mark = [&](int sid) -> int {
        if (!prof) return -1;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecordWithFlags(ev, sid < 0 ? main : S(sid), cudaEventRecordExternal);
        g_prof.ev.push_back(ev);
        return (int)g_prof.ev.size() - 1;
    };
    constexpr int PC = DistProf::kCols;
    std::vector<int> pm((size_t)NT * PC, -1);
    const int m_fork = mark(-1);
    cudaError_t e = cudaEventRecord(D.fork, main);
    if (e != cudaSuccess) return e;
    for (int s = 0; s <= COMM; ++s) {
        e = cudaStreamWaitEvent(S(s), D.fork, 0);
        if (e != cudaSuccess) return e;
    }
    auto acquire = [&](int sid, int i, int j) -> cudaError_t {
        const int w = last[tid(i, j)];
        if (w >= 0 && w != sid) return cudaStreamWaitEvent(S(sid), D.events[tid(i, j)], 0);
        return cudaSuccess;
    };
    auto release = [&](int sid, int i, int j) -> cudaError_t {
        last[tid(i, j)] = sid;
        return cudaEventRecord(D.events[tid(i, j)], S(sid));
    };
#define DAG_CU(x)                 \
    do {                          \
        e = (x);                  \
        if (e != cudaSuccess) return e; \
    } while (0)
    // finished tile (i,k): owner -> everybody
    auto share_tile = [&](int i, int k) -> cudaError_t {
        const size_t rows = (size_t)nblk(i) * kBlk, cols = (size_t)nblk(k) * kBlk;
        double* tile = AT(A, ld, blk0(i), blk0(k));
        const bool root = owner(k) == dist.rank;
        if (rows * cols > dist.staging_count) return cudaErrorInvalidValue;
        if (root) {
            DAG_CU(acquire(COMM, i, k));
            DAG_CU(cudaMemcpy2DAsync(dist.staging, rows * sizeof(double), tile, (size_t)ld * sizeof(double),
                                     rows * sizeof(double), cols, cudaMemcpyDeviceToDevice, dist.stream));
        }
        if (dist_bcast_f64(&dist, dist.staging, rows * cols, owner(k), dist.stream)) return cudaErrorUnknown;
        if (!root) {
            DAG_CU(cudaMemcpy2DAsync(tile, (size_t)ld * sizeof(double), dist.staging, rows * sizeof(double),
                                     rows * sizeof(double), cols, cudaMemcpyDeviceToDevice, dist.stream));
            DAG_CU(release(COMM, i, k));
        }
        return cudaSuccess;
    };
    auto share_diag = [&](int k) -> cudaError_t {
        DAG_CU(share_tile(k, k));
        if (dist_bcast_f64(&dist, LINV(Linv, blk0(k)), (size_t)nblk(k) * kBlk * kBlk, owner(k), dist.stream))
            return cudaErrorUnknown;
        return cudaSuccess;
    };
    static const int diag_order = std::getenv("FEBA_DIST_DIAG") ? std::atoi(std::getenv("FEBA_DIST_DIAG")) : 0;
    for (int k = 0; k < NT; ++k) {
        const bool mine = owner(k) == dist.rank;
        if (mine) {
            DAG_CU(acquire(0, k, k));
            pm[k * PC + 0] = mark(0);
            DAG_CU(rchol(A, ld, Linv, blk0(k), nblk(k), -1, info, D.streams[0], launches));
            DAG_CU(release(0, k, k));
            pm[k * PC + 1] = mark(0);
            for (int i = k + 1; i < NR; ++i) {
                const int sid = (i == k + 1) ? 0 : panel_stream(i);
                DAG_CU(acquire(sid, k, k));
                DAG_CU(acquire(sid, i, k));
                DAG_CU(rtrsm(A, ld, A, ld, Linv, blk0(i), nblk(i), blk0(k), nblk(k), D.streams[sid], launches));
                DAG_CU(release(sid, i, k));
                if (i == k + 1) pm[k * PC + 2] = mark(0);
                if (i == k + 2) pm[k * PC + 3] = mark(sid);
                if (i == NR - 1) pm[k * PC + 4] = mark(sid);
            }
        }
        // the tile the next panel waits for goes first; the diagonal tiles and the inverted diagonal
        // factors are only read by the backward solve and follow after the last panel
        for (int i = k + 1; i < NR; ++i) {
            if (i == k + 1) pm[k * PC + 5] = mark(COMM);
            DAG_CU(share_tile(i, k));
            if (i == k + 1) pm[k * PC + 6] = mark(COMM);
            if (i == k + 2) pm[k * PC + 7] = mark(COMM);
            if (i == k + 1 && diag_order == 1) DAG_CU(share_diag(k));
        }
        if (diag_order == 0) DAG_CU(share_diag(k));
        pm[k * PC + 8] = mark(COMM);
        int last_sid = -1;
        for (int j = k + 1; j < NR; ++j) {
            if (j < NT && owner(j) != dist.rank) continue;
            for (int i = j; i < NR; ++i) {
                if (j == NT && i != NT) continue;
                const int sid = (i == k + 1 && j == k + 1) ? (upd1_bulk() ? 1 : 0)
                                                           : (j == k + 1 && j < NT ? panel_stream(i) : bulk_stream(i, j));
                DAG_CU(acquire(sid, i, k));
                DAG_CU(acquire(sid, j, k));
                DAG_CU(acquire(sid, i, j));
                DAG_CU(gemm_nt(AT(A, ld, blk0(i), blk0(j)), ld, AT(A, ld, blk0(i), blk0(k)), ld,
                               AT(A, ld, blk0(j), blk0(k)), ld, nblk(i), nblk(j), nblk(k), i == j ? 1 : 0,
                               D.streams[sid], launches));
                DAG_CU(release(sid, i, j));
                if (i == k + 1 && j == k + 1) pm[k * PC + 9] = mark(sid);
                if (i == k + 2 && j == k + 1) pm[k * PC + 10] = mark(sid);
                last_sid = sid;
            }
        }
        if (last_sid >= 0) pm[k * PC + 11] = mark(last_sid);
    }
    if (diag_order == 2)
        for (int k = 0; k < NT; ++k) DAG_CU(share_diag(k));
#undef DAG_CU
    for (int s = 0; s <= COMM; ++s) {
        if (prof) g_prof.join.push_back(mark(s));
        e = cudaEventRecord(D.join[s], S(s));
        if (e != cudaSuccess) return e;
        e = cudaStreamWaitEvent(main, D.join[s], 0);
        if (e != cudaSuccess) return e;
    }
    // a failed pivot is only seen by the owner of that column
    if (dist_allreduce_max_i32(&dist, info, 1, main)) return cudaErrorUnknown;
    if (prof) {
        g_prof.m_fork = m_fork;
        g_prof.m_end = mark(-1);
        g_prof.pm = pm;
        g_prof.NT = NT;
        g_prof.T = T;
        g_prof.rank = dist.rank;
        g_prof.world = dist.world;
    }
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------
// Column form of the task graph, for eager (not captured) execution by one GPU or by a group.
// Per supertile column j the matrix is cut into three parts: D(j) the diagonal tile, S(j) the tile
// below it (the next panel's input) and R(j) the rest of the column down to the augmented block row,
// R(j) in up to three row chunks.  Step k:
//   chain stream (own SM partition when the device is split):   DIAG(k), TRSM of S(k), update of D(k+1)
//   panel streams (high priority):   TRSM of the chunks of R(k); look-ahead update of S(k+1), R(k+1)
//   bulk streams:   one tall lower-trapezoid product per remaining column j >= k+2 (rows j..end)
// so a step is ~10 large launches instead of ~100 tile launches, few enough to be issued eagerly while
// the GPU works (the captured tile graph loses its priorities: ready chain nodes queue behind nodes
// that wait for later tiles).  In a group, column k belongs to rank k % world: the owner factorises
// it and broadcasts S(k), the chunks of R(k), then D(k) and the inverted diagonal factors, through
// dist->stream in program order (packed into dist->staging: collectives need contiguous buffers);
// every rank updates the columns it owns, and the 64-row augmented diagonal block.
cudaError_t chol_cols(double* A, int ld, int nb, double* Linv, int* info, const DagStreams& D, DistCtx* dist,
                      cudaStream_t main, int64_t* launches) {
    const int T = D.tile_blocks;
    const int NT = (nb + T - 1) / T;
    const int NR = NT + 1;
    auto blk0 = [&](int t) { return t >= NT ? nb : t * T; };
    auto nblk = [&](int t) { return t == NT ? 1 : (t == NT - 1 ? nb - T * (NT - 1) : T); };
    const bool group = dist && dist->comm && dist->world > 1;
    const int world = group ? dist->world : 1, rank = group ? dist->rank : 0;
    if (NR * NR > D.n_events || D.n_streams < 5) return cudaErrorInvalidValue;
    const int CH = 0, COMM = D.n_streams;
    auto S = [&](int sid) { return sid == COMM ? dist->stream : D.streams[sid]; };
    auto bulk_stream = [&](int j) { return 3 + j % (D.n_streams - 3); };
    auto owner = [&](int k) { return k % world; };
    std::vector<int> last(NR * NR, -1);
    // resources -> event slots
    auto rD = [&](int j) { return j * NR + j; };
    auto rS = [&](int j) { return (j + 1) * NR + j; };
    struct Chunk { int a, b; };                       // supertile rows [a, b) of the rest of a column
    auto chunks = [&](int j) {
        std::vector<Chunk> v;
        const int first = j + 2, n = NR - first;
        if (n <= 0) return v;
        // chunk 0 is the single supertile row the next panel's S part depends on; the remainder in two
        v.push_back({first, first + 1});
        const int m = n - 1, nc = m < 2 ? m : 2;
        for (int c = 0; c < nc; ++c) v.push_back({first + 1 + m * c / nc, first + 1 + m * (c + 1) / nc});
        return v;
    };
    auto rC = [&](int j, const Chunk& c) { return c.a * NR + j; };
    cudaError_t e = cudaEventRecord(D.fork, main);
    if (e != cudaSuccess) return e;
    const int n_used = group ? COMM : COMM - 1;
    for (int s = 0; s <= n_used; ++s) {
        e = cudaStreamWaitEvent(S(s), D.fork, 0);
        if (e != cudaSuccess) return e;
    }
    auto acquire = [&](int sid, int res) -> cudaError_t {
        const int w = last[res];
        if (w >= 0 && w != sid) return cudaStreamWaitEvent(S(sid), D.events[res], 0);
        return cudaSuccess;
    };
    auto release = [&](int sid, int res) -> cudaError_t {
        last[res] = sid;
        return cudaEventRecord(D.events[res], S(sid));
    };
#define DAG_CU(x)                 \
    do {                          \
        e = (x);                  \
        if (e != cudaSuccess) return e; \
    } while (0)
    // finished rows [r0, r0 + mr) (64-blocks) of column k: owner -> everybody
    auto share = [&](int res, int r0, int mr, int k) -> cudaError_t {
        const size_t rows = (size_t)mr * kBlk, cols = (size_t)nblk(k) * kBlk;
        double* src = AT(A, ld, r0, blk0(k));
        const bool root = owner(k) == rank;
        if (rows * cols > dist->staging_count) return cudaErrorInvalidValue;
        if (root) {
            DAG_CU(acquire(COMM, res));
            DAG_CU(cudaMemcpy2DAsync(dist->staging, rows * sizeof(double), src, (size_t)ld * sizeof(double),
                                     rows * sizeof(double), cols, cudaMemcpyDeviceToDevice, dist->stream));
        }
        if (dist_bcast_f64(dist, dist->staging, rows * cols, owner(k), dist->stream)) return cudaErrorUnknown;
        if (!root) {
            DAG_CU(cudaMemcpy2DAsync(src, (size_t)ld * sizeof(double), dist->staging, rows * sizeof(double),
                                     rows * sizeof(double), cols, cudaMemcpyDeviceToDevice, dist->stream));
            DAG_CU(release(COMM, res));
        }
        return cudaSuccess;
    };
    for (int k = 0; k < NT; ++k) {
        const bool mine = owner(k) == rank;
        const std::vector<Chunk> ck = chunks(k);
        if (mine) {
            DAG_CU(acquire(CH, rD(k)));
            DAG_CU(rchol(A, ld, Linv, blk0(k), nblk(k), -1, info, D.streams[CH], launches));
            DAG_CU(release(CH, rD(k)));
            DAG_CU(acquire(CH, rS(k)));
            DAG_CU(rtrsm(A, ld, A, ld, Linv, blk0(k + 1), nblk(k + 1), blk0(k), nblk(k), D.streams[CH], launches));
            DAG_CU(release(CH, rS(k)));
            for (size_t c = 0; c < ck.size(); ++c) {
                const int sid = c == 0 ? CH : 1 + (int)(c & 1);
                const int r0 = blk0(ck[c].a), mr = (ck[c].b >= NR ? nb + 1 : blk0(ck[c].b)) - r0;
                DAG_CU(acquire(sid, rD(k)));
                DAG_CU(acquire(sid, rC(k, ck[c])));
                DAG_CU(rtrsm(A, ld, A, ld, Linv, r0, mr, blk0(k), nblk(k), D.streams[sid], launches));
                DAG_CU(release(sid, rC(k, ck[c])));
            }
        }
        if (group) {
            DAG_CU(share(rS(k), blk0(k + 1), nblk(k + 1), k));
            for (const Chunk& c : ck) {
                const int r0 = blk0(c.a), mr = (c.b >= NR ? nb + 1 : blk0(c.b)) - r0;
                DAG_CU(share(rC(k, c), r0, mr, k));
            }
            DAG_CU(share(rD(k), blk0(k), nblk(k), k));
            if (dist_bcast_f64(dist, LINV(Linv, blk0(k)), (size_t)nblk(k) * kBlk * kBlk, owner(k), dist->stream))
                return cudaErrorUnknown;
        }
        // D(k+1) -= S(k) S(k)'   (k+1 == NT: the augmented diagonal block, kept by every rank)
        if (k + 1 == NT || owner(k + 1) == rank) {
            DAG_CU(acquire(CH, rS(k)));
            DAG_CU(acquire(CH, rD(k + 1)));
            DAG_CU(gemm_nt(AT(A, ld, blk0(k + 1), blk0(k + 1)), ld, AT(A, ld, blk0(k + 1), blk0(k)), ld,
                           AT(A, ld, blk0(k + 1), blk0(k)), ld, nblk(k + 1), nblk(k + 1), nblk(k), 1, D.streams[CH],
                           launches));
            DAG_CU(release(CH, rD(k + 1)));
        }
        if (ck.empty()) continue;
        const int r2 = blk0(k + 2);                         // first row below S(k)
        // look-ahead on column k+1: its S part (one tile, needs only chunk 0 of R(k)), and the rest
        // -= R(k)[rows k+3..] S(k)', on the two panel streams
        if (k + 1 < NT && owner(k + 1) == rank) {
            const int ls = 2;       // not the chain stream: DIAG(k+1) must not queue behind the wait for chunk 0
            DAG_CU(acquire(ls, rS(k)));
            DAG_CU(acquire(ls, rC(k, ck[0])));
            DAG_CU(acquire(ls, rS(k + 1)));
            DAG_CU(gemm_nt(AT(A, ld, r2, blk0(k + 1)), ld, AT(A, ld, r2, blk0(k)), ld, AT(A, ld, blk0(k + 1), blk0(k)), ld,
                           nblk(k + 2), nblk(k + 1), nblk(k), 0, D.streams[ls], launches));
            DAG_CU(release(ls, rS(k + 1)));
            const std::vector<Chunk> cn = chunks(k + 1);
            if (!cn.empty()) {
                const int sid = 1;
                const int r3 = blk0(k + 3), m3 = nb + 1 - r3;
                DAG_CU(acquire(sid, rS(k)));
                for (size_t c = 1; c < ck.size(); ++c) DAG_CU(acquire(sid, rC(k, ck[c])));
                for (const Chunk& c : cn) DAG_CU(acquire(sid, rC(k + 1, c)));
                DAG_CU(gemm_nt(AT(A, ld, r3, blk0(k + 1)), ld, AT(A, ld, r3, blk0(k)), ld, AT(A, ld, blk0(k + 1), blk0(k)),
                               ld, m3, nblk(k + 1), nblk(k), 0, D.streams[sid], launches));
                for (const Chunk& c : cn) DAG_CU(release(sid, rC(k + 1, c)));
            }
        }
        // bulk: column j (rows j..end, lower trapezoid) -= R(k)[rows j..] R(k)[rows of j]'
        for (int j = k + 2; j < NT; ++j) {
            if (owner(j) != rank) continue;
            const int sid = bulk_stream(j);
            const std::vector<Chunk> cj = chunks(j);
            for (const Chunk& c : ck) DAG_CU(acquire(sid, rC(k, c)));
            DAG_CU(acquire(sid, rD(j)));
            DAG_CU(acquire(sid, rS(j)));
            for (const Chunk& c : cj) DAG_CU(acquire(sid, rC(j, c)));
            DAG_CU(gemm_nt(AT(A, ld, blk0(j), blk0(j)), ld, AT(A, ld, blk0(j), blk0(k)), ld, AT(A, ld, blk0(j), blk0(k)),
                           ld, nb + 1 - blk0(j), nblk(j), nblk(k), 1, D.streams[sid], launches));
            DAG_CU(release(sid, rD(j)));
            DAG_CU(release(sid, rS(j)));
            for (const Chunk& c : cj) DAG_CU(release(sid, rC(j, c)));
        }
        {   // augmented diagonal block -= (augmented row of R(k)) (...)'
            const int sid = bulk_stream(NT);
            DAG_CU(acquire(sid, rC(k, ck.back())));
            DAG_CU(acquire(sid, rD(NT)));
            DAG_CU(gemm_nt(AT(A, ld, nb, nb), ld, AT(A, ld, nb, blk0(k)), ld, AT(A, ld, nb, blk0(k)), ld, 1, 1, nblk(k), 1,
                           D.streams[sid], launches));
            DAG_CU(release(sid, rD(NT)));
        }
    }
#undef DAG_CU
    for (int s = 0; s <= n_used; ++s) {
        e = cudaEventRecord(D.join[s], S(s));
        if (e != cudaSuccess) return e;
        e = cudaStreamWaitEvent(main, D.join[s], 0);
        if (e != cudaSuccess) return e;
    }
    // a failed pivot is only seen by the owner of that column
    if (group && dist_allreduce_max_i32(dist, info, 1, main)) return cudaErrorUnknown;
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------
// Border system (7x7) from the Schur complement T = -B' M^-1 B, B = [g Gc]:
//   (Gc' M^-1 Gc) k = -Gc' M^-1 g   <=>   T(1:,1:) k = -T(1:,0).     work[0..6] = k.
__global__ void k_border_solve(const double* __restrict__ A, int ld, int n_pad, double* __restrict__ work,
                               int* __restrict__ info) {
    if (threadIdx.x != 0) return;
    double Q[7][8];
    for (int i = 0; i < 7; ++i) {
        for (int j = 0; j < 7; ++j) {
            const int r = i > j ? i : j, c = i > j ? j : i;      // lower triangle holds T
            Q[i][j] = A[(size_t)(n_pad + 1 + r) + (size_t)ld * (n_pad + 1 + c)];
        }
        Q[i][7] = -A[(size_t)(n_pad + 1 + i) + (size_t)ld * n_pad];
    }
    for (int c = 0; c < 7; ++c) {                                 // Gaussian elimination, partial pivoting
        int p = c;
        for (int r = c + 1; r < 7; ++r)
            if (fabs(Q[r][c]) > fabs(Q[p][c])) p = r;
        if (Q[p][c] == 0.0) { atomicExch(info, 2); return; }
        if (p != c)
            for (int j = 0; j < 8; ++j) { const double t = Q[c][j]; Q[c][j] = Q[p][j]; Q[p][j] = t; }
        for (int r = c + 1; r < 7; ++r) {
            const double f = Q[r][c] / Q[c][c];
            for (int j = c; j < 8; ++j) Q[r][j] -= f * Q[c][j];
        }
    }
    for (int r = 6; r >= 0; --r) {
        double t = Q[r][7];
        for (int j = r + 1; j < 7; ++j) t -= Q[r][j] * work[j];
        work[r] = t / Q[r][r];
    }
}

// Sparse-datum form (feba_sparse.h): B = [g G~ E], T = -B' M_s^-1 B (15 x 15, lower triangle stored);
// work[0..13] = coefficients of the augmented rows 1..14 (k, -t).
__global__ void k_border_solve_sparse(const double* __restrict__ A, int ld, int n_pad, double* __restrict__ work,
                                      int* __restrict__ info) {
    __shared__ double Ts[kSparseAugRows][kSparseAugRows];
    for (int e = threadIdx.x; e < kSparseAugRows * kSparseAugRows; e += blockDim.x) {     // all loads in flight together
        const int i = e / kSparseAugRows, j = e - i * kSparseAugRows;
        const int r = i > j ? i : j, c = i > j ? j : i;                                    // lower triangle holds T
        Ts[i][j] = A[(size_t)(n_pad + r) + (size_t)ld * (n_pad + c)];
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    double T[kSparseAugRows][kSparseAugRows];
    for (int i = 0; i < kSparseAugRows; ++i)
        for (int j = 0; j < kSparseAugRows; ++j) T[i][j] = Ts[i][j];
    double coef[2 * kDatumCols];
    if (!sparse_border_solve(T, coef)) {
        atomicExch(info, 2);
        return;
    }
    for (int m = 0; m < 2 * kDatumCols; ++m) work[m] = coef[m];
}

// y[j] = Y'(0,j) + sum_k kvec[k] Y'(1+k,j),  k < ncoef (0: no border, 7: dense datum, 14: sparse datum)
__global__ void k_combine(const double* __restrict__ A, int ld, int n_pad, int ncoef,
                          const double* __restrict__ kvec, double* __restrict__ y) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_pad) return;
    const double* col = A + (size_t)n_pad + (size_t)ld * j;
    double t = col[0];
    for (int k = 0; k < ncoef; ++k) t += kvec[k] * col[1 + k];
    y[j] = t;
}

// Backward substitution L' x = y, one launch per 64-block step k = nb-1 .. 0:
//   x_k = L_kk^-T y_k (every CTA recomputes it from the inverted diagonal factor -- 4096 FMAs -- so
//   no second launch is needed), CTA 0 stores it, then y_c -= sum_r L[k*64+r][c] x_k[r] for the
//   columns c < k*64 that can be non-zero in block row k, one warp per column.
// The columns come as up to kMaxSeg ranges [c0, c1): a dense system has one range [0, k*64); with a plan
// (feba_order.h) only the column tiles coupled with the row's tile are visited -- supertiles outside the pattern
// were never written and are exactly zero -- which also keeps the steps of independent subtrees of a
// nested-dissection order apart, so that they may run side by side.
constexpr int kMaxSeg = 24;
struct ColSegs {
    int n;
    int c0[kMaxSeg], c1[kMaxSeg];
};

__device__ __forceinline__ void backstep_body(const double* __restrict__ A, int ld, int k,
                                              const double* __restrict__ Linv_k, double* __restrict__ y,
                                              double* __restrict__ x_out, const ColSegs& segs) {
    __shared__ double sx[kBlk];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* yk = y + (size_t)k * kBlk;
    // x[j] = sum_i Linv[i][j] y[i]  (Linv^T y); 4 threads per j
    {
        const int j = tid >> 2, qq = tid & 3;
        double acc = 0.0;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int i = 4 * t + qq;
            acc += Linv_k[i + (size_t)kBlk * j] * yk[i];
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (qq == 0) sx[j] = acc;
    }
    __syncthreads();
    const double x0 = sx[lane], x1 = sx[lane + 32];
    int done = 0;                                   // columns of the earlier ranges
    for (int sgi = 0; sgi < segs.n; ++sgi) {
        const int c0 = segs.c0[sgi], len = segs.c1[sgi] - c0;
        // global column index over all ranges -> CTAs and warps as if the ranges were one
        int first = (blockIdx.x * 8 + warp) - done % (gridDim.x * 8);
        if (first < 0) first += gridDim.x * 8;
        for (int q = first; q < len; q += gridDim.x * 8) {
            const int c = c0 + q;
            const double* colp = A + (size_t)k * kBlk + (size_t)ld * c;
            double acc = colp[lane] * x0 + colp[lane + 32] * x1;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
            if (lane == 0) y[c] -= acc;
        }
        done += len;
    }
    // other CTAs may still be reading y_k, so the solution goes to a separate vector
    if (blockIdx.x == 0 && tid < kBlk) x_out[(size_t)k * kBlk + tid] = sx[tid];
}

__global__ void __launch_bounds__(256) k_backstep(const double* __restrict__ A, int ld, int k,
                                                  const double* __restrict__ Linv_k, double* __restrict__ y,
                                                  double* __restrict__ x_out, ColSegs segs) {
    backstep_body(A, ld, k, Linv_k, y, x_out, segs);
}

// blockIdx.y = system of a batch of equal shape (feba_batch)
__global__ void __launch_bounds__(256) k_backstep_batched(double* const* __restrict__ As, int ld, int k,
                                                          double* const* __restrict__ Linvs, double* const* __restrict__ ys,
                                                          double* const* __restrict__ xs, ColSegs segs) {
    backstep_body(As[blockIdx.y], ld, k, Linvs[blockIdx.y] + (size_t)k * kBlk * kBlk, ys[blockIdx.y], xs[blockIdx.y], segs);
}

// backward substitution of a batch of dense systems of one shape: nb launches for all of them
cudaError_t backsolve_batched(double* const* As, int ld, int nb, double* const* Linvs, double* const* ys, double* const* xs,
                              int n_sys, int sm_count, cudaStream_t st, int64_t* launches) {
    for (int k = nb - 1; k >= 0; --k) {
        ColSegs segs;
        segs.n = 1;
        segs.c0[0] = 0;
        segs.c1[0] = k * kBlk;
        int grid = (k * kBlk + 7) / 8;
        const int cap = (2 * sm_count + n_sys - 1) / n_sys > 1 ? (2 * sm_count + n_sys - 1) / n_sys : 1;
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        k_backstep_batched<<<dim3(grid, n_sys), 256, 0, st>>>(As, ld, k, Linvs, ys, xs, segs);
        ++*launches;
    }
    return cudaGetLastError();
}

// border solve + combination only (the first part of border_and_backsolve), for callers that batch the rest
cudaError_t border_and_combine(double* A, int ld, int nb, int inner, double* work, double* ywork, int* info,
                               cudaStream_t st, int64_t* launches, int sparse_datum) {
    const int n_pad = nb * kBlk;
    if (inner) {
        if (sparse_datum) k_border_solve_sparse<<<1, 32, 0, st>>>(A, ld, n_pad, work, info);
        else k_border_solve<<<1, 32, 0, st>>>(A, ld, n_pad, work, info);
        ++*launches;
    }
    const int ncoef = inner ? (sparse_datum ? 2 * kDatumCols : kDatumCols) : 0;
    k_combine<<<(n_pad + 255) / 256, 256, 0, st>>>(A, ld, n_pad, ncoef, work, ywork);
    ++*launches;
    return cudaGetLastError();
}

// block_owner (host, nb entries, optional; group runs): blocks of subtrees owned by other ranks are skipped --
// their rows of `sol` stay zero and come from the owners through the sum over the ranks that follows.
// V (optional): the plan's tiles and pattern.  While `st` is being captured into a CUDA graph the steps get their
// true dependencies (as chol_tiles_captured): a step waits for the last writers of y in its own tile and in the
// column tiles it updates.
cudaError_t border_and_backsolve(double* A, int ld, int nb, const double* Linv, int inner, double* work,
                                 double* ywork, double* sol, int* info, int sm_count, cudaStream_t st,
                                 int64_t* launches, int sparse_datum, const TileView* V,
                                 const int* block_owner, int rank) {
    const int n_pad = nb * kBlk;
    if (block_owner) {
        cudaError_t e0 = cudaMemsetAsync(sol, 0, (size_t)n_pad * sizeof(double), st);
        if (e0 != cudaSuccess) return e0;
    }
    if (inner) {
        if (sparse_datum) k_border_solve_sparse<<<1, 32, 0, st>>>(A, ld, n_pad, work, info);
        else k_border_solve<<<1, 32, 0, st>>>(A, ld, n_pad, work, info);
        ++*launches;
    }
    const int ncoef = inner ? (sparse_datum ? 2 * kDatumCols : kDatumCols) : 0;
    k_combine<<<(n_pad + 255) / 256, 256, 0, st>>>(A, ld, n_pad, ncoef, work, ywork);
    ++*launches;
    const bool masked = V && V->nz;
    const int NT = V ? V->NT : 1, NR = NT + 1;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    static const bool exact_env = !(std::getenv("FEBA_EXACT_DEPS") && std::atoi(std::getenv("FEBA_EXACT_DEPS")) == 0);
    const bool exact = masked && exact_env && cudaStreamIsCapturing(st, &cap) == cudaSuccess &&
                       cap == cudaStreamCaptureStatusActive;
    typedef std::vector<cudaGraphNode_t> NodeSet;
    NodeSet base, scratch;
    std::vector<NodeSet> last;
    std::vector<char> written;
    auto current = [&](NodeSet& out) -> cudaError_t {
        cudaStreamCaptureStatus s2;
        const cudaGraphNode_t* deps = nullptr;
        size_t nd = 0;
        cudaError_t e = cudaStreamGetCaptureInfo(st, &s2, nullptr, nullptr, &deps, &nd);
        if (e != cudaSuccess) return e;
        out.assign(deps, deps + nd);
        return cudaSuccess;
    };
    if (exact) {
        cudaError_t e = current(base);
        if (e != cudaSuccess) return e;
        last.resize((size_t)NT);
        written.assign((size_t)NT, 0);
    }
    std::vector<int> tiles;                               // column tiles a step updates (plus its own)
    int t_of_k = NT - 1;
    for (int k = nb - 1; k >= 0; --k) {
        if (block_owner && block_owner[k] >= 0 && block_owner[k] != rank) continue;
        ColSegs segs;
        segs.n = 0;
        tiles.clear();
        if (!masked) {
            segs.n = 1;
            segs.c0[0] = 0;
            segs.c1[0] = k * kBlk;
        } else {
            while (t_of_k > 0 && V->b0[t_of_k] > k) --t_of_k;
            const int t = t_of_k;
            auto add = [&](int c0, int c1) {
                if (c1 <= c0) return;
                if (segs.n > 0 && segs.c1[segs.n - 1] == c0) segs.c1[segs.n - 1] = c1;
                else if (segs.n < kMaxSeg) { segs.c0[segs.n] = c0; segs.c1[segs.n] = c1; ++segs.n; }
                else segs.n = kMaxSeg + 1;                  // too many ranges: one range over everything below
            };
            for (int j = 0; j < t && segs.n <= kMaxSeg; ++j)
                if (V->nz[(size_t)t * NR + j] && V->b0[j + 1] > V->b0[j]) {
                    add(V->b0[j] * kBlk, V->b0[j + 1] * kBlk);
                    tiles.push_back(j);
                }
            if (segs.n <= kMaxSeg) add(V->b0[t] * kBlk, k * kBlk);
            if (segs.n > kMaxSeg) {
                int first = 0;
                while (first < t && !V->nz[(size_t)t * NR + first]) ++first;
                segs.n = 1;
                segs.c0[0] = V->b0[first] * kBlk;
                segs.c1[0] = k * kBlk;
                tiles.clear();
                for (int j = first; j < t; ++j) tiles.push_back(j);
            }
            tiles.push_back(t);
        }
        int cols = 0;
        for (int q = 0; q < segs.n; ++q) cols += segs.c1[q] - segs.c0[q];
        int grid = (cols + 7) / 8;
        if (grid > 2 * sm_count) grid = 2 * sm_count;
        if (grid < 1) grid = 1;
        if (exact) {
            scratch.clear();
            for (int j : tiles) {
                const NodeSet& d = written[(size_t)j] ? last[(size_t)j] : base;
                scratch.insert(scratch.end(), d.begin(), d.end());
            }
            std::sort(scratch.begin(), scratch.end());
            scratch.erase(std::unique(scratch.begin(), scratch.end()), scratch.end());
            cudaError_t e = cudaStreamUpdateCaptureDependencies(st, scratch.data(), scratch.size(), cudaStreamSetCaptureDependencies);
            if (e != cudaSuccess) return e;
        }
        k_backstep<<<grid, 256, 0, st>>>(A, ld, k, LINV(Linv, k), ywork, sol, segs);
        ++*launches;
        if (exact) {
            NodeSet now;
            cudaError_t e = current(now);
            if (e != cudaSuccess) return e;
            for (int j : tiles) {
                last[(size_t)j] = now;
                written[(size_t)j] = 1;
            }
        }
    }
    if (exact) {
        scratch = base;
        for (size_t t = 0; t < last.size(); ++t)
            if (written[t]) scratch.insert(scratch.end(), last[t].begin(), last[t].end());
        std::sort(scratch.begin(), scratch.end());
        scratch.erase(std::unique(scratch.begin(), scratch.end()), scratch.end());
        cudaError_t e = cudaStreamUpdateCaptureDependencies(st, scratch.data(), scratch.size(), cudaStreamSetCaptureDependencies);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Covariance stage (main.m:432-444 keeps Cx = NG^-1(1:u,1:u), main.m:468-480 un-scale
// its diagonal, main.m:602 multiplies by sigma02).  From the factor M~ = L L' of the last iteration:
//   U = L^-T (upper; triangular solve on the identity, zero blocks pruned),  Q = M~^-1 = U U'
//   Y = M~^-1 G~ = U (L^-1 G~)   (the augmented rows 1..7 already hold (L^-1 G~)')
//   Qxx_cc = D [ Q - Y (G~' M~^-1 G~)^-1 Y' ] D      (top-left block of the bordered inverse; D = Jacobi scaling)
// Everything reuses the factorisation's kernels (k_trsm_fused, k_gemm_nt); cost ~ u_c^3 flop, once.
__global__ void k_set_identity(double* __restrict__ U, int n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < (size_t)n * n) U[i] = (i / n == i % n) ? 1.0 : 0.0;
}

// X rows [r0, r0+mr) x column blocks [c0, c0+n):  X := X L^-T for X = rows of the identity: rows at or
// beyond the last column block are zero and stay zero.
static cudaError_t rtrsm_identity(double* X, int ldx, double* A, int ld, const double* Linv, int r0, int mr, int c0,
                                  int n, cudaStream_t st, int64_t* launches) {
    if (mr > c0 + n - r0) mr = c0 + n - r0;
    if (mr <= 0) return cudaSuccess;
    if (n <= TF_MAX) return rtrsm(X, ldx, A, ld, Linv, r0, mr, c0, n, st, launches);
    const int n1 = n / 2, n2 = n - n1;
    cudaError_t e = rtrsm_identity(X, ldx, A, ld, Linv, r0, mr, c0, n1, st, launches);
    if (e != cudaSuccess) return e;
    int mr1 = mr;                                   // rows with a non-zero X1
    if (mr1 > c0 + n1 - r0) mr1 = c0 + n1 - r0;
    if (mr1 > 0) {
        e = gemm_nt(AT(X, ldx, r0, c0 + n1), ldx, AT(X, ldx, r0, c0), ldx, AT(A, ld, c0 + n1, c0), ld, mr1, n2, n1, 0,
                    st, launches);
        if (e != cudaSuccess) return e;
    }
    return rtrsm_identity(X, ldx, A, ld, Linv, r0, mr, c0 + n1, n2, st, launches);
}

// Y[i][c] = sum_{k >= i} U(i,k) Zg[k][c],  Zg[k][c] = A[(n_pad+1+c) + ld k]  (c = 0..6); Y: n_pad x 8
__global__ void k_cov_Y(const double* __restrict__ U, int n_pad, const double* __restrict__ A, int ld,
                        double* __restrict__ Y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int k = i; k < n_pad; ++k) {
        const double u = U[(size_t)i + (size_t)n_pad * k];
        const double* z = A + (size_t)(n_pad + 1) + (size_t)ld * k;
#pragma unroll
        for (int c = 0; c < 7; ++c) acc[c] += u * z[c];
    }
#pragma unroll
    for (int c = 0; c < 7; ++c) Y[8 * (size_t)i + c] = acc[c];
    Y[8 * (size_t)i + 7] = 0.0;
}

// T7inv = (G~' M~^-1 G~)^-1 from the Schur block T = -[g G]' M^-1 [g G] of the factorisation.
__global__ void k_cov_T7inv(const double* __restrict__ A, int ld, int n_pad, double* __restrict__ T7inv,
                            int* __restrict__ info) {
    if (threadIdx.x != 0) return;
    double Q[7][14];
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) {
            const int r = i > j ? i : j, c = i > j ? j : i;
            Q[i][j] = -A[(size_t)(n_pad + 1 + r) + (size_t)ld * (n_pad + 1 + c)];
            Q[i][7 + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < 7; ++c) {
        int p = c;
        for (int r = c + 1; r < 7; ++r)
            if (fabs(Q[r][c]) > fabs(Q[p][c])) p = r;
        if (Q[p][c] == 0.0) { atomicExch(info, 2); return; }
        if (p != c)
            for (int j = 0; j < 14; ++j) { const double t = Q[c][j]; Q[c][j] = Q[p][j]; Q[p][j] = t; }
        for (int r = 0; r < 7; ++r) {
            if (r == c) continue;
            const double f = Q[r][c] / Q[c][c];
            for (int j = c; j < 14; ++j) Q[r][j] -= f * Q[c][j];
        }
    }
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) T7inv[7 * i + j] = Q[i][7 + j] / Q[i][i];
}

cudaError_t chol_inverse(double* A, int ld, int nb, const double* Linv, int inner, double* U, double* Q, double* Y,
                         double* T7inv, int* info, cudaStream_t st, int64_t* launches) {
    const int n_pad = nb * kBlk;
    const size_t nn = (size_t)n_pad * n_pad;
    k_set_identity<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(U, n_pad);
    ++*launches;
    cudaError_t e = rtrsm_identity(U, n_pad, A, ld, Linv, 0, nb, 0, nb, st, launches);
    if (e != cudaSuccess) return e;
    e = gemm_nt(Q, n_pad, U, n_pad, U, n_pad, nb, nb, nb, 3, st, launches);
    if (e != cudaSuccess) return e;
    if (inner) {
        k_cov_Y<<<(n_pad + 127) / 128, 128, 0, st>>>(U, n_pad, A, ld, Y);
        k_cov_T7inv<<<1, 32, 0, st>>>(A, ld, n_pad, T7inv, info);
        *launches += 2;
    }
    return cudaGetLastError();
}

}  // namespace feba
