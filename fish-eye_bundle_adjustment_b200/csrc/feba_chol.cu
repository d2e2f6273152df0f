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
#include <cstdio>

#include "feba_dev.h"
#include "feba_kernels.h"

namespace feba {

// ------------------------------------------------------------------------------------------
// C (M x N) -= A (M x K) * B (N x K)'   -- column-major, all dimensions multiples of 64.
// CTA tile 64x64, 4 warps of 32x32, K tile 16, 3-stage cp.async pipeline, DMMA m8n8k4.
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

template <bool LOWER>
__global__ void __launch_bounds__(128) k_gemm_nt(double* __restrict__ C, int ldc, const double* __restrict__ A,
                                                 int lda, const double* __restrict__ B, int ldb, int K) {
    const int bx = blockIdx.x, by = blockIdx.y;
    if (LOWER && by > bx) return;
    extern __shared__ __align__(16) double gsm[];
    double(*As)[GK][GS] = reinterpret_cast<double(*)[GK][GS]>(gsm);
    double(*Bs)[GK][GS] = reinterpret_cast<double(*)[GK][GS]>(gsm + GSTAGES * GK * GS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 1, wn = warp >> 1;
    const double* Ag = A + (size_t)bx * GT;
    const double* Bg = B + (size_t)by * GT;
    const int nk = K / GK;

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
    double* Cg = C + (size_t)bx * GT + (size_t)ldc * by * GT;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = wm * 32 + i * 8 + lr;
            const int c = wn * 32 + j * 8 + 2 * lk;
            Cg[r + (size_t)ldc * c] -= acc[i][j][0];
            Cg[r + (size_t)ldc * (c + 1)] -= acc[i][j][1];
        }
}

// ------------------------------------------------------------------------------------------
// 64x64 Cholesky of a diagonal block (lower), one CTA.
__global__ void __launch_bounds__(256) k_potrf64(double* __restrict__ A, int ld, int* __restrict__ info) {
    __shared__ double s[kBlk][kBlk + 1];
    const int tid = threadIdx.x;
    for (int q = tid; q < kBlk * kBlk; q += 256) {
        const int r = q & 63, c = q >> 6;
        s[r][c] = (r >= c) ? A[r + (size_t)ld * c] : 0.0;
    }
    __syncthreads();
    for (int j = 0; j < kBlk; ++j) {
        const double d = s[j][j];
        if (!(d > 0.0)) {
            if (tid == 0) atomicExch(info, 1);
        }
        const double rd = 1.0 / sqrt(d);
        __syncthreads();
        if (tid >= j && tid < kBlk) s[tid][j] = (tid == j) ? sqrt(d) : s[tid][j] * rd;
        __syncthreads();
        // trailing update: s[r][c] -= s[r][j]*s[c][j], j < c <= r
        const int m = kBlk - 1 - j;
        for (int q = tid; q < m * m; q += 256) {
            const int r = j + 1 + q / m, c = j + 1 + q % m;
            if (c <= r) s[r][c] -= s[r][j] * s[c][j];
        }
        __syncthreads();
    }
    for (int q = tid; q < kBlk * kBlk; q += 256) {
        const int r = q & 63, c = q >> 6;
        if (r >= c) A[r + (size_t)ld * c] = s[r][c];
    }
}

// X (rows x 64) := X * L^-T, L the 64x64 lower-triangular diagonal block.  One thread per row.
__global__ void __launch_bounds__(64) k_trsm64(double* __restrict__ X, int ldx, const double* __restrict__ L,
                                               int ldl) {
    __shared__ double sl[kBlk][kBlk + 1];
    __shared__ double sinv[kBlk];
    const int tid = threadIdx.x;
    for (int q = tid; q < kBlk * kBlk; q += 64) {
        const int r = q & 63, c = q >> 6;
        sl[r][c] = L[r + (size_t)ldl * c];
    }
    __syncthreads();
    sinv[tid] = 1.0 / sl[tid][tid];
    __syncthreads();
    double* xr = X + (size_t)blockIdx.x * kBlk + tid;
    double x[kBlk];
#pragma unroll
    for (int c = 0; c < kBlk; ++c) x[c] = xr[(size_t)ldx * c];
#pragma unroll
    for (int j = 0; j < kBlk; ++j) {
        x[j] *= sinv[j];
#pragma unroll
        for (int k = j + 1; k < kBlk; ++k) x[k] -= x[j] * sl[k][j];
    }
#pragma unroll
    for (int c = 0; c < kBlk; ++c) xr[(size_t)ldx * c] = x[c];
}

static cudaError_t gemm_nt(double* C, int ldc, const double* A, int lda, const double* B, int ldb, int mb,
                           int nbk, int kb, bool lower, cudaStream_t st, int64_t* launches) {
    const size_t smem = 2 * GSTAGES * GK * GS * sizeof(double);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_nt<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_gemm_nt<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid(mb, nbk);
    if (lower) k_gemm_nt<true><<<grid, 128, smem, st>>>(C, ldc, A, lda, B, ldb, kb * kBlk);
    else k_gemm_nt<false><<<grid, 128, smem, st>>>(C, ldc, A, lda, B, ldb, kb * kBlk);
    ++*launches;
    return cudaGetLastError();
}

#define AT(A, ld, br, bc) ((A) + (size_t)(br) * kBlk + (size_t)(ld) * (bc) * kBlk)

// X (mr x n blocks at block (r0, c0)) := X * L^-T with L the n x n block triangle at (c0, c0).
static cudaError_t rtrsm(double* A, int ld, int r0, int mr, int c0, int n, cudaStream_t st, int64_t* launches) {
    if (n == 1) {
        k_trsm64<<<mr, 64, 0, st>>>(AT(A, ld, r0, c0), ld, AT(A, ld, c0, c0), ld);
        ++*launches;
        return cudaGetLastError();
    }
    const int n1 = n / 2, n2 = n - n1;
    cudaError_t e = rtrsm(A, ld, r0, mr, c0, n1, st, launches);
    if (e != cudaSuccess) return e;
    // X2 -= X1 * L21'
    e = gemm_nt(AT(A, ld, r0, c0 + n1), ld, AT(A, ld, r0, c0), ld, AT(A, ld, c0 + n1, c0), ld, mr, n2, n1, false, st,
                launches);
    if (e != cudaSuccess) return e;
    return rtrsm(A, ld, r0, mr, c0 + n1, n2, st, launches);
}

// Factor block range [b0, b0+n); the block with index aug_blk (if inside) is not factorised.
static cudaError_t rchol(double* A, int ld, int b0, int n, int aug_blk, int* info, cudaStream_t st,
                         int64_t* launches) {
    if (n == 1) {
        if (b0 == aug_blk) return cudaSuccess;
        k_potrf64<<<1, 256, 0, st>>>(AT(A, ld, b0, b0), ld, info);
        ++*launches;
        return cudaGetLastError();
    }
    const int n1 = n / 2, n2 = n - n1;
    cudaError_t e = rchol(A, ld, b0, n1, aug_blk, info, st, launches);
    if (e != cudaSuccess) return e;
    e = rtrsm(A, ld, b0 + n1, n2, b0, n1, st, launches);
    if (e != cudaSuccess) return e;
    e = gemm_nt(AT(A, ld, b0 + n1, b0 + n1), ld, AT(A, ld, b0 + n1, b0), ld, AT(A, ld, b0 + n1, b0), ld, n2, n2, n1,
                true, st, launches);
    if (e != cudaSuccess) return e;
    return rchol(A, ld, b0 + n1, n2, aug_blk, info, st, launches);
}

cudaError_t chol_augmented(double* A, int ld, int nb, int* info, cudaStream_t st, int64_t* launches) {
    return rchol(A, ld, 0, nb + 1, nb, info, st, launches);
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

// y[j] = Y'(0,j) + sum_k kvec[k] Y'(1+k,j)
__global__ void k_combine(const double* __restrict__ A, int ld, int n_pad, int inner,
                          const double* __restrict__ kvec, double* __restrict__ y) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_pad) return;
    const double* col = A + (size_t)n_pad + (size_t)ld * j;
    double t = col[0];
    if (inner) {
#pragma unroll
        for (int k = 0; k < 7; ++k) t += kvec[k] * col[1 + k];
    }
    y[j] = t;
}

// Backward substitution L' x = y, one 64-block at a time (right-looking):
//   x_k = L_kk^-T y_k ;  y_j -= L_kj' x_k  for j < k.
__global__ void __launch_bounds__(64) k_trsv_bwd64(const double* __restrict__ L, int ld, double* __restrict__ y) {
    __shared__ double sl[kBlk][kBlk + 1];
    __shared__ double sx[kBlk];
    const int tid = threadIdx.x;
    for (int q = tid; q < kBlk * kBlk; q += 64) {
        const int r = q & 63, c = q >> 6;
        sl[r][c] = L[r + (size_t)ld * c];
    }
    sx[tid] = y[tid];
    __syncthreads();
    for (int j = kBlk - 1; j >= 0; --j) {
        if (tid == j) sx[j] = sx[j] / sl[j][j];
        __syncthreads();
        if (tid < j) sx[tid] -= sl[j][tid] * sx[j];      // (L')[tid][j] = L[j][tid]
        __syncthreads();
    }
    y[tid] = sx[tid];
}

// y[c] -= sum_r L[k*64+r][c] * x[r] for the columns c < k*64; one warp per column.
__global__ void __launch_bounds__(256) k_gemv_bwd(const double* __restrict__ Lrow, int ld, int ncols,
                                                  const double* __restrict__ xk, double* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= ncols) return;
    const double* col = Lrow + (size_t)ld * c;
    double acc = col[lane] * xk[lane] + col[lane + 32] * xk[lane + 32];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) y[c] -= acc;
}

cudaError_t border_and_backsolve(double* A, int ld, int nb, int inner, double* work, double* sol, int* info,
                                 cudaStream_t st, int64_t* launches) {
    const int n_pad = nb * kBlk;
    if (inner) {
        k_border_solve<<<1, 32, 0, st>>>(A, ld, n_pad, work, info);
        ++*launches;
    }
    k_combine<<<(n_pad + 255) / 256, 256, 0, st>>>(A, ld, n_pad, inner, work, sol);
    ++*launches;
    for (int k = nb - 1; k >= 0; --k) {
        k_trsv_bwd64<<<1, 64, 0, st>>>(AT(A, ld, k, k), ld, sol + (size_t)k * kBlk);
        ++*launches;
        if (k > 0) {
            const int ncols = k * kBlk;
            k_gemv_bwd<<<(ncols + 7) / 8, 256, 0, st>>>(A + (size_t)k * kBlk, ld, ncols, sol + (size_t)k * kBlk, sol);
            ++*launches;
        }
    }
    return cudaGetLastError();
}

}  // namespace feba
