// Device-side problem description shared by the kernel translation units (internal header).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace feba {

constexpr int kBlk = 64;          // dense linear-algebra tile; reduced system is padded to it
constexpr int kAugRows = 8;       // rows of the augmented block in use: [g ; G(:,1..7)]  (15 in the sparse-datum form)
constexpr int kRec1 = 18;         // per-observation record 1: Je (2x6), Z (2x3)
constexpr int kCamPart = 104;     // per-warp camera partial: packed NC(NC+1)/2 block + NC rhs (NC <= 13)

// Everything a kernel needs, passed by value (lives in the constant bank).
struct DevProblem {
    int64_t n_obs;
    int n_img, n_cam, n_pts, n_tie, n_seg;
    int type, NK, NC;             // NC = NK + 5 camera columns (xp yp c k1..kNK p1 p2)
    int ui, uc;                   // estimated unknowns per image / per camera (BuildAwG.m:24-25)
    int n_red;                    // u_c = ui*n_img + uc*n_cam  (unknowns of the reduced system, Buildxhat.m:5-15)
    int n_pad;                    // rows of the factorised matrix: u_c + padding, a multiple of kBlk
    int ld;                       // leading dimension of S  (= n_pad + kBlk: augmented rows)
    int off_cam;                  // ROW of the first camera unknown
    int ext_off_cam;              // ui*n_img: index of the first camera unknown in xhat (Buildxhat.m:52-106)
    // row order of the reduced system (feba_order.h): identity = Buildxhat order, or nested dissection
    const int* img_row;           // n_img: row of the first unknown of every image
    const int* row_ext;           // n_pad: index in the EOP/IOP part of xhat of every row, -1 = padding row
    // group of GPUs sharing one adjustment (feba_create_shard): owner rank of every row (-1: shared top row) or
    // null (single GPU); shared rows are initialised (padding diagonal, datum term, G rows) by rank 0 only, so
    // that the sum over the ranks counts them once
    const int* row_owner;
    int rank;
    int inner;                    // Inner_Constraints
    int ecol[6];                  // slot of EOP q inside the image block or -1
    int ccol[16];                 // slot of camera parameter q inside the camera block or -1
    double px, py;                // 1/sigma_x^2, 1/sigma_y^2                     (main.m:396-405)

    // observations, sorted by point (segment s = observations [seg_start[s], seg_start[s+1]))
    const double* ox;
    const double* oy;
    const int* oimg;
    const int* operm;             // sorted position -> PHO row
    const int* seg_start;
    const int* seg_pt;            // CNT row of segment s
    const int* img_cam;
    const int* pt_tie;

    // parameters at the linearisation point of the current / last iteration
    const double* img_tab;        // n_img x kImgStride
    const double* cam_tab;        // n_cam x kCamStride
    double* xyz;                  // n_pts x 3, current object coordinates (updated in place)
    double* xyz_prev;             // coordinates at the linearisation point of the last iteration

    double* S;                    // (n_pad + kBlk) x ld, column-major, lower triangle
    const double* dcam;           // scaled increment of the camera part (length n_pad)
    const double* dcam_unscaled;  // un-scaled increment (main.m:458-482)
    double* dpts;                 // increment of the tie points, n_tie x 3
    double* pt_rec;               // per tie point {L^-1 (6), L^-1 u_p (3), Fc (NC x 3)} of the point pass, stride
                                  // 9 + 3 NC; null: the back-substitution recomputes the Jacobians
    double* partial;              // per-warp partial sums (deterministic final reduction)
    // assembly schedule and per-observation records (feba_assemble.cu)
    double* rec1;                 // n_obs x kRec1
    double* rec2;                 // n_obs x (2 + 2 NC): r (2), H (NC x 2)
    const int* img_start;         // n_img + 1: record ranges of the images
    const int* ipos;              // point-major observation index -> position in the image-major records
    const int2* pairs;            // (a, b) record positions of observation pairs sharing a tie point, by block
    const int4* blocks;           // (image a, image b <= a, first pair, pairs)
    int n_blocks;
    double* cam_part;             // per-warp camera-camera partial sums
    int cam_rec;                  // 1: camera block and right-hand side from the records (k_cam_rec), the point pass adds nothing
    double* Gt;                   // inner-constraint rows, compact: (6 n_img) x 8 row-major (col 7 unused)
    // sparse-datum form (feba_sparse.h, FEBA_SPARSE=1): flags of the datum images, or null (dense M = S + G G')
    int aug_rows;                 // rows of the augmented block in use: kAugRows, or kSparseAugRows
    const unsigned char* datum;   // n_img flags
};

// Opt-in dynamic shared memory above 48 KB is a per-device function attribute: remember per device
// (bit i = device i configured) so that handles on several GPUs of one process all work.
// row of the reduced system of entry e (< n_red) of the EOP/IOP part of xhat
#ifdef __CUDACC__
__device__ __forceinline__ int row_of_ext(const DevProblem& P, int e) {
    if (e < P.ext_off_cam) {
        const int im = e / P.ui;
        return P.img_row[im] + (e - im * P.ui);
    }
    return P.off_cam + (e - P.ext_off_cam);
}
// this rank initialises row r (see row_owner)
__device__ __forceinline__ bool row_init_here(const DevProblem& P, int r) {
    if (P.row_owner == nullptr) return true;
    const int o = P.row_owner[r];
    return o == P.rank || (o < 0 && P.rank == 0);
}
// row r holds an EOP unknown of an image (not a camera unknown, not padding)
__device__ __forceinline__ bool is_image_row(const DevProblem& P, int r) {
    const int e = P.row_ext[r];
    return e >= 0 && e < P.ext_off_cam;
}
#endif

struct SmemOptIn {
    unsigned long long done = 0;
    template <typename F>
    cudaError_t ensure(F func, size_t bytes) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        const unsigned long long bit = 1ull << (dev & 63);
        if (done & bit) return cudaSuccess;
        e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e == cudaSuccess) done |= bit;
        return e;
    }
};

}  // namespace feba
