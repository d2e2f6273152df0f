// Host-callable launchers of the kernel translation units (internal header).
#pragma once
#include <cuda_runtime.h>

#include "feba_dev.h"

namespace feba {

// feba_kernels.cu
cudaError_t launch_tables(const DevProblem& P, const double* eop, const double* iop, const double* cam_box,
                          double* img_tab, double* cam_tab, cudaStream_t st);
cudaError_t launch_xhat_scatter(const DevProblem& P, int sm_count, const double* xhat, double* eop, double* iop,
                                const int* tie_pt, cudaStream_t st);
cudaError_t launch_xhat_gather(const DevProblem& P, int sm_count, double* xhat, const double* eop,
                               const double* iop, const int* tie_pt, cudaStream_t st);
cudaError_t launch_delta_gather(const DevProblem& P, int sm_count, double* delta, cudaStream_t st);
// G rows (inner constraints), padding diagonal, M = S + Gc Gc' and its Jacobi scaling d = diag(M)^-1/2
// blocks: (block row, block column) of the structurally non-zero 64x64 blocks of the lower triangle followed by the
// blocks of the augmented block row (row index = n_pad / 64), as listed by the handle from its plan
cudaError_t launch_border_prepare(const DevProblem& P, const double* eop, double* dg, cudaStream_t st, int64_t* launches);
// gwork: 64 * 56 + 64 doubles of scratch (Gram partials and the conditioning matrix of the inner-constraint rows)
constexpr int kGworkDoubles = 64 * 56 + 64;
cudaError_t launch_border_scale(const DevProblem& P, const double* dg, double* dvec, int* info, const int2* blocks,
                                int n_blocks, double* gwork, cudaStream_t st, int64_t* launches);
cudaError_t launch_ties_pack(int n_own, const int* own_ties, int64_t n_red, double* xhat, double* packed, bool unpack,
                             cudaStream_t st);
cudaError_t launch_ties_copy_owned(int n_own, const int* own_ties, const double* src, double* dst, cudaStream_t st);
cudaError_t launch_keep_own_rows(const DevProblem& P, double* vec, cudaStream_t st);
cudaError_t launch_keep_own_ties(int64_t n_tie, const unsigned char* tie_mine, double* xyz3, cudaStream_t st);
cudaError_t launch_clear_blocks(const DevProblem& P, const int2* blocks, int n_blocks, cudaStream_t st);
// feba_assemble.cu: schedule (once) and the three assembly passes (per iteration)
cudaError_t build_pair_schedule(DevProblem& P, const int* d_oseg, long long* n_pairs_out, void** keep_pairs,
                                void** keep_blocks, cudaStream_t st);
int assemble_warps(const DevProblem& P, int sm_count);
// chunk schedule on the device (feba_chunks.h; arrays owned by the handle)
struct ChunkDev {
    int n_chunks;
    const int* obs0;
    const int* img0;
    const int* slot_obs0;
    const unsigned short* slot_obs;
    const int* blk0;
    const int* bslot_pair0;
    const unsigned int* pairs;
    double* img_part;       // image slots x kImgPart: [0,21) diagonal block (packed lower), [21,27) rhs, [27, 27+6NC) H Je
    double* blk_part;       // block slots x 36
    // final sums
    const int* slot_img;
    const int* timg_ptr;
    const int* timg_slots;
    int n_tblk;
    const int* tblk_a;
    const int* tblk_b;
    const int* tblk_ptr;
    const int* tblk_slots;
};
// chunks == nullptr: the image-major form of round 1 (image pass + image-pair pass over the pair schedule; always
// used with several cameras AND camera unknowns)
// opt: object point (CNT row) of every observation (the direct camera sums run one thread per observation)
cudaError_t launch_assemble(const DevProblem& P, int sm_count, int* info, cudaStream_t st, int64_t* launches,
                            const int* opt, const ChunkDev* chunks = nullptr);
int cam_part_rows(const DevProblem& P, int sm_count);
int backsub_warps(const DevProblem& P, int sm_count);
cudaError_t launch_backsub(const DevProblem& P, int sm_count, cudaStream_t st);
cudaError_t launch_update_cam(const DevProblem& P, const double* sol, const double* dvec, double* dcam,
                              double* dcam_unscaled, double* eop, double* iop, double* out_sumabs, cudaStream_t st);
cudaError_t launch_sum_partials(const double* partial, int n, int stride, int offset, double* out,
                                cudaStream_t st);
int residual_blocks(const DevProblem& P, int sm_count);
cudaError_t launch_residuals(const DevProblem& P, int sm_count, const int* opt, const double* xyz_prev,
                             const double* iop_new, double* v_out, double* rsd_out, cudaStream_t st);

// feba_chol.cu -- dense factorisation of the reduced camera system.
// A: (nb+1)*kBlk square, column-major, leading dimension ld, lower triangle; the last block row is
// the augmented block (right-hand side g in row 0, G columns in rows 1..7) and is not factorised:
// on return A = [L 0; Y' T] with Y' = B' L^-T and T = -B' M^-1 B (Schur complement).
// Linv: nb blocks of kBlk x kBlk (column-major): inverses of the diagonal factors.
// info (device int) is set non-zero when a pivot is not positive.
cudaError_t chol_augmented(double* A, int ld, int nb, double* Linv, int* info, cudaStream_t st, int64_t* launches);
// Batches of systems of one shape (feba_batch): device arrays of pointers, one launch per column / step for all.
cudaError_t chol_augmented_batched(double* const* As, int ld, int nb, double* const* Linvs, int* const* infos, int n_sys,
                                   cudaStream_t st, int64_t* launches);
cudaError_t backsolve_batched(double* const* As, int ld, int nb, double* const* Linvs, double* const* ys, double* const* xs,
                              int n_sys, int sm_count, cudaStream_t st, int64_t* launches);
cudaError_t border_and_combine(double* A, int ld, int nb, int inner, double* work, double* ywork, int* info,
                               cudaStream_t st, int64_t* launches, int sparse_datum);
// Stream pool + events of the task-graph form (owned by the handle).
struct DagStreams {
    int tile_blocks = 8;            // supertile size in 64-blocks
    int n_streams = 0;              // streams[0] has high priority (critical path)
    cudaStream_t streams[32] = {};
    cudaEvent_t fork = nullptr;
    cudaEvent_t join[33] = {};         // [n_streams] belongs to the collective stream of a group run
    cudaEvent_t* events = nullptr;  // one per tile (row supertile, column supertile)
    int n_events = 0;
};
// Supertiles of a plan of the reduced system (feba_order.h) as the task graph needs them (host arrays).
struct TileView {
    int NT = 0;                           // tiles of the factorised part; tile NT is the augmented block row
    const int* b0 = nullptr;              // NT + 1 entries: first 64-block of tile t, b0[NT] = nb
    const unsigned char* nz = nullptr;    // (NT+1) x (NT+1) row-major pattern incl. fill; null = dense
    const int* chain = nullptr;           // per tile: stream slot of its panel chain (0 = critical path); null = 0
    const int* owner = nullptr;           // per tile: owning rank, -1 = every rank; null = everything local
};
cudaError_t chol_tiles(double* A, int ld, int nb, double* Linv, int* info, const DagStreams& D, cudaStream_t main,
                       int64_t* launches, const TileView& V, int k_begin, int k_end, int rank);

// feba_green.cu -- two disjoint SM partitions of one device (green contexts): `chain` for the panel
// chain of the factorisation, `bulk` for its trailing updates.  -1 when the driver cannot provide them.
struct GreenPair {
    void* chain = nullptr;
    void* bulk = nullptr;
    int chain_sms = 0, bulk_sms = 0;
};
int green_create(int device, int reserve_sms, GreenPair* out, char* err, size_t errlen);
int green_stream(void* gctx, int priority, cudaStream_t* out);
void green_destroy(GreenPair* g);

// feba_dist.cu -- one rank of a group of GPUs that factorise the reduced system together.
struct DistCtx {
    int rank = 0, world = 1;
    void* comm = nullptr;            // ncclComm_t
    cudaStream_t stream = nullptr;   // every collective of the factorisation is issued here, in program order
    double* staging = nullptr;       // one supertile, contiguous (collectives need contiguous buffers)
    size_t staging_count = 0;
    char err[256] = {0};
};
const char* dist_load_error();
int dist_unique_id(void* id128);
int dist_comm_init(DistCtx* D, int rank, int world, const void* id128);
void dist_comm_destroy(DistCtx* D);
int dist_bcast_f64(DistCtx* D, double* buf, size_t count, int root, cudaStream_t st);
int dist_reduce_f64(DistCtx* D, double* buf, size_t count, int root, cudaStream_t st);
int dist_allreduce_f64(DistCtx* D, double* buf, size_t count, cudaStream_t st);
int dist_allreduce_max_i32(DistCtx* D, int* buf, size_t count, cudaStream_t st);
int dist_group_start();
int dist_group_end(DistCtx* D);
// Column-cyclic distributed form of chol_dag: supertile column k belongs to rank k % world, which
// factorises it (DIAG, TRSM) and broadcasts the finished tiles; every rank applies the updates of its
// own columns.  On return every rank holds the whole factor.  cudaErrorUnknown: see dist.err.
cudaError_t chol_dag_dist(double* A, int ld, int nb, double* Linv, int* info, const DagStreams& D, DistCtx& dist,
                          cudaStream_t main, int64_t* launches);
// Column form (D / S / R parts per supertile column, ~10 large launches per step), for eager issue;
// dist == nullptr: one GPU.
cudaError_t chol_cols(double* A, int ld, int nb, double* Linv, int* info, const DagStreams& D, DistCtx* dist,
                      cudaStream_t main, int64_t* launches);
void dist_prof_report();   // FEBA_DIST_PROF=1: prints the time stamps of the last replay once
void chain_prof_report();  // FEBA_CHAIN_PROF=1: when the DIAG task of every supertile could start / ended (last replay), once
// ywork (n_pad) := combination of the augmented rows: y = Y'(0,:) + sum_k kvec[k] Y'(1+k,:) where kvec
// solves the 7x7 border system (inner != 0), else y = Y'(0,:).  Then sol := L^-T y.
// sparse_datum != 0: 14 coefficients from the border of the sparse-datum form (feba_sparse.h).
// V (optional): tiles / pattern of the plan -- block rows visit the coupled column tiles only.
// block_owner / rank (optional, group runs): blocks owned by other ranks are skipped, their rows of sol stay 0.
cudaError_t border_and_backsolve(double* A, int ld, int nb, const double* Linv, int inner, double* work,
                                 double* ywork, double* sol, int* info, int sm_count, cudaStream_t st,
                                 int64_t* launches, int sparse_datum = 0, const TileView* V = nullptr,
                                 const int* block_owner = nullptr, int rank = 0);

// Covariance stage: U = L^-T, Q = M~^-1 (n_pad x n_pad, lower valid), Y = M~^-1 G~ (n_pad x 8), T7inv 7x7.
cudaError_t chol_inverse(double* A, int ld, int nb, const double* Linv, int inner, double* U, double* Q, double* Y,
                         double* T7inv, int* info, cudaStream_t st, int64_t* launches);
// cofactor entries of the EOP/IOP part: diag (distortion entries un-scaled, main.m:468-480) and blocks
cudaError_t launch_cov_diag_cam(const DevProblem& P, const double* Q, const double* Y, const double* T7inv,
                                const double* dvec, double* out, cudaStream_t st);
cudaError_t launch_cov_points(const DevProblem& P, int sm_count, const double* Q, const double* Y, const double* T7inv,
                              const double* dvec, double* out, cudaStream_t st);
cudaError_t launch_cov_block(const DevProblem& P, const double* Q, const double* Y, const double* T7inv,
                             const double* dvec, const long long* idx, int k, double* out, cudaStream_t st);

}  // namespace feba
