/*
 * feba.h -- C ABI of the B200-native Gauss-Newton hot path of the equidistant fish-eye
 * bundle adjustment (drop-in for the loop body of the reference's main.m).
 *
 * The reference (MATLAB) has no FFI for this path; its seams are the function calls
 *   main.m:388  [xhaterror,xhat,xhatnames] = Buildxhat(data,EXT,INT,TIE,CNT)
 *   main.m:416  [Awerror,A,w,G,dist_scaling] = BuildAwG(data,xhat)
 *   main.m:424-493  u=A'Pw, N=A'PA, (bordered) inverse, delta, un-scaling, update, sumabs
 *   main.m:569  v = A*delta + w
 *   main.m:571  RSD = BuildRSD(v,data,xhat)
 *   main.m:594-602  RMSx, RMSy, RMS, sigma02
 * Because A (n x u) and N (u x u) are never materialised on the device, the seam sits one level
 * above BuildAwG: feba_iterate() is one pass of the while-body of main.m:412-494.  The MEX
 * gateway (fish-eye_bundle_adjustment_b200/mex/feba_mex.c) and INTEGRATION.md show the
 * MATLAB-side binding.
 *
 * Conventions: every function returns an int status, 0 = ok, non-zero = error (the reference's
 * 0/1 `error` flags, BuildAwG.m:16, Buildxhat.m:3, main.m:23); feba_last_error() gives the text.
 * All pointers are HOST pointers owned by the caller unless a name says `dev`.  Indices are
 * 0-based (reference value minus one).  Calls are synchronous.  One handle must not be used
 * from two host threads at once; different handles are independent.  There is no CPU fallback:
 * without a CUDA device feba_create() fails.
 */
#ifndef FEBA_H
#define FEBA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FEBA_VERSION 200
#define FEBA_MAX_NK 8 /* largest Num_Radial_Distortions supported by the kernels */
/* Limits the reference does not have (BuildAwG.m:110-155): Num_Radial_Distortions <= FEBA_MAX_NK, and, when
 * camera parameters are estimated for more than one camera, an object point may be seen by images of at most
 * FEBA_MAX_CAMS_PER_POINT different cameras.  Both are checked by feba_create (FEBA_ERR_INVALID) before any
 * device state exists.  With several cameras AND camera unknowns the point pass adds the cross-camera terms
 * with FP64 atomics: results are then reproducible to rounding only, not bit for bit (one camera: bit-identical
 * reruns). */
#define FEBA_MAX_CAMS_PER_POINT 4

enum feba_status {
    FEBA_OK = 0,
    FEBA_ERR_INVALID = 1,    /* bad argument / unsupported settings combination        */
    FEBA_ERR_CUDA = 2,       /* CUDA runtime failure (text in feba_last_error)         */
    FEBA_ERR_NUMERIC = 3,    /* reduced normal matrix not positive definite, NaN, ...  */
    FEBA_ERR_STATE = 4       /* call order (e.g. residuals before any iterate)         */
};

/* data.settings (main.m:112-177) -- the subset the hot path depends on. */
typedef struct feba_settings {
    int32_t estimate_eop[6];    /* Estimate_Xc Yc Zc Omega Phi Kappa   (main.m:158-163)  */
    int32_t estimate_xp;        /* main.m:166 */
    int32_t estimate_yp;        /* main.m:167 */
    int32_t estimate_c;         /* main.m:165 */
    int32_t estimate_radial;    /* Estimate_Radial_Distortions          (main.m:169)     */
    int32_t num_radial;         /* Num_Radial_Distortions, clamped to >=1 (BuildAwG.m:18) */
    int32_t estimate_decent;    /* Estimate_Decentering_Distortions     (main.m:171)     */
    int32_t inner_constraints;  /* Inner_Constraints                    (main.m:156)     */
    int32_t type;               /* typeint 0..4: fisheye pinhole equisolid orthographic
                                   stereographic                        (BuildAwG.m:184-208) */
    int32_t iteration_cap;      /* Iteration_Cap   (main.m:153, used by feba_solve only) */
    int32_t plan;               /* row order of the reduced camera system: 0 = automatic (nested dissection when it
                                   pays, csrc/feba_order.h), -1 = the order of Buildxhat.m (dense factorisation;
                                   needed by the covariance stage of a free network and by feba_dist_init),
                                   1 = force nested dissection.  No counterpart in the reference (dense inverse). */
    double sigma_x;             /* Meas_std                             (main.m:123)     */
    double sigma_y;             /* Meas_std_y, = sigma_x when absent    (main.m:397-402) */
    double threshold;           /* Threshold_Value (main.m:154, used by feba_solve only) */
} feba_settings;

/* The reference's `data` struct (main.m:277-384) as numeric structure-of-arrays. */
typedef struct feba_problem {
    int64_t n_obs;              /* data.n / 2                           (main.m:381)      */
    int32_t n_img;              /* data.numImg                          (main.m:379)      */
    int32_t n_cam;              /* data.numCam                          (main.m:380)      */
    int32_t n_pts;              /* rows of CNT                                           */
    int32_t n_tie;              /* data.numtie = rows of TIE            (main.m:383)      */
    const double *obs_x;        /* [n_obs] data.points(i).x             (main.m:282)      */
    const double *obs_y;        /* [n_obs] data.points(i).y             (main.m:283)      */
    const int32_t *obs_img;     /* [n_obs] ext_index - 1                (main.m:298)      */
    const int32_t *obs_pt;      /* [n_obs] cnt_index - 1                (main.m:358)      */
    const int32_t *img_cam;     /* [n_img] cam_num - 1 of the image     (main.m:322)      */
    const double *eop0;         /* [n_img*6] Xc Yc Zc w p k (rad), row-major (main.m:301-307) */
    const double *iop0;         /* [n_cam*(3+NK+2)] xp yp c k1..kNK p1 p2 (main.m:324-330) */
    const double *cam_box;      /* [n_cam*5] y_dir xmin ymin xmax ymax  (main.m:331-343)  */
    const double *xyz0;         /* [n_pts*3] X Y Z                      (main.m:359-361)  */
    const int32_t *pt_tie;      /* [n_pts] tieIndex - 1, or -1          (main.m:362-375)  */
    feba_settings settings;
} feba_problem;

typedef struct feba_handle feba_handle;

/* Multi-GPU runs: one handle per rank on its own device.  Each rank passes the full
 * image / camera tables and ONLY ITS OWN object points (xyz0, pt_tie, n_tie local) and their
 * observations; partial reduced systems are summed between feba_iterate_assemble() and
 * feba_iterate_solve() through feba_reduced_dev(). */

/* Upload the problem to the current CUDA device, sort/segment observations by point, allocate
 * block storage.  Replaces the per-iteration allocations of BuildAwG.m:29-42. */
int feba_create(const feba_problem *problem, feba_handle **out);
/* One rank of a GROUP of GPUs working on ONE adjustment (one process per GPU, each on its own
 * device).  Every rank passes the SAME, COMPLETE problem and the id of feba_dist_unique_id() (created by one rank
 * and shipped to the others); world must be a power of two.  Collective: wraps ncclCommInitRank.  The image block
 * is cut by nested dissection into `world` subtrees plus shared separators (csrc/feba_order.h); a rank keeps the
 * object points whose images lie in its subtree (all observations of a point stay together, so V_p, W_p and the
 * point's Schur contribution are local, main.m:424-425 being a sum over observations), eliminates its subtree
 * locally, and only the shared top part of the reduced system is summed over NVLink (ncclAllReduce) before the
 * replicated top factorisation.  On such a handle feba_iterate / feba_sync / feba_get_xhat / feba_get_delta /
 * feba_residuals are collective calls (every rank makes them in the same order) and work on the GLOBAL xhat and
 * PHO-row layouts: each rank receives the complete result.  Fails with FEBA_ERR_INVALID when the block cannot be
 * cut into `world` parts; callers then use the replicated form below (feba_create per rank on a point shard +
 * feba_iterate_assemble / feba_reduced_dev / feba_iterate_solve). */
int feba_create_shard(const feba_problem *problem, int32_t rank, int32_t world, const void *id, size_t id_bytes,
                      feba_handle **out);
/* Run this handle's kernels and copies on a caller-owned CUDA stream (a cudaStream_t passed as
 * void*; e.g. the stream the caller's NCCL all-reduce is enqueued on).  Default: a private
 * non-blocking stream. */
int feba_set_stream(feba_handle *h, void *cuda_stream);
void feba_destroy(feba_handle *h);
const char *feba_last_error(const feba_handle *h); /* h may be NULL: last create error */

/* Number of unknowns u, and the split u_c (EOP+IOP part) / 3*n_tie (Buildxhat.m:5-15). */
int feba_num_unknowns(const feba_handle *h, int64_t *u, int64_t *u_c);
/* Number of image observations of the handle (data.n / 2, main.m:381): feba_residuals writes 2*n_obs (v)
 * and 5*n_obs (rsd) doubles -- size the outputs from THIS, not from a caller-side count.  -1: null handle. */
int64_t feba_num_obs(const feba_handle *h);

/* xhat in the layout of Buildxhat.m:22-135 (length u). */
int feba_set_xhat(feba_handle *h, const double *xhat, size_t u);
int feba_get_xhat(feba_handle *h, double *xhat, size_t u);

/* Owned-only transfers for a feba_create_shard handle: the EOP/IOP part (replicated) and the coordinates of the tie
 * points THIS rank owns; the other entries of the caller's vector are neither read nor written.  A distributed
 * caller that keeps xhat split over the ranks between iterations uses these per step and the collective
 * feba_get_xhat once at the end.  Not collective.  On a single-GPU handle they equal feba_set_xhat / feba_get_xhat.
 * feba_num_owned_ties: how many tie points that is. */
int feba_set_xhat_owned(feba_handle *h, const double *xhat, size_t u);
int feba_get_xhat_owned(feba_handle *h, double *xhat, size_t u);
int64_t feba_num_owned_ties(const feba_handle *h);

/* One Gauss-Newton step = body of the while loop main.m:412-494:
 * BuildAwG + normal equations + (bordered) solve + un-scaling + xhat += delta.
 * deltasum_out = sumabs(delta) (main.m:487).  */
int feba_iterate(feba_handle *h, double *deltasum_out);

/* The same step in two halves, for multi-GPU runs: _assemble forms this rank's
 * partial reduced camera system (point blocks eliminated) on the device; the caller sums
 * feba_reduced_dev() across ranks (ncclAllReduce / torch.distributed.all_reduce, sum, f64);
 * _solve adds the inner-constraint border, factorises, updates the EOP/IOP part (identical on
 * every rank), back-substitutes THIS rank's points.  deltasum_cam = sum|delta| over the EOP/IOP
 * part, deltasum_pts = over this rank's tie points: sumabs(delta) of main.m:487 is
 * deltasum_cam + sum over ranks of deltasum_pts. */
int feba_iterate_assemble(feba_handle *h);
int feba_reduced_dev(feba_handle *h, double **dev_ptr, size_t *count);
int feba_iterate_solve(feba_handle *h, double *deltasum_cam, double *deltasum_pts);
/* Optional, smaller exchange: only the lower triangle and the augmented block row of the reduced
 * system are read by the solve half.  _pack copies those (per group of 512 columns: the rows from the
 * group's first row down) into one contiguous device buffer of about half the size and returns it;
 * the caller sums THAT buffer across ranks instead of feba_reduced_dev() and calls _unpack before
 * feba_iterate_solve.  Device copies on the handle's stream, asynchronous like the caller's collective. */
int feba_reduced_pack(feba_handle *h, double **dev_ptr, size_t *count);
int feba_reduced_unpack(feba_handle *h);

/* Optional: let the ranks factorise the summed reduced system TOGETHER instead of each rank
 * repeating it (the factorisation is ~70 % of an iteration at u_c = 12,010 and does not shrink with
 * the point shards).  Rank 0 obtains an id, the caller ships the FEBA_DIST_ID_BYTES bytes to every
 * rank (MPI_Bcast, torch.distributed.broadcast, a file ...), then every rank calls feba_dist_init on
 * its handle -- collectively, like ncclCommInitRank, which it wraps.  After that feba_iterate_solve
 * is collective too: supertile columns of the reduced matrix are dealt out cyclically, each owner
 * factorises its panel and broadcasts it (NCCL, loaded at run time from libnccl.so.2 or
 * $FEBA_NCCL_LIB), all ranks end with the same factor and the same EOP/IOP update.  Results are the
 * ones of the replicated solve; small reduced systems (< 96 blocks of 64) keep the replicated solve. */
#define FEBA_DIST_ID_BYTES 128
int feba_dist_unique_id(void *id, size_t bytes);
int feba_dist_init(feba_handle *h, int32_t rank, int32_t world, const void *id, size_t bytes);

/* Last increment delta (un-scaled, main.m:458-482), length u. */
int feba_get_delta(feba_handle *h, double *delta, size_t u);

/* After the loop: v = A*delta + w of the LAST iteration (main.m:569), BuildRSD columns
 * r vx vy vr vt (BuildRSD.m:29-40) in PHO row order, and stats = {RMSx, RMSy, RMS, sigma02,
 * sum vx^2, sum vy^2} (main.m:594-601).  v: [2*n_obs] interleaved x,y; rsd: [n_obs*5] row-major.
 * Either output pointer may be NULL.  sigma02 uses this handle's own n and u (main.m:601, n - u);
 * multi-GPU callers recombine it from stats[4], stats[5] of every rank.  */
int feba_residuals(feba_handle *h, double *v, double *rsd, double stats[6]);

/* Whole loop main.m:412-494 on the device: iterate until deltasum <= threshold or the cap.
 * iterations_out, trace_out[<=cap] (deltasum per iteration, may be NULL). */
int feba_solve(feba_handle *h, int32_t *iterations_out, double *trace_out, size_t trace_cap);

/* BatchRun sweep (BatchRun.m:57-65: main() over many data folders): a batch of independent handles on ONE device
 * advances by one Gauss-Newton step per call.  From the second step on the step of ALL handles is ONE captured CUDA
 * graph (one launch on the host; the blocks run side by side on the device).  Follow with feba_sync() per handle
 * for its deltasum.  Handles that have converged are dropped by creating a new batch of the remaining ones.
 * Results are bit-identical to feba_iterate() per handle. */
typedef struct feba_batch feba_batch;
int feba_batch_create(feba_handle *const *handles, int32_t n, feba_batch **out);
int feba_batch_iterate_async(feba_batch *b);
void feba_batch_destroy(feba_batch *b);

/* Covariance outputs from the normal matrix of the LAST iteration, as the reference
 * keeps Cx = NG^-1(1:u,1:u) of its last loop pass (main.m:432-444).  Values are COFACTORS: multiply
 * by sigma02 for Cx (main.m:602).
 *   feba_cov_diag : diag(Cx)/sigma02 for all u unknowns (tie points: V^-1 + V^-1 W' Qcc W V^-1 per point),
 *                   distortion entries un-scaled as main.m:468-480.
 *   feba_cov_block: k x k block (row-major) for EOP/IOP unknown indices idx[] (0-based xhat positions),
 *                   BEFORE un-scaling -- the values main.m:446-456 normalises into Correlation.
 * feba_cov_prepare builds the inverse of the reduced system once (~u_c^3 flop); the other two call it. */
int feba_cov_prepare(feba_handle *h);
int feba_cov_diag(feba_handle *h, double *qdiag, size_t u);
int feba_cov_block(feba_handle *h, const int64_t *idx, int32_t k, double *out);

/* Timing of the last completed iteration in milliseconds (CUDA events on the handle's stream):
 * ms[0] parameter tables + clearing S, ms[1] fused BuildAwG/normal-blocks/Schur kernel,
 * ms[2] border + Cholesky (includes the caller's all-reduce in a multi-GPU run), ms[3] border
 * solve + backward substitution, ms[4] update + point back-substitution, ms[5] total. */
int feba_last_timing(const feba_handle *h, double ms[6]);
/* The same six entries, then ms[6] = exchange of the shared top part inside ms[2] on a feba_create_shard handle
 * (pack + ncclAllReduce + unpack; 0 otherwise), ms[7] = device time of the last feba_residuals kernel stage. */
int feba_last_timing_ex(const feba_handle *h, double ms[8]);

/* Kernel launches issued by this handle since creation (for bench.py's gpu_launches). */
int64_t feba_launch_count(const feba_handle *h);

/* Block-sparse form of the reduced system (automatic, see feba_settings.plan; csrc/feba_order.h, feba_sparse.h;
 * environment FEBA_SPARSE=0 or FEBA_PLAN=-1 keep the dense form).  info[0] = 1 when active, info[1] /
 * info[2] = structurally non-zero / all lower supertiles of the factorised part (symbolic fill included),
 * info[3] = number of datum images of the sparse-datum form (0: no inner constraints).  The reference has no
 * counterpart: main.m:432,442 invert the dense bordered matrix. */
int feba_sparse_info(const feba_handle *h, int32_t info[4]);
/* Plan of the reduced system in use: info = {nested dissection (1) or identity order (0), rows incl. padding,
 * supertiles, tree nodes, 64-blocks on the longest chain of dependent diagonal factorisations, world, first row
 * of the shared top part (= rows when single GPU), observations held by this handle}; flop = {factorisation flop
 * of the plan's task graph, dense (64 blocks)^3 / 3}.  flop may be NULL. */
int feba_plan_info(const feba_handle *h, int32_t info[8], double flop[2]);

/* Diagnostic for parity tests: the point-eliminated camera system left by a pending
 * feba_iterate_assemble() (S = N_cc - W V^-1 W', g = u_c - W V^-1 u_p; main.m:424-425 reduced),
 * S_out [u_c*u_c] column-major full symmetric, g_out [u_c]; either may be NULL. */
int feba_debug_reduced(feba_handle *h, double *S_out, double *g_out);

/* Device-resident variant for benchmarking with inputs already in HBM: identical to
 * feba_iterate but never copies deltasum back unless asked (deltasum_out may be NULL). */
int feba_iterate_async(feba_handle *h);
int feba_iterate_solve_async(feba_handle *h); /* second half only (after the caller's all-reduce) */
int feba_sync(feba_handle *h, double *deltasum_out);

#ifdef __cplusplus
}
#endif
#endif /* FEBA_H */
