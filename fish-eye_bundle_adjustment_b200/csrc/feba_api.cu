// C ABI of the Gauss-Newton hot path (include/feba.h): handle, uploads, iteration driver.
//
// Replaces the body of the reference's while loop (main.m:412-494) and the residual stage
// (main.m:569-602).  Host work here is limited to argument checks, the one-off ordering of the
// observations by object point (feba_create) and launching kernels; there is no CPU compute path:
// every call fails with FEBA_ERR_CUDA when no CUDA device is usable.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>
#include <algorithm>

#include "../../include/feba.h"
#include "feba_dev.h"
#include "feba_kernels.h"
#include "feba_model.cuh"
#include "feba_chunks.h"
#include "feba_order.h"
#include "feba_sparse.h"

using namespace feba;

static thread_local std::string g_create_error;

struct feba_handle {
    DevProblem P{};
    feba_settings cfg{};
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::vector<void*> allocs;
    // device state
    double *eop = nullptr, *iop = nullptr, *cam_box = nullptr, *img_tab = nullptr, *cam_tab = nullptr;
    double *xhat = nullptr, *sol = nullptr, *dcam = nullptr, *dcam_unscaled = nullptr, *work = nullptr;
    double *ywork = nullptr, *Linv = nullptr, *dvec = nullptr, *dg = nullptr, *gwork = nullptr;
    double *covU = nullptr, *covQ = nullptr, *covY = nullptr, *covT = nullptr;   // covariance stage (lazy)
    bool cov_ready = false;
    // plan of the reduced system (feba_order.h): row order, supertiles, pattern, owners
    ReducedPlan plan;
    std::vector<int> tile_chain, tile_owner, block_owner, row_first;
    int2* blk_list = nullptr;     // non-zero 64x64 blocks of the lower triangle, then the augmented block row
    int n_blk_list = 0;           // all of them (cleared per iteration)
    int n_blk_scale = 0;          // without the last entry, the augmented diagonal block (border / scaling pass)
    DagStreams dag;               // task-graph factorisation (large reduced systems)
    std::vector<cudaEvent_t> dag_events;
    bool use_dag = false;
    DistCtx dist;                 // group of GPUs (feba_create_shard, or the legacy feba_dist_init)
    GreenPair green;              // SM partitions of the task graph (chain | bulk), optional
    bool dist_active = false;     // legacy: column-cyclic shared factorisation of a dense system
    // group run on a nested-dissection plan (feba_create_shard): this rank holds the points of its own subtrees
    bool shard = false;
    int rank = 0, world = 1;
    int64_t n_obs_global = 0;
    unsigned char* tie_mine = nullptr;    // device, n_tie: 1 = tie point owned by this rank
    std::vector<int> own_ties;            // tie indices owned by this rank (ascending); device copy + packed staging
    int* own_ties_dev = nullptr;
    double* own_packed = nullptr;         // device, 3 per owned tie
    double* own_pin = nullptr;            // pinned host staging: u_c + 3 per owned tie
    int top_row0 = 0;                     // first row of the shared top part (summed over the ranks)
    double* xchg = nullptr;               // packed lower trapezoid of the shared top part
    size_t xchg_count = 0;
    unsigned char* datum_dev = nullptr;
    bool dag_cols = false;        // column form of the task graph (chol_cols), issued eagerly unless solve_graph
    bool solve_graph = true;      // capture the solve half into a CUDA graph
    double *scal = nullptr;       // [0] sumabs camera part, [1] sumabs points, [2] sum vx^2, [3] sum vy^2
    double *v_out = nullptr, *rsd_out = nullptr, *delta_out = nullptr;
    double *v_pin = nullptr, *rsd_pin = nullptr;   // pinned staging of the residual stage (chunked D2H)
    size_t pin_chunk = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_ev[2] = {nullptr, nullptr};
    int *opt = nullptr, *tie_pt = nullptr, *info = nullptr;
    double* scal_host = nullptr;  // pinned mirror of scal
    int* info_host = nullptr;
    size_t S_count = 0;
    double* packed = nullptr;     // lower trapezoids of S, contiguous (feba_reduced_pack)
    size_t packed_count = 0;
    long long n_pairs = 0;
    ChunkDev chunks{};            // chunk form of the assembly (feba_chunks.h); n_chunks == 0: image-major form
    bool use_chunks = false;
    int64_t u = 0;
    int n_partial = 0;
    int64_t launches = 0;
    int iterations = 0;           // completed iterations since the last set_xhat
    int phase = 0;                // 0 idle, 1 assembled (waiting for solve)
    // CUDA graphs of the two halves of an iteration (launch-bound inner loops: ~1,900 kernels per
    // iteration at u_c = 12,010).  Call 0 of a phase runs eagerly, call 1 is captured, later calls replay.
    struct GraphSlot {
        cudaGraphExec_t exec = nullptr;
        int64_t launches = 0;
        int calls = 0;
    } g_assemble, g_solve;
    bool use_graph = true;
    bool capturing = false;
    double timing[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool timing_valid = false;
    std::string err;
};

namespace {

int fail(feba_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else g_create_error = buf;
    return code;
}

#define CU(h, call)                                                                                  \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(h, FEBA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                                   \
    } while (0)

#define NC(h, call)                                                                  \
    do {                                                                             \
        if ((call) != 0) return fail(h, FEBA_ERR_CUDA, "%s (%s:%d)", (h)->dist.err, __FILE__, __LINE__); \
    } while (0)

// Event record that also works inside stream capture (external-record node keeps the timestamp).
cudaError_t record(feba_handle* h, int i) {
    return h->capturing ? cudaEventRecordWithFlags(h->ev[i], h->stream, cudaEventRecordExternal)
                        : cudaEventRecord(h->ev[i], h->stream);
}

void drop_graphs(feba_handle* h) {
    for (feba_handle::GraphSlot* g : {&h->g_assemble, &h->g_solve}) {
        if (g->exec) cudaGraphExecDestroy(g->exec);
        g->exec = nullptr;
        if (g->calls > 1) g->calls = 1;
    }
}

template <typename T>
cudaError_t dev_alloc(feba_handle* h, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, (count ? count : 1) * sizeof(T));
    if (e == cudaSuccess) {
        h->allocs.push_back(q);
        *p = static_cast<T*>(q);
    }
    return e;
}

template <typename T>
cudaError_t upload(feba_handle* h, T** p, const T* src, size_t count) {
    cudaError_t e = dev_alloc(h, p, count);
    if (e != cudaSuccess) return e;
    if (count) e = cudaMemcpyAsync(*p, src, count * sizeof(T), cudaMemcpyHostToDevice, h->stream);
    return e;
}

// Stream of the task-graph pool: slot 0 (panel chain) lives in the chain partition when the device
// is split, all others in the bulk partition.
cudaError_t pool_stream(feba_handle* h, int slot, int priority, cudaStream_t* out) {
    if (h->green.chain) {
        if (green_stream(slot == 0 ? h->green.chain : h->green.bulk, priority, out) == 0) return cudaSuccess;
        return cudaErrorNotSupported;
    }
    return cudaStreamCreateWithPriority(out, cudaStreamNonBlocking, priority);
}

void teardown_pool(feba_handle* h) {
    for (int s2 = 0; s2 < h->dag.n_streams; ++s2) {
        cudaStreamDestroy(h->dag.streams[s2]);
        if (h->dag.join[s2]) cudaEventDestroy(h->dag.join[s2]);
        h->dag.streams[s2] = nullptr;
        h->dag.join[s2] = nullptr;
    }
    if (h->dag.join[h->dag.n_streams]) {
        cudaEventDestroy(h->dag.join[h->dag.n_streams]);
        h->dag.join[h->dag.n_streams] = nullptr;
    }
    h->dag.n_streams = 0;
    if (h->dag.fork) cudaEventDestroy(h->dag.fork);
    h->dag.fork = nullptr;
    for (auto& ev : h->dag_events)
        if (ev) cudaEventDestroy(ev);
    h->dag_events.clear();
    green_destroy(&h->green);
    h->use_dag = false;
}

// Stream pool and tile events of the task-graph factorisation.  n_hi streams get the high priority (slot 0
// carries the critical path); n_events = (tiles + 1)^2.
int make_pool(feba_handle* h, int NS, int n_hi, int n_events, int reserve_sms) {
    teardown_pool(h);
    if (NS < 2) NS = 2;
    if (NS > 32) NS = 32;
    int lo = 0, hi = 0;
    CU(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    if (reserve_sms > 0) {
        char why[128];
        const int rc = green_create(h->device, reserve_sms, &h->green, why, sizeof(why));
        if (std::getenv("FEBA_VERBOSE")) {
            if (rc) fprintf(stderr, "[feba] %s; the task graph shares the whole GPU\n", why);
            else fprintf(stderr, "[feba] SM partitions: chain %d, bulk %d\n", h->green.chain_sms, h->green.bulk_sms);
        }
    }
    for (int s2 = 0; s2 < NS; ++s2) {
        CU(h, pool_stream(h, s2, s2 < n_hi ? hi : lo, &h->dag.streams[s2]));
        ++h->dag.n_streams;
        CU(h, cudaEventCreateWithFlags(&h->dag.join[s2], cudaEventDisableTiming));
    }
    CU(h, cudaEventCreateWithFlags(&h->dag.fork, cudaEventDisableTiming));
    h->dag_events.assign((size_t)n_events, nullptr);
    for (auto& ev : h->dag_events) CU(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    h->dag.events = h->dag_events.data();
    h->dag.n_events = (int)h->dag_events.size();
    h->use_dag = true;
    return FEBA_OK;
}

// Pool of the plan's task graph (chol_tiles).  Dense plan, one GPU, measured on u_c = 12,010: supertiles of
// 6/8/12/14/16/20/24 blocks -> 48.7/32.6/26.3/25.3/25.6/27.7/28.1 ms, recursive form 30.2 ms; 4/8/12/16 streams ->
// 28.3/25.6/27.6/27.2 ms.  Nested-dissection plans have many independent chains: 16 streams.
// FEBA_DAG_STREAMS overrides.
int setup_plan_pool(feba_handle* h) {
    const ReducedPlan& pl = h->plan;
    if (pl.NT < 2) {
        teardown_pool(h);
        return FEBA_OK;
    }
    const char* e_s = std::getenv("FEBA_DAG_STREAMS");
    const int NS = e_s ? std::atoi(e_s) : (pl.masked ? 16 : 8);
    // FEBA_DAG_FORM=cols: column form of the dense task graph (chol_cols, uniform supertiles), eager unless
    // FEBA_SOLVE_GRAPH=1; the default is the tile form captured into the iteration's CUDA graph
    const char* e_f = std::getenv("FEBA_DAG_FORM");
    h->dag_cols = !pl.masked && e_f && std::strcmp(e_f, "cols") == 0 && NS >= 5;
    h->solve_graph = !h->dag_cols;
    if (const char* e_sg = std::getenv("FEBA_SOLVE_GRAPH")) h->solve_graph = std::atoi(e_sg) != 0;
    return make_pool(h, NS, 1, (pl.NT + 1) * (pl.NT + 1), 0);
}

// Legacy group form of a DENSE system (feba_dist_init): about 19 uniform supertiles, column form issued eagerly,
// 32 SMs for the panel chain (2 GPUs, whole iteration: tile graph 32.2 ms; column form T=14 33.6, +32 SMs 31.2,
// T=10 30.6 ms).  FEBA_DAG_TILE=t, FEBA_DAG_STREAMS, FEBA_GREEN_SMS=r (multiple of 8, 0 = shared),
// FEBA_DAG_FORM=cols|tiles, FEBA_SOLVE_GRAPH=0|1 override.
int setup_legacy_group_pool(feba_handle* h) {
    const int nb = h->P.n_pad / kBlk;
    const char* e_t = std::getenv("FEBA_DAG_TILE");
    const char* e_s = std::getenv("FEBA_DAG_STREAMS");
    int T = 0;
    if (e_t) {
        T = std::atoi(e_t);
        if (T > 0 && nb < 2 * T) T = 0;
    } else if (nb >= 96) {
        T = (nb + 9) / 19;
        if (T < 8) T = 8;
        if (T > 24) T = 24;
    }
    const int NS = e_s ? std::atoi(e_s) : 8;
    if (T <= 0 || NS < 2 || NS > 16) {
        teardown_pool(h);
        return FEBA_OK;
    }
    h->dag.tile_blocks = T;
    const char* e_f = std::getenv("FEBA_DAG_FORM");
    h->dag_cols = e_f ? std::strcmp(e_f, "cols") == 0 : true;
    const char* e_sg = std::getenv("FEBA_SOLVE_GRAPH");
    h->solve_graph = e_sg ? std::atoi(e_sg) != 0 : !h->dag_cols;
    const char* e_g = std::getenv("FEBA_GREEN_SMS");
    const int reserve = e_g ? std::atoi(e_g) : (h->dag_cols ? 32 : 0);
    const int NT = (nb + T - 1) / T;
    return make_pool(h, NS, h->dag_cols ? 3 : 4, (NT + 1) * (NT + 1), reserve);
}

int check_settings(const feba_problem* pr) {
    const feba_settings& s = pr->settings;
    if (s.type < 0 || s.type > 4)
        return fail(nullptr, FEBA_ERR_INVALID, "BuildAwG, invalid type in data.settings.type (typeint %d)", s.type);
    if (s.num_radial < 1 || s.num_radial > FEBA_MAX_NK)
        return fail(nullptr, FEBA_ERR_INVALID, "Num_Radial_Distortions must be 1..%d (got %d)", FEBA_MAX_NK,
                    s.num_radial);
    int ui = 0;
    for (int q = 0; q < 6; ++q) ui += s.estimate_eop[q] ? 1 : 0;
    if (s.inner_constraints && ui != 6)
        return fail(nullptr, FEBA_ERR_INVALID,
                    "Inner_Constraints needs all six EOPs estimated (Gblock is 6 rows, BuildAwG.m:516-525)");
    if (!(s.sigma_x > 0.0) || !(s.sigma_y > 0.0))
        return fail(nullptr, FEBA_ERR_INVALID, "Meas_std / Meas_std_y must be positive");
    if (pr->n_obs < 0 || pr->n_obs > 2000000000LL || pr->n_img < 1 || pr->n_cam < 1 || pr->n_pts < 0 ||
        pr->n_tie < 0)
        return fail(nullptr, FEBA_ERR_INVALID, "bad problem sizes");
    return FEBA_OK;
}

// Row order, supertiles and pattern of the reduced system.  world > 1 (feba_create_shard) needs a nested-
// dissection plan with one subtree per rank; one GPU takes the masked plan when it shortens the chain of
// diagonal factorisations AND the flop count to less than half of the dense form (BASELINE configs[3]: 42 of 188
// blocks, 3.5e10 of 5.8e11 flop; configs[2], a much denser block: no), else the identity plan.
int choose_plan(feba_handle* h, const feba_problem* pr, int n_seg, const std::vector<int>& seg_start,
                const std::vector<int>& simg, const std::vector<int>& seg_pt, int world) {
    const DevProblem& P = h->P;
    const int cam_rows = P.uc * P.n_cam;
    const int nb_id = (P.ui * P.n_img + cam_rows + kBlk - 1) / kBlk;
    PlanOptions opt = plan_options_from_env();
    if (pr->settings.plan < 0 && world == 1) opt.mode = -1;
    if (pr->settings.plan > 0) opt.mode = 1;
    auto identity = [&]() {
        int T = 0;
        if (const char* e_t = std::getenv("FEBA_DAG_TILE")) {
            T = std::atoi(e_t);
            if (T > 0 && nb_id < 2 * T) T = 0;
        } else if (nb_id >= 96) {
            T = (nb_id + 6) / 13;
            if (T < 8) T = 8;
            if (T > 24) T = 24;
        }
        h->plan = identity_plan(P.n_img, P.ui, cam_rows, T);
        h->dag.tile_blocks = T > 0 ? T : 8;
    };
    const bool can_mask = P.ui > 0 && (!P.inner || P.ui == 6);
    if (world == 1 && (opt.mode < 0 || !can_mask || (opt.mode == 0 && nb_id < opt.min_blocks))) {
        identity();
        return FEBA_OK;
    }
    if (world > 1 && !can_mask)
        return fail(h, FEBA_ERR_INVALID, "feba_create_shard needs EOP unknowns (nothing to dissect)");
    std::vector<unsigned char> seg_tie((size_t)n_seg);
    for (int s2 = 0; s2 < n_seg; ++s2) seg_tie[(size_t)s2] = pr->pt_tie[seg_pt[(size_t)s2]] >= 0;
    std::vector<int> ptr, idx;
    image_adjacency(P.n_img, n_seg, seg_start.data(), simg.data(), seg_tie.data(), world > 1, ptr, idx);
    std::vector<double> pos((size_t)P.n_img * 3);
    for (int i = 0; i < P.n_img; ++i)
        for (int k = 0; k < 3; ++k) pos[3 * (size_t)i + k] = pr->eop0[6 * (size_t)i + k];
    ReducedPlan mp = masked_plan(P.n_img, P.ui, cam_rows, ptr.data(), idx.data(), pos.data(), P.inner != 0, world, opt);
    if (world > 1) {
        if (mp.world != world)
            return fail(h, FEBA_ERR_INVALID, "the image block cannot be cut into %d subtrees: use the replicated form "
                                             "(feba_create + feba_reduced_dev)", world);
        h->plan = std::move(mp);
        return FEBA_OK;
    }
    const double dense_flop = std::pow((double)kBlk * nb_id, 3) / 3;
    const bool pays = mp.flop <= 0.5 * dense_flop && 2 * mp.chain_blocks <= nb_id;
    if (opt.mode > 0 || pays) h->plan = std::move(mp);
    else identity();
    return FEBA_OK;
}

// Everything the solve half needs from the plan: tile arrays for chol_tiles, owners per 64-block, the envelope
// for the backward substitution, the block list on the device, row maps on the device.
int install_plan(feba_handle* h) {
    DevProblem& P = h->P;
    const ReducedPlan& pl = h->plan;
    P.n_pad = pl.n_pad;
    P.ld = P.n_pad + kBlk;
    P.off_cam = pl.off_cam;
    P.ext_off_cam = P.ui * P.n_img;
    const int nb = P.n_pad / kBlk;
    int* d_img_row = nullptr;
    int* d_row_ext = nullptr;
    CU(h, upload(h, &d_img_row, pl.img_row.data(), pl.img_row.size()));
    CU(h, upload(h, &d_row_ext, pl.row_ext.data(), pl.row_ext.size()));
    P.img_row = d_img_row;
    P.row_ext = d_row_ext;
    P.row_owner = nullptr;
    P.rank = h->rank;
    h->tile_chain.assign((size_t)pl.NT, 0);
    h->tile_owner.assign((size_t)pl.NT, -1);
    h->block_owner.assign((size_t)nb, -1);
    if (pl.masked) {
        for (int t = 0; t < pl.NT; ++t) {
            const PlanNode& nd = pl.nodes[(size_t)pl.tile_node[(size_t)t]];
            // slot 0 = critical path (root node); other nodes by depth level and position inside the level
            h->tile_chain[(size_t)t] = pl.tile_node[(size_t)t] == 0 ? 0 : 1 + nd.lane + 5 * nd.depth;
            h->tile_owner[(size_t)t] = h->shard ? nd.owner : -1;
            for (int b = pl.tile_b0[(size_t)t]; b < pl.tile_b0[(size_t)t + 1]; ++b) h->block_owner[(size_t)b] = h->tile_owner[(size_t)t];
        }
        h->row_first = pl.row_first_block();
    } else {
        h->row_first.clear();
    }
    if (h->shard) {
        std::vector<int> row_owner((size_t)P.n_pad);
        for (int r = 0; r < P.n_pad; ++r) row_owner[(size_t)r] = h->block_owner[(size_t)(r / kBlk)];
        int* d_ro = nullptr;
        CU(h, upload(h, &d_ro, row_owner.data(), row_owner.size()));
        P.row_owner = d_ro;
        h->top_row0 = pl.top_tile0 < pl.NT ? pl.tile_b0[(size_t)pl.top_tile0] * kBlk : P.n_pad;
    }
    // block list: lower blocks of the non-zero tiles (every lower block for a dense plan), then the augmented row
    std::vector<int2> list;
    for (int ti = 0; ti < pl.NT; ++ti)
        for (int tj = 0; tj <= ti; ++tj) {
            if (!pl.at(ti, tj)) continue;
            if (h->shard) {           // tiles of other ranks' subtrees are never touched here
                const int oi = h->tile_owner[(size_t)ti], oj = h->tile_owner[(size_t)tj];
                if ((oi >= 0 && oi != h->rank) || (oj >= 0 && oj != h->rank)) continue;
            }
            for (int bi = pl.tile_b0[(size_t)ti]; bi < pl.tile_b0[(size_t)ti + 1]; ++bi)
                for (int bj = pl.tile_b0[(size_t)tj]; bj < pl.tile_b0[(size_t)tj + 1] && bj <= bi; ++bj)
                    list.push_back(make_int2(bi, bj));
        }
    for (int bj = 0; bj < nb; ++bj) list.push_back(make_int2(nb, bj));
    h->n_blk_scale = (int)list.size();
    list.push_back(make_int2(nb, nb));      // augmented diagonal block T = -B' M^-1 B: cleared, never scaled
    h->n_blk_list = (int)list.size();
    CU(h, upload(h, &h->blk_list, list.data(), list.size()));
    CU(h, cudaStreamSynchronize(h->stream));
    if (std::getenv("FEBA_VERBOSE")) {
        int nzt = 0, allt = 0;
        for (int i = 0; i < pl.NT; ++i)
            for (int j = 0; j <= i; ++j) {
                ++allt;
                nzt += pl.at(i, j) ? 1 : 0;
            }
        fprintf(stderr, "[feba] plan: %s, %d rows (u_c %d), %d supertiles, %d of %d lower supertiles non-zero, chain %d of %d "
                        "blocks, %.2e flop (dense %.2e), %zu nodes, %zu datum images, world %d\n",
                pl.masked ? "nested dissection" : "identity", pl.n_pad, P.n_red, pl.NT, nzt, allt, pl.chain_blocks, nb,
                pl.flop, pl.flop_dense, pl.nodes.size(), pl.datum.size(), h->world);
    }
    return FEBA_OK;
}

}  // namespace

extern "C" {

const char* feba_last_error(const feba_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void feba_destroy(feba_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void* p : h->allocs) cudaFree(p);
    if (h->scal_host) cudaFreeHost(h->scal_host);
    if (h->info_host) cudaFreeHost(h->info_host);
    if (h->own_pin) cudaFreeHost(h->own_pin);
    if (h->v_pin) cudaFreeHost(h->v_pin);
    if (h->rsd_pin) cudaFreeHost(h->rsd_pin);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (auto& e : h->copy_ev)
        if (e) cudaEventDestroy(e);
    drop_graphs(h);
    teardown_pool(h);
    if (h->dist.comm) dist_comm_destroy(&h->dist);
    if (h->dist.stream) cudaStreamDestroy(h->dist.stream);
    for (auto& e : h->ev)
        if (e) cudaEventDestroy(e);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

static int create_impl(const feba_problem* pr, int rank, int world, const void* id, feba_handle** out) {
    if (!pr || !out) return fail(nullptr, FEBA_ERR_INVALID, "null argument");
    *out = nullptr;
    int rc = check_settings(pr);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, FEBA_ERR_CUDA, "no CUDA device: the hot path has no CPU fallback");
    feba_handle* h = new (std::nothrow) feba_handle();
    if (!h) return fail(nullptr, FEBA_ERR_INVALID, "out of host memory");
    struct Guard {
        feba_handle* h;
        bool keep = false;
        ~Guard() {
            if (!keep) {
                g_create_error = h->err;
                feba_destroy(h);
            }
        }
    } guard{h};
    h->rank = rank;
    h->world = world;
    h->shard = world > 1;
    CU(h, cudaGetDevice(&h->device));
    CU(h, cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device));
    CU(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
    for (auto& e : h->ev) CU(h, cudaEventCreate(&e));
    h->cfg = pr->settings;
    h->use_graph = std::getenv("FEBA_NO_GRAPH") == nullptr;
    const feba_settings& s = pr->settings;

    DevProblem& P = h->P;
    P.n_img = pr->n_img;
    P.n_cam = pr->n_cam;
    P.n_pts = pr->n_pts;
    P.n_tie = pr->n_tie;
    P.type = s.type;
    P.NK = s.num_radial;
    P.NC = P.NK + 5;
    P.inner = s.inner_constraints ? 1 : 0;
    P.aug_rows = kAugRows;
    P.datum = nullptr;
    P.px = 1.0 / (s.sigma_x * s.sigma_x);      // main.m:396-405
    P.py = 1.0 / (s.sigma_y * s.sigma_y);
    // slots of the estimated parameters (BuildAwG.m:24-25, :52-93, :110-155)
    int k = 0;
    for (int q = 0; q < 6; ++q) P.ecol[q] = s.estimate_eop[q] ? k++ : -1;
    P.ui = k;
    k = 0;
    for (int q = 0; q < 16; ++q) P.ccol[q] = -1;
    P.ccol[0] = s.estimate_xp ? k++ : -1;
    P.ccol[1] = s.estimate_yp ? k++ : -1;
    P.ccol[2] = s.estimate_c ? k++ : -1;
    for (int j = 0; j < P.NK; ++j) P.ccol[3 + j] = s.estimate_radial ? k++ : -1;
    for (int j = 0; j < 2; ++j) P.ccol[3 + P.NK + j] = s.estimate_decent ? k++ : -1;
    P.uc = k;
    P.n_red = P.ui * P.n_img + P.uc * P.n_cam;
    if (P.n_red < 1) return fail(h, FEBA_ERR_INVALID, "no EOP/IOP unknowns: nothing to adjust");
    h->u = (int64_t)P.n_red + 3 * (int64_t)pr->n_tie;
    h->n_obs_global = pr->n_obs;

    // ---- order the observations by object point (stable counting sort, PHO order inside a point)
    const int64_t n = pr->n_obs;
    std::vector<int> start((size_t)pr->n_pts + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
        const int p = pr->obs_pt[i], im = pr->obs_img[i];
        if (p < 0 || p >= pr->n_pts) return fail(h, FEBA_ERR_INVALID, "obs_pt[%lld] out of range", (long long)i);
        if (im < 0 || im >= pr->n_img) return fail(h, FEBA_ERR_INVALID, "obs_img[%lld] out of range", (long long)i);
        ++start[(size_t)p + 1];
    }
    std::vector<int> seg_start, seg_pt;
    seg_start.reserve((size_t)pr->n_pts + 1);
    seg_pt.reserve((size_t)pr->n_pts);
    for (int p = 0; p < pr->n_pts; ++p) {
        if (start[(size_t)p + 1] > 0) {
            seg_start.push_back(start[p]);
            seg_pt.push_back(p);
        }
        start[(size_t)p + 1] += start[p];
    }
    seg_start.push_back((int)n);
    int n_seg = (int)seg_pt.size();
    std::vector<int> perm((size_t)n), simg((size_t)n), spt((size_t)n);
    std::vector<double> sx((size_t)n), sy((size_t)n);
    {
        std::vector<int> cur(start.begin(), start.end() - 1);
        for (int64_t i = 0; i < n; ++i) {
            const int p = pr->obs_pt[i];
            const int d = cur[p]++;
            perm[d] = (int)i;
            simg[d] = pr->obs_img[i];
            spt[d] = p;
            sx[d] = pr->obs_x[i];
            sy[d] = pr->obs_y[i];
        }
    }
    // tie index <-> CNT row (main.m:362-375, Buildxhat.m:108-135)
    std::vector<int> tie_pt((size_t)pr->n_tie, -1);
    for (int p = 0; p < pr->n_pts; ++p) {
        const int t = pr->pt_tie[p];
        if (t < -1 || t >= pr->n_tie) return fail(h, FEBA_ERR_INVALID, "pt_tie[%d] out of range", p);
        if (t >= 0) {
            if (tie_pt[t] >= 0) return fail(h, FEBA_ERR_INVALID, "tie index %d used by two points", t);
            tie_pt[t] = p;
        }
    }
    for (int j = 0; j < pr->n_img; ++j)
        if (pr->img_cam[j] < 0 || pr->img_cam[j] >= pr->n_cam)
            return fail(h, FEBA_ERR_INVALID, "img_cam[%d] out of range", j);
    for (int c = 0; c < pr->n_cam; ++c)
        if (std::fabs(pr->cam_box[5 * c]) != 1.0)
            return fail(h, FEBA_ERR_INVALID, "y_dir should be +-1 only (main.m:334-337)");
    // the multi-camera point pass keeps the per-camera blocks of a point in shared memory for at most
    // kMaxCamPt (4) different cameras: refuse the problem HERE, before any state exists (the reference has no
    // such limit, BuildAwG.m:110-155; documented in feba.h)
    if (P.uc > 0 && pr->n_cam > FEBA_MAX_CAMS_PER_POINT) {
        for (int sg = 0; sg < n_seg; ++sg) {
            int cams[FEBA_MAX_CAMS_PER_POINT], nc = 0;
            for (int o = seg_start[sg]; o < seg_start[sg + 1]; ++o) {
                const int c = pr->img_cam[simg[o]];
                bool found = false;
                for (int kk = 0; kk < nc; ++kk) found = found || cams[kk] == c;
                if (found) continue;
                if (nc == FEBA_MAX_CAMS_PER_POINT)
                    return fail(h, FEBA_ERR_INVALID,
                                "object point %d (CNT row, 0-based) is observed by images of more than %d different "
                                "cameras while camera parameters are estimated: unsupported (feba.h)", seg_pt[sg],
                                FEBA_MAX_CAMS_PER_POINT);
                cams[nc++] = c;
            }
        }
    }

    // ---- plan of the reduced system; a group keeps only the points of this rank's subtrees
    rc = choose_plan(h, pr, n_seg, seg_start, simg, seg_pt, world);
    if (rc) return rc;
    std::vector<unsigned char> tie_mine;
    if (h->shard) {
        std::vector<int> owner((size_t)n_seg);
        if (!plan_point_owner(h->plan, n_seg, seg_start.data(), simg.data(), owner.data()))
            return fail(h, FEBA_ERR_INVALID, "internal: a point's images do not lie on one path of the dissection tree");
        std::vector<int> l_start, l_pt, l_perm, l_simg, l_spt;
        std::vector<double> l_sx, l_sy;
        tie_mine.assign((size_t)pr->n_tie, 0);
        for (int sg = 0; sg < n_seg; ++sg) {
            if (owner[(size_t)sg] != rank) continue;
            l_start.push_back((int)l_perm.size());
            l_pt.push_back(seg_pt[(size_t)sg]);
            const int t = pr->pt_tie[seg_pt[(size_t)sg]];
            if (t >= 0) tie_mine[(size_t)t] = 1;
            for (int o = seg_start[(size_t)sg]; o < seg_start[(size_t)sg + 1]; ++o) {
                l_perm.push_back(perm[(size_t)o]);
                l_simg.push_back(simg[(size_t)o]);
                l_spt.push_back(spt[(size_t)o]);
                l_sx.push_back(sx[(size_t)o]);
                l_sy.push_back(sy[(size_t)o]);
            }
        }
        l_start.push_back((int)l_perm.size());
        // tie points without any observation: rank 0 keeps them (their xhat entries only pass through)
        for (int t = 0; t < pr->n_tie; ++t) {
            const int p = tie_pt[(size_t)t];
            if (rank == 0 && p >= 0 && start[(size_t)p + 1] == start[(size_t)p]) tie_mine[(size_t)t] = 1;
        }
        seg_start.swap(l_start);
        seg_pt.swap(l_pt);
        perm.swap(l_perm);
        simg.swap(l_simg);
        spt.swap(l_spt);
        sx.swap(l_sx);
        sy.swap(l_sy);
        n_seg = (int)seg_pt.size();
    }
    // ---- processing order of the object points: along a space-filling (Morton) curve through the initial
    // coordinates, so that consecutive points are seen by the same images (the per-image tables, the records of an
    // image and the increments of its unknowns are touched together).  Outputs keep the PHO / TIE order.
    if (n_seg > 1 && !(std::getenv("FEBA_POINT_ORDER") && std::atoi(std::getenv("FEBA_POINT_ORDER")) == 0)) {
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int sg = 0; sg < n_seg; ++sg)
            for (int k2 = 0; k2 < 3; ++k2) {
                const double v = pr->xyz0[3 * (size_t)seg_pt[(size_t)sg] + k2];
                if (v < lo[k2]) lo[k2] = v;
                if (v > hi[k2]) hi[k2] = v;
            }
        auto spread = [](unsigned long long v) {          // 21 bits -> every third bit
            v &= 0x1fffffULL;
            v = (v | v << 32) & 0x1f00000000ffffULL;
            v = (v | v << 16) & 0x1f0000ff0000ffULL;
            v = (v | v << 8) & 0x100f00f00f00f00fULL;
            v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
            v = (v | v << 2) & 0x1249249249249249ULL;
            return v;
        };
        std::vector<std::pair<unsigned long long, int>> key((size_t)n_seg);
        // one scale for the three axes (the largest extent): a flat block (aerial: little relief) is ordered by its
        // two long axes -- scaling every axis to its own extent would make the curve 3-D there and a run of 64 points
        // five times wider (measured on BASELINE configs[3]: 31 instead of ~13 images per chunk of the assembly)
        double ext = 0.0;
        for (int k2 = 0; k2 < 3; ++k2) ext = std::max(ext, hi[k2] - lo[k2]);
        for (int sg = 0; sg < n_seg; ++sg) {
            unsigned long long code = 0;
            for (int k2 = 0; k2 < 3; ++k2) {
                const double t = ext > 0 ? (pr->xyz0[3 * (size_t)seg_pt[(size_t)sg] + k2] - lo[k2]) / ext : 0.0;
                unsigned long long q = std::isfinite(t) ? (unsigned long long)(t * 2097151.0) : 0ULL;
                code |= spread(q) << k2;
            }
            key[(size_t)sg] = {code, sg};
        }
        std::sort(key.begin(), key.end());
        std::vector<int> n_start, n_pt, n_perm((size_t)perm.size()), n_simg((size_t)perm.size()), n_spt((size_t)perm.size());
        std::vector<double> n_sx((size_t)perm.size()), n_sy((size_t)perm.size());
        n_start.reserve((size_t)n_seg + 1);
        n_pt.reserve((size_t)n_seg);
        size_t w = 0;
        for (int q = 0; q < n_seg; ++q) {
            const int sg = key[(size_t)q].second;
            n_start.push_back((int)w);
            n_pt.push_back(seg_pt[(size_t)sg]);
            for (int o = seg_start[(size_t)sg]; o < seg_start[(size_t)sg + 1]; ++o, ++w) {
                n_perm[w] = perm[(size_t)o];
                n_simg[w] = simg[(size_t)o];
                n_spt[w] = spt[(size_t)o];
                n_sx[w] = sx[(size_t)o];
                n_sy[w] = sy[(size_t)o];
            }
        }
        n_start.push_back((int)w);
        seg_start.swap(n_start);
        seg_pt.swap(n_pt);
        perm.swap(n_perm);
        simg.swap(n_simg);
        spt.swap(n_spt);
        sx.swap(n_sx);
        sy.swap(n_sy);
    }
    const int64_t nl = (int64_t)perm.size();          // observations held by this handle
    P.n_obs = nl;
    P.n_seg = n_seg;
    std::vector<int> oseg((size_t)nl);
    for (int sg = 0; sg < n_seg; ++sg)
        for (int o = seg_start[sg]; o < seg_start[sg + 1]; ++o) oseg[o] = sg;
    // image-major record positions (stable: point-major order is kept inside an image)
    std::vector<int> img_start((size_t)pr->n_img + 1, 0), ipos((size_t)nl);
    for (int64_t o = 0; o < nl; ++o) ++img_start[(size_t)simg[o] + 1];
    for (int j = 0; j < pr->n_img; ++j) img_start[(size_t)j + 1] += img_start[j];
    {
        std::vector<int> cur(img_start.begin(), img_start.end() - 1);
        for (int64_t o = 0; o < nl; ++o) ipos[o] = cur[simg[o]]++;
    }

    // ---- uploads
    {
        const int rc_plan = install_plan(h);
        if (rc_plan) return rc_plan;
    }
    double *ox, *oy, *xyz, *xyz_prev;
    int *oimg, *operm, *dseg_start, *dseg_pt, *dimg_cam, *dpt_tie, *dimg_start, *diobs, *doseg;
    CU(h, upload(h, &ox, sx.data(), (size_t)nl));
    CU(h, upload(h, &oy, sy.data(), (size_t)nl));
    CU(h, upload(h, &oimg, simg.data(), (size_t)nl));
    CU(h, upload(h, &operm, perm.data(), (size_t)nl));
    CU(h, upload(h, &h->opt, spt.data(), (size_t)nl));
    CU(h, upload(h, &dseg_start, seg_start.data(), seg_start.size()));
    CU(h, upload(h, &dseg_pt, seg_pt.data(), seg_pt.size()));
    CU(h, upload(h, &dimg_cam, pr->img_cam, (size_t)pr->n_img));
    CU(h, upload(h, &dpt_tie, pr->pt_tie, (size_t)pr->n_pts));
    CU(h, upload(h, &h->tie_pt, tie_pt.data(), tie_pt.size()));
    CU(h, upload(h, &dimg_start, img_start.data(), img_start.size()));
    CU(h, upload(h, &diobs, ipos.data(), ipos.size()));
    CU(h, upload(h, &doseg, oseg.data(), oseg.size()));
    if (h->shard) {
        CU(h, upload(h, &h->tie_mine, tie_mine.data(), tie_mine.size()));
        for (int t = 0; t < pr->n_tie; ++t)
            if (tie_mine[(size_t)t]) h->own_ties.push_back(t);
        CU(h, upload(h, &h->own_ties_dev, h->own_ties.data(), h->own_ties.size()));
        CU(h, dev_alloc(h, &h->own_packed, 3 * h->own_ties.size()));
        CU(h, cudaMallocHost((void**)&h->own_pin, ((size_t)P.n_red + 3 * h->own_ties.size() + 1) * sizeof(double)));
    }
    CU(h, dev_alloc(h, &P.rec1, (size_t)nl * kRec1));
    CU(h, dev_alloc(h, &P.rec2, (size_t)nl * (2 + 2 * P.NC)));
    CU(h, upload(h, &h->eop, pr->eop0, (size_t)pr->n_img * 6));
    CU(h, upload(h, &h->iop, pr->iop0, (size_t)pr->n_cam * P.NC));
    CU(h, upload(h, &h->cam_box, pr->cam_box, (size_t)pr->n_cam * 5));
    CU(h, upload(h, &xyz, pr->xyz0, (size_t)pr->n_pts * 3));
    CU(h, upload(h, &xyz_prev, pr->xyz0, (size_t)pr->n_pts * 3));
    CU(h, dev_alloc(h, &h->img_tab, (size_t)pr->n_img * kImgStride));
    CU(h, dev_alloc(h, &h->cam_tab, (size_t)pr->n_cam * kCamStride));
    h->S_count = (size_t)P.ld * (size_t)P.ld;
    CU(h, dev_alloc(h, &P.S, h->S_count));
    CU(h, cudaMemsetAsync(P.S, 0, h->S_count * sizeof(double), h->stream));
    CU(h, dev_alloc(h, &h->xhat, (size_t)h->u));
    CU(h, dev_alloc(h, &h->sol, (size_t)P.n_pad));
    CU(h, dev_alloc(h, &h->dcam, (size_t)P.n_pad));
    CU(h, dev_alloc(h, &h->dcam_unscaled, (size_t)P.n_pad));
    CU(h, dev_alloc(h, &P.dpts, (size_t)pr->n_tie * 3));
    P.pt_rec = nullptr;
    {
        // per-point record of the point pass for the record-based back-substitution (k_backsub_rec); the
        // multi-camera point pass keeps per-camera Fc blocks: there the Jacobians are recomputed (k_backsub)
        const char* e = std::getenv("FEBA_BACKSUB_REC");
        const bool multi_cam = pr->n_cam > 1 && P.uc > 0;
        if (pr->n_tie > 0 && !multi_cam && !(e && e[0] == '0'))
            CU(h, dev_alloc(h, &P.pt_rec, (size_t)pr->n_tie * (size_t)(9 + 3 * (P.NK + 5))));
        // camera block and right-hand side from the records (k_cam_rec) instead of a Jacobian pass (k_cam_direct)
        const char* c = std::getenv("FEBA_CAM_REC");
        P.cam_rec = (P.uc > 0 && !multi_cam && !(c && c[0] == '0')) ? 1 : 0;
    }
    CU(h, dev_alloc(h, &h->work, 64));
    CU(h, dev_alloc(h, &h->ywork, (size_t)P.n_pad));
    CU(h, dev_alloc(h, &h->dvec, (size_t)P.n_pad));
    CU(h, dev_alloc(h, &h->dg, (size_t)P.n_pad));
    CU(h, dev_alloc(h, &h->gwork, (size_t)kGworkDoubles));
    CU(h, dev_alloc(h, &h->Linv, (size_t)(P.n_pad / kBlk) * kBlk * kBlk));
    CU(h, dev_alloc(h, &P.Gt, (size_t)P.n_pad * 8));
    CU(h, cudaMemsetAsync(P.Gt, 0, (size_t)P.n_pad * 8 * sizeof(double), h->stream));
    CU(h, dev_alloc(h, &h->scal, 8));
    CU(h, dev_alloc(h, &h->info, 1));
    CU(h, cudaMallocHost((void**)&h->scal_host, 8 * sizeof(double)));
    CU(h, cudaMallocHost((void**)&h->info_host, sizeof(int)));
    if (pr->n_tie > 0) CU(h, cudaMemsetAsync(P.dpts, 0, (size_t)pr->n_tie * 3 * sizeof(double), h->stream));
    CU(h, cudaMemsetAsync(h->dcam, 0, (size_t)P.n_pad * sizeof(double), h->stream));
    CU(h, cudaMemsetAsync(h->dcam_unscaled, 0, (size_t)P.n_pad * sizeof(double), h->stream));
    CU(h, cudaMemsetAsync(h->Linv, 0, (size_t)(P.n_pad / kBlk) * kBlk * kBlk * sizeof(double), h->stream));
    P.ox = ox;
    P.oy = oy;
    P.oimg = oimg;
    P.operm = operm;
    P.seg_start = dseg_start;
    P.seg_pt = dseg_pt;
    P.img_cam = dimg_cam;
    P.pt_tie = dpt_tie;
    P.img_start = dimg_start;
    P.ipos = diobs;
    P.img_tab = h->img_tab;
    P.cam_tab = h->cam_tab;
    P.xyz = xyz;
    P.xyz_prev = xyz_prev;
    P.dcam = h->dcam;
    P.dcam_unscaled = h->dcam_unscaled;
    {
        const int a = backsub_warps(P, h->sm_count), b = 2 * residual_blocks(P, h->sm_count);
        h->n_partial = a > b ? a : b;
    }
    CU(h, dev_alloc(h, &P.partial, (size_t)h->n_partial));
    {
        const int rc_pool = setup_plan_pool(h);
        if (rc_pool) return rc_pool;
    }
    CU(h, dev_alloc(h, &P.cam_part, (size_t)cam_part_rows(P, h->sm_count) * kCamPart));
    // ---- assembly schedule, static for the life of the handle: image-major form (pair schedule built on the device),
    // or the chunk form (feba_chunks.h) on request
    {
        const bool mc = P.uc > 0 && P.n_cam > 1;
        const char* e_c = std::getenv("FEBA_CHUNKS");
        // opt-in (FEBA_CHUNKS=1): measured on BASELINE configs[3] the chunk form cuts the DRAM traffic of the image /
        // image-pair stage from 20 GB to 4.3 GB but is instruction-bound (2.9e9 warp instructions: index handling and
        // cross-lane sums per 21-pair block) -- 5.8 ms against 4.5 ms for the image-major passes (profiles/r2d_*)
        if (!mc && n_seg > 0 && e_c && std::atoi(e_c) != 0) {
            std::vector<unsigned char> seg_tie((size_t)n_seg);
            for (int sg = 0; sg < n_seg; ++sg) seg_tie[(size_t)sg] = pr->pt_tie[seg_pt[(size_t)sg]] >= 0;
            ChunkSchedule cs = build_chunks(P.n_img, n_seg, seg_start.data(), simg.data(), seg_tie.data(),
                                            h->plan.img_row.data());
            if (cs.ok && cs.n_chunks > 0) {
                ChunkDev& D = h->chunks;
                int *d_obs0, *d_img0, *d_slot_obs0, *d_blk0, *d_pair0, *d_slot_img, *d_tptr, *d_tslots, *d_ba, *d_bb, *d_bptr, *d_bslots;
                unsigned short* d_slot_obs;
                unsigned int* d_pairs;
                CU(h, upload(h, &d_obs0, cs.obs0.data(), cs.obs0.size()));
                CU(h, upload(h, &d_img0, cs.img0.data(), cs.img0.size()));
                CU(h, upload(h, &d_slot_obs0, cs.slot_obs0.data(), cs.slot_obs0.size()));
                CU(h, upload(h, &d_slot_obs, reinterpret_cast<const unsigned short*>(cs.slot_obs.data()), cs.slot_obs.size()));
                CU(h, upload(h, &d_blk0, cs.blk0.data(), cs.blk0.size()));
                CU(h, upload(h, &d_pair0, cs.bslot_pair0.data(), cs.bslot_pair0.size()));
                CU(h, upload(h, &d_pairs, reinterpret_cast<const unsigned int*>(cs.pairs.data()), cs.pairs.size()));
                CU(h, upload(h, &d_slot_img, cs.slot_img.data(), cs.slot_img.size()));
                CU(h, upload(h, &d_tptr, cs.timg_ptr.data(), cs.timg_ptr.size()));
                CU(h, upload(h, &d_tslots, cs.timg_slots.data(), cs.timg_slots.size()));
                CU(h, upload(h, &d_ba, cs.tblk_a.data(), cs.tblk_a.size()));
                CU(h, upload(h, &d_bb, cs.tblk_b.data(), cs.tblk_b.size()));
                CU(h, upload(h, &d_bptr, cs.tblk_ptr.data(), cs.tblk_ptr.size()));
                CU(h, upload(h, &d_bslots, cs.tblk_slots.data(), cs.tblk_slots.size()));
                CU(h, dev_alloc(h, &D.img_part, cs.slot_img.size() * (size_t)kImgPart));
                CU(h, dev_alloc(h, &D.blk_part, cs.bslot_a.size() * (size_t)kBlkPart));
                CU(h, cudaStreamSynchronize(h->stream));        // the schedule's host vectors go out of scope
                D.n_chunks = cs.n_chunks;
                D.obs0 = d_obs0; D.img0 = d_img0; D.slot_obs0 = d_slot_obs0; D.slot_obs = d_slot_obs;
                D.blk0 = d_blk0; D.bslot_pair0 = d_pair0; D.pairs = d_pairs;
                D.slot_img = d_slot_img; D.timg_ptr = d_tptr; D.timg_slots = d_tslots;
                D.n_tblk = (int)cs.tblk_a.size();
                D.tblk_a = d_ba; D.tblk_b = d_bb; D.tblk_ptr = d_bptr; D.tblk_slots = d_bslots;
                h->n_pairs = cs.n_pairs;
                h->use_chunks = true;
                P.ipos = nullptr;                                // records at the observations' own positions
                if (std::getenv("FEBA_VERBOSE"))
                    fprintf(stderr, "[feba] assembly: %d chunks, %zu image slots, %zu block slots (%d distinct image pairs), "
                                    "%lld observation pairs\n", cs.n_chunks, cs.slot_img.size(), cs.bslot_a.size(),
                            D.n_tblk, cs.n_pairs);
            }
        }
        if (!h->use_chunks) {
            // image-major form: image-pair schedule of the Schur blocks on the device
            void *kp = nullptr, *kb = nullptr;
            CU(h, build_pair_schedule(P, doseg, &h->n_pairs, &kp, &kb, h->stream));
            if (kp) h->allocs.push_back(kp);
            if (kb) h->allocs.push_back(kb);
        }
    }
    // sparse-datum form (feba_sparse.h) on a masked plan of a free network: flags of the datum images
    if (h->plan.masked && P.inner) {
        std::vector<unsigned char> flags((size_t)P.n_img, 0);
        for (int im : h->plan.datum) flags[(size_t)im] = 1;
        CU(h, upload(h, &h->datum_dev, flags.data(), flags.size()));
        CU(h, cudaStreamSynchronize(h->stream));
        P.datum = h->datum_dev;
        P.aug_rows = kSparseAugRows;
    }
    if (h->shard) {
        if (dist_comm_init(&h->dist, rank, world, id)) return fail(h, FEBA_ERR_CUDA, "%s", h->dist.err);
        // packed lower trapezoid of the shared top part: per group of kPackCols block columns the rows from the
        // group's first row down to the end of the augmented block row
        const size_t ld = (size_t)P.ld, w = 8 * (size_t)kBlk;
        size_t total = 0;
        for (size_t c0 = (size_t)h->top_row0; c0 < ld; c0 += w) total += (ld - c0) * std::min(w, ld - c0);
        h->xchg_count = total;
        CU(h, dev_alloc(h, &h->xchg, total));
    }
    // xhat = Buildxhat of the uploaded tables (Buildxhat.m:22-135)
    CU(h, cudaMemsetAsync(h->xhat, 0, (size_t)h->u * sizeof(double), h->stream));
    CU(h, launch_xhat_gather(P, h->sm_count, h->xhat, h->eop, h->iop, h->tie_pt, h->stream));
    ++h->launches;
    CU(h, cudaStreamSynchronize(h->stream));   // host staging vectors go out of scope
    guard.keep = true;
    *out = h;
    return FEBA_OK;
}

int feba_create(const feba_problem* pr, feba_handle** out) { return create_impl(pr, 0, 1, nullptr, out); }

int feba_create_shard(const feba_problem* pr, int32_t rank, int32_t world, const void* id, size_t bytes,
                      feba_handle** out) {
    if (world < 1 || rank < 0 || rank >= world || (world & (world - 1)) != 0)
        return fail(nullptr, FEBA_ERR_INVALID, "feba_create_shard: world must be a power of two and 0 <= rank < world");
    if (world == 1) return create_impl(pr, 0, 1, nullptr, out);
    if (!id || bytes != FEBA_DIST_ID_BYTES)
        return fail(nullptr, FEBA_ERR_INVALID, "feba_create_shard: id must be %d bytes (feba_dist_unique_id)", FEBA_DIST_ID_BYTES);
    return create_impl(pr, rank, world, id, out);
}

int feba_set_stream(feba_handle* h, void* stream) {
    if (!h) return FEBA_ERR_INVALID;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    drop_graphs(h);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    h->stream = static_cast<cudaStream_t>(stream);
    h->own_stream = false;
    // stream capture is not permitted on the legacy default stream: run eagerly there
    if (h->stream == nullptr || h->stream == cudaStreamLegacy) h->use_graph = false;
    return FEBA_OK;
}

int feba_dist_unique_id(void* id, size_t bytes) {
    if (!id || bytes != FEBA_DIST_ID_BYTES) return fail(nullptr, FEBA_ERR_INVALID, "id buffer must be %d bytes", FEBA_DIST_ID_BYTES);
    if (dist_unique_id(id)) return fail(nullptr, FEBA_ERR_CUDA, "%s", dist_load_error());
    return FEBA_OK;
}

int feba_dist_init(feba_handle* h, int32_t rank, int32_t world, const void* id, size_t bytes) {
    if (!h || !id || bytes != FEBA_DIST_ID_BYTES || world < 1 || rank < 0 || rank >= world)
        return fail(h, FEBA_ERR_INVALID, "feba_dist_init: bad arguments");
    if (h->dist.comm) return fail(h, FEBA_ERR_STATE, "feba_dist_init called twice (or on a feba_create_shard handle)");
    if (h->plan.masked && world > 1)
        return fail(h, FEBA_ERR_STATE, "feba_dist_init needs the identity row order on every rank: create the handle with "
                                       "settings.plan = -1 (a nested-dissection plan of a group comes from feba_create_shard)");
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    // the communicator is created by every rank even when this problem keeps the replicated solve
    if (dist_comm_init(&h->dist, rank, world, id)) return fail(h, FEBA_ERR_CUDA, "%s", h->dist.err);
    if (!h->use_dag || world == 1) return FEBA_OK;
    drop_graphs(h);
    {
        const int rc_pool = setup_legacy_group_pool(h);
        if (rc_pool) return rc_pool;
    }
    if (!h->use_dag || h->dag.n_streams < 6) return FEBA_OK;
    int lo = 0, hi = 0;
    CU(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU(h, cudaStreamCreateWithPriority(&h->dist.stream, cudaStreamNonBlocking, hi));
    CU(h, cudaEventCreateWithFlags(&h->dag.join[h->dag.n_streams], cudaEventDisableTiming));
    const size_t tile = (size_t)h->dag.tile_blocks * kBlk;
    h->dist.staging_count = tile * (size_t)(h->P.n_pad + kBlk);      // up to one supertile column
    CU(h, dev_alloc(h, &h->dist.staging, h->dist.staging_count));
    h->dist_active = true;
    return FEBA_OK;
}

int feba_num_unknowns(const feba_handle* h, int64_t* u, int64_t* u_c) {
    if (!h) return FEBA_ERR_INVALID;
    if (u) *u = h->u;
    if (u_c) *u_c = h->P.n_red;
    return FEBA_OK;
}

int64_t feba_num_obs(const feba_handle* h) { return h ? h->n_obs_global : -1; }

int feba_set_xhat(feba_handle* h, const double* xhat, size_t u) {
    if (!h || !xhat) return FEBA_ERR_INVALID;
    if ((int64_t)u != h->u) return fail(h, FEBA_ERR_INVALID, "xhat has %zu entries, expected %lld", u, (long long)h->u);
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaMemcpyAsync(h->xhat, xhat, u * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(h, launch_xhat_scatter(h->P, h->sm_count, h->xhat, h->eop, h->iop, h->tie_pt, h->stream));
    ++h->launches;
    CU(h, cudaStreamSynchronize(h->stream));
    h->iterations = 0;
    h->phase = 0;
    return FEBA_OK;
}

int feba_get_xhat(feba_handle* h, double* xhat, size_t u) {
    if (!h || !xhat) return FEBA_ERR_INVALID;
    if ((int64_t)u != h->u) return fail(h, FEBA_ERR_INVALID, "xhat has %zu entries, expected %lld", u, (long long)h->u);
    CU(h, cudaSetDevice(h->device));
    CU(h, launch_xhat_gather(h->P, h->sm_count, h->xhat, h->eop, h->iop, h->tie_pt, h->stream));
    ++h->launches;
    if (h->shard && h->P.n_tie > 0) {
        // every rank returns the whole vector: tie coordinates come from their owners (collective)
        CU(h, launch_keep_own_ties(h->P.n_tie, h->tie_mine, h->xhat + h->P.n_red, h->stream));
        NC(h, dist_allreduce_f64(&h->dist, h->xhat + h->P.n_red, 3 * (size_t)h->P.n_tie, h->stream));
        ++h->launches;
    }
    CU(h, cudaMemcpyAsync(xhat, h->xhat, u * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return FEBA_OK;
}

// host copy between the caller's xhat (tie part, scattered) and the packed staging buffer, a few threads
static void owned_host_copy(const std::vector<int>& own, double* packed, double* xhat_ties, bool to_packed) {
    const size_t no = own.size();
    const int nth = no > 20000 ? 4 : 1;
    auto work = [&](int t) {
        const size_t lo = no * (size_t)t / nth, hi = no * (size_t)(t + 1) / nth;
        for (size_t i = lo; i < hi; ++i) {
            double* a = packed + 3 * i;
            double* b = xhat_ties + 3 * (size_t)own[i];
            if (to_packed) { a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; }
            else { b[0] = a[0]; b[1] = a[1]; b[2] = a[2]; }
        }
    };
    if (nth == 1) {
        work(0);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 1; t < nth; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
}

// device address of a caller's buffer when it is page-locked host memory the device can address, else null
static double* mapped_host(const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    static const bool off = [] { const char* e = std::getenv("FEBA_ZERO_COPY"); return e && e[0] == '0'; }();
    return off ? nullptr : static_cast<double*>(at.devicePointer);
}

// Owned-only transfers of a group handle (feba_create_shard): the EOP/IOP part and the tie points THIS rank owns.
// A distributed caller keeps xhat split over the ranks between iterations; the full-vector calls above move and
// all-reduce 8 u bytes on every rank.  On a single-GPU handle they are feba_set_xhat / feba_get_xhat.
int feba_set_xhat_owned(feba_handle* h, const double* xhat, size_t u) {
    if (!h || !xhat) return FEBA_ERR_INVALID;
    if (!h->shard) return feba_set_xhat(h, xhat, u);
    if ((int64_t)u != h->u) return fail(h, FEBA_ERR_INVALID, "xhat has %zu entries, expected %lld", u, (long long)h->u);
    CU(h, cudaSetDevice(h->device));
    const size_t nr = (size_t)h->P.n_red, no = h->own_ties.size();
    if (double* dx = mapped_host(xhat)) {
        // page-locked caller buffer: the device reads the owned entries where they lie
        CU(h, cudaMemcpyAsync(h->xhat, xhat, nr * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CU(h, launch_ties_copy_owned((int)no, h->own_ties_dev, dx + nr, h->xhat + nr, h->stream));
        CU(h, launch_xhat_scatter(h->P, h->sm_count, h->xhat, h->eop, h->iop, h->tie_pt, h->stream));
        h->launches += 2;
        CU(h, cudaStreamSynchronize(h->stream));
        h->iterations = 0;
        h->phase = 0;
        return FEBA_OK;
    }
    std::memcpy(h->own_pin, xhat, nr * sizeof(double));
    owned_host_copy(h->own_ties, h->own_pin + nr, const_cast<double*>(xhat) + nr, true);
    CU(h, cudaMemcpyAsync(h->xhat, h->own_pin, nr * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (no) {
        CU(h, cudaMemcpyAsync(h->own_packed, h->own_pin + nr, 3 * no * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CU(h, launch_ties_pack((int)no, h->own_ties_dev, (int64_t)nr, h->xhat, h->own_packed, true, h->stream));
    }
    // coordinates of other ranks' tie points in the device copy of xhat are whatever they were: no observation of
    // this rank uses them, and the collective feba_get_xhat takes every tie from its owner
    CU(h, launch_xhat_scatter(h->P, h->sm_count, h->xhat, h->eop, h->iop, h->tie_pt, h->stream));
    h->launches += 2;
    CU(h, cudaStreamSynchronize(h->stream));
    h->iterations = 0;
    h->phase = 0;
    return FEBA_OK;
}

int feba_get_xhat_owned(feba_handle* h, double* xhat, size_t u) {
    if (!h || !xhat) return FEBA_ERR_INVALID;
    if (!h->shard) return feba_get_xhat(h, xhat, u);
    if ((int64_t)u != h->u) return fail(h, FEBA_ERR_INVALID, "xhat has %zu entries, expected %lld", u, (long long)h->u);
    CU(h, cudaSetDevice(h->device));
    const size_t nr = (size_t)h->P.n_red, no = h->own_ties.size();
    CU(h, launch_xhat_gather(h->P, h->sm_count, h->xhat, h->eop, h->iop, h->tie_pt, h->stream));
    ++h->launches;
    if (double* dx = mapped_host(xhat)) {
        CU(h, cudaMemcpyAsync(xhat, h->xhat, nr * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(h, launch_ties_copy_owned((int)no, h->own_ties_dev, h->xhat + nr, dx + nr, h->stream));
        ++h->launches;
        CU(h, cudaStreamSynchronize(h->stream));
        return FEBA_OK;
    }
    CU(h, cudaMemcpyAsync(h->own_pin, h->xhat, nr * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (no) {
        CU(h, launch_ties_pack((int)no, h->own_ties_dev, (int64_t)nr, h->xhat, h->own_packed, false, h->stream));
        ++h->launches;
        CU(h, cudaMemcpyAsync(h->own_pin + nr, h->own_packed, 3 * no * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    CU(h, cudaStreamSynchronize(h->stream));
    std::memcpy(xhat, h->own_pin, nr * sizeof(double));
    owned_host_copy(h->own_ties, h->own_pin + nr, xhat + nr, false);
    return FEBA_OK;
}

// number of tie points this handle owns (all of them on a single-GPU handle)
int64_t feba_num_owned_ties(const feba_handle* h) {
    if (!h) return -1;
    return h->shard ? (int64_t)h->own_ties.size() : (int64_t)h->P.n_tie;
}

int feba_get_delta(feba_handle* h, double* delta, size_t u) {
    if (!h || !delta) return FEBA_ERR_INVALID;
    if ((int64_t)u != h->u) return fail(h, FEBA_ERR_INVALID, "delta has %zu entries, expected %lld", u, (long long)h->u);
    if (h->iterations < 1) return fail(h, FEBA_ERR_STATE, "feba_get_delta before any iteration");
    CU(h, cudaSetDevice(h->device));
    if (!h->delta_out) CU(h, dev_alloc(h, &h->delta_out, (size_t)h->u));
    CU(h, launch_delta_gather(h->P, h->sm_count, h->delta_out, h->stream));
    ++h->launches;
    if (h->shard && h->P.n_tie > 0)       // increments of other ranks' tie points are zero here
        NC(h, dist_allreduce_f64(&h->dist, h->delta_out + h->P.n_red, 3 * (size_t)h->P.n_tie, h->stream));
    CU(h, cudaMemcpyAsync(delta, h->delta_out, u * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return FEBA_OK;
}

// ---- one Gauss-Newton step, first half: tables, zero S, fused BuildAwG + normal blocks + Schur.
static int enqueue_assemble(feba_handle* h) {
    DevProblem& P = h->P;
    CU(h, record(h, 0));
    CU(h, launch_tables(P, h->eop, h->iop, h->cam_box, h->img_tab, h->cam_tab, h->stream));
    if (h->plan.masked) {
        // only the structurally non-zero blocks are ever written (assembly, fill): the rest stays zero from create
        CU(h, launch_clear_blocks(P, h->blk_list, h->n_blk_list, h->stream));
        ++h->launches;
    } else {
        CU(h, cudaMemsetAsync(P.S, 0, h->S_count * sizeof(double), h->stream));
    }
    CU(h, cudaMemsetAsync(h->info, 0, sizeof(int), h->stream));
    CU(h, record(h, 1));
    ++h->launches;
    CU(h, launch_assemble(P, h->sm_count, h->info, h->stream, &h->launches, h->opt, h->use_chunks ? &h->chunks : nullptr));
    CU(h, record(h, 2));
    return FEBA_OK;
}

// Run one phase eagerly (first call: kernels set their attributes), capture it (second call) or
// replay its graph.
static int run_phase(feba_handle* h, feba_handle::GraphSlot& g, int (*fn)(feba_handle*), bool allow_graph = true) {
    const int call = g.calls++;
    if (!h->use_graph || !allow_graph || call == 0) return fn(h);
    if (!g.exec) {
        const int64_t before = h->launches;
        CU(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        h->capturing = true;
        const int rc = fn(h);
        h->capturing = false;
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        CU(h, e);
        g.launches = h->launches - before;
        const cudaError_t ei = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        CU(h, ei);
        CU(h, cudaGraphLaunch(g.exec, h->stream));
        return FEBA_OK;
    }
    CU(h, cudaGraphLaunch(g.exec, h->stream));
    h->launches += g.launches;
    return FEBA_OK;
}

int feba_iterate_assemble(feba_handle* h) {
    if (!h) return FEBA_ERR_INVALID;
    CU(h, cudaSetDevice(h->device));
    const int rc = run_phase(h, h->g_assemble, enqueue_assemble);
    if (rc) return rc;
    h->phase = 1;
    h->cov_ready = false;
    return FEBA_OK;
}

int feba_reduced_dev(feba_handle* h, double** dev_ptr, size_t* count) {
    if (!h || !dev_ptr || !count) return FEBA_ERR_INVALID;
    *dev_ptr = h->P.S;
    *count = h->S_count;
    return FEBA_OK;
}

// Packed form of the exchange buffer (opt-in, FEBA_PACKED_REDUCE=1 in shard.py): the solve half reads the
// lower triangle and the augmented block row only, so per group of kPackCols 64-column blocks the rows from
// the group's first row down to the end of the buffer are copied (strided device copies, no kernel) into
// one contiguous buffer of ~(1/2 + 1/(2 groups)) of the square; the caller sums THAT across ranks and
// feba_reduced_unpack copies it back.  Entries above a group's first row keep this rank's partial sums:
// nothing reads them.
static constexpr int kPackCols = 8;

int feba_reduced_pack(feba_handle* h, double** dev_ptr, size_t* count) {
    if (!h || !dev_ptr || !count) return FEBA_ERR_INVALID;
    if (h->phase != 1) return fail(h, FEBA_ERR_STATE, "feba_reduced_pack needs a pending feba_iterate_assemble");
    CU(h, cudaSetDevice(h->device));
    const DevProblem& P = h->P;
    const size_t ld = (size_t)P.ld, w = (size_t)kPackCols * kBlk;
    if (!h->packed) {
        size_t total = 0;
        for (size_t c0 = 0; c0 < ld; c0 += w) total += (ld - c0) * std::min(w, ld - c0);
        h->packed_count = total;
        CU(h, dev_alloc(h, &h->packed, total));
    }
    size_t off = 0;
    for (size_t c0 = 0; c0 < ld; c0 += w) {
        const size_t rows = ld - c0, cols = std::min(w, ld - c0);
        CU(h, cudaMemcpy2DAsync(h->packed + off, rows * sizeof(double), P.S + c0 + ld * c0, ld * sizeof(double),
                                rows * sizeof(double), cols, cudaMemcpyDeviceToDevice, h->stream));
        off += rows * cols;
    }
    *dev_ptr = h->packed;
    *count = h->packed_count;
    return FEBA_OK;
}

int feba_reduced_unpack(feba_handle* h) {
    if (!h) return FEBA_ERR_INVALID;
    if (h->phase != 1 || !h->packed) return fail(h, FEBA_ERR_STATE, "feba_reduced_unpack without feba_reduced_pack");
    CU(h, cudaSetDevice(h->device));
    const DevProblem& P = h->P;
    const size_t ld = (size_t)P.ld, w = (size_t)kPackCols * kBlk;
    size_t off = 0;
    for (size_t c0 = 0; c0 < ld; c0 += w) {
        const size_t rows = ld - c0, cols = std::min(w, ld - c0);
        CU(h, cudaMemcpy2DAsync(P.S + c0 + ld * c0, ld * sizeof(double), h->packed + off, rows * sizeof(double),
                                rows * sizeof(double), cols, cudaMemcpyDeviceToDevice, h->stream));
        off += rows * cols;
    }
    return FEBA_OK;
}

// Copy the lower trapezoid of S from row/column `first` on (incl. the augmented block row) into one contiguous
// buffer (pack) or back (unpack): per group of 8 block columns the rows from the group's first row down.
static int trapezoid_copy(feba_handle* h, double* buf, size_t first, bool pack) {
    const DevProblem& P = h->P;
    const size_t ld = (size_t)P.ld, w = 8 * (size_t)kBlk;
    size_t off = 0;
    for (size_t c0 = first; c0 < ld; c0 += w) {
        const size_t rows = ld - c0, cols = std::min(w, ld - c0);
        double* sp = P.S + c0 + ld * c0;
        if (pack)
            CU(h, cudaMemcpy2DAsync(buf + off, rows * sizeof(double), sp, ld * sizeof(double), rows * sizeof(double), cols,
                                    cudaMemcpyDeviceToDevice, h->stream));
        else
            CU(h, cudaMemcpy2DAsync(sp, ld * sizeof(double), buf + off, rows * sizeof(double), rows * sizeof(double), cols,
                                    cudaMemcpyDeviceToDevice, h->stream));
        off += rows * cols;
    }
    return FEBA_OK;
}

static TileView plan_view(const feba_handle* h) {
    TileView V;
    V.NT = h->plan.NT;
    V.b0 = h->plan.tile_b0.data();
    V.nz = h->plan.masked ? h->plan.nz.data() : nullptr;
    V.chain = h->plan.masked ? h->tile_chain.data() : nullptr;
    V.owner = h->shard ? h->tile_owner.data() : nullptr;
    return V;
}

// ---- second half: inner-constraint border, factorisation, solve, update, back-substitution.
// Group on a nested-dissection plan (feba_create_shard), exchanges on the handle's stream (NCCL):
//   (1) diag(S), n_pad doubles: conditioning of G and Jacobi scaling must be identical on every rank;
//   (2) after every rank has eliminated its own subtrees: the shared top part of S (lower trapezoid from
//       top_row0 on, incl. the augmented block row) -- the only large message of an iteration;
//   (3) the solution of the reduced system, n_pad doubles (own rows + rank 0's copy of the shared rows).
static int enqueue_solve_pre(feba_handle* h) {
    DevProblem& P = h->P;
    CU(h, launch_border_prepare(P, h->eop, h->dg, h->stream, &h->launches));
    if (h->shard) NC(h, dist_allreduce_f64(&h->dist, h->dg, (size_t)P.n_pad, h->stream));
    CU(h, launch_border_scale(P, h->dg, h->dvec, h->info, h->blk_list, h->n_blk_scale, h->gwork, h->stream, &h->launches));
    return FEBA_OK;
}

static int enqueue_solve_post(feba_handle* h) {
    DevProblem& P = h->P;
    CU(h, launch_update_cam(P, h->sol, h->dvec, h->dcam, h->dcam_unscaled, h->eop, h->iop, h->scal, h->stream));
    ++h->launches;
    if (P.n_tie > 0 && P.n_seg > 0) {
        CU(h, launch_backsub(P, h->sm_count, h->stream));
        CU(h, launch_sum_partials(P.partial, backsub_warps(P, h->sm_count), 1, 0, h->scal + 1, h->stream));
        h->launches += 2;
    } else {
        CU(h, cudaMemsetAsync(h->scal + 1, 0, sizeof(double), h->stream));
    }
    CU(h, record(h, 5));
    return FEBA_OK;
}

static int enqueue_solve(feba_handle* h) {
    DevProblem& P = h->P;
    const int nb = P.n_pad / kBlk;
    {
        const int rc_pre = enqueue_solve_pre(h);
        if (rc_pre) return rc_pre;
    }
    if (h->use_dag && h->dag_cols) {
        const cudaError_t ed = chol_cols(P.S, P.ld, nb, h->Linv, h->info, h->dag, h->dist_active ? &h->dist : nullptr,
                                         h->stream, &h->launches);
        if (ed == cudaErrorUnknown && h->dist.err[0]) return fail(h, FEBA_ERR_CUDA, "%s", h->dist.err);
        CU(h, ed);
    } else if (h->dist_active) {
        const cudaError_t ed = chol_dag_dist(P.S, P.ld, nb, h->Linv, h->info, h->dag, h->dist, h->stream, &h->launches);
        if (ed == cudaErrorUnknown && h->dist.err[0]) return fail(h, FEBA_ERR_CUDA, "%s", h->dist.err);
        CU(h, ed);
    } else if (h->use_dag) {
        const TileView V = plan_view(h);
        const int top = h->shard ? h->plan.top_tile0 : V.NT;
        CU(h, chol_tiles(P.S, P.ld, nb, h->Linv, h->info, h->dag, h->stream, &h->launches, V, 0, top, h->rank));
        if (h->shard) {
            CU(h, record(h, 6));
            const int rc = trapezoid_copy(h, h->xchg, (size_t)h->top_row0, true);
            if (rc) return rc;
            NC(h, dist_allreduce_f64(&h->dist, h->xchg, h->xchg_count, h->stream));
            const int rc2 = trapezoid_copy(h, h->xchg, (size_t)h->top_row0, false);
            if (rc2) return rc2;
            CU(h, record(h, 7));
            CU(h, chol_tiles(P.S, P.ld, nb, h->Linv, h->info, h->dag, h->stream, &h->launches, V, top, V.NT, h->rank));
        }
    } else CU(h, chol_augmented(P.S, P.ld, nb, h->Linv, h->info, h->stream, &h->launches));
    CU(h, record(h, 3));
    {
        const TileView Vb = plan_view(h);
        CU(h, border_and_backsolve(P.S, P.ld, nb, h->Linv, P.inner, h->work, h->ywork, h->sol, h->info, h->sm_count,
                                   h->stream, &h->launches, P.datum != nullptr, h->plan.masked ? &Vb : nullptr,
                                   h->shard ? h->block_owner.data() : nullptr, h->rank));
    }
    if (h->shard) {
        CU(h, launch_keep_own_rows(P, h->sol, h->stream));
        ++h->launches;
        NC(h, dist_allreduce_f64(&h->dist, h->sol, (size_t)P.n_pad, h->stream));
    }
    CU(h, record(h, 4));
    return enqueue_solve_post(h);
}

static int solve_async(feba_handle* h) {
    if (h->phase != 1) return fail(h, FEBA_ERR_STATE, "feba_iterate_solve without feba_iterate_assemble");
    const int rc = run_phase(h, h->g_solve, enqueue_solve, h->solve_graph);
    if (rc) return rc;
    h->phase = 0;
    ++h->iterations;
    return FEBA_OK;
}

static int finish_iteration(feba_handle* h, double* dcam_sum, double* dpts_sum) {
    if (h->shard) {
        // sum|delta| over the tie points of all ranks (main.m:487) and the worst status: collective, on the stream
        NC(h, dist_allreduce_f64(&h->dist, h->scal + 1, 1, h->stream));
        NC(h, dist_allreduce_max_i32(&h->dist, h->info, 1, h->stream));
    }
    CU(h, cudaMemcpyAsync(h->scal_host, h->scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(h->info_host, h->info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    bool ok = true;
    double tot = 0.0;
    for (int i = 0; i < 5 && ok; ++i) {
        float ms = 0.f;
        ok = cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]) == cudaSuccess;
        h->timing[i] = ms;
        tot += ms;
    }
    if (ok) h->timing[5] = tot;
    h->timing[6] = 0.0;
    if (ok && h->shard && h->use_dag) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]) == cudaSuccess) h->timing[6] = ms;
        else cudaGetLastError();
    }
    h->timing_valid = ok;
    if (h->dist_active && !h->dag_cols && h->g_solve.exec && h->g_solve.calls >= 5) dist_prof_report();
    if (h->g_solve.exec && h->g_solve.calls >= 5) chain_prof_report();
    if (dcam_sum) *dcam_sum = h->scal_host[0];
    if (dpts_sum) *dpts_sum = h->scal_host[1];
    if (*h->info_host == 1)
        return fail(h, FEBA_ERR_NUMERIC, "reduced normal matrix is not positive definite");
    if (*h->info_host == 2) return fail(h, FEBA_ERR_NUMERIC, "inner-constraint border system is singular");
    if (*h->info_host == 3)
        return fail(h, FEBA_ERR_INVALID, "an object point is observed by more than 4 different cameras (unsupported)");
    if (!std::isfinite(h->scal_host[0]) || !std::isfinite(h->scal_host[1]))
        return fail(h, FEBA_ERR_NUMERIC, "non-finite increment (R = 0 on the optical axis? BuildAwG.m:186)");
    return FEBA_OK;
}

int feba_iterate_solve(feba_handle* h, double* deltasum_cam, double* deltasum_pts) {
    if (!h) return FEBA_ERR_INVALID;
    CU(h, cudaSetDevice(h->device));
    int rc = solve_async(h);
    if (rc) return rc;
    return finish_iteration(h, deltasum_cam, deltasum_pts);
}

int feba_iterate_solve_async(feba_handle* h) {
    if (!h) return FEBA_ERR_INVALID;
    CU(h, cudaSetDevice(h->device));
    return solve_async(h);
}

int feba_iterate_async(feba_handle* h) {
    int rc = feba_iterate_assemble(h);
    if (rc) return rc;
    return solve_async(h);
}

int feba_sync(feba_handle* h, double* deltasum_out) {
    if (!h) return FEBA_ERR_INVALID;
    CU(h, cudaSetDevice(h->device));
    double a = 0, b = 0;
    int rc = finish_iteration(h, &a, &b);
    if (deltasum_out) *deltasum_out = a + b;      // main.m:487
    return rc;
}

int feba_iterate(feba_handle* h, double* deltasum_out) {
    int rc = feba_iterate_async(h);
    if (rc) return rc;
    return feba_sync(h, deltasum_out);
}

// ---- BatchRun sweep (BatchRun.m:57-65): many small independent adjustments on one GPU.
// One Gauss-Newton step of EVERY handle of the batch is captured into ONE CUDA graph (the handles' streams fork
// from the first handle's stream and join it again), so a step of the whole sweep costs one graph launch on the
// host instead of two per block; on the device the blocks run side by side.  Each block's arithmetic is untouched:
// results are bit-identical to feba_iterate per handle.
struct feba_batch {
    std::vector<feba_handle*> hs;
    // systems of one shape and small enough for the single-chain form: factorisation and backward substitution
    // of ALL of them run as batched launches (blockIdx.y = system), nb + 1 and nb launches per step in total
    bool fused = false;
    double **dA = nullptr, **dLinv = nullptr, **dy = nullptr, **dsol = nullptr;
    int** dinfo = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaEvent_t fork = nullptr, done = nullptr;
    std::vector<cudaEvent_t> join;
    std::vector<int64_t> launches;        // kernel launches of one step, per handle
};

void feba_batch_destroy(feba_batch* b) {
    if (!b) return;
    if (b->exec) cudaGraphExecDestroy(b->exec);
    if (b->fork) cudaEventDestroy(b->fork);
    if (b->done) cudaEventDestroy(b->done);
    for (void* p : {(void*)b->dA, (void*)b->dLinv, (void*)b->dy, (void*)b->dsol, (void*)b->dinfo})
        if (p) cudaFree(p);
    for (auto& e : b->join)
        if (e) cudaEventDestroy(e);
    delete b;
}

int feba_batch_create(feba_handle* const* handles, int32_t n, feba_batch** out) {
    if (!handles || !out || n < 1) return fail(nullptr, FEBA_ERR_INVALID, "feba_batch_create: bad arguments");
    *out = nullptr;
    for (int i = 0; i < n; ++i) {
        feba_handle* h = handles[i];
        if (!h || h->shard || h->dist_active || h->device != handles[0]->device)
            return fail(nullptr, FEBA_ERR_INVALID, "feba_batch_create: handle %d is null, on another device or part of a group", i);
        if (h->stream == nullptr || h->stream == cudaStreamLegacy || !h->use_graph)
            return fail(nullptr, FEBA_ERR_INVALID, "feba_batch_create: handle %d cannot be captured (legacy default stream or FEBA_NO_GRAPH)", i);
        for (int j = 0; j < i; ++j)
            if (handles[j] == h || handles[j]->stream == h->stream)
                return fail(nullptr, FEBA_ERR_INVALID, "feba_batch_create: handles %d and %d are the same or share a stream", j, i);
    }
    feba_batch* b = new (std::nothrow) feba_batch();
    if (!b) return fail(nullptr, FEBA_ERR_INVALID, "out of host memory");
    b->hs.assign(handles, handles + n);
    b->join.assign((size_t)n, nullptr);
    b->launches.assign((size_t)n, 0);
    bool ok = cudaSetDevice(handles[0]->device) == cudaSuccess &&
              cudaEventCreateWithFlags(&b->fork, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&b->done, cudaEventDisableTiming) == cudaSuccess;
    for (auto& e : b->join) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        feba_batch_destroy(b);
        return fail(nullptr, FEBA_ERR_CUDA, "feba_batch_create: event creation failed");
    }
    // one shape, single-chain form, dense datum: the factorisation can be batched
    b->fused = n > 1 && !(std::getenv("FEBA_BATCH_FUSED") && std::atoi(std::getenv("FEBA_BATCH_FUSED")) == 0);
    for (int i = 0; i < n && b->fused; ++i) {
        const feba_handle* h = handles[i];
        b->fused = !h->use_dag && !h->plan.masked && h->P.n_pad == handles[0]->P.n_pad && h->P.inner == handles[0]->P.inner &&
                   h->P.datum == nullptr && h->P.n_pad / kBlk <= 32;
    }
    if (b->fused) {
        std::vector<double*> vA, vL, vy, vs;
        std::vector<int*> vi;
        for (int i = 0; i < n; ++i) {
            vA.push_back(handles[i]->P.S);
            vL.push_back(handles[i]->Linv);
            vy.push_back(handles[i]->ywork);
            vs.push_back(handles[i]->sol);
            vi.push_back(handles[i]->info);
        }
        auto up = [&](void** dst, const void* src, size_t bytes) {
            return cudaMalloc(dst, bytes) == cudaSuccess && cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
        };
        const size_t pb = (size_t)n * sizeof(void*);
        if (!(up((void**)&b->dA, vA.data(), pb) && up((void**)&b->dLinv, vL.data(), pb) && up((void**)&b->dy, vy.data(), pb) &&
              up((void**)&b->dsol, vs.data(), pb) && up((void**)&b->dinfo, vi.data(), pb))) {
            feba_batch_destroy(b);
            return fail(nullptr, FEBA_ERR_CUDA, "feba_batch_create: device allocation failed");
        }
    }
    *out = b;
    return FEBA_OK;
}

int feba_batch_iterate_async(feba_batch* b) {
    if (!b) return FEBA_ERR_INVALID;
    feba_handle* h0 = b->hs[0];
    CU(h0, cudaSetDevice(h0->device));
    const size_t n = b->hs.size();
    bool warm = true;
    for (feba_handle* h : b->hs) {
        if (h->phase != 0) return fail(h0, FEBA_ERR_STATE, "feba_batch_iterate_async: a handle has a pending feba_iterate_assemble");
        warm = warm && h->g_assemble.calls > 0 && h->g_solve.calls > 0;
    }
    if (!warm) {      // first step of a handle runs eagerly (kernel attributes are set on first use): no capture yet
        for (feba_handle* h : b->hs) {
            const int rc = feba_iterate_async(h);
            if (rc) return fail(h0, rc, "%s", h->err.c_str());
        }
        return FEBA_OK;
    }
    // the step starts after whatever the handles have pending on their own streams
    for (size_t i = 1; i < n; ++i) {
        CU(h0, cudaEventRecord(b->join[i], b->hs[i]->stream));
        CU(h0, cudaStreamWaitEvent(h0->stream, b->join[i], 0));
    }
    if (!b->exec) {
        CU(h0, cudaStreamBeginCapture(h0->stream, cudaStreamCaptureModeThreadLocal));
        int rc = FEBA_OK;
        cudaError_t e = cudaEventRecord(b->fork, h0->stream);
        for (size_t i = 1; i < n && e == cudaSuccess; ++i) e = cudaStreamWaitEvent(b->hs[i]->stream, b->fork, 0);
        std::vector<int64_t> before(n);
        for (size_t i = 0; i < n; ++i) {
            before[i] = b->hs[i]->launches;
            b->hs[i]->capturing = true;
        }
        // every handle's stream runs `stage`, then the first handle's stream waits for all of them
        auto all_handles = [&](int (*stage)(feba_handle*)) {
            for (size_t i = 0; i < n && e == cudaSuccess && rc == FEBA_OK; ++i) rc = stage(b->hs[i]);
            for (size_t i = 1; i < n && e == cudaSuccess && rc == FEBA_OK; ++i) {
                e = cudaEventRecord(b->join[i], b->hs[i]->stream);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(h0->stream, b->join[i], 0);
            }
        };
        // the other streams continue after what the first one has done so far
        auto release = [&]() {
            if (e == cudaSuccess && rc == FEBA_OK) e = cudaEventRecord(b->fork, h0->stream);
            for (size_t i = 1; i < n && e == cudaSuccess && rc == FEBA_OK; ++i) e = cudaStreamWaitEvent(b->hs[i]->stream, b->fork, 0);
        };
        if (!b->fused) {
            all_handles([](feba_handle* h) {
                const int r1 = enqueue_assemble(h);
                return r1 ? r1 : enqueue_solve(h);
            });
        } else {
            const int nb = h0->P.n_pad / kBlk;
            all_handles([](feba_handle* h) {
                const int r1 = enqueue_assemble(h);
                return r1 ? r1 : enqueue_solve_pre(h);
            });
            if (e == cudaSuccess && rc == FEBA_OK)
                e = chol_augmented_batched(b->dA, h0->P.ld, nb, b->dLinv, b->dinfo, (int)n, h0->stream, &h0->launches);
            release();
            all_handles([](feba_handle* h) {
                cudaError_t e3 = record(h, 3);
                if (e3 == cudaSuccess)
                    e3 = border_and_combine(h->P.S, h->P.ld, h->P.n_pad / kBlk, h->P.inner, h->work, h->ywork, h->info, h->stream,
                                            &h->launches, 0);
                return e3 == cudaSuccess ? (int)FEBA_OK : fail(h, FEBA_ERR_CUDA, "border stage: %s", cudaGetErrorString(e3));
            });
            if (e == cudaSuccess && rc == FEBA_OK)
                e = backsolve_batched(b->dA, h0->P.ld, nb, b->dLinv, b->dy, b->dsol, (int)n, h0->sm_count, h0->stream, &h0->launches);
            release();
            all_handles([](feba_handle* h) {
                const cudaError_t e4 = record(h, 4);
                if (e4 != cudaSuccess) return fail(h, FEBA_ERR_CUDA, "event: %s", cudaGetErrorString(e4));
                return enqueue_solve_post(h);
            });
        }
        for (size_t i = 0; i < n; ++i) {
            b->hs[i]->capturing = false;
            b->launches[i] = b->hs[i]->launches - before[i];
        }
        cudaGraph_t graph = nullptr;
        const cudaError_t ee = cudaStreamEndCapture(h0->stream, &graph);
        if (rc || e != cudaSuccess || ee != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            if (rc) return rc;
            return fail(h0, FEBA_ERR_CUDA, "feba_batch_iterate_async: capture failed: %s",
                        cudaGetErrorString(e != cudaSuccess ? e : ee));
        }
        const cudaError_t ei = cudaGraphInstantiate(&b->exec, graph, 0);
        cudaGraphDestroy(graph);
        CU(h0, ei);
    } else {
        for (size_t i = 0; i < n; ++i) b->hs[i]->launches += b->launches[i];
    }
    CU(h0, cudaGraphLaunch(b->exec, h0->stream));
    // feba_sync(h) waits on h's own stream: order it after the step
    CU(h0, cudaEventRecord(b->done, h0->stream));
    for (size_t i = 1; i < n; ++i) CU(h0, cudaStreamWaitEvent(b->hs[i]->stream, b->done, 0));
    for (feba_handle* h : b->hs) {
        ++h->iterations;
        h->cov_ready = false;
    }
    return FEBA_OK;
}

int feba_solve(feba_handle* h, int32_t* iterations_out, double* trace_out, size_t trace_cap) {
    if (!h) return FEBA_ERR_INVALID;
    // main.m:407-494: deltasum = 100; while deltasum > threshold; ...; if count >= cap, break
    double deltasum = 100.0;
    int count = 0;
    while (deltasum > h->cfg.threshold) {
        ++count;
        int rc = feba_iterate(h, &deltasum);
        if (rc) {
            if (iterations_out) *iterations_out = count;
            return rc;
        }
        if (trace_out && (size_t)(count - 1) < trace_cap) trace_out[count - 1] = deltasum;
        if (count >= h->cfg.iteration_cap) break;
    }
    if (iterations_out) *iterations_out = count;
    return FEBA_OK;
}

// Device -> host copy of a large result.  A destination in pinned (page-locked or registered) memory is written
// by one DMA; pageable memory goes through two pinned staging buffers: chunk k+1 crosses PCIe on the copy stream
// while the host moves chunk k into the caller's array (the driver's own pageable path serialises the two and
// reached 2.2 GB/s on the 560 MB of BASELINE configs[3]: 249 of the 250 ms of the residual stage).
static int copy_out(feba_handle* h, double* dst, const double* src_dev, size_t count) {
    if (count == 0) return FEBA_OK;
    cudaPointerAttributes at{};
    const bool pinned = cudaPointerGetAttributes(&at, dst) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) {
        CU(h, cudaMemcpyAsync(dst, src_dev, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        return FEBA_OK;
    }
    constexpr size_t kChunk = (size_t)4 << 20;                 // doubles per staging buffer (32 MB)
    if (!h->v_pin) {
        CU(h, cudaMallocHost((void**)&h->v_pin, kChunk * sizeof(double)));
        CU(h, cudaMallocHost((void**)&h->rsd_pin, kChunk * sizeof(double)));
        CU(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (auto& e : h->copy_ev) CU(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->pin_chunk = kChunk;
    }
    CU(h, cudaStreamSynchronize(h->stream));                   // the result is complete
    double* pin[2] = {h->v_pin, h->rsd_pin};
    const size_t nchunk = (count + kChunk - 1) / kChunk;
    auto issue = [&](size_t c) -> cudaError_t {
        const size_t off = c * kChunk, len = std::min(kChunk, count - off);
        cudaError_t e = cudaMemcpyAsync(pin[c & 1], src_dev + off, len * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(h->copy_ev[c & 1], h->copy_stream);
        return e;
    };
    CU(h, issue(0));
    for (size_t c = 0; c < nchunk; ++c) {
        CU(h, cudaEventSynchronize(h->copy_ev[c & 1]));
        if (c + 1 < nchunk) CU(h, issue(c + 1));               // the other buffer: its last reader finished an iteration ago
        const size_t off = c * kChunk, len = std::min(kChunk, count - off);
        std::memcpy(dst + off, pin[c & 1], len * sizeof(double));
    }
    return FEBA_OK;
}

int feba_residuals(feba_handle* h, double* v, double* rsd, double stats[6]) {
    if (!h) return FEBA_ERR_INVALID;
    if (h->iterations < 1) return fail(h, FEBA_ERR_STATE, "feba_residuals before any iteration (main.m:569 uses the last A, w, delta)");
    if (h->phase != 0)
        return fail(h, FEBA_ERR_STATE, "feba_residuals between feba_iterate_assemble and feba_iterate_solve: the tables "
                                       "already belong to the next linearisation point");
    CU(h, cudaSetDevice(h->device));
    DevProblem& P = h->P;
    const size_t n = (size_t)h->n_obs_global;               // rows of the PHO table (a shard fills its own rows)
    // a group computes v / RSD on every rank when ANY output is wanted anywhere: callers pass the same arguments
    if ((v || h->shard) && !h->v_out) CU(h, dev_alloc(h, &h->v_out, 2 * n));
    if ((rsd || h->shard) && !h->rsd_out) CU(h, dev_alloc(h, &h->rsd_out, 5 * n));
    double* vd = (v || h->shard) ? h->v_out : nullptr;
    double* rd = (rsd || h->shard) ? h->rsd_out : nullptr;
    if (h->shard) {
        CU(h, cudaMemsetAsync(vd, 0, 2 * n * sizeof(double), h->stream));
        CU(h, cudaMemsetAsync(rd, 0, 5 * n * sizeof(double), h->stream));
    }
    CU(h, record(h, 6));
    CU(h, launch_residuals(P, h->sm_count, h->opt, P.xyz_prev, h->iop, vd, rd, h->stream));
    const int nblk = residual_blocks(P, h->sm_count);
    CU(h, launch_sum_partials(P.partial, nblk, 2, 0, h->scal + 2, h->stream));
    CU(h, launch_sum_partials(P.partial, nblk, 2, 1, h->scal + 3, h->stream));
    h->launches += 3;
    CU(h, record(h, 7));
    if (h->shard) {
        // every rank returns the whole result: rows of other ranks' observations come from their owners
        NC(h, dist_allreduce_f64(&h->dist, h->scal + 2, 2, h->stream));
        NC(h, dist_allreduce_f64(&h->dist, vd, 2 * n, h->stream));
        NC(h, dist_allreduce_f64(&h->dist, rd, 5 * n, h->stream));
    }
    CU(h, cudaMemcpyAsync(h->scal_host + 2, h->scal + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (v) {
        const int rc = copy_out(h, v, vd, 2 * n);
        if (rc) return rc;
    }
    if (rsd) {
        const int rc = copy_out(h, rsd, rd, 5 * n);
        if (rc) return rc;
    }
    CU(h, cudaStreamSynchronize(h->stream));
    {
        float ms = 0.f;
        h->timing[7] = cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]) == cudaSuccess ? ms : 0.0;
        cudaGetLastError();
    }
    if (stats) {
        const double sxx = h->scal_host[2], syy = h->scal_host[3];
        const double nn = (double)h->n_obs_global;
        const double rmsx = std::sqrt(sxx / nn), rmsy = std::sqrt(syy / nn);     // main.m:594-597, :998-1002
        stats[0] = rmsx;
        stats[1] = rmsy;
        stats[2] = std::sqrt(rmsx * rmsx + rmsy * rmsy);                         // main.m:598
        stats[3] = (sxx * P.px + syy * P.py) / (2.0 * nn - (double)h->u);         // main.m:601 (n - u)
        stats[4] = sxx;
        stats[5] = syy;
    }
    return FEBA_OK;
}

// ---- covariance stage
int feba_cov_prepare(feba_handle* h) {
    if (!h) return FEBA_ERR_INVALID;
    if (h->iterations < 1 || h->phase != 0)
        return fail(h, FEBA_ERR_STATE, "feba_cov_prepare needs a completed iteration (Cx comes from the last N, main.m:432-444)");
    if (h->cov_ready) return FEBA_OK;
    if (h->shard)
        return fail(h, FEBA_ERR_STATE, "the covariance stage is not available on a feba_create_shard handle: run it on one "
                                       "GPU (feba_create) from the converged xhat");
    if (h->P.datum)
        return fail(h, FEBA_ERR_STATE, "the covariance stage needs the dense datum form: create the handle with "
                                       "settings.plan = -1 (identity order) for a free network");
    CU(h, cudaSetDevice(h->device));
    DevProblem& P = h->P;
    const size_t nn = (size_t)P.n_pad * (size_t)P.n_pad;
    if (!h->covU) {
        CU(h, dev_alloc(h, &h->covU, nn));
        CU(h, dev_alloc(h, &h->covQ, nn));
        CU(h, dev_alloc(h, &h->covY, (size_t)P.n_pad * 8));
        CU(h, dev_alloc(h, &h->covT, 64));
    }
    CU(h, chol_inverse(P.S, P.ld, P.n_pad / kBlk, h->Linv, P.inner, h->covU, h->covQ, h->covY, h->covT, h->info,
                       h->stream, &h->launches));
    CU(h, cudaMemcpyAsync(h->info_host, h->info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (*h->info_host) return fail(h, FEBA_ERR_NUMERIC, "covariance stage: singular border system");
    h->cov_ready = true;
    return FEBA_OK;
}

int feba_cov_diag(feba_handle* h, double* qdiag, size_t u) {
    if (!h || !qdiag) return FEBA_ERR_INVALID;
    if ((int64_t)u != h->u) return fail(h, FEBA_ERR_INVALID, "qdiag has %zu entries, expected %lld", u, (long long)h->u);
    int rc = feba_cov_prepare(h);
    if (rc) return rc;
    DevProblem& P = h->P;
    if (!h->delta_out) CU(h, dev_alloc(h, &h->delta_out, (size_t)h->u));
    // tie entries without observations keep NaN (the reference's N is singular for them)
    CU(h, cudaMemsetAsync(h->delta_out, 0xff, (size_t)h->u * sizeof(double), h->stream));
    CU(h, launch_cov_diag_cam(P, h->covQ, h->covY, h->covT, h->dvec, h->delta_out, h->stream));
    ++h->launches;
    if (P.n_tie > 0 && P.n_seg > 0) {
        CU(h, launch_cov_points(P, h->sm_count, h->covQ, h->covY, h->covT, h->dvec, h->delta_out + P.n_red, h->stream));
        ++h->launches;
    }
    CU(h, cudaMemcpyAsync(qdiag, h->delta_out, (size_t)h->u * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return FEBA_OK;
}

int feba_cov_block(feba_handle* h, const int64_t* idx, int32_t k, double* out) {
    if (!h || !idx || !out || k < 1) return FEBA_ERR_INVALID;
    for (int i = 0; i < k; ++i)
        if (idx[i] < 0 || idx[i] >= h->P.n_red)
            return fail(h, FEBA_ERR_INVALID, "feba_cov_block: index %lld is not an EOP/IOP unknown", (long long)idx[i]);
    int rc = feba_cov_prepare(h);
    if (rc) return rc;
    long long* didx = nullptr;
    double* dout = nullptr;
    CU(h, cudaMalloc((void**)&didx, (size_t)k * sizeof(long long)));
    cudaError_t e = cudaMalloc((void**)&dout, (size_t)k * k * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(didx, idx, (size_t)k * sizeof(long long), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = launch_cov_block(h->P, h->covQ, h->covY, h->covT, h->dvec, didx, k, dout, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, (size_t)k * k * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(didx);
    if (dout) cudaFree(dout);
    ++h->launches;
    CU(h, e);
    return FEBA_OK;
}

int feba_last_timing(const feba_handle* h, double ms[6]) {
    if (!h || !ms) return FEBA_ERR_INVALID;
    if (!h->timing_valid) return FEBA_ERR_STATE;
    for (int i = 0; i < 6; ++i) ms[i] = h->timing[i];
    return FEBA_OK;
}

int feba_last_timing_ex(const feba_handle* h, double ms[8]) {
    if (!h || !ms) return FEBA_ERR_INVALID;
    if (!h->timing_valid) return FEBA_ERR_STATE;
    for (int i = 0; i < 8; ++i) ms[i] = h->timing[i];
    return FEBA_OK;
}

int64_t feba_launch_count(const feba_handle* h) { return h ? h->launches : 0; }

int feba_sparse_info(const feba_handle* h, int32_t info[4]) {
    if (!h || !info) return FEBA_ERR_INVALID;
    const ReducedPlan& pl = h->plan;
    info[0] = pl.masked ? 1 : 0;
    info[1] = info[2] = 0;
    for (int i = 0; i < pl.NT; ++i)
        for (int j = 0; j <= i; ++j) {
            ++info[2];
            if (pl.at(i, j)) ++info[1];
        }
    info[3] = h->P.datum ? (int32_t)pl.datum.size() : 0;
    return FEBA_OK;
}

int feba_plan_info(const feba_handle* h, int32_t info[8], double flop[2]) {
    if (!h || !info) return FEBA_ERR_INVALID;
    const ReducedPlan& pl = h->plan;
    info[0] = pl.masked ? 1 : 0;
    info[1] = pl.n_pad;
    info[2] = pl.NT;
    info[3] = (int32_t)pl.nodes.size();
    info[4] = pl.chain_blocks;
    info[5] = h->world;
    info[6] = h->shard ? h->top_row0 : pl.n_pad;
    info[7] = (int32_t)h->P.n_obs;
    if (flop) {
        flop[0] = pl.flop;
        flop[1] = pl.flop_dense;
    }
    return FEBA_OK;
}

// Diagnostic: the reduced camera system as assembled by the last feba_iterate_assemble (before the
// inner-constraint border is added): S_out n_red x n_red column-major, full symmetric; g_out n_red.
int feba_debug_reduced(feba_handle* h, double* S_out, double* g_out) {
    if (!h) return FEBA_ERR_INVALID;
    if (h->phase != 1) return fail(h, FEBA_ERR_STATE, "feba_debug_reduced needs a pending feba_iterate_assemble");
    CU(h, cudaSetDevice(h->device));
    const DevProblem& P = h->P;
    const size_t nr = (size_t)P.n_red;
    // xhat order -> rows of the plan
    std::vector<int> row(nr);
    for (int r = 0; r < P.n_pad; ++r)
        if (h->plan.row_ext[(size_t)r] >= 0) row[(size_t)h->plan.row_ext[(size_t)r]] = r;
    std::vector<double> col((size_t)P.ld);
    CU(h, cudaStreamSynchronize(h->stream));
    for (size_t c = 0; c < nr; ++c) {
        const size_t rc = (size_t)row[c];
        CU(h, cudaMemcpy(col.data(), P.S + (size_t)P.ld * rc, (size_t)P.ld * sizeof(double), cudaMemcpyDeviceToHost));
        if (S_out)
            for (size_t r = 0; r < nr; ++r)
                if ((size_t)row[r] >= rc) {                  // stored lower triangle: rows at or below the diagonal
                    S_out[r + nr * c] = col[(size_t)row[r]];
                    S_out[c + nr * r] = col[(size_t)row[r]];
                }
        if (g_out) g_out[c] = col[(size_t)P.n_pad];
    }
    return FEBA_OK;
}

}  // extern "C"
