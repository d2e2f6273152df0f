// TEST INFRASTRUCTURE: the CUDA path's per-observation source (csrc/feba_model.cuh: table rows,
// projection, distortion, analytic Jacobian, misclosure) compiled for the HOST with g++, so that the very
// statements the kernels inline can be checked in a container without a GPU against the reference's own
// BuildAwG.m (tests/test_cuda_model_source_on_host.py).  Nothing in the product links or calls this.
#include "../../fish-eye_bundle_adjustment_b200/csrc/feba_model.cuh"

namespace {

template <int NK>
void run(int type, double x, double y, const double* it, const double* ct, const double* xyz, double* Je, double* Jc,
         double* Jt, double* w) {
    feba::ObsJac<NK> o;
    feba::observation<NK, true>(type, x, y, it, ct, xyz[0], xyz[1], xyz[2], o);
    for (int r = 0; r < 2; ++r) {
        for (int k = 0; k < 6; ++k) Je[6 * r + k] = o.Je[r][k];
        for (int k = 0; k < NK + 5; ++k) Jc[(NK + 5) * r + k] = o.Jc[r][k];
        for (int k = 0; k < 3; ++k) Jt[3 * r + k] = o.Jt[r][k];
        w[r] = o.w[r];
    }
}

}  // namespace

// eop[6] (angles in radians), iop[NK+5] = xp yp c k1..kNK p1 p2, box[5] = y_dir xmin ymin xmax ymax.
// Outputs row-major: Je[2][6], Jc[2][NK+5] (columns xp yp c k1..kNK p1 p2, distortion columns divided by
// r_max^(2j) / r_max^2 as BuildAwG.m:428-445), Jt[2][3], w[2].  Returns 0, or 1 for an unsupported NK.
extern "C" int feba_host_observation(int type, int NK, double x, double y, const double* eop, const double* iop,
                                     const double* box, const double* xyz, double* Je, double* Jc, double* Jt,
                                     double* w) {
    double it[feba::kImgStride], ct[feba::kCamStride];
    feba::image_table_row(eop, it);
    feba::camera_table_row(NK, iop, box, ct);
    switch (NK) {
        case 1: run<1>(type, x, y, it, ct, xyz, Je, Jc, Jt, w); break;
        case 2: run<2>(type, x, y, it, ct, xyz, Je, Jc, Jt, w); break;
        case 3: run<3>(type, x, y, it, ct, xyz, Je, Jc, Jt, w); break;
        case 4: run<4>(type, x, y, it, ct, xyz, Je, Jc, Jt, w); break;
        case 5: run<5>(type, x, y, it, ct, xyz, Je, Jc, Jt, w); break;
        case 6: run<6>(type, x, y, it, ct, xyz, Je, Jc, Jt, w); break;
        case 7: run<7>(type, x, y, it, ct, xyz, Je, Jc, Jt, w); break;
        case 8: run<8>(type, x, y, it, ct, xyz, Je, Jc, Jt, w); break;
        default: return 1;
    }
    return 0;
}

namespace {

template <int NK>
void run_residual(int type, bool has_cam, double x, double y, const double* it, const double* ct, const double* xyz,
                  const int* ecol, const int* ccol, const double* d_img, const double* d_cam, const double* d_pt,
                  double xp_new, double yp_new, double* v, double* rsd) {
    feba::ObsJac<NK> o;
    if (has_cam) {
        feba::observation<NK, true>(type, x, y, it, ct, xyz[0], xyz[1], xyz[2], o);
        feba::residual_of<NK, true>(o, ecol, ccol, d_img, d_cam, d_pt, v);
    } else {
        feba::observation<NK, false>(type, x, y, it, ct, xyz[0], xyz[1], xyz[2], o);
        feba::residual_of<NK, false>(o, ecol, ccol, d_img, d_cam, d_pt, v);
    }
    feba::rsd_row(x, y, xp_new, yp_new, v, rsd);
}

}  // namespace

// What k_residuals does for one observation: Jacobian blocks at the last linearisation point (eop, iop, xyz
// BEFORE the last update), v = w + J * (un-scaled last increment), then the BuildRSD row with the post-update
// xp, yp.  d_img / d_cam: the image's / camera's slice of the increment in xhat order; d_pt may be NULL.
extern "C" int feba_host_residual(int type, int NK, int has_cam, double x, double y, const double* eop,
                                  const double* iop, const double* box, const double* xyz, const int* ecol,
                                  const int* ccol, const double* d_img, const double* d_cam, const double* d_pt,
                                  double xp_new, double yp_new, double* v, double* rsd) {
    double it[feba::kImgStride], ct[feba::kCamStride];
    feba::image_table_row(eop, it);
    feba::camera_table_row(NK, iop, box, ct);
#define FEBA_CASE(N) case N: run_residual<N>(type, has_cam != 0, x, y, it, ct, xyz, ecol, ccol, d_img, d_cam, d_pt, xp_new, yp_new, v, rsd); break;
    switch (NK) {
        FEBA_CASE(1) FEBA_CASE(2) FEBA_CASE(3) FEBA_CASE(4) FEBA_CASE(5) FEBA_CASE(6) FEBA_CASE(7) FEBA_CASE(8)
        default: return 1;
    }
#undef FEBA_CASE
    return 0;
}

// Gblock of one image (k_G_rows): G[6][7] row-major from eop[6].
extern "C" void feba_host_inner_constraint_rows(const double* eop, double* G42) {
    double G[6][7];
    feba::inner_constraint_rows(eop, G);
    for (int q = 0; q < 6; ++q)
        for (int c = 0; c < 7; ++c) G42[7 * q + c] = G[q][c];
}

// ---- block-sparse reduced system (csrc/feba_sparse.h, opt-in FEBA_SPARSE=1): the 14x14 border of the
// sparse-datum form and the supertile pattern with symbolic fill, as the library computes them.
#include "../../fish-eye_bundle_adjustment_b200/csrc/feba_sparse.h"

// T225: 15x15 row-major (symmetric).  coef14 out.  Returns 1 when singular.
extern "C" int feba_host_sparse_border(const double* T225, double* coef14) {
    double T[feba::kSparseAugRows][feba::kSparseAugRows];
    for (int i = 0; i < feba::kSparseAugRows; ++i)
        for (int j = 0; j < feba::kSparseAugRows; ++j) T[i][j] = T225[feba::kSparseAugRows * i + j];
    return feba::sparse_border_solve(T, coef14) ? 0 : 1;
}

// datum_out: up to feba::kDatumImages (8) image indices, returns their number.
extern "C" int feba_host_sparse_datum(int n_img, int ui, int nb, int T, int* datum_out) {
    const std::vector<int> d = feba::sparse_datum_images(n_img, ui, nb, T);
    for (size_t i = 0; i < d.size(); ++i) datum_out[i] = d[i];
    return (int)d.size();
}

// nz_out: (NT+1) x (NT+1) row-major; returns NT.
extern "C" int feba_host_sparse_pattern(int nb, int T, int ui, int n_img, int off_cam, int n_red, int n_blocks,
                                        const int* blocks_ab, int n_datum, const int* datum, unsigned char* nz_out) {
    const std::vector<int> d(datum, datum + n_datum);
    const feba::SparsePattern P = feba::sparse_supertile_pattern(nb, T, ui, n_img, off_cam, n_red, n_blocks, blocks_ab, d);
    for (size_t i = 0; i < P.nz.size(); ++i) nz_out[i] = P.nz[i];
    return P.NT;
}

// Envelope of a pattern (backward substitution): first 64-block column that can be non-zero per 64-block row.
extern "C" void feba_host_sparse_row_first(int nb, int T, const unsigned char* nz, int* out_nb) {
    feba::SparsePattern P;
    P.T = T;
    P.NT = (nb + T - 1) / T;
    P.nz.assign(nz, nz + (size_t)(P.NT + 1) * (P.NT + 1));
    const std::vector<int> v = P.row_first_block(nb);
    for (int k = 0; k < nb; ++k) out_nb[k] = v[(size_t)k];
}

// ---- plan of the reduced system (csrc/feba_order.h): nested-dissection row order, supertiles, pattern, owners
#include "../../fish-eye_bundle_adjustment_b200/csrc/feba_order.h"

// adjacency from observations sorted by point (what feba_create does): returns nnz, fills ptr (n_img+1); idx via
// a second call with idx_out != null (capacity cap)
extern "C" int feba_host_adjacency(int n_img, int n_seg, const int* seg_start, const int* simg, const unsigned char* seg_tie,
                                   int couple_control, int* ptr_out, int* idx_out, int cap) {
    std::vector<int> ptr, idx;
    feba::image_adjacency(n_img, n_seg, seg_start, simg, seg_tie, couple_control != 0, ptr, idx);
    for (size_t i = 0; i < ptr.size(); ++i) ptr_out[i] = ptr[i];
    if (idx_out)
        for (size_t i = 0; i < idx.size() && (int)i < cap; ++i) idx_out[i] = idx[i];
    return (int)idx.size();
}

extern "C" void* feba_host_plan_masked(int n_img, int ui, int cam_rows, const int* adj_ptr, const int* adj_idx,
                                       const double* pos, int inner, int world, int max_depth, int leaf_images,
                                       int tile_max) {
    feba::PlanOptions o;
    if (max_depth >= 0) o.max_depth = max_depth;
    if (leaf_images > 0) o.leaf_images = leaf_images;
    if (tile_max > 0) o.tile_max = tile_max;
    return new feba::ReducedPlan(feba::masked_plan(n_img, ui, cam_rows, adj_ptr, adj_idx, pos, inner != 0, world, o));
}
extern "C" void* feba_host_plan_identity(int n_img, int ui, int cam_rows, int tile_blocks) {
    return new feba::ReducedPlan(feba::identity_plan(n_img, ui, cam_rows, tile_blocks));
}
extern "C" void feba_host_plan_free(void* p) { delete static_cast<feba::ReducedPlan*>(p); }
// info: n_pad, NT, off_cam, n_nodes, n_datum, world (-1: group not possible), top_tile0, chain_blocks
extern "C" void feba_host_plan_info(const void* p, int* info8, double* flop2) {
    const feba::ReducedPlan& P = *static_cast<const feba::ReducedPlan*>(p);
    info8[0] = P.n_pad; info8[1] = P.NT; info8[2] = P.off_cam; info8[3] = (int)P.nodes.size();
    info8[4] = (int)P.datum.size(); info8[5] = P.world; info8[6] = P.top_tile0; info8[7] = P.chain_blocks;
    flop2[0] = P.flop; flop2[1] = P.flop_dense;
}
// img_row[n_img], row_ext[n_pad], tile_b0[NT+1], nz[(NT+1)^2], tile_node[NT], datum[n_datum], img_node[n_img],
// node_info[n_nodes*6] = parent depth owner tile0 n_tiles lane, first_block[nb]
extern "C" void feba_host_plan_arrays(const void* p, int* img_row, int* row_ext, int* tile_b0, unsigned char* nz,
                                      int* tile_node, int* datum, int* img_node, int* node_info, int* first_block) {
    const feba::ReducedPlan& P = *static_cast<const feba::ReducedPlan*>(p);
    std::copy(P.img_row.begin(), P.img_row.end(), img_row);
    std::copy(P.row_ext.begin(), P.row_ext.end(), row_ext);
    std::copy(P.tile_b0.begin(), P.tile_b0.end(), tile_b0);
    std::copy(P.nz.begin(), P.nz.end(), nz);
    std::copy(P.tile_node.begin(), P.tile_node.end(), tile_node);
    std::copy(P.datum.begin(), P.datum.end(), datum);
    std::copy(P.img_node.begin(), P.img_node.end(), img_node);
    for (size_t i = 0; i < P.nodes.size(); ++i) {
        const feba::PlanNode& n = P.nodes[i];
        const int v[6] = {n.parent, n.depth, n.owner, n.tile0, n.n_tiles, n.lane};
        std::copy(v, v + 6, node_info + 6 * i);
    }
    const std::vector<int> fb = P.row_first_block();
    std::copy(fb.begin(), fb.end(), first_block);
}
extern "C" int feba_host_plan_point_owner(const void* p, int n_seg, const int* seg_start, const int* simg, int* owner) {
    return feba::plan_point_owner(*static_cast<const feba::ReducedPlan*>(p), n_seg, seg_start, simg, owner) ? 0 : 1;
}

// ---- chunk schedule of the assembly (csrc/feba_chunks.h)
#include "../../fish-eye_bundle_adjustment_b200/csrc/feba_chunks.h"

extern "C" void* feba_host_chunks(int n_img, int n_seg, const int* seg_start, const int* simg, const unsigned char* seg_tie,
                                  const int* img_row) {
    return new feba::ChunkSchedule(feba::build_chunks(n_img, n_seg, seg_start, simg, seg_tie, img_row));
}
extern "C" void feba_host_chunks_free(void* p) { delete static_cast<feba::ChunkSchedule*>(p); }
// sizes: n_chunks, image slots, block slots, pairs, distinct image pairs, ok
extern "C" void feba_host_chunks_sizes(const void* p, long long* out6) {
    const feba::ChunkSchedule& C = *static_cast<const feba::ChunkSchedule*>(p);
    out6[0] = C.n_chunks; out6[1] = (long long)C.slot_img.size(); out6[2] = (long long)C.bslot_a.size();
    out6[3] = C.n_pairs; out6[4] = (long long)C.tblk_a.size(); out6[5] = C.ok ? 1 : 0;
}
extern "C" void feba_host_chunks_arrays(const void* p, int* obs0, int* img0, int* slot_img, int* slot_obs0, int* slot_obs,
                                        int* blk0, int* bslot_a, int* bslot_b, int* bslot_pair0, unsigned int* pairs,
                                        int* timg_ptr, int* timg_slots, int* tblk_a, int* tblk_b, int* tblk_ptr,
                                        int* tblk_slots) {
    const feba::ChunkSchedule& C = *static_cast<const feba::ChunkSchedule*>(p);
    std::copy(C.obs0.begin(), C.obs0.end(), obs0);
    std::copy(C.img0.begin(), C.img0.end(), img0);
    std::copy(C.slot_img.begin(), C.slot_img.end(), slot_img);
    std::copy(C.slot_obs0.begin(), C.slot_obs0.end(), slot_obs0);
    for (size_t i = 0; i < C.slot_obs.size(); ++i) slot_obs[i] = C.slot_obs[i];
    std::copy(C.blk0.begin(), C.blk0.end(), blk0);
    std::copy(C.bslot_a.begin(), C.bslot_a.end(), bslot_a);
    std::copy(C.bslot_b.begin(), C.bslot_b.end(), bslot_b);
    std::copy(C.bslot_pair0.begin(), C.bslot_pair0.end(), bslot_pair0);
    std::copy(C.pairs.begin(), C.pairs.end(), pairs);
    std::copy(C.timg_ptr.begin(), C.timg_ptr.end(), timg_ptr);
    std::copy(C.timg_slots.begin(), C.timg_slots.end(), timg_slots);
    std::copy(C.tblk_a.begin(), C.tblk_a.end(), tblk_a);
    std::copy(C.tblk_b.begin(), C.tblk_b.end(), tblk_b);
    std::copy(C.tblk_ptr.begin(), C.tblk_ptr.end(), tblk_ptr);
    std::copy(C.tblk_slots.begin(), C.tblk_slots.end(), tblk_slots);
}
