// Plan of the reduced camera system (host side, plain C++17): row order, supertiles, structural pattern
// with fill, elimination tree and -- for a group of GPUs -- which rank owns which subtree.
//
// The reference inverts the dense (bordered) normal matrix (main.m:432,442).  The point-eliminated camera
// system S couples two images only when they share a tie point, so for a photogrammetric block it is as
// sparse as the image adjacency graph.  Here that graph is cut by NESTED DISSECTION: a part of the block is
// split into two halves and the images of one half that see points of the other (the separator); halves are
// numbered first, separators after them, recursively.  Unknowns of different halves are never coupled, so
//   * their factorisations are independent chains (the dense form is ONE chain of u_c/64 sequential
//     64x64 diagonal factorisations; here the longest root-to-leaf path decides, 45 instead of 188 blocks on
//     BASELINE configs[3]);
//   * a group of GPUs can own one subtree each: all points a rank owns touch its subtree and the shared top
//     separators only, the subtree is factorised locally and only the top part of S crosses NVLink.
// Every tree node is padded to whole 64-row blocks (unit diagonal on the padding rows), cut into supertiles of
// at most tile_max blocks, and the supertile pattern (adjacency + symbolic fill) tells the task graph
// (feba_chol.cu::chol_tiles) which TRSM / UPDATE tasks exist.
//
// Row layout of a masked plan:   [ leaves ... | separators, deepest first ... | root separator, datum images,
//                                  camera unknowns ]   then the augmented block row.
// The identity plan (small or dense systems) keeps the order of Buildxhat.m:22-106 with no interior padding.
//
// Plain C++ so that tests/host_model compiles it with g++ (tests/test_reduced_plan_host.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <numeric>
#include <thread>
#include <vector>

namespace feba {

struct PlanOptions {
    int mode = 0;            // 0 automatic, 1 force a masked plan, -1 force the identity plan
    int max_depth = 12;      // deepest level of the dissection
    int leaf_images = 96;    // parts of at most this many images are not cut further
    int tile_max = 6;        // supertile cap in 64-row blocks (config 4, factorisation: 6 / 8 / 12 / 16 -> 2.8 / 3.1 / 3.1 / 3.7 ms)
    double sep_frac = 0.30;  // a cut whose separator exceeds this fraction of the part is rejected
    int identity_tile = 0;   // identity plan: uniform supertiles of this many blocks (0: one tile)
    int min_blocks = 48;     // automatic mode: smaller reduced systems keep the identity plan
};

struct PlanNode {
    int parent = -1, depth = 0;
    int child[2] = {-1, -1};
    std::vector<int> imgs;          // images of this node in row order
    int tile0 = 0, n_tiles = 0;     // its supertiles
    int row0 = 0, rows = 0;         // first row, rows incl. padding
    int rank_lo = 0, rank_hi = 1;   // ranks whose subtrees lie below this node [lo, hi)
    int owner = -1;                 // rank that factorises this node, -1: shared top node (every rank)
    int lane = 0;                   // position inside its depth level (stream choice of the task graph)
};

struct ReducedPlan {
    bool masked = false;
    int n_img = 0, ui = 0, cam_rows = 0, n_red = 0;
    std::vector<int> img_row;       // first row of every image
    int off_cam = 0;                // first row of the camera unknowns
    int n_pad = 0;                  // rows of the factorised part (multiple of 64)
    std::vector<int> row_ext;       // row -> index in the EOP/IOP part of xhat (Buildxhat order), -1 padding
    int NT = 0;                     // supertiles of the factorised part; tile NT is the augmented block row
    std::vector<int> tile_b0;       // NT + 1 entries: first 64-block of tile t; tile_b0[NT] = n_pad / 64
    std::vector<unsigned char> nz;  // (NT+1) x (NT+1) row-major, lower triangle: structural pattern incl. fill
    std::vector<int> tile_node;
    std::vector<PlanNode> nodes;    // nodes[0] is the root
    std::vector<int> img_node;      // node of every image
    std::vector<int> datum;         // datum images of the sparse-datum form (masked plan of a free network)
    int world = 1;
    int top_tile0 = 0;              // tiles >= top_tile0 belong to shared top nodes (world > 1), else NT
    double flop = 0, flop_dense = 0;
    int chain_blocks = 0;           // 64-blocks on the longest dependent chain of diagonal factorisations
    int nb() const { return n_pad / 64; }
    int NR() const { return NT + 1; }
    bool at(int i, int j) const { return nz[(size_t)i * NR() + j] != 0; }
    void set(int i, int j) {
        if (i < j) std::swap(i, j);
        nz[(size_t)i * NR() + j] = 1;
    }
    int tile_blocks(int t) const { return t == NT ? 1 : tile_b0[(size_t)t + 1] - tile_b0[(size_t)t]; }
    int tile_of_row(int row) const {
        const int b = row / 64;
        int lo = 0, hi = NT - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) / 2;
            if (tile_b0[(size_t)mid] <= b) lo = mid;
            else hi = mid - 1;
        }
        return lo;
    }
    int tile_owner(int t) const { return t >= NT ? -1 : nodes[(size_t)tile_node[(size_t)t]].owner; }
    // envelope: first 64-block column that can be non-zero in each 64-block row
    std::vector<int> row_first_block() const {
        std::vector<int> v((size_t)nb(), 0);
        for (int t = 0; t < NT; ++t) {
            int J = 0;
            while (J < t && !at(t, J)) ++J;
            for (int b = tile_b0[(size_t)t]; b < tile_b0[(size_t)t + 1]; ++b) v[(size_t)b] = tile_b0[(size_t)J];
        }
        return v;
    }
};

// ---- image adjacency: images a != b are neighbours when an object point is seen by both.  Only tie points
// couple images in S; with couple_control the control points do as well (group runs: all observations of a
// point stay on one rank, so its images must lie on one root-to-leaf path).  Observations sorted by point:
// segment s = [seg_start[s], seg_start[s+1]), image of observation o = simg[o].
inline void image_adjacency(int n_img, int n_seg, const int* seg_start, const int* simg, const unsigned char* seg_tie,
                            bool couple_control, std::vector<int>& ptr, std::vector<int>& idx) {
    std::vector<int> cnt((size_t)n_img + 1, 0);
    auto use = [&](int s) { return couple_control || seg_tie[s]; };
    for (int s = 0; s < n_seg; ++s)
        if (use(s))
            for (int o = seg_start[s]; o < seg_start[s + 1]; ++o) ++cnt[(size_t)simg[o] + 1];
    for (int i = 0; i < n_img; ++i) cnt[(size_t)i + 1] += cnt[(size_t)i];
    std::vector<int> segs((size_t)cnt[(size_t)n_img]);
    {
        std::vector<int> cur(cnt.begin(), cnt.end() - 1);
        for (int s = 0; s < n_seg; ++s)
            if (use(s))
                for (int o = seg_start[s]; o < seg_start[s + 1]; ++o) segs[(size_t)cur[(size_t)simg[o]]++] = s;
    }
    std::vector<std::vector<int>> nbr((size_t)n_img);
    unsigned hw = std::thread::hardware_concurrency();
    int nth = (int)(hw ? hw : 1);
    if (nth > 16) nth = 16;
    if ((long long)segs.size() < 200000) nth = 1;
    auto work = [&](int t) {
        std::vector<int> mark((size_t)n_img, -1);
        for (int a = t; a < n_img; a += nth) {
            std::vector<int>& out = nbr[(size_t)a];
            mark[(size_t)a] = a;
            for (int q = cnt[(size_t)a]; q < cnt[(size_t)a + 1]; ++q) {
                const int s = segs[(size_t)q];
                for (int o = seg_start[s]; o < seg_start[s + 1]; ++o) {
                    const int b = simg[o];
                    if (mark[(size_t)b] != a) {
                        mark[(size_t)b] = a;
                        out.push_back(b);
                    }
                }
            }
            std::sort(out.begin(), out.end());
        }
    };
    if (nth == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nth; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    ptr.assign((size_t)n_img + 1, 0);
    for (int a = 0; a < n_img; ++a) ptr[(size_t)a + 1] = ptr[(size_t)a] + (int)nbr[(size_t)a].size();
    idx.resize((size_t)ptr[(size_t)n_img]);
    for (int a = 0; a < n_img; ++a) std::copy(nbr[(size_t)a].begin(), nbr[(size_t)a].end(), idx.begin() + ptr[(size_t)a]);
}

namespace plan_detail {

struct Graph {
    int n;
    const int* ptr;
    const int* idx;
    const double* pos;   // n x 3 (camera stations), may be null
};

inline int longest_axis(const Graph& g, const std::vector<int>& part) {
    if (!g.pos) return -1;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int v : part)
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], g.pos[3 * (size_t)v + k]);
            hi[k] = std::max(hi[k], g.pos[3 * (size_t)v + k]);
        }
    int ax = 0;
    for (int k = 1; k < 3; ++k)
        if (hi[k] - lo[k] > hi[ax] - lo[ax]) ax = k;
    return ax;
}

// images of a part sorted along its longest axis (index order when there are no positions)
inline void sort_along(const Graph& g, std::vector<int>& part, int ax) {
    if (ax < 0) {
        std::sort(part.begin(), part.end());
        return;
    }
    std::sort(part.begin(), part.end(), [&](int a, int b) {
        const double pa = g.pos[3 * (size_t)a + ax], pb = g.pos[3 * (size_t)b + ax];
        return pa < pb || (pa == pb && a < b);
    });
}

struct Cut {
    std::vector<int> sep, A, B;
    bool ok = false;
    bool from_first = true;      // the separator was taken from the first group
};

// Cut `part` across axis `ax` after its first `na` images (sorted along ax): side[] = 1 for the first group, 2 for
// the second; the separator is the smaller of the two boundaries.  `side` is a scratch array of n zeros.
// force: -1 = the smaller boundary, 1 / 0 = the boundary of the first / second group.
inline Cut cut_at(const Graph& g, const std::vector<int>& sorted, int na, std::vector<unsigned char>& side, int force = -1) {
    Cut c;
    const int n = (int)sorted.size();
    for (int k = 0; k < n; ++k) side[(size_t)sorted[(size_t)k]] = k < na ? 1 : 2;
    std::vector<int> bA, bB;
    for (int k = 0; k < n; ++k) {
        const int v = sorted[(size_t)k];
        const unsigned char other = side[(size_t)v] == 1 ? 2 : 1;
        bool touch = false;
        for (int q = g.ptr[v]; q < g.ptr[v + 1] && !touch; ++q) touch = side[(size_t)g.idx[q]] == other;
        if (touch) (side[(size_t)v] == 1 ? bA : bB).push_back(v);
    }
    const bool fromA = force < 0 ? bA.size() <= bB.size() : force == 1;
    c.from_first = fromA;
    c.sep = fromA ? bA : bB;
    for (int v : c.sep) side[(size_t)v] = 3;
    for (int k = 0; k < n; ++k) {
        const int v = sorted[(size_t)k];
        if (side[(size_t)v] == 1) c.A.push_back(v);
        else if (side[(size_t)v] == 2) c.B.push_back(v);
    }
    for (int k = 0; k < n; ++k) side[(size_t)sorted[(size_t)k]] = 0;
    c.ok = !c.A.empty() && !c.B.empty();
    return c;
}

inline Cut best_cut(const Graph& g, const std::vector<int>& part, std::vector<unsigned char>& side) {
    Cut best;
    const int n = (int)part.size();
    // candidate axes: the two longest extents of the part (or index order without positions)
    int axes[2] = {-1, -1};
    if (g.pos) {
        double ext[3];
        for (int k = 0; k < 3; ++k) {
            double lo = 1e300, hi = -1e300;
            for (int v : part) {
                lo = std::min(lo, g.pos[3 * (size_t)v + k]);
                hi = std::max(hi, g.pos[3 * (size_t)v + k]);
            }
            ext[k] = hi - lo;
        }
        int o[3] = {0, 1, 2};
        std::sort(o, o + 3, [&](int a, int b) { return ext[a] > ext[b] || (ext[a] == ext[b] && a < b); });
        axes[0] = o[0];
        axes[1] = ext[o[1]] > 0.0 ? o[1] : -1;
    }
    // candidates: both axes x both sides of the cut; the cut position is moved so that the halves WITHOUT the
    // separator balance.  Score: separator size, penalised by the imbalance of the halves (a group of GPUs waits
    // for its largest subtree; on one GPU the longest chain does).
    double best_score = 1e300;
    for (int c = 0; c < 2; ++c) {
        if (c == 1 && axes[1] < 0) break;
        std::vector<int> sorted = part;
        sort_along(g, sorted, axes[c]);
        for (int from_first = 1; from_first >= 0; --from_first) {
            Cut first = cut_at(g, sorted, n / 2, side, from_first);
            if (!first.ok) continue;
            int na = n / 2 + (from_first ? (int)first.sep.size() / 2 : -(int)first.sep.size() / 2);
            if (na < 1) na = 1;
            if (na > n - 1) na = n - 1;
            Cut second = cut_at(g, sorted, na, side, from_first);
            for (Cut* cand : {&first, &second}) {
                if (!cand->ok) continue;
                const double imb = std::fabs((double)cand->A.size() - (double)cand->B.size()) / n;
                const double score = (double)cand->sep.size() * (1.0 + 3.0 * imb);
                if (score < best_score) {
                    best_score = score;
                    best = *cand;
                }
            }
        }
    }
    return best;
}

}  // namespace plan_detail

// Eight well-spread datum images (farthest-point sampling on the camera stations; index order without
// positions): the sparse-datum form factorises S + E E' with E = G~ restricted to these images, which live in
// the root node so that E E' adds no fill.
inline std::vector<int> plan_datum_images(int n_img, const double* pos, int want = 8) {
    std::vector<int> v;
    if (n_img <= want) {
        for (int i = 0; i < n_img; ++i) v.push_back(i);
        return v;
    }
    if (!pos) {
        for (int k = 0; k < want; ++k) v.push_back((int)((long long)(n_img - 1) * k / (want - 1)));
        return v;
    }
    double c[3] = {0, 0, 0};
    for (int i = 0; i < n_img; ++i)
        for (int k = 0; k < 3; ++k) c[k] += pos[3 * (size_t)i + k];
    for (int k = 0; k < 3; ++k) c[k] /= n_img;
    std::vector<double> d((size_t)n_img);
    int first = 0;
    double far = -1.0;
    for (int i = 0; i < n_img; ++i) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += (pos[3 * (size_t)i + k] - c[k]) * (pos[3 * (size_t)i + k] - c[k]);
        if (s > far) { far = s; first = i; }
        d[(size_t)i] = 1e300;
    }
    int cur = first;
    for (int q = 0; q < want; ++q) {
        v.push_back(cur);
        int nxt = -1;
        double best = -1.0;
        for (int i = 0; i < n_img; ++i) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += (pos[3 * (size_t)i + k] - pos[3 * (size_t)cur + k]) * (pos[3 * (size_t)i + k] - pos[3 * (size_t)cur + k]);
            if (s < d[(size_t)i]) d[(size_t)i] = s;
            if (d[(size_t)i] > best) { best = d[(size_t)i]; nxt = i; }
        }
        cur = nxt;
    }
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    return v;
}

// flop of the masked supertile factorisation (DIAG n^3/3, TRSM m n^2, UPDATE 2 m_i m_j n, m^2 n on the diagonal)
inline double plan_flop(const ReducedPlan& P, bool dense) {
    double fl = 0;
    const int NT = P.NT;
    auto size = [&](int t) { return 64.0 * P.tile_blocks(t); };
    for (int k = 0; k < NT; ++k) {
        const double nk = size(k);
        fl += nk * nk * nk / 3;
        for (int i = k + 1; i <= NT; ++i) {
            if (!dense && !P.at(i, k)) continue;
            fl += size(i) * nk * nk;
            for (int j = k + 1; j <= i; ++j) {
                if (!dense && !P.at(j, k)) continue;
                if (j == NT && i != NT) continue;
                fl += (i != j ? 2.0 : 1.0) * size(i) * size(j) * nk;
            }
        }
    }
    return fl;
}

namespace plan_detail {

inline void finish_tiles_and_pattern(ReducedPlan& P, const Graph& g, const std::vector<std::pair<int, int>>& node_blocks,
                                     int tile_max, bool dense_pattern) {
    // node_blocks: (node, blocks) in row order
    P.tile_b0.clear();
    P.tile_node.clear();
    int b = 0;
    for (const auto& nbk : node_blocks) {
        PlanNode& nd = P.nodes[(size_t)nbk.first];
        const int nblk = nbk.second;
        const int nt = tile_max > 0 ? (nblk + tile_max - 1) / tile_max : 1;
        nd.tile0 = (int)P.tile_b0.size();
        nd.n_tiles = nt;
        for (int t = 0; t < nt; ++t) {
            P.tile_b0.push_back(b);
            P.tile_node.push_back(nbk.first);
            b += nblk / nt + (t < nblk % nt ? 1 : 0);
        }
    }
    P.NT = (int)P.tile_b0.size();
    P.tile_b0.push_back(b);
    const int NT = P.NT, NR = NT + 1;
    P.nz.assign((size_t)NR * NR, dense_pattern ? 1 : 0);
    if (!dense_pattern) {
        for (int t = 0; t < NT; ++t) P.set(t, t);
        auto tiles_of_img = [&](int v, int& t0, int& t1) {
            t0 = P.tile_of_row(P.img_row[(size_t)v]);
            t1 = P.tile_of_row(P.img_row[(size_t)v] + P.ui - 1);
        };
        for (int v = 0; v < P.n_img && P.ui > 0; ++v) {
            int a0, a1;
            tiles_of_img(v, a0, a1);
            P.set(a1, a0);
            for (int q = g.ptr[v]; q < g.ptr[v + 1]; ++q) {
                int b0, b1;
                tiles_of_img(g.idx[q], b0, b1);
                for (int a = a0; a <= a1; ++a)
                    for (int c = b0; c <= b1; ++c) P.set(a, c);
            }
        }
        for (size_t x = 0; x < P.datum.size() && P.ui > 0; ++x)
            for (size_t y = 0; y <= x; ++y) {
                int a0, a1, b0, b1;
                tiles_of_img(P.datum[x], a0, a1);
                tiles_of_img(P.datum[y], b0, b1);
                for (int a = a0; a <= a1; ++a)
                    for (int c = b0; c <= b1; ++c) P.set(a, c);
            }
        if (P.cam_rows > 0)                                  // camera unknowns: coupled with every image
            for (int i = P.tile_of_row(P.off_cam); i <= P.tile_of_row(P.off_cam + P.cam_rows - 1); ++i)
                for (int j = 0; j <= i; ++j) P.set(i, j);
        for (int j = 0; j <= NT; ++j) P.set(NT, j);          // augmented block row
        for (int k = 0; k < NT; ++k)                         // symbolic factorisation
            for (int i = k + 1; i < NR; ++i) {
                if (!P.at(i, k)) continue;
                for (int j = k + 1; j <= i; ++j)
                    if (P.at(j, k)) P.set(i, j);
            }
    }
    P.flop = plan_flop(P, false);
    P.flop_dense = std::pow(64.0 * P.nb(), 3) / 3;
}

}  // namespace plan_detail

// The identity plan: rows in Buildxhat order, padding only at the end, dense pattern.
inline ReducedPlan identity_plan(int n_img, int ui, int cam_rows, int tile_blocks) {
    ReducedPlan P;
    P.masked = false;
    P.n_img = n_img;
    P.ui = ui;
    P.cam_rows = cam_rows;
    P.n_red = ui * n_img + cam_rows;
    P.img_row.resize((size_t)n_img);
    for (int i = 0; i < n_img; ++i) P.img_row[(size_t)i] = ui * i;
    P.off_cam = ui * n_img;
    P.n_pad = (P.n_red + 63) / 64 * 64;
    P.row_ext.assign((size_t)P.n_pad, -1);
    for (int r = 0; r < P.n_red; ++r) P.row_ext[(size_t)r] = r;
    P.nodes.resize(1);
    P.nodes[0].imgs.resize((size_t)n_img);
    std::iota(P.nodes[0].imgs.begin(), P.nodes[0].imgs.end(), 0);
    P.nodes[0].rows = P.n_pad;
    P.img_node.assign((size_t)n_img, 0);
    plan_detail::Graph g{n_img, nullptr, nullptr, nullptr};
    plan_detail::finish_tiles_and_pattern(P, g, {{0, P.n_pad / 64}}, tile_blocks, true);
    P.top_tile0 = P.NT;
    P.chain_blocks = P.nb();
    return P;
}

// adjacency in CSR form (image_adjacency), pos: n_img x 3 camera stations (or null), inner: free network
// (datum images wanted), world: number of ranks (power of two) that will share the factorisation.
// Returns a masked plan; the caller compares flop / chain_blocks with the identity plan (choose_plan below).
inline ReducedPlan masked_plan(int n_img, int ui, int cam_rows, const int* adj_ptr, const int* adj_idx, const double* pos,
                               bool inner, int world, const PlanOptions& opt) {
    using namespace plan_detail;
    ReducedPlan P;
    P.masked = true;
    P.n_img = n_img;
    P.ui = ui;
    P.cam_rows = cam_rows;
    P.n_red = ui * n_img + cam_rows;
    P.world = world;
    Graph g{n_img, adj_ptr, adj_idx, pos};
    int need_depth = 0;
    while ((1 << need_depth) < world) ++need_depth;
    if (inner) P.datum = plan_datum_images(n_img, pos);
    std::vector<unsigned char> is_datum((size_t)n_img, 0);
    for (int v : P.datum) is_datum[(size_t)v] = 1;

    // ---- recursive bisection (explicit stack; children are created in order, so ranks run left to right)
    std::vector<unsigned char> side((size_t)n_img, 0);
    struct Item { int node; std::vector<int> part; };
    std::vector<Item> stack;
    P.nodes.emplace_back();
    {
        std::vector<int> all;
        for (int v = 0; v < n_img; ++v)
            if (!is_datum[(size_t)v]) all.push_back(v);
        P.nodes[0].rank_lo = 0;
        P.nodes[0].rank_hi = world;
        stack.push_back({0, std::move(all)});
    }
    while (!stack.empty()) {
        Item it = std::move(stack.back());
        stack.pop_back();
        const int id = it.node;
        const int depth = P.nodes[(size_t)id].depth;
        const int n = (int)it.part.size();
        const bool must = depth < need_depth;                        // a group needs 2^depth subtrees
        bool split = (must || n > opt.leaf_images) && depth < opt.max_depth && n >= 2;
        Cut c;
        if (split) {
            c = best_cut(g, it.part, side);
            split = c.ok && (must || (double)c.sep.size() <= opt.sep_frac * n);
        }
        if (!split) {
            P.nodes[(size_t)id].imgs = std::move(it.part);
            continue;
        }
        P.nodes[(size_t)id].imgs = std::move(c.sep);
        const int lo = P.nodes[(size_t)id].rank_lo, hi = P.nodes[(size_t)id].rank_hi;
        const int mid = hi - lo > 1 ? (lo + hi) / 2 : hi;
        for (int s = 0; s < 2; ++s) {
            PlanNode ch;
            ch.parent = id;
            ch.depth = depth + 1;
            ch.rank_lo = hi - lo > 1 ? (s == 0 ? lo : mid) : lo;
            ch.rank_hi = hi - lo > 1 ? (s == 0 ? mid : hi) : hi;
            const int cid = (int)P.nodes.size();
            P.nodes.push_back(ch);
            P.nodes[(size_t)id].child[s] = cid;
        }
        // push B first so that A is expanded first (left-to-right numbering is by node id order anyway)
        stack.push_back({P.nodes[(size_t)id].child[1], std::move(c.B)});
        stack.push_back({P.nodes[(size_t)id].child[0], std::move(c.A)});
    }
    // owners: a node whose rank range is a single rank belongs to it when world > 1; wider ranges are shared
    bool group_ok = true;
    for (PlanNode& nd : P.nodes) {
        if (world == 1) nd.owner = -1;
        else if (nd.rank_hi - nd.rank_lo == 1) nd.owner = nd.rank_lo;
        else {
            nd.owner = -1;
            if (nd.child[0] < 0) group_ok = false;               // could not be cut deep enough for the group
        }
    }
    if (!group_ok) {
        P.world = -1;                                            // caller falls back to the replicated form
    }
    // ---- row order: deepest nodes first, the root last; inside a node along its longest axis
    std::vector<int> order((size_t)P.nodes.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const bool ta = P.nodes[(size_t)a].owner < 0 && world > 1, tb = P.nodes[(size_t)b].owner < 0 && world > 1;
        if (ta != tb) return tb;                                 // owned subtrees before the shared top
        return P.nodes[(size_t)a].depth > P.nodes[(size_t)b].depth;
    });
    {
        std::vector<int> per_depth(64, 0);
        for (size_t i = 0; i < P.nodes.size(); ++i) {
            const int d = std::min(P.nodes[i].depth, 63);
            P.nodes[i].lane = per_depth[(size_t)d]++;
        }
    }
    P.img_row.assign((size_t)n_img, -1);
    P.img_node.assign((size_t)n_img, 0);
    std::vector<std::pair<int, int>> node_blocks;
    int rows = 0;
    for (int id : order) {
        PlanNode& nd = P.nodes[(size_t)id];
        sort_along(g, nd.imgs, longest_axis(g, nd.imgs));
        if (id == 0)
            for (int v : P.datum) nd.imgs.push_back(v);          // datum images: tail of the root node
        nd.row0 = rows;
        int r = rows;
        for (int v : nd.imgs) {
            P.img_row[(size_t)v] = r;
            P.img_node[(size_t)v] = id;
            r += ui;
        }
        if (id == 0) {
            P.off_cam = r;
            r += cam_rows;
        }
        const int nblk = (r - rows + 63) / 64;
        if (nblk == 0) continue;                                 // empty node (no images): no rows
        nd.rows = nblk * 64;
        node_blocks.push_back({id, nblk});
        rows += nblk * 64;
    }
    P.n_pad = rows;
    P.row_ext.assign((size_t)rows, -1);
    for (int v = 0; v < n_img; ++v)
        for (int k = 0; k < ui; ++k) P.row_ext[(size_t)P.img_row[(size_t)v] + k] = ui * v + k;
    for (int k = 0; k < cam_rows; ++k) P.row_ext[(size_t)P.off_cam + k] = ui * n_img + k;
    finish_tiles_and_pattern(P, g, node_blocks, opt.tile_max, false);
    // shared top tiles are the tail of the order
    P.top_tile0 = P.NT;
    if (world > 1 && P.world > 0)
        for (int t = P.NT - 1; t >= 0 && P.nodes[(size_t)P.tile_node[(size_t)t]].owner < 0; --t) P.top_tile0 = t;
    // longest chain of dependent diagonal blocks: blocks of a node + the longest chain among its children
    {
        std::vector<int> chain(P.nodes.size(), 0);
        for (int i = (int)P.nodes.size() - 1; i >= 0; --i) {     // children have larger ids than parents
            const PlanNode& nd = P.nodes[(size_t)i];
            int c = 0;
            for (int s = 0; s < 2; ++s)
                if (nd.child[s] >= 0) c = std::max(c, chain[(size_t)nd.child[s]]);
            chain[(size_t)i] = c + nd.rows / 64;
        }
        P.chain_blocks = chain[0];
    }
    return P;
}

// Which rank owns every object point (segment) of a group run: the rank of the deepest node among the point's
// images; points that only touch shared top nodes are dealt out over the ranks below the deepest such node.
// Returns false when a point's images do not lie on one root-to-leaf path (adjacency built without it).
inline bool plan_point_owner(const ReducedPlan& P, int n_seg, const int* seg_start, const int* simg, int* owner_out) {
    bool ok = true;
    const int world = P.world > 0 ? P.world : 1;
    std::vector<long long> load((size_t)world, 0);
    std::vector<int> deepest_of((size_t)n_seg, -1);
    // pass 1: points of owned subtrees
    for (int s = 0; s < n_seg; ++s) {
        int deepest = -1;
        for (int o = seg_start[s]; o < seg_start[s + 1]; ++o) {
            const int nd = P.img_node[(size_t)simg[o]];
            if (deepest < 0 || P.nodes[(size_t)nd].depth > P.nodes[(size_t)deepest].depth) deepest = nd;
        }
        deepest_of[(size_t)s] = deepest;
        if (deepest < 0) continue;
        // every image of the point must sit in an ancestor-or-self of the deepest node
        for (int o = seg_start[s]; o < seg_start[s + 1] && ok; ++o) {
            int a = deepest;
            const int nd = P.img_node[(size_t)simg[o]];
            while (a >= 0 && a != nd) a = P.nodes[(size_t)a].parent;
            if (a < 0) ok = false;
        }
        const PlanNode& d = P.nodes[(size_t)deepest];
        if (d.owner >= 0) {
            owner_out[s] = d.owner;
            load[(size_t)d.owner] += seg_start[s + 1] - seg_start[s];
        }
    }
    // pass 2: points that touch shared top nodes only go to the least loaded rank below their deepest node
    // (in point order: deterministic), which also evens out subtrees of different size
    for (int s = 0; s < n_seg; ++s) {
        const int deepest = deepest_of[(size_t)s];
        int lo = 0, hi = world;
        if (deepest >= 0) {
            const PlanNode& d = P.nodes[(size_t)deepest];
            if (d.owner >= 0) continue;
            lo = d.rank_lo;
            hi = d.rank_hi;
        }
        int pick = lo;
        for (int r = lo + 1; r < hi; ++r)
            if (load[(size_t)r] < load[(size_t)pick]) pick = r;
        owner_out[s] = pick;
        load[(size_t)pick] += seg_start[s + 1] - seg_start[s];
    }
    return ok;
}

inline PlanOptions plan_options_from_env() {
    PlanOptions o;
    auto geti = [](const char* name, int dflt) {
        const char* e = std::getenv(name);
        return e ? std::atoi(e) : dflt;
    };
    o.mode = geti("FEBA_PLAN", 0);              // 1 force masked, -1 force identity (dense), 0 automatic
    if (const char* e = std::getenv("FEBA_SPARSE")) {   // round-1 switch kept: 0 = dense
        if (std::atoi(e) == 0) o.mode = -1;
    }
    o.max_depth = geti("FEBA_ND_DEPTH", o.max_depth);
    o.leaf_images = geti("FEBA_ND_LEAF", o.leaf_images);
    o.tile_max = geti("FEBA_TILE_MAX", o.tile_max);
    o.min_blocks = geti("FEBA_PLAN_MIN_BLOCKS", o.min_blocks);
    if (o.tile_max < 1) o.tile_max = 1;
    return o;
}

}  // namespace feba
