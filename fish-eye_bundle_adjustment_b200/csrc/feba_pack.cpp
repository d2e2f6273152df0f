// Host-side problem build (include/feba_pack.h): the reference's text files ->
// the numeric structure-of-arrays of feba_problem.
//
// What it stands for in the reference: ReadFiles.m:49 (readmatrix as a string table), main.m:196-258
// (str2double, degrees -> radians, missing distortion terms -> 0), main.m:260-264 (Estimate_AllGCP)
// and main.m:277-384 (one linear strcmp scan over EXT, INT, CNT and TIE per observation).  Here the
// small tables are tokenised once, their ID columns go into open-addressing hash maps that keep the
// FIRST row of every ID (the scans `break` on the first hit), and the .pho file -- the only large
// one, ~40 B per observation -- is cut at line boundaries into one piece per host thread; each
// thread tokenises its piece, converts x and y, and resolves the two IDs in the same sweep.
// No GPU code: plain C++17, compiled into libfeba.so next to the CUDA translation units.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

#include "../../include/feba_pack.h"

namespace {

using sv = std::string_view;

thread_local std::string g_pack_error;

int pack_fail(int code, const char* fmt, ...) {
    char buf[600];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_pack_error = buf;
    return code;
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- a whole file in memory (read-only mapping; empty files are legal)
struct FileBuf {
    const char* data = nullptr;
    size_t size = 0;
    bool mapped = false;
    FileBuf() = default;
    FileBuf(const FileBuf&) = delete;
    FileBuf& operator=(const FileBuf&) = delete;
    ~FileBuf() {
        if (mapped && data) munmap(const_cast<char*>(data), size);
    }
    bool open(const char* path) {
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
            ::close(fd);
            return false;
        }
        size = (size_t)st.st_size;
        if (size) {
            void* p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p == MAP_FAILED) {
                ::close(fd);
                return false;
            }
            data = static_cast<const char*>(p);
            mapped = true;
        }
        ::close(fd);
        return true;
    }
};

inline bool is_eol(char c) { return c == '\n' || c == '\r'; }
inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\v' || c == '\f'; }

// One line of a table: tokens between runs of blanks, up to a '#' (ReadFiles.m:49).  Returns the
// position after the line terminator; tokens land in tok[0..*nt) (at most cap are kept, the count
// keeps running).
inline const char* next_line(const char* p, const char* end, sv* tok, int cap, int* nt) {
    int n = 0;
    bool comment = false;
    while (p < end && !is_eol(*p)) {
        if (comment) {
            ++p;
            continue;
        }
        if (*p == '#') {
            comment = true;
            ++p;
            continue;
        }
        if (is_blank(*p)) {
            ++p;
            continue;
        }
        const char* b = p;
        while (p < end && !is_eol(*p) && !is_blank(*p) && *p != '#') ++p;
        if (n < cap) tok[n] = sv(b, (size_t)(p - b));
        ++n;
    }
    if (p < end) ++p;   // the terminator ("\r\n" leaves an empty line behind, which callers skip)
    *nt = n;
    return p;
}

// str2double (main.m:199-256): NaN unless the whole token is a number.
inline double to_double(sv t) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    if (t.empty()) return nan;
    const char* b = t.data();
    const char* e = b + t.size();
    if (*b == '+') {
        ++b;
        if (b == e || *b == '+' || *b == '-') return nan;
    }
    for (const char* q = b; q < e; ++q)
        if (*q == '(' || *q == '_') return nan;            // "nan(...)" is from_chars-only syntax
    double v = 0.0;
    const auto r = std::from_chars(b, e, v, std::chars_format::general);
    if (r.ptr != e) return nan;
    if (r.ec == std::errc::result_out_of_range) {
        // overflow -> +-inf, underflow -> +-0 / subnormal, as strtod and str2double do
        const std::string z(b, e);
        return std::strtod(z.c_str(), nullptr);
    }
    if (r.ec != std::errc()) return nan;
    return v;
}

// ---- ID -> first row, open addressing
struct IdMap {
    std::vector<sv> key;
    std::vector<int32_t> val;
    size_t mask = 0;
    static uint64_t hash(sv s) {
        uint64_t h = 1469598103934665603ull;
        for (unsigned char c : s) {
            h ^= c;
            h *= 1099511628211ull;
        }
        return h ^ (h >> 29);
    }
    void build(const std::vector<sv>& ids) {
        size_t cap = 16;
        while (cap < 2 * ids.size() + 2) cap <<= 1;
        key.assign(cap, sv());
        val.assign(cap, -1);
        mask = cap - 1;
        for (size_t k = 0; k < ids.size(); ++k) {
            size_t s = hash(ids[k]) & mask;
            while (val[s] >= 0 && key[s] != ids[k]) s = (s + 1) & mask;
            if (val[s] < 0) {                               // later duplicates never win
                key[s] = ids[k];
                val[s] = (int32_t)k;
            }
        }
    }
    int32_t find(sv id) const {
        if (val.empty()) return -1;
        size_t s = hash(id) & mask;
        while (val[s] >= 0) {
            if (key[s] == id) return val[s];
            s = (s + 1) & mask;
        }
        return -1;
    }
};

// A small table: rows of up to `width` tokens (missing ones are empty views with data()==nullptr).
struct SmallTable {
    FileBuf file;
    std::vector<sv> tok;
    int width = 0;
    size_t rows = 0;
    sv at(size_t r, int c) const { return tok[r * (size_t)width + (size_t)c]; }
    bool has(size_t r, int c) const { return at(r, c).data() != nullptr; }
    bool read(const char* path, int w) {
        width = w;
        if (!file.open(path)) return false;
        const char* p = file.data;
        const char* end = p + file.size;
        std::vector<sv> line((size_t)w);
        while (p < end) {
            int nt = 0;
            std::fill(line.begin(), line.end(), sv());
            p = next_line(p, end, line.data(), w, &nt);
            if (nt == 0) continue;
            tok.insert(tok.end(), line.begin(), line.end());
            ++rows;
        }
        return true;
    }
};

}  // namespace

struct feba_pack {
    FileBuf pho;
    SmallTable ext, cnt, intr, tie;
    int32_t nk = 1, n_img = 0, n_cam = 0, n_pts = 0, n_tie = 0;
    std::vector<double> obs_x, obs_y, eop0, iop0, cam_box, xyz0;
    std::vector<int32_t> obs_img, obs_pt, img_cam, pt_tie, tie_pt;
    std::vector<sv> point_ids, image_ids, camera_ids, tie_ids;
    double sec[3] = {0, 0, 0};
};

namespace {

struct Piece {
    const char* b = nullptr;
    const char* e = nullptr;
    std::vector<double> x, y;
    std::vector<int32_t> img, pt;
    // first failed look-up of the piece: kind 1 image, 2 target; row = local row
    int bad_kind = 0;
    size_t bad_row = 0;
    std::string bad_id;
};

void resolve_piece(Piece* pc, const IdMap* ext_map, const IdMap* cnt_map) {
    const size_t guess = (size_t)(pc->e - pc->b) / 24 + 16;
    pc->x.reserve(guess);
    pc->y.reserve(guess);
    pc->img.reserve(guess);
    pc->pt.reserve(guess);
    const char* p = pc->b;
    sv t[4];
    while (p < pc->e) {
        int nt = 0;
        t[0] = t[1] = t[2] = t[3] = sv();
        p = next_line(p, pc->e, t, 4, &nt);
        if (nt == 0) continue;
        const int32_t j = ext_map->find(t[1]);               // main.m:286-291
        const int32_t c = cnt_map->find(t[0]);               // main.m:346-351
        if (!pc->bad_kind && (j < 0 || c < 0)) {
            // main.m looks the image up first (:293), then the target (:352)
            pc->bad_kind = j < 0 ? 1 : 2;
            pc->bad_row = pc->x.size();
            pc->bad_id = std::string(j < 0 ? t[1] : t[0]);
        }
        pc->img.push_back(j);
        pc->pt.push_back(c);
        pc->x.push_back(to_double(t[2]));                    // main.m:202-203
        pc->y.push_back(to_double(t[3]));
    }
}

int build_pack(feba_pack* P, const char* pho, const char* ext, const char* cnt, const char* intr, const char* tie,
               int all_gcp, int threads) {
    const double t0 = now_s();
    const int nk = P->nk;
    if (!P->ext.read(ext, 8)) return pack_fail(FEBA_PACK_ERR_IO, "Error reading files (%s)", ext);
    if (!P->cnt.read(cnt, 4)) return pack_fail(FEBA_PACK_ERR_IO, "Error reading files (%s)", cnt);
    if (!P->intr.read(intr, std::max(6, 5 + nk))) return pack_fail(FEBA_PACK_ERR_IO, "Error reading files (%s)", intr);
    const bool read_tie = tie && !all_gcp;
    if (read_tie && !P->tie.read(tie, 1)) return pack_fail(FEBA_PACK_ERR_IO, "Error reading files (%s)", tie);
    if (!P->pho.open(pho)) return pack_fail(FEBA_PACK_ERR_IO, "Error reading files (%s)", pho);
    if (P->intr.rows % 2) return pack_fail(FEBA_PACK_ERR_FORMAT, ".int needs two rows per camera (main.m:231-256)");
    if (P->ext.rows > 0x7fffffff || P->cnt.rows > 0x7fffffff) return pack_fail(FEBA_PACK_ERR_FORMAT, "table too long");

    // ---- EXT: imageID cameraID Xc Yc Zc w p k, angles in degrees (main.m:206-219)
    const size_t n_ext = P->ext.rows, n_cnt = P->cnt.rows, n_int = P->intr.rows / 2;
    std::vector<sv> ext_img(n_ext), ext_cam(n_ext), cnt_id(n_cnt), int_id(n_int);
    std::vector<double> eop(n_ext * 6);
    for (size_t r = 0; r < n_ext; ++r) {
        ext_img[r] = P->ext.at(r, 0);
        ext_cam[r] = P->ext.at(r, 1);
        for (int q = 0; q < 3; ++q) eop[r * 6 + q] = to_double(P->ext.at(r, 2 + q));
        for (int q = 3; q < 6; ++q) eop[r * 6 + q] = to_double(P->ext.at(r, 2 + q)) * M_PI / 180.0;
    }
    // ---- CNT: ID X Y Z (main.m:221-228)
    P->xyz0.resize(n_cnt * 3);
    for (size_t r = 0; r < n_cnt; ++r) {
        cnt_id[r] = P->cnt.at(r, 0);
        for (int q = 0; q < 3; ++q) P->xyz0[r * 3 + q] = to_double(P->cnt.at(r, 1 + q));
    }
    // ---- INT: row 1 cameraID y_dir xmin ymin xmax ymax, row 2 xp yp c k1..kNK p1 p2 (main.m:230-256)
    const int ncol = 3 + nk + 2;
    std::vector<double> box(n_int * 5), iop(n_int * (size_t)ncol);
    for (size_t c = 0; c < n_int; ++c) {
        const size_t r1 = 2 * c, r2 = 2 * c + 1;
        int_id[c] = P->intr.at(r1, 0);
        for (int q = 0; q < 5; ++q) box[c * 5 + q] = to_double(P->intr.at(r1, 1 + q));
        for (int q = 0; q < 3; ++q) iop[c * ncol + q] = to_double(P->intr.at(r2, q));
        for (int q = 3; q < ncol; ++q)                       // absent distortion terms are 0 (main.m:243-253)
            iop[c * ncol + q] = P->intr.has(r2, q) ? to_double(P->intr.at(r2, q)) : 0.0;
    }
    IdMap ext_map, cnt_map, cam_map;
    ext_map.build(ext_img);
    cnt_map.build(cnt_id);
    cam_map.build(int_id);
    const double t1 = now_s();

    // ---- PHO: pointID imageID x y, one piece per thread
    int nthr = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nthr = std::max(1, std::min(nthr, 32));
    if (P->pho.size < (size_t)(1 << 20)) nthr = 1;
    std::vector<Piece> pieces((size_t)nthr);
    {
        const char* b = P->pho.data;
        const char* end = b + P->pho.size;
        const char* cur = b;
        for (int k = 0; k < nthr; ++k) {
            const char* stop = k + 1 == nthr ? end : b + P->pho.size * (size_t)(k + 1) / (size_t)nthr;
            if (stop < cur) stop = cur;
            while (stop < end && !is_eol(*stop)) ++stop;      // cut at a line boundary
            pieces[k].b = cur;
            pieces[k].e = stop;
            cur = stop;
        }
        std::vector<std::thread> pool;
        for (int k = 1; k < nthr; ++k) pool.emplace_back(resolve_piece, &pieces[k], &ext_map, &cnt_map);
        resolve_piece(&pieces[0], &ext_map, &cnt_map);
        for (auto& th : pool) th.join();
    }
    for (const Piece& pc : pieces)                           // first failure in file order
        if (pc.bad_kind) {
            return pc.bad_kind == 1
                       ? pack_fail(FEBA_PACK_ERR_LOOKUP,
                                   "Could not find image %s from .pho in .ext. Check that the image ID exists in both files",
                                   pc.bad_id.c_str())
                       : pack_fail(FEBA_PACK_ERR_LOOKUP,
                                   "Could not find target %s from .pho in .cnt. Check that the target ID exists in both files",
                                   pc.bad_id.c_str());
        }
    size_t n_obs = 0;
    for (const Piece& pc : pieces) n_obs += pc.x.size();
    P->obs_x.resize(n_obs);
    P->obs_y.resize(n_obs);
    P->obs_img.resize(n_obs);
    P->obs_pt.resize(n_obs);
    {
        std::vector<size_t> off(pieces.size() + 1, 0);
        for (size_t k = 0; k < pieces.size(); ++k) off[k + 1] = off[k] + pieces[k].x.size();
        auto gather = [&](size_t k) {
            Piece& pc = pieces[k];
            if (pc.x.empty()) return;
            std::memcpy(&P->obs_x[off[k]], pc.x.data(), pc.x.size() * sizeof(double));
            std::memcpy(&P->obs_y[off[k]], pc.y.data(), pc.y.size() * sizeof(double));
            std::memcpy(&P->obs_img[off[k]], pc.img.data(), pc.img.size() * sizeof(int32_t));
            std::memcpy(&P->obs_pt[off[k]], pc.pt.data(), pc.pt.size() * sizeof(int32_t));
            Piece().x.swap(pc.x);
            Piece().y.swap(pc.y);
            Piece().img.swap(pc.img);
            Piece().pt.swap(pc.pt);
        };
        std::vector<std::thread> pool;
        for (size_t k = 1; k < pieces.size(); ++k) pool.emplace_back(gather, k);
        gather(0);
        for (auto& th : pool) th.join();
    }
    const double t2 = now_s();

    // ---- every camera named by .ext must exist in .int (main.m:310-320)
    std::vector<int32_t> img_cam_all(n_ext);
    for (size_t r = 0; r < n_ext; ++r) {
        img_cam_all[r] = cam_map.find(ext_cam[r]);
        if (img_cam_all[r] < 0)
            return pack_fail(FEBA_PACK_ERR_LOOKUP,
                             "Could not find camera %.*s from .ext in .int. Check that the camera ID exists in both files",
                             (int)ext_cam[r].size(), ext_cam[r].data());
    }
    // ---- data.numImg / data.numCam = distinct image / camera IDs over the observations (main.m:379-380)
    std::vector<char> img_used(n_ext, 0), pt_used(n_cnt, 0), cam_used(n_int, 0);
    int32_t max_img = -1;
    for (size_t i = 0; i < n_obs; ++i) {
        img_used[(size_t)P->obs_img[i]] = 1;
        pt_used[(size_t)P->obs_pt[i]] = 1;
        max_img = std::max(max_img, P->obs_img[i]);
    }
    int32_t num_img = 0, num_cam = 0;
    for (size_t r = 0; r < n_ext; ++r)
        if (img_used[r]) {
            ++num_img;
            cam_used[(size_t)img_cam_all[r]] = 1;
        }
    for (size_t c = 0; c < n_int; ++c) num_cam += cam_used[c];
    // Buildxhat.m:22-30 takes EXT rows 1..numImg and INT cameras 1..numCam as the parameter slots
    int32_t max_cam = -1;
    for (int32_t r = 0; r < num_img && (size_t)r < n_ext; ++r) max_cam = std::max(max_cam, img_cam_all[(size_t)r]);
    if (n_obs == 0 || max_img >= num_img || max_cam >= num_cam)
        return pack_fail(FEBA_PACK_ERR_LAYOUT, "EXT/INT must list exactly the images/cameras used in PHO, first");
    P->n_img = num_img;
    P->n_cam = num_cam;
    P->n_pts = (int32_t)n_cnt;
    P->img_cam.assign(img_cam_all.begin(), img_cam_all.begin() + num_img);
    P->eop0.assign(eop.begin(), eop.begin() + (size_t)num_img * 6);
    P->iop0.assign(iop.begin(), iop.begin() + (size_t)num_cam * ncol);
    P->cam_box.assign(box.begin(), box.begin() + (size_t)num_cam * 5);
    P->point_ids = cnt_id;
    P->image_ids.assign(ext_img.begin(), ext_img.begin() + num_img);
    P->camera_ids.assign(int_id.begin(), int_id.begin() + num_cam);

    // ---- TIE (main.m:180-188, :260-264, :362-375)
    if (all_gcp) {
        // unique(PHO(:,1)): the distinct target IDs, sorted
        for (size_t r = 0; r < n_cnt; ++r)
            if (pt_used[r]) P->tie_ids.push_back(cnt_id[r]);
        std::sort(P->tie_ids.begin(), P->tie_ids.end());
    } else if (read_tie) {
        P->tie_ids.resize(P->tie.rows);
        for (size_t r = 0; r < P->tie.rows; ++r) P->tie_ids[r] = P->tie.at(r, 0);
    }
    if (P->tie_ids.size() > 0x7fffffff) return pack_fail(FEBA_PACK_ERR_FORMAT, "table too long");
    P->n_tie = (int32_t)P->tie_ids.size();
    IdMap tie_map;
    tie_map.build(P->tie_ids);
    P->pt_tie.assign(n_cnt, -1);
    P->tie_pt.assign(P->tie_ids.size(), -1);
    for (size_t t = 0; t < P->tie_ids.size(); ++t) {
        const int32_t c = cnt_map.find(P->tie_ids[t]);
        P->tie_pt[t] = c;                                    // Buildxhat.m:110-122
        // isTie is decided per observation by its target ID, first TIE row wins (main.m:362-375)
        if (c >= 0 && pt_used[(size_t)c] && tie_map.find(P->tie_ids[t]) == (int32_t)t) P->pt_tie[(size_t)c] = (int32_t)t;
    }
    P->sec[0] = t1 - t0;
    P->sec[1] = t2 - t1;
    P->sec[2] = now_s() - t2;
    return FEBA_PACK_OK;
}

}  // namespace

extern "C" {

const char* feba_pack_last_error(void) { return g_pack_error.c_str(); }

void feba_pack_free(feba_pack* p) { delete p; }

int feba_pack_read(const char* pho, const char* ext, const char* cnt, const char* intr, const char* tie,
                   int32_t num_radial, int32_t all_gcp, int32_t threads, feba_pack** out) {
    if (out) *out = nullptr;
    if (!pho || !ext || !cnt || !intr || !out) return pack_fail(FEBA_PACK_ERR_FORMAT, "feba_pack_read: null argument");
    if (num_radial > 64) return pack_fail(FEBA_PACK_ERR_FORMAT, "feba_pack_read: Num_Radial_Distortions too large");
    feba_pack* P = new (std::nothrow) feba_pack();
    if (!P) return pack_fail(FEBA_PACK_ERR_IO, "out of host memory");
    P->nk = num_radial < 1 ? 1 : num_radial;                  // BuildAwG.m:18-20
    int rc;
    try {
        rc = build_pack(P, pho, ext, cnt, intr, tie, all_gcp, threads);
    } catch (const std::exception& ex) {
        rc = pack_fail(FEBA_PACK_ERR_IO, "feba_pack_read: %s", ex.what());
    }
    if (rc) {
        delete P;
        return rc;
    }
    *out = P;
    return FEBA_PACK_OK;
}

int feba_pack_get(const feba_pack* p, feba_pack_view* v) {
    if (!p || !v) return pack_fail(FEBA_PACK_ERR_FORMAT, "feba_pack_get: null argument");
    v->n_obs = (int64_t)p->obs_x.size();
    v->n_img = p->n_img;
    v->n_cam = p->n_cam;
    v->n_pts = p->n_pts;
    v->n_tie = p->n_tie;
    v->n_iop_cols = 3 + p->nk + 2;
    v->reserved = 0;
    v->obs_x = p->obs_x.data();
    v->obs_y = p->obs_y.data();
    v->obs_img = p->obs_img.data();
    v->obs_pt = p->obs_pt.data();
    v->img_cam = p->img_cam.data();
    v->eop0 = p->eop0.data();
    v->iop0 = p->iop0.data();
    v->cam_box = p->cam_box.data();
    v->xyz0 = p->xyz0.data();
    v->pt_tie = p->pt_tie.data();
    v->tie_pt = p->tie_pt.data();
    return FEBA_PACK_OK;
}

int feba_pack_problem(const feba_pack* p, const feba_settings* s, feba_problem* pr) {
    if (!p || !s || !pr) return pack_fail(FEBA_PACK_ERR_FORMAT, "feba_pack_problem: null argument");
    if ((s->num_radial < 1 ? 1 : s->num_radial) != p->nk)
        return pack_fail(FEBA_PACK_ERR_FORMAT, "feba_pack_problem: the pack was read with Num_Radial_Distortions = %d", p->nk);
    feba_pack_view v;
    feba_pack_get(p, &v);
    pr->n_obs = v.n_obs;
    pr->n_img = v.n_img;
    pr->n_cam = v.n_cam;
    pr->n_pts = v.n_pts;
    pr->n_tie = v.n_tie;
    pr->obs_x = v.obs_x;
    pr->obs_y = v.obs_y;
    pr->obs_img = v.obs_img;
    pr->obs_pt = v.obs_pt;
    pr->img_cam = v.img_cam;
    pr->eop0 = v.eop0;
    pr->iop0 = v.iop0;
    pr->cam_box = v.cam_box;
    pr->xyz0 = v.xyz0;
    pr->pt_tie = v.pt_tie;
    pr->settings = *s;
    return FEBA_PACK_OK;
}

int feba_pack_ids(const feba_pack* p, int32_t which, char* out, size_t cap, size_t* need) {
    if (!p || !need) return pack_fail(FEBA_PACK_ERR_FORMAT, "feba_pack_ids: null argument");
    const std::vector<sv>* ids = which == FEBA_PACK_IDS_TARGET   ? &p->point_ids
                                 : which == FEBA_PACK_IDS_IMAGE  ? &p->image_ids
                                 : which == FEBA_PACK_IDS_CAMERA ? &p->camera_ids
                                 : which == FEBA_PACK_IDS_TIE    ? &p->tie_ids
                                                                 : nullptr;
    if (!ids) return pack_fail(FEBA_PACK_ERR_FORMAT, "feba_pack_ids: unknown table %d", which);
    size_t total = 0;
    for (const sv& s : *ids) total += s.size() + 1;
    *need = total;
    if (!out || cap < total) return FEBA_PACK_OK;
    char* q = out;
    for (const sv& s : *ids) {
        if (!s.empty()) std::memcpy(q, s.data(), s.size());
        q += s.size();
        *q++ = '\n';
    }
    return FEBA_PACK_OK;
}

int feba_pack_timing(const feba_pack* p, double sec[3]) {
    if (!p || !sec) return FEBA_PACK_ERR_FORMAT;
    sec[0] = p->sec[0];
    sec[1] = p->sec[1];
    sec[2] = p->sec[2];
    return FEBA_PACK_OK;
}

}  // extern "C"
