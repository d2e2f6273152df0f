// Chunk schedule of the assembly (host side, plain C++17; built once per handle).
//
// The object points are processed in the order of a space-filling curve (feba_api.cu), so consecutive points are
// seen by the same images.  The sorted observations are cut into CHUNKS of consecutive points (at most
// kChunkObs observations / kChunkPts points); the point pass writes the per-observation records at the
// observation's own position, i.e. chunk-major and contiguous, and one CTA per chunk then forms everything the
// chunk contributes to the reduced system from ITS OWN ~200 KB of records (one LANE per image or image pair):
//   per image seen in the chunk  : diagonal block Je'(P - Z Z')Je, right-hand side Je' r, camera x image block H Je
//   per image pair (row(a) >= row(b)) sharing tie points in the chunk : sum Je_a' Z_a Z_b' Je_b
// as PARTIAL blocks (one writer each), which a last pass sums per target block in a fixed order -- no atomics,
// bit-identical reruns.  What the schedule saves over the image-major form of round 1: every record is read from
// HBM once (it was gathered ~9 times by the image-pair pass: 14.2 GB of the assembly's 23.6 GB DRAM traffic on
// BASELINE configs[3]), and the pair list shrinks to two 16-bit chunk-local indices per pair.
//
// Replaces the same reference lines as feba_assemble.cu (BuildAwG.m:46-512, main.m:424-425).
#pragma once
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

namespace feba {

constexpr int kChunkObs = 416;    // observations per chunk (soft cap: a single point may exceed it); 416 records of
                                  // 33 doubles = 107 KB of shared memory: two chunks in flight per SM
constexpr int kChunkPts = 64;     // points per chunk
constexpr int kImgPart = 108;     // doubles per (chunk, image) partial: 21 + 6 + 6 (NK + 5) <= 105, padded
constexpr int kBlkPart = 36;      // doubles per (chunk, image pair) partial

struct ChunkSchedule {
    int n_chunks = 0;
    std::vector<int> seg0;            // n_chunks + 1: first segment (point) of every chunk
    std::vector<int> obs0;            // n_chunks + 1: first observation
    // images of a chunk
    std::vector<int> img0;            // n_chunks + 1 -> image slots
    std::vector<int> slot_img;        // image of every image slot
    std::vector<int> slot_obs0;       // slots + 1 -> slot_obs
    std::vector<uint16_t> slot_obs;   // chunk-local observation index, ascending
    // image pairs of a chunk
    std::vector<int> blk0;            // n_chunks + 1 -> block slots
    std::vector<int> bslot_a, bslot_b;  // images of every block slot, row(a) >= row(b)
    std::vector<int> bslot_pair0;     // block slots + 1 -> pairs
    std::vector<uint32_t> pairs;      // chunk-local observations: a | b << 16
    // final sums: per image the image slots, per distinct image pair the block slots (ascending = chunk order)
    std::vector<int> timg_ptr, timg_slots;              // n_img + 1
    std::vector<int> tblk_a, tblk_b, tblk_ptr, tblk_slots;
    long long n_pairs = 0;
    bool ok = true;                   // false: a chunk does not fit 16-bit local indices (use the image-major form)
};

// seg_start / simg: observations sorted by point (segment s = one point), seg_tie[s]: the point is a tie point,
// img_row: row of each image in the reduced system (orientation of the pairs: the lower triangle is stored).
inline ChunkSchedule build_chunks(int n_img, int n_seg, const int* seg_start, const int* simg,
                                  const unsigned char* seg_tie, const int* img_row) {
    ChunkSchedule C;
    // ---- cut
    C.seg0.push_back(0);
    C.obs0.push_back(0);
    {
        int s0 = 0;
        while (s0 < n_seg) {
            int s1 = s0 + 1;
            while (s1 < n_seg && s1 - s0 < kChunkPts && seg_start[s1 + 1] - seg_start[s0] <= kChunkObs) ++s1;
            if (seg_start[s1] - seg_start[s0] > 65535) C.ok = false;
            C.seg0.push_back(s1);
            C.obs0.push_back(seg_start[s1]);
            s0 = s1;
        }
    }
    C.n_chunks = (int)C.seg0.size() - 1;
    if (!C.ok || C.n_chunks == 0) return C;
    // ---- per chunk: images, their observations, image pairs and their observation pairs (threads over chunks)
    struct Local {
        std::vector<int> img0, slot_img, slot_cnt, blk0, ba, bb, bcnt;
        std::vector<uint16_t> slot_obs;
        std::vector<uint32_t> pairs;
    };
    unsigned hw = std::thread::hardware_concurrency();
    int nth = (int)(hw ? hw : 1);
    if (nth > 16) nth = 16;
    if (C.n_chunks < 64) nth = 1;
    std::vector<Local> loc((size_t)nth);
    auto work = [&](int t) {
        Local& L = loc[(size_t)t];
        const int c_lo = (int)((long long)C.n_chunks * t / nth), c_hi = (int)((long long)C.n_chunks * (t + 1) / nth);
        std::vector<int> imgs, lidx((size_t)n_img, -1);
        struct PairKey { uint32_t key; uint32_t pr; };
        std::vector<PairKey> pk;
        std::vector<std::pair<size_t, int>> runs;
        for (int c = c_lo; c < c_hi; ++c) {
            const int o0 = C.obs0[(size_t)c], o1 = C.obs0[(size_t)c + 1];
            imgs.clear();
            for (int o = o0; o < o1; ++o)
                if (lidx[(size_t)simg[o]] < 0) {
                    lidx[(size_t)simg[o]] = 0;
                    imgs.push_back(simg[o]);
                }
            std::sort(imgs.begin(), imgs.end());
            for (size_t k = 0; k < imgs.size(); ++k) lidx[(size_t)imgs[k]] = (int)k;
            L.img0.push_back((int)L.slot_img.size());
            const size_t base = L.slot_img.size();
            for (int im : imgs) {
                L.slot_img.push_back(im);
                L.slot_cnt.push_back(0);
            }
            for (int o = o0; o < o1; ++o) ++L.slot_cnt[base + (size_t)lidx[(size_t)simg[o]]];
            {
                std::vector<size_t> cur(imgs.size());
                size_t w = L.slot_obs.size();
                for (size_t k = 0; k < imgs.size(); ++k) {
                    cur[k] = w;
                    w += (size_t)L.slot_cnt[base + k];
                }
                L.slot_obs.resize(w);
                for (int o = o0; o < o1; ++o) L.slot_obs[cur[(size_t)lidx[(size_t)simg[o]]]++] = (uint16_t)(o - o0);
            }
            // pairs of the tie points: (o, b) with the image of b earlier in the row order, or the same image
            pk.clear();
            for (int s = C.seg0[(size_t)c]; s < C.seg0[(size_t)c + 1]; ++s) {
                if (!seg_tie[s]) continue;
                for (int o = seg_start[s]; o < seg_start[s + 1]; ++o) {
                    const int ia = simg[o], ra = img_row[ia];
                    for (int b = seg_start[s]; b < seg_start[s + 1]; ++b) {
                        const int ib = simg[b];
                        if (img_row[ib] < ra || (ib == ia && b != o))
                            pk.push_back({(uint32_t)lidx[(size_t)ia] << 16 | (uint32_t)lidx[(size_t)ib],
                                          (uint32_t)(o - o0) | (uint32_t)(b - o0) << 16});
                    }
                }
            }
            std::stable_sort(pk.begin(), pk.end(), [](const PairKey& x, const PairKey& y) { return x.key < y.key; });
            L.blk0.push_back((int)L.ba.size());
            // runs of equal keys = image pairs; longest first: one LANE works through one image pair, so lanes of a
            // warp should have about the same number of observation pairs
            runs.clear();
            for (size_t q = 0; q < pk.size(); ++q) {
                if (q == 0 || pk[q].key != pk[q - 1].key) runs.push_back({q, 0});
                ++runs.back().second;
            }
            std::stable_sort(runs.begin(), runs.end(),
                             [](const std::pair<size_t, int>& x, const std::pair<size_t, int>& y) { return x.second > y.second; });
            for (const auto& r : runs) {
                L.ba.push_back(imgs[pk[r.first].key >> 16]);
                L.bb.push_back(imgs[pk[r.first].key & 0xffffu]);
                L.bcnt.push_back(r.second);
                for (int q = 0; q < r.second; ++q) L.pairs.push_back(pk[r.first + (size_t)q].pr);
            }
            for (int im : imgs) lidx[(size_t)im] = -1;
        }
    };
    if (nth == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nth; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    // ---- concatenate
    for (const Local& L : loc) {
        const int s_off = (int)C.slot_img.size(), b_off = (int)C.bslot_a.size();
        for (int v : L.img0) C.img0.push_back(v + s_off);
        for (int v : L.blk0) C.blk0.push_back(v + b_off);
        C.slot_img.insert(C.slot_img.end(), L.slot_img.begin(), L.slot_img.end());
        C.bslot_a.insert(C.bslot_a.end(), L.ba.begin(), L.ba.end());
        C.bslot_b.insert(C.bslot_b.end(), L.bb.begin(), L.bb.end());
    }
    C.img0.push_back((int)C.slot_img.size());
    C.blk0.push_back((int)C.bslot_a.size());
    {
        C.slot_obs0.assign(C.slot_img.size() + 1, 0);
        C.bslot_pair0.assign(C.bslot_a.size() + 1, 0);
        size_t si = 0, bi = 0;
        for (const Local& L : loc) {
            for (int cnt : L.slot_cnt) {
                C.slot_obs0[si + 1] = C.slot_obs0[si] + cnt;
                ++si;
            }
            for (int cnt : L.bcnt) {
                C.bslot_pair0[bi + 1] = C.bslot_pair0[bi] + cnt;
                ++bi;
            }
            C.slot_obs.insert(C.slot_obs.end(), L.slot_obs.begin(), L.slot_obs.end());
            C.pairs.insert(C.pairs.end(), L.pairs.begin(), L.pairs.end());
        }
    }
    C.n_pairs = (long long)C.pairs.size();
    // ---- final sums: image slots by image (counting sort keeps chunk order), block slots by image pair
    C.timg_ptr.assign((size_t)n_img + 1, 0);
    for (int im : C.slot_img) ++C.timg_ptr[(size_t)im + 1];
    for (int i = 0; i < n_img; ++i) C.timg_ptr[(size_t)i + 1] += C.timg_ptr[(size_t)i];
    C.timg_slots.resize(C.slot_img.size());
    {
        std::vector<int> cur(C.timg_ptr.begin(), C.timg_ptr.end() - 1);
        for (size_t s = 0; s < C.slot_img.size(); ++s) C.timg_slots[(size_t)cur[(size_t)C.slot_img[s]]++] = (int)s;
    }
    {
        std::vector<std::pair<unsigned long long, int>> key(C.bslot_a.size());
        for (size_t s = 0; s < key.size(); ++s)
            key[s] = {(unsigned long long)C.bslot_a[s] * (unsigned long long)n_img + (unsigned long long)C.bslot_b[s], (int)s};
        std::sort(key.begin(), key.end());
        C.tblk_ptr.push_back(0);
        for (size_t q = 0; q < key.size(); ++q) {
            if (q == 0 || key[q].first != key[q - 1].first) {
                if (q) C.tblk_ptr.push_back((int)q);
                C.tblk_a.push_back((int)(key[q].first / (unsigned long long)n_img));
                C.tblk_b.push_back((int)(key[q].first % (unsigned long long)n_img));
            }
            C.tblk_slots.push_back(key[q].second);
        }
        C.tblk_ptr.push_back((int)key.size());
        if (key.empty()) C.tblk_ptr.assign(1, 0);
    }
    return C;
}

}  // namespace feba
