// Host-side optimisation of the K1 tile schedule: bank-aware numbering of the staged elements of every tile.
//
// The incidence loop of assemble_kernel (tefem_assemble.cu) gathers, per incidence (element e, local nodes a, b), seven
// (K) / ten (with D) doubles from the element's record in shared memory: rec[stride * e + offset].  The lanes of a
// half-warp read pseudo-random records, and the shared-memory pipe - the unit that bounds the kernel - serialises lanes
// that fall into the same 8-byte bank with different addresses (14 of 33 wavefronts per element, profiles/README.md).
// Which bank a record starts in is decided by the TILE-LOCAL NUMBER of its element only: bank = (stride * e + offset)
// mod 16 with an odd stride, so e mod 16 (the element's "class") is a free parameter per staged element.  Two lanes x, y
// of one half-warp that execute the same load with offsets o_x, o_y collide iff
//
//        class_x - class_y  ==  inv(stride) * (o_y - o_x)      (mod 16),
//
// a pairwise condition.  This file replays the thread mapping of the kernel for every tile, accumulates for each pair of
// co-accessed elements how many loads collide at each class difference, and improves the assignment by exchanging the
// classes of two elements inside a block of 16 consecutive staged elements while that lowers the count (so the numbering
// stays a bijection and phase 1 of the kernel is untouched).  Nothing else changes: same lists, same order, same values
// (bitwise) - the kernel just finds its records in other banks.  The record strides of the two kernel families (13: no D, 17: with D) have
// different inverses mod 16 (5 and 1); `strides` selects which collisions are counted (both by default).
//
// Plain C++ (threads over tiles); called once per mesh by symbolic.py with host copies of the schedule arrays.
#include "tefem_common.cuh"

#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

namespace tefem {

namespace {

enum { H_NITEMS = 0, H_NELEMS = 1, H_NINC = 2, H_OITEMS = 5, H_OELEMS = 6, H_OINC = 7, H_NPIECE = 11, HDR_STRIDE = 12 };

// Tiles with the same items and incidence entries (structured meshes repeat a few hundred tile patterns) share one search:
// FNV-1a over the tile's item and incidence segments; a hit is verified byte by byte before its numbering is copied.
uint64_t fnv1a(uint64_t h, const void* data, size_t n) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

struct TileKey {
    const int32_t* hdr;
    const int32_t* item_inc;
    const int32_t* item_meta;
    const uint16_t* inc;
    uint64_t hash() const {
        uint64_t h = 1469598103934665603ull;
        const int32_t dims[4] = {hdr[H_NITEMS], hdr[H_NELEMS], hdr[H_NINC], hdr[H_NPIECE]};
        h = fnv1a(h, dims, sizeof(dims));
        h = fnv1a(h, item_inc + hdr[H_OITEMS], sizeof(int32_t) * (size_t)hdr[H_NITEMS]);
        h = fnv1a(h, item_meta + hdr[H_OITEMS], sizeof(int32_t) * (size_t)hdr[H_NITEMS]);
        return fnv1a(h, inc + hdr[H_OINC], sizeof(uint16_t) * (size_t)hdr[H_NINC]);
    }
    bool same(const TileKey& o) const {
        return hdr[H_NITEMS] == o.hdr[H_NITEMS] && hdr[H_NELEMS] == o.hdr[H_NELEMS] && hdr[H_NINC] == o.hdr[H_NINC] &&
               hdr[H_NPIECE] == o.hdr[H_NPIECE] &&
               memcmp(item_inc + hdr[H_OITEMS], o.item_inc + o.hdr[H_OITEMS], sizeof(int32_t) * (size_t)hdr[H_NITEMS]) == 0 &&
               memcmp(item_meta + hdr[H_OITEMS], o.item_meta + o.hdr[H_OITEMS], sizeof(int32_t) * (size_t)hdr[H_NITEMS]) == 0 &&
               memcmp(inc + hdr[H_OINC], o.inc + o.hdr[H_OINC], sizeof(uint16_t) * (size_t)hdr[H_NINC]) == 0;
    }
};

struct PairW {
    int32_t other;
    int32_t w[16];   // w[d]: loads that collide when class(this) - class(other) == d (mod 16)
};

// offsets (doubles inside the record) read for incidence (a, b); returns the number of loads
inline int record_offsets(int a, int b, bool with_d, int* off) {
    int n = 0;
    off[n++] = 3 * a; off[n++] = 3 * a + 1; off[n++] = 3 * a + 2;
    off[n++] = 12;
    if (a != b) { off[n++] = 3 * b; off[n++] = 3 * b + 1; off[n++] = 3 * b + 2; }
    if (with_d) {
        for (int c = 0; c < 3; ++c) off[n++] = (b == 0) ? 13 + c : 3 * (c + 1) + b - 1;   // GT[c][b], tefem_element.cuh
    }
    return n;
}

struct TileWork {
    std::vector<std::vector<PairW>> nbr;   // per element: collision weights against co-accessed elements
    std::vector<int32_t> cls;
};

// list[other].w[(sign * d) mod 16] += hist[d]; `slot` = position of `other` in the list + 1 (0: not yet there), a cell
// of the tile's dense (element x element) index map
void add_pair(std::vector<PairW>& list, int16_t& slot, int32_t other, const int* hist, bool negate) {
    if (slot == 0) {
        PairW p;
        p.other = other;
        for (int k = 0; k < 16; ++k) p.w[k] = 0;
        list.push_back(p);
        slot = (int16_t)list.size();
    }
    PairW& q = list[slot - 1];
    for (int d = 0; d < 16; ++d)
        if (hist[d]) q.w[negate ? (16 - d) & 15 : d] += hist[d];
}

// One tile: returns the new local number of every staged element in new_local[0 .. n_el) and the number of record slots.
int optimise_tile(const int32_t* hdr, const int32_t* item_inc, const int32_t* item_meta, const uint16_t* inc, int strides,
                  int n_sweeps, int block_size, int32_t* new_local, int64_t* cost_before, int64_t* cost_after) {
    const int n_items = hdr[H_NITEMS], n_el = hdr[H_NELEMS], n_piece = hdr[H_NPIECE];
    const int32_t* it_inc = item_inc + hdr[H_OITEMS];
    const int32_t* it_meta = item_meta + hdr[H_OITEMS];
    const uint16_t* inc_s = inc + hdr[H_OINC];
    TileWork T;
    T.nbr.assign(n_el, std::vector<PairW>());
    std::vector<int16_t> index((size_t)n_el * n_el, 0);   // (e, f) -> position of f in nbr[e] + 1
    // ---- replay: items [lo, hi) of a pass are mapped to consecutive threads, a half-warp is 16 consecutive items, a warp
    // iterates to its longest list (pieces-first layout: the piece pass, then the block pass whose combine items gather nothing)
    const int range_lo[2] = {0, n_piece}, range_hi[2] = {n_piece, n_items};
    for (int pass = (n_piece > 0 ? 0 : 1); pass < 2; ++pass) {
        const int lo = range_lo[pass], hi = range_hi[pass];
        for (int w0 = lo; w0 < hi; w0 += 16) {   // one half-warp
            int len[16], start[16], maxlen = 0;
            for (int l = 0; l < 16; ++l) {
                const int it = w0 + l;
                len[l] = 0; start[l] = 0;
                if (it < hi) {
                    const int32_t meta = it_meta[it];
                    const bool combine = (pass == 1 && n_piece > 0 && ((meta >> 8) & 0xffff) != 0);
                    len[l] = combine ? 0 : (meta & 0xff);
                    start[l] = (int)((uint32_t)it_inc[it] & 0xffffu);
                }
                maxlen = std::max(maxlen, len[l]);
            }
            for (int k = 0; k < maxlen; ++k) {
                int e[16], a[16], b[16];
                bool act[16];
                for (int l = 0; l < 16; ++l) {
                    act[l] = len[l] > k;
                    if (act[l]) {
                        const uint32_t w = inc_s[start[l] + k];
                        e[l] = (int)(w >> 4); a[l] = (int)((w >> 2) & 3); b[l] = (int)(w & 3);
                    }
                }
                for (int fam = 0; fam < 2; ++fam) {   // record family: 0 = stride 13 (no D), 1 = stride 17 (with D)
                    if (!(strides & (1 << fam))) continue;
                    const int inv = fam == 0 ? 5 : 1;   // 13 * 5 = 65 = 1, 17 = 1 (mod 16)
                    int off[16][10], n_off[16];
                    for (int l = 0; l < 16; ++l) n_off[l] = act[l] ? record_offsets(a[l], b[l], fam == 1, off[l]) : 0;
                    // load j of lane x and load j of lane y are the same instruction when both lanes execute it: the a-side
                    // loads (0..3) always, the b-side loads (4..6) unless a == b, then the GT loads
                    for (int x = 0; x < 16; ++x) {
                        if (!act[x]) continue;
                        for (int y = x + 1; y < 16; ++y) {
                            if (!act[y] || e[x] == e[y]) continue;   // same record: distinct offsets below 16 never share a bank
                            const bool bx = a[x] != b[x], by = a[y] != b[y];
                            int hist[16] = {0};   // loads of this half-warp step that collide at class difference d
                            for (int j = 0; j < 4; ++j) hist[((inv * (off[y][j] - off[x][j])) % 16 + 16) % 16] += 1;
                            if (bx && by)
                                for (int j = 4; j < 7; ++j) hist[((inv * (off[y][j] - off[x][j])) % 16 + 16) % 16] += 1;
                            if (fam == 1) {
                                const int gx = bx ? 7 : 4, gy = by ? 7 : 4;
                                for (int j = 0; j < 3; ++j) hist[((inv * (off[y][gy + j] - off[x][gx + j])) % 16 + 16) % 16] += 1;
                            }
                            add_pair(T.nbr[e[x]], index[(size_t)e[x] * n_el + e[y]], e[y], hist, false);
                            add_pair(T.nbr[e[y]], index[(size_t)e[y] * n_el + e[x]], e[x], hist, true);
                        }
                    }
                }
            }
        }
    }
    // ---- local search on the pairwise collision count.  Only the classes of the 16 consecutive staged elements of a block
    // are permuted among themselves: the new numbering is a bijection onto 0 .. n_el - 1 (no holes), and the half-warps of
    // phase 1 (thread le <-> staged element le) keep their element sets, so the coordinate gathers and the record stores of
    // phase 1 cost what they cost before (an unconstrained re-numbering would scatter the elements a half-warp stages:
    // modelled +2.4 wavefronts per element in phase 1, half of the gain).
    T.cls.resize(n_el);
    std::vector<int32_t> pos(n_el);                        // new local number of staged element i
    for (int i = 0; i < n_el; ++i) { pos[i] = i; T.cls[i] = i & 15; }
    auto total_cost = [&]() {
        int64_t c = 0;
        for (int i = 0; i < n_el; ++i)
            for (const PairW& p : T.nbr[i]) c += p.w[(T.cls[i] - T.cls[p.other] + 16) & 15];
        return c / 2;
    };
    auto mutual = [&](int i, int j) -> const int32_t* {   // w_ij or nullptr
        const int16_t slot = index[(size_t)i * n_el + j];
        return slot ? T.nbr[i][slot - 1].w : nullptr;
    };
    if (cost_before) *cost_before = total_cost();
    const int BS = block_size;                             // 16: phase 1 untouched; 32 / 64: its half-warps exchange neighbours
    std::vector<int64_t> C((size_t)BS * 16);
    for (int sweep = 0; sweep < n_sweeps; ++sweep) {
        int swaps = 0;
        for (int b0 = 0; b0 < n_el; b0 += BS) {
            const int nb = std::min(BS, n_el - b0);
            for (int round = 0; round < 8 * (BS / 16); ++round) {
                // C[m][r]: collisions of member m if it were in class r (others fixed)
                for (int m = 0; m < nb; ++m) {
                    int64_t* c = &C[(size_t)m * 16];
                    for (int r = 0; r < 16; ++r) c[r] = 0;
                    for (const PairW& p : T.nbr[b0 + m]) {
                        const int rf = T.cls[p.other];
                        for (int d = 0; d < 16; ++d) c[(d + rf) & 15] += p.w[d];
                    }
                }
                int64_t best = 0;
                int bi = -1, bj = -1;
                for (int x = 0; x < nb; ++x)
                    for (int y = x + 1; y < nb; ++y) {
                        const int i = b0 + x, j = b0 + y, ri = T.cls[i], rj = T.cls[j];
                        if (ri == rj) continue;
                        int64_t delta = C[(size_t)x * 16 + rj] - C[(size_t)x * 16 + ri] + C[(size_t)y * 16 + ri] - C[(size_t)y * 16 + rj];
                        const int32_t* w = mutual(i, j);
                        if (w) {
                            // C[x][rj] assumed j stays in rj (difference 0), C[y][ri] likewise; after the swap the pair sits at
                            // difference rj - ri; before it sat at ri - rj (counted once in C[x][ri] and once in C[y][rj])
                            delta += 2 * ((int64_t)w[(rj - ri + 16) & 15] - (int64_t)w[0]);
                        }
                        if (delta < best) { best = delta; bi = i; bj = j; }
                    }
                if (bi < 0) break;
                std::swap(T.cls[bi], T.cls[bj]);
                std::swap(pos[bi], pos[bj]);
                swaps += 1;
            }
        }
        if (swaps == 0) break;
    }
    if (cost_after) *cost_after = total_cost();
    // ---- new local numbers: a permutation of the positions inside every block
    for (int i = 0; i < n_el; ++i) new_local[i] = pos[i];
    return n_el;
}

}  // namespace

}  // namespace tefem

using namespace tefem;

// Bank-aware tile-local numbering of the staged elements (host function; all pointers are HOST pointers).
//   h_cta_hdr, h_item_inc, h_item_meta, h_inc   the schedule arrays of tefem_assemble (tile-packed, see include/tefem.h)
//   strides      bit 0: count the collisions of the kernels without D (record stride 13), bit 1: with D (stride 17)
//   h_new_local  [same indexing as cta_elems: o_elems of the tile + old local number] <- new local number
//   h_n_slots    [n_cta] <- number of record slots the tile needs (= n_elems: the numbering is a bijection)
//   h_cost       optional [2]: pairwise collision count summed over the tiles before / after
extern "C" int tefem_schedule_bank_numbering(int32_t n_cta, const int32_t* h_cta_hdr, const int32_t* h_item_inc,
                                             const int32_t* h_item_meta, const uint16_t* h_inc, int strides, int n_sweeps,
                                             int n_threads, int32_t* h_new_local, int32_t* h_n_slots, int64_t* h_cost) {
    TEFEM_REQUIRE(h_cta_hdr && h_item_inc && h_item_meta && h_inc && h_new_local && h_n_slots, "null pointer");
    TEFEM_REQUIRE(n_cta >= 0 && (strides & 3) != 0 && n_sweeps >= 0, "arguments");
    const int block_size = 16;   // classes are exchanged inside blocks of 16 staged elements: phase 1 keeps its half-warp sets
    if (n_threads <= 0) n_threads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 64u);
    n_threads = std::min(n_threads, std::max(1, n_cta));
    std::atomic<int32_t> next{0};
    std::atomic<int64_t> before{0}, after{0}, searched{0};
    struct Done { int32_t tile; int64_t cb, ca; };
    std::unordered_map<uint64_t, Done> done;       // hash -> a finished tile with that content
    std::mutex done_mutex;
    std::atomic<int> failed{0};
    auto worker = [&]() {
        try {
        for (;;) {
            const int32_t c = next.fetch_add(1);
            if (c >= n_cta || failed.load()) break;
            const int32_t* hdr = h_cta_hdr + (size_t)HDR_STRIDE * c;
            const TileKey key = {hdr, h_item_inc, h_item_meta, h_inc};
            const uint64_t h = key.hash();
            Done rep = {-1, 0, 0};
            {
                std::lock_guard<std::mutex> lock(done_mutex);
                auto it = done.find(h);
                if (it != done.end()) rep = it->second;
            }
            if (rep.tile >= 0) {
                const int32_t* rhdr = h_cta_hdr + (size_t)HDR_STRIDE * rep.tile;
                const TileKey rkey = {rhdr, h_item_inc, h_item_meta, h_inc};
                if (key.same(rkey)) {   // the finished tile's numbering is read-only from now on
                    memcpy(h_new_local + hdr[H_OELEMS], h_new_local + rhdr[H_OELEMS], sizeof(int32_t) * (size_t)hdr[H_NELEMS]);
                    h_n_slots[c] = hdr[H_NELEMS];
                    before.fetch_add(rep.cb);
                    after.fetch_add(rep.ca);
                    continue;
                }
            }
            int64_t cb = 0, ca = 0;
            h_n_slots[c] = optimise_tile(hdr, h_item_inc, h_item_meta, h_inc, strides, n_sweeps, block_size, h_new_local + hdr[H_OELEMS], &cb, &ca);
            before.fetch_add(cb);
            after.fetch_add(ca);
            searched.fetch_add(1);
            std::lock_guard<std::mutex> lock(done_mutex);
            done.emplace(h, Done{c, cb, ca});
        }
        } catch (...) {   // out of memory in a worker: report, the caller keeps the staged order
            failed.store(1);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) {
        try {
            pool.emplace_back(worker);
        } catch (...) {   // thread limit of the container: go on with the threads that started
            break;
        }
    }
    worker();
    for (std::thread& t : pool) t.join();
    if (failed.load()) {
        set_error("bank numbering: a worker failed (out of memory?)");
        return TEFEM_ERR_INVALID;
    }
    if (h_cost) { h_cost[0] = before.load(); h_cost[1] = after.load(); }
    if (getenv("TEFEM_SCHEDULE_VERBOSE"))
        fprintf(stderr, "[tefem schedule] bank numbering: %lld of %d tiles searched, the others share a pattern\n",
                (long long)searched.load(), (int)n_cta);
    return TEFEM_OK;
}
