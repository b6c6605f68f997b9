// backtrack.cpp -- host stage of the chaining path: chain extraction from (f, p) and anchor compaction.
//
// Same results as the reference's mg_chain_backtrack + compact_a (lchain.c:27-111), which mm2-gb also runs on the host
// after its kernels (gpu/plchain.cu:99-150).  Not a fallback for the device DP: this stage has no device version yet
// (SURVEY.md 8f N1).  Works on int32 predecessors (the device output) and plain caller buffers, so it can run on any
// thread; only the final copy into the driver's kalloc arena has to happen on the owning thread.
//
// Bit-exactness hinges on the order in which equal-score chain ends are visited: the reference sorts (score, index)
// pairs by score only with an UNSTABLE in-place MSD radix sort (ksort.h:98-151, 8-bit digits from bit 56 down,
// insertion sort for <= 64 elements), so that permutation is reproduced here step for step.
#include "../../include/mm2gb_chain.h"

#include <cstring>
#include <vector>

namespace {

struct Key { uint64_t key, val; };

void insertion_by_key(Key *beg, Key *end) // ksort.h:105-115 (stable for equal keys)
{
    for (Key *i = beg + 1; i < end; ++i) {
        if (i->key < (i - 1)->key) {
            Key tmp = *i, *j;
            for (j = i; j > beg && tmp.key < (j - 1)->key; --j) *j = *(j - 1);
            *j = tmp;
        }
    }
}

// ksort.h:116-145: American-flag permutation on digit (key >> shift) & 255, buckets visited in ascending order,
// each displaced element chased until something that belongs to the current bucket comes back.
void flag_pass(Key *beg, Key *end, int shift)
{
    struct Span { Key *cur, *end; } bk[256];
    for (auto &b : bk) b.cur = b.end = beg;
    for (Key *it = beg; it != end; ++it) ++bk[(it->key >> shift) & 255].end;
    for (int k = 1; k < 256; ++k) {
        bk[k].end += bk[k - 1].end - beg;
        bk[k].cur = bk[k - 1].end;
    }
    for (int k = 0; k < 256;) {
        Span &home = bk[k];
        if (home.cur == home.end) { ++k; continue; }
        int d = (int)((home.cur->key >> shift) & 255);
        if (d == k) { ++home.cur; continue; }
        Key carried = *home.cur;
        do {
            Key evicted = *bk[d].cur;
            *bk[d].cur++ = carried;
            carried = evicted;
            d = (int)((carried.key >> shift) & 255);
        } while (d != k);
        *home.cur++ = carried;
    }
    if (!shift) return;
    const int next = shift > 8 ? shift - 8 : 0;
    Key *lo = beg;
    for (int k = 0; k < 256; ++k) {
        Key *hi = bk[k].end;
        if (hi - lo > 64) flag_pass(lo, hi, next);
        else if (hi - lo > 1) insertion_by_key(lo, hi);
        lo = hi;
    }
}

// ksort.h:146-150.  The reference always starts at the top byte (shift 56).  A pass in which every key has the same
// digit is the identity (one bucket, nothing is displaced, and the recursion then continues on the whole range with the
// next digit), so starting at the highest digit in which the keys actually differ yields the same permutation while
// skipping the 5-6 no-op passes that 64-bit keys holding 17-bit scores would otherwise pay for.
void sort_by_key(Key *beg, Key *end)
{
    if (end - beg <= 64) { insertion_by_key(beg, end); return; }
    uint64_t all_or = 0, all_and = ~0ULL;
    for (Key *it = beg; it != end; ++it) all_or |= it->key, all_and &= it->key;
    const uint64_t diff = all_or ^ all_and;
    if (!diff) return; // all keys equal: every pass is the identity
    int shift = 56;
    while (((diff >> shift) & 255) == 0) shift -= 8;
    flag_pass(beg, end, shift);
}

// lchain.c:9-25: from chain end z, walk predecessors until a used anchor, the root, or a drop > max_drop;
// returns the anchor at which the chain is cut (exclusive), i.e. where score-so-far peaked.
int64_t find_cut(int32_t max_drop, const Key &z, const int32_t *f, const int32_t *p, int32_t *t)
{
    int64_t i = (int64_t)z.val, stop = -1, cut = i;
    int32_t top = 0;
    if (t[i] != 0) return i;
    do {
        t[i] = 2;
        stop = i = p[i];
        const int32_t s = i < 0 ? (int32_t)z.key : (int32_t)z.key - f[i];
        if (s > top) top = s, cut = i;
        else if (top - s > max_drop) break;
    } while (i >= 0 && t[i] == 0);
    for (i = (int64_t)z.val; i >= 0 && i != stop; i = p[i]) t[i] = 0;
    return cut;
}

} // namespace

extern "C" int32_t mm2gb_backtrack(int64_t n, const int32_t *f, const int32_t *p, const mm2gb_anchor_t *a, int32_t min_cnt,
                                   int32_t min_sc, int32_t max_drop, uint64_t *u, mm2gb_anchor_t *b, int64_t *n_b)
{
    if (n_b) *n_b = 0;
    if (n <= 0) return 0;
    // per-thread scratch, reused across reads (a fresh 24 B/anchor allocation per read costs more than the walk itself)
    static thread_local std::vector<Key> z, w;
    static thread_local std::vector<int32_t> t, v;
    static thread_local std::vector<uint64_t> uu;
    static thread_local std::vector<int64_t> start;
    // chain ends: every anchor scoring >= min_sc, visited from the highest score down (lchain.c:33-41)
    z.clear();
    for (int64_t i = 0; i < n; ++i)
        if (f[i] >= min_sc) z.push_back({(uint64_t)(int64_t)f[i], (uint64_t)i});
    if (z.empty()) return 0;
    sort_by_key(z.data(), z.data() + z.size());

    t.assign((size_t)n, 0);
    v.clear();
    uu.clear();
    for (int64_t k = (int64_t)z.size() - 1; k >= 0; --k) { // lchain.c:58-72
        if (t[z[k].val] != 0) continue;
        const size_t v0 = v.size();
        const int64_t cut = find_cut(max_drop, z[k], f, p, t.data());
        int64_t i;
        for (i = (int64_t)z[k].val; i != cut; i = p[i]) v.push_back((int32_t)i), t[i] = 1;
        const int32_t sc = i < 0 ? (int32_t)z[k].key : (int32_t)z[k].key - f[i];
        const int64_t cnt = (int64_t)(v.size() - v0);
        if (sc >= min_sc && cnt > 0 && cnt >= min_cnt) uu.push_back((uint64_t)sc << 32 | (uint64_t)cnt);
        else v.resize(v0);
    }
    const int32_t n_u = (int32_t)uu.size();
    if (n_u == 0) return 0;

    // lchain.c:78-111: each chain was collected end-first; flip it, then order chains by the x of their first anchor
    // with the same unstable sort (ties between chains starting at the same x follow it too).
    w.resize((size_t)n_u);
    start.resize((size_t)n_u);
    int64_t k = 0;
    for (int32_t c = 0; c < n_u; ++c) {
        const int32_t cnt = (int32_t)uu[(size_t)c];
        start[(size_t)c] = k;
        w[(size_t)c] = {a[v[(size_t)(k + cnt - 1)]].x, (uint64_t)k << 32 | (uint64_t)c};
        k += cnt;
    }
    sort_by_key(w.data(), w.data() + n_u);
    int64_t out = 0;
    for (int32_t c = 0; c < n_u; ++c) {
        const int32_t src = (int32_t)w[(size_t)c].val, cnt = (int32_t)uu[(size_t)src];
        const int64_t s0 = start[(size_t)src];
        u[c] = uu[(size_t)src];
        for (int32_t j = 0; j < cnt; ++j) b[out + j] = a[v[(size_t)(s0 + cnt - 1 - j)]];
        out += cnt;
    }
    if (n_b) *n_b = out;
    return n_u;
}
