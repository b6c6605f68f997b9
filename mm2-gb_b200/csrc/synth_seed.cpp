// synth_seed.cpp -- synthetic workload generator for the chaining benchmarks (libmm2gb_synth.so).
//
// Produces what the chaining stage is fed by minimap2's seeding (map.c:295-331 collect_seed_hits): per read, an
// x-sorted array of 16-byte anchors.  The pipeline is minimap2-shaped -- random reference (optionally with planted,
// diverged repeat copies), ONT-like reads (substitutions / deletions / insertions, half of them reverse-complemented),
// (w,k)-minimizers of reference and reads, an occurrence filter, one anchor per minimizer hit -- but it is an
// independent implementation for generating INPUT data at benchmark scale in seconds on the host; it is not part of
// the chaining path and makes no claim to reproduce minimap2's seeds bit for bit (seeding stays on the CPU in the
// reference driver).  Anchor packing follows minimap.h:72 / map.c:311-325 so the arrays are
// valid input for lchain.c and for the device path alike.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Rng { // splitmix64
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { uint64_t z = (s += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return n ? next() % n : 0; }
};

inline uint64_t mix_kmer(uint64_t key, uint64_t mask) // invertible-style scramble of a 2k-bit k-mer code
{
    key = (key ^ (key >> 17)) * 0xed5ad4bbULL & mask;
    key = (key ^ (key >> 11)) * 0xac4c1b51ULL & mask;
    key = (key ^ (key >> 15)) * 0x31848babULL & mask;
    return (key ^ (key >> 14)) & mask;
}

struct Mz { uint64_t h; uint32_t pos; uint8_t strand; }; // pos = index of the k-mer's last base

// (w,k)-minimizers of seq[0..n): for every window of w consecutive k-mers the smallest hash (rightmost on ties),
// each reported once.  Strand = which of forward / reverse-complement k-mer is canonical; palindromes are skipped.
void sketch(const uint8_t *seq, int64_t n, int k, int w, std::vector<Mz> &out)
{
    const uint64_t mask = (1ULL << (2 * k)) - 1, shift = 2 * (k - 1);
    uint64_t fw = 0, rv = 0;
    std::vector<Mz> ring((size_t)w);
    std::vector<uint8_t> valid((size_t)w, 0);
    int64_t last_pos = -1;
    for (int64_t i = 0, l = 0; i < n; ++i) {
        const uint64_t c = seq[i] & 3;
        fw = (fw << 2 | c) & mask;
        rv = rv >> 2 | (3ULL ^ c) << shift;
        ++l;
        const int slot = (int)(i % w);
        valid[(size_t)slot] = 0;
        if (l >= k && fw != rv) {
            const uint8_t s = fw < rv ? 0 : 1;
            ring[(size_t)slot] = {mix_kmer(s ? rv : fw, mask), (uint32_t)i, s};
            valid[(size_t)slot] = 1;
        }
        if (l >= k + w - 1) {
            int best = -1;
            for (int j = 0; j < w; ++j) {
                const int q = (int)((i + 1 + j) % w); // oldest .. newest
                if (valid[(size_t)q] && (best < 0 || ring[(size_t)q].h <= ring[(size_t)best].h)) best = q;
            }
            if (best >= 0 && (int64_t)ring[(size_t)best].pos != last_pos) {
                out.push_back(ring[(size_t)best]);
                last_pos = ring[(size_t)best].pos;
            }
        }
    }
}

struct Index {
    int bucket_bits = 24, k = 15;
    std::vector<uint32_t> bstart;     // bucket -> first entry
    std::vector<uint64_t> key;        // hash per entry (sorted inside a bucket)
    std::vector<uint32_t> val;        // pos << 1 | strand
    int mid_occ = 10;
};

struct Workload {
    std::vector<uint64_t> a; // x,y interleaved
    std::vector<int64_t> off;
};

template <class F>
void parallel_for(int64_t n, int n_threads, F fn)
{
    std::atomic<int64_t> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([&, t]() { for (;;) { int64_t i = next.fetch_add(1); if (i >= n) break; fn(i, t); } });
    for (auto &x : th) x.join();
}

} // namespace

extern "C" {

// Optional shape parameters of a workload (mm2gb_synth_create_ex); all zero = the plain workload of mm2gb_synth_create.
struct mm2gb_synth_extra_t {
    // "human-scale" hit mix without building a 3 Gb index: every read minimizer additionally gets Poisson(lambda) hits at uniformly
    // random positions of bg_contigs virtual contigs (ids after the real ones, 50-250 Mb each, bg_len bp in total), on a random
    // strand; lambda = (minimizers per bp of the real reference) x bg_len / 4^k, i.e. the chance hits a random reference of that
    // size produces (~0.5 per minimizer for 3 Gb at k = 15, w = 10).
    int64_t bg_len;
    int bg_contigs;
    // a tandem array (tandem_copies x tandem_unit bp, each copy diverged by tandem_div) planted in the middle of contig 0; a fraction
    // tandem_read_frac of the reads is drawn from inside it.  With mid_occ > 0 the occurrence filter is fixed at that value instead of
    // the 2e-4 quantile (minimap2 -f <large>): repetitive seeds are kept, windows overflow max_iter.
    int tandem_copies, tandem_unit;
    double tandem_div, tandem_read_frac;
    int mid_occ;
};

void *mm2gb_synth_create_ex(uint64_t seed, uint64_t read_seed, int64_t ref_len, int64_t contig_len, int n_repeat_copies, int repeat_unit, double repeat_div,
                            int n_reads, int len_lo, int len_hi, double err, int k, int w, int n_threads, const mm2gb_synth_extra_t *extra);

// Generates the workload (`seed` fixes the reference, `read_seed` the reads); returns an opaque handle (NULL on bad arguments).  n_total / n_reads via the getters.
void *mm2gb_synth_create(uint64_t seed, uint64_t read_seed, int64_t ref_len, int64_t contig_len, int n_repeat_copies, int repeat_unit, double repeat_div,
                         int n_reads, int len_lo, int len_hi, double err, int k, int w, int n_threads)
{
    return mm2gb_synth_create_ex(seed, read_seed, ref_len, contig_len, n_repeat_copies, repeat_unit, repeat_div, n_reads, len_lo, len_hi, err, k, w,
                                 n_threads, nullptr);
}

void *mm2gb_synth_create_ex(uint64_t seed, uint64_t read_seed, int64_t ref_len, int64_t contig_len, int n_repeat_copies, int repeat_unit, double repeat_div,
                            int n_reads, int len_lo, int len_hi, double err, int k, int w, int n_threads, const mm2gb_synth_extra_t *extra)
{
    mm2gb_synth_extra_t X;
    memset(&X, 0, sizeof(X));
    if (extra) X = *extra;
    if (ref_len < 1000 || n_reads < 0 || k < 8 || k > 28 || w < 1 || w > 64 || len_lo < k || len_hi < len_lo || ref_len >= (1LL << 32))
        return nullptr;
    if (n_threads < 1) n_threads = 1;
    if (contig_len <= 0 || contig_len > ref_len) contig_len = ref_len;
    if (len_hi >= contig_len) len_hi = (int)contig_len - 1;
    if (len_lo > len_hi) len_lo = len_hi;
    // ---- reference
    std::vector<uint8_t> ref((size_t)ref_len);
    {
        const int64_t chunk = 1 << 20, nchunk = (ref_len + chunk - 1) / chunk;
        parallel_for(nchunk, n_threads, [&](int64_t c, int) {
            Rng r(seed * 0x100000001b3ULL + 7919 * (uint64_t)c + 1);
            const int64_t e = std::min(ref_len, (c + 1) * chunk);
            for (int64_t i = c * chunk; i < e; i += 32) {
                uint64_t v = r.next();
                for (int64_t j = i; j < std::min(e, i + 32); ++j, v >>= 2) ref[(size_t)j] = v & 3;
            }
        });
        if (n_repeat_copies > 0 && repeat_unit > 0 && repeat_unit < ref_len) {
            Rng r(seed ^ 0xabcdef12345ULL);
            std::vector<uint8_t> unit((size_t)repeat_unit);
            for (auto &b : unit) b = r.next() & 3;
            for (int c = 0; c < n_repeat_copies; ++c) {
                const int64_t pos = (int64_t)r.below((uint64_t)(ref_len - repeat_unit));
                for (int j = 0; j < repeat_unit; ++j)
                    ref[(size_t)(pos + j)] = r.uniform() < repeat_div ? (uint8_t)((unit[(size_t)j] + 1 + r.below(3)) & 3) : unit[(size_t)j];
            }
        }
    }
    int64_t tandem_lo = 0, tandem_hi = 0;
    if (X.tandem_copies > 0 && X.tandem_unit > 0 && (int64_t)X.tandem_copies * X.tandem_unit < contig_len / 2) {
        Rng r(seed ^ 0x7a4de3ULL);
        std::vector<uint8_t> unit((size_t)X.tandem_unit);
        for (auto &b : unit) b = r.next() & 3;
        const int64_t span = (int64_t)X.tandem_copies * X.tandem_unit;
        tandem_lo = (contig_len - span) / 2;
        tandem_hi = tandem_lo + span;
        for (int64_t i = 0; i < span; ++i) {
            const uint8_t b = unit[(size_t)(i % X.tandem_unit)];
            ref[(size_t)(tandem_lo + i)] = r.uniform() < X.tandem_div ? (uint8_t)((b + 1 + r.below(3)) & 3) : b;
        }
    }
    // ---- index: minimizers of the reference, bucketed by the top bits of the hash
    Index idx;
    idx.k = k;
    {
        const int64_t chunk = 4 << 20, nchunk = (ref_len + chunk - 1) / chunk;
        std::vector<std::vector<Mz>> part((size_t)nchunk);
        parallel_for(nchunk, n_threads, [&](int64_t c, int) {
            const int64_t s = std::max<int64_t>(0, c * chunk - (k + w - 2)), e = std::min(ref_len, (c + 1) * chunk);
            std::vector<Mz> tmp;
            sketch(ref.data() + s, e - s, k, w, tmp);
            auto &dst = part[(size_t)c];
            for (auto &m : tmp) { // keep minimizers whose k-mer ends inside this chunk (overlap region belongs to the previous one)
                const int64_t gp = s + m.pos;
                if (gp >= c * chunk) dst.push_back({m.h, (uint32_t)gp, m.strand});
            }
        });
        size_t n_mz = 0;
        for (auto &p : part) n_mz += p.size();
        const int hb = 2 * k;
        idx.bucket_bits = std::min(24, hb);
        const int sh = hb - idx.bucket_bits;
        const size_t nb = (size_t)1 << idx.bucket_bits;
        idx.bstart.assign(nb + 1, 0);
        for (auto &p : part) for (auto &m : p) ++idx.bstart[(size_t)(m.h >> sh) + 1];
        for (size_t b = 0; b < nb; ++b) idx.bstart[b + 1] += idx.bstart[b];
        idx.key.resize(n_mz); idx.val.resize(n_mz);
        std::vector<uint32_t> cur(idx.bstart.begin(), idx.bstart.end() - 1);
        for (auto &p : part) for (auto &m : p) {
            const uint32_t at = cur[(size_t)(m.h >> sh)]++;
            idx.key[at] = m.h; idx.val[at] = m.pos << 1 | m.strand;
        }
        part.clear(); part.shrink_to_fit();
        // sort inside buckets by (hash, pos) and histogram the occurrence counts for the mid_occ cut-off
        std::vector<std::vector<uint32_t>> occ_hist((size_t)n_threads, std::vector<uint32_t>(4096, 0));
        const int64_t bchunk = 1 << 14, nbchunk = ((int64_t)nb + bchunk - 1) / bchunk;
        parallel_for(nbchunk, n_threads, [&](int64_t c, int t) {
            std::vector<std::pair<uint64_t, uint32_t>> tmp;
            for (size_t b = (size_t)(c * bchunk); b < std::min(nb, (size_t)((c + 1) * bchunk)); ++b) {
                const uint32_t s = idx.bstart[b], e = idx.bstart[b + 1];
                if (e - s > 1) {
                    tmp.clear();
                    for (uint32_t i = s; i < e; ++i) tmp.push_back({idx.key[i], idx.val[i]});
                    std::sort(tmp.begin(), tmp.end());
                    for (uint32_t i = s; i < e; ++i) { idx.key[i] = tmp[i - s].first; idx.val[i] = tmp[i - s].second; }
                }
                for (uint32_t i = s; i < e;) {
                    uint32_t j = i + 1;
                    while (j < e && idx.key[j] == idx.key[i]) ++j;
                    ++occ_hist[(size_t)t][std::min<uint32_t>(j - i, 4095)];
                    i = j;
                }
            }
        });
        // like mm_idx_cal_max_occ with -f 2e-4: the occurrence count exceeded by only 0.02% of distinct minimizers, >= 10
        std::vector<uint64_t> hist(4096, 0);
        uint64_t distinct = 0;
        for (auto &h : occ_hist) for (int i = 0; i < 4096; ++i) hist[(size_t)i] += h[(size_t)i], distinct += h[(size_t)i];
        uint64_t above = 0, lim = (uint64_t)(distinct * 2e-4);
        int cut = 4095;
        for (; cut > 0; --cut) { above += hist[(size_t)cut]; if (above > lim) break; }
        idx.mid_occ = std::max(10, cut + 1);
        if (X.mid_occ > 0) idx.mid_occ = X.mid_occ;
    }
    // background hits of a reference much larger than the one that was built
    std::vector<int64_t> bg_start;     // cumulative lengths of the virtual contigs
    double bg_lambda = 0.0;
    if (X.bg_len > 0 && X.bg_contigs > 0) {
        Rng r(seed ^ 0x5eedb6ULL);
        std::vector<double> len((size_t)X.bg_contigs);
        double sum = 0;
        for (auto &l : len) { l = 50e6 + 200e6 * r.uniform(); sum += l; }
        bg_start.assign((size_t)X.bg_contigs + 1, 0);
        for (int c = 0; c < X.bg_contigs; ++c)
            bg_start[(size_t)c + 1] = bg_start[(size_t)c] + std::min<int64_t>((int64_t)(len[(size_t)c] * (double)X.bg_len / sum), (1LL << 31) - 1);
        double space = 1.0;
        for (int i = 0; i < k; ++i) space *= 4.0;
        bg_lambda = (double)idx.key.size() / (double)ref_len * (double)bg_start.back() / space;
    }
    // ---- reads -> anchors
    Workload *wl = new Workload();
    std::vector<std::vector<uint64_t>> per_read((size_t)n_reads);
    const int64_t n_contigs = (ref_len + contig_len - 1) / contig_len;
    const int sh = 2 * k - idx.bucket_bits;
    parallel_for(n_reads, n_threads, [&](int64_t r, int) {
        Rng rng(read_seed * 6364136223846793005ULL + 1442695040888963407ULL * (uint64_t)(r + 1));
        const int64_t ln = len_lo + (int64_t)rng.below((uint64_t)(len_hi - len_lo + 1));
        int64_t contig = (int64_t)rng.below((uint64_t)n_contigs);
        const bool from_tandem = tandem_hi > tandem_lo && rng.uniform() < X.tandem_read_frac;
        if (from_tandem) contig = 0;
        const int64_t c0 = contig * contig_len, c1 = std::min(ref_len, c0 + contig_len);
        const int64_t span = std::min(ln, c1 - c0 - 1);
        int64_t st = c0 + (int64_t)rng.below((uint64_t)(c1 - c0 - span));
        if (from_tandem)   // anywhere that keeps the read inside the array (or centred on it if the read is longer)
            st = span < tandem_hi - tandem_lo ? tandem_lo + (int64_t)rng.below((uint64_t)(tandem_hi - tandem_lo - span))
                                              : std::max<int64_t>(c0, std::min<int64_t>(c1 - span - 1, (tandem_lo + tandem_hi - span) / 2));
        std::vector<uint8_t> q;
        q.reserve((size_t)(span + span / 8));
        for (int64_t i = 0; i < span; ++i) {
            const double u = rng.uniform();
            const uint8_t b = ref[(size_t)(st + i)];
            if (u < 0.4 * err) q.push_back((uint8_t)((b + 1 + rng.below(3)) & 3));   // substitution
            else if (u < 0.7 * err) continue;                                          // deletion
            else if (u < err) { q.push_back(b); q.push_back((uint8_t)(rng.next() & 3)); } // insertion
            else q.push_back(b);
        }
        if (rng.next() & 1) { std::reverse(q.begin(), q.end()); for (auto &b : q) b = 3 - b; }
        const int64_t qlen = (int64_t)q.size();
        std::vector<Mz> mz;
        sketch(q.data(), qlen, k, w, mz);
        std::vector<std::pair<uint64_t, uint64_t>> hits;
        if (bg_lambda > 0.0) {
            const double p0 = std::exp(-bg_lambda);
            const int64_t bg_total = bg_start.back();
            for (auto &m : mz) {
                double u = rng.uniform(), pk = p0, cum = p0;   // Poisson(lambda) by inversion
                int nh = 0;
                while (u > cum && nh < 64) { ++nh; pk *= bg_lambda / nh; cum += pk; }
                for (int h = 0; h < nh; ++h) {
                    const int64_t g = (int64_t)rng.below((uint64_t)bg_total);
                    const int64_t c = (int64_t)(std::upper_bound(bg_start.begin(), bg_start.end(), g) - bg_start.begin()) - 1;
                    const uint64_t rid = (uint64_t)(n_contigs + c), rpos = (uint64_t)(g - bg_start[(size_t)c]);
                    if (rng.next() & 1) hits.push_back({rid << 32 | rpos, (uint64_t)k << 32 | m.pos});
                    else hits.push_back({1ULL << 63 | rid << 32 | rpos, (uint64_t)k << 32 | (uint64_t)(qlen - ((int64_t)m.pos + 1 - k) - 1)});
                }
            }
        }
        for (auto &m : mz) {
            const size_t b = (size_t)(m.h >> sh);
            uint32_t s = idx.bstart[b], e = idx.bstart[b + 1];
            while (s < e && idx.key[s] < m.h) ++s;
            uint32_t t = s;
            while (t < e && idx.key[t] == m.h) ++t;
            if (t == s || (int)(t - s) > idx.mid_occ) continue;
            for (uint32_t i = s; i < t; ++i) {
                const uint64_t gpos = idx.val[i] >> 1, rid = gpos / (uint64_t)contig_len, rpos = gpos % (uint64_t)contig_len;
                uint64_t x, y;
                if ((idx.val[i] & 1) == m.strand) { // same strand (map.c:311-313)
                    x = rid << 32 | rpos;
                    y = (uint64_t)k << 32 | m.pos;
                } else {                            // opposite strand (map.c:314-316)
                    x = 1ULL << 63 | rid << 32 | rpos;
                    y = (uint64_t)k << 32 | (uint64_t)(qlen - ((int64_t)m.pos + 1 - k) - 1);
                }
                hits.push_back({x, y});
            }
        }
        std::sort(hits.begin(), hits.end(), [](const std::pair<uint64_t, uint64_t> &p, const std::pair<uint64_t, uint64_t> &q2) { return p.first < q2.first; });
        auto &dst = per_read[(size_t)r];
        dst.reserve(hits.size() * 2);
        for (auto &h : hits) { dst.push_back(h.first); dst.push_back(h.second); }
    });
    wl->off.assign((size_t)n_reads + 1, 0);
    for (int r = 0; r < n_reads; ++r) wl->off[(size_t)r + 1] = wl->off[(size_t)r] + (int64_t)per_read[(size_t)r].size() / 2;
    wl->a.resize((size_t)wl->off[(size_t)n_reads] * 2);
    parallel_for(n_reads, n_threads, [&](int64_t r, int) {
        if (!per_read[(size_t)r].empty())
            memcpy(wl->a.data() + 2 * wl->off[(size_t)r], per_read[(size_t)r].data(), per_read[(size_t)r].size() * 8);
        std::vector<uint64_t>().swap(per_read[(size_t)r]);
    });
    return wl;
}

int64_t mm2gb_synth_n_anchors(void *h) { return h ? ((Workload *)h)->off.back() : 0; }
int64_t mm2gb_synth_n_reads(void *h) { return h ? (int64_t)((Workload *)h)->off.size() - 1 : 0; }
void mm2gb_synth_copy(void *h, uint64_t *a, int64_t *off)
{
    Workload *w = (Workload *)h;
    if (!w) return;
    if (a && !w->a.empty()) memcpy(a, w->a.data(), w->a.size() * 8);
    if (off) memcpy(off, w->off.data(), w->off.size() * 8);
}
void mm2gb_synth_free(void *h) { delete (Workload *)h; }

} // extern "C"
