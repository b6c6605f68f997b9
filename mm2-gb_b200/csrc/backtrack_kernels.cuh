// backtrack_kernels.cuh -- device version of the chain extraction + anchor compaction stage.
//
// What is computed is the reference's mg_chain_backtrack + compact_a (lchain.c:27-111), bit for bit, INCLUDING the
// order in which equal-score chain ends are visited: the reference sorts (score, index) pairs with an unstable in-place
// MSD radix sort (ksort.h:98-151: 8-bit American-flag passes from the top byte down, insertion sort for buckets of
// <= 64), so that permutation is reproduced here -- not approximated by a stable sort (SURVEY.md trap T4).  The
// reference (and mm2-gb after its kernels, gpu/plchain.cu:99-150) runs this on one host thread.
//
// How: one warp per read, the read's working set in shared memory (two kernels, see below).
//   * American-flag pass = a deterministic walk over the ORIGINAL array: every bucket is a queue of its original
//     elements; placing an element into bucket d pops the element that sat at d's cursor, which is placed next.  Runs of
//     elements that already sit in their own bucket are finalised (home bucket) or shifted by one (visited bucket) as a
//     whole, found with warp ballots -- so the sequential part is one step per MISPLACED element, and score arrays that
//     are nearly sorted (a chain's scores grow along the read) cost almost nothing.
//   * buckets of <= 64 elements: insertion sort == stable sort, done as a parallel rank sort.
//   * chain walks: the predecessor chase is serial, everything else (claimed test, score drop test of
//     mg_chain_bk_end, prefix maxima) is evaluated for 32 path nodes at a time.
// Three tiers by read length (the host bins the reads, chain_core.cu):
//   * up to 8192 anchors: k_bt_sort<CAP> / k_bt_walk<CAP>, the read's working set in shared memory (32-bit packed keys);
//   * up to 196608 anchors: k_bt_sort_mid / k_bt_walk_mid -- 64-bit keys, f[] and z[] in global scratch, but everything the
//     SERIAL parts touch in shared memory at one byte per anchor: the radix digits of the foreign elements (the pass is
//     reformulated as a walk over them, bt_flag_pass_fq) and the predecessor links as distances; short walks run
//     lane-parallel in batches and are committed in visiting order;
//   * anything else (longer reads; reads a shared-memory kernel hands over through a device-side overflow list: scores
//     >= 2^19, more chains than its key buffer holds): the first tier's code on global-memory scratch with 64-bit keys
//     (k_bt_sort_big / k_bt_walk_big).
// Compacted anchors and chains of the whole batch are PACKED (positions handed out by atomic cursors, recorded per read),
// so that only what was produced is moved to the host (k_drain writes it straight into mapped pinned memory).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "chain_kernels.cuh"

namespace mm2gb {

constexpr int kBtIdxBits = 13;                          // anchors per read handled in shared memory: 2^13
constexpr int kBtMaxAnchors = 1 << kBtIdxBits;
constexpr unsigned kBtIdxMask = kBtMaxAnchors - 1;
constexpr int kBtMaxScore = 1 << (32 - kBtIdxBits);     // packed key = f << 13 | i must fit 32 bits
constexpr int kBtLevels = 8;                            // radix levels of a 64-bit key
constexpr int kBtRow = 260;                             // per level: start[0..256], next-bucket cursor, shift

struct BtParams { int min_cnt, min_sc, max_drop; };

// ---- key traits ---------------------------------------------------------------------------------------------------
struct ZKey {   // (score, index) pair packed in 32 bits, ordered by score only (radix_sort_128x on mm128_t.x = f)
    typedef unsigned T;
    static constexpr bool kPosKey = true;   // positions fit the index field: (score, position) compares as one word
    __device__ static __forceinline__ unsigned poskey(T k, int pos) { return (k & ~kBtIdxMask) | (unsigned)pos; }
    __device__ static __forceinline__ unsigned digit(T k, int shift) { return ((k >> kBtIdxBits) >> shift) & 255u; }
    __device__ static __forceinline__ bool less(T a, T b) { return (a >> kBtIdxBits) < (b >> kBtIdxBits); }
    __device__ static __forceinline__ unsigned long long key64(T k) { return (unsigned long long)(k >> kBtIdxBits); }
    __device__ static __forceinline__ int idx(T k) { return (int)(k & kBtIdxMask); }
    __device__ static __forceinline__ int score(T k) { return (int)(k >> kBtIdxBits); }
    __device__ static __forceinline__ T make(int f, int i) { return ((unsigned)f << kBtIdxBits) | (unsigned)i; }
};
struct ZKey64 { // the same pair in 64 bits (any read length, any score): score << 32 | index
    typedef unsigned long long T;
    static constexpr bool kPosKey = false;
    __device__ static __forceinline__ unsigned poskey(T, int) { return 0; }
    __device__ static __forceinline__ unsigned digit(T k, int shift) { return ((unsigned)(k >> 32) >> shift) & 255u; }
    __device__ static __forceinline__ bool less(T a, T b) { return (unsigned)(a >> 32) < (unsigned)(b >> 32); }
    __device__ static __forceinline__ unsigned long long key64(T k) { return k >> 32; }
    __device__ static __forceinline__ int idx(T k) { return (int)(unsigned)k; }
    __device__ static __forceinline__ int score(T k) { return (int)(unsigned)(k >> 32); }
    __device__ static __forceinline__ T make(int f, int i) { return ((unsigned long long)(unsigned)f << 32) | (unsigned)i; }
};
struct WKey {   // chain start position x (64 bit); the chain id travels in a separate payload array
    typedef unsigned long long T;
    static constexpr bool kPosKey = false;
    __device__ static __forceinline__ unsigned poskey(T, int) { return 0; }
    __device__ static __forceinline__ unsigned digit(T k, int shift) { return (unsigned)(k >> shift) & 255u; }
    __device__ static __forceinline__ bool less(T a, T b) { return a < b; }
    __device__ static __forceinline__ unsigned long long key64(T k) { return k; }
};

// scratch shared by the sorts of one warp; POS = type of a position inside the sorted array
template <class POS>
struct BtSortScratch {
    unsigned *cnt;      // [256] histogram, then the bucket cursors of the running pass
    POS *start;         // [levels][kBtRow]
};

// rank sort (== the reference's insertion sort, ksort.h:105-115: stable) of every bucket of 2..64 elements of one pass;
// start = bucket boundaries of that pass (257 entries), or nullptr to sort [lo, hi) as ONE bucket.
template <class KO, bool PAY, class PAYT, class POS>
__device__ void bt_rank_sort(typename KO::T *A, PAYT *pay, typename KO::T *tmpA, PAYT *tmpPay, int lo, int hi,
                             int shift, const POS *start, int lane, unsigned *dirty = nullptr)
{
    typedef typename KO::T K;
    // Most buckets of real score arrays are sorted already (scores grow along a chain): with `dirty` ([256] words, free after
    // the pass) a first sweep marks the buckets that hold an element smaller than its left neighbour, and only those are
    // ranked and moved -- nothing at all if there is none.
    if (start && dirty) {
        for (int d = lane; d < 256; d += 32) dirty[d] = 0;
        __syncwarp();
        bool any = false;
        for (int e0 = lo; e0 < hi; e0 += 32) {
            const int e = e0 + lane;
            if (e < hi) {
                const K key = A[e];
                const unsigned d = KO::digit(key, shift);
                if (e > (int)start[d] && KO::less(key, A[e - 1])) { dirty[d] = 1; any = true; }
            }
        }
        if (!__any_sync(0xffffffffu, any)) return;
        __syncwarp();
    }
    for (int e0 = lo; e0 < hi; e0 += 32) {
        const int e = e0 + lane;
        if (e < hi) {
            const K key = A[e];
            int bs = lo, be = hi;
            bool todo = true;
            if (start) {
                const unsigned d = KO::digit(key, shift);
                bs = (int)start[d];
                be = (int)start[d + 1];
                if (dirty) todo = dirty[d] != 0;
            }
            const int m = be - bs;
            if (todo && m >= 2 && m <= 64) {
                int r = 0;
                if constexpr (KO::kPosKey) { // stable order = order of (score, position): one compare per bucket mate
                    const unsigned ce = KO::poskey(key, e);
                    for (int j = bs; j < be; ++j) r += KO::poskey(A[j], j) < ce ? 1 : 0;
                } else {
                    for (int j = bs; j < be; ++j) {
                        const K kj = A[j];
                        r += (KO::less(kj, key) || (!KO::less(key, kj) && j < e)) ? 1 : 0;
                    }
                }
                tmpA[bs + r] = key;
                if (PAY) tmpPay[bs + r] = pay[e];
            } else {
                tmpA[e] = key;
                if (PAY) tmpPay[e] = pay[e];
            }
        }
    }
    __syncwarp();
    for (int e = lo + lane; e < hi; e += 32) {
        A[e] = tmpA[e];
        if (PAY) pay[e] = tmpPay[e];
    }
    __syncwarp();
}

// One American-flag pass over A[lo, hi) on digit `shift` (ksort.h:116-139), bucket boundaries -> st[0..256].
template <class KO, bool PAY, class PAYT, class POS>
__device__ void bt_flag_pass(typename KO::T *A, PAYT *pay, int lo, int hi, int shift, unsigned *cnt, POS *st, int lane)
{
    typedef typename KO::T K;
    const unsigned full = 0xffffffffu;
    for (int d = lane; d < 256; d += 32) cnt[d] = 0;
    __syncwarp();
    for (int e = lo + lane; e < hi; e += 32) atomicAdd(&cnt[KO::digit(A[e], shift)], 1u);
    __syncwarp();
    {   // exclusive scan of 256 counts: 8 per lane
        unsigned c[8], sum = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { c[q] = cnt[lane * 8 + q]; sum += c[q]; }
        unsigned incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned y = __shfl_up_sync(full, incl, d);
            if (lane >= d) incl += y;
        }
        unsigned run = (unsigned)lo + incl - sum;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            st[lane * 8 + q] = (POS)run;
            cnt[lane * 8 + q] = run;        // cursor of the bucket
            run += c[q];
        }
        if (lane == 31) st[256] = (POS)hi;
    }
    __syncwarp();
    // buckets in ascending order; everything below is warp-uniform (all lanes read the same values)
    for (int k = 0; k < 256; ++k) {
        const int endk = (int)st[k + 1];
        int c = (int)cnt[k];
        while (c < endk) {
            // elements that already sit in their home bucket stay where they are
            const int e = c + lane;
            const bool mis = e < endk && KO::digit(A[e], shift) != (unsigned)k;
            const unsigned mm = __ballot_sync(full, mis);
            if (!mm) { c = min(c + 32, endk); continue; }
            c += __ffs(mm) - 1;
            K carried = A[c];
            PAYT cpay = PAY ? pay[c] : (PAYT)0;
            unsigned d = KO::digit(carried, shift);
            do {
                const int pos = (int)cnt[d], endd = (int)st[d + 1];
                {   // common case in shuffled data: the element at the cursor is foreign -- a plain swap, nothing to shift
                    // (every lane reads the same words: no ballot, two dependent loads per step)
                    const K front = A[pos];
                    const unsigned df = KO::digit(front, shift);
                    if (df != d) {
                        const PAYT fpay = PAY ? pay[pos] : (PAYT)0;
                        __syncwarp();
                        if (lane == 0) { A[pos] = carried; if (PAY) pay[pos] = cpay; cnt[d] = (unsigned)(pos + 1); }
                        __syncwarp();
                        carried = front;
                        cpay = fpay;
                        d = df;
                        continue;
                    }
                }
                // elements of bucket d sitting at its cursor are pushed one slot to the right (each is evicted by the
                // arriving element and re-placed at the next slot); the first foreign element after them is evicted for good
                int L = 0;
                for (;;) {
                    const int q = pos + L + lane;
                    const bool own = q < endd && KO::digit(A[q], shift) == d;
                    const unsigned nm = __ballot_sync(full, !own);
                    if (!nm) { L += 32; continue; }
                    L += __ffs(nm) - 1;
                    break;
                }
                const K evicted = A[pos + L];
                const PAYT epay = PAY ? pay[pos + L] : (PAYT)0;
                for (int top = L; top > 0; top -= 32) { // shift A[pos .. pos+L) up by one, highest chunk first
                    const int base = max(0, top - 32), idx = base + lane;
                    K val = 0;
                    PAYT pv = 0;
                    if (idx < top) { val = A[pos + idx]; if (PAY) pv = pay[pos + idx]; }
                    __syncwarp();
                    if (idx < top) { A[pos + idx + 1] = val; if (PAY) pay[pos + idx + 1] = pv; }
                    __syncwarp();
                }
                if (lane == 0) {
                    A[pos] = carried;
                    if (PAY) pay[pos] = cpay;
                    cnt[d] = (unsigned)(pos + L + 1);
                }
                __syncwarp();
                carried = evicted;
                cpay = epay;
                d = KO::digit(carried, shift);
            } while (d != (unsigned)k);
            if (lane == 0) { A[c] = carried; if (PAY) pay[c] = cpay; }
            __syncwarp();
            ++c;
        }
    }
    __syncwarp();
}

// radix_sort_128x (ksort.h:146-150) of A[0, n) by KO's key, payload in tandem.  tmpA/tmpPay: scratch of the same size.
template <class KO, bool PAY, class PAYT, class POS>
__device__ void bt_sort(typename KO::T *A, PAYT *pay, typename KO::T *tmpA, PAYT *tmpPay, int n, BtSortScratch<POS> sc, int lane)
{
    const unsigned full = 0xffffffffu;
    if (n <= 1) return;
    if (n <= 64) { bt_rank_sort<KO, PAY, PAYT, POS>(A, pay, tmpA, tmpPay, 0, n, 0, nullptr, lane); return; }
    // Passes in which every key has the same digit are the identity (one bucket, nothing moves, recursion continues on the
    // whole range): start at the highest digit in which the keys differ (same shortcut as backtrack.cpp).
    unsigned long long o = 0, an = ~0ULL;
    for (int e = lane; e < n; e += 32) { const unsigned long long k = KO::key64(A[e]); o |= k; an &= k; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { o |= __shfl_xor_sync(full, o, d); an &= __shfl_xor_sync(full, an, d); }
    const unsigned long long diff = o ^ an;
    if (!diff) return;
    int shift = 56;
    while (((diff >> shift) & 255ULL) == 0) shift -= 8;
    // depth-first over buckets of more than 64 elements; the explicit stack lives in the per-level rows of `start`
    int lv = 0;
    POS *row = sc.start;
    bt_flag_pass<KO, PAY, PAYT, POS>(A, pay, 0, n, shift, sc.cnt, row, lane);
    if (shift) bt_rank_sort<KO, PAY, PAYT, POS>(A, pay, tmpA, tmpPay, 0, n, shift, row, lane, sc.cnt);
    if (lane == 0) { row[257] = 0; row[258] = (POS)shift; }
    __syncwarp();
    while (lv >= 0) {
        row = sc.start + lv * kBtRow;
        const int sh = (int)row[258];
        int k = (int)row[257];
        if (sh == 0 || k >= 256) { --lv; continue; }
        // next bucket of this level with more than 64 elements
        int found = -1;
        while (k < 256) {
            const int kk = k + lane;
            const bool big = kk < 256 && (int)row[kk + 1] - (int)row[kk] > 64;
            const unsigned bm = __ballot_sync(full, big);
            if (bm) { found = k + __ffs(bm) - 1; break; }
            k += 32;
        }
        __syncwarp();
        if (found < 0) { if (lane == 0) row[257] = 256; __syncwarp(); --lv; continue; }
        if (lane == 0) row[257] = (POS)(found + 1);
        const int blo = (int)row[found], bhi = (int)row[found + 1];
        const int nsh = sh > 8 ? sh - 8 : 0;
        ++lv;
        POS *crow = sc.start + lv * kBtRow;
        __syncwarp();
        bt_flag_pass<KO, PAY, PAYT, POS>(A, pay, blo, bhi, nsh, sc.cnt, crow, lane);
        if (nsh) bt_rank_sort<KO, PAY, PAYT, POS>(A, pay, tmpA, tmpPay, blo, bhi, nsh, crow, lane, sc.cnt);
        if (lane == 0) { crow[257] = 0; crow[258] = (POS)nsh; }
        __syncwarp();
    }
}

// ---- the same sort for reads too long for shared-memory keys ------------------------------------------------------------
// One American-flag pass over A[lo, hi) in GLOBAL memory, reformulated so that its serial part touches two shared-memory
// words per step and nothing in the pass costs more than O(1) per element.
//
// The reference's pass (ksort.h:116-139) is a deterministic walk: every bucket's region is consumed front to back by a
// cursor, positions at or behind a cursor still hold their original element, and an element that sits in its own region
// ("own") never changes the walk -- an arriving element is dropped at the cursor, the own elements behind it move up by one
// and the first FOREIGN element after them is carried on.  So:
//   * control flow depends on the foreign elements only.  They are compacted, in position order, into tokens
//     (digit in shared memory, position in global memory); region r owns tokens [fst[r], fen[r]).
//   * the walk: "carried element with digit d arrives in region d, evicts that region's next foreign token" is
//     g2 = cur[d]++; next digit = D[g2]  -- two dependent shared-memory loads.  During region k's own turn (the outer
//     loop) its remaining foreign tokens start cycles and are replaced IN PLACE by the element that closes the cycle.
//   * where everything lands follows in parallel afterwards: the j-th arrival in region d sits at the region start (j = 0) or
//     one behind the (j-1)-th evicted token; an own element moves up by one iff it lies before the last token evicted by an
//     arrival (ecut[d]); a cycle-closing element takes the position of the token that opened the cycle.
// The O(run length) shifting of the literal algorithm (the hot spot of the first version of this kernel) is gone
// (tests/test_flagpass_model.py checks the reformulation against the literal pass).
// A CTA of NT threads works on one read: the parallel phases (digits + histogram, placement, copy back, the rank sort
// of small buckets) use every thread, the token compaction and the walk run on warp 0.  Every thread takes the same path
// (all control values come from shared memory behind a barrier).
// Scratch (global, one unsigned per element each): tok (position -> token or ~0 for own), fpos (token -> position),
// nxt (token -> token it evicted on arrival, or 2^31 | opening token).  Bucket boundaries -> st[0..256] (absolute positions).
constexpr int kBtMidThreads = 512;    // largest CTA of k_bt_sort_mid<NT> (NT = 128 / 256 / 512 by size class, see bt_mid_threads)
// the fewer reads of a class fit an SM, the more threads each gets for the parallel phases (registers: 40 x NT x resident reads)
__host__ __device__ inline int bt_mid_threads(int cap) { return cap <= 20480 ? 128 : cap <= 49152 ? 256 : 512; }

struct BtFqScratch {
    unsigned *cnt, *rows, *fst, *fen;   // shared: [256], [4 * kBtRow], [256], [256]
    int *ecut;                          // shared: [256]
    unsigned *misc;                     // shared: [8] broadcast words
    unsigned *tok, *fpos, *nxt;         // global, per element
};

template <class KO, int NT>
__device__ void bt_flag_pass_fq(typename KO::T *A, typename KO::T *tmpA, const BtFqScratch &q, int lo, int hi, int shift, unsigned *st,
                                unsigned char *D, int tid)
{
    typedef typename KO::T K;
    const int lane = tid & 31;
    const bool w0 = tid < 32;
    const unsigned full = 0xffffffffu, lt = (1u << lane) - 1u;
    const int m = hi - lo;
    K *Al = A + lo, *Tl = tmpA + lo;
    unsigned *tk = q.tok + lo, *fp = q.fpos + lo, *nx = q.nxt + lo;
    unsigned *cnt = q.cnt, *fst = q.fst, *fen = q.fen;
    int *ecut = q.ecut;
    for (int d = tid; d < 256; d += NT) cnt[d] = 0;
    __syncthreads();
    for (int e0 = 0; e0 < m; e0 += 4 * NT) { // 4 coalesced key loads per thread in flight
        K kv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { const int e = e0 + t * NT + tid; kv[t] = e < m ? Al[e] : (K)0; }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int e = e0 + t * NT + tid;
            if (e < m) { const unsigned d = KO::digit(kv[t], shift); D[e] = (unsigned char)d; atomicAdd(&cnt[d], 1u); }
        }
    }
    __syncthreads();
    if (w0) {
        {   // exclusive scan of 256 counts: 8 per lane (positions relative to lo)
            unsigned c[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { c[j] = cnt[lane * 8 + j]; sum += c[j]; }
            unsigned incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned y = __shfl_up_sync(full, incl, d);
                if (lane >= d) incl += y;
            }
            unsigned run = incl - sum;
#pragma unroll
            for (int j = 0; j < 8; ++j) { st[lane * 8 + j] = run; run += c[j]; }
            if (lane == 31) st[256] = (unsigned)m;
        }
        __syncwarp();
        // foreign elements -> tokens, region by region (= ascending positions); digits compacted in place (token <= position)
        int g = 0;
        for (int rb = 0; rb < 256; rb += 32) {
            const unsigned s0l = st[rb + lane], s1l = st[rb + lane + 1];
            unsigned mask = __ballot_sync(full, s1l > s0l);
            while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                const int r = rb + j;
                const int s0 = (int)__shfl_sync(full, s0l, j), s1 = (int)__shfl_sync(full, s1l, j);
                if (lane == 0) { fst[r] = (unsigned)g; cnt[r] = (unsigned)g; }   // cnt[r]: the region's next foreign token
                for (int c = s0; c < s1; c += 32) {
                    const int e = c + lane;
                    const unsigned d = e < s1 ? (unsigned)D[e] : (unsigned)r;
                    const bool fo = d != (unsigned)r;
                    const unsigned fm = __ballot_sync(full, fo);
                    const int t = g + __popc(fm & lt);
                    if (fo) { D[t] = (unsigned char)d; fp[t] = (unsigned)e; }
                    if (e < s1) tk[e] = fo ? (unsigned)t : 0xffffffffu;
                    g += __popc(fm);
                }
                if (lane == 0) fen[r] = (unsigned)g;
            }
        }
        __syncwarp();
        if (g > 0) {
            // the walk, by lane 0 alone.  A region's cursor and the digit of the token under it travel in ONE word,
            // cnt[d] = cursor | digit << 24 (cursors stay below 2^24: reads of at most 196608 anchors), so a step of the
            // dependent chain is a single shared-memory load: the digit of the evicted token comes with the cursor, and the
            // word for the region's next visit (cursor + 1 and the digit there) is prepared off the critical path.
            for (int r = lane; r < 256; r += 32) {
                const unsigned c0 = cnt[r];                     // = fst[r] (0 for an empty region): c0 <= g <= m, D[m] is slack
                cnt[r] = c0 | ((unsigned)D[c0] << 24);
            }
            __syncwarp();
            for (int rb = 0; rb < 256; rb += 32) {
                const unsigned s0l = st[rb + lane], s1l = st[rb + lane + 1];
                unsigned mask = __ballot_sync(full, s1l > s0l);
                while (mask) {
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const unsigned k = (unsigned)(rb + j);
                    if (lane == 0) {
                        const unsigned fe = fen[k];
                        unsigned h = cnt[k] & 0xffffffu;
                        ecut[k] = (int)h;   // tokens below h were evicted by arrivals; turned into a position below
                        while (h < fe) {
                            const unsigned c0 = h++;
                            unsigned carried = c0;
                            unsigned d = D[c0];
                            while (d != k) {
                                const unsigned w = cnt[d];
                                const unsigned g2 = w & 0xffffffu;
                                nx[carried] = g2;
                                cnt[d] = (g2 + 1u) | ((unsigned)D[g2 + 1u] << 24);
                                carried = g2;
                                d = w >> 24;
                            }
                            nx[carried] = 0x80000000u | c0;
                        }
                    }
                }
            }
            __syncwarp();
            for (int r = lane; r < 256; r += 32) {
                int ec = -1;
                if (st[r + 1] > st[r]) { const unsigned h = (unsigned)ecut[r]; if (h > fst[r]) ec = (int)fp[h - 1]; }
                ecut[r] = ec;
            }
        }
        if (lane == 0) q.misc[0] = (unsigned)g;
    }
    __syncthreads();
    if (q.misc[0] > 0) {
        for (int e0 = 0; e0 < m; e0 += 4 * NT) { // placement: up to three dependent loads per element, 4 elements per thread in flight
            K kv[4];
            unsigned tv[4], v[4], dst[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) { const int e = e0 + t * NT + tid; kv[t] = e < m ? Al[e] : (K)0; tv[t] = e < m ? tk[e] : 0xffffffffu; }
#pragma unroll
            for (int t = 0; t < 4; ++t) v[t] = tv[t] != 0xffffffffu ? nx[tv[t]] : 0u;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int e = e0 + t * NT + tid;
                const unsigned d = KO::digit(kv[t], shift);
                if (tv[t] == 0xffffffffu) dst[t] = (unsigned)e + (e < ecut[d] ? 1u : 0u);
                else if (v[t] & 0x80000000u) dst[t] = fp[v[t] & 0x7fffffffu];
                else dst[t] = v[t] == fst[d] ? st[d] : fp[v[t] - 1] + 1u;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) { const int e = e0 + t * NT + tid; if (e < m) Tl[dst[t]] = kv[t]; }
        }
        __syncthreads();
        for (int e0 = 0; e0 < m; e0 += 4 * NT) {
            K kv[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) { const int e = e0 + t * NT + tid; kv[t] = e < m ? Tl[e] : (K)0; }
#pragma unroll
            for (int t = 0; t < 4; ++t) { const int e = e0 + t * NT + tid; if (e < m) Al[e] = kv[t]; }
        }
    }
    __syncthreads();
    for (int d = tid; d <= 256; d += NT) st[d] += (unsigned)lo;
    __syncthreads();
}

// rank sort of every bucket of 2..64 elements of one pass over A[lo, hi) in global memory, by the whole CTA (bt_rank_sort is
// the one-warp version): rank = number of bucket mates that order before the element as (score, position)
template <class KO, int NT>
__device__ void bt_rank_sort_blk(typename KO::T *A, typename KO::T *tmpA, int lo, int hi, int shift, const unsigned *start, int tid,
                                 unsigned *dirty, unsigned *any_dirty)
{
    typedef typename KO::T K;
    // only buckets that hold an element smaller than its left neighbour need ranking (see bt_rank_sort)
    for (int d = tid; d < 256; d += NT) dirty[d] = 0;
    if (tid == 0) *any_dirty = 0;
    __syncthreads();
    for (int e = lo + tid; e < hi; e += NT) {
        const K key = A[e];
        const unsigned d = KO::digit(key, shift);
        if (e > (int)start[d] && KO::less(key, A[e - 1])) { dirty[d] = 1; *any_dirty = 1; }
    }
    __syncthreads();
    if (*any_dirty == 0) return;
    for (int e = lo + tid; e < hi; e += NT) {
        const K key = A[e];
        const unsigned d = KO::digit(key, shift);
        if (!dirty[d]) continue;
        const int bs = (int)start[d], be = (int)start[d + 1];
        const int m = be - bs;
        int r = e - bs;
        if (m >= 2 && m <= 64) {
            r = 0;
            for (int j = bs; j < be; ++j) {
                const K kj = A[j];
                r += (KO::less(kj, key) || (!KO::less(key, kj) && j < e)) ? 1 : 0;
            }
        }
        tmpA[bs + r] = key;
    }
    __syncthreads();
    for (int e = lo + tid; e < hi; e += NT) {
        const K key = A[e];
        if (dirty[KO::digit(key, shift)]) A[e] = tmpA[e];
    }
    __syncthreads();
}

// radix_sort_128x of A[0, n) (global memory) for reads of up to `cap` anchors, by a CTA of NT threads: every pass is
// bt_flag_pass_fq; only buckets of at most kcap (<= cap / 8) elements are copied to shared memory (KA, which aliases D) and
// finished there by warp 0 with bt_sort -- for a few hundred elements the literal algorithm beats the fixed cost of a pass
// through global memory.  rows: 4 levels (32-bit scores) of kBtRow entries.
template <class KO, int NT>
__device__ void bt_sort_mid(typename KO::T *A, typename KO::T *tmpA, int n, const BtFqScratch &q, unsigned char *D, int kcap, int tid)
{
    typedef typename KO::T K;
    const int lane = tid & 31;
    const bool w0 = tid < 32;
    const unsigned full = 0xffffffffu;
    K *KA = reinterpret_cast<K *>(D);
    if (n <= 1) return;
    BtSortScratch<unsigned> sc;
    sc.cnt = q.cnt;
    sc.start = q.rows;
    if (n <= kcap) {
        if (w0) {
            for (int e = lane; e < n; e += 32) KA[e] = A[e];
            __syncwarp();
            bt_sort<KO, false, unsigned, unsigned>(KA, nullptr, tmpA, nullptr, n, sc, lane);
            __syncwarp();
            for (int e = lane; e < n; e += 32) A[e] = KA[e];
        }
        __syncthreads();
        return;
    }
    // highest digit in which the keys differ (32-bit scores: two words of OR / AND, reduced through shared memory)
    if (tid == 0) { q.misc[2] = 0u; q.misc[3] = 0xffffffffu; }
    __syncthreads();
    {
        unsigned o = 0, an = 0xffffffffu;
        for (int e = tid; e < n; e += NT) { const unsigned k = (unsigned)KO::key64(A[e]); o |= k; an &= k; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { o |= __shfl_xor_sync(full, o, d); an &= __shfl_xor_sync(full, an, d); }
        if (lane == 0) { atomicOr(&q.misc[2], o); atomicAnd(&q.misc[3], an); }
    }
    __syncthreads();
    const unsigned diff = q.misc[2] ^ q.misc[3];
    if (!diff) return;
    int shift = 24;
    while (((diff >> shift) & 255u) == 0) shift -= 8;
    int lv = 0;
    unsigned *row = q.rows;
    bt_flag_pass_fq<KO, NT>(A, tmpA, q, 0, n, shift, row, D, tid);
    if (shift) bt_rank_sort_blk<KO, NT>(A, tmpA, 0, n, shift, row, tid, q.cnt, q.misc + 4);
    if (tid == 0) { row[257] = 0; row[258] = (unsigned)shift; }
    __syncthreads();
    while (lv >= 0) {
        row = q.rows + lv * kBtRow;
        const int sh = (int)row[258];
        int k = (int)row[257];
        if (sh == 0 || k >= 256) { --lv; continue; }
        int found = -1;
        while (k < 256) { // every warp finds the same bucket
            const int kk = k + lane;
            const bool big = kk < 256 && (int)row[kk + 1] - (int)row[kk] > 64;
            const unsigned bm = __ballot_sync(full, big);
            if (bm) { found = k + __ffs(bm) - 1; break; }
            k += 32;
        }
        int blo = 0, bhi = 0;
        if (found >= 0) { blo = (int)row[found]; bhi = (int)row[found + 1]; }
        __syncthreads();   // everyone has read this level's cursor before it moves
        if (tid == 0) row[257] = (unsigned)(found < 0 ? 256 : found + 1);
        if (found < 0) { __syncthreads(); --lv; continue; }
        const int nsh = sh > 8 ? sh - 8 : 0;
        if (bhi - blo <= kcap) { // the whole subtree in shared memory (identity passes are skipped by bt_sort itself)
            if (w0) {
                const int m = bhi - blo;
                for (int e = lane; e < m; e += 32) KA[e] = A[blo + e];
                __syncwarp();
                sc.start = q.rows + (lv + 1) * kBtRow;
                bt_sort<KO, false, unsigned, unsigned>(KA, nullptr, tmpA + blo, nullptr, m, sc, lane);
                __syncwarp();
                for (int e = lane; e < m; e += 32) A[blo + e] = KA[e];
            }
            __syncthreads();
            continue;
        }
        ++lv;
        unsigned *crow = q.rows + lv * kBtRow;
        bt_flag_pass_fq<KO, NT>(A, tmpA, q, blo, bhi, nsh, crow, D, tid);
        if (nsh) bt_rank_sort_blk<KO, NT>(A, tmpA, blo, bhi, nsh, crow, tid, q.cnt, q.misc + 4);
        if (tid == 0) { crow[257] = 0; crow[258] = (unsigned)nsh; }
        __syncthreads();
    }
}

// a read the shared-memory kernels cannot finish goes to the global-memory ones through this list
__device__ __forceinline__ void bt_overflow(int r, int *ovf_list, Counters *ctr)
{
    ovf_list[atomicAdd(&ctr->ovf_cnt, 1)] = r;     // the list holds one entry per read of the batch: it cannot overflow
}

// ---------------------------------------------------------------------------------------------------------------------
// The stage is two kernels per size class, one warp (= one CTA) per read each, because the two halves want different
// amounts of shared memory: the sort needs 8 bytes per anchor for a few microseconds, the walk only ~3.5 bytes per anchor
// for much longer (it is a serial pointer chase) -- so three times as many walks as sorts fit on an SM.
//   k_bt_sort : z[] = anchors scoring >= min_sc, sorted as the reference sorts them  -> zs_scr (global), nz_out
//   k_bt_walk : chain extraction in that order + compaction                          -> n_u, n_b, b_pack, u_pack
// ---------------------------------------------------------------------------------------------------------------------
template <int CAP>
struct BtSortSmem {
    unsigned zk[CAP];               // (score << 13 | index), sorted in place
    unsigned cnt[256];
    unsigned short start[3 * kBtRow];   // scores are below 2^19: at most three radix levels
};
// The rank-sort scratch (the second key array) is NOT here: it is the read's slice of zs_scr, where the sorted keys go in
// the end anyway.  The rank sort is the parallel part of the sort, so the global round trip costs little, and 4 instead of
// 8 bytes per anchor doubles the reads resident on an SM -- these one-warp kernels are latency bound.

// z[]: anchors scoring >= min_sc, in index order (lchain.c:33-40); 4 coalesced loads per lane in flight.  Returns nz; the
// largest kept score in fmax.
template <class ZK>
__device__ __forceinline__ int bt_collect(const int *__restrict__ fr, int n, int min_sc, typename ZK::T *zk, int lane, int &fmax)
{
    const unsigned full = 0xffffffffu;
    int nz = 0;
    fmax = 0;
    for (int i0 = 0; i0 < n; i0 += 256) {   // 8 coalesced loads per lane in flight: the loop is bound by their latency
        int fv[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) { const int i = i0 + t * 32 + lane; fv[t] = i < n ? fr[i] : INT32_MIN; }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int i = i0 + t * 32 + lane;
            const bool keep = fv[t] >= min_sc && i < n;
            const unsigned m = __ballot_sync(full, keep);
            if (keep) zk[nz + __popc(m & ((1u << lane) - 1u))] = ZK::make(fv[t], i);
            nz += __popc(m);
            fmax = max(fmax, keep ? fv[t] : 0);
        }
    }
    fmax = __reduce_max_sync(full, fmax);
    return nz;
}

// the same by every warp of a CTA: warp w takes the w-th slice of the read, counts, and writes behind the slices before it
// (wcnt: one shared word per warp).  Same output as bt_collect.
template <class ZK, int NT>
__device__ __forceinline__ int bt_collect_blk(const int *__restrict__ fr, int n, int min_sc, typename ZK::T *zk, int tid, unsigned *wcnt)
{
    constexpr int NW = NT / 32;
    const unsigned full = 0xffffffffu;
    const int lane = tid & 31, w = tid >> 5;
    const int per = ((n + NW - 1) / NW + 255) & ~255;   // slice length: whole 256-anchor rounds
    const int s0 = min(n, w * per), s1 = min(n, s0 + per);
    int cntw = 0;
    for (int i0 = s0; i0 < s1; i0 += 256) {
        int fv[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) { const int i = i0 + t * 32 + lane; fv[t] = i < s1 ? fr[i] : INT32_MIN; }
#pragma unroll
        for (int t = 0; t < 8; ++t) cntw += (fv[t] >= min_sc && i0 + t * 32 + lane < s1) ? 1 : 0;
    }
    cntw = __reduce_add_sync(full, cntw);
    if (lane == 0) wcnt[w] = (unsigned)cntw;
    __syncthreads();
    int nz = 0, total = 0;
    for (int k = 0; k < NW; ++k) { const int c = (int)wcnt[k]; if (k < w) nz += c; total += c; }
    for (int i0 = s0; i0 < s1; i0 += 256) {
        int fv[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) { const int i = i0 + t * 32 + lane; fv[t] = i < s1 ? fr[i] : INT32_MIN; }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int i = i0 + t * 32 + lane;
            const bool keep = fv[t] >= min_sc && i < s1;
            const unsigned m = __ballot_sync(full, keep);
            if (keep) zk[nz + __popc(m & ((1u << lane) - 1u))] = ZK::make(fv[t], i);
            nz += __popc(m);
        }
    }
    __syncthreads();
    return total;
}

template <int CAP>
__global__ void __launch_bounds__(32)
k_bt_sort(const int *__restrict__ f, const long long *__restrict__ off, const int *__restrict__ read_list, int n_list, BtParams bp,
          unsigned *__restrict__ zs_scr, int *__restrict__ nz_out, int *__restrict__ ovf_list, Counters *__restrict__ ctr)
{
    extern __shared__ int4 bt_raw[];
    BtSortSmem<CAP> &S = *reinterpret_cast<BtSortSmem<CAP> *>(bt_raw);
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= n_list) return;
    const int r = read_list[blockIdx.x];
    const long long o0 = off[r];
    const int n = (int)(off[r + 1] - o0);
    const int *fr = f + o0;
    if (n > CAP || bp.min_sc < 0) { if (lane == 0) nz_out[r] = -1; return; }
    int fmax;
    const int nz = bt_collect<ZKey>(fr, n, bp.min_sc, S.zk, lane, fmax);
    if (fmax >= kBtMaxScore) { // does not pack into 32 bits: the 64-bit kernels take the read
        if (lane == 0) { nz_out[r] = -1; bt_overflow(r, ovf_list, ctr); }
        return;
    }
    __syncwarp();
    BtSortScratch<unsigned short> sc;
    sc.cnt = S.cnt;
    sc.start = S.start;
    unsigned *zo = zs_scr + o0;
    bt_sort<ZKey, false, unsigned short, unsigned short>(S.zk, nullptr, zo, nullptr, nz, sc, lane);
    __syncwarp();
    for (int e = lane; e < nz; e += 32) zo[e] = S.zk[e];
    if (lane == 0) nz_out[r] = nz;
}

// the read of CTA b: from the host's list, or (ovf != nullptr) from the device-side overflow list
__device__ __forceinline__ int bt_big_read(const int *__restrict__ read_list, int n_list, const int *ovf_list, const Counters *ctr)
{
    const int b = (int)blockIdx.x;
    if (ovf_list) return b < ctr->ovf_cnt ? ovf_list[b] : -1;
    return b < n_list ? read_list[b] : -1;
}

// The same for reads of any size / score range: keys are 64 bit and live in global scratch (zk, zk2: one u64 per anchor each).
__global__ void __launch_bounds__(32)
k_bt_sort_big(const int *__restrict__ f, const long long *__restrict__ off, const int *__restrict__ read_list, int n_list,
              const int *ovf_list, const Counters *ctr, BtParams bp, unsigned long long *zk_scr, unsigned long long *zk2_scr,
              int *__restrict__ nz_out)
{
    __shared__ unsigned s_cnt[256];
    __shared__ unsigned s_start[4 * kBtRow];    // scores are below 2^31: at most four radix levels
    const int lane = threadIdx.x;
    const int r = bt_big_read(read_list, n_list, ovf_list, ctr);
    if (r < 0) return;
    const long long o0 = off[r];
    const int n = (int)(off[r + 1] - o0);
    if (bp.min_sc < 0) { if (lane == 0) nz_out[r] = -1; return; }
    unsigned long long *zk = zk_scr + o0, *zk2 = zk2_scr + o0;
    int fmax;
    const int nz = bt_collect<ZKey64>(f + o0, n, bp.min_sc, zk, lane, fmax);
    __syncwarp();
    BtSortScratch<unsigned> sc;
    sc.cnt = s_cnt;
    sc.start = s_start;
    bt_sort<ZKey64, false, unsigned, unsigned>(zk, nullptr, zk2, nullptr, nz, sc, lane);
    if (lane == 0) nz_out[r] = nz;
}

// Reads of 8193 .. 196608 anchors ("mid" classes): 64-bit keys in global scratch as in k_bt_sort_big, but the serial part of
// every pass runs on one-byte tokens in shared memory (bt_flag_pass_fq).  Dynamic shared memory: cap bytes (digits / tokens;
// keys of buckets of <= 512 elements).
template <int NT>
__global__ void __launch_bounds__(NT)
k_bt_sort_mid(const int *__restrict__ f, const long long *__restrict__ off, const int *__restrict__ read_list, int n_list, BtParams bp,
              unsigned long long *zk_scr, unsigned long long *zk2_scr, unsigned *tok_scr, unsigned *fpos_scr, unsigned *nxt_scr,
              int *__restrict__ nz_out, int cap)
{
    extern __shared__ int4 bt_raw[];
    __shared__ unsigned s_cnt[256];
    __shared__ unsigned s_rows[4 * kBtRow];     // scores are below 2^31: at most four radix levels
    __shared__ unsigned s_fst[256], s_fen[256], s_misc[8];
    __shared__ int s_ecut[256];
    const int tid = threadIdx.x;
    if ((int)blockIdx.x >= n_list) return;
    const int r = read_list[blockIdx.x];
    const long long o0 = off[r];
    const int n = (int)(off[r + 1] - o0);
    if (bp.min_sc < 0 || n > cap) { if (tid == 0) nz_out[r] = -1; return; }
    unsigned long long *zk = zk_scr + o0, *zk2 = zk2_scr + o0;
    const int nz = bt_collect_blk<ZKey64, NT>(f + o0, n, bp.min_sc, zk, tid, s_cnt);
    BtFqScratch q;
    q.cnt = s_cnt; q.rows = s_rows; q.fst = s_fst; q.fen = s_fen; q.ecut = s_ecut; q.misc = s_misc;
    q.tok = tok_scr + o0; q.fpos = fpos_scr + o0; q.nxt = nxt_scr + o0;
    bt_sort_mid<ZKey64, NT>(zk, zk2, nz, q, reinterpret_cast<unsigned char *>(bt_raw), min(cap / 8, 512), tid);
    if (tid == 0) nz_out[r] = nz;
}

template <int CAP>
struct BtWalkSmem {
    static constexpr int WC = CAP / 16 < 64 ? 64 : CAP / 16;   // chains whose start keys can be sorted here
    // Two phases, one footprint: the links and claimed bits of the chain extraction are dead when the compaction sorts the chain
    // starts, so its scratch lies over them.  (Side by side they were 26.8 KB for the 6144 class = 8 reads per SM; 13.9 KB = 15
    // reads, and these one-warp kernels are as fast as the number of reads resident per SM.)
    struct Walk {
        // f[] is read from global memory (one 32-wide gather per chase batch): keeping it here would double the footprint and
        // halve the reads resident on an SM, and this one-warp kernel is bound by latency, not by bandwidth
        unsigned short ps[CAP + 2];             // p[] by anchor index; "none" is the sentinel index CAP, whose own entry is CAP
        unsigned short path[32];                // the nodes of one chase batch
        unsigned tb[CAP / 32 + 1];              // claimed bits (lchain.c: t[]); the sentinel's bit is never set
        unsigned gp[CAP / 32];                  // bit i: f[i] - f[p[i]] > 0 (f[i] > 0 for a root), the sign a one-step walk needs
    };
    struct Compact {
        unsigned long long wk[WC], wtmp[WC];    // chain-start keys (x of the first anchor) + sort scratch
        unsigned short wpay[WC], wpay2[WC];     // chain ids + sort scratch
        unsigned cnt[256];
        unsigned short start[kBtLevels * kBtRow];
    };
    union {
        Walk w;
        Compact c;
    };
};

// ---- where the walk keeps its state: shared memory (reads of <= CAP anchors) ...
template <int CAP>
struct WalkSmall {
    typedef ZKey ZK;
    typedef unsigned short IDX;     // anchor index / chain id
    typedef unsigned short POS;
    BtWalkSmem<CAP> &S;
    const unsigned *zs;
    __device__ __forceinline__ WalkSmall(BtWalkSmem<CAP> &s, const unsigned *z) : S(s), zs(z) {}
    __device__ __forceinline__ int sent() const { return CAP; }
    __device__ __forceinline__ int wc() const { return BtWalkSmem<CAP>::WC; }
    __device__ __forceinline__ void prepare_compaction(int) {}
    static constexpr bool kLanePar = false;     // (its scratch would cost these kernels a resident read per SM)
    static constexpr int kZq = 6;               // groups of sorted ends in flight ahead of the scan
    static constexpr int kZpf = 0;
    __device__ __forceinline__ void zprefetch(int) const {}
    __device__ __forceinline__ unsigned *lp_z() { return nullptr; }
    __device__ __forceinline__ unsigned short *lp_path() { return nullptr; }
    __device__ __forceinline__ void init(int n, const int *__restrict__ fr, const int *__restrict__ pr, int lane)
    {
        // 4 coalesced loads of each array per lane, then 4 gathers of f[p[i]]; the loads of the NEXT 128 anchors are issued
        // before the gathers are consumed, so an iteration costs one memory latency, not two
        int pn[4], fn[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { const int i = t * 32 + lane; pn[t] = i < n ? pr[i] : -1; fn[t] = i < n ? fr[i] : 0; }
        for (int i0 = 0; i0 < n; i0 += 128) {
            int pv[4], fv[4], fp[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) { pv[t] = pn[t]; fv[t] = fn[t]; }
#pragma unroll
            for (int t = 0; t < 4; ++t) fp[t] = pv[t] >= 0 ? fr[pv[t]] : 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) { const int i = i0 + 128 + t * 32 + lane; pn[t] = i < n ? pr[i] : -1; fn[t] = i < n ? fr[i] : 0; }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int i = i0 + t * 32 + lane;
                if (i < n) S.w.ps[i] = pv[t] < 0 ? (unsigned short)CAP : (unsigned short)pv[t];
                const unsigned g = __ballot_sync(0xffffffffu, i < n && fv[t] - fp[t] > 0);
                if (lane == 0 && i0 + t * 32 < n) { S.w.gp[(i0 >> 5) + t] = g; S.w.tb[(i0 >> 5) + t] = 0; }
            }
        }
        if (lane == 0) { S.w.ps[CAP] = (unsigned short)CAP; S.w.tb[CAP / 32] = 0; }
    }
    __device__ __forceinline__ int fat(int i, const int *__restrict__ fr) const { return i < CAP ? fr[i] : 0; }
    __device__ __forceinline__ int nextp(int cur) const { return S.w.ps[cur]; }
    __device__ __forceinline__ int nextp_blind(int cur, unsigned &) const { return S.w.ps[cur]; }
    __device__ __forceinline__ bool claimed(int i) const { return ((S.w.tb[i >> 5] >> (i & 31)) & 1u) != 0; }
    __device__ __forceinline__ void claim(int i) { atomicOr(&S.w.tb[i >> 5], 1u << (i & 31)); }
    __device__ __forceinline__ bool gain(int i, const int *, const int *) const { return ((S.w.gp[i >> 5] >> (i & 31)) & 1u) != 0; }
    __device__ __forceinline__ unsigned zat(int e) const { return zs[e]; }
    __device__ __forceinline__ IDX *path() { return S.w.path; }
    __device__ __forceinline__ unsigned long long *wk() { return S.c.wk; }
    __device__ __forceinline__ unsigned long long *wtmp() { return S.c.wtmp; }
    __device__ __forceinline__ IDX *wpay() { return S.c.wpay; }
    __device__ __forceinline__ IDX *wpay2() { return S.c.wpay2; }
    __device__ __forceinline__ unsigned *cnt() { return S.c.cnt; }
    __device__ __forceinline__ POS *start() { return S.c.start; }
};

// ---- ... or global scratch (any read).  p[] is read in place, the claimed bits sit in a global bit array, the sorted z[]
//      and (after the walks, when z[] is dead) the chain-start keys use the read's zk / zk2 scratch.
struct WalkBig {
    typedef ZKey64 ZK;
    typedef int IDX;
    typedef unsigned POS;
    int n;
    const int *pr;
    unsigned *tb;
    unsigned long long *zk, *zk2;
    unsigned *pay, *pay2;
    int *path_s;
    unsigned *cnt_s, *start_s;
    __device__ __forceinline__ int sent() const { return n; }
    __device__ __forceinline__ int wc() const { return n; }
    __device__ __forceinline__ void prepare_compaction(int) {}
    static constexpr bool kLanePar = false;
    static constexpr int kZq = 6;
    static constexpr int kZpf = 0;
    __device__ __forceinline__ void zprefetch(int) const {}
    __device__ __forceinline__ unsigned long long *lp_z() { return nullptr; }
    __device__ __forceinline__ unsigned short *lp_path() { return nullptr; }
    __device__ __forceinline__ void init(int n_, const int *, const int *, int lane)
    {
        for (int w = lane; w <= (n_ >> 5); w += 32) tb[w] = 0;
    }
    __device__ __forceinline__ int nextp(int cur) const
    {
        int nx = n;
        if (cur < n) { const int q = pr[cur]; nx = q < 0 ? n : q; }
        return nx;
    }
    __device__ __forceinline__ int nextp_blind(int cur, unsigned &) const { return nextp(cur); }
    __device__ __forceinline__ int fat(int i, const int *fr) const { return i < n ? fr[i] : 0; }
    __device__ __forceinline__ bool claimed(int i) const { return ((tb[i >> 5] >> (i & 31)) & 1u) != 0; }
    __device__ __forceinline__ void claim(int i) { atomicOr(&tb[i >> 5], 1u << (i & 31)); }
    __device__ __forceinline__ bool gain(int i, const int *fr, const int *prr) const
    {
        const int q = prr[i];
        return fr[i] - (q >= 0 ? fr[q] : 0) > 0;
    }
    __device__ __forceinline__ unsigned long long zat(int e) const { return zk[e]; }
    __device__ __forceinline__ IDX *path() { return path_s; }
    __device__ __forceinline__ unsigned long long *wk() { return zk; }
    __device__ __forceinline__ unsigned long long *wtmp() { return zk2; }
    __device__ __forceinline__ IDX *wpay() { return reinterpret_cast<int *>(pay); }
    __device__ __forceinline__ IDX *wpay2() { return reinterpret_cast<int *>(pay2); }
    __device__ __forceinline__ unsigned *cnt() { return cnt_s; }
    __device__ __forceinline__ POS *start() { return start_s; }
};

// ---- ... or in between (reads of up to 196608 anchors): the predecessor links sit in shared memory as ONE BYTE per anchor,
//      the distance i - p[i] (0 = none, 255 = "255 or more": look p[i] up in global memory), next to the claimed bits.  The
//      pointer chase -- the serial part of every walk -- then runs at shared-memory latency; f[], the sorted z[] and the
//      chain-start keys stay in global scratch as in WalkBig.  The sort scratch of the compaction aliases the links (dead by then).
struct WalkMid {
    typedef ZKey64 ZK;
    typedef int IDX;
    typedef unsigned POS;
    int n;
    const int *pr;
    unsigned char *rel;     // [n + 1], rel[n] = 0: the sentinel points at itself
    unsigned *tb;           // [n / 32 + 1]
    unsigned long long *zk, *zk2;
    unsigned *pay, *pay2;
    int *path_s;
    unsigned *cnt_s, *start_s;
    size_t smem_bytes;      // dynamic shared memory of the CTA
    unsigned short *lp_path_s;   // [32][33]: the paths of the lane-parallel walks, as distances below the lane's chain end
    static constexpr bool kLanePar = true;
    // the sorted ends sit in global scratch (an L2 round trip per group) and a run of claimed windows is skipped in a few dozen
    // cycles each: 16 groups in flight (the kernel has registers to spare: shared memory caps it at a few warps per SM)
    static constexpr int kZq = 14;
    // ... and the lines 64 groups (16 KB) further down are pulled into L2 meanwhile: the sorted ends of a batch of long reads
    // (8 B per anchor) do not survive in L2 between the sort and the walk kernel, and a DRAM round trip is ~3 groups of scanning
    static constexpr int kZpf = 64;
    __device__ __forceinline__ void zprefetch(int e) const
    {
        if (e >= 0 && (threadIdx.x & 15) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(zk + e));   // 32 lanes x 8 B = two 128-byte lines
    }
    unsigned long long *lp_z_s;   // [32] pending ends
    __device__ __forceinline__ unsigned long long *lp_z() { return lp_z_s; }
    __device__ __forceinline__ unsigned short *lp_path() { return lp_path_s; }
    __device__ __forceinline__ int sent() const { return n; }
    __device__ __forceinline__ int wc() const { return n; }
    // the chain-start keys of the compaction (sorted with the serial flag passes) go to shared memory when they fit behind
    // the sort scratch: the links are dead by then
    __device__ __forceinline__ void prepare_compaction(int n_u)
    {
        const size_t need = (size_t)(256 + kBtLevels * kBtRow) * 4 + (size_t)n_u * 24;
        if (need > smem_bytes) return;
        unsigned long long *base = reinterpret_cast<unsigned long long *>(cnt_s + 256 + kBtLevels * kBtRow);
        zk = base;
        zk2 = base + n_u;
        pay = reinterpret_cast<unsigned *>(base + 2 * (size_t)n_u);
        pay2 = pay + n_u;
    }
    __device__ __forceinline__ void init(int n_, const int *, const int *__restrict__ prr, int lane)
    {
        for (int i0 = 0; i0 < n_; i0 += 256) {   // 8 coalesced loads per lane in flight
            int pv[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) { const int i = i0 + t * 32 + lane; pv[t] = i < n_ ? prr[i] : -1; }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int i = i0 + t * 32 + lane;
                if (i < n_) rel[i] = pv[t] < 0 ? (unsigned char)0 : (unsigned char)min(i - pv[t], 255);
            }
        }
        if (lane == 0) rel[n_] = 0;
        for (int w = lane; w <= (n_ >> 5); w += 32) tb[w] = 0;
    }
    __device__ __forceinline__ int nextp(int cur) const
    {
        const unsigned d = rel[cur];
        int nx = cur - (int)d;
        if (d == 0u) nx = n;
        else if (d == 255u) nx = pr[cur];
        return nx;
    }
    // the same step for a blind chase of many nodes: no branch on the dependent chain (load, compare, select).  A link of 255
    // or more is followed wrongly and flagged in esc (esc == 255 afterwards: redo that stretch with nextp)
    __device__ __forceinline__ int nextp_blind(int cur, unsigned &esc) const
    {
        const unsigned d = rel[cur];
        esc = max(esc, d);
        return d ? cur - (int)d : n;
    }
    __device__ __forceinline__ int fat(int i, const int *fr) const { return i < n ? fr[i] : 0; }
    __device__ __forceinline__ bool claimed(int i) const { return ((tb[i >> 5] >> (i & 31)) & 1u) != 0; }
    __device__ __forceinline__ void claim(int i) { atomicOr(&tb[i >> 5], 1u << (i & 31)); }
    __device__ __forceinline__ bool gain(int i, const int *fr, const int *prr) const
    {
        const int q = prr[i];
        return fr[i] - (q >= 0 ? fr[q] : 0) > 0;
    }
    __device__ __forceinline__ unsigned long long zat(int e) const { return zk[e]; }
    __device__ __forceinline__ IDX *path() { return path_s; }
    __device__ __forceinline__ unsigned long long *wk() { return zk; }
    __device__ __forceinline__ unsigned long long *wtmp() { return zk2; }
    __device__ __forceinline__ IDX *wpay() { return reinterpret_cast<int *>(pay); }
    __device__ __forceinline__ IDX *wpay2() { return reinterpret_cast<int *>(pay2); }
    __device__ __forceinline__ unsigned *cnt() { return cnt_s; }
    __device__ __forceinline__ POS *start() { return start_s; }
};

// Chain extraction + compaction of one read by one warp; W = where the state lives (above).
// inputs : ar / fr / pr = the read's anchors, scores, predecessors (index inside the read, -1 none); nz sorted ends in W
// scratch: vr (int per anchor; the chains' anchor indices in emission order), ur (u64 per anchor), vsr (int per anchor)
// outputs: n_u (-1 = handed to the global-memory kernels), n_b, u_pos, b_pos of read r; the INDICES (inside the read) of the n_b
//          compacted anchors at v_pack[b_pos ..] -- compact_a's a'[k] = a[v_pack[b_pos + k]], a gather the host does from the
//          anchors it still holds, so 4 instead of 16 bytes per chain anchor leave the device -- and the n_u chains
//          (score << 32 | count) at u_pack[u_pos ..]: packed arrays shared by the batch, slots handed out by atomic cursors,
//          so only what was produced has to leave the device
template <class W>
__device__ __forceinline__ void bt_walk_body(W &S, int r, int n, int nz, const uint4 *__restrict__ ar, const int *__restrict__ fr,
                                             const int *__restrict__ pr, BtParams bp, int *vr, unsigned long long *ur, int *vsr,
                                             int *__restrict__ v_pack, unsigned long long *__restrict__ u_pack, int u_cap,
                                             int *__restrict__ n_u_out, int *__restrict__ n_b_out, int *__restrict__ u_pos,
                                             int *__restrict__ b_pos, int *ovf_list, Counters *ctr, int lane)
{
    typedef typename W::ZK ZK;
    typedef typename W::IDX IDX;
    typedef typename W::POS POS;
    const unsigned full = 0xffffffffu;
    const int SENT = S.sent();       // "no predecessor"
    S.init(n, fr, pr, lane);
    __syncwarp();

    // ---- chain extraction, best end first (lchain.c:42-72 with mg_chain_bk_end :9-25) --------------------------------------
    int n_v = 0, n_u = 0;
    int k = nz - 1;
    bool nothing_claimed = true;    // true until the first chain is claimed: its walk needs no claim tests at all
    // the sorted ends are read through a sliding window of two register groups: zc = z[B - lane], zn = z[B - 32 - lane]
    // (fetched one group ahead); the 32 ends below any k in (B - 32, B] come out of them by shuffles, so the array is loaded
    // once per 32 ends and not once per walk
    typedef typename ZK::T ZT;
    int B = k;
    ZT zc = (B - lane >= 0) ? S.zat(B - lane) : (ZT)0;
    ZT zn = (B - 32 - lane >= 0) ? S.zat(B - 32 - lane) : (ZT)0;
    // further groups in flight (zq[g] = z[B - 64 - 32 g - lane]): a run of claimed windows is skipped much faster than a load
    // returns, so the scan needs about latency / (time per window) loads outstanding
    constexpr int kZq = W::kZq;
    ZT zq[kZq];
#pragma unroll
    for (int g = 0; g < kZq; ++g) zq[g] = (B - 64 - 32 * g - lane >= 0) ? S.zat(B - 64 - 32 * g - lane) : (ZT)0;
    IDX *path = S.path();
    // one cooperative walk from the end zkk = (score, index): the whole warp chases its path and evaluates it
    auto walk_one = [&](typename ZK::T zkk) {
        const int i0 = ZK::idx(zkk), key = ZK::score(zkk);
        // path n_0 = i0, n_1 = p[n_0], ...; node n_j (j >= 1) is "evaluated": s_j = key - f[n_j] (key if n_j is the sentinel).
        // cutj = largest evaluated j whose s_j is a strict new maximum (0 if none): the chain is n_0 .. n_{cutj-1}.
        // The predecessor chase is the serial part: 32 nodes per batch, one load per node (the sentinel points at itself, so
        // the chase needs no end test); where the walk ends is found afterwards for all 32 nodes at once.
        int cur = i0, max_s = 0, cutj = 0, cutf = 0, j0 = 0;   // cutf = f of the node the chain is cut at
        int mine0 = SENT; // this lane's node of the first batch (enough to mark chains of <= 32 nodes without re-reading)
        bool have = false;  // path[] already holds the next batch (chased ahead, below)
        // 32 nodes from cur on into path[], blind (the sentinel points at itself, so no end test)
        auto chase32 = [&]() {
            const int cur0 = cur;
            unsigned esc = 0;
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                if (lane == 0) path[b] = (IDX)cur;
                cur = S.nextp_blind(cur, esc);
            }
            if (esc == 255u) {   // (WalkMid) a link that does not fit a byte was followed wrongly: the same stretch, looking it up
                cur = cur0;
                for (int b = 0; b < 32; ++b) {
                    if (lane == 0) path[b] = (IDX)cur;
                    cur = S.nextp(cur);
                }
            }
        };
        for (;;) {
            int nb = 32;
            if (have) {
                have = false;
            } else if (j0 == 0 && !nothing_claimed) {
                // first batch of a later walk: most of them end within a few nodes (at a claimed anchor), so look as we go
                nb = 0;
#pragma unroll 4
                for (int b = 0; b < 32; ++b) {
                    if (lane == 0) path[b] = (IDX)cur;
                    nb = b + 1;
                    const bool cl = S.claimed(cur);                 // both loads depend on cur only: issued together
                    const int nxt = S.nextp(cur);
                    if (cur == SENT || (cl && b >= 1)) break;
                    cur = nxt;
                }
            } else {
                chase32();
            }
            __syncwarp();
            const int mine = lane < nb ? (int)path[lane] : SENT;
            int fmine = 0;
            if (lane < nb) fmine = S.fat(mine, fr);      // 0 for the sentinel; a gather from global memory
            __syncwarp();
            if (nb == 32) {
                // a full batch: the walk most likely goes on, so the next 32 nodes are chased (shared memory only) while the
                // gather is in flight; if the walk ends in this batch the chase was for nothing
                chase32();
                have = true;
            }
            if (j0 == 0) mine0 = mine;
            const int j = j0 + lane;
            const bool ev = j >= 1 && lane < nb;
            int s = INT32_MIN;
            // the walk stops after evaluating a node that is the root's "predecessor" or already claimed (lchain.c:22)
            const bool stop = ev && (mine == SENT || S.claimed(mine));
            if (ev) s = key - fmine;
            if (lane < nb && mine != SENT && n_v + j < n) vr[n_v + j] = mine;   // speculative: only the first cutj entries count
            // prefix maxima (max_s carried in), exclusive for the tests of lchain.c:20-21
            int pm = ev ? s : INT32_MIN;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(full, pm, d);
                if (lane >= d) pm = max(pm, y);
            }
            int pmprev = __shfl_up_sync(full, pm, 1);
            if (lane == 0) pmprev = INT32_MIN;
            pmprev = max(pmprev, max_s);
            const bool newmax = ev && s > pmprev;
            const bool brk = ev && !newmax && (long long)pmprev - (long long)s > (long long)bp.max_drop;
            const unsigned endm = __ballot_sync(full, stop || brk);
            const unsigned upto = endm ? (0xffffffffu >> (31 - (__ffs(endm) - 1))) : full;   // lanes evaluated in this batch
            const unsigned nm = __ballot_sync(full, newmax) & upto;
            if (nm) {
                const int l = 31 - __clz(nm);
                cutj = j0 + l;
                cutf = __shfl_sync(full, fmine, l);
            }
            if (endm) break;
            max_s = max(max_s, __shfl_sync(full, pm, 31));
            j0 += 32;
        }
        const int cnt = cutj;
        if (cnt > 0) { // claim n_0 .. n_{cnt-1}  (stays claimed even if the chain is rejected below, as in the reference)
            nothing_claimed = false;
            if (cnt <= 32) {
                if (lane < cnt) S.claim(mine0);
            } else {
                __syncwarp();
                for (int q0 = 0; q0 < cnt; q0 += 256) {   // 8 loads of the recorded path per lane in flight
                    int nd[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) { const int q = q0 + t * 32 + lane; nd[t] = q < cnt ? vr[n_v + q] : -1; }
#pragma unroll
                    for (int t = 0; t < 8; ++t) if (nd[t] >= 0) S.claim(nd[t]);
                }
            }
        }
        __syncwarp();
        const int scv = key - cutf;
        if (scv >= bp.min_sc && cnt > 0 && cnt >= bp.min_cnt) {
            if (lane == 0) { ur[n_u] = ((unsigned long long)(unsigned)scv << 32) | (unsigned)cnt; vsr[n_u] = n_v; }
            ++n_u;
            n_v += cnt;
        }
    };
    // ---- lane-parallel walks (W::kLanePar: reads full of short side chains) ----------------------------------------------------
    // Unclaimed ends are not walked as they are met: they are collected, in visiting order, into a batch of up to 32 (the
    // claimed bits do not change meanwhile, so "unclaimed" stays true).  Then every lane walks ITS end on its own against the
    // claimed bits of now -- chase (<= 32 nodes, shared memory), one gather of f per path position, evaluation in the lane --
    // and the results are committed in visiting order.  A result is used iff no node the lane evaluated was claimed by an
    // earlier end of the batch in the meantime (then the walk is what the sequential algorithm does; an end that was itself
    // claimed is dropped); otherwise, and for paths of more than 32 nodes, the end is walked by the whole warp at its turn.
    int nb_pend = 0;
    auto flush_batch = [&]() {
        const bool unc = lane < nb_pend;
        const unsigned m = nb_pend >= 32 ? full : ((1u << nb_pend) - 1u);
        const ZT z = unc ? S.lp_z()[lane] : (ZT)0;
        const int i0l = ZK::idx(z);
        // a path is recorded as 16-bit distances below its end (0xffff = the sentinel); an end whose path leaves that range is
        // walked by the whole warp like any long one.  (As ints the 32 paths cost the 64 k class a resident read per SM.)
        unsigned short *lp = S.lp_path() + lane * 33;
        auto lp_node = [&](const unsigned short *q, int t, int end_i) { const int v = q[t]; return v == 0xffff ? SENT : end_i - v; };
        int len = 0;
        bool longw = false;
        if (unc) {
            int cur = i0l;
            for (int t = 0; t < 32; ++t) {
                if (cur != SENT && i0l - cur > 0xfffe) { longw = true; break; }   // (t >= 1 here: lp[0] is set)
                lp[t] = cur == SENT ? (unsigned short)0xffff : (unsigned short)(i0l - cur);
                len = t + 1;
                if (t >= 1 && (cur == SENT || S.claimed(cur))) break;
                if (t == 31) { longw = true; break; }
                cur = S.nextp(cur);
            }
        }
        // f of every recorded node, 8 independent loads per lane in flight, evaluated as it arrives (lchain.c:16-22 per lane)
        int cutl = 0, maxs = 0, nev = 0;
        {
            const int glen = (unc && !longw) ? len : 0;
            const int maxlen = __reduce_max_sync(full, glen);
            const int keyl = ZK::score(z);
            bool live = true;
            for (int t0 = 1; t0 < maxlen; t0 += 8) {
                int fv[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) fv[q] = t0 + q < glen ? S.fat(lp_node(lp, t0 + q, i0l), fr) : 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (live && t0 + q < glen) {
                        const int sj = keyl - fv[q];
                        nev = t0 + q;
                        if (sj > maxs) { maxs = sj; cutl = t0 + q; }
                        else if ((long long)maxs - (long long)sj > (long long)bp.max_drop) live = false;
                    }
                }
            }
        }
        __syncwarp();
        unsigned todo = m;
        while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const int len_l = __shfl_sync(full, len, l), nev_l = __shfl_sync(full, nev, l), cut_l = __shfl_sync(full, cutl, l);
            const int sc_l = __shfl_sync(full, maxs, l);
            const bool long_l = __shfl_sync(full, longw ? 1 : 0, l) != 0;
            const unsigned short *rp = S.lp_path() + l * 33;
            const int end_l = __shfl_sync(full, i0l, l);
            const int node = (long_l ? lane == 0 : lane <= nev_l) ? lp_node(rp, lane, end_l) : SENT;
            // nodes 0 .. len-2 were unclaimed when the lane walked, the last one is where it stopped
            const bool nowc = node != SENT && (long_l || lane < len_l - 1) && S.claimed(node);
            const unsigned chg = __ballot_sync(full, nowc);
            if (chg & 1u) continue;                       // the end itself has been claimed since: not a chain end any more
            if (long_l || chg) { walk_one(__shfl_sync(full, z, l)); continue; }
            if (cut_l > 0) {
                nothing_claimed = false;
                if (lane < cut_l) S.claim(node);
                if (sc_l >= bp.min_sc && cut_l >= bp.min_cnt) {
                    if (lane == 0) { ur[n_u] = ((unsigned long long)(unsigned)sc_l << 32) | (unsigned)cut_l; vsr[n_u] = n_v; }
                    if (lane < cut_l) vr[n_v + lane] = node;
                    ++n_u;
                    n_v += cut_l;
                }
            }
            __syncwarp();
        }
        nb_pend = 0;
    };
    while (k >= 0) {
        typename ZK::T zkk;
        {   // next chain end that is not claimed yet.  Ends whose predecessor is already claimed (or absent) are one-step
            // walks that can only claim themselves (lchain.c:16-22 evaluates p[i] once and stops): a run of them is settled
            // here in parallel, in visiting order; the first end that needs a real walk goes through the general code below.
            const int e = k - lane;
            while (k <= B - 32) {                          // the window left group zc behind
                zc = zn;
                zn = zq[0];
#pragma unroll
                for (int g = 0; g + 1 < kZq; ++g) zq[g] = zq[g + 1];
                B -= 32;
                zq[kZq - 1] = (B - 64 - 32 * (kZq - 1) - lane >= 0) ? S.zat(B - 64 - 32 * (kZq - 1) - lane) : (ZT)0;
                S.zprefetch(B - 32 * W::kZpf - lane);
            }
            ZT z = zc;
            const int sft = B - k;                         // warp-uniform, 0 .. 31
            if (sft) {
                const int srcl = lane + sft;
                const ZT a0 = __shfl_sync(full, zc, srcl & 31), a1 = __shfl_sync(full, zn, srcl & 31);
                z = srcl < 32 ? a0 : a1;
            }
            bool unc = false, simple = false;
            int i0l = 0, n1 = SENT;
            if (e >= 0) {
                i0l = ZK::idx(z);
                unc = !S.claimed(i0l);
                if (unc) {
                    n1 = S.nextp(i0l);
                    simple = n1 == SENT || S.claimed(n1);
                }
            }
            const unsigned m = __ballot_sync(full, unc);
            if (!m) { k -= 32; continue; }
            const unsigned hard = __ballot_sync(full, unc && !simple);
            if (W::kLanePar) {
                const int cntw = __popc(m);
                if (nb_pend + cntw > 32) { flush_batch(); continue; }   // claims changed: look at this window again
                if (unc) S.lp_z()[nb_pend + __popc(m & ((1u << lane) - 1u))] = z;
                nb_pend += cntw;
                k -= 32;
                continue;
            }
            const int nfast = hard ? __ffs(hard) - 1 : 32;
            const unsigned fastm = nfast >= 32 ? m : (m & ((1u << nfast) - 1u));
            if (fastm) {
                const bool mine_f = ((fastm >> lane) & 1u) != 0;
                const bool claim = mine_f && S.gain(i0l, fr, pr);   // s_1 > 0: cut = p[i0], chain = {i0}
                if (claim) S.claim(i0l);
                if (__any_sync(full, claim)) nothing_claimed = false;
                if (bp.min_cnt <= 1) { // single-anchor chains can be accepted: needs the value of s_1
                    const int keyl = ZK::score(z);
                    const int s1 = claim ? keyl - S.fat(n1, fr) : 0;
                    const bool acc = claim && s1 >= bp.min_sc;
                    const unsigned am = __ballot_sync(full, acc);
                    if (am) {
                        const int rank = __popc(am & ((1u << lane) - 1u));
                        if (acc) {
                            ur[n_u + rank] = ((unsigned long long)(unsigned)s1 << 32) | 1ULL;
                            vsr[n_u + rank] = n_v + rank;
                            vr[n_v + rank] = i0l;
                        }
                        n_u += __popc(am);
                        n_v += __popc(am);
                    }
                }
                __syncwarp();
            }
            k -= nfast;
            if (!hard) continue;
            zkk = __shfl_sync(full, z, nfast);     // the end that needs a real walk was fetched by lane nfast
        }
        walk_one(zkk);
        --k;
    }
    if (W::kLanePar && nb_pend) flush_batch();
    __syncwarp();
    if (n_u == 0) { if (lane == 0) { n_u_out[r] = 0; n_b_out[r] = 0; u_pos[r] = 0; b_pos[r] = 0; } return; }

    // ---- compact_a (lchain.c:78-111): chains flipped to ascending order, then ordered by the x of their first anchor with
    //      the same unstable sort (w[i].x = b[k].x, payload = chain id) --------------------------------------------------------
    if (n_u > S.wc()) { // more chains than the key buffer holds: the global-memory kernels take the read
        if (lane == 0) { n_u_out[r] = -1; n_b_out[r] = 0; if (ovf_list) bt_overflow(r, ovf_list, ctr); }
        return;
    }
    int upos = 0, bpos = 0;
    if (lane == 0) upos = atomicAdd(&ctr->u_cur, n_u);
    upos = __shfl_sync(full, upos, 0);
    if (upos + n_u > u_cap) { if (lane == 0) { n_u_out[r] = -1; n_b_out[r] = 0; } return; }   // cannot happen: u_cap = anchors of the batch
    if (lane == 0) bpos = atomicAdd(&ctr->b_cur, n_v);
    bpos = __shfl_sync(full, bpos, 0);
    unsigned long long *uo = u_pack + upos;
    int *bo = v_pack + bpos;
    S.prepare_compaction(n_u);
    unsigned long long *wk = S.wk();
    IDX *wpay = S.wpay();
    for (int c = lane; c < n_u; c += 32) {
        const int cntc = (int)(unsigned)ur[c], s0 = vsr[c];
        const uint4 av = ar[vr[s0 + cntc - 1]];
        wk[c] = ((unsigned long long)av.y << 32) | av.x;
        wpay[c] = (IDX)c;
    }
    __syncwarp();
    BtSortScratch<POS> sc;
    sc.cnt = S.cnt();
    sc.start = S.start();
    bt_sort<WKey, true, IDX, POS>(wk, wpay, S.wtmp(), S.wpay2(), n_u, sc, lane);
    __syncwarp();
    int out = 0;
    if (n_u < 8) { // a few chains (the usual read): chain by chain, 4 independent gathers per lane in flight
        for (int c = 0; c < n_u; ++c) {
            const int src = (int)wpay[c];
            const unsigned long long uv = ur[src];
            const int cntc = (int)(unsigned)uv, s0 = vsr[src];
            if (lane == 0) uo[c] = uv;
            // the chain was collected end-first: flipped copy of its indices, 8 coalesced loads per lane in flight
            for (int q0 = 0; q0 < cntc; q0 += 256) {
                int idx[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) { const int q = q0 + t * 32 + lane; idx[t] = q < cntc ? vr[s0 + cntc - 1 - q] : -1; }
#pragma unroll
                for (int t = 0; t < 8; ++t) if (idx[t] >= 0) bo[out + q0 + t * 32 + lane] = idx[t];
            }
            out += cntc;
        }
        if (lane == 0) { n_u_out[r] = n_u; n_b_out[r] = out; u_pos[r] = upos; b_pos[r] = bpos; }
        return;
    }
    // Where every chain (in sorted order) starts in the output and where its anchors sit in vr[]: the key array is done with, it
    // now holds  out start | (vr index of the chain's last = smallest anchor) << 32.  The anchors are then copied by ONE flat
    // loop over output positions (4 per lane in flight) instead of chain by chain: a read with hundreds of short chains would
    // pay two dependent global loads per chain otherwise.
    for (int c0 = 0; c0 < n_u; c0 += 32) {
        const int c = c0 + lane;
        int cntc = 0, vb = 0;
        if (c < n_u) {
            const int src = (int)wpay[c];
            const unsigned long long uv = ur[src];
            cntc = (int)(unsigned)uv;
            vb = vsr[src] + cntc - 1;
            uo[c] = uv;
        }
        int incl = cntc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(full, incl, d);
            if (lane >= d) incl += y;
        }
        if (c < n_u) wk[c] = (unsigned long long)(unsigned)(out + incl - cntc) | ((unsigned long long)(unsigned)vb << 32);
        out += __shfl_sync(full, incl, 31);
    }
    __syncwarp();
    int clo = 0;    // chain of position o0 (chain starts ascend with the chain number, so the range of a chunk only moves up)
    for (int o0 = 0; o0 < out; o0 += 128) {
        // chains [clo, chi] cover the 128 positions of this chunk: chi by a warp-wide scan from clo (a chunk usually holds a
        // handful of chains), then every position looks its chain up among those few instead of among all n_u
        const int oend = min(out, o0 + 128) - 1;
        int chi = clo;
        for (;;) {
            const int c = chi + 1 + lane;
            const bool le = c < n_u && (int)(unsigned)wk[c] <= oend;
            const unsigned b = __ballot_sync(full, le);     // a prefix of the lanes
            chi += __popc(b);
            if (b != full) break;
        }
        int idx[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int o = o0 + t * 32 + lane;
            idx[t] = -1;
            if (o < out) {
                int lo = clo, hi = chi + 1;       // the chain of position o: start[lo] <= o < start[hi]
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if ((int)(unsigned)wk[mid] <= o) lo = mid; else hi = mid;
                }
                const unsigned long long rc = wk[lo];
                idx[t] = vr[(int)(rc >> 32) - (o - (int)(unsigned)rc)];
            }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) if (idx[t] >= 0) bo[o0 + t * 32 + lane] = idx[t];
        clo = chi;
    }
    if (lane == 0) { n_u_out[r] = n_u; n_b_out[r] = out; u_pos[r] = upos; b_pos[r] = bpos; }
}

template <int CAP>
__global__ void __launch_bounds__(32)
k_bt_walk(const uint4 *__restrict__ a, const int *__restrict__ f, const int *__restrict__ p, const long long *__restrict__ off,
          const int *__restrict__ read_list, int n_list, BtParams bp, const unsigned *__restrict__ zs_scr, const int *__restrict__ nz_in,
          int *__restrict__ v_scr, unsigned long long *__restrict__ u_scr, int *__restrict__ vs_scr, int *__restrict__ v_pack,
          unsigned long long *__restrict__ u_pack, int u_cap, int *__restrict__ n_u_out, int *__restrict__ n_b_out,
          int *__restrict__ u_pos, int *__restrict__ b_pos, int *__restrict__ ovf_list, Counters *ctr)
{
    extern __shared__ int4 bt_raw[];
    BtWalkSmem<CAP> &SM = *reinterpret_cast<BtWalkSmem<CAP> *>(bt_raw);
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= n_list) return;
    const int r = read_list[blockIdx.x];
    const long long o0 = off[r];
    const int n = (int)(off[r + 1] - o0);
    const int nz = nz_in[r];
    if (nz < 0) { if (lane == 0) { n_u_out[r] = -1; n_b_out[r] = 0; } return; }
    if (nz == 0) { if (lane == 0) { n_u_out[r] = 0; n_b_out[r] = 0; u_pos[r] = 0; b_pos[r] = 0; } return; }
    WalkSmall<CAP> S(SM, zs_scr + o0);
    bt_walk_body(S, r, n, nz, a + o0, f + o0, p + o0, bp, v_scr + o0, u_scr + o0, vs_scr + o0, v_pack, u_pack, u_cap, n_u_out, n_b_out,
                 u_pos, b_pos, ovf_list, ctr, lane);
}

// tb_scr: claimed bits, read r uses the words from (off[r] >> 5) + 2 r on; pay / pay2: one unsigned per anchor each
__global__ void __launch_bounds__(32)
k_bt_walk_big(const uint4 *__restrict__ a, const int *__restrict__ f, const int *__restrict__ p, const long long *__restrict__ off,
              const int *__restrict__ read_list, int n_list, const int *ovf_list, BtParams bp, unsigned long long *zk_scr,
              unsigned long long *zk2_scr, const int *__restrict__ nz_in, unsigned *tb_scr, unsigned *pay_scr, unsigned *pay2_scr,
              int *__restrict__ v_scr, unsigned long long *__restrict__ u_scr, int *__restrict__ vs_scr, int *__restrict__ v_pack,
              unsigned long long *__restrict__ u_pack, int u_cap, int *__restrict__ n_u_out, int *__restrict__ n_b_out,
              int *__restrict__ u_pos, int *__restrict__ b_pos, Counters *ctr)
{
    __shared__ unsigned s_cnt[256];
    __shared__ unsigned s_start[kBtLevels * kBtRow];
    __shared__ int s_path[32];
    const int lane = threadIdx.x;
    const int r = bt_big_read(read_list, n_list, ovf_list, ctr);
    if (r < 0) return;
    const long long o0 = off[r];
    const int n = (int)(off[r + 1] - o0);
    const int nz = nz_in[r];
    if (nz < 0) { if (lane == 0) { n_u_out[r] = -1; n_b_out[r] = 0; } return; }
    if (nz == 0) { if (lane == 0) { n_u_out[r] = 0; n_b_out[r] = 0; u_pos[r] = 0; b_pos[r] = 0; } return; }
    WalkBig S;
    S.n = n;
    S.pr = p + o0;
    S.tb = tb_scr + (o0 >> 5) + 2 * (long long)r;
    S.zk = zk_scr + o0;
    S.zk2 = zk2_scr + o0;
    S.pay = pay_scr + o0;
    S.pay2 = pay2_scr + o0;
    S.path_s = s_path;
    S.cnt_s = s_cnt;
    S.start_s = s_start;
    bt_walk_body(S, r, n, nz, a + o0, f + o0, p + o0, bp, v_scr + o0, u_scr + o0, vs_scr + o0, v_pack, u_pack, u_cap, n_u_out, n_b_out,
                 u_pos, b_pos, nullptr, ctr, lane);
}

// dynamic shared memory of k_bt_walk_mid for reads of up to cap anchors: links + claimed bits (the compaction's sort scratch,
// (256 + kBtLevels * kBtRow) words, reuses the front of it: cap >= 16384)
__host__ __device__ inline size_t bt_walk_mid_smem(int cap) { return (((size_t)cap + 1 + 15) & ~(size_t)15) + ((size_t)cap / 32 + 1) * 4 + 16; }

__global__ void __launch_bounds__(32)
k_bt_walk_mid(const uint4 *__restrict__ a, const int *__restrict__ f, const int *__restrict__ p, const long long *__restrict__ off,
              const int *__restrict__ read_list, int n_list, BtParams bp, unsigned long long *zk_scr, unsigned long long *zk2_scr,
              const int *__restrict__ nz_in, unsigned *pay_scr, unsigned *pay2_scr, int *__restrict__ v_scr,
              unsigned long long *__restrict__ u_scr, int *__restrict__ vs_scr, int *__restrict__ v_pack,
              unsigned long long *__restrict__ u_pack, int u_cap, int *__restrict__ n_u_out, int *__restrict__ n_b_out,
              int *__restrict__ u_pos, int *__restrict__ b_pos, Counters *ctr, int cap)
{
    extern __shared__ int4 bt_raw[];
    __shared__ int s_path[32];
    __shared__ unsigned short s_lp_path[32 * 33];
    __shared__ unsigned long long s_lp_z[32];
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= n_list) return;
    const int r = read_list[blockIdx.x];
    const long long o0 = off[r];
    const int n = (int)(off[r + 1] - o0);
    const int nz = nz_in[r];
    if (nz < 0 || n > cap) { if (lane == 0) { n_u_out[r] = -1; n_b_out[r] = 0; } return; }
    if (nz == 0) { if (lane == 0) { n_u_out[r] = 0; n_b_out[r] = 0; u_pos[r] = 0; b_pos[r] = 0; } return; }
    WalkMid S;
    S.n = n;
    S.pr = p + o0;
    S.rel = reinterpret_cast<unsigned char *>(bt_raw);
    S.tb = reinterpret_cast<unsigned *>(S.rel + (((size_t)cap + 1 + 15) & ~(size_t)15));
    S.zk = zk_scr + o0;
    S.zk2 = zk2_scr + o0;
    S.pay = pay_scr + o0;
    S.pay2 = pay2_scr + o0;
    S.path_s = s_path;
    S.cnt_s = reinterpret_cast<unsigned *>(bt_raw);
    S.start_s = S.cnt_s + 256;
    S.smem_bytes = bt_walk_mid_smem(cap);
    S.lp_path_s = s_lp_path;
    S.lp_z_s = s_lp_z;
    bt_walk_body(S, r, n, nz, a + o0, f + o0, p + o0, bp, v_scr + o0, u_scr + o0, vs_scr + o0, v_pack, u_pack, u_cap, n_u_out, n_b_out,
                 u_pos, b_pos, nullptr, ctr, lane);
}

constexpr int kDrainThreads = 64;   // small CTAs (2 warps x 32 registers) fit next to a resident score kernel of another slot

// Packed results -> (mapped, pinned) host memory.  The amounts are only known on the device (the cursors), so this is a
// kernel rather than a copy-engine transfer of the worst case: a few CTAs keep enough 16-byte stores in flight for PCIe.
// The indices occupy [first, b_cur) of src_v and go to the same positions of dst_v; both bases are 16-byte aligned (the
// device packs from `first` = the misalignment of the landing area on), so the body moves four indices per store.
__global__ void __launch_bounds__(kDrainThreads)
k_drain(const int *__restrict__ src_v, int *__restrict__ dst_v, int first, const unsigned long long *__restrict__ src_u,
        unsigned long long *__restrict__ dst_u, const Counters *__restrict__ ctr)
{
    const long long nb = ctr->b_cur, nu = ctr->u_cur;
    const long long stride = (long long)gridDim.x * blockDim.x, t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long q0 = first ? 1 : 0, nq = nb >> 2;     // whole 16-byte groups [q0, nq)
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src_v);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst_v);
    if (first && t < 4 && t >= first && t < nb) dst_v[t] = __ldg(src_v + t);   // the partial first group
    long long i = q0 + t;
    for (; i + 3 * stride < nq; i += 4 * stride) {
        const uint4 v0 = __ldg(s4 + i), v1 = __ldg(s4 + i + stride), v2 = __ldg(s4 + i + 2 * stride), v3 = __ldg(s4 + i + 3 * stride);
        d4[i] = v0; d4[i + stride] = v1; d4[i + 2 * stride] = v2; d4[i + 3 * stride] = v3;
    }
    for (; i < nq; i += stride) d4[i] = __ldg(s4 + i);
    for (long long k = max(nq << 2, (long long)(first ? 4 : 0)) + t; k < nb; k += stride) dst_v[k] = __ldg(src_v + k);
    for (long long k = t; k < nu; k += stride) dst_u[k] = __ldg(src_u + k);
}

// Results of a DEVICE-RESIDENT batch (anchors produced on the device, e.g. by the seeding kernels: the host holds no copy to
// gather from) -> mapped pinned host memory: the compacted anchors themselves, b[b_pos[r] + k] = a[off[r] + v[b_pos[r] + k]]
// (compact_a's gather, lchain.c:100-105, fused into the transfer; one warp per read, 16-byte stores), and the packed chains.
__global__ void __launch_bounds__(256)
k_drain_anchors(const uint4 *__restrict__ a, const long long *__restrict__ off, const int *__restrict__ v, const int *__restrict__ n_b,
                const int *__restrict__ b_pos, int n_reads, uint4 *__restrict__ dst_b, const unsigned long long *__restrict__ src_u,
                unsigned long long *__restrict__ dst_u, const Counters *__restrict__ ctr)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < n_reads; r += n_warps) {
        const int nb = n_b[r], bp = b_pos[r];
        const long long o = off[r];
        for (int t = lane; t < nb; t += 32) dst_b[bp + t] = __ldg(a + o + __ldg(v + bp + t));
    }
    const long long nu = ctr->u_cur, stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nu; k += stride) dst_u[k] = __ldg(src_u + k);
}


} // namespace mm2gb
