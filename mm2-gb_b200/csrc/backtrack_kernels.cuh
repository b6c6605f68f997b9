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
// Reads that do not fit (more than kBtMaxAnchors anchors, scores >= 2^19) are declined (n_u = -1) and go through the
// host implementation (backtrack.cpp), which is the same algorithm.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mm2gb {

constexpr int kBtIdxBits = 13;                          // anchors per read handled on the device: 2^13
constexpr int kBtMaxAnchors = 1 << kBtIdxBits;
constexpr unsigned kBtIdxMask = kBtMaxAnchors - 1;
constexpr int kBtMaxScore = 1 << (32 - kBtIdxBits);     // packed key = f << 13 | i must fit 32 bits
constexpr int kBtLevels = 8;                            // radix levels of a 64-bit key
constexpr int kBtRow = 260;                             // per level: start[0..256], next-bucket cursor, shift

struct BtParams { int min_cnt, min_sc, max_drop; };

// ---- key traits ---------------------------------------------------------------------------------------------------
struct ZKey {   // (score, index) pair packed in 32 bits, ordered by score only (radix_sort_128x on mm128_t.x = f)
    typedef unsigned T;
    __device__ static __forceinline__ unsigned digit(T k, int shift) { return ((k >> kBtIdxBits) >> shift) & 255u; }
    __device__ static __forceinline__ bool less(T a, T b) { return (a >> kBtIdxBits) < (b >> kBtIdxBits); }
    __device__ static __forceinline__ unsigned long long key64(T k) { return (unsigned long long)(k >> kBtIdxBits); }
};
struct WKey {   // chain start position x (64 bit); the chain id travels in a separate 16-bit payload array
    typedef unsigned long long T;
    __device__ static __forceinline__ unsigned digit(T k, int shift) { return (unsigned)(k >> shift) & 255u; }
    __device__ static __forceinline__ bool less(T a, T b) { return a < b; }
    __device__ static __forceinline__ unsigned long long key64(T k) { return k; }
};

// scratch shared by the sorts of one warp
struct BtSortScratch {
    unsigned *cnt;              // [256] histogram, then the bucket cursors of the running pass
    unsigned short *start;      // [kBtLevels][kBtRow]
};

// rank sort (== the reference's insertion sort, ksort.h:105-115: stable) of every bucket of 2..64 elements of one pass;
// start = bucket boundaries of that pass (257 entries), or nullptr to sort [lo, hi) as ONE bucket.
template <class KO, bool PAY>
__device__ void bt_rank_sort(typename KO::T *A, unsigned short *pay, typename KO::T *tmpA, unsigned short *tmpPay, int lo, int hi,
                             int shift, const unsigned short *start, int lane)
{
    typedef typename KO::T K;
    for (int e0 = lo; e0 < hi; e0 += 32) {
        const int e = e0 + lane;
        if (e < hi) {
            const K key = A[e];
            int bs = lo, be = hi;
            if (start) { const unsigned d = KO::digit(key, shift); bs = start[d]; be = start[d + 1]; }
            const int m = be - bs;
            if (m >= 2 && m <= 64) {
                int r = 0;
                for (int j = bs; j < be; ++j) {
                    const K kj = A[j];
                    r += (KO::less(kj, key) || (!KO::less(key, kj) && j < e)) ? 1 : 0;
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
template <class KO, bool PAY>
__device__ void bt_flag_pass(typename KO::T *A, unsigned short *pay, int lo, int hi, int shift, unsigned *cnt, unsigned short *st, int lane)
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
            st[lane * 8 + q] = (unsigned short)run;
            cnt[lane * 8 + q] = run;        // cursor of the bucket
            run += c[q];
        }
        if (lane == 31) st[256] = (unsigned short)hi;
    }
    __syncwarp();
    // buckets in ascending order; everything below is warp-uniform (all lanes read the same shared values)
    for (int k = 0; k < 256; ++k) {
        const int endk = st[k + 1];
        int c = (int)cnt[k];
        while (c < endk) {
            // elements that already sit in their home bucket stay where they are
            const int e = c + lane;
            const bool mis = e < endk && KO::digit(A[e], shift) != (unsigned)k;
            const unsigned mm = __ballot_sync(full, mis);
            if (!mm) { c = min(c + 32, endk); continue; }
            c += __ffs(mm) - 1;
            K carried = A[c];
            unsigned short cpay = PAY ? pay[c] : (unsigned short)0;
            unsigned d = KO::digit(carried, shift);
            do {
                const int pos = (int)cnt[d], endd = st[d + 1];
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
                const unsigned short epay = PAY ? pay[pos + L] : (unsigned short)0;
                for (int top = L; top > 0; top -= 32) { // shift A[pos .. pos+L) up by one, highest chunk first
                    const int base = max(0, top - 32), idx = base + lane;
                    K val = 0;
                    unsigned short pv = 0;
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
template <class KO, bool PAY>
__device__ void bt_sort(typename KO::T *A, unsigned short *pay, typename KO::T *tmpA, unsigned short *tmpPay, int n, BtSortScratch sc, int lane)
{
    const unsigned full = 0xffffffffu;
    if (n <= 1) return;
    if (n <= 64) { bt_rank_sort<KO, PAY>(A, pay, tmpA, tmpPay, 0, n, 0, nullptr, lane); return; }
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
    unsigned short *row = sc.start;
    bt_flag_pass<KO, PAY>(A, pay, 0, n, shift, sc.cnt, row, lane);
    if (shift) bt_rank_sort<KO, PAY>(A, pay, tmpA, tmpPay, 0, n, shift, row, lane);
    if (lane == 0) { row[257] = 0; row[258] = (unsigned short)shift; }
    __syncwarp();
    while (lv >= 0) {
        row = sc.start + lv * kBtRow;
        const int sh = row[258];
        int k = row[257];
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
        if (lane == 0) row[257] = (unsigned short)(found + 1);
        const int blo = row[found], bhi = row[found + 1];
        const int nsh = sh > 8 ? sh - 8 : 0;
        ++lv;
        unsigned short *crow = sc.start + lv * kBtRow;
        __syncwarp();
        bt_flag_pass<KO, PAY>(A, pay, blo, bhi, nsh, sc.cnt, crow, lane);
        if (nsh) bt_rank_sort<KO, PAY>(A, pay, tmpA, tmpPay, blo, bhi, nsh, crow, lane);
        if (lane == 0) { crow[257] = 0; crow[258] = (unsigned short)nsh; }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The stage is two kernels per size class, one warp (= one CTA) per read each, because the two halves want different
// amounts of shared memory: the sort needs 8 bytes per anchor for a few microseconds, the walk only ~3.5 bytes per anchor
// for much longer (it is a serial pointer chase) -- so three times as many walks as sorts fit on an SM.
//   k_bt_sort : z[] = anchors scoring >= min_sc, sorted as the reference sorts them  -> zs_scr (global), nz_out
//   k_bt_walk : chain extraction in that order + compaction                          -> n_u, n_b, b_out, u_pack
// ---------------------------------------------------------------------------------------------------------------------
template <int CAP>
struct BtSortSmem {
    unsigned zk[CAP];               // (score << 13 | index), sorted in place
    unsigned zk2[CAP];              // rank-sort scratch
    unsigned cnt[256];
    unsigned short start[3 * kBtRow];   // scores are below 2^19: at most three radix levels
};

template <int CAP>
__global__ void __launch_bounds__(32)
k_bt_sort(const int *__restrict__ f, const long long *__restrict__ off, const int *__restrict__ read_list, int n_list, BtParams bp,
          unsigned *__restrict__ zs_scr, int *__restrict__ nz_out)
{
    extern __shared__ int4 bt_raw[];
    BtSortSmem<CAP> &S = *reinterpret_cast<BtSortSmem<CAP> *>(bt_raw);
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= n_list) return;
    const int r = read_list[blockIdx.x];
    const long long o0 = off[r];
    const int n = (int)(off[r + 1] - o0);
    const int *fr = f + o0;
    if (n > CAP || bp.min_sc < 0) { if (lane == 0) nz_out[r] = -1; return; }
    // z[]: anchors scoring >= min_sc, in index order (lchain.c:33-40); 4 coalesced loads per lane in flight
    int nz = 0, fmax = 0;
    for (int i0 = 0; i0 < n; i0 += 128) {
        int fv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { const int i = i0 + t * 32 + lane; fv[t] = i < n ? fr[i] : INT32_MIN; }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int i = i0 + t * 32 + lane;
            const bool keep = fv[t] >= bp.min_sc && i < n;
            const unsigned m = __ballot_sync(full, keep);
            if (keep) S.zk[nz + __popc(m & ((1u << lane) - 1u))] = ((unsigned)fv[t] << kBtIdxBits) | (unsigned)i;
            nz += __popc(m);
            fmax = max(fmax, keep ? fv[t] : 0);
        }
    }
    fmax = __reduce_max_sync(full, fmax);
    if (fmax >= kBtMaxScore) { if (lane == 0) nz_out[r] = -1; return; }   // does not pack into 32 bits: host path
    __syncwarp();
    BtSortScratch sc;
    sc.cnt = S.cnt;
    sc.start = S.start;
    bt_sort<ZKey, false>(S.zk, nullptr, S.zk2, nullptr, nz, sc, lane);
    unsigned *zo = zs_scr + o0;
    for (int e = lane; e < nz; e += 32) zo[e] = S.zk[e];
    if (lane == 0) nz_out[r] = nz;
}

template <int CAP>
struct BtWalkSmem {
    static constexpr int WC = CAP / 16 < 64 ? 64 : CAP / 16;   // chains whose start keys can be sorted here
    unsigned long long wk[WC], wtmp[WC];    // chain-start keys (x of the first anchor) + sort scratch
    unsigned short ps[CAP + 2];             // p[] by anchor index; "none" is the sentinel index CAP, whose own entry is CAP
    unsigned short path[32];                // the nodes of one chase batch
    unsigned tb[CAP / 32 + 1];              // claimed bits (lchain.c: t[]); the sentinel's bit is never set
    unsigned gp[CAP / 32];                  // bit i: f[i] - f[p[i]] > 0 (f[i] > 0 for a root), the sign a one-step walk needs
    unsigned short wpay[WC], wpay2[WC];     // chain ids + sort scratch
    unsigned cnt[256];
    unsigned short start[kBtLevels * kBtRow];
};

// inputs : a, f, p (p = predecessor index inside the read, -1 none), off, zs_scr / nz (from k_bt_sort)
// scratch: v_scr (int per anchor; the chains' anchor indices in emission order), u_scr (u64 per anchor), vs_scr (int per anchor)
// outputs: per read r  n_u[r] (-1 = declined), n_b[r];  b_out[off[r] .. off[r] + n_b) = compacted anchors;  the n_u chains
//          (score << 32 | count) at u_pack[u_pos[r] ..], a packed array shared by the batch (slots handed out by an atomic
//          cursor, so only a short prefix has to be downloaded); a read whose chains do not fit u_cap is declined
template <int CAP>
__global__ void __launch_bounds__(32)
k_bt_walk(const uint4 *__restrict__ a, const int *__restrict__ f, const int *__restrict__ p, const long long *__restrict__ off,
          const int *__restrict__ read_list, int n_list, BtParams bp, const unsigned *__restrict__ zs_scr, const int *__restrict__ nz_in,
          int *__restrict__ v_scr, unsigned long long *__restrict__ u_scr, int *__restrict__ vs_scr, uint4 *__restrict__ b_out,
          int *__restrict__ n_u_out, int *__restrict__ n_b_out, unsigned long long *__restrict__ u_pack, int u_cap, int *__restrict__ u_cur,
          int *__restrict__ u_pos)
{
    extern __shared__ int4 bt_raw[];
    BtWalkSmem<CAP> &S = *reinterpret_cast<BtWalkSmem<CAP> *>(bt_raw);
    constexpr int WC = BtWalkSmem<CAP>::WC;
    constexpr int SENT = CAP;       // "no predecessor"
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= n_list) return;
    const int r = read_list[blockIdx.x];
    const long long o0 = off[r];
    const int n = (int)(off[r + 1] - o0);
    const int nz = nz_in[r];
    if (nz < 0) { if (lane == 0) { n_u_out[r] = -1; n_b_out[r] = 0; } return; }
    if (nz == 0) { if (lane == 0) { n_u_out[r] = 0; n_b_out[r] = 0; } return; }
    const int *fr = f + o0, *pr = p + o0;
    const uint4 *ar = a + o0;
    const unsigned *zs = zs_scr + o0;
    int *vr = v_scr + o0, *vsr = vs_scr + o0;
    unsigned long long *ur = u_scr + o0;
    uint4 *bo = b_out + o0;

    // ---- p by index, sign of the link gains, cleared claim bits ------------------------------------------------------------
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        int pi = -1, fi = 0, fp = 0;
        if (i < n) { pi = pr[i]; fi = fr[i]; }
        if (pi >= 0) fp = fr[pi];
        if (i < n) S.ps[i] = pi < 0 ? (unsigned short)SENT : (unsigned short)pi;
        const unsigned g = __ballot_sync(full, i < n && fi - fp > 0);
        if (lane == 0) { S.gp[i0 >> 5] = g; S.tb[i0 >> 5] = 0; }
    }
    if (lane == 0) { S.ps[SENT] = (unsigned short)SENT; S.tb[CAP / 32] = 0; }
    __syncwarp();

    // ---- chain extraction, best end first (lchain.c:42-72 with mg_chain_bk_end :9-25) --------------------------------------
    int n_v = 0, n_u = 0;
    int k = nz - 1;
    bool nothing_claimed = true;    // true until the first chain is claimed: its walk needs no claim tests at all
    unsigned znext = (k - lane >= 0) ? zs[k - lane] : 0u;   // the 32 ends below k, fetched one round ahead
    int knext = k;
    while (k >= 0) {
        {   // next chain end that is not claimed yet.  Ends whose predecessor is already claimed (or absent) are one-step
            // walks that can only claim themselves (lchain.c:16-22 evaluates p[i] once and stops): a run of them is settled
            // here in parallel, in visiting order; the first end that needs a real walk goes through the general code below.
            const int e = k - lane;
            unsigned z = 0;
            if (knext == k) z = znext;                     // the prefetched group is exactly this one
            else if (e >= 0) z = zs[e];
            knext = k - 32;                                // prefetch the group a full step further down
            znext = (knext - lane >= 0) ? zs[knext - lane] : 0u;
            bool unc = false, simple = false;
            int i0l = 0, n1 = SENT;
            if (e >= 0) {
                i0l = (int)(z & kBtIdxMask);
                unc = ((S.tb[i0l >> 5] >> (i0l & 31)) & 1u) == 0;
                if (unc) {
                    n1 = S.ps[i0l];
                    simple = n1 == SENT || ((S.tb[n1 >> 5] >> (n1 & 31)) & 1u) != 0;
                }
            }
            const unsigned m = __ballot_sync(full, unc);
            if (!m) { k -= 32; continue; }
            const unsigned hard = __ballot_sync(full, unc && !simple);
            const int nfast = hard ? __ffs(hard) - 1 : 32;
            const unsigned fastm = nfast >= 32 ? m : (m & ((1u << nfast) - 1u));
            if (fastm) {
                const bool mine_f = ((fastm >> lane) & 1u) != 0;
                const bool claim = mine_f && ((S.gp[i0l >> 5] >> (i0l & 31)) & 1u) != 0;   // s_1 > 0: cut = p[i0], chain = {i0}
                if (claim) atomicOr(&S.tb[i0l >> 5], 1u << (i0l & 31));
                if (__any_sync(full, claim)) nothing_claimed = false;
                if (bp.min_cnt <= 1) { // single-anchor chains can be accepted: needs the value of s_1
                    const int keyl = (int)(z >> kBtIdxBits);
                    const int s1 = claim ? (n1 == SENT ? keyl : keyl - fr[n1]) : 0;
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
        }
        const unsigned zkk = zs[k];
        const int i0 = (int)(zkk & kBtIdxMask), key = (int)(zkk >> kBtIdxBits);
        // path n_0 = i0, n_1 = p[n_0], ...; node n_j (j >= 1) is "evaluated": s_j = key - f[n_j] (key if n_j is the sentinel).
        // cutj = largest evaluated j whose s_j is a strict new maximum (0 if none): the chain is n_0 .. n_{cutj-1}.
        // The predecessor chase is the serial part: 32 nodes per batch, one shared-memory load per node (the sentinel points
        // at itself, so the chase needs no end test); where the walk ends is found afterwards for all 32 nodes at once.
        int cur = i0, max_s = 0, cutj = 0, cutnode = i0, j0 = 0;
        int mine0 = SENT; // this lane's node of the first batch (enough to mark chains of <= 32 nodes without re-reading)
        for (;;) {
            int nb = 32;
            if (j0 == 0 && !nothing_claimed) {
                // first batch of a later walk: most of them end within a few nodes (at a claimed anchor), so look as we go
                nb = 0;
#pragma unroll 4
                for (int b = 0; b < 32; ++b) {
                    if (lane == 0) S.path[b] = (unsigned short)cur;
                    nb = b + 1;
                    const unsigned tw = S.tb[cur >> 5];             // both loads depend on cur only: issued together
                    const int nxt = S.ps[cur];
                    if (cur == SENT || (((tw >> (cur & 31)) & 1u) != 0 && b >= 1)) break;
                    cur = nxt;
                }
            } else {
#pragma unroll
                for (int b = 0; b < 32; ++b) {
                    if (lane == 0) S.path[b] = (unsigned short)cur;
                    cur = S.ps[cur];
                }
            }
            __syncwarp();
            const int mine = lane < nb ? (int)S.path[lane] : SENT;
            int fmine = 0;
            if (lane < nb && mine != SENT) fmine = fr[mine];
            __syncwarp();
            if (j0 == 0) mine0 = mine;
            const int j = j0 + lane;
            const bool ev = j >= 1 && lane < nb;
            int s = INT32_MIN;
            // the walk stops after evaluating a node that is the root's "predecessor" or already claimed (lchain.c:22)
            const bool stop = ev && (mine == SENT || ((S.tb[mine >> 5] >> (mine & 31)) & 1u) != 0);
            if (ev) s = mine == SENT ? key : key - fmine;
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
                cutnode = __shfl_sync(full, mine, l);
            }
            if (endm) break;
            max_s = max(max_s, __shfl_sync(full, pm, 31));
            j0 += 32;
        }
        const int cnt = cutj;
        if (cnt > 0) { // claim n_0 .. n_{cnt-1}  (stays claimed even if the chain is rejected below, as in the reference)
            nothing_claimed = false;
            if (cnt <= 32) {
                if (lane < cnt) atomicOr(&S.tb[mine0 >> 5], 1u << (mine0 & 31));
            } else {
                __syncwarp();
                for (int q = lane; q < cnt; q += 32) { const int nd = vr[n_v + q]; atomicOr(&S.tb[nd >> 5], 1u << (nd & 31)); }
            }
        }
        __syncwarp();
        const int scv = cutnode == SENT ? key : key - fr[cutnode];
        if (scv >= bp.min_sc && cnt > 0 && cnt >= bp.min_cnt) {
            if (lane == 0) { ur[n_u] = ((unsigned long long)(unsigned)scv << 32) | (unsigned)cnt; vsr[n_u] = n_v; }
            ++n_u;
            n_v += cnt;
        }
        --k;
    }
    __syncwarp();
    if (n_u == 0) { if (lane == 0) { n_u_out[r] = 0; n_b_out[r] = 0; } return; }

    // ---- compact_a (lchain.c:78-111): chains flipped to ascending order, then ordered by the x of their first anchor with
    //      the same unstable sort (w[i].x = b[k].x, payload = chain id) --------------------------------------------------------
    if (n_u > WC) { if (lane == 0) { n_u_out[r] = -1; n_b_out[r] = 0; } return; }   // more chains than the key buffer holds: host path
    int upos = 0;
    if (lane == 0) upos = atomicAdd(u_cur, n_u);
    upos = __shfl_sync(full, upos, 0);
    if (upos + n_u > u_cap) { if (lane == 0) { n_u_out[r] = -1; n_b_out[r] = 0; } return; }   // packed chain buffer full: host path
    unsigned long long *uo = u_pack + upos;
    for (int c = lane; c < n_u; c += 32) {
        const int cntc = (int)(unsigned)ur[c], s0 = vsr[c];
        const uint4 av = ar[vr[s0 + cntc - 1]];
        S.wk[c] = ((unsigned long long)av.y << 32) | av.x;
        S.wpay[c] = (unsigned short)c;
    }
    __syncwarp();
    BtSortScratch sc;
    sc.cnt = S.cnt;
    sc.start = S.start;
    bt_sort<WKey, true>(S.wk, S.wpay, S.wtmp, S.wpay2, n_u, sc, lane);
    int out = 0;
    for (int c = 0; c < n_u; ++c) {
        const int src = S.wpay[c];
        const unsigned long long uv = ur[src];
        const int cntc = (int)(unsigned)uv, s0 = vsr[src];
        if (lane == 0) uo[c] = uv;
        for (int q0 = 0; q0 < cntc; q0 += 128) { // 4 independent gathers per lane in flight
            int idx[4];
            uint4 val[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) { const int q = q0 + t * 32 + lane; idx[t] = q < cntc ? vr[s0 + cntc - 1 - q] : -1; }
#pragma unroll
            for (int t = 0; t < 4; ++t) if (idx[t] >= 0) val[t] = ar[idx[t]];
#pragma unroll
            for (int t = 0; t < 4; ++t) if (idx[t] >= 0) bo[out + q0 + t * 32 + lane] = val[t];
        }
        out += cntc;
    }
    if (lane == 0) { n_u_out[r] = n_u; n_b_out[r] = out; u_pos[r] = upos; }
}

} // namespace mm2gb
