// chain_kernels.cuh -- sm_100a device code of the anchor-chaining path.
//
// What is computed is minimap2-v2.24's mg_lchain_dp forward pass (reference lchain.c:148-207) at
// max-chain-skip = infinity, bit for bit:
//     f[i] = max( q_span(i), max_{st_i <= j < i} f[j] + comput_sc(a[i], a[j]) ),  p[i] = argmax (largest j on ties, -1 if
//     nothing beats q_span(i)),   st_i = max( first j with same rid/strand and x_i - x_j <= max_dist_x , i - max_iter ).
// How it is computed is NOT the reference's GPU layer (gpu/plrange.cu, gpu/plscore.cu: push DP, one barrier per anchor,
// operands in global memory, an integer-log2 score that deviates from lchain.c).  Here:
//   * the whole batch is one flat anchor array; k_range finds st_i by gallop + binary search, marks the independent-unit
//     cuts (st_i == i) and the max_iter-clipped windows, and counts pairs;
//   * units (runs between selected cuts) are scored by one warp each (k_score_units) or by a whole CTA (k_score_long),
//     in tiles of 32 anchors: lane l owns anchor t0+l; predecessors older than the tile are swept with warp-uniform
//     16-byte shared-memory loads (no cross-lane reduction at all), the 32x32 in-tile triangle is pre-scored in
//     registers and resolved with one shuffle per step;
//   * the gap penalty (int)(gap*dd + .5f*mg_log2(dd+1)) is an integer table when chn_pen_skip == 0 (all presets), and
//     is otherwise evaluated with non-contracted fp32 ops in exactly lchain.c's order;
//   * units that contain a clipped window run the exact max_ii state machine of lchain.c:189-205.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mm2gb {

constexpr int kRangeThreads = 256;          // anchors per k_range block
constexpr int kGroupsPerBlock = kRangeThreads / 32;
constexpr int kNeg = -(1 << 30);            // "rejected pair" (lchain.c returns INT32_MIN); never wins a max
constexpr int kLutMax = 1024;               // largest bw served by the integer penalty table (static shared: 2*bw+1 bytes)

// chaining parameters after the adjustments of lchain.c:160-161
struct DevParams {
    int max_iter, max_dist_x, max_dist_y, bw, is_cdna, n_seg;
    int maxd_q;       // min(max_dist_x, max_dist_y): the dq bound when both anchors share a segment id
    int lut_n;        // 2*bw + 1 when the byte penalty table is valid (chn_pen_skip == 0, bw <= kLutMax, max entry <= 255), else 0
    float pen_gap, pen_skip;
};

// device-side counters of one batch
struct Counters {
    int next_unit;          // work-queue cursor of k_score_units
    int n_units;            // written by k_scan
    int multi_sid;          // 1 if some read mixes segment ids or has a zero q_span (forces the general score path)
    int n_exact;            // units scored by the max_ii path
    int next_long;          // work-queue cursor of k_score_long
    int n_long;             // units routed to k_score_long
    int big_cnt[5];         // units of >= 8192 / 4096 / 2048 / 1024 / 512 anchors, queued first (longest-first scheduling)
    int qs_max;             // largest q_span in the batch (bounds the chain scores: f <= unit length * qs_max)
    int ovf_cnt;            // reads the shared-memory chain-extraction kernels handed to the global-memory ones
    int u_cur, b_cur;       // cursors of the packed chain / compacted-anchor outputs of the batch (k_bt_walk)
    int scan_done;          // chunks of k_scan that have finished (the last one combines them)
    int exact_cnt;          // clipped units of >= kExactMin anchors, queued by k_score_units for k_score_exact
    int next_exact;         // work-queue cursor of k_score_exact
    unsigned long long n_pairs;
};

constexpr int kBigMin = 512;                // smallest unit that goes through the longest-first lists
constexpr int kExactMin = 1024;             // clipped units of at least this many anchors get a CTA of their own (k_score_exact)
constexpr int kBigClasses = 5;
__device__ __forceinline__ int big_class(int len) { return len >= 8192 ? 0 : len >= 4096 ? 1 : len >= 2048 ? 2 : len >= 1024 ? 3 : 4; }
// smallest unit of the first `n` classes (the ones k_score_long takes when long_classes = n)
__host__ __device__ __forceinline__ int big_class_min(int n) { return n <= 0 ? INT32_MAX : (8192 >> (n - 1)); }

// Which size classes k_score_long takes for this batch (both score kernels evaluate the same rule on the same counters).
// Units of >= 8192 anchors always (the packed keys of the one-warp kernel do not hold them).  The CTA-cooperative kernel
// spends ~1.5x the instructions per pair of the packed one-warp path but finishes a unit ~2x sooner, so the 4096 and 2048
// classes go to it only while all long units of the batch fit one wave of its CTAs (`wave`; <= 0: always) -- a batch with
// thousands of such units keeps every warp busy in the one-warp kernel anyway.
__device__ __forceinline__ int long_classes_eff(const Counters *ctr, int big_cap, int long_classes, int wave)
{
    if (long_classes <= 1 || wave <= 0) return long_classes;
    int acc = min(ctr->big_cnt[0], big_cap), n = 1;
    for (int c = 1; c < long_classes; ++c) {
        acc += min(ctr->big_cnt[c], big_cap);
        if (acc > wave) break;
        n = c + 1;
    }
    return n;
}

// ---------------------------------------------------------------------------------------------------------------------
// k_expand: packed wire format -> 16-byte anchors in HBM               (replaces the H2D of gpu/plmem.cu:200-236)
//   The upload carries 8 bytes per anchor (low words of x and y) plus one 16-byte record per run of equal high words
//   (csrc/wire.h).  One thread per anchor: blk_run[] brackets the runs a 256-anchor block can touch (almost always one or
//   two), so finding the run is a 0-2 step search, and the anchor is written back as one 16-byte store.  Reads 8 B, writes
//   16 B per anchor: HBM-bound, ~0.12 ms for 31.7 M anchors.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_expand(const uint2 *__restrict__ pk, const int *__restrict__ blk_run, const uint4 *__restrict__ runs, int n_runs, int n_total,
         uint4 *__restrict__ a)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_total) return;
    const uint2 v = __ldg(pk + i);
    int lo = __ldg(blk_run + blockIdx.x), hi = min(__ldg(blk_run + blockIdx.x + 1), n_runs - 1);
    uint4 r = __ldg(runs + lo);
    while (lo < hi) {   // largest run in [lo, hi] that starts at or before i
        const int mid = (lo + hi + 1) >> 1;
        const uint4 rm = __ldg(runs + mid);
        if ((int)rm.x <= i) { lo = mid; r = rm; } else hi = mid - 1;
    }
    a[i] = make_uint4(v.x, r.y, v.y, r.z);
}

// ---------------------------------------------------------------------------------------------------------------------
// k_range: window start, cuts, clipped windows, pair count            (replaces gpu/plrange.cu:38-76)
//   one thread per anchor of the flat batch; a block first finds the read that holds its first anchor.
// ---------------------------------------------------------------------------------------------------------------------
// window start of anchor g by gallop + binary search in global memory (any window length): first index in [lo0, g] whose
// x is >= lower
__device__ __forceinline__ int window_start_global(const ulonglong2 *__restrict__ a, int g, int lo0, unsigned long long lower)
{
    int hi = g, bad = lo0 - 1, step = 1;
    for (;;) { // gallop back from g
        int probe = hi - step;
        if (probe <= lo0) {
            if (hi > lo0) { if (a[lo0].x >= lower) hi = lo0; else bad = lo0; }
            break;
        }
        if (a[probe].x >= lower) { hi = probe; step <<= 1; }
        else { bad = probe; break; }
    }
    while (hi - bad > 1) {
        int mid = (hi + bad) >> 1;
        if (a[mid].x >= lower) hi = mid; else bad = mid;
    }
    return hi;
}

constexpr int kRangeHist = 1024;            // most x history (anchors before the block) staged in shared memory

// block_read[b] = the read that owns the first anchor of k_range block b (largest r with off[r] <= 256 b): one thread per
// block here, so that k_range itself starts without a 14-step dependent binary search in front of every block
__global__ void __launch_bounds__(256)
k_block_reads(const long long *__restrict__ off, int n_reads, int n_blocks, int *__restrict__ block_read)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const long long g0 = (long long)b * kRangeThreads;
    int lo = 0, hi = n_reads;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= g0) lo = mid; else hi = mid;
    }
    block_read[b] = lo;
}

// what a k_range thread does once the x history of its block is staged: XS(j) = x of anchor j for j in [max(h0, 0), g0 + 256)
template <class XS>
__device__ __forceinline__ void range_tail(const ulonglong2 *__restrict__ a, const long long *__restrict__ off, int r0, int n_total,
                                           const DevParams &prm, const ulonglong2 &ai, unsigned long long yprev, int g0, int h0, XS xs,
                                           int *__restrict__ st, unsigned *__restrict__ selmask, unsigned *__restrict__ clipmask,
                                           int *__restrict__ block_cnt, unsigned long long *__restrict__ block_pairs, Counters *__restrict__ ctr,
                                           int *s_cnt, unsigned long long *s_pairs)
{
    const int g = g0 + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool act = g < n_total;
    const unsigned mdx = (unsigned)prm.max_dist_x;
    bool cut = false, clipped = false, rstart = false;
    int npair = 0, qspan = 0;
    if (act) {
        int r = r0;
        while (off[r + 1] <= (long long)g) ++r;
        const int rs = (int)off[r];
        const unsigned long long xi = ai.x;
        const unsigned lo32 = (unsigned)xi;
        // lchain.c:172: j is in the window iff same rid/strand and x_i <= x_j + max_dist_x  <=>  x_j >= lower
        const unsigned long long lower = (xi & 0xffffffff00000000ULL) | (unsigned long long)(lo32 > mdx ? lo32 - mdx : 0u);
        const long long lo_ll = (long long)g - (long long)prm.max_iter;   // lchain.c:173
        const int lo0 = lo_ll > (long long)rs ? (int)lo_ll : rs;
        int hi;
        const int lo_s = max(lo0, max(h0, 0));   // oldest candidate available in shared memory
        if (g > lo_s && xs(g - 1) < lower) {
            hi = g;     // the previous anchor is already out of reach (x is sorted): empty window.  An isolated hit -- most of the
                        // anchors of a large reference's hit mix -- costs one probe instead of a binary search
        } else if (lo_s > lo0 && xs(lo_s) >= lower) {
            hi = window_start_global(a, g, lo0, lower);   // window reaches beyond the staged history (rare)
        } else {
            // first index in [lo_s, g] with x >= lower; x[g] itself qualifies
            int bad = lo_s - 1;
            hi = g;
            while (hi - bad > 1) {
                const int mid = (hi + bad) >> 1;
                if (xs(mid) >= lower) hi = mid; else bad = mid;
            }
        }
        st[g] = hi;
        npair = g - hi;
        cut = hi == g;
        rstart = g == rs;
        clipped = lo0 > rs && hi == lo0 && a[lo0 - 1].x >= lower;
        // lchain.c:115-116 compares the segment ids of the two anchors; one id per read is the common case
        qspan = (int)((ai.y >> 32) & 0xff);
        // (the table path also assumes q_span > 0, which every real seed satisfies)
        if ((g > rs && (unsigned)((ai.y >> 48) & 0xff) != (unsigned)((yprev >> 48) & 0xff)) || ((ai.y >> 32) & 0xff) == 0) atomicOr(&ctr->multi_sid, 1);
    }
    const unsigned cutm = __ballot_sync(0xffffffffu, cut);
    const unsigned rsm = __ballot_sync(0xffffffffu, rstart);
    const unsigned clm = __ballot_sync(0xffffffffu, clipped);
    // unit boundaries: the first cut of every 32-anchor group, plus every read start (so no unit spans two reads)
    const unsigned sel = (cutm & (0u - cutm)) | rsm;
    // pairs of this block: summed by k_scan, not by a million same-address atomics.  npair < 2^31, so the warp sum is
    // taken in two 16-bit halves to stay inside 32-bit redux
    const unsigned long long psum = (unsigned long long)__reduce_add_sync(0xffffffffu, (unsigned)npair & 0xffffu) +
                                    ((unsigned long long)__reduce_add_sync(0xffffffffu, (unsigned)npair >> 16) << 16);
    const int qmax = __reduce_max_sync(0xffffffffu, qspan);
    if (lane == 0) {
        if (qmax > ctr->qs_max) atomicMax(&ctr->qs_max, qmax);
        const int grp = g >> 5;
        if (g0 + wid * 32 < n_total) { selmask[grp] = sel; clipmask[grp] = clm; }
        s_cnt[wid] = __popc(sel);
        s_pairs[wid] = psum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int c = 0;
        unsigned long long ps = 0;
#pragma unroll
        for (int w = 0; w < kGroupsPerBlock; ++w) c += s_cnt[w], ps += s_pairs[w];
        block_cnt[blockIdx.x] = c;
        block_pairs[blockIdx.x] = ps;
    }
}

__global__ void __launch_bounds__(kRangeThreads)
k_range(const ulonglong2 *__restrict__ a, const long long *__restrict__ off, const int *__restrict__ block_read, int n_total, DevParams prm,
        int *__restrict__ st, unsigned *__restrict__ selmask, unsigned *__restrict__ clipmask, int *__restrict__ block_cnt,
        unsigned long long *__restrict__ block_pairs, Counters *__restrict__ ctr)
{
    // x of the anchors [g0 - hist, g0 + 256): the block's own anchors plus as much history as the first anchor's window
    // needs (grown 256 at a time, coalesced), so the per-anchor binary search runs on shared memory
    __shared__ unsigned long long s_x[kRangeHist + kRangeThreads];
    __shared__ int s_more;
    __shared__ int s_cnt[kGroupsPerBlock];
    __shared__ unsigned long long s_pairs[kGroupsPerBlock];
    const int g0 = blockIdx.x * kRangeThreads;
    const int g = g0 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int r0 = block_read[blockIdx.x];
    const bool act = g < n_total;
    ulonglong2 ai = make_ulonglong2(0, 0);
    if (act) ai = a[g];
    // y of the previous anchor (lane 0 fetches it, the others get it by shuffle): segment ids are compared between neighbours
    unsigned long long yprev = (lane == 0 && act && g > 0) ? a[g - 1].y : 0ULL;
    {
        const unsigned long long yup = __shfl_up_sync(0xffffffffu, ai.y, 1);
        if (lane) yprev = yup;
    }
    // the first chunk of history and the start of the first anchor's read are always needed: their loads are issued here,
    // together with the block's own anchors, not one latency later (the kernel is bound by dependent-load latency)
    const int gpre = g0 - kRangeThreads + (int)threadIdx.x;
    const unsigned long long xpre = gpre >= 0 ? a[gpre].x : 0ULL;
    const long long rs0 = off[r0];
    s_x[kRangeHist + threadIdx.x] = ai.x;
    __syncthreads();
    // history: stop once the oldest staged x is outside the first anchor's window (x sorted => outside everyone's), or the
    // read of the first anchor starts inside the staged part (a later read starts later still)
    const unsigned mdx = (unsigned)prm.max_dist_x;
    int hist = 0;
    {
        const unsigned long long x0 = s_x[kRangeHist];
        const unsigned x0lo = (unsigned)x0;
        const unsigned long long lower0 = (x0 & 0xffffffff00000000ULL) | (unsigned long long)(x0lo > mdx ? x0lo - mdx : 0u);
        while (hist < kRangeHist) {
            const int base = g0 - hist - kRangeThreads;    // next chunk [base, base + 256)
            const int gi = base + (int)threadIdx.x;
            const unsigned long long xv = hist == 0 ? xpre : (gi >= 0 ? a[gi].x : 0ULL);
            s_x[kRangeHist - hist - kRangeThreads + threadIdx.x] = xv;
            if (threadIdx.x == 0) s_more = (gi > rs0 && xv >= lower0) ? 1 : 0;   // oldest of the chunk still inside the window
            hist += kRangeThreads;
            __syncthreads();
            const int more = s_more;
            __syncthreads();
            if (!more) break;
        }
    }
    const int h0 = g0 - hist;   // s_x[kRangeHist + (j - g0)] is valid for j in [max(h0, 0), g0 + 256)
    const unsigned long long *xb = s_x + (kRangeHist - g0);   // xb[j] = x of anchor j
    range_tail(a, off, r0, n_total, prm, ai, yprev, g0, h0, [xb](int j) { return xb[j]; }, st, selmask, clipmask, block_cnt, block_pairs, ctr,
               s_cnt, s_pairs);
}

// ---- the same kernel with the staging done by the bulk-copy engine (TMA, 1-D): thread 0 issues ONE cp.async.bulk for the block's own
//      256 anchors plus the first 256 of history (8 KB of whole 16-byte anchors, contiguous in the flat array), every thread waits
//      on the mbarrier, and further history comes 4 KB at a time the same way.  No LDG -> STS hop through registers and a single
//      wait per chunk instead of two block barriers; the price is 16 instead of 8 bytes of shared memory per staged anchor and
//      a binary search over 16-byte strides.  Selected with MM2GB_RANGE_TMA=1 (measured against the plain kernel in profiles/).
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

__global__ void __launch_bounds__(kRangeThreads)
k_range_tma(const ulonglong2 *__restrict__ a, const long long *__restrict__ off, const int *__restrict__ block_read, int n_total, DevParams prm,
            int *__restrict__ st, unsigned *__restrict__ selmask, unsigned *__restrict__ clipmask, int *__restrict__ block_cnt,
            unsigned long long *__restrict__ block_pairs, Counters *__restrict__ ctr)
{
    __shared__ __align__(128) ulonglong2 s_a[kRangeHist + kRangeThreads];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int s_cnt[kGroupsPerBlock];
    __shared__ unsigned long long s_pairs[kGroupsPerBlock];
    const int g0 = blockIdx.x * kRangeThreads;
    const int g = g0 + threadIdx.x;
    const int r0 = block_read[blockIdx.x];
    const bool act = g < n_total;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    unsigned phase = 0;
    // own anchors + the first chunk of history: [max(0, g0 - 256), min(n_total, g0 + 256))
    {
        const int lo = max(0, g0 - kRangeThreads), hi = min(n_total, g0 + kRangeThreads);
        if (threadIdx.x == 0) {
            const unsigned bytes = (unsigned)(hi - lo) * 16u;
            mbar_expect_tx(&s_bar, bytes);
            bulk_g2s(&s_a[kRangeHist + (lo - g0)], a + lo, bytes, &s_bar);
        }
    }
    const long long rs0 = off[r0];
    mbar_wait(&s_bar, phase);
    phase ^= 1u;
    const ulonglong2 *ab = s_a + (kRangeHist - g0);   // ab[j] = anchor j
    ulonglong2 ai = make_ulonglong2(0, 0);
    if (act) ai = ab[g];
    const unsigned long long yprev = (act && g > 0) ? ab[g - 1].y : 0ULL;
    const unsigned mdx = (unsigned)prm.max_dist_x;
    int hist = kRangeThreads;
    {
        const unsigned long long x0 = ab[g0].x;
        const unsigned x0lo = (unsigned)x0;
        const unsigned long long lower0 = (x0 & 0xffffffff00000000ULL) | (unsigned long long)(x0lo > mdx ? x0lo - mdx : 0u);
        for (;;) {
            // the oldest staged anchor (g0 - hist) still inside the first anchor's window and read?  every thread decides alike
            const int oldest = g0 - hist;
            if (!(oldest > rs0 && ab[oldest].x >= lower0) || hist >= kRangeHist) break;
            const int lo = max(0, oldest - kRangeThreads);
            __syncthreads();    // everyone has passed the previous wait before the barrier is armed again
            if (threadIdx.x == 0) {
                const unsigned bytes = (unsigned)(oldest - lo) * 16u;
                mbar_expect_tx(&s_bar, bytes);
                bulk_g2s(&s_a[kRangeHist + (lo - g0)], a + lo, bytes, &s_bar);
            }
            mbar_wait(&s_bar, phase);
            phase ^= 1u;
            hist += kRangeThreads;
        }
    }
    const int h0 = g0 - hist;
    range_tail(a, off, r0, n_total, prm, ai, yprev, g0, h0, [ab](int j) { return ab[j].x; }, st, selmask, clipmask, block_cnt, block_pairs, ctr,
               s_cnt, s_pairs);
}

// k_scan: exclusive prefix of the per-block unit counts, total -> ctr->n_units.  One CTA per chunk of 8192 entries (8 consecutive
//         entries per thread): block_base[i] = prefix inside the chunk; the CTA that finishes last scans the chunk totals into
//         chunk_base[] (k_units adds the two) and sums the pair counts.  (A single CTA walking all ~120 k entries of a
//         30 M-anchor batch took 125 us.)
constexpr int kScanPer = 8;
constexpr int kScanChunk = 1024 * kScanPer;
__global__ void __launch_bounds__(1024)
k_scan(const int *__restrict__ block_cnt, const unsigned long long *__restrict__ block_pairs, int n_blocks, int *__restrict__ block_base,
       int *chunk_tot, unsigned long long *chunk_pairs, int *__restrict__ chunk_base, Counters *ctr)
{
    __shared__ int s_warp[32];
    __shared__ unsigned long long s_pw[32];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i0 = blockIdx.x * kScanChunk + threadIdx.x * kScanPer;
    int v[kScanPer];
    unsigned long long pv[kScanPer];
    // all 16 loads first (clamped index, no branch around them), so that they are in flight together
#pragma unroll
    for (int q = 0; q < kScanPer; ++q) {
        const int ic = min(i0 + q, n_blocks - 1);
        v[q] = __ldg(block_cnt + ic);
        pv[q] = __ldg(block_pairs + ic);
    }
    int x = 0;
    unsigned long long pairs = 0;
#pragma unroll
    for (int q = 0; q < kScanPer; ++q) {
        if (i0 + q >= n_blocks) { v[q] = 0; pv[q] = 0; }
        pairs += pv[q];
        x += v[q];
    }
    const int mine = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) pairs += __shfl_xor_sync(0xffffffffu, pairs, d);
    if (lane == 31) s_warp[wid] = x;
    if (lane == 0) s_pw[wid] = pairs;
    __syncthreads();
    if (wid == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += y;
        }
        s_warp[lane] = w; // inclusive
    }
    __syncthreads();
    const int incl = x + (wid ? s_warp[wid - 1] : 0);
    int run = incl - mine;
#pragma unroll
    for (int q = 0; q < kScanPer; ++q) {
        if (i0 + q < n_blocks) block_base[i0 + q] = run;
        run += v[q];
    }
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 32; ++w) t += s_pw[w];
        chunk_tot[blockIdx.x] = s_warp[31];
        chunk_pairs[blockIdx.x] = t;
        __threadfence();
        s_last = atomicAdd(&ctr->scan_done, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last || wid != 0) return;
    // the last chunk to finish: prefix over the chunk totals (a few hundred at most), total pair count = sum_i (i - st_i)
    // (the reference's n_iter, lchain.c:177)
    __threadfence();
    int base = 0;
    unsigned long long tp = 0;
    for (int c0 = 0; c0 < (int)gridDim.x; c0 += 32) {
        const int c = c0 + lane;
        const int tv = c < (int)gridDim.x ? __ldcg(chunk_tot + c) : 0;
        if (c < (int)gridDim.x) tp += __ldcg(chunk_pairs + c);
        int in = tv;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, in, d);
            if (lane >= d) in += y;
        }
        if (c < (int)gridDim.x) chunk_base[c] = base + in - tv;
        base += __shfl_sync(0xffffffffu, in, 31);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) tp += __shfl_xor_sync(0xffffffffu, tp, d);
    if (lane == 0) { ctr->n_pairs = tp; ctr->n_units = base; }
}

// k_units: ordered scatter of the selected cuts -> unit_start[k], unit_rbase[k] (first anchor of the owning read),
//          sentinel unit_start[n_units] = n_total.  One thread per 32-anchor group.
__global__ void __launch_bounds__(256)
k_units(const unsigned *__restrict__ selmask, const int *__restrict__ block_base, const int *__restrict__ chunk_base,
        const long long *__restrict__ off, int n_reads, int n_total, int n_groups, int *__restrict__ unit_start,
        int *__restrict__ unit_rbase, unsigned *__restrict__ unit_clip, const Counters *__restrict__ ctr)
{
    const int grp = blockIdx.x * blockDim.x + threadIdx.x;
    if (grp == 0) unit_start[ctr->n_units] = n_total;
    if (grp >= n_groups) return;
    unsigned m = selmask[grp];
    if (!m) return;
    const int b = grp / kGroupsPerBlock;
    int k = block_base[b] + chunk_base[b / kScanChunk];
    for (int g2 = b * kGroupsPerBlock; g2 < grp; ++g2) k += __popc(selmask[g2]);
    while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1;
        const int g = grp * 32 + bit;
        int lo = 0, hi = n_reads;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (off[mid] <= (long long)g) lo = mid; else hi = mid;
        }
        unit_start[k] = g;
        unit_rbase[k] = (int)off[lo];
        unit_clip[k] = 0;
        ++k;
    }
}

// k_unit_clip: unit_clip[k] = 1 for every unit that holds a window clipped by max_iter (those need the max_ii fallback of
//              lchain.c:189-205).  One thread per 32-anchor group of clipmask; a group without clipped windows -- every group of
//              ordinary reads -- costs one load, so the score kernels read one flag per unit instead of scanning the unit's mask.
__global__ void __launch_bounds__(256)
k_unit_clip(const unsigned *__restrict__ clipmask, const int *__restrict__ unit_start, int n_groups, unsigned *__restrict__ unit_clip,
            const Counters *__restrict__ ctr)
{
    const int grp = blockIdx.x * blockDim.x + threadIdx.x;
    if (grp >= n_groups) return;
    unsigned m = clipmask[grp];
    if (!m) return;
    const int n_units = ctr->n_units;
    int k = -1;
    while (m) {
        const int i = grp * 32 + __ffs(m) - 1;
        m &= m - 1;
        if (k < 0 || i >= unit_start[k + 1]) {      // the unit of anchor i: largest k with unit_start[k] <= i
            int lo = 0, hi = n_units;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (unit_start[mid] <= i) lo = mid; else hi = mid;
            }
            k = lo;
            unit_clip[k] = 1u;
        }
    }
}

// k_order: units of >= kBigMin anchors are listed per size class so that the score kernel starts the longest units first
// (a 5000-anchor unit is ~1 ms of one warp's time: started last it would be the tail of the launch).
// Clipped units of >= kExactMin anchors are queued for k_score_exact (behind the lists in big_order) and appear in no list.
__global__ void __launch_bounds__(256)
k_order(const int *__restrict__ unit_start, const unsigned *__restrict__ unit_clip, int *__restrict__ big_order, int big_cap,
        Counters *__restrict__ ctr, int fast)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ctr->n_units) return;
    const int len = unit_start[k + 1] - unit_start[k];
    if (len < kBigMin) return;
    // (only when the table path scores the batch: the general path -- float penalties, mixed segment ids -- keeps them in the lists)
    if (fast && ctr->multi_sid == 0 && len >= kExactMin && unit_clip[k]) {
        big_order[kBigClasses * big_cap + atomicAdd(&ctr->exact_cnt, 1)] = k;
        atomicAdd(&ctr->n_exact, 1);
        return;
    }
    const int c = big_class(len);
    const int pos = atomicAdd(&ctr->big_cnt[c], 1);
    if (pos < big_cap) big_order[c * big_cap + pos] = k;
}

// ---------------------------------------------------------------------------------------------------------------------
// pair score                                                              (reference lchain.c:113-138, mmpriv.h:118-126)
// ---------------------------------------------------------------------------------------------------------------------

// shared-memory / register record of a scored anchor: x, y = low 32 bits of mm128_t.x / .y, f = chain score,
// q = q_span | seg_id << 8
struct __align__(16) Rec { int x, y, f, q; };

__device__ __forceinline__ float mg_log2_dev(float x) // mmpriv.h:118-126, no FMA contraction
{
    unsigned z = __float_as_uint(x);
    float r = (float)((int)((z >> 23) & 255u) - 128);
    z &= ~(255u << 23);
    z += 127u << 23;
    const float zf = __uint_as_float(z);
    const float poly = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(-0.34484843f, zf), 2.02466578f), zf), 0.67487759f);
    return __fadd_rn(r, poly);
}

// every branch of comput_sc; returns kNeg for a rejected pair
__device__ __forceinline__ int pair_general(int xi, int yi, int sidi, const Rec &r, const DevParams &P)
{
    const int dq = yi - r.y;
    const int sidj = (r.q >> 8) & 0xff, qs = r.q & 0xff;
    const bool same = sidi == sidj;
    if (dq <= 0 || dq > P.max_dist_x) return kNeg;                                   // lchain.c:118
    const int dr = xi - r.x;                                                          // :119
    if (same && (dr == 0 || dq > P.max_dist_y)) return kNeg;                          // :120
    const int dd = dr > dq ? dr - dq : dq - dr;                                       // :121
    if (same && dd > P.bw) return kNeg;                                               // :122
    if (P.n_seg > 1 && !P.is_cdna && same && dr > P.max_dist_y) return kNeg;          // :123
    const int dg = dr < dq ? dr : dq;                                                 // :124
    int sc = qs < dg ? qs : dg;                                                       // :126
    if (dd || dg > qs) {                                                              // :127
        const float lin = __fadd_rn(__fmul_rn(P.pen_gap, (float)dd), __fmul_rn(P.pen_skip, (float)dg));
        const float lg = dd >= 1 ? mg_log2_dev((float)(dd + 1)) : 0.0f;
        if (P.is_cdna || !same) {                                                     // :131
            if (!same && dr == 0) ++sc;
            else if (dr > dq || !same) sc -= __float2int_rz(lin < lg ? lin : lg);
            else sc -= __float2int_rz(__fadd_rn(lin, __fmul_rn(.5f, lg)));
        } else sc -= __float2int_rz(__fadd_rn(lin, __fmul_rn(.5f, lg)));              // :135
    }
    return sc;
}

// single-segment, non-cDNA, chn_pen_skip == 0, every q_span > 0: the penalty is a byte table in shared memory, stored
// symmetrically around bw (entry k holds pen(|k - bw|), k = dr - dq + bw in [0, 2bw]) so neither |.| nor an index clamp is
// needed: one unsigned compare gives the band test (lchain.c:121-122) and the load predicate.  lut_s = 32-bit shared address
// of entry 0.  Returns validity, score in sc.  `pen` is caller-owned scratch, rewritten only inside the band (a rejected
// pair never uses it).  dr >= 0 always (x-sorted, same rid/strand inside a window), so min(dr,dq,q_span) > 0 <=> dq > 0 &&
// dr != 0 (lchain.c:118,120).  In FAST kernels Rec.q holds q_span only.
__device__ __forceinline__ bool pair_fast(int xi, int yi, const Rec &r, int maxd_q, unsigned bw, unsigned lut_s, int &pen, int &sc)
{
    const int dr = xi - r.x, dq = yi - r.y;
    const unsigned tb = (unsigned)dr - (unsigned)dq + bw;
    const bool band = tb <= 2u * bw;
    if (band) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(pen) : "r"(lut_s + tb));
    const int m = min(min(dr, dq), r.q);
    sc = m - pen;
    bool ok = band;
    ok = ok && m > 0;
    ok = ok && dq <= maxd_q;
    return ok;
}

template <bool FAST>
__device__ __forceinline__ bool pair_score(int xi, int yi, int sidi, const Rec &r, const DevParams &P, unsigned lut_s, int &pen, int &sc)
{
    if (FAST) return pair_fast(xi, yi, r, P.maxd_q, (unsigned)P.bw, lut_s, pen, sc);
    sc = pair_general(xi, yi, sidi, r, P);
    return sc != kNeg;
}

template <bool FAST>
__device__ __forceinline__ Rec make_rec(const uint4 &v, int f)
{
    Rec r;
    r.x = (int)v.x; r.y = (int)v.z; r.f = f;
    r.q = FAST ? (int)(v.w & 0xffu) : (int)((v.w & 0xffu) | (((v.w >> 16) & 0xffu) << 8));
    return r;
}

// predecessor record straight from global memory (anchors are read-only; f[] was written by this CTA)
template <bool FAST>
__device__ __forceinline__ Rec fetch_global(const uint4 *__restrict__ a, const int *f, int j)
{
    return make_rec<FAST>(__ldg(a + j), f[j]);
}

// ---------------------------------------------------------------------------------------------------------------------
// exact path for units with a max_iter-clipped window: lchain.c:169-207 row by row, including the max_ii fallback
// (lchain.c:189-205).  One warp, lanes stride over the window.
// ---------------------------------------------------------------------------------------------------------------------
template <bool FAST>
__device__ void score_unit_exact(const uint4 *__restrict__ a, const int *__restrict__ st, int *f, int *__restrict__ p,
                                 int u0, int u1, int rbase, const DevParams &P, unsigned lut_s, int lane)
{
    const unsigned full = 0xffffffffu;
    int pen = 0;
    const unsigned long long *ax = reinterpret_cast<const unsigned long long *>(a);
    const unsigned long long mdx = (unsigned long long)(long long)P.max_dist_x;
    int max_ii = -1;
    for (int i = u0; i < u1; ++i) {
        const uint4 ai = __ldg(a + i);
        const int xi = (int)ai.x, yi = (int)ai.z, qsi = (int)(ai.w & 0xffu), sidi = (int)((ai.w >> 16) & 0xffu);
        const unsigned long long x64 = ax[2 * (size_t)i];
        const int sti = st[i];
        int bv = kNeg, bj = -1;
        for (int j = sti + lane; j < i; j += 32) {
            const Rec r = fetch_global<FAST>(a, f, j);
            int s;
            if (pair_score<FAST>(xi, yi, sidi, r, P, lut_s, pen, s) && s + r.f >= bv) bv = s + r.f, bj = j;
        }
        const int m = __reduce_max_sync(full, bv);
        const int jm = __reduce_max_sync(full, bv == m ? bj : -1);
        int best = qsi, arg = -1;
        if (m != kNeg && m > best) best = m, arg = jm;                                 // strict '>' against the init value
        if (max_ii < 0 || x64 - ax[2 * (size_t)max_ii] > mdx) {                        // lchain.c:189-194
            int fv = INT32_MIN, fj = -1;
            for (int j = sti + lane; j < i; j += 32) {
                const int fjv = f[j];
                if (fjv >= fv) fv = fjv, fj = j;
            }
            const int fm = __reduce_max_sync(full, fv);
            max_ii = __reduce_max_sync(full, (fv == fm && fj >= 0) ? fj : -1);
        }
        if (max_ii >= 0 && max_ii < sti - 1) {                                         // :196 (end_j = st - 1 at max_skip = inf)
            const Rec r = fetch_global<FAST>(a, f, max_ii);
            int s;
            if (pair_score<FAST>(xi, yi, sidi, r, P, lut_s, pen, s) && best < s + r.f) best = s + r.f, arg = max_ii;
        }
        if (lane == 0) { f[i] = best; p[i] = arg < 0 ? -1 : arg - rbase; }
        __syncwarp();
        if (max_ii < 0 || (x64 - ax[2 * (size_t)max_ii] <= mdx && f[max_ii] < best)) max_ii = i; // :204
    }
}

// does [u0,u1) contain an anchor whose window was clipped by max_iter?
__device__ __forceinline__ bool unit_has_clip(const unsigned *__restrict__ clipmask, int u0, int u1, int lane)
{
    const int g0 = u0 >> 5, g1 = (u1 - 1) >> 5;
    bool any = false;
    for (int g = g0 + lane; g <= g1; g += 32) {
        unsigned m = clipmask[g];
        if (g == g0) m &= 0xffffffffu << (u0 & 31);
        if (g == g1 && ((u1 & 31) != 0)) m &= 0xffffffffu >> (32 - (u1 & 31));
        any |= m != 0;
    }
    return __any_sync(0xffffffffu, any);
}

// ---------------------------------------------------------------------------------------------------------------------
// tiled scoring of one unit by one warp
//   ring: this warp's R-entry shared-memory window of Rec, indexed (i - u0) & (R-1).
// ---------------------------------------------------------------------------------------------------------------------
template <int R, bool FAST>
__device__ void score_unit_tiled(const uint4 *__restrict__ a, const int *__restrict__ st, int *f, int *__restrict__ p,
                                 int u0, int u1, int rbase, const DevParams &P, unsigned lut_s, Rec *ring, int lane)
{
    const unsigned full = 0xffffffffu;
    int pen = 0;
    for (int t0 = u0; t0 < u1; t0 += 32) {
        const int i = t0 + lane;
        const bool act = i < u1;
        uint4 ai = make_uint4(0, 0, 0, 0);
        int sti = INT32_MAX; // inactive lanes: empty window
        if (act) { ai = __ldg(a + i); sti = st[i]; }
        const int xi = (int)ai.x, yi = (int)ai.z, qsi = (int)(ai.w & 0xffu), sidi = (int)((ai.w >> 16) & 0xffu);
        const int wmin = __shfl_sync(full, sti, 0);                       // st is non-decreasing, lane 0 is always active
        const int wfull = min(t0, __reduce_max_sync(full, act ? sti : 0)); // from here on every active lane's window is open
        const int jring = max(u0, t0 - R);                                 // oldest predecessor still in the ring
        int thr = qsi + 1, bj = -1; // a candidate wins iff val >= thr: '>' against q_span(i), '>=' afterwards (ascending j)

        // phase A.0: predecessors that have left the ring (window longer than R): global / L1
        for (int j = wmin; j < min(jring, t0); ++j) {
            const Rec r = fetch_global<FAST>(a, f, j);
            int s;
            const bool ok = pair_score<FAST>(xi, yi, sidi, r, P, lut_s, pen, s);
            const int val = s + r.f;
            if (ok && j >= sti && val >= thr) thr = val, bj = j;
        }
        // phase A.1 / A.2: the ring, walked in address-contiguous runs (a run never wraps, so loads are [base + imm]).
        // A.1 = windows still opening (needs the j >= st_i test), A.2 = every active lane's window is open.
        int j = max(wmin, jring);
        while (j < wfull) {
            const int pos = (j - u0) & (R - 1);
            const int len = min(wfull - j, R - pos);
            const Rec *rp = ring + pos;
#pragma unroll 4
            for (int k = 0; k < len; ++k) {
                const Rec r = rp[k];
                int s;
                const bool ok = pair_score<FAST>(xi, yi, sidi, r, P, lut_s, pen, s);
                const int val = s + r.f;
                if (ok && j + k >= sti && val >= thr) thr = val, bj = j + k;
            }
            j += len;
        }
        while (j < t0) {
            const int pos = (j - u0) & (R - 1);
            const int len = min(t0 - j, R - pos);
            const Rec *rp = ring + pos;
#pragma unroll 4
            for (int k = 0; k < len; ++k) {
                const Rec r = rp[k];
                int s;
                const bool ok = pair_score<FAST>(xi, yi, sidi, r, P, lut_s, pen, s);
                const int val = s + r.f;
                if (ok && val >= thr) thr = val, bj = j + k;
            }
            j += len;
        }
        __syncwarp();
        // publish this tile's static fields (overwrites ring entries older than t0 - R)
        Rec *tile = ring + ((t0 - u0) & (R - 1)); // 32-aligned: a tile never wraps
        if (act) tile[lane] = make_rec<FAST>(ai, 0);
        __syncwarp();

        // phase B: in-tile triangle.  static part first (independent of f), then the serial chain
        const int nact = min(32, u1 - t0);
        int fcur = bj >= 0 ? thr : qsi;
        // the tile slots s that are predecessors of this lane's anchor (s < lane and t0 + s >= st_i), one bit test per slot below
        const int vlo = sti - t0;
        const unsigned vm = vlo >= lane ? 0u : (((1u << lane) - 1u) & ~((1u << max(vlo, 0)) - 1u));
#pragma unroll
        for (int h = 0; h < 2; ++h) { // two halves keep the pre-scored set at 16 registers
            if (h == 1 && nact <= 17) break; // warp-uniform: a short last tile has no candidates s >= 16
            int w[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int s = h * 16 + q;
                if (s < 31) {
                    int sc;
                    const bool ok = pair_score<FAST>(xi, yi, sidi, tile[s], P, lut_s, pen, sc);
                    w[q] = (ok && ((vm >> s) & 1u)) ? sc : kNeg;
                }
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int s = h * 16 + q;
                if (s < 31) {
                    const int fs = __shfl_sync(full, fcur, s);
                    const int val = fs + w[q];                  // kNeg + f stays far below any threshold
                    if (val >= thr) thr = val, fcur = val, bj = t0 + s;
                }
            }
        }
        if (act) {
            f[i] = fcur;
            p[i] = bj < 0 ? -1 : bj - rbase;
            tile[lane].f = fcur;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_score_units: persistent grid, one warp per unit taken from a device-side queue
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kScoreWarps = 4;


// ---------------------------------------------------------------------------------------------------------------------
// packed-key variant of the tiled scoring (table path only).  For a unit of at most 8192 anchors whose scores fit 18 bits
// (length * max q_span < 2^18 -- every 10-100 kb ONT read), score and predecessor travel in ONE register:
//     key = f << 13 | (j - u0)
// A candidate is key = F_j + (sc << 13) with F_j = f_j << 13 | slot_j, and the running best is thr = max(thr, key): the max
// picks the best score and, among equal scores, the largest j (lchain.c:174-181 scans j downward with a strict '>').  thr
// starts at q_span << 13 | 8191, so a candidate that merely equals q_span(i) loses (max_j stays -1).
//
// The ring record is laid out for the common case:  e = x - y (the diagonal),  g = F_j + (q_span_j << 13),  y,  q.
// With D_i = x_i - y_i + bw, one subtraction gives tb = D_i - e_j = dr - dq + bw, the band test (lchain.c:121-122) is the
// unsigned compare tb <= 2bw and tb is also the index into the symmetric penalty table.  Predecessors are then split, per
// tile and warp-uniformly, by how far back they are (x is sorted inside a window):
//   MID   dr > bw + max q_span for every lane of the tile  =>  inside the band dq >= dr - bw > q_span_j, so
//         min(dr, dq, q_span_j) = q_span_j > 0 and sc = q_span_j - pen: already folded into g.  If also dr <= maxd_q - bw for
//         every lane, dq <= dr + bw <= maxd_q holds too and the whole pair is  sub, setp, ld.u8, mad, max  (+ one 8-byte
//         warp-uniform record load).
//   FAR   as MID but dr may exceed maxd_q - bw (or the window of some lanes is still opening): adds the dq <= maxd_q test
//         (and j >= st_i while the windows open).
//   GEN   the near predecessors (dr <= bw + max q_span): the full min(dr, dq, q_span) of lchain.c:124-126.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kSlotBits = 13, kSlotMask = (1 << kSlotBits) - 1;
constexpr int kNegKey = -(1 << 30);

struct __align__(16) RecP { int e, g, y, q; };

enum { MODE_MID = 0, MODE_FAR = 1, MODE_GEN = 2 };

// One candidate.  The validity tests stay ONE predicate chain and the update is a single predicated max.
// D = x_i - y_i + bw; jk / sti only used when CHECK (window still opening for some lane).
template <int MODE, bool CHECK, int SH = kSlotBits>
__device__ __forceinline__ void packed_upd(int &thr, int &pen, int D, int yi, int re, int rg, int ry, int rq, int maxd_q, unsigned bw,
                                           unsigned bw2, unsigned lut_s, int jk, int sti)
{
    constexpr int NEGMUL = -(1 << SH);   // the score sits above SH index bits
    if (MODE == MODE_MID) {
        asm volatile("{\n\t"
            ".reg .pred p;\n\t"
            ".reg .s32 key;\n\t"
            ".reg .u32 tb, ad;\n\t"
            "sub.s32 tb, %2, %3;\n\t"
            "setp.le.u32 p, tb, %5;\n\t"
            "add.u32 ad, tb, %6;\n\t"
            "@p ld.shared.u8 %1, [ad];\n\t"
            "mad.lo.s32 key, %1, %7, %4;\n\t"
            "@p max.s32 %0, %0, key;\n\t"
            "}"
            : "+r"(thr), "+r"(pen) : "r"(D), "r"(re), "r"(rg), "r"(bw2), "r"(lut_s), "n"(NEGMUL));
    } else if (MODE == MODE_FAR) {
        asm volatile("{\n\t"
            ".reg .pred p;\n\t"
            ".reg .s32 key, dq;\n\t"
            ".reg .u32 tb, ad;\n\t"
            "sub.s32 tb, %2, %3;\n\t"
            "setp.le.u32 p, tb, %5;\n\t"
            "add.u32 ad, tb, %6;\n\t"
            "@p ld.shared.u8 %1, [ad];\n\t"
            "sub.s32 dq, %7, %8;\n\t"
            "setp.le.and.s32 p, dq, %9, p;\n\t"
            "setp.ge.and.s32 p, %10, %11, p;\n\t"
            "mad.lo.s32 key, %1, %12, %4;\n\t"
            "@p max.s32 %0, %0, key;\n\t"
            "}"
            : "+r"(thr), "+r"(pen) : "r"(D), "r"(re), "r"(rg), "r"(bw2), "r"(lut_s), "r"(yi), "r"(ry), "r"(maxd_q),
              "r"(CHECK ? jk : 1), "r"(CHECK ? sti : 0), "n"(NEGMUL));
    } else {
        asm volatile("{\n\t"
            ".reg .pred p;\n\t"
            ".reg .s32 key, dq, dr, m, d;\n\t"
            ".reg .u32 tb, ad;\n\t"
            "sub.s32 tb, %2, %3;\n\t"
            "setp.le.u32 p, tb, %5;\n\t"
            "add.u32 ad, tb, %6;\n\t"
            "@p ld.shared.u8 %1, [ad];\n\t"
            "sub.s32 dq, %7, %8;\n\t"
            "add.s32 dr, dq, tb;\n\t"
            "sub.s32 dr, dr, %12;\n\t"
            "min.s32 m, dr, dq;\n\t"
            "min.s32 m, m, %13;\n\t"
            "setp.gt.and.s32 p, m, 0, p;\n\t"
            "setp.le.and.s32 p, dq, %9, p;\n\t"
            "setp.ge.and.s32 p, %10, %11, p;\n\t"
            "sub.s32 d, %13, m;\n\t"
            "add.s32 d, d, %1;\n\t"
            "mad.lo.s32 key, d, %14, %4;\n\t"
            "@p max.s32 %0, %0, key;\n\t"
            "}"
            : "+r"(thr), "+r"(pen) : "r"(D), "r"(re), "r"(rg), "r"(bw2), "r"(lut_s), "r"(yi), "r"(ry), "r"(maxd_q),
              "r"(CHECK ? jk : 1), "r"(CHECK ? sti : 0), "r"(bw), "r"(rq), "n"(NEGMUL));
    }
}

// predecessors [j0, j1) of the ring, walked in address-contiguous runs (a run never wraps, so loads are [base + imm])
template <int R, int MODE, bool CHECK>
__device__ __forceinline__ void packed_walk(int &thr, int &pen, int j0, int j1, int u0, const RecP *ring, int D, int yi, int maxd_q,
                                            unsigned bw, unsigned bw2, unsigned lut_s, int sti)
{
    int j = j0;
    while (j < j1) {
        const int pos = (j - u0) & (R - 1);
        const int len = min(j1 - j, R - pos);
        const RecP *rp = ring + pos;
#pragma unroll(MODE == MODE_MID ? 8 : 4)
        for (int k = 0; k < len; ++k) {
            if (MODE == MODE_MID) {
                const int2 r = *reinterpret_cast<const int2 *>(rp + k);
                packed_upd<MODE, CHECK>(thr, pen, D, yi, r.x, r.y, 0, 0, maxd_q, bw, bw2, lut_s, j + k, sti);
            } else {
                const int4 r = *reinterpret_cast<const int4 *>(rp + k);
                packed_upd<MODE, CHECK>(thr, pen, D, yi, r.x, r.y, r.z, r.w, maxd_q, bw, bw2, lut_s, j + k, sti);
            }
        }
        j += len;
    }
}

// in-tile candidate, f-independent part: (m - pen) << 13 if the pair is valid and bit `BIT` of vm is set (vm: which tile
// slots are predecessors of this lane's anchor), else kNegKey.  The vm test (bit = 1 << slot, an immediate once the slot loop is unrolled) opens the predicate chain, so it costs one LOP3.
__device__ __forceinline__ int packed_static(int &pen, int D, int yi, int re, int ry, int rq, int maxd_q, unsigned bw, unsigned bw2,
                                             unsigned lut_s, unsigned vm, unsigned bit)
{
    int w;
    asm volatile("{\n\t"
        ".reg .pred p;\n\t"
        ".reg .s32 dq, dr, m, d;\n\t"
        ".reg .u32 tb, ad, vb;\n\t"
        "and.b32 vb, %9, %12;\n\t"
        "setp.ne.u32 p, vb, 0;\n\t"
        "sub.s32 tb, %2, %3;\n\t"
        "setp.le.and.u32 p, tb, %4, p;\n\t"
        "add.u32 ad, tb, %5;\n\t"
        "@p ld.shared.u8 %1, [ad];\n\t"
        "sub.s32 dq, %6, %7;\n\t"
        "add.s32 dr, dq, tb;\n\t"
        "sub.s32 dr, dr, %10;\n\t"
        "min.s32 m, dr, dq;\n\t"
        "min.s32 m, m, %11;\n\t"
        "setp.gt.and.s32 p, m, 0, p;\n\t"
        "setp.le.and.s32 p, dq, %8, p;\n\t"
        "sub.s32 d, m, %1;\n\t"
        "shl.b32 d, d, 13;\n\t"
        "selp.s32 %0, d, -1073741824, p;\n\t"
        "}"
        : "=r"(w), "+r"(pen)
        : "r"(D), "r"(re), "r"(bw2), "r"(lut_s), "r"(yi), "r"(ry), "r"(maxd_q), "r"(vm), "r"(bw), "r"(rq), "r"(bit));
    return w;
}

// first c in [from, t0) whose ring x (= e + y) is >= X (signed; x is sorted over [from, t0)), t0 if none
template <int R>
__device__ __forceinline__ int ring_lower_bound(const RecP *ring, int u0, int from, int t0, int X, int lane)
{
    int j = from;
    while (j < t0) {
        const int c = j + lane;
        bool ge = true;
        if (c < t0) {
            const RecP &r = ring[(c - u0) & (R - 1)];
            ge = r.e + r.y >= X;
        }
        const unsigned b = __ballot_sync(0xffffffffu, ge);
        if (b) return min(t0, j + __ffs(b) - 1);
        j += 32;
    }
    return t0;
}

#ifndef MM2GB_SCORE_PREFETCH
#define MM2GB_SCORE_PREFETCH 2     // measured (profiles/r5b_ab.jsonl): 0 -> 2.104 ms, 1 -> 2.117 ms, 2 -> 2.097 ms per launch on configs[1]
#endif
constexpr int kStageBytes = MM2GB_SCORE_PREFETCH == 2 ? 32 * 16 + 32 * 4 : 0;   // per warp, behind the rings

template <int R>
__device__ void score_unit_packed(const uint4 *__restrict__ a, const int *__restrict__ st, int *f, int *__restrict__ p,
                                  int u0, int u1, int rbase, const DevParams &P, int qs_max, unsigned lut_s, RecP *ring, int lane,
                                  unsigned char *stage)
{
    const unsigned full = 0xffffffffu;
    const unsigned bw = (unsigned)P.bw, bw2 = 2u * (unsigned)P.bw;
    const int maxd_q = P.maxd_q;
    const int near_d = P.bw + qs_max;       // dr >  near_d  =>  min(dr, dq, q_span_j) = q_span_j inside the band
    const int far_d = maxd_q - P.bw;        // dr <= far_d   =>  dq <= maxd_q inside the band
    int pen = 0; // scratch of the table load, only rewritten inside the band
    int jA = u0, jB = u0; // first predecessor with x >= x_last - far_d / x >= x_first - near_d (monotone inside a rid/strand run)
    // MM2GB_SCORE_PREFETCH (build-time experiment, profiles/r5*_prefetch*): how a tile's own anchors + window starts reach the
    // lanes.  0 = loaded at the top of the tile (other warps cover the latency); 1 = the next tile's loads are issued into
    // registers before this tile's predecessors are walked; 2 = the same through cp.async (LDGSTS) into a 640-byte per-warp
    // staging area behind the rings, so the prefetch costs no registers (the kernel sits at its 80-register cap).
#if MM2GB_SCORE_PREFETCH == 1
    uint4 an = make_uint4(0, 0, 0, 0);
    int stn = INT32_MAX;
    if (u0 + lane < u1) { an = __ldg(a + u0 + lane); stn = st[u0 + lane]; }
#elif MM2GB_SCORE_PREFETCH == 2
    uint4 *stage_a = reinterpret_cast<uint4 *>(stage);
    int *stage_st = reinterpret_cast<int *>(stage_a + 32);
    const unsigned sa_s = (unsigned)__cvta_generic_to_shared(stage_a + lane), ss_s = (unsigned)__cvta_generic_to_shared(stage_st + lane);
    if (u0 + lane < u1) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa_s), "l"(a + u0 + lane) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ss_s), "l"(st + u0 + lane) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
    for (int t0 = u0; t0 < u1; t0 += 32) {
        const int i = t0 + lane;
        const bool act = i < u1;
        uint4 ai = make_uint4(0, 0, 0, 0);
        int sti = INT32_MAX; // inactive lanes: empty window
#if MM2GB_SCORE_PREFETCH == 1
        ai = an; sti = stn;
        an = make_uint4(0, 0, 0, 0); stn = INT32_MAX;
        if (i + 32 < u1) { an = __ldg(a + i + 32); stn = st[i + 32]; }
#elif MM2GB_SCORE_PREFETCH == 2
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (act) { ai = stage_a[lane]; sti = stage_st[lane]; }
        // a lane only ever reads and refills its own slot; the refill is issued after the reads above have returned (in-order
        // issue: the instructions between consume ai / sti) and lands hundreds of cycles later
        if (i + 32 < u1 && (int)ai.x + sti != INT32_MIN + 1) {   // (always true: ties the refill to the values just read)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa_s), "l"(a + i + 32) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ss_s), "l"(st + i + 32) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
#else
        if (act) { ai = __ldg(a + i); sti = st[i]; }
#endif
        const int xi = (int)ai.x, yi = (int)ai.z, qsi = (int)(ai.w & 0xffu);
        const int D = xi - yi + (int)bw;
        const int nact = min(32, u1 - t0);
        const int wmin = __shfl_sync(full, sti, 0);
        const int wfull = min(t0, __reduce_max_sync(full, act ? sti : 0));
        const int jring = max(u0, t0 - R);
        const int thr0 = (qsi << kSlotBits) | kSlotMask;
        int thr = thr0;

        for (int j = wmin; j < min(jring, t0); ++j) { // window longer than the ring: global / L1
            const uint4 v = __ldg(a + j);
            const int fj = f[j];
            const int rq = (int)(v.w & 0xffu);
            packed_upd<MODE_GEN, true>(thr, pen, D, yi, (int)v.x - (int)v.z, ((fj + rq) << kSlotBits) | (j - u0), (int)v.z, rq, maxd_q, bw, bw2,
                                       lut_s, j, sti);
        }
        const int lo = max(wmin, jring);        // first predecessor taken from the ring
        const int open = min(t0, max(wfull, lo)); // from here on every active lane's window is open
        // region boundaries (warp-uniform).  Everything in [open, t0) shares rid/strand with the whole tile, x ascending.
        jA = jB = t0;
        if (open < t0) {
            const int x_first = __shfl_sync(full, xi, 0), x_last = __shfl_sync(full, xi, nact - 1);
            jA = ring_lower_bound<R>(ring, u0, open, t0, x_last - far_d, lane);
            jB = ring_lower_bound<R>(ring, u0, open, t0, x_first - near_d, lane);
        }
        // windows still opening: [lo, open).  If the near region starts after `open` (jB > open), x_open < x_first - near_d
        // and x is ascending inside a lane's window, so every predecessor older than `open` is far for every lane that
        // has it in its window; otherwise (young unit, or a tile that straddles rid/strand runs) score them in full.
        if (jB > open) packed_walk<R, MODE_FAR, true>(thr, pen, lo, open, u0, ring, D, yi, maxd_q, bw, bw2, lut_s, sti);
        else packed_walk<R, MODE_GEN, true>(thr, pen, lo, open, u0, ring, D, yi, maxd_q, bw, bw2, lut_s, sti);
        // open windows: [open, t0) = FAR [open, min(jA, jB)) | MID [jA, jB) (if jA < jB) | GEN [jB, t0)
        const int e1 = min(jA, jB);
        packed_walk<R, MODE_FAR, false>(thr, pen, open, e1, u0, ring, D, yi, maxd_q, bw, bw2, lut_s, sti);
        packed_walk<R, MODE_MID, false>(thr, pen, jA, jB, u0, ring, D, yi, maxd_q, bw, bw2, lut_s, sti);
        packed_walk<R, MODE_GEN, false>(thr, pen, jB, t0, u0, ring, D, yi, maxd_q, bw, bw2, lut_s, sti);
        __syncwarp();
        RecP *tile = ring + ((t0 - u0) & (R - 1));
        if (act) { RecP r; r.e = xi - yi; r.g = 0; r.y = yi; r.q = qsi; tile[lane] = r; }
        __syncwarp();

        // phase B in two halves (s = 0..15, 16..30): the f-independent parts of 16 in-tile candidates are pre-scored into
        // registers, then resolved by the serial chain; halving keeps the live set at 16 values instead of 31.
        const int slot = i - u0;
        const int qk = qsi << kSlotBits;
        // this anchor as an in-tile predecessor: current score, own slot.  (Ring records carry score + q_span so that the MID
        // path needs no add; inside the tile the plain score saves one operation on every step of the serial chain.)
        int gown = (thr & ~kSlotMask) | slot;
        // bit s: anchor t0 + s is inside the window of its successor t0 + s + 1.  Window starts never decrease along a unit, so a
        // clear bit means NO later anchor of the tile has t0 + s in its window either: a half without set bits has no in-tile
        // candidates at all (tiles of isolated hits -- the chance hits of a large reference -- skip the triangle altogether)
        const unsigned needm = __ballot_sync(full, act && sti < i) >> 1;
        // vm: the tile slots s that are predecessors of this lane's anchor, s < lane and t0 + s >= st_i (none for an inactive lane)
        const int vlo = sti - t0;
        const unsigned vm = vlo >= lane ? 0u : (((1u << lane) - 1u) & ~((1u << max(vlo, 0)) - 1u));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h == 1 && nact <= 17) break; // warp-uniform: a short last tile has no candidates s >= 16
            if (((needm >> (16 * h)) & 0xffffu) == 0) continue;
            int w[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int s = h * 16 + q;
                if (s < 31) {
                    const int4 r = *reinterpret_cast<const int4 *>(tile + s);
                    w[q] = packed_static(pen, D, yi, r.x, r.z, r.w, maxd_q, bw, bw2, lut_s, vm, 1u << s);
                }
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int s = h * 16 + q;
                if (s < 31) {
                    const int gs = __shfl_sync(full, gown, s);
                    thr = max(thr, gs + w[q]);
                    gown = (thr & ~kSlotMask) | slot;
                }
            }
        }
        if (act) {
            f[i] = thr >> kSlotBits;
            p[i] = thr == thr0 ? -1 : u0 + (thr & kSlotMask) - rbase;
            tile[lane].g = gown + qk;
        }
        __syncwarp();
    }
}
template <int R, bool FAST>
__global__ void __launch_bounds__(kScoreWarps * 32, 6) // 6 CTAs/SM = what the 512-entry rings allow; caps registers at 80
k_score_units(const uint4 *__restrict__ a, const int *__restrict__ st, const int *__restrict__ unit_start,
              const int *__restrict__ unit_rbase, const unsigned *__restrict__ clipmask, int *f, int *__restrict__ p,
              const int *__restrict__ big_order, int big_cap, Counters *ctr, DevParams P, const unsigned char *__restrict__ lut_g,
              int run_mode, int long_classes, int long_wave)
{
    // the penalty table sits in STATIC shared memory so that its address is a compile-time constant: table loads are
    // LDS.U8 [tb + const] with no address arithmetic
    __shared__ __align__(16) unsigned char lut[2 * kLutMax + 16];
    extern __shared__ int4 smem_raw[];
    Rec *ring = reinterpret_cast<Rec *>(smem_raw) + (threadIdx.x >> 5) * R;
    const unsigned lut_s = (unsigned)__cvta_generic_to_shared(lut);
    const int lane = threadIdx.x & 31;
    // run_mode 0: always; 1: only if no read mixes segment ids (FAST is valid); 2: only if some read does
    if (run_mode == 1 && ctr->multi_sid != 0) return;
    if (run_mode == 2 && ctr->multi_sid == 0) return;
    for (int k = threadIdx.x; k < P.lut_n; k += blockDim.x) lut[k] = lut_g[k];
    __syncthreads();
    const int n_units = ctr->n_units;
    const int qs_max = max(ctr->qs_max, 1);
    int bb[kBigClasses + 1]; // list boundaries in queue order
    bb[0] = 0;
#pragma unroll
    for (int c = 0; c < kBigClasses; ++c) bb[c + 1] = bb[c] + min(ctr->big_cnt[c], big_cap);
    const int n_big = bb[kBigClasses];
    const int long_min = big_class_min(long_classes_eff(ctr, big_cap, long_classes, long_wave));
    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(&ctr->next_unit, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        int k;
        if (w < n_big) { // longest-first lists
            int c = 0, base = 0;
#pragma unroll
            for (int q = 1; q < kBigClasses; ++q) if (w >= bb[q]) c = q, base = bb[q];
            k = big_order[c * big_cap + (w - base)];
        } else {
            k = w - n_big;
            if (k >= n_units) break;
        }
        const int u0 = unit_start[k], u1 = unit_start[k + 1], rbase = unit_rbase[k];
        if (w >= n_big && u1 - u0 >= kBigMin) continue; // already taken from a list
        const bool clip = unit_has_clip(clipmask, u0, u1, lane);
        if (u1 - u0 >= long_min && !clip) continue; // k_score_long's
        if (clip) {
            if (lane == 0) atomicAdd(&ctr->n_exact, 1);
            score_unit_exact<FAST>(a, st, f, p, u0, u1, rbase, P, lut_s, lane);
        } else if (FAST && u1 - u0 <= (1 << kSlotBits) && (u1 - u0 + 1) * qs_max < (1 << 18)) {
            score_unit_packed<R>(a, st, f, p, u0, u1, rbase, P, qs_max, lut_s, reinterpret_cast<RecP *>(ring), lane,
                                 reinterpret_cast<unsigned char *>(smem_raw) + (size_t)kScoreWarps * R * sizeof(Rec) + (threadIdx.x >> 5) * kStageBytes);
        } else {
            score_unit_tiled<R, FAST>(a, st, f, p, u0, u1, rbase, P, lut_s, ring, lane);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_score_long: one CTA per long unit (>= long_min anchors), its warps pipelined over the unit's 32-anchor tiles
//                                                                       (replaces gpu/plscore.cu:370-451, the long kernel)
//
// A long independent segment cannot be split -- f[i] needs f of the whole window before it -- and one warp walking it
// alone leaves the unit's latency at (window x issue + in-tile chain) per tile.  Here warp w of the CTA owns tiles
// w, w + NW, w + 2NW, ...: it sweeps the part of its window that earlier tiles have already finished while the tiles just
// before its own are still being resolved by the other warps, and only the last step -- the newest predecessor tile and
// the 32x32 in-tile triangle -- stays on the critical path.  Tiles finish in order; `s_done` (tiles finished, shared
// memory, release/acquire) is the only synchronisation.  All warps share ONE ring of RL records (64 KB for 4096), so
// windows of up to RL - 32 NW anchors are served from shared memory (the one-warp kernel has 512 per warp).
//
// Scores and predecessors are kept in two registers here (no f << 13 | slot packing), so neither the unit length nor
// the chain score is bounded.  Records are in diagonal form as in score_unit_packed: e = x - y, g = f + q_span, y, q_span;
// predecessor tiles are classified MID / FAR / GEN per (tile, predecessor tile) exactly as there.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef MM2GB_LONG_SLEEP_NEAR
#define MM2GB_LONG_SLEEP_NEAR 20     // ns between polls while waiting for the tile right before this warp's own ...
#define MM2GB_LONG_SLEEP_FAR 200     // ... and for an older one
#endif
constexpr int kLongWarps = 8;
constexpr int kLongRing = 4096;

struct __align__(16) RecL { int e, g, y, q; };

__device__ __forceinline__ int ld_acquire_shared(const int *p)
{
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_shared(int *p, int v)
{
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// one candidate j = jb + K for the anchor of this lane: (thr, bj) <- (val, j) iff the pair is valid and val >= thr.
// As packed_upd, the validity tests stay ONE predicate chain; the update is two predicated moves.  K is an immediate so
// that the predecessor index costs one predicated add and no register.
template <int MODE, bool CHECK, int K>
__device__ __forceinline__ void long_upd(int &thr, int &bj, int &pen, int D, int yi, int re, int rg, int ry, int rq, int maxd_q, unsigned bw,
                                         unsigned bw2, unsigned lut_s, int jb, int sti)
{
    if (MODE == MODE_MID) {
        asm volatile("{\n\t"
            ".reg .pred p;\n\t"
            ".reg .s32 val;\n\t"
            ".reg .u32 tb, ad;\n\t"
            "sub.s32 tb, %3, %4;\n\t"
            "setp.le.u32 p, tb, %6;\n\t"
            "add.u32 ad, tb, %7;\n\t"
            "@p ld.shared.u8 %2, [ad];\n\t"
            "sub.s32 val, %5, %2;\n\t"
            "setp.ge.and.s32 p, val, %0, p;\n\t"
            "@p mov.s32 %0, val;\n\t"
            "@p add.s32 %1, %8, %9;\n\t"
            "}"
            : "+r"(thr), "+r"(bj), "+r"(pen) : "r"(D), "r"(re), "r"(rg), "r"(bw2), "r"(lut_s), "r"(jb), "n"(K));
    } else if (MODE == MODE_FAR) {
        asm volatile("{\n\t"
            ".reg .pred p;\n\t"
            ".reg .s32 val, dq, jk;\n\t"
            ".reg .u32 tb, ad;\n\t"
            "sub.s32 tb, %3, %4;\n\t"
            "setp.le.u32 p, tb, %6;\n\t"
            "add.u32 ad, tb, %7;\n\t"
            "@p ld.shared.u8 %2, [ad];\n\t"
            "sub.s32 dq, %10, %11;\n\t"
            "setp.le.and.s32 p, dq, %12, p;\n\t"
            "add.s32 jk, %8, %9;\n\t"
            "setp.ge.and.s32 p, jk, %13, p;\n\t"
            "sub.s32 val, %5, %2;\n\t"
            "setp.ge.and.s32 p, val, %0, p;\n\t"
            "@p mov.s32 %0, val;\n\t"
            "@p mov.s32 %1, jk;\n\t"
            "}"
            : "+r"(thr), "+r"(bj), "+r"(pen) : "r"(D), "r"(re), "r"(rg), "r"(bw2), "r"(lut_s), "r"(jb), "n"(K), "r"(yi), "r"(ry), "r"(maxd_q),
              "r"(CHECK ? sti : INT32_MIN));
    } else {
        asm volatile("{\n\t"
            ".reg .pred p;\n\t"
            ".reg .s32 val, dq, dr, m, d, jk;\n\t"
            ".reg .u32 tb, ad;\n\t"
            "sub.s32 tb, %3, %4;\n\t"
            "setp.le.u32 p, tb, %6;\n\t"
            "add.u32 ad, tb, %7;\n\t"
            "@p ld.shared.u8 %2, [ad];\n\t"
            "sub.s32 dq, %10, %11;\n\t"
            "add.s32 dr, dq, tb;\n\t"
            "sub.s32 dr, dr, %14;\n\t"
            "min.s32 m, dr, dq;\n\t"
            "min.s32 m, m, %15;\n\t"
            "setp.gt.and.s32 p, m, 0, p;\n\t"
            "setp.le.and.s32 p, dq, %12, p;\n\t"
            "add.s32 jk, %8, %9;\n\t"
            "setp.ge.and.s32 p, jk, %13, p;\n\t"
            "sub.s32 d, m, %15;\n\t"
            "sub.s32 d, d, %2;\n\t"
            "add.s32 val, %5, d;\n\t"
            "setp.ge.and.s32 p, val, %0, p;\n\t"
            "@p mov.s32 %0, val;\n\t"
            "@p mov.s32 %1, jk;\n\t"
            "}"
            : "+r"(thr), "+r"(bj), "+r"(pen) : "r"(D), "r"(re), "r"(rg), "r"(bw2), "r"(lut_s), "r"(jb), "n"(K), "r"(yi), "r"(ry), "r"(maxd_q),
              "r"(CHECK ? sti : INT32_MIN), "r"(bw), "r"(rq));
    }
}

template <int MODE, bool CHECK, int K>
__device__ __forceinline__ void long_step(int &thr, int &bj, int &pen, const RecL *rp, int jb, int D, int yi, int maxd_q, unsigned bw,
                                          unsigned bw2, unsigned lut_s, int sti)
{
    if (MODE == MODE_MID) {
        const int2 r = *reinterpret_cast<const int2 *>(rp + K);
        long_upd<MODE, CHECK, K>(thr, bj, pen, D, yi, r.x, r.y, 0, 0, maxd_q, bw, bw2, lut_s, jb, sti);
    } else {
        const int4 r = *reinterpret_cast<const int4 *>(rp + K);
        long_upd<MODE, CHECK, K>(thr, bj, pen, D, yi, r.x, r.y, r.z, r.w, maxd_q, bw, bw2, lut_s, jb, sti);
    }
}

// the 32 records of one predecessor tile (never wraps in the ring: tiles are 32-aligned)
template <int MODE, bool CHECK>
__device__ __forceinline__ void long_walk32(int &thr, int &bj, int &pen, const RecL *rp, int j0, int D, int yi, int maxd_q, unsigned bw,
                                            unsigned bw2, unsigned lut_s, int sti)
{
#pragma unroll 1
    for (int kk = 0; kk < 32; kk += 8) {
        const RecL *r8 = rp + kk;
        const int jb = j0 + kk;
#define MM2GB_LSTEP(K) long_step<MODE, CHECK, K>(thr, bj, pen, r8, jb, D, yi, maxd_q, bw, bw2, lut_s, sti);
        MM2GB_LSTEP(0) MM2GB_LSTEP(1) MM2GB_LSTEP(2) MM2GB_LSTEP(3) MM2GB_LSTEP(4) MM2GB_LSTEP(5) MM2GB_LSTEP(6) MM2GB_LSTEP(7)
#undef MM2GB_LSTEP
    }
}

// The same 32 records with the candidates of ONE predecessor tile reduced in a packed register first:
//     key = (f_j + sc) << 5 | (j & 31),   tkey = max(tkey, key)
// (the predicated single max of score_unit_packed; 5 index bits are enough inside a tile, so scores up to 2^26 fit), then
// one merge per tile into (thr, bj) with the '>=' rule.  Inside the tile the max prefers the larger j among equal scores,
// across tiles '>=' in ascending order does: the result is that of the one-candidate-at-a-time update.
constexpr int kTileBits = 5;
template <int MODE, bool CHECK>
__device__ __forceinline__ void long_walk32_pk(int &thr, int &bj, int &pen, const RecL *rp, int j0, int D, int yi, int maxd_q, unsigned bw,
                                               unsigned bw2, unsigned lut_s, int sti)
{
    int tkey = kNegKey;
#pragma unroll 1
    for (int kk = 0; kk < 32; kk += 8) {
        const RecL *r8 = rp + kk;
        const int jb = j0 + kk;
#pragma unroll
        for (int K = 0; K < 8; ++K) {
            if (MODE == MODE_MID) {
                const int2 r = *reinterpret_cast<const int2 *>(r8 + K);
                packed_upd<MODE, CHECK, kTileBits>(tkey, pen, D, yi, r.x, r.y, 0, 0, maxd_q, bw, bw2, lut_s, jb + K, sti);
            } else {
                const int4 r = *reinterpret_cast<const int4 *>(r8 + K);
                packed_upd<MODE, CHECK, kTileBits>(tkey, pen, D, yi, r.x, r.y, r.z, r.w, maxd_q, bw, bw2, lut_s, jb + K, sti);
            }
        }
    }
    const int val = tkey >> kTileBits;
    if (val >= thr) { thr = val; bj = j0 + (tkey & 31); }
}

struct ExactState;
template <int RL>
__device__ __forceinline__ void exact_row(ExactState &M, int s, int t0, const uint4 *__restrict__ a, const int *f, const uint4 &ai, int sti,
                                          int &fcur, int &bj, const DevParams &P, unsigned lut_s, int &pen, int lane);

// EXACT: the unit has windows clipped by max_iter, so the max_ii fallback of lchain.c:189-205 (SURVEY.md trap T3) runs as well.
// Its state -- max_ii and that anchor's x, y, q_span, f -- is one more serial dependency from row to row: inside a tile it is
// carried in warp-uniform registers through the steps of the in-tile chain (row s is final, fallback candidate included,
// before its score is broadcast to the rows behind it), between tiles through s_mi[] (written before the release of s_done).
// The rescan of a whole window (max_ii has left the max_dist_x range) is a warp-wide argmax over f[].
struct ExactState { int mi, xhi, xlo, y, q, f; };

// Row t0 + s of an EXACT tile, executed by the whole warp when the in-tile chain reaches it: lane s holds the row's best
// (fcur, bj) over its window; M is the max_ii state after row t0 + s - 1 (warp-uniform).  lchain.c:189-205:
//   rescan  if max_ii < 0 or x_i - x[max_ii] > max_dist_x:  max_ii = argmax f over the window (largest j among equals, -1 if empty)
//   try     if 0 <= max_ii < st_i - 1:  the pair (i, max_ii) as one more candidate (strictly better wins)
//   keep    f[i] is final; if max_ii < 0 or (x_i - x[max_ii] <= max_dist_x and f[max_ii] < f[i]):  max_ii = i
template <int RL>
__device__ __forceinline__ void exact_row(ExactState &M, int s, int t0, const uint4 *__restrict__ a, const int *f, const uint4 &ai, int sti,
                                          int &fcur, int &bj, const DevParams &P, unsigned lut_s, int &pen, int lane)
{
    const unsigned full = 0xffffffffu;
    const int i = t0 + s;
    const int xlo = __shfl_sync(full, (int)ai.x, s), xhi = __shfl_sync(full, (int)ai.y, s), st_s = __shfl_sync(full, sti, s);
    // x is sorted, so x_i >= x[max_ii]: another rid/strand word means a distance of at least 2^32
    bool far = M.mi < 0 || xhi != M.xhi || (unsigned)(xlo - M.xlo) > (unsigned)P.max_dist_x;
    if (far) {
        int fv = INT32_MIN, fj = -1;
        for (int j0 = st_s; j0 < min(i, t0); j0 += 32) {       // finished tiles: from global memory (L2)
            const int j = j0 + lane;
            if (j < min(i, t0)) { const int v = __ldcg(f + j); if (v >= fv) fv = v, fj = j; }
        }
        if (lane < s && t0 + lane >= st_s && fcur >= fv) fv = fcur, fj = t0 + lane;    // rows of this tile that are final already
        const int fm = __reduce_max_sync(full, fv);
        M.mi = __reduce_max_sync(full, (fv == fm && fj >= 0) ? fj : -1);
        if (M.mi >= 0) {
            M.f = fm;
            if (M.mi >= t0) {
                const int l = M.mi - t0;
                M.xlo = __shfl_sync(full, (int)ai.x, l); M.xhi = __shfl_sync(full, (int)ai.y, l);
                M.y = __shfl_sync(full, (int)ai.z, l); M.q = __shfl_sync(full, (int)(ai.w & 0xffu), l);
            } else {
                const uint4 v = __ldg(a + M.mi);
                M.xlo = (int)v.x; M.xhi = (int)v.y; M.y = (int)v.z; M.q = (int)(v.w & 0xffu);
            }
            far = xhi != M.xhi || (unsigned)(xlo - M.xlo) > (unsigned)P.max_dist_x;
        }
    }
    int fs = fcur;
    if (M.mi >= 0 && M.mi < st_s - 1 && lane == s) {
        Rec r; r.x = M.xlo; r.y = M.y; r.f = M.f; r.q = M.q;
        int sc;
        if (pair_fast((int)ai.x, (int)ai.z, r, P.maxd_q, (unsigned)P.bw, lut_s, pen, sc) && fcur < sc + M.f) { fcur = sc + M.f; bj = M.mi; }
        fs = fcur;
    }
    fs = __shfl_sync(full, fs, s);      // f[i], final
    if (M.mi < 0 || (!far && M.f < fs)) {
        M.mi = i; M.xlo = xlo; M.xhi = xhi; M.f = fs;
        M.y = __shfl_sync(full, (int)ai.z, s); M.q = __shfl_sync(full, (int)(ai.w & 0xffu), s);
    }
}

template <int RL, int NW, bool PACK, bool EXACT = false>
__device__ void score_unit_long(const uint4 *__restrict__ a, const int *__restrict__ st, int *f, int *__restrict__ p, int u0, int u1,
                                int rbase, const DevParams &P, int qs_max, unsigned lut_s, RecL *ring, int *s_done, int warp, int lane,
                                ExactState *s_mi = nullptr)
{
    const unsigned full = 0xffffffffu;
    const unsigned bw = (unsigned)P.bw, bw2 = 2u * (unsigned)P.bw;
    const int maxd_q = P.maxd_q;
    const int near_d = P.bw + qs_max;       // dr >  near_d  =>  min(dr, dq, q_span_j) = q_span_j inside the band
    const int far_d = maxd_q - P.bw;        // dr <= far_d   =>  dq <= maxd_q inside the band
    const int ntiles = (u1 - u0 + 31) >> 5;
    int pen = 0;
    int done = 0;                           // tiles known to be finished
    for (int t = warp; t < ntiles; t += NW) {
        const int t0 = u0 + 32 * t;
        const int i = t0 + lane;
        const bool act = i < u1;
        uint4 ai = make_uint4(0, 0, 0, 0);
        int sti = INT32_MAX; // inactive lanes: empty window
        if (act) { ai = __ldg(a + i); sti = st[i]; }
        const int xi = (int)ai.x, yi = (int)ai.z, qsi = (int)(ai.w & 0xffu);
        const int D = xi - yi + (int)bw;
        const int nact = min(32, u1 - t0);
        const int wmin = __shfl_sync(full, sti, 0);                        // st is non-decreasing, lane 0 is always active
        const int wfull = min(t0, __reduce_max_sync(full, act ? sti : 0)); // from here on every active lane's window is open
        const int x_first = __shfl_sync(full, xi, 0), x_last = __shfl_sync(full, xi, nact - 1);
        int thr = qsi + 1, bj = -1; // a candidate wins iff val >= thr: '>' against q_span(i), '>=' afterwards (ascending j)
        // tiles this warp may still find in the ring: the NW - 1 tiles in flight behind it may overwrite anything older
        const int tring = max(0, t - RL / 32 + NW);
        const int jring = u0 + 32 * tring;
        // (rare) window older than the ring: from global memory; those tiles were finished before this warp's previous tile
        for (int j = wmin; j < min(jring, t0); ++j) {
            const uint4 v = __ldg(a + j);
            const int fj = __ldcg(f + j);
            const int rq = (int)(v.w & 0xffu);
            long_upd<MODE_GEN, true, 0>(thr, bj, pen, D, yi, (int)v.x - (int)v.z, fj + rq, (int)v.z, rq, maxd_q, bw, bw2, lut_s, j, sti);
        }
        const int kfirst = max(tring, (max(wmin, u0) - u0) >> 5);
        for (int k = kfirst; k < t; ++k) {
            if (k >= done) { // wait for tile k (tiles finish in order)
                for (;;) {
                    done = ld_acquire_shared(s_done);
                    if (k < done) break;
                    __nanosleep(t - done > 1 ? MM2GB_LONG_SLEEP_FAR : MM2GB_LONG_SLEEP_NEAR);
                }
            }
            const int j0 = u0 + 32 * k;
            const RecL *rp = ring + ((32 * k) & (RL - 1));
#define MM2GB_LWALK(MODE, CHECK)                                                                                         \
    do {                                                                                                                \
        if (PACK) long_walk32_pk<MODE, CHECK>(thr, bj, pen, rp, j0, D, yi, maxd_q, bw, bw2, lut_s, sti);                \
        else long_walk32<MODE, CHECK>(thr, bj, pen, rp, j0, D, yi, maxd_q, bw, bw2, lut_s, sti);                        \
    } while (0)
            if (j0 < wfull) { // some lane's window opens inside or after this tile (or the tile straddles rid/strand runs)
                MM2GB_LWALK(MODE_GEN, true);
            } else {          // inside every lane's window: same rid/strand as the whole tile, x ascending
                const int xk_first = rp[0].e + rp[0].y, xk_last = rp[31].e + rp[31].y;
                if (x_first - xk_last > near_d) {
                    if (x_last - xk_first <= far_d) MM2GB_LWALK(MODE_MID, false);
                    else MM2GB_LWALK(MODE_FAR, false);
                } else MM2GB_LWALK(MODE_GEN, false);
            }
#undef MM2GB_LWALK
        }
        // publish this tile's static fields (nobody reads them as a predecessor before s_done passes t)
        RecL *tile = ring + ((32 * t) & (RL - 1));
        if (act) { RecL r; r.e = xi - yi; r.g = 0; r.y = yi; r.q = qsi; tile[lane] = r; }
        __syncwarp();
        // phase B: the in-tile triangle, f-independent parts first, then the serial chain (one shuffle per step)
        int fcur = bj >= 0 ? thr : qsi;
        ExactState M = {-1, 0, 0, 0, 0, 0};
        if (EXACT) {    // the state after row t0 - 1: tile t - 1 has to be finished (it usually is: it is in every window)
            while (done < t) done = ld_acquire_shared(s_done);
            M = *s_mi;
        }
        // the tile slots s that are predecessors of this lane's anchor (s < lane and t0 + s >= st_i), one bit test per slot below
        const int vlo = sti - t0;
        const unsigned vm = vlo >= lane ? 0u : (((1u << lane) - 1u) & ~((1u << max(vlo, 0)) - 1u));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h == 1 && nact <= 17) break; // warp-uniform: a short last tile has no candidates s >= 16
            int w[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int s = h * 16 + q;
                if (s < 31) {
                    const int4 r = *reinterpret_cast<const int4 *>(tile + s);
                    const unsigned tb = (unsigned)(D - r.x);
                    bool ok = ((vm >> s) & 1u) && tb <= bw2;
                    if (ok) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(pen) : "r"(lut_s + tb));
                    const int dq = yi - r.z;
                    const int dr = dq + (int)tb - (int)bw;
                    const int m = min(min(dr, dq), r.w);
                    ok = ok && m > 0 && dq <= maxd_q;
                    w[q] = ok ? m - pen : kNeg;
                }
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int s = h * 16 + q;
                if (s < 31) {
                    if (EXACT && s < nact) exact_row<RL>(M, s, t0, a, f, ai, sti, fcur, bj, P, lut_s, pen, lane);
                    const int fs = __shfl_sync(full, fcur, s);
                    const int val = fs + w[q];                  // kNeg + f stays far below any threshold
                    if (val >= thr) thr = val, fcur = val, bj = t0 + s;
                }
            }
        }
        if (EXACT) {
            // rows the loops above did not finalise: row 31 always, rows 16 .. 30 if the second half was skipped (short last tile)
            for (int s = (nact <= 17 ? 16 : 31); s < nact; ++s) exact_row<RL>(M, s, t0, a, f, ai, sti, fcur, bj, P, lut_s, pen, lane);
        }
        if (act) {
            f[i] = fcur;
            p[i] = bj < 0 ? -1 : bj - rbase;
            tile[lane].g = PACK ? (((fcur + qsi) << kTileBits) | lane) : fcur + qsi;
        }
        if (EXACT && lane == 0) *s_mi = M;
        // tiles finish in order: a tile whose window does not reach back into tile t - 1 (a cut inside the unit) has not
        // waited for it yet
        while (done < t) done = ld_acquire_shared(s_done);
        __syncwarp();
        if (lane == 0) st_release_shared(s_done, t + 1);
        done = t + 1;
    }
}

template <int RL, int NW>
__global__ void __launch_bounds__(NW * 32, 3)
k_score_long(const uint4 *__restrict__ a, const int *__restrict__ st, const int *__restrict__ unit_start, const int *__restrict__ unit_rbase,
             const unsigned *__restrict__ clipmask, int *f, int *__restrict__ p, const int *__restrict__ big_order, int big_cap, Counters *ctr,
             DevParams P, const unsigned char *__restrict__ lut_g, int long_classes_max, int long_wave)
{
    __shared__ __align__(16) unsigned char lut[2 * kLutMax + 16];
    __shared__ int s_unit, s_done;
    extern __shared__ int4 smem_raw[];
    RecL *ring = reinterpret_cast<RecL *>(smem_raw);
    unsigned lut_s = (unsigned)__cvta_generic_to_shared(lut);
    asm volatile("" : "+r"(lut_s)); // keep the table address in a register (otherwise it is rematerialised per pair)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (ctr->multi_sid != 0) return; // the table path is not valid for this batch: k_score_units<false> takes every unit
    const int long_classes = long_classes_eff(ctr, big_cap, long_classes_max, long_wave);
    int n_list = 0;
    for (int c = 0; c < long_classes; ++c) n_list += min(ctr->big_cnt[c], big_cap);
    if ((int)blockIdx.x >= n_list) return;
    for (int k = threadIdx.x; k < P.lut_n; k += blockDim.x) lut[k] = lut_g[k];
    const int qs_max = max(ctr->qs_max, 1);
    for (;;) {
        __syncthreads(); // everyone is done with s_unit / s_done / the ring of the previous unit
        if (threadIdx.x == 0) { s_unit = atomicAdd(&ctr->next_long, 1); s_done = 0; }
        __syncthreads();
        const int w = s_unit;
        if (w >= n_list) break;
        int c = 0, base = 0, acc = 0;
        for (int q = 0; q < long_classes; ++q) { if (w >= acc) c = q, base = acc; acc += min(ctr->big_cnt[q], big_cap); }
        const int k = big_order[c * big_cap + (w - base)];
        const int u0 = unit_start[k], u1 = unit_start[k + 1], rbase = unit_rbase[k];
        if (unit_has_clip(clipmask, u0, u1, lane)) continue; // exact max_ii path, k_score_units
        if (threadIdx.x == 0) atomicAdd(&ctr->n_long, 1);
        // packed per-tile keys hold scores below 2^26 (f <= unit length x largest q_span)
        if ((long long)(u1 - u0 + 1) * qs_max < (1LL << 26))
            score_unit_long<RL, NW, true>(a, st, f, p, u0, u1, rbase, P, qs_max, lut_s, ring, &s_done, warp, lane);
        else
            score_unit_long<RL, NW, false>(a, st, f, p, u0, u1, rbase, P, qs_max, lut_s, ring, &s_done, warp, lane);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_score_exact: the clipped units of >= kExactMin anchors that k_score_units queued (tandem repeats, dense repeat families:
// more than max_iter anchors inside max_dist_x), one CTA each, score_unit_long in EXACT mode.  The ring holds 8192 records, so
// windows of max_iter = 5000 anchors plus the tiles in flight are served from shared memory.  Launched behind k_score_units on
// the same stream; exits at once when nothing was queued (the usual case).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kExactRing = 8192;
constexpr int kExactWarps = 16;            // tiles in flight per unit (the CTA has the SM to itself: 128 KB of ring)

template <int RL, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
k_score_exact(const uint4 *__restrict__ a, const int *__restrict__ st, const int *__restrict__ unit_start, const int *__restrict__ unit_rbase,
              int *f, int *__restrict__ p, const int *__restrict__ exact_list, Counters *ctr, DevParams P, const unsigned char *__restrict__ lut_g)
{
    __shared__ __align__(16) unsigned char lut[2 * kLutMax + 16];
    __shared__ int s_unit, s_done;
    __shared__ ExactState s_mi;
    extern __shared__ int4 smem_raw[];
    RecL *ring = reinterpret_cast<RecL *>(smem_raw);
    unsigned lut_s = (unsigned)__cvta_generic_to_shared(lut);
    asm volatile("" : "+r"(lut_s));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (ctr->multi_sid != 0) return; // the table path is not valid for this batch: k_score_units<false> takes every unit
    const int n_list = ctr->exact_cnt;
    if ((int)blockIdx.x >= n_list) return;
    for (int k = threadIdx.x; k < P.lut_n; k += blockDim.x) lut[k] = lut_g[k];
    const int qs_max = max(ctr->qs_max, 1);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            s_unit = atomicAdd(&ctr->next_exact, 1);
            s_done = 0;
            s_mi.mi = -1; s_mi.xhi = s_mi.xlo = s_mi.y = s_mi.q = s_mi.f = 0;
        }
        __syncthreads();
        const int w = s_unit;
        if (w >= n_list) break;
        const int k = exact_list[w];
        const int u0 = unit_start[k], u1 = unit_start[k + 1], rbase = unit_rbase[k];
        if ((long long)(u1 - u0 + 1) * qs_max < (1LL << 26))
            score_unit_long<RL, NW, true, true>(a, st, f, p, u0, u1, rbase, P, qs_max, lut_s, ring, &s_done, warp, lane, &s_mi);
        else
            score_unit_long<RL, NW, false, true>(a, st, f, p, u0, u1, rbase, P, qs_max, lut_s, ring, &s_done, warp, lane, &s_mi);
    }
}

// the same queue, one WARP per unit with the row-by-row path (score_unit_exact): what these units cost before k_score_exact
// existed; kept as the A/B switch MM2GB_EXACT_BIG=0
__global__ void __launch_bounds__(256)
k_score_exact_warp(const uint4 *__restrict__ a, const int *__restrict__ st, const int *__restrict__ unit_start, const int *__restrict__ unit_rbase,
                   int *f, int *__restrict__ p, const int *__restrict__ exact_list, Counters *ctr, DevParams P, const unsigned char *__restrict__ lut_g)
{
    __shared__ __align__(16) unsigned char lut[2 * kLutMax + 16];
    const unsigned lut_s = (unsigned)__cvta_generic_to_shared(lut);
    const int lane = threadIdx.x & 31;
    const int n_list = ctr->exact_cnt;
    if (n_list == 0 || ctr->multi_sid != 0) return;
    for (int k = threadIdx.x; k < P.lut_n; k += blockDim.x) lut[k] = lut_g[k];
    __syncthreads();
    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(&ctr->next_exact, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_list) break;
        const int k = exact_list[w];
        score_unit_exact<true>(a, st, f, p, unit_start[k], unit_start[k + 1], unit_rbase[k], P, lut_s, lane);
    }
}

} // namespace mm2gb
