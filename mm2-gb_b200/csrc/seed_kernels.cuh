// seed_kernels.cuh -- device seeding: mm_map_seed (map.c:355-391) as data-parallel kernels for sm_100a.
//
//   k_sketch          mm_sketch (sketch.c:77-143): every rule of the sequential winnowing loop depends only on the k-mer hashes
//                     of the last w + 1 positions and on the length of the current run of unambiguous bases, so one thread per
//                     position evaluates it from a shared-memory tile (count pass, scan, write pass: the output stays ordered).
//   k_qocc_*          mm_seed_mz_flt (seed.c:5-29): occurrences of a minimizer inside its own read, counted in a per-read
//                     open-addressing table in HBM.
//   k_lookup          mm_seed_collect_all (seed.c:31-53): mm_idx_get (index.c:81-97) against the device hash table + is_tandem.
//   k_select          mm_seed_select (seed.c:57-96) in closed form (the binary heap keeps the k smallest (n, j) pairs of a streak
//                     of high-occurrence seeds) + the flt rule of mm_collect_matches (seed.c:106-113) and rep_len (:117-121,128).
//   k_expand          anchor construction of collect_seed_hits (map.c:303-325).
//   k_seed_sort       radix_sort_128x (ksort.h:98-151) per read: the reference's in-place MSD radix sort is NOT stable, and the
//                     order it leaves among anchors of equal x decides chaining ties downstream, so its American-flag passes
//                     are replayed step for step on 4-byte (digit, index) words in shared memory; buckets of <= 64 elements
//                     (stable insertion sort in the reference) are ranked by a warp.
//
//   k_hpc_*           homopolymer compression (MM_I_HPC, sketch.c:92-101) as a stream compaction: one element per run of equal bases;
//                     k_sketch<u64, true> then runs the same rules on the elements, every record carrying its own span.
//
// Integer / byte work, HBM- and latency-bound: no tensor cores here.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace mm2gb_seed {

typedef unsigned long long u64;
typedef unsigned int u32;

#ifndef MM2GB_SKETCH_TILE
#define MM2GB_SKETCH_TILE 512
#endif
constexpr int kTile = MM2GB_SKETCH_TILE;   // positions per sketch tile = threads per CTA (a multiple of 32)
constexpr int kHalo = 96;            // staged positions ahead of a tile (>= w + k, a multiple of 32)
constexpr int kMaxW = 32;
constexpr u64 kNone = ~0ULL;

// sketch.c:9-26 (seq_nt4_table): ACGT / acgt -> 0..3, U/u -> 3, everything else 4
__device__ __forceinline__ int nt4(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    case 0: case 1: case 2: case 3: return c;     // the table maps bytes 0..3 to themselves
    default: return 4;
    }
}

// sketch.c:28-38
__host__ __device__ __forceinline__ u64 hash64(u64 key, u64 mask)
{
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

// ---- sketch ---------------------------------------------------------------------------------------------------------------
//
// Restated for odd k without homopolymer compression (a k-mer then never equals its reverse complement, so the loop never
// takes its `continue`, sketch.c:104, and buffer slot = position mod w).  Let info(j) be the (hash << 8 | k, position, strand)
// record of the k-mer ending at j, or NONE if a base of it is ambiguous or j < k - 1; l(i) the number of unambiguous bases
// in a row ending at i.  The loop keeps  min = the rightmost minimum of info over the last w positions, and at position i:
//   (P1) l(i) == w+k-1 and min(i-1) != NONE: emit the records of [i-w+1, i-1] equal to min(i-1).x other than min(i-1) itself;
//   (P2) info(i).x <= min(i-1).x: emit min(i-1) if l(i) >= w+k and it is not NONE;
//   (P3) else if min(i-1) sits at i-w: emit it if l(i) >= w+k-1; then, if l(i) >= w+k-1 and the new minimum min(i) of
//        [i-w+1, i] is not NONE, emit the other records of that window equal to min(i).x, oldest first;
//   (P4) after the last position emit the final minimum if it is not NONE.
// One pass: a tile's minimizer count is published to the tiles behind it through a chained scan with decoupled look-back
// (ticket-ordered tiles; status word = flag << 62 | value), so the ordered output offset is known without a counting pass.
// Bases are staged as 2-bit codes packed 32 per 64-bit word (earlier base = higher bits), so a k-mer is one funnel shift and
// its reverse complement one bit reversal; for k <= 16 the hash and all window comparisons are 32-bit (HT = u32).
template <typename HT>
struct SketchTile {
    u64 pk[(kHalo + kTile) / 32];
    u32 nmask[(kHalo + kTile) / 32];
    HT ih[kMaxW + kTile];             // hash of the k-mer ending at positions t0 - w .. t0 + kTile - 1 (all ones: none)
    unsigned char iz[kMaxW + kTile];  // strand bit
    u32 warp_sum[kTile / 32];
    u64 excl;
    int seq, tile;
};

__device__ __forceinline__ u64 spread32(u32 v)
{
    u64 x = v;
    x = (x | x << 16) & 0x0000FFFF0000FFFFULL;
    x = (x | x << 8) & 0x00FF00FF00FF00FFULL;
    x = (x | x << 4) & 0x0F0F0F0F0F0F0F0FULL;
    x = (x | x << 2) & 0x3333333333333333ULL;
    x = (x | x << 1) & 0x5555555555555555ULL;
    return x;
}

__device__ __forceinline__ u32 hash32(u32 key, u32 mask)    // hash64 (sketch.c:28-38) for masks of at most 32 bits
{
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

#define MM2GB_FLAG_AGG (1ULL << 62)
#define MM2GB_FLAG_PREFIX (2ULL << 62)
#define MM2GB_VAL_MASK ((1ULL << 62) - 1ULL)

// A launch covers the tiles [tile_begin, tile_end) (a batch is sketched in a few launches so that the upload of its later reads
// overlaps the sketch of its earlier ones; launches of a batch run in stream order, so every predecessor tile of an earlier launch
// is complete); *ticket = that launch's ticket counter, scan_state[tile] = status word; all zero before the first launch
// HPC (MM_I_HPC, sketch.c:92-101): the kernel then runs on the ELEMENTS of the homopolymer-compressed sequence (k_hpc_* below: one
// element per run of equal unambiguous bases, one per ambiguous base; `seqs` holds their codes, `seq_off` their offsets, `epos` the
// position of every element's last base) -- the loop does on elements exactly what it does on positions without HPC, with the
// k-mer span = distance between the ends of the k-th previous element and this one, and only spans below 256 are valid.
template <typename HT, bool HPC>
__global__ void __launch_bounds__(kTile)
k_sketch(const unsigned char *__restrict__ seqs, const long long *__restrict__ seq_off, const int *__restrict__ tile_first, const int *__restrict__ tile_seq, int n_seq,
         int n_tiles, int w, int k, int rid_is_seq, u64 *__restrict__ ticket, int tile_begin, int tile_end, u64 *__restrict__ scan_state, long long cap, u64 *__restrict__ mv_x, u64 *__restrict__ mv_y,
         u32 *__restrict__ mv_seq, u64 *__restrict__ tile_excl, const u32 *__restrict__ epos)
{
    __shared__ SketchTile<HT> S;
    __shared__ u32 s_epos[HPC ? kHalo + kTile : 1];
    constexpr HT NONE = (HT)~(HT)0;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        const int tile = tile_begin + (int)atomicAdd(ticket, 1ULL);
        S.seq = tile_seq[tile];
        S.tile = tile;
    }
    __syncthreads();
    const int s = S.seq, tile = S.tile;
    const long long base = seq_off[s];
    const int len = (int)(seq_off[s + 1] - base);
    const int t0 = (tile - tile_first[s]) * kTile;
    // stage positions t0 - kHalo .. t0 + kTile - 1 (outside the sequence: ambiguous)
    for (int j = tid; j < kHalo + kTile; j += kTile) {
        const int pos = t0 - kHalo + j;
        const int c = (pos >= 0 && pos < len) ? nt4(__ldg(seqs + base + pos)) : 4;
        const u32 b0 = __ballot_sync(0xffffffffu, c & 1), b1 = __ballot_sync(0xffffffffu, c & 2), bn = __ballot_sync(0xffffffffu, c == 4);
        if (lane == 0) {
            S.pk[j >> 5] = spread32(__brev(b1)) << 1 | spread32(__brev(b0));
            S.nmask[j >> 5] = bn;
        }
        if (HPC) s_epos[j] = (pos >= 0 && pos < len) ? __ldg(epos + base + pos) : 0xffffffffu;   // before the first element: -1
    }
    __syncthreads();
    const u64 mask = (1ULL << (2 * k)) - 1;
    // hashes of positions t0 - w .. t0 + kTile - 1
    for (int j = tid; j < w + kTile; j += kTile) {
        const int sj = kHalo - w + j, pos = t0 - w + j;
        HT h = NONE;
        unsigned char z = 0;
        if (pos >= k - 1 && pos < len) {
            const int q = sj >> 5, r = sj & 31;
            // no ambiguous base among the last k positions: bits r-k+1 .. r of the mask words
            const u64 nm = ((u64)S.nmask[q] << 32 | S.nmask[q - 1]) >> (r + 1);          // bit 31 = position sj, bit 31 - t = position sj - t
            const bool ok = ((nm << 32 >> 32) >> (32 - k)) == 0;                          // k <= 28
            if (ok) {
                const int sft = 2 * (31 - r);
                const u64 f64 = sft ? (S.pk[q] >> sft) | (S.pk[q - 1] << (64 - sft)) : S.pk[q];
                const u64 f = f64 & mask;
                u64 y = __brevll(f);
                y = ((y & 0x5555555555555555ULL) << 1) | ((y >> 1) & 0x5555555555555555ULL);
                const u64 rc = ((~y) >> (64 - 2 * k)) & mask;
                if (f != rc) {
                    z = f < rc ? 0 : 1;
                    const u64 km = z ? rc : f;
                    h = sizeof(HT) == 4 ? (HT)hash32((u32)km, (u32)mask) : (HT)hash64(km, mask);
                    if (HPC) {       // the record compared by the loop is hash << 8 | span (sketch.c:109), and spans of 256 or more are no records
                        const u32 span = s_epos[sj] - s_epos[sj - k];
                        h = span < 256u ? (HT)((u64)h << 8 | span) : NONE;
                    }
                }
            }
        }
        S.ih[j] = h;
        S.iz[j] = z;
    }
    __syncthreads();
    // the rules of the loop for position i = t0 + tid
    const int i = t0 + tid, jj = w + tid;
    int l = 0;
    HT cur = NONE, mprev_x = NONE, mx = NONE;
    int mprev_p = -1, mp = -1;
    const bool in_range = i < len;
    int mode = 0;                                  // 2: rule P2, 3: rule P3
    const int T1 = w + k - 1;
    if (in_range) {
        // run of unambiguous bases ending at i (capped at 97, more than any threshold below)
        const int sj = kHalo + tid, q = sj >> 5, r = sj & 31;
        const u32 w2 = S.nmask[q] & (r == 31 ? 0xffffffffu : ((2u << r) - 1u));
        if (w2) l = r - (31 - __clz(w2));
        else {
            const u32 w1 = S.nmask[q - 1];
            if (w1) l = r + 1 + __clz(w1);
            else {
                const u32 w0 = S.nmask[q - 2];
                l = w0 ? r + 33 + __clz(w0) : 97;
            }
        }
        cur = S.ih[jj];
        for (int d = w; d >= 1; --d) {           // oldest -> newest, `<=` keeps the rightmost minimum
            const HT x = S.ih[jj - d];
            if (x <= mprev_x) mprev_x = x, mprev_p = i - d;
        }
        if (cur <= mprev_x) mode = 2;
        else if (mprev_p == i - w) {
            mode = 3;
            for (int d = w - 1; d >= 0; --d) {
                const HT x = S.ih[jj - d];
                if (x <= mx) mx = x, mp = i - d;
            }
        }
    }
    u64 wpos = 0;
    for (int pass = 0; pass < 2; ++pass) {       // the same rules twice: count, then (after the scans) write
        int c = 0;
        auto emit = [&](int p) {
            if (pass == 1 && (long long)(wpos + c) < cap) {
                const int q = w + (p - t0);
                mv_x[wpos + c] = HPC ? (u64)S.ih[q] : ((u64)S.ih[q] << 8 | (u64)k);
                mv_y[wpos + c] = (rid_is_seq ? (u64)s << 32 : 0ULL) | (u64)(HPC ? s_epos[kHalo + (p - t0)] : (u32)p) << 1 | (u64)S.iz[q];
                mv_seq[wpos + c] = (u32)s;
            }
            ++c;
        };
        if (in_range) {
            if (l == T1 && mprev_x != NONE)                                    // P1
                for (int d = w - 1; d >= 1; --d)
                    if (S.ih[jj - d] == mprev_x && i - d != mprev_p) emit(i - d);
            if (mode == 2) {                                                  // P2
                if (l >= T1 + 1 && mprev_x != NONE) emit(mprev_p);
            } else if (mode == 3) {                                           // P3
                if (l >= T1) emit(mprev_p);
                if (l >= T1 && mx != NONE)
                    for (int d = w - 1; d >= 0; --d)
                        if (S.ih[jj - d] == mx && i - d != mp) emit(i - d);
            }
            if (i == len - 1) {                                               // P4
                const HT fx = mode == 2 ? cur : mode == 3 ? mx : mprev_x;
                const int fp = mode == 2 ? i : mode == 3 ? mp : mprev_p;
                if (fx != NONE) emit(fp);
            }
        }
        if (pass == 0) {
            // block exclusive scan of the counts
            u32 v = (u32)c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
            if (lane == 31) S.warp_sum[wid] = v;
            __syncthreads();
            if (wid == 0) {
                u32 t = lane < kTile / 32 ? S.warp_sum[lane] : 0u;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
                if (lane < kTile / 32) S.warp_sum[lane] = t;
                // chained scan across tiles: publish the aggregate, look back for the prefix, publish the inclusive prefix
                const u64 tot = __shfl_sync(0xffffffffu, t, 31);
                volatile u64 *st = scan_state;
                if (lane == 0) {
                    __threadfence();
                    st[tile] = (tile == 0 ? MM2GB_FLAG_PREFIX : MM2GB_FLAG_AGG) | tot;
                }
                u64 excl = 0;
                if (tile > 0) {
                    int look = tile - 1;
                    for (;;) {
                        const int idx = look - lane;
                        u64 sv = idx >= 0 ? st[idx] : MM2GB_FLAG_PREFIX;
                        while (__any_sync(0xffffffffu, (sv >> 62) == 0)) sv = idx >= 0 ? st[idx] : MM2GB_FLAG_PREFIX;
                        const u32 pref = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
                        const int first = pref ? __ffs(pref) - 1 : 32;      // nearest predecessor holding a full prefix
                        u64 add = lane <= first ? (sv & MM2GB_VAL_MASK) : 0;
#pragma unroll
                        for (int o = 16; o >= 1; o >>= 1) add += __shfl_xor_sync(0xffffffffu, add, o);
                        excl += add;
                        if (pref) break;
                        look -= 32;
                    }
                    if (lane == 0) { __threadfence(); st[tile] = MM2GB_FLAG_PREFIX | (excl + tot); }
                }
                if (lane == 0) {
                    S.excl = excl;
                    tile_excl[tile] = excl;
                    if (tile == n_tiles - 1) tile_excl[n_tiles] = excl + tot;
                }
            }
            __syncthreads();
            wpos = S.excl + (v - (u32)c) + (wid ? S.warp_sum[wid - 1] : 0u);
        }
    }
}

// ---- the same for k <= 15 (hashes of at most 30 bits; map-ont), restructured for instruction count ---------------------------
// ncu of the kernel above (profiles/r6e_sketch_ncu.md): issue-bound, ~1000 warp instructions per 32 positions, most of them in
// the two w-step window scans and in evaluating the rules twice.  Here the window minima come from prefix / suffix minima over
// chunks of w positions (one thread per chunk; window [j-w+1, j] = suffix of one chunk + prefix of the next): keys
// hash << 11 | (2047 - local position) give the RIGHTMOST minimum, hash << 11 | local position the leftmost one -- the two differ
// exactly when the minimal hash occurs more than once in the window, the only case in which the "identical k-mer" loops of
// rules P1 / P3 can emit anything.  The rules are evaluated once; a position emits at most two records unless such duplicates
// exist (then it is re-evaluated when writing).
struct SketchTile32 {
    u64 pk[(kHalo + kTile) / 32];
    u32 nmask[(kHalo + kTile) / 32];
    u32 ih[kMaxW + kTile];
    unsigned char iz[kMaxW + kTile];
    u64 pre_r[kMaxW + kTile], suf_r[kMaxW + kTile], pre_l[kMaxW + kTile], suf_l[kMaxW + kTile];
    u32 warp_sum[kTile / 32];
    u64 excl;
    int seq, tile;
};

__device__ __forceinline__ int nt4_fast(u32 ch)
{
    const u32 up = ch & 0xDFu, t = up - 'A';
    const u32 x = (ch >> 1) & 3u;
    const bool letter = t < 26u && ((1u << t) & ((1u << 0) | (1u << 2) | (1u << 6) | (1u << 19) | (1u << 20)));
    return ch < 4u ? (int)ch : letter ? (int)(x ^ (x >> 1)) : 4;
}

__global__ void __launch_bounds__(kTile)
k_sketch32(const unsigned char *__restrict__ seqs, const long long *__restrict__ seq_off, const int *__restrict__ tile_first, const int *__restrict__ tile_seq, int n_seq,
           int n_tiles, int w, int k, int rid_is_seq, u64 *__restrict__ ticket, int tile_begin, int tile_end, u64 *__restrict__ scan_state, long long cap, u64 *__restrict__ mv_x, u64 *__restrict__ mv_y,
           u32 *__restrict__ mv_seq, u64 *__restrict__ tile_excl)
{
    __shared__ SketchTile32 S;
    constexpr u32 NONE = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        const int tile = tile_begin + (int)atomicAdd(ticket, 1ULL);
        S.seq = tile_seq[tile];
        S.tile = tile;
    }
    __syncthreads();
    const int s = S.seq, tile = S.tile;
    const long long base = seq_off[s];
    const int len = (int)(seq_off[s + 1] - base);
    const int t0 = (tile - tile_first[s]) * kTile;
    for (int j = tid; j < kHalo + kTile; j += kTile) {
        const int pos = t0 - kHalo + j;
        const int c = (pos >= 0 && pos < len) ? nt4_fast(__ldg(seqs + base + pos)) : 4;
        const u32 b0 = __ballot_sync(0xffffffffu, c & 1), b1 = __ballot_sync(0xffffffffu, c & 2), bn = __ballot_sync(0xffffffffu, c == 4);
        if (lane == 0) {
            S.pk[j >> 5] = spread32(__brev(b1)) << 1 | spread32(__brev(b0));
            S.nmask[j >> 5] = bn;
        }
    }
    __syncthreads();
    const u32 mask = (u32)((1ULL << (2 * k)) - 1);
    const int nloc = w + kTile;                    // local index 0 = position t0 - w
    for (int j = tid; j < nloc; j += kTile) {
        const int sj = kHalo - w + j, pos = t0 - w + j;
        u32 h = NONE;
        unsigned char z = 0;
        if (pos >= k - 1 && pos < len) {
            const int q = sj >> 5, r = sj & 31;
            const u64 nm = ((u64)S.nmask[q] << 32 | S.nmask[q - 1]) >> (r + 1);
            if ((((u32)nm) >> (32 - k)) == 0) {
                const int sft = 2 * (31 - r);
                const u64 f64 = sft ? (S.pk[q] >> sft) | (S.pk[q - 1] << (64 - sft)) : S.pk[q];
                const u32 f = (u32)f64 & mask;
                u32 y = __brev(f);
                y = ((y & 0x55555555u) << 1) | ((y >> 1) & 0x55555555u);
                const u32 rc = ((~y) >> (32 - 2 * k)) & mask;
                if (f != rc) {
                    z = f < rc ? 0 : 1;
                    h = hash32(z ? rc : f, mask);
                }
            }
        }
        S.ih[j] = h;
        S.iz[j] = z;
    }
    __syncthreads();
    // prefix / suffix minima over chunks of w local positions
    for (int cidx = tid; cidx * w < nloc; cidx += kTile) {
        const int c0 = cidx * w, c1 = min(c0 + w, nloc);
        u64 mr = ~0ULL, ml = ~0ULL;
        for (int j = c0; j < c1; ++j) {
            const u32 h = S.ih[j];
            const u64 kr = h == NONE ? ~0ULL : ((u64)h << 11 | (u64)(2047 - j)), kl = h == NONE ? ~0ULL : ((u64)h << 11 | (u64)j);
            mr = min(mr, kr); ml = min(ml, kl);
            S.pre_r[j] = mr; S.pre_l[j] = ml;
        }
        mr = ~0ULL; ml = ~0ULL;
        for (int j = c1 - 1; j >= c0; --j) {
            const u32 h = S.ih[j];
            const u64 kr = h == NONE ? ~0ULL : ((u64)h << 11 | (u64)(2047 - j)), kl = h == NONE ? ~0ULL : ((u64)h << 11 | (u64)j);
            mr = min(mr, kr); ml = min(ml, kl);
            S.suf_r[j] = mr; S.suf_l[j] = ml;
        }
    }
    __syncthreads();
    const int i = t0 + tid, jj = w + tid;
    const bool in_range = i < len;
    const int T1 = w + k - 1;
    int l = 0, mode = 0, mprev_j = -1, mp_j = -1;  // local indices of min(i-1) and (rule P3) min(i)
    u32 cur = NONE, mprev_x = NONE, mx = NONE;
    bool dup1 = false, dup0 = false;               // the minimal hash of window (i-1) / (i) occurs more than once
    if (in_range) {
        const int sj = kHalo + tid, q = sj >> 5, r = sj & 31;
        const u32 w2 = S.nmask[q] & (r == 31 ? 0xffffffffu : ((2u << r) - 1u));
        if (w2) l = r - (31 - __clz(w2));
        else {
            const u32 w1 = S.nmask[q - 1];
            if (w1) l = r + 1 + __clz(w1);
            else { const u32 w0 = S.nmask[q - 2]; l = w0 ? r + 33 + __clz(w0) : 97; }
        }
        cur = S.ih[jj];
        // window (i-1) = local [jj - w, jj - 1]; the rightmost minimum with `<=` semantics over all-NONE windows is its newest slot
        const u64 m1r = min(S.suf_r[jj - w], S.pre_r[jj - 1]), m1l = min(S.suf_l[jj - w], S.pre_l[jj - 1]);
        if (m1r != ~0ULL) { mprev_x = (u32)(m1r >> 11); mprev_j = 2047 - (int)(m1r & 2047u); dup1 = (int)(m1l & 2047u) != mprev_j; }
        else mprev_j = jj - 1;
        if (cur <= mprev_x) mode = 2;
        else if (mprev_j == jj - w) {
            mode = 3;
            const u64 m0r = min(S.suf_r[jj - w + 1], S.pre_r[jj]), m0l = min(S.suf_l[jj - w + 1], S.pre_l[jj]);
            if (m0r != ~0ULL) { mx = (u32)(m0r >> 11); mp_j = 2047 - (int)(m0r & 2047u); dup0 = (int)(m0l & 2047u) != mp_j; }
        }
    }
    auto rules = [&](auto &&emit) {                // emit(local index)
        if (!in_range) return;
        if (l == T1 && mprev_x != NONE && dup1)                                // P1
            for (int d = w - 1; d >= 1; --d)
                if (S.ih[jj - d] == mprev_x && jj - d != mprev_j) emit(jj - d);
        if (mode == 2) {                                                      // P2
            if (l >= T1 + 1 && mprev_x != NONE) emit(mprev_j);
        } else if (mode == 3) {                                               // P3
            if (l >= T1) emit(mprev_j);
            if (l >= T1 && mx != NONE && dup0)
                for (int d = w - 1; d >= 0; --d)
                    if (S.ih[jj - d] == mx && jj - d != mp_j) emit(jj - d);
        }
        if (i == len - 1) {                                                   // P4
            const u32 fx = mode == 2 ? cur : mode == 3 ? mx : mprev_x;
            const int fj = mode == 2 ? jj : mode == 3 ? mp_j : mprev_j;
            if (fx != NONE) emit(fj);
        }
    };
    int c = 0, e0 = 0, e1 = 0;
    rules([&](int j) { if (c == 0) e0 = j; else if (c == 1) e1 = j; ++c; });
    // block exclusive scan of the counts, chained scan across tiles
    u32 v = (u32)c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane == 31) S.warp_sum[wid] = v;
    __syncthreads();
    if (wid == 0) {
        u32 t = lane < kTile / 32 ? S.warp_sum[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
        if (lane < kTile / 32) S.warp_sum[lane] = t;
        const u64 tot = __shfl_sync(0xffffffffu, t, 31);
        volatile u64 *st = scan_state;
        if (lane == 0) { __threadfence(); st[tile] = (tile == 0 ? MM2GB_FLAG_PREFIX : MM2GB_FLAG_AGG) | tot; }
        u64 excl = 0;
        if (tile > 0) {
            int look = tile - 1;
            for (;;) {
                const int idx = look - lane;
                u64 sv = idx >= 0 ? st[idx] : MM2GB_FLAG_PREFIX;
                while (__any_sync(0xffffffffu, (sv >> 62) == 0)) sv = idx >= 0 ? st[idx] : MM2GB_FLAG_PREFIX;
                const u32 pref = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
                const int first = pref ? __ffs(pref) - 1 : 32;
                u64 add = lane <= first ? (sv & MM2GB_VAL_MASK) : 0;
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) add += __shfl_xor_sync(0xffffffffu, add, o);
                excl += add;
                if (pref) break;
                look -= 32;
            }
            if (lane == 0) { __threadfence(); st[tile] = MM2GB_FLAG_PREFIX | (excl + tot); }
        }
        if (lane == 0) {
            S.excl = excl;
            tile_excl[tile] = excl;
            if (tile == n_tiles - 1) tile_excl[n_tiles] = excl + tot;
        }
    }
    __syncthreads();
    const u64 wpos = S.excl + (v - (u32)c) + (wid ? S.warp_sum[wid - 1] : 0u);
    const u64 ridbits = rid_is_seq ? (u64)s << 32 : 0ULL;
    auto put = [&](u64 at, int j) {
        if ((long long)at >= cap) return;
        mv_x[at] = (u64)S.ih[j] << 8 | (u64)k;
        mv_y[at] = ridbits | (u64)(u32)(t0 - w + j) << 1 | (u64)S.iz[j];
        mv_seq[at] = (u32)s;
    };
    if (c <= 2) {
        if (c >= 1) put(wpos, e0);
        if (c == 2) put(wpos + 1, e1);
    } else {
        int n = 0;
        rules([&](int j) { put(wpos + n, j); ++n; });
    }
}

// ---- k_sketch32 as a persistent, software-pipelined kernel --------------------------------------------------------------------
// ncu of k_sketch32 (profiles/r6h_sketch32_ncu.md): 35 % of the stall samples sit behind the look-back of the chained scan (the whole
// CTA waits for one warp's global round trips), 16 % behind the ticket at the start of every tile.  Here a CTA loops over tickets;
// the ticket of the next tile is fetched while the current one is computed, and the WRITE phase of a tile is deferred by one
// iteration: the CTA publishes the tile's count, computes the next tile, and only then resolves the first tile's prefix -- by
// then its predecessors have published theirs, so the look-back (done by the whole CTA at once: one status word per thread, up
// to kTile tiles back in one round trip) does not wait.  Every CTA publishes the aggregate of a tile before it waits for anything,
// so ticket order guarantees progress.  A tile in which some position emits more than two records (duplicate minimal k-mers)
// is written at once instead of deferred (its rules are re-evaluated when writing).
struct SketchTile32P {
    u64 pk[(kHalo + kTile) / 32];
    u32 nmask[(kHalo + kTile) / 32];
    u32 ih[2][kMaxW + kTile];
    unsigned char iz[2][kMaxW + kTile];
    u64 pre_r[kMaxW + kTile], suf_r[kMaxW + kTile], pre_l[kMaxW + kTile], suf_l[kMaxW + kTile];
    u32 warp_sum[2][kTile / 32];
    u64 red[kTile / 32];
    int first[kTile / 32];
    u64 excl;
    int ticket[2];
};

// prefix of tile `tile` (> 0) by the whole CTA; publishes it; returns it to every thread.  `tot` = the tile's own count.
__device__ __forceinline__ u64 cta_lookback(volatile u64 *st, int tile, u64 tot, SketchTile32P &S)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    u64 excl = 0;
    int look = tile - 1;
    for (;;) {
        const int idx = look - tid;
        u64 sv = idx >= 0 ? st[idx] : MM2GB_FLAG_PREFIX;
        while ((sv >> 62) == 0) sv = st[idx];
        const u32 pref = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
        if (lane == 0) S.first[wid] = pref ? wid * 32 + __ffs(pref) - 1 : 1 << 30;
        __syncthreads();
        int first = 1 << 30;
#pragma unroll
        for (int q = 0; q < kTile / 32; ++q) first = min(first, S.first[q]);
        u64 add = tid <= first ? (sv & MM2GB_VAL_MASK) : 0;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) add += __shfl_xor_sync(0xffffffffu, add, o);
        if (lane == 0) S.red[wid] = add;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kTile / 32; ++q) excl += S.red[q];
        __syncthreads();
        if (first != 1 << 30) break;
        look -= kTile;
    }
    if (tid == 0) { __threadfence(); st[tile] = MM2GB_FLAG_PREFIX | (excl + tot); }
    return excl;
}

#ifndef MM2GB_SKETCHP_MIN_CTAS
#define MM2GB_SKETCHP_MIN_CTAS 3
#endif
#ifndef MM2GB_SKETCHP_V2
#define MM2GB_SKETCHP_V2 1     // 0: the r7 form of the two phases marked below (A/B: profiles/r8d_sketch_ab.txt)
#endif
__global__ void __launch_bounds__(kTile, MM2GB_SKETCHP_MIN_CTAS)
k_sketch32p(const unsigned char *__restrict__ seqs, const long long *__restrict__ seq_off, const int *__restrict__ tile_first, const int *__restrict__ tile_seq,
            int n_seq, int n_tiles, int w, int k, int rid_is_seq, u64 *__restrict__ ticket, int tile_begin, int tile_end, u64 *__restrict__ scan_state, long long cap, u64 *__restrict__ mv_x,
            u64 *__restrict__ mv_y, u32 *__restrict__ mv_seq, u64 *__restrict__ tile_excl)
{
    __shared__ SketchTile32P S;
    constexpr u32 NONE = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    volatile u64 *st = scan_state;
    const u32 mask = (u32)((1ULL << (2 * k)) - 1);
    const int nloc = w + kTile, T1 = w + k - 1;
    if (tid == 0) S.ticket[0] = tile_begin + (int)atomicAdd(ticket, 1ULL);
    // the deferred tile
    bool have_prev = false;
    int p_tile = 0, p_seq = 0, p_t0 = 0, p_c = 0, p_e0 = 0, p_e1 = 0, p_buf = 0;
    u32 p_off = 0, p_tot = 0;
    auto put = [&](int buf, int sq, int t0, u64 at, int j) {
        if ((long long)at >= cap) return;
        mv_x[at] = (u64)S.ih[buf][j] << 8 | (u64)k;
        mv_y[at] = (rid_is_seq ? (u64)sq << 32 : 0ULL) | (u64)(u32)(t0 - w + j) << 1 | (u64)S.iz[buf][j];
        mv_seq[at] = (u32)sq;
    };
    for (int it = 0;; ++it) {
        const int buf = it & 1;
        __syncthreads();
        const int tile = S.ticket[buf];
        const bool valid = tile < tile_end;
        if (tid == 0 && valid) S.ticket[buf ^ 1] = tile_begin + (int)atomicAdd(ticket, 1ULL);     // in flight while this tile is computed
        int s = 0, t0 = 0, c = 0, e0 = 0, e1 = 0, i = 0, jj = 0, l = 0, mode = 0, mprev_j = -1, mp_j = -1, len = 0;
        u32 cur = NONE, mprev_x = NONE, mx = NONE, v = 0, tot = 0;
        bool dup1 = false, dup0 = false, in_range = false;
        auto rules = [&](auto &&emit) {                // emit(local index)
            if (!in_range) return;
            if (l == T1 && mprev_x != NONE && dup1)                                // P1
                for (int d = w - 1; d >= 1; --d)
                    if (S.ih[buf][jj - d] == mprev_x && jj - d != mprev_j) emit(jj - d);
            if (mode == 2) {                                                      // P2
                if (l >= T1 + 1 && mprev_x != NONE) emit(mprev_j);
            } else if (mode == 3) {                                               // P3
                if (l >= T1) emit(mprev_j);
                if (l >= T1 && mx != NONE && dup0)
                    for (int d = w - 1; d >= 0; --d)
                        if (S.ih[buf][jj - d] == mx && jj - d != mp_j) emit(jj - d);
            }
            if (i == len - 1) {                                                   // P4
                const u32 fx = mode == 2 ? cur : mode == 3 ? mx : mprev_x;
                const int fj = mode == 2 ? jj : mode == 3 ? mp_j : mprev_j;
                if (fx != NONE) emit(fj);
            }
        };
        if (valid) {
            s = tile_seq[tile];
            const long long base = seq_off[s];
            len = (int)(seq_off[s + 1] - base);
            t0 = (tile - tile_first[s]) * kTile;
            // staging: thread t < (kHalo + kTile) / 4 packs four consecutive bases (SWAR: all four are letters of ACGTU in the common
            // case), eight neighbouring lanes assemble a 32-base word
            if (wid < (kHalo + kTile + 127) / 128) {
                const int j4 = 4 * tid;
                u32 pk8 = 0, n4 = 0;
                if (j4 < kHalo + kTile) {
                    const int pos0 = t0 - kHalo + j4;
                    u32 wv = 0x4E4E4E4Eu;                                   // "NNNN"
                    if (pos0 >= 0 && pos0 + 3 < len) {
                        const unsigned char *p = seqs + base + pos0;
                        wv = (u32)__ldg(p) | (u32)__ldg(p + 1) << 8 | (u32)__ldg(p + 2) << 16 | (u32)__ldg(p + 3) << 24;
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (pos0 + q >= 0 && pos0 + q < len) wv = (wv & ~(0xffu << (8 * q))) | (u32)__ldg(seqs + base + pos0 + q) << (8 * q);
                    }
                    const u32 up = wv & 0xDFDFDFDFu;
                    u32 ok = 0;
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        const u32 cc = (q == 0 ? 0x41u : q == 1 ? 0x43u : q == 2 ? 0x47u : q == 3 ? 0x54u : 0x55u) * 0x01010101u;
                        const u32 v = up ^ cc;
                        ok |= ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;   // bit 7 of every byte that is zero
                    }
                    u32 code;
                    if (ok == 0x80808080u) {
                        const u32 x = (wv >> 1) & 0x03030303u;
                        code = x ^ ((x >> 1) & 0x01010101u);
                    } else {                                                // an ambiguous base (or the table's 0..3 bytes): one by one
                        code = 0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int c1 = nt4_fast((wv >> (8 * q)) & 0xffu);
                            code |= (u32)(c1 & 3) << (8 * q);
                            n4 |= (c1 == 4 ? 1u : 0u) << q;
                        }
                    }
                    pk8 = (code * 0x40100401u) >> 24;                       // b0 << 6 | b1 << 4 | b2 << 2 | b3
                }
                u64 v64 = (u64)pk8 << (8 * (7 - (lane & 7)));
                u32 nm = n4 << (4 * (lane & 7));
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) { v64 |= __shfl_xor_sync(0xffffffffu, v64, o); nm |= __shfl_xor_sync(0xffffffffu, nm, o); }
                if ((lane & 7) == 0 && j4 < kHalo + kTile) { S.pk[j4 >> 5] = v64; S.nmask[j4 >> 5] = nm; }
            }
            __syncthreads();
            for (int j = tid; j < nloc; j += kTile) {
                const int sj = kHalo - w + j, pos = t0 - w + j;
                u32 h = NONE;
                unsigned char z = 0;
                if (pos >= k - 1 && pos < len) {
                    const int q = sj >> 5, r = sj & 31;
                    const u64 nm = ((u64)S.nmask[q] << 32 | S.nmask[q - 1]) >> (r + 1);
                    if ((((u32)nm) >> (32 - k)) == 0) {
                        const int sft = 2 * (31 - r);
                        const u64 f64 = sft ? (S.pk[q] >> sft) | (S.pk[q - 1] << (64 - sft)) : S.pk[q];
                        const u32 f = (u32)f64 & mask;
                        u32 y = __brev(f);
                        y = ((y & 0x55555555u) << 1) | ((y >> 1) & 0x55555555u);
                        const u32 rc = ((~y) >> (32 - 2 * k)) & mask;
                        if (f != rc) {
                            z = f < rc ? 0 : 1;
                            h = hash32(z ? rc : f, mask);
                        }
                    }
                }
                S.ih[buf][j] = h;
                S.iz[buf][j] = z;
            }
            __syncthreads();
#if MM2GB_SKETCHP_V2
            {   // the four running minima (prefix / suffix x rightmost / leftmost key) of a chunk on four different warps: the CTA waits
                // for w dependent steps instead of 2 w steps of twice the work (ncu r8c: 18 % of the stall samples behind this barrier)
                const int nch = (nloc + w - 1) / w, R = (nch + 31) & ~31;
                for (int idx = tid; idx < 4 * R; idx += kTile) {
                    const int role = idx / R, cidx = idx - role * R;       // warp-uniform: R and kTile are multiples of 32
                    if (cidx >= nch) continue;
                    const int c0 = cidx * w, c1 = min(c0 + w, nloc);
                    u64 *dst = role == 0 ? S.pre_r : role == 1 ? S.pre_l : role == 2 ? S.suf_r : S.suf_l;
                    const bool right = (role & 1) == 0;
                    u64 m = ~0ULL;
                    if (role < 2) {
                        for (int j = c0; j < c1; ++j) {
                            const u32 h = S.ih[buf][j];
                            const u64 key = h == NONE ? ~0ULL : ((u64)h << 11 | (u64)(right ? 2047 - j : j));
                            m = min(m, key);
                            dst[j] = m;
                        }
                    } else {
                        for (int j = c1 - 1; j >= c0; --j) {
                            const u32 h = S.ih[buf][j];
                            const u64 key = h == NONE ? ~0ULL : ((u64)h << 11 | (u64)(right ? 2047 - j : j));
                            m = min(m, key);
                            dst[j] = m;
                        }
                    }
                }
            }
#else
            for (int cidx = tid; cidx * w < nloc; cidx += kTile) {
                const int c0 = cidx * w, c1 = min(c0 + w, nloc);
                u64 mr = ~0ULL, ml = ~0ULL;
                for (int j = c0; j < c1; ++j) {
                    const u32 h = S.ih[buf][j];
                    const u64 kr = h == NONE ? ~0ULL : ((u64)h << 11 | (u64)(2047 - j)), kl = h == NONE ? ~0ULL : ((u64)h << 11 | (u64)j);
                    mr = min(mr, kr); ml = min(ml, kl);
                    S.pre_r[j] = mr; S.pre_l[j] = ml;
                }
                mr = ~0ULL; ml = ~0ULL;
                for (int j = c1 - 1; j >= c0; --j) {
                    const u32 h = S.ih[buf][j];
                    const u64 kr = h == NONE ? ~0ULL : ((u64)h << 11 | (u64)(2047 - j)), kl = h == NONE ? ~0ULL : ((u64)h << 11 | (u64)j);
                    mr = min(mr, kr); ml = min(ml, kl);
                    S.suf_r[j] = mr; S.suf_l[j] = ml;
                }
            }
#endif
            __syncthreads();
            i = t0 + tid; jj = w + tid;
            in_range = i < len;
            if (in_range) {
                const int sj = kHalo + tid, q = sj >> 5, r = sj & 31;
                const u32 w2 = S.nmask[q] & (r == 31 ? 0xffffffffu : ((2u << r) - 1u));
                if (w2) l = r - (31 - __clz(w2));
                else {
                    const u32 w1 = S.nmask[q - 1];
                    if (w1) l = r + 1 + __clz(w1);
                    else { const u32 w0 = S.nmask[q - 2]; l = w0 ? r + 33 + __clz(w0) : 97; }
                }
                cur = S.ih[buf][jj];
                const u64 m1r = min(S.suf_r[jj - w], S.pre_r[jj - 1]), m1l = min(S.suf_l[jj - w], S.pre_l[jj - 1]);
                if (m1r != ~0ULL) { mprev_x = (u32)(m1r >> 11); mprev_j = 2047 - (int)(m1r & 2047u); dup1 = (int)(m1l & 2047u) != mprev_j; }
                else mprev_j = jj - 1;
                if (cur <= mprev_x) mode = 2;
                else if (mprev_j == jj - w) {
                    mode = 3;
                    const u64 m0r = min(S.suf_r[jj - w + 1], S.pre_r[jj]), m0l = min(S.suf_l[jj - w + 1], S.pre_l[jj]);
                    if (m0r != ~0ULL) { mx = (u32)(m0r >> 11); mp_j = 2047 - (int)(m0r & 2047u); dup0 = (int)(m0l & 2047u) != mp_j; }
                }
            }
            rules([&](int j) { if (c == 0) e0 = j; else if (c == 1) e1 = j; ++c; });
            v = (u32)c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
            if (lane == 31) S.warp_sum[buf][wid] = v;
        }
        const bool many = __syncthreads_or(valid && c > 2);
#if MM2GB_SKETCHP_V2
        if (valid) {
            // every warp scans the kTile / 32 warp totals itself (they are final after the barrier above): no second barrier, no
            // round trip through shared memory (ncu r8c: 10 % of the stall samples sat behind it); warp 0 publishes the aggregate
            u32 t = lane < kTile / 32 ? S.warp_sum[buf][lane] : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
            tot = __shfl_sync(0xffffffffu, t, 31);
            const u32 before = __shfl_sync(0xffffffffu, t, (wid + 31) & 31);
            if (wid == 0 && lane == 31) { __threadfence(); st[tile] = (tile == 0 ? MM2GB_FLAG_PREFIX : MM2GB_FLAG_AGG) | (u64)tot; }
            v = v - (u32)c + (wid ? before : 0u);                        // exclusive offset of this thread inside the tile
        }
#else
        if (valid) {
            if (wid == 0) {
                u32 t = lane < kTile / 32 ? S.warp_sum[buf][lane] : 0u;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
                if (lane < kTile / 32) S.warp_sum[buf][lane] = t;
                if (lane == 31) { __threadfence(); st[tile] = (tile == 0 ? MM2GB_FLAG_PREFIX : MM2GB_FLAG_AGG) | (u64)t; }
            }
            __syncthreads();
            tot = S.warp_sum[buf][kTile / 32 - 1];
            v = v - (u32)c + (wid ? S.warp_sum[buf][wid - 1] : 0u);      // exclusive offset of this thread inside the tile
        }
#endif
        // finish the deferred tile: its predecessors have had a whole tile's time to publish
        if (have_prev) {
            const u64 excl = p_tile ? cta_lookback(st, p_tile, p_tot, S) : 0ULL;
            if (tid == 0) { tile_excl[p_tile] = excl; if (p_tile == n_tiles - 1) tile_excl[n_tiles] = excl + p_tot; }
            if (p_c >= 1) put(p_buf, p_seq, p_t0, excl + p_off, p_e0);
            if (p_c == 2) put(p_buf, p_seq, p_t0, excl + p_off + 1, p_e1);
            have_prev = false;
        }
        if (!valid) break;
        if (many) {                                   // written at once: the rules are re-evaluated with the state still in registers
            const u64 excl = tile ? cta_lookback(st, tile, tot, S) : 0ULL;
            if (tid == 0) { tile_excl[tile] = excl; if (tile == n_tiles - 1) tile_excl[n_tiles] = excl + tot; }
            int n = 0;
            rules([&](int j) { put(buf, s, t0, excl + v + n, j); ++n; });
        } else {
            have_prev = true;
            p_tile = tile; p_seq = s; p_t0 = t0; p_c = c; p_e0 = e0; p_e1 = e1; p_buf = buf; p_off = v; p_tot = tot;
        }
    }
}

// ---- homopolymer compression of a batch (sketch.c:92-101 as a stream compaction) ------------------------------------------------
// flag[g] = 1 if an element ends at global base g: the last base of a run of equal unambiguous bases, or an ambiguous base.
__global__ void __launch_bounds__(256)
k_hpc_flags(const unsigned char *__restrict__ seqs, const long long *__restrict__ seq_off, const int *__restrict__ tile_first,
            const int *__restrict__ tile_seq, int n_tiles, unsigned char *__restrict__ flag)
{
    const int tile = blockIdx.x;
    if (tile >= n_tiles) return;
    const int s = tile_seq[tile];
    const long long base = seq_off[s];
    const int len = (int)(seq_off[s + 1] - base);
    const int t0 = (tile - tile_first[s]) * kTile;
    for (int i = t0 + threadIdx.x; i < min(len, t0 + kTile); i += blockDim.x) {
        const int c = nt4(__ldg(seqs + base + i));
        const int cn = i + 1 < len ? nt4(__ldg(seqs + base + i + 1)) : 5;
        flag[base + i] = (c == 4 || cn != c) ? 1 : 0;
    }
}

// element e = eidx[g] of every flagged base g: its code and the position (inside its sequence) of its last base
__global__ void __launch_bounds__(256)
k_hpc_write(const unsigned char *__restrict__ seqs, const long long *__restrict__ seq_off, const int *__restrict__ tile_first,
            const int *__restrict__ tile_seq, int n_tiles, const unsigned char *__restrict__ flag, const u64 *__restrict__ eidx,
            unsigned char *__restrict__ ecode, u32 *__restrict__ epos)
{
    const int tile = blockIdx.x;
    if (tile >= n_tiles) return;
    const int s = tile_seq[tile];
    const long long base = seq_off[s];
    const int len = (int)(seq_off[s + 1] - base);
    const int t0 = (tile - tile_first[s]) * kTile;
    for (int i = t0 + threadIdx.x; i < min(len, t0 + kTile); i += blockDim.x) {
        if (!flag[base + i]) continue;
        const u64 e = eidx[base + i];
        ecode[e] = (unsigned char)nt4(__ldg(seqs + base + i));
        epos[e] = (u32)i;
    }
}

// element offsets per sequence
__global__ void k_hpc_offsets(const long long *__restrict__ seq_off, const u64 *__restrict__ eidx, int n_seq, long long *__restrict__ eoff)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s <= n_seq) eoff[s] = (long long)eidx[seq_off[s]];
}

// ---- exclusive scan of a u32 array into u64 (three kernels; n up to 2^40) -------------------------------------------------
constexpr int kScanThreads = 256, kScanItems = 16, kScanChunk = kScanThreads * kScanItems;

__device__ __forceinline__ u64 block_scan_excl(u64 v, u64 *warp_sum, u64 *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    u64 x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u64 t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += t; }
    if (lane == 31) warp_sum[wid] = x;
    __syncthreads();
    if (wid == 0) {
        u64 t = lane < nw ? warp_sum[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u64 y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
        if (lane < nw) warp_sum[lane] = t;
    }
    __syncthreads();
    const u64 excl = x - v + (wid ? warp_sum[wid - 1] : 0);
    if (total) *total = warp_sum[nw - 1];
    __syncthreads();
    return excl;
}

template <typename TI>
__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const TI *__restrict__ in, long long n, u64 *__restrict__ part)
{
    __shared__ u64 ws[32];
    const long long b0 = (long long)blockIdx.x * kScanChunk;
    u64 s = 0;
    for (int t = 0; t < kScanItems; ++t) {
        const long long i = b0 + (long long)t * kScanThreads + threadIdx.x;
        if (i < n) s += in[i];
    }
    u64 tot;
    block_scan_excl(s, ws, &tot);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}

// one CTA: exclusive scan of the partial sums in place; total -> part[n_part]
__global__ void __launch_bounds__(1024) k_scan_top(u64 *__restrict__ part, int n_part)
{
    __shared__ u64 ws[32];
    __shared__ u64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b = 0; b < n_part; b += 1024) {
        const int i = b + threadIdx.x;
        const u64 v = i < n_part ? part[i] : 0;
        u64 tot;
        const u64 e = block_scan_excl(v, ws, &tot);
        if (i < n_part) part[i] = carry + e;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) part[n_part] = carry;
}

template <typename TI>
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const TI *__restrict__ in, long long n, const u64 *__restrict__ part,
                                                             u64 *__restrict__ out)
{
    __shared__ u64 ws[32];
    const long long b0 = (long long)blockIdx.x * kScanChunk + (long long)threadIdx.x * kScanItems;
    u32 v[kScanItems];
    u64 s = 0;
#pragma unroll
    for (int t = 0; t < kScanItems; ++t) { const long long i = b0 + t; v[t] = i < n ? in[i] : 0u; s += v[t]; }
    u64 e = block_scan_excl(s, ws, nullptr) + part[blockIdx.x];
#pragma unroll
    for (int t = 0; t < kScanItems; ++t) { const long long i = b0 + t; if (i < n) out[i] = e; e += v[t]; }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) out[n] = part[gridDim.x];
}

// tile -> sequence (one thread per sequence fills the entries of its tiles)
__global__ void k_tile_map(const int *__restrict__ tile_first, int n_seq, int *__restrict__ tile_seq)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seq) return;
    for (int t = tile_first[s]; t < tile_first[s + 1]; ++t) tile_seq[t] = s;
}

// ---- per-sequence minimizer offsets -------------------------------------------------------------------------------------
__global__ void k_seq_mv_off(const int *__restrict__ tile_first, const u64 *__restrict__ tile_base, int n_seq, u64 *__restrict__ mv_off)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s <= n_seq) mv_off[s] = tile_base[tile_first[s]];
}

// ---- mm_seed_mz_flt (seed.c:5-29) ---------------------------------------------------------------------------------------
// Read s counts its minimizer values (the full x, span included) in the table region [2 * mv_off[s], 2 * mv_off[s+1]).
__device__ __forceinline__ u64 mix64(u64 h)
{
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
    return h;
}

__global__ void k_qocc_count(const u64 *__restrict__ mv_x, const u32 *__restrict__ mv_seq, const u64 *__restrict__ mv_off, long long n_mv,
                             int q_occ_max, u64 *__restrict__ tab_key, u32 *__restrict__ tab_cnt)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_mv) return;
    const u32 s = mv_seq[i];
    const u64 b = mv_off[s], n = mv_off[s + 1] - b;
    if (n <= (u64)q_occ_max) return;                 // seed.c:9: nothing is filtered in this read
    const u64 m = 2 * n, x = mv_x[i];
    u64 h = mix64(x) % m;
    for (;;) {
        const u64 old = atomicCAS(tab_key + 2 * b + h, kNone, x);
        if (old == kNone || old == x) { atomicAdd(tab_cnt + 2 * b + h, 1u); return; }
        if (++h == m) h = 0;
    }
}

// keep[i] = 1 unless seed.c:17-19 zeroes the minimizer
__global__ void k_qocc_flag(const u64 *__restrict__ mv_x, const u32 *__restrict__ mv_seq, const u64 *__restrict__ mv_off, long long n_mv,
                            int q_occ_max, float q_occ_frac, const u64 *__restrict__ tab_key, const u32 *__restrict__ tab_cnt,
                            unsigned char *__restrict__ keep)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_mv) return;
    const u32 s = mv_seq[i];
    const u64 b = mv_off[s], n = mv_off[s + 1] - b;
    unsigned char kp = 1;
    if (n > (u64)q_occ_max) {
        const u64 m = 2 * n, x = mv_x[i];
        u64 h = mix64(x) % m;
        while (tab_key[2 * b + h] != x) if (++h == m) h = 0;
        const int cnt = (int)tab_cnt[2 * b + h];
        // `cnt > mv->n * q_occ_frac`: size_t * float -> float, int -> float
        if (cnt > q_occ_max && (float)cnt > __fmul_rn((float)n, q_occ_frac)) kp = 0;
    }
    keep[i] = kp;
}

// ---- mm_seed_collect_all (seed.c:31-53) ---------------------------------------------------------------------------------
// Index: open addressing, slot = (key, off << 28 | cnt), empty key = ~0.  Per minimizer: n_occ (0: dropped or not in the index),
// offset of its occurrence list, is_tandem.
struct DevIndex {
    const u64 *key;
    const u64 *val;
    const u64 *occ;
    u64 mask;        // slots - 1
};

__device__ __forceinline__ bool index_get(const DevIndex &ix, u64 minier, u64 *off, u32 *n)
{
    u64 h = mix64(minier) & ix.mask;
    for (;;) {
        const u64 kk = __ldg(ix.key + h);
        if (kk == minier) { const u64 v = __ldg(ix.val + h); *off = v >> 28; *n = (u32)(v & 0xfffffffu); return true; }
        if (kk == kNone) return false;
        h = (h + 1) & ix.mask;
    }
}

__global__ void k_lookup(DevIndex ix, const u64 *__restrict__ mv_x, const u32 *__restrict__ mv_seq, const unsigned char *__restrict__ keep,
                         long long n_mv, u32 *__restrict__ occ_n, u64 *__restrict__ occ_off, unsigned char *__restrict__ tandem,
                         u32 *__restrict__ has)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_mv) return;
    u32 n = 0;
    u64 off = 0;
    unsigned char td = 0;
    if (keep[i]) {
        const u64 x = mv_x[i], mz = x >> 8;
        const u32 s = mv_seq[i];
        if (index_get(ix, mz, &off, &n) && n > 0) {
            // neighbours in the filtered minimizer list of the same read (seed.c:48-49)
            long long j = i - 1;
            while (j >= 0 && mv_seq[j] == s && !keep[j]) --j;
            if (j >= 0 && mv_seq[j] == s && (mv_x[j] >> 8) == mz) td = 1;
            j = i + 1;
            while (j < n_mv && mv_seq[j] == s && !keep[j]) ++j;
            if (j < n_mv && mv_seq[j] == s && (mv_x[j] >> 8) == mz) td = 1;
        } else n = 0;
    }
    occ_n[i] = n; occ_off[i] = off; tandem[i] = td; has[i] = n > 0 ? 1u : 0u;
}

// seeds (the m[] array of seed.c) compacted in order
struct Seeds {
    u32 *n;          // occurrences
    u32 *q_pos;      // pos << 1 | strand
    u64 *off;        // occurrence list
    u32 *seq;        // read
    unsigned char *tandem;
    unsigned char *flt;
    unsigned char *span;     // q_span = the low byte of the minimizer record (k unless the index is homopolymer-compressed)
};

__global__ void k_compact_seeds(const u64 *__restrict__ mv_x, const u64 *__restrict__ mv_y, const u32 *__restrict__ mv_seq, const u32 *__restrict__ occ_n,
                                const u64 *__restrict__ occ_off, const unsigned char *__restrict__ tandem, const u64 *__restrict__ m_idx,
                                long long n_mv, Seeds m)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_mv || occ_n[i] == 0) return;
    const u64 q = m_idx[i];
    m.n[q] = occ_n[i]; m.q_pos[q] = (u32)mv_y[i]; m.off[q] = occ_off[i]; m.seq[q] = mv_seq[i]; m.tandem[q] = tandem[i];
    m.span[q] = (unsigned char)(mv_x[i] & 0xffu);
}

// ---- mm_seed_select + the filter / rep_len part of mm_collect_matches -----------------------------------------------------
// m_off[s] = first seed of read s.  Output per seed: flt, cnt_a (= n if kept else 0), kept (0/1); rep_len[s] accumulated.
__global__ void k_select(Seeds m, const u64 *__restrict__ m_idx, const u64 *__restrict__ mv_off, const long long *__restrict__ seq_off,
                         long long n_m, int max_occ, int max_max_occ, int dist, int q_span, u32 *__restrict__ cnt_a, u32 *__restrict__ kept)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_m) return;
    const u32 s = m.seq[i];
    const long long b = (long long)m_idx[mv_off[s]], e = (long long)m_idx[mv_off[s + 1]];   // seeds of the read: [b, e)
    const int n0 = (int)(e - b);
    const u32 ni = m.n[i];
    unsigned char flt = 0;
    if (dist > 0 && max_max_occ > max_occ) {
        // seed.c:65-69: nothing happens with fewer than two seeds or without a high-occurrence seed; the second test is
        // implied for a seed that is not high itself (its flt stays 0 either way)
        if (n0 >= 2 && ni > (u32)max_occ) {
            flt = 1;
            if (ni <= (u32)max_max_occ) {          // seed.c:91-93: above max_max_occ the seed is filtered whatever its rank
                long long st = i, en = i + 1;
                int rank = 0;
                // a seed with 128 (MAX_MAX_HIGH_OCC, seed.c:55) smaller ones in its streak is filtered whatever the streak's bounds
                // are: stop scanning there (a read that is one long repeat is one long streak)
                while (rank < 128 && st > b && m.n[st - 1] > (u32)max_occ) { --st; if (m.n[st] <= ni) ++rank; }      // (n, j) < (ni, i) with j < i
                while (rank < 128 && en < e && m.n[en] > (u32)max_occ) { if (m.n[en] < ni) ++rank; ++en; }           // j > i: strictly smaller n
                if (rank < 128) {
                    const int len = (int)(seq_off[s + 1] - seq_off[s]);
                    const int ps = st > b ? (int)(m.q_pos[st - 1] >> 1) : 0;
                    const int pe = en < e ? (int)(m.q_pos[en] >> 1) : len;
                    int max_high_occ = (int)((double)(pe - ps) / dist + .499);
                    if (max_high_occ > 128) max_high_occ = 128;
                    flt = (max_high_occ > 0 && rank < max_high_occ) ? 0 : 1;
                }
            }
        }
    } else if (ni > (u32)max_occ) flt = 1;
    m.flt[i] = flt;
    cnt_a[i] = flt ? 0u : ni;
    kept[i] = flt ? 0u : 1u;
}

// rep_len (seed.c:117-121,128): the filtered seeds of a read in order; each adds what its interval [en - q_span, en) extends
// beyond the previous filtered seed's end (positions increase along the read)
__global__ void k_rep_len(Seeds m, const u64 *__restrict__ m_idx, const u64 *__restrict__ mv_off, long long n_m, int q_span,
                          int *__restrict__ rep_len)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_m || !m.flt[i]) return;
    const u32 s = m.seq[i];
    const long long b = (long long)m_idx[mv_off[s]];
    (void)q_span;
    const int en = (int)(m.q_pos[i] >> 1) + 1, st = en - (int)m.span[i];
    long long j = i - 1;
    while (j >= b && !m.flt[j]) --j;
    const int prev_en = j >= b ? (int)(m.q_pos[j] >> 1) + 1 : 0;
    const int add = en - (st > prev_en ? st : prev_en);
    if (add) atomicAdd(rep_len + s, add);
}

// per read: anchor offsets, mini_pos offsets
__global__ void k_read_offsets(const u64 *__restrict__ m_idx, const u64 *__restrict__ mv_off, const u64 *__restrict__ a_pos,
                               const u64 *__restrict__ mp_pos, int n_reads, long long *__restrict__ a_off, long long *__restrict__ mp_off)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_reads) return;
    const u64 q = m_idx[mv_off[s]];
    a_off[s] = (long long)a_pos[q];
    mp_off[s] = (long long)mp_pos[q];
}

// ---- anchors (map.c:303-325) ----------------------------------------------------------------------------------------------
constexpr u64 kSeedTandem = 1ULL << 42;   // mmpriv.h:20

__global__ void k_expand(Seeds m, const u64 *__restrict__ occ, const u64 *__restrict__ a_pos, const u64 *__restrict__ mp_pos,
                         const long long *__restrict__ seq_off, long long n_m, int q_span, uint4 *__restrict__ a, u64 *__restrict__ mini_pos)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_m || m.flt[i]) return;
    const u32 s = m.seq[i], qp = m.q_pos[i], n = m.n[i];
    const int qlen = (int)(seq_off[s + 1] - seq_off[s]);
    const u64 *r = occ + m.off[i];
    const u64 fl = m.tandem[i] ? kSeedTandem : 0ULL;
    (void)q_span;
    const int sp = (int)m.span[i];
    const u64 yf = (u64)sp << 32 | (u64)(qp >> 1) | fl;
    const u64 yr = (u64)sp << 32 | (u64)(u32)(qlen - ((int)(qp >> 1) + 1 - sp) - 1) | fl;
    uint4 *o = a + a_pos[i];
    for (u32 t = 0; t < n; ++t) {
        const u64 rr = __ldg(r + t);
        const u64 rpos = (u64)((u32)rr >> 1);
        u64 x, y;
        if ((rr & 1) == (u64)(qp & 1)) { x = (rr & 0xffffffff00000000ULL) | rpos; y = yf; }
        else { x = 1ULL << 63 | (rr & 0xffffffff00000000ULL) | rpos; y = yr; }
        o[t] = make_uint4((u32)x, (u32)(x >> 32), (u32)y, (u32)(y >> 32));
    }
    if (mini_pos) mini_pos[mp_pos[i]] = (u64)sp << 32 | (u64)(qp >> 1);
}

// ---- radix_sort_128x (ksort.h:98-151) replayed per read --------------------------------------------------------------------
//
// One warp per read (reads binned by size so that the digit array of a read fits the shared memory of its class; several reads
// per SM).  A segment [beg, end) at byte `sh` lives in one of two anchor buffers (ping-pong):
//   1. digits: D[i] = byte `sh` of x, histogram (a pass in which all digits agree is the identity: next byte, nothing moves);
//   2. destinations: the American-flag permutation (ksort.h:125-138) only ever reads slots that still hold their ORIGINAL
//      element (a bucket's cursor never passes a slot twice), so it is replayed on the digits alone -- lane 0, two dependent
//      shared-memory loads per moved element -- and emits dest[src]; a pass with exactly two non-empty buckets (the strand
//      split at byte 7, always) has a closed form evaluated by all lanes (oracle/seed_model.py: pass_dest_walk / pass_dest_two);
//   3. scatter: other[dest[i]] = this[i], all lanes, 16-byte elements;
//   4. buckets of more than 64 elements are pushed for the next byte; the others (stable insertion sort in the reference,
//      ksort.h:105-115,143) are ranked on the full key inside 32-element windows and written to their final place.
// At byte 0 nothing follows (ksort.h:140).  The sorted read ends up in the second buffer, which the chaining kernels read.
struct SortShared {
    u32 cur[256];
    u32 end[256];
    u32 nonempty[64];      // 256 bytes: digits of the non-empty buckets of the segment, ascending
};

__device__ __forceinline__ u64 key_x(const uint4 *p)
{
    const uint2 v = *reinterpret_cast<const uint2 *>(p);
    return (u64)v.y << 32 | v.x;
}

// stable rank sort of src[beg, end) (at most 64 elements) on the full key into out[beg, end), by one warp (src may alias out)
__device__ __forceinline__ void sort64_out(const uint4 *src, uint4 *out, int beg, int end, int lane)
{
    const int m = end - beg;
    if (m <= 0) return;
    uint4 e0 = make_uint4(0, 0, 0, 0), e1 = e0;
    if (lane < m) e0 = src[beg + lane];
    if (lane + 32 < m) e1 = src[beg + lane + 32];
    const u64 k0 = (u64)e0.y << 32 | e0.x, k1 = (u64)e1.y << 32 | e1.x;
    int r0 = 0, r1 = 0;
    for (int t = 0; t < m; ++t) {
        const u64 kt = t < 32 ? __shfl_sync(0xffffffffu, k0, t) : __shfl_sync(0xffffffffu, k1, t - 32);
        r0 += (kt < k0 || (kt == k0 && t < lane)) ? 1 : 0;
        r1 += (kt < k1 || (kt == k1 && t < lane + 32)) ? 1 : 0;
    }
    __syncwarp();
    if (lane < m) out[beg + r0] = e0;
    if (lane + 32 < m) out[beg + r1] = e1;
    __syncwarp();
}

template <bool DSMEM>     // digits of the read in shared memory (explicit LDS in the walk) or, for reads above the largest class, in HBM
__global__ void __launch_bounds__(128)
k_seed_sort(uint4 *__restrict__ buf_a, uint4 *__restrict__ buf_b, const long long *__restrict__ a_off, const int *__restrict__ list, int n_list,
            int cap, unsigned char *__restrict__ g_dig, u32 *__restrict__ g_dest, u32 *__restrict__ g_lst, int4 *__restrict__ g_stack)
{
    extern __shared__ __align__(16) unsigned char sort_dyn[];      // per warp: cap digit bytes, then SortShared
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int li = blockIdx.x * wpb + wid;
    if (li >= n_list) return;
    const int r = list[li];
    const long long o0 = a_off[r];
    const int n = (int)(a_off[r + 1] - o0);
    if (n == 0) return;
    uint4 *const XA = buf_a + o0, *const OUT = buf_b + o0;
    if (n <= 64) { sort64_out(XA, OUT, 0, n, lane); return; }        // ksort.h:148
    unsigned char *mine = sort_dyn + (size_t)wid * ((size_t)cap + sizeof(SortShared));
    unsigned char *D;
    if (DSMEM) D = mine; else D = g_dig + o0;
    SortShared *SSw = reinterpret_cast<SortShared *>(mine + cap);
    u32 *dest = g_dest + o0, *lst = g_lst + o0;
    int4 *stack = g_stack + (o0 / 64 + 16LL * r);
    u32 *cur = SSw->cur, *en = SSw->end;
    int sp = 0;
    if (lane == 0) stack[0] = make_int4(0, n, 0, 56);
    sp = 1;
    __syncwarp();
    while (sp > 0) {
        --sp;
        const int4 seg = stack[sp];
        const int beg = seg.x, end = seg.y, buf = seg.z, m = end - beg;
        int sh = seg.w;
        const uint4 *src = buf ? OUT : XA;
        uint4 *dst = buf ? XA : OUT;
        u32 c[8];
        int n_nonempty = 0;
        bool done = sh < 0;
        u64 diff = 0;                                                  // bits in which the keys of the segment differ
        while (!done) {
            for (int t = lane; t < 256; t += 32) cur[t] = 0;
            __syncwarp();
            u64 k_or = 0, k_and = ~0ULL;
            for (int i0 = beg; i0 < end; i0 += 128) {
                u64 kx[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) { const int i = i0 + t * 32 + lane; kx[t] = i < end ? key_x(src + i) : 0ULL; }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int i = i0 + t * 32 + lane;
                    const bool act = i < end;
                    u32 d = 256;
                    if (act) {
                        d = (u32)(kx[t] >> sh) & 255u;
                        D[i] = (unsigned char)d;
                        dest[i] = (u32)i;
                        k_or |= kx[t]; k_and &= kx[t];
                    }
                    const u32 peers = __match_any_sync(0xffffffffu, d);
                    if (act && lane == __ffs(peers) - 1) cur[d] += __popc(peers);
                    __syncwarp();
                }
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) { k_or |= __shfl_xor_sync(0xffffffffu, k_or, o); k_and &= __shfl_xor_sync(0xffffffffu, k_and, o); }
            diff = k_or ^ k_and;
            bool single = false;
            n_nonempty = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) { c[t] = cur[lane * 8 + t]; single = single || c[t] == (u32)m; n_nonempty += c[t] ? 1 : 0; }
            single = __any_sync(0xffffffffu, single);
            if (!single) break;
            // identity pass (nothing moves): on to the highest lower byte in which the keys differ at all
            const u64 low = sh ? diff & ((1ULL << sh) - 1ULL) : 0ULL;
            if (!low) { done = true; break; }                          // all keys equal from here down
            sh = (63 - __clzll(low)) & ~7;
        }
        if (done) {
            if (buf != 1) for (int i = beg + lane; i < end; i += 32) OUT[i] = src[i];
            __syncwarp();
            continue;
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) n_nonempty += __shfl_xor_sync(0xffffffffu, n_nonempty, o);
        u32 sum = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) sum += c[t];
        u32 incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        u32 run = (u32)beg + incl - sum;
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 8; ++t) { cur[lane * 8 + t] = run; run += c[t]; en[lane * 8 + t] = run; }
        __syncwarp();
        if (n_nonempty == 2) {
            // closed form.  A = lower digit, region [beg, mid); B = [mid, end).  p_j: B-elements in A's region, q_j: A-elements in B's
            // region (equally many).  p_j -> q_{j-1} + 1 (mid for j = 0), q_j -> p_j, B-elements of B's region before q_last move up by one.
            u32 dA = 256;
#pragma unroll
            for (int t = 7; t >= 0; --t) if (c[t]) dA = (u32)(lane * 8 + t);
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) dA = min(dA, __shfl_xor_sync(0xffffffffu, dA, o));
            const int mid = (int)en[dA];
            int cntq = 0;
            for (int i0 = mid; i0 < end; i0 += 32) {
                const int i = i0 + lane;
                const bool q = i < end && D[i] == (unsigned char)dA;
                const u32 bal = __ballot_sync(0xffffffffu, q);
                if (q) lst[cntq + __popc(bal & ((1u << lane) - 1u))] = (u32)i;
                cntq += __popc(bal);
            }
            __syncwarp();
            const int q_last = cntq ? (int)lst[cntq - 1] : -1;
            for (int i = mid + lane; i < end; i += 32)
                if (D[i] != (unsigned char)dA && i < q_last) dest[i] = (u32)i + 1u;
            int cntp = 0;
            for (int i0 = beg; i0 < mid; i0 += 32) {
                const int i = i0 + lane;
                const bool p = i < mid && D[i] != (unsigned char)dA;
                const u32 bal = __ballot_sync(0xffffffffu, p);
                if (p) {
                    const int j = cntp + __popc(bal & ((1u << lane) - 1u));
                    dest[i] = j ? lst[j - 1] + 1u : (u32)mid;
                    dest[lst[j]] = (u32)i;
                }
                cntp += __popc(bal);
            }
        } else {                                                       // ksort.h:125-138 on the digits
            // the non-empty buckets in ascending order (the walk visits only those: segments of a few hundred anchors are common
            // below repeat copies, and 256 empty-bucket tests per segment would cost more than their elements)
            int mine = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) mine += c[t] ? 1 : 0;
            int before = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, before, o); if (lane >= o) before += t; }
            before -= mine;
            unsigned char *nb = reinterpret_cast<unsigned char *>(SSw->nonempty);
#pragma unroll
            for (int t = 0; t < 8; ++t) if (c[t]) nb[before++] = (unsigned char)(lane * 8 + t);
            __syncwarp();
            if (lane == 0) {
                for (int q = 0; q < n_nonempty; ++q) {
                    const u32 kk = nb[q];
                    const u32 ke = en[kk];
                    u32 kb = cur[kk];
                    while (kb != ke) {
                        u32 d = D[kb];
                        if (d == kk) { ++kb; continue; }
                        u32 from = kb;
                        do {
                            const u32 pos = cur[d];
                            cur[d] = pos + 1;
                            dest[from] = pos;
                            from = pos;
                            d = D[pos];
                        } while (d != kk);
                        dest[from] = kb;
                        ++kb;
                    }
                }
            }
        }
        __syncwarp();
        // scatter into the other buffer
        for (int i0 = beg; i0 < end; i0 += 128) {
            uint4 e[4];
            u32 dd[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) { const int i = i0 + t * 32 + lane; if (i < end) { e[t] = src[i]; dd[t] = dest[i]; } }
#pragma unroll
            for (int t = 0; t < 4; ++t) { const int i = i0 + t * 32 + lane; if (i < end) dst[dd[t]] = e[t]; }
        }
        __syncwarp();
        const int nbuf = buf ^ 1;
        if (sh == 0) {                                                 // ksort.h:140: nothing follows the last byte
            if (nbuf != 1) for (int i = beg + lane; i < end; i += 32) OUT[i] = dst[i];
            __syncwarp();
            continue;
        }
        // buckets of more than 64 elements: next byte -- directly the highest lower byte in which the keys of this segment differ (the
        // passes in between would be identities for every subset); none left: the bucket is final as it stands (sh = -8: copied out)
        const u64 low_diff = diff & ((1ULL << sh) - 1ULL);
        const int child_sh = low_diff ? ((63 - __clzll(low_diff)) & ~7) : -8;
        for (int t = 0; t < 8; ++t) {
            const bool big = c[t] > 64u;
            u32 bal = __ballot_sync(0xffffffffu, big);
            while (bal) {
                const int l = __ffs(bal) - 1;
                bal &= bal - 1;
                if (lane == l) { const int kk = lane * 8 + t; stack[sp] = make_int4((int)(en[kk] - c[t]), (int)en[kk], nbuf, child_sh); }
                ++sp;
            }
        }
        // the rest: ranked on the full key (stable) inside windows of 32 consecutive elements, written to their final place
        uint4 e = make_uint4(0, 0, 0, 0);
        if (beg + lane < end) e = dst[beg + lane];
        for (int s = beg; s < end;) {
            const int last = min(s + 32, end), i = s + lane;
            int bs = -1, be = -1;
            if (i < last) {
                const u32 d = (u32)((((u64)e.y << 32 | e.x) >> sh) & 255u);
                be = (int)en[d];
                bs = d ? (int)en[d - 1] : beg;
            }
            const u64 key = (u64)e.y << 32 | e.x;
            const int l_bs = __shfl_sync(0xffffffffu, bs, last - 1 - s), l_be = __shfl_sync(0xffffffffu, be, last - 1 - s);
            const int cut = l_be > last ? l_bs : last;
            const int f_be = __shfl_sync(0xffffffffu, be, 0);
            const int next_s = cut == s ? f_be : cut;
            uint4 e_next = make_uint4(0, 0, 0, 0);
            if (next_s + lane < end) e_next = dst[next_s + lane];      // in flight while this window is ranked (disjoint from its writes)
            if (cut == s) {                                            // the bucket starting here does not fit the window
                if (f_be - s <= 64) sort64_out(dst, OUT, s, f_be, lane);
            } else {
                // rank inside the own bucket (stable): elements before me count if their key is <= mine, elements behind me if it is
                // smaller.  The members of a bucket sit in adjacent lanes, so offsets 1 .. (largest bucket of the window) - 1 cover
                // every pair; below byte 4 the keys of a bucket differ in their low 32 bits only.
                int rk = 0;
                int mb = (i < cut) ? be - bs : 0;
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) mb = max(mb, __shfl_xor_sync(0xffffffffu, mb, o));
                const int lo_lane = bs - s, hi_lane = be - s;          // my bucket = lanes [lo_lane, hi_lane)
                if (sh < 32) {
                    const u32 k32 = e.x;
                    for (int o = 1; o < mb; ++o) {
                        const u32 kd = __shfl_up_sync(0xffffffffu, k32, o), ku = __shfl_down_sync(0xffffffffu, k32, o);
                        rk += (lane - o >= lo_lane && kd <= k32) ? 1 : 0;
                        rk += (lane + o < hi_lane && ku < k32) ? 1 : 0;
                    }
                } else {
                    for (int o = 1; o < mb; ++o) {
                        const u64 kd = __shfl_up_sync(0xffffffffu, key, o), ku = __shfl_down_sync(0xffffffffu, key, o);
                        rk += (lane - o >= lo_lane && kd <= key) ? 1 : 0;
                        rk += (lane + o < hi_lane && ku < key) ? 1 : 0;
                    }
                }
                __syncwarp();
                if (i < cut) OUT[bs + rk] = e;
                __syncwarp();
            }
            s = next_s;
            e = e_next;
        }
        __syncwarp();
    }
}

} // namespace mm2gb_seed
