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
// Integer / byte work, HBM- and latency-bound: no tensor cores here.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace mm2gb_seed {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int kTile = 1024;          // positions per sketch tile = threads per CTA
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
struct SketchTile {
    unsigned char code[kHalo + kTile];
    u32 nmask[(kHalo + kTile) / 32];
    u64 ix[kMaxW + kTile];            // info.x of positions t0 - w .. t0 + kTile - 1
    unsigned char iz[kMaxW + kTile];  // strand bit
    u32 warp_sum[kTile / 32];
    int seq;
};

template <bool WRITE>
__global__ void __launch_bounds__(kTile)
k_sketch(const unsigned char *__restrict__ seqs, const long long *__restrict__ seq_off, const int *__restrict__ tile_first, int n_seq,
         int w, int k, int rid_is_seq, u32 *__restrict__ tile_cnt, const u64 *__restrict__ tile_base, u64 *__restrict__ mv_x,
         u64 *__restrict__ mv_y, u32 *__restrict__ mv_seq)
{
    __shared__ SketchTile S;
    const int tid = threadIdx.x, tile = blockIdx.x;
    if (tid == 0) {
        int lo = 0, hi = n_seq;   // last s with tile_first[s] <= tile
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tile_first[mid] <= tile) lo = mid; else hi = mid; }
        S.seq = lo;
    }
    __syncthreads();
    const int s = S.seq;
    const long long base = seq_off[s];
    const int len = (int)(seq_off[s + 1] - base);
    const int t0 = (tile - tile_first[s]) * kTile;
    // stage the codes of positions t0 - kHalo .. t0 + kTile - 1 (outside the sequence: ambiguous)
    for (int j = tid; j < kHalo + kTile; j += kTile) {
        const int pos = t0 - kHalo + j;
        const int c = (pos >= 0 && pos < len) ? nt4(__ldg(seqs + base + pos)) : 4;
        S.code[j] = (unsigned char)c;
        const u32 m = __ballot_sync(0xffffffffu, c == 4);
        if ((tid & 31) == 0) S.nmask[j >> 5] = m;
    }
    __syncthreads();
    const u64 mask = (1ULL << (2 * k)) - 1;
    const int shift1 = 2 * (k - 1);
    // info of positions t0 - w .. t0 + kTile - 1
    for (int j = tid; j < w + kTile; j += kTile) {
        const int sj = kHalo - w + j;          // staged index of the position
        const int pos = t0 - w + j;
        u64 x = kNone;
        unsigned char z = 0;
        if (pos >= k - 1 && pos < len) {
            u64 f = 0, r = 0;
            bool ok = true;
            for (int t = k - 1; t >= 0; --t) {
                const int c = S.code[sj - t];
                ok = ok && c < 4;
                f = (f << 2 | (u64)(c & 3)) & mask;
                r = (r >> 2) | (u64)(3 ^ (c & 3)) << shift1;
            }
            if (ok && f != r) {
                z = f < r ? 0 : 1;
                x = hash64(z ? r : f, mask) << 8 | (u64)k;
            }
        }
        S.ix[j] = x;
        S.iz[j] = z;
    }
    __syncthreads();
    // the rules of the loop for position i = t0 + tid
    const int i = t0 + tid;
    int cnt = 0;
    u64 wpos = 0;
    // two passes over the same rules: count, then (after the block scan) write
    const int jj = w + tid;   // index of position i in S.ix
    int l = 0;
    u64 cur = kNone, mprev_x = kNone, mx = kNone;
    int mprev_p = -1, mp = -1;
    bool in_range = i < len;
    if (in_range) {
        // run of unambiguous bases ending at i (capped at 96, more than any threshold below)
        const int sj = kHalo + tid, q = sj >> 5, r = sj & 31;
        const u32 w2 = S.nmask[q] & (r == 31 ? 0xffffffffu : ((2u << r) - 1u));
        if (w2) l = r - (31 - __clz(w2));
        else {
            const u32 w1 = q >= 1 ? S.nmask[q - 1] : 0xffffffffu;
            if (w1) l = r + 1 + __clz(w1);
            else {
                const u32 w0 = q >= 2 ? S.nmask[q - 2] : 0xffffffffu;
                l = w0 ? r + 33 + __clz(w0) : 97;
            }
        }
        cur = S.ix[jj];
        for (int d = w; d >= 1; --d) {           // oldest -> newest, `<=` keeps the rightmost minimum
            const u64 x = S.ix[jj - d];
            if (x <= mprev_x) mprev_x = x, mprev_p = i - d;
        }
    }
    const int T1 = w + k - 1;
    int mode = 0;                                  // 2: rule P2, 3: rule P3
    if (in_range) {
        if (cur <= mprev_x) mode = 2;
        else if (mprev_p == i - w) {
            mode = 3;
            for (int d = w - 1; d >= 0; --d) {
                const u64 x = S.ix[jj - d];
                if (x <= mx) mx = x, mp = i - d;
            }
        }
    }
    for (int pass = 0; pass < (WRITE ? 2 : 1); ++pass) {
        int c = 0;
        auto emit = [&](int p) {
            if (pass == 1) {
                const int q = w + (p - t0);
                mv_x[wpos + c] = S.ix[q];
                mv_y[wpos + c] = (rid_is_seq ? (u64)s << 32 : 0ULL) | (u64)(u32)p << 1 | (u64)S.iz[q];
                mv_seq[wpos + c] = (u32)s;
            }
            ++c;
        };
        if (in_range) {
            if (l == T1 && mprev_x != kNone)                                   // P1
                for (int d = w - 1; d >= 1; --d)
                    if (S.ix[jj - d] == mprev_x && i - d != mprev_p) emit(i - d);
            if (mode == 2) {                                                  // P2
                if (l >= T1 + 1 && mprev_x != kNone) emit(mprev_p);
            } else if (mode == 3) {                                           // P3
                if (l >= T1) emit(mprev_p);
                if (l >= T1 && mx != kNone)
                    for (int d = w - 1; d >= 0; --d)
                        if (S.ix[jj - d] == mx && i - d != mp) emit(i - d);
            }
            if (i == len - 1) {                                               // P4
                const u64 fx = mode == 2 ? cur : mode == 3 ? mx : mprev_x;
                const int fp = mode == 2 ? i : mode == 3 ? mp : mprev_p;
                if (fx != kNone) emit(fp);
            }
        }
        if (pass == 0) {
            cnt = c;
            // block exclusive scan of cnt
            u32 v = (u32)cnt;
            const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
            if (lane == 31) S.warp_sum[wid] = v;
            __syncthreads();
            if (wid == 0) {
                u32 t = S.warp_sum[lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
                S.warp_sum[lane] = t;
            }
            __syncthreads();
            const u32 excl = v - (u32)cnt + (wid ? S.warp_sum[wid - 1] : 0u);
            if (!WRITE) { if (tid == kTile - 1) tile_cnt[tile] = excl + (u32)cnt; }
            else wpos = tile_base[tile] + excl;
        }
    }
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

__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const u32 *__restrict__ in, long long n, u64 *__restrict__ part)
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

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const u32 *__restrict__ in, long long n, const u64 *__restrict__ part,
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
};

__global__ void k_compact_seeds(const u64 *__restrict__ mv_y, const u32 *__restrict__ mv_seq, const u32 *__restrict__ occ_n,
                                const u64 *__restrict__ occ_off, const unsigned char *__restrict__ tandem, const u64 *__restrict__ m_idx,
                                long long n_mv, Seeds m)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_mv || occ_n[i] == 0) return;
    const u64 q = m_idx[i];
    m.n[q] = occ_n[i]; m.q_pos[q] = (u32)mv_y[i]; m.off[q] = occ_off[i]; m.seq[q] = mv_seq[i]; m.tandem[q] = tandem[i];
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
            long long st = i, en = i + 1;
            int rank = 0;
            while (st > b && m.n[st - 1] > (u32)max_occ) { --st; if (m.n[st] <= ni) ++rank; }          // (n, j) < (ni, i) with j < i
            while (en < e && m.n[en] > (u32)max_occ) { if (m.n[en] < ni) ++rank; ++en; }               // j > i: strictly smaller n
            const int len = (int)(seq_off[s + 1] - seq_off[s]);
            const int ps = st > b ? (int)(m.q_pos[st - 1] >> 1) : 0;
            const int pe = en < e ? (int)(m.q_pos[en] >> 1) : len;
            int max_high_occ = (int)((double)(pe - ps) / dist + .499);
            if (max_high_occ > 128) max_high_occ = 128;
            flt = (max_high_occ > 0 && rank < max_high_occ) ? 0 : 1;
            if (ni > (u32)max_max_occ) flt = 1;
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
    const int en = (int)(m.q_pos[i] >> 1) + 1, st = en - q_span;
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
    const u64 yf = (u64)q_span << 32 | (u64)(qp >> 1) | fl;
    const u64 yr = (u64)q_span << 32 | (u64)(u32)(qlen - ((int)(qp >> 1) + 1 - q_span) - 1) | fl;
    uint4 *o = a + a_pos[i];
    for (u32 t = 0; t < n; ++t) {
        const u64 rr = __ldg(r + t);
        const u64 rpos = (u64)((u32)rr >> 1);
        u64 x, y;
        if ((rr & 1) == (u64)(qp & 1)) { x = (rr & 0xffffffff00000000ULL) | rpos; y = yf; }
        else { x = 1ULL << 63 | (rr & 0xffffffff00000000ULL) | rpos; y = yr; }
        o[t] = make_uint4((u32)x, (u32)(x >> 32), (u32)y, (u32)(y >> 32));
    }
    if (mini_pos) mini_pos[mp_pos[i]] = (u64)q_span << 32 | (u64)(qp >> 1);
}

// ---- radix_sort_128x (ksort.h:98-151) replayed per read --------------------------------------------------------------------
//
// One CTA per read.  W[i] = digit << 24 | index of the element now at position i (index into the read's unsorted anchors).
// A segment [beg, end) at byte `sh` is taken by one warp: histogram of the digits (a pass in which all digits agree is the
// identity and moves on to the next byte), bucket bounds, then lane 0 replays the American-flag permutation (ksort.h:125-138)
// on the words; buckets of more than 64 elements are queued for the next byte, smaller ones (stable insertion sort in the
// reference, ksort.h:105-115) are ranked by the warp on the full key.  At byte 0 nothing follows (ksort.h:140).
constexpr int kSortWarps = 8;
constexpr int kSortThreads = kSortWarps * 32;
constexpr u32 kIdxMask = 0xffffffu;

struct SortShared {
    u32 cur[kSortWarps][256];
    u32 end[kSortWarps][256];
    int q_n[2];
    int q_take;
};

__device__ __forceinline__ u64 key_of(const uint4 *__restrict__ in, u32 idx)
{
    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(in + idx));
    return (u64)v.y << 32 | v.x;
}

// stable rank sort of W[beg, end) (at most 64 elements) on the full key, by one warp
__device__ __forceinline__ void small_sort(u32 *W, const uint4 *__restrict__ in, int beg, int end, int lane)
{
    const int m = end - beg;
    if (m <= 1) return;
    const u32 i0 = lane < m ? (W[beg + lane] & kIdxMask) : 0u, i1 = lane + 32 < m ? (W[beg + lane + 32] & kIdxMask) : 0u;
    const u64 k0 = lane < m ? key_of(in, i0) : kNone, k1 = lane + 32 < m ? key_of(in, i1) : kNone;
    int r0 = 0, r1 = 0;
    for (int t = 0; t < m; ++t) {
        const u64 kt = t < 32 ? __shfl_sync(0xffffffffu, k0, t) : __shfl_sync(0xffffffffu, k1, t - 32);
        r0 += (kt < k0 || (kt == k0 && t < lane)) ? 1 : 0;
        r1 += (kt < k1 || (kt == k1 && t < lane + 32)) ? 1 : 0;
    }
    __syncwarp();
    if (lane < m) W[beg + r0] = i0;
    if (lane + 32 < m) W[beg + r1] = i1;
    __syncwarp();
}

__global__ void __launch_bounds__(kSortThreads)
k_seed_sort(const uint4 *__restrict__ a_in, uint4 *__restrict__ a_out, const long long *__restrict__ a_off, const int *__restrict__ order,
            int n_reads, int smem_words, u32 *__restrict__ g_words, int2 *__restrict__ g_queue)
{
    extern __shared__ u32 sort_dyn[];
    __shared__ SortShared S;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int r = order ? order[blockIdx.x] : blockIdx.x;
    if (r >= n_reads) return;
    const long long o0 = a_off[r];
    const int n = (int)(a_off[r + 1] - o0);
    if (n == 0) return;
    const uint4 *in = a_in + o0;
    uint4 *out = a_out + o0;
    u32 *W = n <= smem_words ? sort_dyn : g_words + o0;
    // segment queues of the read (two levels, ping-pong): entries (beg, end); at most n / 65 + 1 segments per level
    int2 *Q[2];
    const long long qcap = n / 64 + 2;
    Q[0] = g_queue + 2 * (o0 / 64 + 2LL * r);
    Q[1] = Q[0] + qcap;
    for (int i = tid; i < n; i += kSortThreads) W[i] = (u32)i;
    if (tid == 0) { S.q_n[0] = S.q_n[1] = 0; S.q_take = 0; }
    __syncthreads();
    if (n <= 64) {                                   // ksort.h:148
        if (wid == 0) small_sort(W, in, 0, n, lane);
    } else {
        if (tid == 0) { Q[0][0] = make_int2(0, n); S.q_n[0] = 1; }
        __syncthreads();
        int level = 0;
        for (int sh = 56; sh >= 0; sh -= 8, level ^= 1) {
            const int nq = S.q_n[level];
            if (nq == 0) break;
            for (;;) {
                int qi = 0;
                if (lane == 0) qi = atomicAdd(&S.q_take, 1);
                qi = __shfl_sync(0xffffffffu, qi, 0);
                if (qi >= nq) break;
                const int2 seg = Q[level][qi];
                const int beg = seg.x, end = seg.y, m = end - beg;
                u32 *cur = S.cur[wid], *en = S.end[wid];
                for (int t = lane; t < 256; t += 32) cur[t] = 0;
                __syncwarp();
                for (int i = beg + lane; i < end; i += 32) {
                    const u32 idx = W[i] & kIdxMask;
                    const u32 d = (u32)(key_of(in, idx) >> sh) & 255u;
                    W[i] = d << 24 | idx;
                    atomicAdd(&cur[d], 1u);
                }
                __syncwarp();
                // bucket bounds: lane t owns digits 8t .. 8t+7
                u32 c[8], sum = 0;
                bool single = false;
#pragma unroll
                for (int t = 0; t < 8; ++t) { c[t] = cur[lane * 8 + t]; sum += c[t]; single = single || c[t] == (u32)m; }
                single = __any_sync(0xffffffffu, single);
                if (single) {                        // identity pass: the whole segment moves on to the next byte
                    if (sh > 0 && lane == 0) { const int q = atomicAdd(&S.q_n[level ^ 1], 1); Q[level ^ 1][q] = seg; }
                    continue;
                }
                u32 incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                u32 run = (u32)beg + incl - sum;
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 8; ++t) { cur[lane * 8 + t] = run; run += c[t]; en[lane * 8 + t] = run; }
                __syncwarp();
                if (lane == 0) {                     // ksort.h:125-138
                    for (int kk = 0; kk < 256; ++kk) {
                        const u32 ke = en[kk];
                        u32 kb = cur[kk];
                        while (kb != ke) {
                            u32 carried = W[kb];
                            u32 d = carried >> 24;
                            if (d == (u32)kk) { ++kb; continue; }
                            do {
                                const u32 pos = cur[d];
                                cur[d] = pos + 1;
                                const u32 ev = W[pos];
                                W[pos] = carried;
                                carried = ev;
                                d = carried >> 24;
                            } while (d != (u32)kk);
                            W[kb++] = carried;
                        }
                        cur[kk] = kb;
                    }
                }
                __syncwarp();
                if (sh > 0) {                        // ksort.h:140-145
                    for (int kk = 0; kk < 256; ++kk) {
                        const int be = (int)en[kk], bb = kk ? (int)en[kk - 1] : beg;
                        const int sz = be - bb;
                        if (sz > 64) { if (lane == 0) { const int q = atomicAdd(&S.q_n[level ^ 1], 1); Q[level ^ 1][q] = make_int2(bb, be); } }
                        else if (sz > 1) small_sort(W, in, bb, be, lane);
                    }
                }
                __syncwarp();
            }
            __syncthreads();
            if (tid == 0) { S.q_n[level] = 0; S.q_take = 0; }
            __syncthreads();
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += kSortThreads) out[i] = __ldg(in + (W[i] & kIdxMask));
}

} // namespace mm2gb_seed
