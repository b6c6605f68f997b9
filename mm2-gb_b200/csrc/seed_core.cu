// seed_core.cu -- host side of the device seeding stage (include/mm2gb_seed.h), part of libmm2gb_chain.so.
//
// Replaces what the reference does per read on a host thread before its GPU path starts -- mm_map_seed (map.c:355-391): mm_sketch
// (sketch.c:77), mm_seed_mz_flt (seed.c:5), mm_collect_matches (seed.c:98) with mm_idx_get (index.c:81), anchor construction and
// radix_sort_128x (map.c:295-331) -- by a pipeline of kernels over the whole batch (seed_kernels.cuh) whose output, the x-sorted
// anchor arrays, is consumed by the chaining kernels where it lies in HBM.  The index (mm_idx_t: minimizer -> positions sorted
// ascending, index.c:213-266) is sketched on the device by the same kernel, grouped on the host once per index part, and kept
// as an open-addressing table in HBM.
#include "seed_kernels.cuh"
#include "../../include/mm2gb_seed.h"
#include "host_place.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

using namespace mm2gb_seed;

extern "C" void mm2gb_internal_set_error(const char *msg);

namespace {

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    mm2gb_internal_set_error(buf);
    return code;
}

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) return fail(MM2GB_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

// flags of mm_mapopt_t (minimap.h:11-40) that change what collect_seed_hits / skip_seed do
constexpr int64_t F_NO_DIAG = 0x001, F_NO_DUAL = 0x002, F_FOR_ONLY = 0x100000, F_REV_ONLY = 0x200000, F_HEAP_SORT = 0x400000,
                  F_QSTRAND = 0x100000000LL;

template <class T> int dalloc(T *&p, size_t n)
{
    p = nullptr;
    CK(cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T)));
    return MM2GB_OK;
}

} // namespace

// ---- index -----------------------------------------------------------------------------------------------------------------

struct mm2gb_index {
    int device = 0, w = 0, k = 0;
    bool hpc = false;          // MM_I_HPC: minimizers of the homopolymer-compressed sequences
    // host copy: keys ascending, occurrences of keys[i] at occ[off[i] .. off[i+1]) ascending
    std::vector<uint64_t> keys, off, occ;
    bool keys_sorted = true;               // false: built from an enumerated hash table (mm2gb_index_from_lists); lookups go through hk / hv
    std::vector<uint64_t> hk, hv;          // host copy of the device table
    // device table
    u64 *d_key = nullptr, *d_val = nullptr, *d_occ = nullptr;
    uint64_t mask = 0;
};

// Streams the size classes of the x-sort are spread over (run_sort), round robin.  kSortStreams is what a seeder owns, the default in
// use is four (MM2GB_SORT_STREAMS).  One stream per class is faster for the sort on its own (2.91 -> 2.6 ms on 3000 reads of 10-100 kb,
// profiles/r8f_sort_streams.txt) but slower for the fused step on batches that use many classes (10 k reads with repeat-rich
// outliers: 62.8 ms with four streams, 67.8 ms with fourteen, profiles/r8j-r8l_seed_full*.json): most likely the chain-extraction
// classes that follow on the context's own streams then share hardware queues with them (not isolated further).
constexpr int kSortStreams = 14;
constexpr int kSortStreamsDefault = 4;

struct mm2gb_seeder {
    const mm2gb_index *idx = nullptr;
    int device = 0;
    int64_t max_bases = 0, max_mv = 0, max_anchors = 0;
    int max_reads = 0, max_tiles = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[MM2GB_SEED_NTIMERS + 1] = {nullptr};
    cudaEvent_t ev_join = nullptr;
    cudaStream_t stage_stream[8] = {nullptr};
    cudaEvent_t stage_ev[8][4] = {{nullptr}}, stage_done[8] = {nullptr};
    int stage_pos[8] = {0};            // next window of a staging thread's ring (persists across jobs, see stage_slice)
    // staging workers (pageable sources): created on first use, parked on a condition variable between jobs
    std::vector<std::thread> pool;
    std::mutex pool_mu;
    std::condition_variable pool_cv, pool_done_cv;
    unsigned pool_gen = 0;
    int pool_pending = 0, pool_nt = 0;
    bool pool_stop = false;
    const char *job_seqs = nullptr;
    size_t job_b0 = 0, job_b1 = 0;
    cudaError_t job_err[8] = {cudaSuccess};
    // sequences
    unsigned char *d_seq = nullptr;
    long long *d_seq_off = nullptr;
    int *d_tile_first = nullptr;
    u32 *d_tile_cnt = nullptr;
    u64 *d_tile_base = nullptr, *d_part = nullptr, *d_scan_state = nullptr;
    // homopolymer compression (HPC indices only): element-end flags, element index of every base, element codes / end positions / offsets
    unsigned char *d_hflag = nullptr, *d_ecode = nullptr;
    u64 *d_eidx = nullptr;
    u32 *d_epos = nullptr;
    long long *d_eoff = nullptr;
    // minimizers
    u64 *d_mv_x = nullptr, *d_mv_y = nullptr, *d_mv_off = nullptr;
    u32 *d_mv_seq = nullptr;
    unsigned char *d_keep = nullptr, *d_tandem = nullptr;
    u64 *d_tab_key = nullptr;
    u32 *d_tab_cnt = nullptr;
    u32 *d_occ_n = nullptr, *d_has = nullptr;
    u64 *d_occ_off = nullptr, *d_m_idx = nullptr;
    // seeds
    Seeds m{};
    u32 *d_cnt_a = nullptr, *d_kept = nullptr;
    u64 *d_a_pos = nullptr, *d_mp_pos = nullptr, *d_mini_pos = nullptr;
    // per read
    long long *d_a_off = nullptr, *d_mp_off = nullptr;
    int *d_rep_len = nullptr;
    // anchors
    uint4 *d_a_tmp = nullptr, *d_a = nullptr;
    u32 *d_dest = nullptr, *d_lst = nullptr;       // x-sort: destinations of a pass, scratch lists
    unsigned char *d_dig = nullptr;                // digits of reads too long for shared memory
    int4 *d_stack = nullptr;
    int *d_sort_list = nullptr, *h_sort_list = nullptr;
    cudaStream_t sort_stream[kSortStreams] = {nullptr};
    cudaEvent_t sort_fork = nullptr, sort_join[kSortStreams] = {nullptr};
    int *d_f = nullptr, *d_p = nullptr;
    int sort_max_cap = 0;
    bool sketch_persistent = true;     // k_sketch32p (MM2GB_SKETCH_PERSISTENT=0: one CTA per tile, k_sketch32)
    int sketch_grid = 0;
    // pinned host
    long long *h_a_off = nullptr, *h_mp_off = nullptr;
    int *h_rep_len = nullptr;
    u64 *h_tot = nullptr;
    mm2gb_anchor_t *h_b = nullptr;
    uint64_t *h_u = nullptr;
    unsigned char *h_seq = nullptr;      // staging for pageable sequences
    std::vector<int> tile_first;
    std::vector<int64_t> a_off_copy;
    // last batch
    long long n_mv = 0, n_m = 0, n_a = 0, n_mp = 0, n_bases = 0;
    int n_tiles = 0;
    float ms[MM2GB_SEED_NTIMERS] = {0};
    bool timed = false;
};

namespace {

template <typename TI>
int scan_any(cudaStream_t st, const TI *in, long long n, u64 *out, u64 *part)
{
    if (n <= 0) { CK(cudaMemsetAsync(out, 0, sizeof(u64), st)); return MM2GB_OK; }
    const int nb = (int)((n + kScanChunk - 1) / kScanChunk);
    k_scan_reduce<TI><<<nb, kScanThreads, 0, st>>>(in, n, part);
    k_scan_top<<<1, 1024, 0, st>>>(part, nb);
    k_scan_apply<TI><<<nb, kScanThreads, 0, st>>>(in, n, part, out);
    CK(cudaGetLastError());
    return MM2GB_OK;
}
int scan_u32(cudaStream_t st, const u32 *in, long long n, u64 *out, u64 *part) { return scan_any<u32>(st, in, n, out, part); }

inline unsigned grid_for(long long n, int threads) { return (unsigned)std::max<long long>(1, (n + threads - 1) / threads); }

// tiles per sequence; -1 if the batch does not fit
int plan_tiles(mm2gb_seeder *sd, const int64_t *seq_off, int n_seq)
{
    sd->tile_first.resize((size_t)n_seq + 1);
    long long t = 0;
    for (int s = 0; s < n_seq; ++s) {
        sd->tile_first[(size_t)s] = (int)t;
        const int64_t len = seq_off[s + 1] - seq_off[s];
        if (len < 0 || len > INT32_MAX - 2 * kTile) return -1;
        t += (len + kTile - 1) / kTile;
        if (t > sd->max_tiles) return -1;
    }
    sd->tile_first[(size_t)n_seq] = (int)t;
    return (int)t;
}

// ---- sketch of a batch: prepare, one or several launches over consecutive tile ranges, finish -------------------------------
constexpr int kMaxChunks = 12;      // launches per batch; d_scan_state = 16 ticket counters, then one status word per tile

int sketch_prepare(mm2gb_seeder *sd, int n_seq)
{
    cudaStream_t st = sd->stream;
    const int nt = sd->n_tiles;
    sd->n_mv = 0;
    if (nt == 0) { CK(cudaMemsetAsync(sd->d_mv_off, 0, ((size_t)n_seq + 1) * sizeof(u64), st)); return MM2GB_OK; }
    CK(cudaMemsetAsync(sd->d_scan_state, 0, ((size_t)nt + 17) * sizeof(u64), st));
    k_tile_map<<<grid_for(n_seq, 256), 256, 0, st>>>(sd->d_tile_first, n_seq, (int *)sd->d_tile_cnt);
    CK(cudaGetLastError());
    return MM2GB_OK;
}

// tiles [t0, t1) of the batch on the main stream (launch number c of the batch)
int sketch_launch(mm2gb_seeder *sd, int n_seq, int rid_is_seq, int c, int t0, int t1)
{
    if (t1 <= t0) return MM2GB_OK;
    cudaStream_t st = sd->stream;
    const mm2gb_index *ix = sd->idx;
    const int nt = sd->n_tiles, n = t1 - t0;
    u64 *ticket = sd->d_scan_state + c, *status = sd->d_scan_state + 16;
    const int *tile_seq = (const int *)sd->d_tile_cnt;
    if (ix->hpc) {
        // elements of the homopolymer-compressed sequences first (flags, scan, write), then the generic kernel on them; the tiles are
        // planned on the original lengths (an upper bound of the element counts: surplus tiles publish a count of zero)
        const long long nb = sd->n_bases;
        k_hpc_flags<<<nt, 256, 0, st>>>(sd->d_seq, sd->d_seq_off, sd->d_tile_first, tile_seq, nt, sd->d_hflag);
        int rc = scan_any<unsigned char>(st, sd->d_hflag, nb, sd->d_eidx, sd->d_part);
        if (rc) return rc;
        k_hpc_write<<<nt, 256, 0, st>>>(sd->d_seq, sd->d_seq_off, sd->d_tile_first, tile_seq, nt, sd->d_hflag, sd->d_eidx, sd->d_ecode, sd->d_epos);
        k_hpc_offsets<<<grid_for(n_seq + 1, 256), 256, 0, st>>>(sd->d_seq_off, sd->d_eidx, n_seq, sd->d_eoff);
        k_sketch<u64, true><<<n, kTile, 0, st>>>(sd->d_ecode, sd->d_eoff, sd->d_tile_first, tile_seq, n_seq, nt, ix->w, ix->k, rid_is_seq, ticket, t0, t1, status,
                                                 (long long)sd->max_mv, sd->d_mv_x, sd->d_mv_y, sd->d_mv_seq, sd->d_tile_base, sd->d_epos);
    } else if (ix->k <= 15 && sd->sketch_persistent)
        k_sketch32p<<<std::min(n, sd->sketch_grid), kTile, 0, st>>>(sd->d_seq, sd->d_seq_off, sd->d_tile_first, tile_seq, n_seq, nt, ix->w, ix->k, rid_is_seq, ticket,
                                                                   t0, t1, status, (long long)sd->max_mv, sd->d_mv_x, sd->d_mv_y, sd->d_mv_seq, sd->d_tile_base);
    else if (ix->k <= 15)
        k_sketch32<<<n, kTile, 0, st>>>(sd->d_seq, sd->d_seq_off, sd->d_tile_first, tile_seq, n_seq, nt, ix->w, ix->k, rid_is_seq, ticket, t0, t1, status,
                                        (long long)sd->max_mv, sd->d_mv_x, sd->d_mv_y, sd->d_mv_seq, sd->d_tile_base);
    else
        k_sketch<u64, false><<<n, kTile, 0, st>>>(sd->d_seq, sd->d_seq_off, sd->d_tile_first, tile_seq, n_seq, nt, ix->w, ix->k, rid_is_seq, ticket, t0, t1, status,
                                           (long long)sd->max_mv, sd->d_mv_x, sd->d_mv_y, sd->d_mv_seq, sd->d_tile_base, nullptr);
    CK(cudaGetLastError());
    return MM2GB_OK;
}

// per-sequence offsets, the minimizer count on the host (one synchronisation: capacity check, grids of the later kernels)
int sketch_finish(mm2gb_seeder *sd, int n_seq)
{
    cudaStream_t st = sd->stream;
    const int nt = sd->n_tiles;
    if (nt == 0) return MM2GB_OK;
    CK(cudaMemcpyAsync(sd->h_tot, sd->d_tile_base + nt, sizeof(u64), cudaMemcpyDeviceToHost, st));
    k_seq_mv_off<<<grid_for(n_seq + 1, 256), 256, 0, st>>>(sd->d_tile_first, sd->d_tile_base, n_seq, sd->d_mv_off);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    sd->n_mv = (long long)sd->h_tot[0];
    if (sd->n_mv > sd->max_mv)
        return fail(MM2GB_ECAP, "batch has %lld minimizers, the seeder holds %lld (max_bases too small for this sequence content)",
                    sd->n_mv, (long long)sd->max_mv);
    return MM2GB_OK;
}

// sequences already on the device -> minimizers in d_mv_*
int run_sketch(mm2gb_seeder *sd, int n_seq, int rid_is_seq)
{
    int rc = sketch_prepare(sd, n_seq);
    if (rc) return rc;
    rc = sketch_launch(sd, n_seq, rid_is_seq, 0, 0, sd->n_tiles);
    if (rc) return rc;
    return sketch_finish(sd, n_seq);
}

int upload_offsets(mm2gb_seeder *sd, const int64_t *seq_off, int n_seq)
{
    if (n_seq < 0 || n_seq > sd->max_reads) return fail(MM2GB_ECAP, "batch of %d sequences exceeds capacity %d", n_seq, sd->max_reads);
    if (seq_off[0] != 0) return fail(MM2GB_EARG, "seq_off[0] must be 0");
    if (seq_off[n_seq] > sd->max_bases) return fail(MM2GB_ECAP, "batch of %lld bases exceeds capacity %lld", (long long)seq_off[n_seq], (long long)sd->max_bases);
    const int nt = plan_tiles(sd, seq_off, n_seq);
    if (nt < 0) return fail(MM2GB_ECAP, "batch does not fit the seeder (tiles)");
    sd->n_tiles = nt;
    sd->n_bases = seq_off[n_seq];
    CK(cudaMemcpyAsync(sd->d_seq_off, seq_off, ((size_t)n_seq + 1) * sizeof(long long), cudaMemcpyHostToDevice, sd->stream));
    CK(cudaMemcpyAsync(sd->d_tile_first, sd->tile_first.data(), ((size_t)n_seq + 1) * sizeof(int), cudaMemcpyHostToDevice, sd->stream));
    return MM2GB_OK;
}

// Sequences -> device.  Pinned sources are DMA'd as they are.  Pageable ones are staged through pinned windows by kStageThreads
// host threads, each with its own stream and two windows (the copy into one overlaps the DMA of the other): a single thread's
// memcpy (~10 GB/s) would be five times slower than the link.
// A thread's windows form a ring that carries over from one job (chunk) to the next: a chunk's slice is smaller than a window or
// two, and a thread that started every job at window 0 would first wait for the DMA of its previous slice -- copy and DMA in
// lockstep, ~19 GB/s from pageable memory (profiles/r8d_sketch_ab.txt: the sketch stage was bound by this, not by the kernel).
constexpr int kStageThreads = 8;
constexpr int kStageRing = 4;
constexpr size_t kStageWindow = (size_t)1 << 20;

static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return pinned;
}

// one worker's share of the job: its slice of [job_b0, job_b1) through its two pinned windows and its own stream
static void stage_slice(mm2gb_seeder *sd, int t, int nt)
{
    cudaError_t &err = sd->job_err[t];
    err = cudaSetDevice(sd->device);
    const size_t n_bases = sd->job_b1 - sd->job_b0;
    const size_t lo = sd->job_b0 + n_bases * (size_t)t / (size_t)nt, hi = sd->job_b0 + n_bases * (size_t)(t + 1) / (size_t)nt;
    int which = sd->stage_pos[t];
    for (size_t done = lo; done < hi && err == cudaSuccess; which = (which + 1) % kStageRing) {
        const size_t n = std::min(kStageWindow, hi - done);
        unsigned char *stg = sd->h_seq + ((size_t)t * kStageRing + (size_t)which) * kStageWindow;
        cudaEvent_t ev = sd->stage_ev[t][which];
        if ((err = cudaEventSynchronize(ev)) != cudaSuccess) break;     // the window's previous DMA (a never-recorded event is complete)
        memcpy(stg, sd->job_seqs + done, n);
        if ((err = cudaMemcpyAsync(sd->d_seq + done, stg, n, cudaMemcpyHostToDevice, sd->stage_stream[t])) != cudaSuccess) break;
        err = cudaEventRecord(ev, sd->stage_stream[t]);
        done += n;
    }
    sd->stage_pos[t] = which;
    if (err == cudaSuccess) err = cudaEventRecord(sd->stage_done[t], sd->stage_stream[t]);
}

static void stage_worker(mm2gb_seeder *sd, int t)
{
    unsigned seen = 0;
    for (;;) {
        int nt;
        {
            std::unique_lock<std::mutex> lk(sd->pool_mu);
            sd->pool_cv.wait(lk, [&] { return sd->pool_stop || sd->pool_gen != seen; });
            if (sd->pool_stop) return;
            seen = sd->pool_gen;
            nt = sd->pool_nt;
        }
        if (t < nt) stage_slice(sd, t, nt);
        {
            std::lock_guard<std::mutex> lk(sd->pool_mu);
            if (--sd->pool_pending == 0) sd->pool_done_cv.notify_one();
        }
    }
}

// bases [b0, b1) of a pageable source -> d_seq, staged by up to kStageThreads host threads (the caller + parked workers); afterwards
// the main stream waits for their copies
int stage_range(mm2gb_seeder *sd, const char *seqs, size_t b0, size_t b1)
{
    if (b1 <= b0) return MM2GB_OK;
    const size_t n_bases = b1 - b0;
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    int want = hw / 2;      // MM2GB_STAGE_THREADS: host threads per seeder for staging (several ranks / driver threads share the cores)
    if (const char *e = getenv("MM2GB_STAGE_THREADS")) want = atoi(e);
    const int nt = n_bases < 2 * kStageWindow ? 1 : std::max(1, std::min(kStageThreads, want));
    if (nt > 1 && sd->pool.empty())
        for (int t = 1; t < kStageThreads; ++t) sd->pool.emplace_back(stage_worker, sd, t);
    sd->job_seqs = seqs; sd->job_b0 = b0; sd->job_b1 = b1;
    if (nt > 1) {
        { std::lock_guard<std::mutex> lk(sd->pool_mu); sd->pool_nt = nt; sd->pool_pending = (int)sd->pool.size(); ++sd->pool_gen; }
        sd->pool_cv.notify_all();
    }
    stage_slice(sd, 0, nt);
    if (nt > 1) {
        std::unique_lock<std::mutex> lk(sd->pool_mu);
        sd->pool_done_cv.wait(lk, [&] { return sd->pool_pending == 0; });
    }
    for (int t = 0; t < nt; ++t) {
        if (sd->job_err[t] != cudaSuccess) return fail(MM2GB_ECUDA, "staging the sequences: %s", cudaGetErrorString(sd->job_err[t]));
        CK(cudaStreamWaitEvent(sd->stream, sd->stage_done[t], 0));
    }
    return MM2GB_OK;
}

// the staging streams start behind whatever still reads d_seq on the main stream
int stage_begin(mm2gb_seeder *sd)
{
    CK(cudaEventRecord(sd->ev_join, sd->stream));
    for (int t = 0; t < kStageThreads; ++t) CK(cudaStreamWaitEvent(sd->stage_stream[t], sd->ev_join, 0));
    return MM2GB_OK;
}

int upload_seqs(mm2gb_seeder *sd, const char *seqs, int64_t n_bases)
{
    if (n_bases <= 0) return MM2GB_OK;
    if (is_pinned_host(seqs)) { CK(cudaMemcpyAsync(sd->d_seq, seqs, (size_t)n_bases, cudaMemcpyHostToDevice, sd->stream)); return MM2GB_OK; }
    int rc = stage_begin(sd);
    if (rc) return rc;
    return stage_range(sd, seqs, 0, (size_t)n_bases);
}

// Host sequences -> minimizers: the batch is cut into a few chunks of whole reads; chunk c is sketched while chunk c + 1 is still
// being staged / crossing PCIe (the chained scan of the sketch continues across the launches).
int upload_and_sketch(mm2gb_seeder *sd, const char *seqs, const int64_t *seq_off, int n_seq, int rid_is_seq)
{
    int rc = sketch_prepare(sd, n_seq);
    if (rc) return rc;
    const int64_t total = seq_off[n_seq];
    if (total > 0 && sd->idx->hpc) {      // the homopolymer compression runs over the whole batch: upload first, one launch
        rc = upload_seqs(sd, seqs, total);
        if (rc) return rc;
        rc = sketch_launch(sd, n_seq, rid_is_seq, 0, 0, sd->n_tiles);
        if (rc) return rc;
    } else if (total > 0) {
        const bool pinned = is_pinned_host(seqs);
        rc = stage_begin(sd);
        if (rc) return rc;
        const int64_t target = std::max<int64_t>((int64_t)8 << 20, (total + kMaxChunks - 1) / kMaxChunks);
        int r0 = 0, c = 0;
        while (r0 < n_seq) {
            int r1 = r0;
            while (r1 < n_seq && (r1 == r0 || seq_off[r1 + 1] - seq_off[r0] <= target || c == kMaxChunks - 1)) ++r1;
            const size_t b0 = (size_t)seq_off[r0], b1 = (size_t)seq_off[r1];
            if (pinned) {
                if (b1 > b0) CK(cudaMemcpyAsync(sd->d_seq + b0, seqs + b0, b1 - b0, cudaMemcpyHostToDevice, sd->stage_stream[0]));
                CK(cudaEventRecord(sd->stage_done[0], sd->stage_stream[0]));
                CK(cudaStreamWaitEvent(sd->stream, sd->stage_done[0], 0));
            } else {
                rc = stage_range(sd, seqs, b0, b1);
                if (rc) return rc;
            }
            rc = sketch_launch(sd, n_seq, rid_is_seq, c, sd->tile_first[(size_t)r0], sd->tile_first[(size_t)r1]);
            if (rc) return rc;
            r0 = r1;
            ++c;
        }
    }
    return sketch_finish(sd, n_seq);
}

int check_params(const mm2gb_seed_params_t *p)
{
    if (!p) return fail(MM2GB_EARG, "null parameters");
    if (p->flag & (F_NO_DIAG | F_NO_DUAL | F_FOR_ONLY | F_REV_ONLY | F_HEAP_SORT | F_QSTRAND))
        return fail(MM2GB_EARG, "map flag 0x%llx changes seed collection (no-diag / no-dual / for-only / rev-only / heap-sort / qstrand): "
                                "not supported by the device seeding path", (long long)p->flag);
    if (p->sdust_thres > 0) return fail(MM2GB_EARG, "sdust masking of query minimizers is not supported by the device seeding path");
    if (p->mid_occ <= 0) return fail(MM2GB_EARG, "mid_occ must be positive (run mm_mapopt_update / mm2gb_index_cal_max_occ first)");
    return MM2GB_OK;
}

// x-sort of every read (k_seed_sort): reads binned by anchor count into shared-memory size classes (digit bytes per read,
// warps = reads per CTA), the classes side by side on auxiliary streams, longest reads first inside a class
int run_sort(mm2gb_seeder *sd, int n_reads)
{
    struct Cls { int cap, warps; };
    // small reads: several warps (= reads) per CTA; above that one warp per CTA with caps chosen so that exactly 16, 12, 10, 8, 6, 5, 4, 3, 2
    // CTAs fit the shared memory of an SM (227 KB; digits + SortShared + the 1 KB the runtime reserves per CTA) -- what matters for
    // these latency-bound single warps is how many reads are resident
    static const Cls base[] = {{2048, 4}, {4096, 4}, {8192, 2}, {11200, 1}, {16032, 1}, {19904, 1}, {25728, 1}, {35408, 1}, {43152, 1},
                               {54784, 1}, {74144, 1}, {112896, 1}};
    std::vector<Cls> cls;
    for (const Cls &c : base) if (c.cap < sd->sort_max_cap) cls.push_back(c);
    cls.push_back({sd->sort_max_cap, 1});
    const int nc = (int)cls.size();            // class nc = reads whose digits stay in HBM
    std::vector<std::vector<int>> bin((size_t)nc + 1);
    const long long *off = sd->h_a_off;
    for (int r = 0; r < n_reads; ++r) {
        const long long n = off[r + 1] - off[r];
        if (n <= 0) continue;
        int k = 0;
        while (k < nc && n > cls[(size_t)k].cap) ++k;
        bin[(size_t)k].push_back(r);
    }
    int pos = 0;
    std::vector<int> start((size_t)nc + 2, 0);
    for (int k = nc; k >= 0; --k) {
        auto &b = bin[(size_t)k];
        std::sort(b.begin(), b.end(), [&](int x, int y) { return off[x + 1] - off[x] > off[y + 1] - off[y]; });
        start[(size_t)k] = pos;
        for (int r : b) sd->h_sort_list[pos++] = r;
    }
    if (!pos) return MM2GB_OK;
    cudaStream_t st = sd->stream;
    CK(cudaMemcpyAsync(sd->d_sort_list, sd->h_sort_list, (size_t)pos * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(sd->sort_fork, st));
    static const int n_streams = [] {       // MM2GB_SORT_STREAMS: see kSortStreams
        const char *e = getenv("MM2GB_SORT_STREAMS");
        return std::max(1, std::min(kSortStreams, e ? atoi(e) : kSortStreamsDefault));
    }();
    // experiment switch: the HBM class after all others, alone on the GPU (145-148 ms on the workload above: steadier, slower)
    static const bool hbm_alone = getenv("MM2GB_SORT_HBM_ALONE") && atoi(getenv("MM2GB_SORT_HBM_ALONE")) != 0;
    int used = 0;
    for (int k = nc; k >= 0; --k) {
        const int cnt = (int)bin[(size_t)k].size();
        if (!cnt) continue;
        if (hbm_alone && k == nc) continue;
        const int cap = k < nc ? cls[(size_t)k].cap : 0, warps = k < nc ? cls[(size_t)k].warps : 1;
        cudaStream_t ss = sd->sort_stream[used % n_streams];
        CK(cudaStreamWaitEvent(ss, sd->sort_fork, 0));
        const size_t smem = (size_t)warps * ((size_t)cap + sizeof(SortShared));
        if (k < nc)
            k_seed_sort<true><<<(cnt + warps - 1) / warps, warps * 32, smem, ss>>>(sd->d_a_tmp, sd->d_a, sd->d_a_off, sd->d_sort_list + start[(size_t)k], cnt, cap,
                                                                                    sd->d_dig, sd->d_dest, sd->d_lst, sd->d_stack);
        else
            k_seed_sort<false><<<cnt, 32, sizeof(SortShared), ss>>>(sd->d_a_tmp, sd->d_a, sd->d_a_off, sd->d_sort_list + start[(size_t)k], cnt, 0,
                                                                     sd->d_dig, sd->d_dest, sd->d_lst, sd->d_stack);
        CK(cudaGetLastError());
        ++used;
    }
    for (int i = 0; i < std::min(used, n_streams); ++i) {
        CK(cudaEventRecord(sd->sort_join[i], sd->sort_stream[i]));
        CK(cudaStreamWaitEvent(st, sd->sort_join[i], 0));
    }
    if (hbm_alone && !bin[(size_t)nc].empty()) {      // after every other class has left the GPU
        k_seed_sort<false><<<(int)bin[(size_t)nc].size(), 32, sizeof(SortShared), st>>>(sd->d_a_tmp, sd->d_a, sd->d_a_off, sd->d_sort_list + start[(size_t)nc],
                                                                                       (int)bin[(size_t)nc].size(), 0, sd->d_dig, sd->d_dest, sd->d_lst, sd->d_stack);
        CK(cudaGetLastError());
    }
    return MM2GB_OK;
}

// the stages behind the sketch, up to the x-sorted anchors in sd->d_a and the per-read offsets on the host (sd->h_a_off)
int run_seed(mm2gb_seeder *sd, const mm2gb_seed_params_t *prm, const int64_t *seq_off, int n_reads, bool want_mini_pos, const char *host_seqs)
{
    cudaStream_t st = sd->stream;
    const mm2gb_index *ix = sd->idx;
    const int T = 256;
    sd->n_m = sd->n_a = sd->n_mp = 0;
    CK(cudaEventRecord(sd->ev[0], st));
    int rc = host_seqs ? upload_and_sketch(sd, host_seqs, seq_off, n_reads, 0) : run_sketch(sd, n_reads, 0);
    if (rc) return rc;
    const long long n_mv = sd->n_mv;
    CK(cudaEventRecord(sd->ev[1], st));
    CK(cudaMemsetAsync(sd->d_rep_len, 0, (size_t)std::max(n_reads, 1) * sizeof(int), st));
    // opt->max_qlen (map.c:376): such reads get no anchors -- drop their minimizers through the keep flags
    const bool qflt = prm->q_occ_frac > 0.0f && prm->mid_occ > 0;
    if (n_mv) {
        if (qflt) {
            CK(cudaMemsetAsync(sd->d_tab_key, 0xff, (size_t)2 * n_mv * sizeof(u64), st));
            CK(cudaMemsetAsync(sd->d_tab_cnt, 0, (size_t)2 * n_mv * sizeof(u32), st));
            k_qocc_count<<<grid_for(n_mv, T), T, 0, st>>>(sd->d_mv_x, sd->d_mv_seq, sd->d_mv_off, n_mv, prm->mid_occ, sd->d_tab_key, sd->d_tab_cnt);
            k_qocc_flag<<<grid_for(n_mv, T), T, 0, st>>>(sd->d_mv_x, sd->d_mv_seq, sd->d_mv_off, n_mv, prm->mid_occ, prm->q_occ_frac, sd->d_tab_key,
                                                         sd->d_tab_cnt, sd->d_keep);
        } else CK(cudaMemsetAsync(sd->d_keep, 1, (size_t)n_mv, st));
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(sd->ev[2], st));
    if (n_mv) {
        DevIndex di{ix->d_key, ix->d_val, ix->d_occ, ix->mask};
        k_lookup<<<grid_for(n_mv, T), T, 0, st>>>(di, sd->d_mv_x, sd->d_mv_seq, sd->d_keep, n_mv, sd->d_occ_n, sd->d_occ_off, sd->d_tandem, sd->d_has);
        CK(cudaGetLastError());
    }
    rc = scan_u32(st, sd->d_has, n_mv, sd->d_m_idx, sd->d_part);
    if (rc) return rc;
    CK(cudaMemcpyAsync(sd->h_tot, sd->d_m_idx + n_mv, sizeof(u64), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const long long n_m = sd->n_m = (long long)sd->h_tot[0];
    if (n_m) {
        k_compact_seeds<<<grid_for(n_mv, T), T, 0, st>>>(sd->d_mv_x, sd->d_mv_y, sd->d_mv_seq, sd->d_occ_n, sd->d_occ_off, sd->d_tandem, sd->d_m_idx, n_mv, sd->m);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(sd->ev[3], st));
    if (n_m) {
        k_select<<<grid_for(n_m, T), T, 0, st>>>(sd->m, sd->d_m_idx, sd->d_mv_off, sd->d_seq_off, n_m, prm->mid_occ, prm->max_max_occ, prm->occ_dist,
                                                 ix->k, sd->d_cnt_a, sd->d_kept);
        k_rep_len<<<grid_for(n_m, T), T, 0, st>>>(sd->m, sd->d_m_idx, sd->d_mv_off, n_m, ix->k, sd->d_rep_len);
        CK(cudaGetLastError());
    }
    rc = scan_u32(st, sd->d_cnt_a, n_m, sd->d_a_pos, sd->d_part);
    if (rc) return rc;
    rc = scan_u32(st, sd->d_kept, n_m, sd->d_mp_pos, sd->d_part);
    if (rc) return rc;
    k_read_offsets<<<grid_for(n_reads + 1, T), T, 0, st>>>(sd->d_m_idx, sd->d_mv_off, sd->d_a_pos, sd->d_mp_pos, n_reads, sd->d_a_off, sd->d_mp_off);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(sd->h_a_off, sd->d_a_off, ((size_t)n_reads + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(sd->h_mp_off, sd->d_mp_off, ((size_t)n_reads + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(sd->h_rep_len, sd->d_rep_len, (size_t)std::max(n_reads, 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const long long n_a = sd->n_a = sd->h_a_off[n_reads];
    sd->n_mp = sd->h_mp_off[n_reads];
    if (n_a > sd->max_anchors) return fail(MM2GB_ECAP, "batch seeds %lld anchors, the seeder holds %lld", n_a, (long long)sd->max_anchors);
    CK(cudaEventRecord(sd->ev[4], st));
    if (n_m) {
        k_expand<<<grid_for(n_m, T), T, 0, st>>>(sd->m, ix->d_occ, sd->d_a_pos, sd->d_mp_pos, sd->d_seq_off, n_m, ix->k, sd->d_a_tmp,
                                                 want_mini_pos ? sd->d_mini_pos : nullptr);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(sd->ev[5], st));
    if (n_a) {
        rc = run_sort(sd, n_reads);
        if (rc) return rc;
    }
    CK(cudaEventRecord(sd->ev[6], st));
    sd->timed = true;
    (void)seq_off;
    return MM2GB_OK;
}

} // namespace

// ---- C ABI: seeder ---------------------------------------------------------------------------------------------------------

extern "C" int mm2gb_seeder_create(mm2gb_seeder_t **out, const mm2gb_index_t *idx, int64_t max_bases, int max_reads, int64_t max_anchors)
{
    if (!out || !idx || max_bases <= 0 || max_reads <= 0 || max_anchors <= 0) return fail(MM2GB_EARG, "bad argument");
    if (max_anchors > (int64_t)INT32_MAX - 1024) return fail(MM2GB_EARG, "max_anchors must be below 2^31");
    *out = nullptr;
    CK(cudaSetDevice(idx->device));
    mm2gb_seeder *sd = new mm2gb_seeder();
    sd->idx = idx;
    sd->device = idx->device;
    sd->max_bases = max_bases;
    sd->max_reads = max_reads;
    sd->max_anchors = max_anchors;
    sd->max_tiles = (int)std::min<int64_t>(INT32_MAX - 1, max_bases / kTile + max_reads + 1);
    // expected density of (w, k)-minimizers is 2 / (w + 1) per base; repeats of short period exceed it, hence the margin
    sd->max_mv = std::max<int64_t>(1024, (int64_t)((double)max_bases * 3.0 / (idx->w + 1)) + 64 * (int64_t)max_reads);
    const size_t M = (size_t)sd->max_mv, A = (size_t)max_anchors, R = (size_t)max_reads + 2, NT = (size_t)sd->max_tiles + 2;
    int rc = MM2GB_OK;
    mm2gb::NearGpu near_gpu(sd->device);   // the pinned staging windows and counters below land on the GPU's NUMA node
#define TRY(x) do { if ((rc = (x)) != MM2GB_OK) { mm2gb_seeder_destroy(sd); return rc; } } while (0)
#define TRYC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail(MM2GB_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); mm2gb_seeder_destroy(sd); return rc; } } while (0)
    TRYC(cudaStreamCreateWithFlags(&sd->stream, cudaStreamNonBlocking));
    for (auto &e : sd->ev) TRYC(cudaEventCreate(&e));
    TRYC(cudaEventCreateWithFlags(&sd->ev_join, cudaEventDisableTiming));
    for (int t = 0; t < 8; ++t) {
        TRYC(cudaStreamCreateWithFlags(&sd->stage_stream[t], cudaStreamNonBlocking));
        TRYC(cudaEventCreateWithFlags(&sd->stage_done[t], cudaEventDisableTiming));
        for (auto &e : sd->stage_ev[t]) TRYC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    TRY(dalloc(sd->d_seq, (size_t)max_bases + 16));
    TRY(dalloc(sd->d_seq_off, R)); TRY(dalloc(sd->d_tile_first, R));
    TRY(dalloc(sd->d_tile_cnt, NT)); TRY(dalloc(sd->d_tile_base, NT)); TRY(dalloc(sd->d_scan_state, NT + 18));
    TRY(dalloc(sd->d_part, std::max(std::max(M, NT), (size_t)max_bases) / kScanChunk + 4));
    if (idx->hpc) {
        TRY(dalloc(sd->d_hflag, (size_t)max_bases + 16)); TRY(dalloc(sd->d_ecode, (size_t)max_bases + 16)); TRY(dalloc(sd->d_eidx, (size_t)max_bases + 2));
        TRY(dalloc(sd->d_epos, (size_t)max_bases + 16)); TRY(dalloc(sd->d_eoff, R));
    }
    TRY(dalloc(sd->d_mv_x, M)); TRY(dalloc(sd->d_mv_y, M)); TRY(dalloc(sd->d_mv_seq, M)); TRY(dalloc(sd->d_mv_off, R));
    TRY(dalloc(sd->d_keep, M)); TRY(dalloc(sd->d_tandem, M));
    TRY(dalloc(sd->d_tab_key, 2 * M)); TRY(dalloc(sd->d_tab_cnt, 2 * M));
    TRY(dalloc(sd->d_occ_n, M)); TRY(dalloc(sd->d_has, M)); TRY(dalloc(sd->d_occ_off, M)); TRY(dalloc(sd->d_m_idx, M + 1));
    TRY(dalloc(sd->m.n, M)); TRY(dalloc(sd->m.q_pos, M)); TRY(dalloc(sd->m.off, M)); TRY(dalloc(sd->m.seq, M));
    TRY(dalloc(sd->m.tandem, M)); TRY(dalloc(sd->m.flt, M)); TRY(dalloc(sd->m.span, M));
    TRY(dalloc(sd->d_cnt_a, M)); TRY(dalloc(sd->d_kept, M)); TRY(dalloc(sd->d_a_pos, M + 1)); TRY(dalloc(sd->d_mp_pos, M + 1));
    TRY(dalloc(sd->d_mini_pos, M));
    TRY(dalloc(sd->d_a_off, R)); TRY(dalloc(sd->d_mp_off, R)); TRY(dalloc(sd->d_rep_len, R));
    TRY(dalloc(sd->d_a_tmp, A)); TRY(dalloc(sd->d_a, A)); TRY(dalloc(sd->d_dest, A)); TRY(dalloc(sd->d_lst, A)); TRY(dalloc(sd->d_dig, A));
    TRY(dalloc(sd->d_stack, A / 64 + 16 * R + 16)); TRY(dalloc(sd->d_sort_list, R));
    TRYC(cudaHostAlloc((void **)&sd->h_sort_list, R * sizeof(int), cudaHostAllocDefault));
    for (auto &x : sd->sort_stream) TRYC(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
    TRYC(cudaEventCreateWithFlags(&sd->sort_fork, cudaEventDisableTiming));
    for (auto &x : sd->sort_join) TRYC(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
    TRY(dalloc(sd->d_f, A)); TRY(dalloc(sd->d_p, A));
    TRYC(cudaHostAlloc((void **)&sd->h_a_off, R * sizeof(long long), cudaHostAllocDefault));
    TRYC(cudaHostAlloc((void **)&sd->h_mp_off, R * sizeof(long long), cudaHostAllocDefault));
    TRYC(cudaHostAlloc((void **)&sd->h_rep_len, R * sizeof(int), cudaHostAllocDefault));
    TRYC(cudaHostAlloc((void **)&sd->h_tot, 64, cudaHostAllocDefault));
    TRYC(cudaHostAlloc((void **)&sd->h_seq, (size_t)kStageThreads * kStageRing * kStageWindow, cudaHostAllocDefault));
    {
        // the x-sort keeps one digit byte per anchor of a read in shared memory; reads above the largest class use HBM for them
        int dev_smem = 0;
        TRYC(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, sd->device));
        sd->sort_max_cap = ((dev_smem - (int)sizeof(SortShared) - 1024) / 16) * 16;
        if (const char *e = getenv("MM2GB_SEED_SORT_CAP")) sd->sort_max_cap = std::max(64, std::min(sd->sort_max_cap, atoi(e) / 16 * 16));
        TRYC(cudaFuncSetAttribute(k_seed_sort<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sd->sort_max_cap + (int)sizeof(SortShared)));
        // The class whose digits stay in HBM (reads above sort_max_cap anchors) is one serial chain of dependent loads per read and
        // needs next to no shared memory: it asks for the largest L1.  Without the preference its CTAs ran with whatever carve-out
        // the SM had (usually the near-full shared memory of the other classes): 214-219 ms against 123 ms for the x-sort of 300
        // reads of 190 k anchors, with both figures (and 94 ms, an SM of its own) appearing from launch to launch
        // (profiles/r8g_sort_probe.txt, r8h_sort_probe.txt; MM2GB_SORT_L1=0 switches the preference off)
        if (!(getenv("MM2GB_SORT_L1") && atoi(getenv("MM2GB_SORT_L1")) == 0))
            if (cudaFuncSetAttribute(k_seed_sort<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1) != cudaSuccess)
                cudaGetLastError();      // a preference only
        int n_sm = 0, per_sm = 0;
        TRYC(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, sd->device));
        TRYC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sketch32p, kTile, 0));
        sd->sketch_grid = std::max(1, n_sm * std::max(1, per_sm));
        if (const char *e = getenv("MM2GB_SKETCH_PERSISTENT")) sd->sketch_persistent = atoi(e) != 0;
    }
#undef TRY
#undef TRYC
    *out = sd;
    return MM2GB_OK;
}

extern "C" void mm2gb_seeder_destroy(mm2gb_seeder_t *sd)
{
    if (!sd) return;
    cudaSetDevice(sd->device);
    if (sd->stream) cudaStreamSynchronize(sd->stream);
    void *dev[] = {sd->d_seq, sd->d_seq_off, sd->d_tile_first, sd->d_tile_cnt, sd->d_tile_base, sd->d_scan_state, sd->d_part, sd->d_mv_x, sd->d_mv_y, sd->d_mv_seq,
                   sd->d_mv_off, sd->d_keep, sd->d_tandem, sd->d_tab_key, sd->d_tab_cnt, sd->d_occ_n, sd->d_has, sd->d_occ_off, sd->d_m_idx,
                   sd->m.n, sd->m.q_pos, sd->m.off, sd->m.seq, sd->m.tandem, sd->m.flt, sd->m.span, sd->d_hflag, sd->d_ecode, sd->d_eidx, sd->d_epos, sd->d_eoff, sd->d_cnt_a, sd->d_kept, sd->d_a_pos, sd->d_mp_pos,
                   sd->d_mini_pos, sd->d_a_off, sd->d_mp_off, sd->d_rep_len, sd->d_a_tmp, sd->d_a, sd->d_dest, sd->d_lst, sd->d_dig, sd->d_stack, sd->d_sort_list, sd->d_f, sd->d_p};
    for (void *p : dev) if (p) cudaFree(p);
    void *pin[] = {sd->h_a_off, sd->h_mp_off, sd->h_rep_len, sd->h_tot, sd->h_b, sd->h_u, sd->h_seq, sd->h_sort_list};
    for (void *p : pin) if (p) cudaFreeHost(p);
    for (auto &e : sd->ev) if (e) cudaEventDestroy(e);
    if (!sd->pool.empty()) {
        { std::lock_guard<std::mutex> lk(sd->pool_mu); sd->pool_stop = true; }
        sd->pool_cv.notify_all();
        for (auto &t : sd->pool) t.join();
        sd->pool.clear();
    }
    for (int t = 0; t < 8; ++t) {
        if (sd->stage_stream[t]) { cudaStreamSynchronize(sd->stage_stream[t]); cudaStreamDestroy(sd->stage_stream[t]); }
        if (sd->stage_done[t]) cudaEventDestroy(sd->stage_done[t]);
        for (auto &e : sd->stage_ev[t]) if (e) cudaEventDestroy(e);
    }
    if (sd->ev_join) cudaEventDestroy(sd->ev_join);
    if (sd->sort_fork) cudaEventDestroy(sd->sort_fork);
    for (auto &e : sd->sort_join) if (e) cudaEventDestroy(e);
    for (auto &x : sd->sort_stream) if (x) { cudaStreamSynchronize(x); cudaStreamDestroy(x); }
    if (sd->stream) cudaStreamDestroy(sd->stream);
    cudaGetLastError();
    delete sd;
}

extern "C" int mm2gb_sketch_host(mm2gb_seeder_t *sd, const char *seqs, const int64_t *seq_off, int n_seq, int rid_is_seq, uint64_t *out_xy,
                                 int64_t cap, int64_t *mv_off)
{
    if (!sd || !seqs || !seq_off || !mv_off) return fail(MM2GB_EARG, "bad argument");
    CK(cudaSetDevice(sd->device));
    int rc = upload_offsets(sd, seq_off, n_seq);
    if (rc) return rc;
    rc = upload_seqs(sd, seqs, seq_off[n_seq]);
    if (rc) return rc;
    rc = run_sketch(sd, n_seq, rid_is_seq);
    if (rc) return rc;
    CK(cudaMemcpyAsync(mv_off, sd->d_mv_off, ((size_t)n_seq + 1) * sizeof(u64), cudaMemcpyDeviceToHost, sd->stream));
    const long long n = std::min<long long>(sd->n_mv, cap);
    std::vector<u64> x((size_t)n), y((size_t)n);
    if (n && out_xy) {
        CK(cudaMemcpyAsync(x.data(), sd->d_mv_x, (size_t)n * sizeof(u64), cudaMemcpyDeviceToHost, sd->stream));
        CK(cudaMemcpyAsync(y.data(), sd->d_mv_y, (size_t)n * sizeof(u64), cudaMemcpyDeviceToHost, sd->stream));
    }
    CK(cudaStreamSynchronize(sd->stream));
    if (out_xy) for (long long i = 0; i < n; ++i) out_xy[2 * i] = x[(size_t)i], out_xy[2 * i + 1] = y[(size_t)i];
    return MM2GB_OK;
}

extern "C" int mm2gb_seed_host(mm2gb_seeder_t *sd, const mm2gb_seed_params_t *prm, const char *seqs, const int64_t *seq_off, int n_reads,
                               mm2gb_anchor_t *a, int64_t a_cap, int64_t *a_off, int32_t *rep_len, uint64_t *mini_pos, int64_t mp_cap,
                               int64_t *mp_off)
{
    if (!sd || !seqs || !seq_off || !a_off) return fail(MM2GB_EARG, "bad argument");
    int rc = check_params(prm);
    if (rc) return rc;
    CK(cudaSetDevice(sd->device));
    rc = upload_offsets(sd, seq_off, n_reads);
    if (rc) return rc;
    rc = run_seed(sd, prm, seq_off, n_reads, mini_pos != nullptr, seqs);
    if (rc) return rc;
    if (a && sd->n_a > a_cap) return fail(MM2GB_ECAP, "%lld anchors do not fit the output (%lld)", sd->n_a, (long long)a_cap);
    if (mini_pos && sd->n_mp > mp_cap) return fail(MM2GB_ECAP, "%lld mini_pos entries do not fit the output (%lld)", sd->n_mp, (long long)mp_cap);
    if (a && sd->n_a) CK(cudaMemcpyAsync(a, sd->d_a, (size_t)sd->n_a * sizeof(mm2gb_anchor_t), cudaMemcpyDeviceToHost, sd->stream));
    if (mini_pos && sd->n_mp) CK(cudaMemcpyAsync(mini_pos, sd->d_mini_pos, (size_t)sd->n_mp * sizeof(u64), cudaMemcpyDeviceToHost, sd->stream));
    CK(cudaStreamSynchronize(sd->stream));
    // max_qlen (map.c:376) is applied by the caller of mm_map_seed's equivalent: reads above it are not in the batch
    for (int r = 0; r <= n_reads; ++r) a_off[r] = sd->h_a_off[r];
    if (mp_off) for (int r = 0; r <= n_reads; ++r) mp_off[r] = sd->h_mp_off[r];
    if (rep_len) for (int r = 0; r < n_reads; ++r) rep_len[r] = sd->h_rep_len[r];
    return MM2GB_OK;
}

static int seed_chain_common(mm2gb_seeder_t *sd, mm2gb_ctx_t *ctx, const mm2gb_seed_params_t *prm, const int64_t *seq_off, int n_reads, const char *host_seqs)
{
    int rc = run_seed(sd, prm, seq_off, n_reads, true, host_seqs);
    if (rc) return rc;
    // the chaining kernels run on the context's own stream: order them behind the seeding stream
    cudaStream_t cs = (cudaStream_t)mm2gb_stream(ctx, 0);
    if (!cs) return fail(MM2GB_EARG, "chaining context has no slot 0");
    CK(cudaEventRecord(sd->ev_join, sd->stream));
    CK(cudaStreamWaitEvent(cs, sd->ev_join, 0));
    sd->a_off_copy.assign(sd->h_a_off, sd->h_a_off + n_reads + 1);
    return mm2gb_chain_device_slot(ctx, 0, sd->d_a, sd->d_a_off, sd->a_off_copy.data(), n_reads, sd->n_a, sd->d_f, sd->d_p);
}

extern "C" int mm2gb_seed_chain(mm2gb_seeder_t *sd, mm2gb_ctx_t *ctx, const mm2gb_seed_params_t *prm, const char *seqs, const int64_t *seq_off,
                                int n_reads, mm2gb_seed_chain_result_t *res)
{
    if (!sd || !ctx || !seqs || !seq_off || !res) return fail(MM2GB_EARG, "bad argument");
    int rc = check_params(prm);
    if (rc) return rc;
    CK(cudaSetDevice(sd->device));
    rc = upload_offsets(sd, seq_off, n_reads);
    if (rc) return rc;
    if (!sd->h_b) {   // result landing buffers (pinned, mapped): allocated on first use, the parity / device-resident entries never need them
        mm2gb::NearGpu near_gpu(sd->device);   // pinned on the GPU's NUMA node
        CK(cudaHostAlloc((void **)&sd->h_b, (size_t)sd->max_anchors * sizeof(mm2gb_anchor_t), cudaHostAllocMapped));
        CK(cudaHostAlloc((void **)&sd->h_u, (size_t)sd->max_anchors * sizeof(uint64_t), cudaHostAllocMapped));
    }
    rc = seed_chain_common(sd, ctx, prm, seq_off, n_reads, seqs);
    if (rc) return rc;
    rc = mm2gb_chain_device_fetch(ctx, 0, sd->d_a, sd->d_a_off, n_reads, sd->n_a, sd->h_b, sd->h_u);
    if (rc) return rc;
    memset(res, 0, sizeof(*res));
    rc = mm2gb_chain_device_results(ctx, 0, &res->n_u, &res->u_pos, &res->n_b, &res->b_pos, &res->n_chains, &res->n_chain_anchors, &res->stats);
    if (rc) return rc;
    res->n_reads = n_reads;
    res->n_anchors = sd->n_a;
    res->a_off = (const int64_t *)sd->h_a_off;
    res->rep_len = sd->h_rep_len;
    res->u = sd->h_u;
    res->b = sd->h_b;
    res->h2d_bytes = seq_off[n_reads] + ((int64_t)n_reads + 1) * 12;
    res->d2h_bytes = res->n_chain_anchors * 16 + res->n_chains * 8 + ((int64_t)n_reads + 1) * (16 + 8 + 8 + 4) + 24;
    return MM2GB_OK;
}

extern "C" int mm2gb_seed_chain_device(mm2gb_seeder_t *sd, mm2gb_ctx_t *ctx, const mm2gb_seed_params_t *prm, const void *d_seqs,
                                       const int64_t *seq_off, int n_reads, int64_t *n_anchors)
{
    if (!sd || !ctx || !d_seqs || !seq_off) return fail(MM2GB_EARG, "bad argument");
    int rc = check_params(prm);
    if (rc) return rc;
    CK(cudaSetDevice(sd->device));
    rc = upload_offsets(sd, seq_off, n_reads);
    if (rc) return rc;
    unsigned char *own = sd->d_seq;
    sd->d_seq = (unsigned char *)const_cast<void *>(d_seqs);
    rc = seed_chain_common(sd, ctx, prm, seq_off, n_reads, nullptr);
    sd->d_seq = own;
    if (n_anchors) *n_anchors = sd->n_a;
    return rc;
}

extern "C" int mm2gb_seed_last_mini_pos(mm2gb_seeder_t *sd, int n_reads, uint64_t *mini_pos, int64_t cap, int64_t *mp_off)
{
    if (!sd || !mp_off || n_reads < 0 || n_reads > sd->max_reads) return fail(MM2GB_EARG, "bad argument");
    if (!sd->timed) return fail(MM2GB_ESTATE, "no batch has been seeded yet");
    CK(cudaSetDevice(sd->device));
    if (mini_pos && sd->n_mp > cap) return fail(MM2GB_ECAP, "%lld mini_pos entries do not fit the output (%lld)", sd->n_mp, (long long)cap);
    if (mini_pos && sd->n_mp) CK(cudaMemcpyAsync(mini_pos, sd->d_mini_pos, (size_t)sd->n_mp * sizeof(u64), cudaMemcpyDeviceToHost, sd->stream));
    CK(cudaStreamSynchronize(sd->stream));
    for (int r = 0; r <= n_reads; ++r) mp_off[r] = sd->h_mp_off[r];
    return MM2GB_OK;
}

extern "C" int mm2gb_seed_profile(mm2gb_seeder_t *sd, float ms[MM2GB_SEED_NTIMERS], int64_t *n_minimizers, int64_t *n_seeds)
{
    if (!sd || !ms) return fail(MM2GB_EARG, "bad argument");
    if (!sd->timed) return fail(MM2GB_ESTATE, "no batch has been seeded yet");
    CK(cudaSetDevice(sd->device));
    CK(cudaEventSynchronize(sd->ev[MM2GB_SEED_NTIMERS]));
    for (int i = 0; i < MM2GB_SEED_NTIMERS; ++i) CK(cudaEventElapsedTime(&ms[i], sd->ev[i], sd->ev[i + 1]));
    if (n_minimizers) *n_minimizers = sd->n_mv;
    if (n_seeds) *n_seeds = sd->n_m;
    return MM2GB_OK;
}

// ---- C ABI: index ------------------------------------------------------------------------------------------------------------

extern "C" void mm2gb_index_destroy(mm2gb_index_t *ix)
{
    if (!ix) return;
    cudaSetDevice(ix->device);
    if (ix->d_key) cudaFree(ix->d_key);
    if (ix->d_val) cudaFree(ix->d_val);
    if (ix->d_occ) cudaFree(ix->d_occ);
    cudaGetLastError();
    delete ix;
}

// host lists (keys ascending, off, occ) -> open-addressing table + occurrence array in HBM; destroys ix on failure
static int index_upload(mm2gb_index *ix)
{
    const size_t nk = ix->keys.size();
    size_t slots = 1024;
    while (slots < 2 * nk) slots <<= 1;
    ix->mask = slots - 1;
    std::vector<uint64_t> &hk = ix->hk, &hv = ix->hv;
    hk.assign(slots, kNone); hv.assign(slots, 0);
    for (size_t i = 0; i < nk; ++i) {
        const u64 cnt = ix->off[i + 1] - ix->off[i];
        if (cnt >= (1ULL << 28) || ix->off[i] >= (1ULL << 36)) { delete ix; return fail(MM2GB_ECAP, "index too large for the table encoding"); }
        u64 key = ix->keys[i], h = key;
        h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;   // mix64 of seed_kernels.cuh
        h &= ix->mask;
        while (hk[h] != kNone) h = (h + 1) & ix->mask;
        hk[h] = key; hv[h] = ix->off[i] << 28 | cnt;
    }
    auto bad = [&](cudaError_t e, const char *what) { int rc = fail(MM2GB_ECUDA, "%s: %s", what, cudaGetErrorString(e)); mm2gb_index_destroy(ix); return rc; };
    cudaError_t e;
    if ((e = cudaMalloc((void **)&ix->d_key, slots * sizeof(u64))) != cudaSuccess) return bad(e, "cudaMalloc(index keys)");
    if ((e = cudaMalloc((void **)&ix->d_val, slots * sizeof(u64))) != cudaSuccess) return bad(e, "cudaMalloc(index values)");
    if ((e = cudaMalloc((void **)&ix->d_occ, std::max<size_t>(ix->occ.size(), 1) * sizeof(u64))) != cudaSuccess) return bad(e, "cudaMalloc(index occurrences)");
    if ((e = cudaMemcpy(ix->d_key, hk.data(), slots * sizeof(u64), cudaMemcpyHostToDevice)) != cudaSuccess) return bad(e, "cudaMemcpy(index keys)");
    if ((e = cudaMemcpy(ix->d_val, hv.data(), slots * sizeof(u64), cudaMemcpyHostToDevice)) != cudaSuccess) return bad(e, "cudaMemcpy(index values)");
    if (!ix->occ.empty() && (e = cudaMemcpy(ix->d_occ, ix->occ.data(), ix->occ.size() * sizeof(u64), cudaMemcpyHostToDevice)) != cudaSuccess)
        return bad(e, "cudaMemcpy(index occurrences)");
    return MM2GB_OK;
}

extern "C" int mm2gb_index_build(mm2gb_index_t **out, int device, const char *seqs, const int64_t *seq_off, int n_seq, int w, int k, int is_hpc,
                                 int bucket_bits)
{
    (void)bucket_bits;
    if (!out || !seqs || !seq_off || n_seq <= 0) return fail(MM2GB_EARG, "bad argument");
    *out = nullptr;
    if (k < 1 || k > 28 || !(k & 1)) return fail(MM2GB_EARG, "device seeding needs an odd k <= 28 (got %d)", k);
    if (w < 1 || w > kMaxW) return fail(MM2GB_EARG, "device seeding needs 1 <= w <= %d (got %d)", kMaxW, w);
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(MM2GB_EARG, "no CUDA device %d (have %d)", device, ndev);
    CK(cudaSetDevice(device));
    mm2gb_index *ix = new mm2gb_index();
    ix->device = device; ix->w = w; ix->k = k; ix->hpc = is_hpc != 0;
    // sketch every sequence with the read kernel (rid = sequence number), in pieces of at most 256 M bases
    std::vector<std::pair<u64, u64>> mz;     // (minimizer = x >> 8, y)
    {
        const int64_t piece = (int64_t)256 << 20;
        int s0 = 0;
        while (s0 < n_seq) {
            int s1 = s0;
            int64_t bases = 0;
            while (s1 < n_seq && (s1 == s0 || bases + (seq_off[s1 + 1] - seq_off[s1]) <= piece)) { bases += seq_off[s1 + 1] - seq_off[s1]; ++s1; }
            mm2gb_seeder *sd = nullptr;
            // a throw-away seeder sized for this piece (only its sketch buffers matter: ask for the smallest anchor capacity)
            int rc = mm2gb_seeder_create(&sd, ix, std::max<int64_t>(bases, 1), s1 - s0, 1024);
            if (rc) { delete ix; return rc; }
            std::vector<int64_t> off((size_t)(s1 - s0) + 1), mvo((size_t)(s1 - s0) + 1);
            for (int s = s0; s <= s1; ++s) off[(size_t)(s - s0)] = seq_off[s] - seq_off[s0];
            std::vector<uint64_t> xy((size_t)2 * sd->max_mv);
            rc = mm2gb_sketch_host(sd, seqs + seq_off[s0], off.data(), s1 - s0, 1, xy.data(), sd->max_mv, mvo.data());
            const long long n = sd->n_mv;
            mm2gb_seeder_destroy(sd);
            if (rc) { delete ix; return rc; }
            mz.reserve(mz.size() + (size_t)n);
            for (long long i = 0; i < n; ++i) mz.emplace_back(xy[2 * i] >> 8, xy[2 * i + 1] + ((u64)s0 << 32));
            s0 = s1;
        }
    }
    // group: by minimizer, positions ascending (index.c:224,253).  Sorted in slices on host threads (set-up, once per index)
    {
        const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        const size_t n = mz.size();
        if (n < (1u << 16) || nt == 1) std::sort(mz.begin(), mz.end());
        else {
            std::vector<size_t> cut(nt + 1);
            for (unsigned t = 0; t <= nt; ++t) cut[t] = n * t / nt;
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t]() { std::sort(mz.begin() + (long)cut[t], mz.begin() + (long)cut[t + 1]); });
            for (auto &x : th) x.join();
            for (unsigned step = 1; step < nt; step *= 2) {
                std::vector<std::thread> mt;
                for (unsigned t = 0; t + step < nt; t += 2 * step)
                    mt.emplace_back([&, t, step]() {
                        std::inplace_merge(mz.begin() + (long)cut[t], mz.begin() + (long)cut[t + step], mz.begin() + (long)cut[std::min(nt, t + 2 * step)]);
                    });
                for (auto &x : mt) x.join();
            }
        }
    }
    ix->occ.resize(mz.size());
    for (size_t i = 0; i < mz.size(); ++i) {
        if (i == 0 || mz[i].first != mz[i - 1].first) { ix->keys.push_back(mz[i].first); ix->off.push_back(i); }
        ix->occ[i] = mz[i].second;
    }
    ix->off.push_back(mz.size());
    std::vector<std::pair<u64, u64>>().swap(mz);
    {
        const int rc = index_upload(ix);
        if (rc) return rc;
    }
    *out = ix;
    return MM2GB_OK;
}

extern "C" int mm2gb_index_from_lists(mm2gb_index_t **out, int device, int w, int k, int is_hpc, int64_t n_keys, const uint64_t *keys,
                                      const int64_t *off, const uint64_t *occ)
{
    if (!out || n_keys < 0 || (n_keys && (!keys || !off || !occ))) return fail(MM2GB_EARG, "bad argument");
    *out = nullptr;
    if (k < 1 || k > 28 || !(k & 1)) return fail(MM2GB_EARG, "device seeding needs an odd k <= 28 (got %d)", k);
    if (w < 1 || w > kMaxW) return fail(MM2GB_EARG, "device seeding needs 1 <= w <= %d (got %d)", kMaxW, w);
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(MM2GB_EARG, "no CUDA device %d (have %d)", device, ndev);
    CK(cudaSetDevice(device));
    mm2gb_index *ix = new mm2gb_index();
    ix->device = device; ix->w = w; ix->k = k; ix->hpc = is_hpc != 0;
    // keys arrive in the enumeration order of the host hash tables and stay in it: the device side is a hash table anyway, and the
    // host-side lookups (mm2gb_index_get: tests) go through the host copy of that table
    ix->keys.assign(keys, keys + n_keys);
    ix->off.resize((size_t)n_keys + 1);
    for (int64_t i = 0; i <= n_keys; ++i) ix->off[(size_t)i] = n_keys ? (uint64_t)off[i] : 0;
    ix->occ.assign(occ, occ + (n_keys ? off[n_keys] : 0));
    ix->keys_sorted = false;
    const int rc = index_upload(ix);
    if (rc) return rc;
    *out = ix;
    return MM2GB_OK;
}

extern "C" int32_t mm2gb_index_cal_max_occ(const mm2gb_index_t *ix, float f)
{
    if (!ix || f <= 0.f) return INT32_MAX;           // index.c:192
    const size_t n = ix->keys.size();
    if (n == 0) return INT32_MAX;
    std::vector<uint32_t> a(n);
    for (size_t i = 0; i < n; ++i) a[i] = (uint32_t)(ix->off[i + 1] - ix->off[i]);
    size_t kk = (uint32_t)((1. - f) * n);           // index.c:204 (double arithmetic, truncated to 32 bits)
    if (kk >= n) kk = n - 1;
    std::nth_element(a.begin(), a.begin() + (long)kk, a.end());
    return (int32_t)(a[kk] + 1);
}

extern "C" int64_t mm2gb_index_get(const mm2gb_index_t *ix, uint64_t minier, uint64_t *out, int64_t cap)
{
    if (!ix || ix->hk.empty()) return 0;
    uint64_t h = minier;
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;   // the probe k_lookup does
    h &= ix->mask;
    while (ix->hk[h] != minier) { if (ix->hk[h] == kNone) return 0; h = (h + 1) & ix->mask; }
    const uint64_t o = ix->hv[h] >> 28;
    const int64_t n = (int64_t)(ix->hv[h] & 0xfffffffu);
    for (int64_t t = 0; t < n && t < cap && out; ++t) out[t] = ix->occ[o + (size_t)t];
    return n;
}

extern "C" int64_t mm2gb_index_n_keys(const mm2gb_index_t *ix) { return ix ? (int64_t)ix->keys.size() : 0; }
extern "C" int64_t mm2gb_index_n_occ(const mm2gb_index_t *ix) { return ix ? (int64_t)ix->occ.size() : 0; }
