// plchain_dropin.cpp -- the four entry points of minimap2's --gpu-chain boundary on top of libmm2gb_chain.
//
// Replaces the host half of the reference's GPU layer: gpu/plchain.cu:292-560 (stream state machine, micro-batching,
// host sort of long segments, host backtracking) and gpu/plmem.cu:373-540 (JSON config, buffer sizing).
// What is kept is the protocol the unmodified driver relies on (map.c:1016-1071, kthread.c:52-55):
//   * chain_stream_gpu launches the incoming batch asynchronously and returns the batch launched by the previous call of
//     the same thread_id, fully chained (a / u / n_u set from the arena `km`, post_chaining_helper run);
//   * finish_stream_gpu drains; free_stream_gpu tears down and is a no-op when nothing was initialised.
// What is different:
//   * the new batch is submitted BEFORE the previous one is finished on the host, so the GPU chains batch k+1 while this
//     thread backtracks batch k (the reference synchronises first, plchain.cu:299-305);
//   * every thread_id owns a context (2 slots: stream + pinned staging + device buffers) on GPU thread_id % n_gpus, so
//     `-t N` and several GPUs work (the reference is limited to -t 1 / one stream, README.md:46-47);
//   * anchors are gathered straight from chain_read_t.a into pinned memory -- no AoS->SoA repack (plmem.cu:154-198);
//   * no read is ever handed back for CPU chaining (plchain.cu:421-423): oversized batches grow the context instead;
//   * --max-chain-skip is ignored, i.e. true infinity (SURVEY.md trap T1), as in the reference's kernels;
//   * chain extraction + compaction (lchain.c:27-111) run on the device behind the DP kernels (k_bt_sort* / k_bt_walk*); the calling
//     thread only copies the results into the kalloc arena (kalloc is not thread-safe).  "host_backtrack": 1 in the config
//     (or MM2GB_HOST_BACKTRACK=1) moves that stage to a small host thread pool instead (csrc/backtrack.cpp).
#include "../../include/mm2gb_plchain.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr int kMaxThreads = 256;

struct Config {
    size_t max_total_n = 64u << 20; // anchors per launched batch
    int max_read = 200000;          // reads per batch
    int min_n = 0;                  // plumbed through, unused downstream (map.c:1314)
    int n_gpus = 0;                 // 0 = all visible
    int n_slots = 2;
    int host_threads = 8;
    int host_backtrack = 0;         // 1 = chain extraction on host threads instead of the device kernel
};

struct ThreadState {
    mm2gb_ctx_t *ctx = nullptr;
    int device = 0;
    size_t cap_anchors = 0;
    int cap_reads = 0;
    Misc_abi misc;
    bool has_misc = false;
    // the batch in flight
    bool busy = false;
    mm2gb_chain_read_t *reads = nullptr;
    int n_reads = 0;
    int slot = 0;
    bool submitted = false; // false for batches without anchors
    // scratch of the host stage
    std::vector<uint64_t> su;
    std::vector<mm2gb_anchor_t> sb;
    std::vector<int32_t> s_nu;
    std::vector<int64_t> s_nb, s_off;
    std::vector<const mm2gb_anchor_t *> ptrs;
    std::vector<int64_t> ns;
};

Config g_cfg;
bool g_inited = false;
Misc_abi g_misc;
std::mutex g_mu;
ThreadState *g_state[kMaxThreads];

[[noreturn]] void die(const char *what)
{
    fprintf(stderr, "[ERROR] mm2gb chaining: %s: %s\n", what, mm2gb_last_error());
    exit(1);
}

// minimal reader for the flat numeric keys of gpu/gpu_config.json-style files
bool json_number(const std::string &txt, const char *key, double *out)
{
    const std::string pat = std::string("\"") + key + "\"";
    size_t at = 0;
    while ((at = txt.find(pat, at)) != std::string::npos) {
        size_t q = at + pat.size();
        while (q < txt.size() && (txt[q] == ' ' || txt[q] == '\t' || txt[q] == '\n' || txt[q] == '\r')) ++q;
        if (q < txt.size() && txt[q] == ':') {
            char *end = nullptr;
            const double v = strtod(txt.c_str() + q + 1, &end);
            if (end != txt.c_str() + q + 1) { *out = v; return true; }
        }
        at += pat.size();
    }
    return false;
}

void load_config(const char *path)
{
    g_cfg = Config();
    std::string txt;
    if (path && *path) {
        if (FILE *fp = fopen(path, "rb")) {
            char buf[4096];
            size_t n;
            while ((n = fread(buf, 1, sizeof(buf), fp)) > 0) txt.append(buf, n);
            fclose(fp);
        } else {
            fprintf(stderr, "[WARNING] mm2gb chaining: cannot open gpu config '%s'; using built-in defaults\n", path);
        }
    }
    double v;
    if (json_number(txt, "max_total_n", &v) && v >= 1) g_cfg.max_total_n = (size_t)v;
    if (json_number(txt, "max_read", &v) && v >= 1) g_cfg.max_read = (int)v;
    if (json_number(txt, "min_n", &v) && v >= 0) g_cfg.min_n = (int)v;
    if (json_number(txt, "n_gpus", &v) && v >= 0) g_cfg.n_gpus = (int)v;
    if (json_number(txt, "n_slots", &v) && v >= 2 && v <= 4) g_cfg.n_slots = (int)v;
    if (json_number(txt, "host_threads", &v) && v >= 1) g_cfg.host_threads = (int)v;
    if (json_number(txt, "host_backtrack", &v)) g_cfg.host_backtrack = v != 0;
    if (const char *e = getenv("MM2GB_HOST_BACKTRACK")) g_cfg.host_backtrack = atoi(e) != 0;
    if (const char *e = getenv("MM2GB_HOST_THREADS")) g_cfg.host_threads = atoi(e) > 0 ? atoi(e) : g_cfg.host_threads;
    if (const char *e = getenv("MM2GB_N_GPUS")) g_cfg.n_gpus = atoi(e) > 0 ? atoi(e) : g_cfg.n_gpus;
    if (g_cfg.max_total_n > ((size_t)1 << 31) - 2048) g_cfg.max_total_n = ((size_t)1 << 31) - 2048;
    const int ndev = mm2gb_device_count();
    if (ndev <= 0) { fprintf(stderr, "[ERROR] mm2gb chaining: --gpu-chain needs a CUDA device (no CPU fallback)\n"); exit(1); }
    if (g_cfg.n_gpus <= 0 || g_cfg.n_gpus > ndev) g_cfg.n_gpus = ndev;
}

bool same_misc(const Misc_abi &a, const Misc_abi &b) { return memcmp(&a, &b, sizeof(Misc_abi)) == 0; }

void make_ctx(ThreadState &S, size_t cap_anchors, int cap_reads, const Misc_abi &misc)
{
    if (S.ctx) mm2gb_ctx_destroy(S.ctx);
    S.ctx = nullptr;
    if (mm2gb_ctx_create(&S.ctx, S.device, cap_anchors, cap_reads, g_cfg.n_slots, &misc) != MM2GB_OK) die("cannot create the chaining context");
    S.cap_anchors = cap_anchors;
    S.cap_reads = cap_reads;
    S.misc = misc;
    S.has_misc = true;
}

ThreadState &state_of(int tid)
{
    if (tid < 0 || tid >= kMaxThreads) { fprintf(stderr, "[ERROR] mm2gb chaining: thread id %d out of range\n", tid); exit(1); }
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_inited) { fprintf(stderr, "[ERROR] mm2gb chaining: chain_stream_gpu before init_stream_gpu\n"); exit(1); }
    if (!g_state[tid]) {
        g_state[tid] = new ThreadState();
        g_state[tid]->device = tid % g_cfg.n_gpus;
    }
    return *g_state[tid];
}

// finish the batch in flight: wait for the device, publish the chains into the arena, run the driver's helper.
// Chain extraction + compaction (lchain.c:27-111) ran on the device behind the DP kernels (g_cfg.host_backtrack == 0,
// default) or runs here on a small thread pool (host_backtrack == 1: the reference's arrangement, plchain.cu:99-150).
void complete_inflight(ThreadState &S, const mm2gb_idx_t *mi, const mm2gb_mapopt_t *opt, const Misc_abi &misc, void *km)
{
    mm2gb_chain_read_t *reads = S.reads;
    const int n_reads = S.n_reads;
    const int64_t *off = nullptr;
    const uint64_t *const *dev_u = nullptr;
    const int32_t *dev_nu = nullptr, *dev_nb = nullptr;
    const mm2gb_anchor_t *const *dev_b = nullptr;
    if (S.submitted && !g_cfg.host_backtrack) {
        if (mm2gb_wait_chains(S.ctx, S.slot, &dev_u, &dev_nu, &dev_b, &dev_nb, &off, nullptr) != MM2GB_OK) die("waiting for a chaining batch");
    } else if (S.submitted) {
        const int32_t *f = nullptr, *p = nullptr;
        if (mm2gb_wait(S.ctx, S.slot, &f, &p, &off, nullptr) != MM2GB_OK) die("waiting for a chaining batch");
        const int64_t total = off[n_reads];
        if ((int64_t)S.su.size() < total) { S.su.resize((size_t)total); S.sb.resize((size_t)total); }
        S.s_nu.assign((size_t)n_reads, 0);
        S.s_nb.assign((size_t)n_reads, 0);
        const int32_t max_drop = misc.is_cdna ? INT32_MAX : misc.bw; // lchain.c:151,162
        std::atomic<int> next(0);
        auto work = [&]() {
            for (;;) {
                const int r = next.fetch_add(1);
                if (r >= n_reads) return;
                const int64_t s = off[r], n = off[r + 1] - s;
                if (n <= 0) continue;
                S.s_nu[(size_t)r] = mm2gb_backtrack(n, f + s, p + s, reads[r].a, misc.min_cnt, misc.min_score, max_drop,
                                                    S.su.data() + s, S.sb.data() + s, &S.s_nb[(size_t)r]);
            }
        };
        const int nt = std::max(1, std::min(g_cfg.host_threads, n_reads));
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; ++t) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
    }
    for (int r = 0; r < n_reads; ++r) { // arena traffic stays on the calling thread
        mm2gb_chain_read_t &rd = reads[r];
        int32_t n_u = 0;
        int64_t n_b = 0;
        const uint64_t *src_u = nullptr;
        const mm2gb_anchor_t *src_b = nullptr;
        if (S.submitted && dev_nu) { n_u = dev_nu[r]; n_b = dev_nb[r]; src_u = dev_u[r]; src_b = dev_b[r]; }
        else if (S.submitted) { n_u = S.s_nu[(size_t)r]; n_b = S.s_nb[(size_t)r]; src_u = S.su.data() + off[r]; src_b = S.sb.data() + off[r]; }
        if (n_u > 0) {
            uint64_t *u = (uint64_t *)kmalloc(km, (size_t)n_u * sizeof(uint64_t));
            memcpy(u, src_u, (size_t)n_u * sizeof(uint64_t));
            mm2gb_anchor_t *b = (mm2gb_anchor_t *)kmalloc(km, (size_t)n_b * sizeof(mm2gb_anchor_t));
            memcpy(b, src_b, (size_t)n_b * sizeof(mm2gb_anchor_t));
            kfree(km, rd.a); // compact_a frees the oversized input array (lchain.c:107-109)
            rd.a = b; rd.u = u; rd.n_u = n_u;
        } else {             // lchain.c:212-215 / plchain.cu:135-143
            kfree(km, rd.a);
            rd.a = nullptr; rd.u = nullptr; rd.n_u = 0;
        }
        post_chaining_helper(mi, opt, &rd, misc, km);
    }
    S.busy = false;
    S.reads = nullptr;
    S.n_reads = 0;
}

} // namespace

extern "C" void init_stream_gpu(size_t *max_total_n, int *max_reads, int *min_n, char gpu_config_file[], Misc_abi misc)
{
    std::lock_guard<std::mutex> lk(g_mu);
    load_config(gpu_config_file);
    g_misc = misc;
    g_inited = true;
    if (max_total_n) *max_total_n = g_cfg.max_total_n;
    if (max_reads) *max_reads = g_cfg.max_read;
    if (min_n) *min_n = g_cfg.min_n;
}

extern "C" void chain_stream_gpu(const mm2gb_idx_t *mi, const mm2gb_mapopt_t *opt, mm2gb_chain_read_t **in_arr_, int *n_read_, int thread_id, void *km)
{
    ThreadState &S = state_of(thread_id);
    const Misc_abi misc = build_misc(mi, opt, 0, 1); // single segment, qlen_sum irrelevant (plchain.cu:498-500)
    mm2gb_chain_read_t *in = in_arr_ ? *in_arr_ : nullptr;
    const int n_in = (n_read_ && in) ? *n_read_ : 0;

    int64_t total = 0;
    S.ptrs.resize((size_t)n_in);
    S.ns.resize((size_t)n_in);
    for (int r = 0; r < n_in; ++r) {
        S.ptrs[(size_t)r] = in[r].a;
        S.ns[(size_t)r] = in[r].a ? in[r].n : 0;
        total += S.ns[(size_t)r];
    }
    // (re)size the context: first use, a batch beyond the configured limits, or new chaining parameters
    const bool need_grow = !S.ctx || (size_t)total > S.cap_anchors || n_in > S.cap_reads;
    const bool new_misc = S.ctx && !same_misc(S.misc, misc);
    mm2gb_chain_read_t *prev = S.busy ? S.reads : nullptr;
    const int n_prev = S.busy ? S.n_reads : 0;
    const bool had_prev = S.busy;
    if ((need_grow || new_misc) && S.busy) complete_inflight(S, mi, opt, S.misc, km); // old buffers / parameters still in use
    if (need_grow) {
        size_t cap = std::max(g_cfg.max_total_n, (size_t)total + (size_t)total / 2);
        if (cap > ((size_t)1 << 31) - 2048) cap = ((size_t)1 << 31) - 2048;
        if ((size_t)total > cap) { fprintf(stderr, "[ERROR] mm2gb chaining: a batch of %lld anchors cannot be indexed with 31 bits\n", (long long)total); exit(1); }
        make_ctx(S, cap, std::max(g_cfg.max_read, n_in + n_in / 2) + 1, misc);
    } else if (new_misc) {
        if (mm2gb_ctx_set_misc(S.ctx, &misc) != MM2GB_OK) die("updating the chaining parameters");
        S.misc = misc;
    }
    // launch the new batch first, so the device works while this thread finishes the previous one
    const int slot = S.busy ? (S.slot + 1) % g_cfg.n_slots : 0;
    bool submitted = false;
    if (in && total > 0) {
        const int rc = g_cfg.host_backtrack ? mm2gb_submit_gather(S.ctx, slot, S.ptrs.data(), S.ns.data(), n_in)
                                            : mm2gb_submit_gather_chains(S.ctx, slot, S.ptrs.data(), S.ns.data(), n_in);
        if (rc != MM2GB_OK) die("launching a chaining batch");
        submitted = true;
    }
    if (S.busy) complete_inflight(S, mi, opt, misc, km);
    if (in) {
        S.busy = true; S.reads = in; S.n_reads = n_in; S.slot = slot; S.submitted = submitted;
    }
    if (in_arr_) *in_arr_ = had_prev ? prev : nullptr;
    if (n_read_) *n_read_ = had_prev ? n_prev : 0;
}

extern "C" void finish_stream_gpu(const mm2gb_idx_t *mi, const mm2gb_mapopt_t *opt, mm2gb_chain_read_t **reads_, int *n_read_, int thread_id, void *km)
{
    ThreadState &S = state_of(thread_id);
    if (!S.busy) {
        if (reads_) *reads_ = nullptr;
        if (n_read_) *n_read_ = 0;
        return;
    }
    mm2gb_chain_read_t *prev = S.reads;
    const int n_prev = S.n_reads;
    const Misc_abi misc = build_misc(mi, opt, 0, 1);
    complete_inflight(S, mi, opt, misc, km);
    if (reads_) *reads_ = prev;
    if (n_read_) *n_read_ = n_prev;
}

extern "C" void free_stream_gpu(int n_threads)
{
    (void)n_threads;
    std::lock_guard<std::mutex> lk(g_mu);
    for (int t = 0; t < kMaxThreads; ++t) {
        if (!g_state[t]) continue;
        if (g_state[t]->ctx) mm2gb_ctx_destroy(g_state[t]->ctx);
        delete g_state[t];
        g_state[t] = nullptr;
    }
    g_inited = false;
}
