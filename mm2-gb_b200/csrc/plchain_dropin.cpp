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
//     thread publishes batch k (the reference synchronises first, plchain.cu:299-305);
//   * a batch is cut into up to four sub-batches that go through their own slots (stream + pinned staging + device buffers):
//     while this thread gathers sub-batch j+1 into pinned memory, sub-batch j is already uploading / being chained, and on the
//     way back sub-batch j is published while j+1 is still on the GPU;
//   * every thread_id owns a context on GPU thread_id % n_gpus, so `-t N` and several GPUs work (the reference is limited to
//     -t 1 / one stream, README.md:46-47);
//   * the gather pass over chain_read_t.a writes the packed 8-byte wire format (csrc/wire.h) -- the reference's pass is an
//     AoS->SoA repack of 13 B/anchor (plmem.cu:154-198) -- and what comes back are chains and the INDICES of their anchors:
//     compact_a's gather (lchain.c:100-105) runs here, from the read's own array, straight into the kmalloc'd result;
//   * no read is ever handed back for CPU chaining (plchain.cu:421-423) and there is no host chaining or backtracking path:
//     oversized batches grow the context instead; without a CUDA device every call fails;
//   * --max-chain-skip is ignored, i.e. true infinity (SURVEY.md trap T1), as in the reference's kernels.
#include "../../include/mm2gb_plchain.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

namespace {

constexpr int kMaxSub = 4;          // sub-batches of one batch (2 batches in flight x 4 = the 8 slots of a context)

struct Config {
    size_t max_total_n = 4u << 20;  // anchors per launched batch
    int max_read = 200000;          // reads per batch
    int min_n = 0;                  // plumbed through, unused downstream (map.c:1314)
    int n_gpus = 0;                 // 0 = all visible
    int gpu_base = 0;               // first GPU used (thread_id t runs on GPU gpu_base + t % n_gpus)
    int sub_batches = kMaxSub;      // 1..4
    int threads_per_gpu = 4;        // sizing hint: how many driver threads share a GPU (minimap2 -t N / n_gpus)
    int64_t sub_min = 1 << 19;      // a batch is only cut into sub-batches of at least this many anchors (enough to fill the GPU)
    int sub_min_reads = 64;         // ... and at least this many reads
};

struct Sub { int slot, r0, r1; };

struct ThreadState {
    mm2gb_ctx_t *ctx = nullptr;
    int device = 0;
    size_t cap_anchors = 0;         // capacity of ONE slot
    int cap_reads = 0;
    Misc_abi misc;
    bool has_misc = false;
    // the batch in flight
    bool busy = false;
    mm2gb_chain_read_t *reads = nullptr;
    int n_reads = 0;
    int half = 0;                   // which half of the slots it uses
    std::vector<Sub> subs;          // its sub-batches (empty for a batch without anchors)
    // scratch
    std::vector<const mm2gb_anchor_t *> ptrs;
    std::vector<int64_t> ns;
    // where this thread's time at the boundary goes (seconds; reported by free_stream_gpu with MM2GB_VERBOSE=1)
    double t_ctx = 0, t_submit = 0, t_wait = 0, t_publish = 0;
    long long n_batches = 0, n_anchors = 0, n_ctx = 0;
};

inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
const double g_t_load = now_s();    // when the library was loaded (= process start for a statically linked driver)
int verbose_level() { static const int v = getenv("MM2GB_VERBOSE") ? atoi(getenv("MM2GB_VERBOSE")) : 0; return v; }
#define VLOG(lvl, ...) do { if (verbose_level() >= (lvl)) { fprintf(stderr, "[mm2gb %8.3f] ", now_s() - g_t_load); fprintf(stderr, __VA_ARGS__); fputc('\n', stderr); } } while (0)

Config g_cfg;
bool g_inited = false;
Misc_abi g_misc;
std::mutex g_mu;
std::vector<ThreadState *> g_state;
std::atomic<long long> g_h2d_bytes(0), g_d2h_bytes(0);   // bytes that crossed PCIe since the last reset (mm2gb_dropin_traffic)
constexpr int kMaxDev = 64;
std::atomic<long long> g_dev_batches[kMaxDev];           // batches launched per GPU since init_stream_gpu (mm2gb_dropin_device_batches)

// ---- start-up off the critical path -------------------------------------------------------------------------------------
// Bringing CUDA up (driver initialisation, a primary context per GPU, loading the kernels) takes 1-3 s per process and a
// chaining context ~50 ms more per driver thread, all serialised inside the driver -- on a 10 k-read input that was most of the
// chaining stage (profiles/r4h_driver_ont.json: 3.2 s per thread).  None of it depends on the reads, so it runs on background
// threads while the driver is still busy elsewhere:
//   * early_warm: started by a load-time constructor when the process was started with --gpu-chain (or MM2GB_EARLY_INIT=1):
//     initialises every visible GPU while the driver parses options and loads / builds the index;
//   * the context pool: started by init_stream_gpu (the batch limits are known then): creates threads_per_gpu contexts per GPU,
//     sized for the configured batch limit, while the driver reads and seeds its first mini-batch.  A driver thread takes a
//     ready context on its first batch (waiting for one that is still being created) and only creates its own when the pool
//     is used up or the batch does not fit.
std::thread *g_early = nullptr, *g_pool_thread = nullptr;   // heap objects, never destroyed: exit(1) on an error path must not meet a joinable std::thread
std::mutex g_pool_mu;
std::condition_variable g_pool_cv;
struct PoolCtx { mm2gb_ctx_t *ctx; size_t cap_anchors; int cap_reads; };
std::vector<PoolCtx> g_ready[kMaxDev];
int g_planned[kMaxDev];            // contexts the pool thread will still deliver, per GPU
bool g_pool_stop = false;
std::atomic<bool> g_early_done(true);   // false while the early warm-up thread is still bringing the GPUs up

Misc_abi warm_misc()
{
    Misc_abi m;
    memset(&m, 0, sizeof(m));
    m.max_iter = 5000; m.max_dist_x = 5000; m.max_dist_y = 5000; m.max_skip = 25; m.bw = 500; m.min_cnt = 3; m.min_score = 40; m.n_seg = 1;
    m.chn_pen_gap = 0.12f; m.chn_pen_skip = 0.0f;
    return m;
}

void warm_device(int d)
{
    mm2gb_ctx_t *tmp = nullptr;
    const Misc_abi m = warm_misc();
    if (mm2gb_ctx_create_ex(&tmp, d, 1 << 16, 16, 1, &m, MM2GB_CTX_NO_FP_STAGING) == MM2GB_OK) mm2gb_ctx_destroy(tmp);
}

void early_warm()
{
    const int ndev = mm2gb_device_count();
    int n = ndev;
    if (const char *e = getenv("MM2GB_N_GPUS")) if (atoi(e) > 0) n = std::min(ndev, atoi(e));
    const int base = getenv("MM2GB_GPU_BASE") ? std::max(0, atoi(getenv("MM2GB_GPU_BASE"))) : 0;
    VLOG(2, "early warm-up: CUDA is up, %d device(s)", ndev);
    for (int d = base; d < std::min(ndev, base + n); ++d) warm_device(d);
    VLOG(2, "early warm-up: done");
    g_early_done = true;
}

struct EarlyInit {
    EarlyInit()
    {
        bool want = getenv("MM2GB_EARLY_INIT") && atoi(getenv("MM2GB_EARLY_INIT")) != 0;
        if (!want && !getenv("MM2GB_EARLY_INIT")) {
            if (FILE *fp = fopen("/proc/self/cmdline", "rb")) {
                char buf[8192];
                const size_t n = fread(buf, 1, sizeof(buf) - 1, fp);
                fclose(fp);
                buf[n] = 0;
                for (size_t i = 0; i < n; i += strlen(buf + i) + 1)
                    if (!strcmp(buf + i, "--gpu-chain")) { want = true; break; }
            }
        }
        if (want) { g_early_done = false; g_early = new std::thread(early_warm); }
    }
} g_early_init;

std::mutex g_early_mu;
void join_early()
{
    std::lock_guard<std::mutex> lk(g_early_mu);     // callable from every driver thread at once
    if (g_early && g_early->joinable()) g_early->join();
}

void stop_pool()
{
    {
        std::lock_guard<std::mutex> pl(g_pool_mu);
        g_pool_stop = true;
        g_pool_cv.notify_all();
    }
    if (g_pool_thread && g_pool_thread->joinable()) g_pool_thread->join();
    delete g_pool_thread;
    g_pool_thread = nullptr;
    std::lock_guard<std::mutex> pl(g_pool_mu);
    for (int d = 0; d < kMaxDev; ++d) {
        for (auto &pc : g_ready[d]) mm2gb_ctx_destroy(pc.ctx);
        g_ready[d].clear();
        g_planned[d] = 0;
    }
}

// Fatal error: message already on stderr, exit status 1 -- as the reference (hipify.cuh:47-55).  Other threads (the context pool, the
// driver threads of other GPUs) may be inside CUDA calls: running the atexit handlers / static destructors under them tears the
// runtime down while it is in use (seen as a SIGSEGV instead of status 1), so every stdio stream is flushed and the process ends
// without them.
[[noreturn]] void fatal_exit()
{
    fflush(nullptr);
    _exit(1);
}

[[noreturn]] void die(const char *what)
{
    fprintf(stderr, "[ERROR] mm2gb chaining: %s: %s\n", what, mm2gb_last_error());
    fatal_exit();
}

// ---- gpu config file: the reference's gpu/*.json (parsed there with cJSON, gpu/plmem.cu:373-451) --------------------------
// A small recursive-descent JSON reader that collects the NUMERIC members of the TOP-LEVEL object only; nested objects
// ("range_kernel", "score_kernel": the tuning keys of the old kernels) and arrays are parsed and skipped, strings may hold
// anything, so a key of the same name inside a nested object or a string value is never picked up.
struct JsonReader {
    const char *p, *end;
    bool ok = true;
    std::vector<std::pair<std::string, double>> top;

    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    bool lit(const char *s) { const size_t n = strlen(s); if ((size_t)(end - p) >= n && !strncmp(p, s, n)) { p += n; return true; } return false; }
    bool string(std::string *out)
    {
        if (p >= end || *p != '"') return ok = false;
        ++p;
        while (p < end && *p != '"') {
            if (*p == '\\') { if (++p >= end) return ok = false; }   // escaped character (\uXXXX digits pass as plain characters)
            if (out) out->push_back(*p);
            ++p;
        }
        if (p >= end) return ok = false;
        ++p;
        return true;
    }
    bool number(double *out)
    {
        char *e = nullptr;
        const double v = strtod(p, &e);
        if (e == p || e > end) return ok = false;
        p = e;
        if (out) *out = v;
        return true;
    }
    bool value(int depth, const std::string *key)
    {
        ws();
        if (p >= end || depth > 64) return ok = false;
        if (*p == '{') {
            ++p; ws();
            if (p < end && *p == '}') { ++p; return true; }
            for (;;) {
                ws();
                std::string k;
                if (!string(&k)) return false;
                ws();
                if (p >= end || *p != ':') return ok = false;
                ++p;
                if (!value(depth + 1, depth == 0 ? &k : nullptr)) return false;
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; return true; }
                return ok = false;
            }
        }
        if (*p == '[') {
            ++p; ws();
            if (p < end && *p == ']') { ++p; return true; }
            for (;;) {
                if (!value(depth + 1, nullptr)) return false;
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; return true; }
                return ok = false;
            }
        }
        if (*p == '"') return string(nullptr);
        if (lit("true") || lit("false") || lit("null")) return true;
        double v;
        if (!number(&v)) return false;
        if (key) top.emplace_back(*key, v);
        return true;
    }
    bool parse(const std::string &txt)
    {
        p = txt.data(); end = p + txt.size();
        if (!value(0, nullptr)) return false;
        ws();
        return ok && p == end;
    }
    bool get(const char *key, double *out) const
    {
        for (const auto &kv : top) if (kv.first == key) { *out = kv.second; return true; }
        return false;
    }
};

// bytes per anchor of batch capacity a driver thread's context costs (two batches in flight): see DESIGN.md section 2
constexpr size_t kDevBytesPerAnchor = 2 * 90, kPinnedBytesPerAnchor = 2 * 28;

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
    JsonReader js;
    if (!txt.empty() && !js.parse(txt)) {
        fprintf(stderr, "[ERROR] mm2gb chaining: gpu config '%s' is not valid JSON\n", path);   // plmem.cu:390-413 exits too
        fatal_exit();
    }
    double v;
    if (js.get("max_total_n", &v) && v >= 1) g_cfg.max_total_n = (size_t)v;
    if (js.get("max_read", &v) && v >= 1) g_cfg.max_read = (int)v;
    if (js.get("min_n", &v) && v >= 0) g_cfg.min_n = (int)v;
    if (js.get("n_gpus", &v) && v >= 0) g_cfg.n_gpus = (int)v;
    if (js.get("sub_batches", &v) && v >= 1 && v <= kMaxSub) g_cfg.sub_batches = (int)v;
    if (js.get("threads_per_gpu", &v) && v >= 1) g_cfg.threads_per_gpu = (int)v;
    if (js.get("gpu_base", &v) && v >= 0) g_cfg.gpu_base = (int)v;
    if (const char *e = getenv("MM2GB_N_GPUS")) g_cfg.n_gpus = atoi(e) > 0 ? atoi(e) : g_cfg.n_gpus;
    if (const char *e = getenv("MM2GB_GPU_BASE")) g_cfg.gpu_base = std::max(0, atoi(e));
    if (const char *e = getenv("MM2GB_SUB_BATCHES")) g_cfg.sub_batches = std::min(kMaxSub, std::max(1, atoi(e)));
    if (const char *e = getenv("MM2GB_THREADS_PER_GPU")) g_cfg.threads_per_gpu = std::max(1, atoi(e));
    if (const char *e = getenv("MM2GB_SUB_MIN")) g_cfg.sub_min = std::max<int64_t>(1, atoll(e));
    if (const char *e = getenv("MM2GB_SUB_MIN_READS")) g_cfg.sub_min_reads = std::max(1, atoi(e));
    if (g_cfg.max_total_n > ((size_t)1 << 31) - 2048) g_cfg.max_total_n = ((size_t)1 << 31) - 2048;
    const int ndev = mm2gb_device_count();
    if (ndev <= 0) { fprintf(stderr, "[ERROR] mm2gb chaining: --gpu-chain needs a CUDA device (no CPU fallback)\n"); fatal_exit(); }
    if (g_cfg.gpu_base >= ndev) g_cfg.gpu_base = 0;
    if (g_cfg.n_gpus <= 0 || g_cfg.n_gpus > ndev - g_cfg.gpu_base) g_cfg.n_gpus = ndev - g_cfg.gpu_base;
    // Every driver thread owns a context with room for two batches, so the batch limit handed to the driver has to fit the
    // memory that is actually there: 80 % of the GPU's free memory shared by threads_per_gpu threads, half of the host's
    // available memory (pinned staging) shared by all of them.  The reference's own gpu_config.json asks for 500 M anchors
    // (sized for its 13 B/anchor buffers and one thread); that is shrunk here with a warning instead of failing in cudaMalloc.
    size_t dev_free = 0, dev_total = 0;
    // while the GPUs are still being brought up in the background the exact free memory would mean waiting for that (a context
    // on the device): 90 % of the total, which needs no context, is close enough for a sanity clamp
    const bool have_mem = g_early_done.load() ? mm2gb_device_memory(g_cfg.gpu_base, &dev_free, &dev_total) == MM2GB_OK
                                              : (mm2gb_device_total_memory(g_cfg.gpu_base, &dev_total) == MM2GB_OK && ((dev_free = (size_t)(0.9 * (double)dev_total)), true));
    if (have_mem && dev_free) {
        const size_t by_dev = (size_t)(0.8 * (double)dev_free) / ((size_t)g_cfg.threads_per_gpu * kDevBytesPerAnchor);
        const long pages = sysconf(_SC_AVPHYS_PAGES), psz = sysconf(_SC_PAGESIZE);
        size_t by_host = SIZE_MAX;
        if (pages > 0 && psz > 0)
            by_host = (size_t)(0.5 * (double)pages * (double)psz) / ((size_t)g_cfg.threads_per_gpu * (size_t)g_cfg.n_gpus * kPinnedBytesPerAnchor);
        const size_t cap = std::max<size_t>(1u << 20, std::min(by_dev, by_host));
        if (g_cfg.max_total_n > cap) {
            fprintf(stderr, "[WARNING] mm2gb chaining: max_total_n %zu shrunk to %zu anchors (%zu B of device and %zu B of pinned memory per "
                            "anchor and driver thread, %d thread(s) per GPU assumed; set \"threads_per_gpu\" in the gpu config)\n",
                    g_cfg.max_total_n, cap, kDevBytesPerAnchor, kPinnedBytesPerAnchor, g_cfg.threads_per_gpu);
            g_cfg.max_total_n = cap;
        }
    }
}

bool same_misc(const Misc_abi &a, const Misc_abi &b) { return memcmp(&a, &b, sizeof(Misc_abi)) == 0; }

void make_ctx(ThreadState &S, size_t cap_anchors, int cap_reads, const Misc_abi &misc)
{
    const double t0 = now_s();
    if (S.ctx) mm2gb_ctx_destroy(S.ctx);
    S.ctx = nullptr;
    {   // a context from the pool, if one that fits is ready or on its way
        std::unique_lock<std::mutex> lk(g_pool_mu);
        const int d = S.device;
        for (;;) {
            auto &v = g_ready[d];
            bool took = false;
            for (size_t i = 0; i < v.size(); ++i)
                if (v[i].cap_anchors >= cap_anchors && v[i].cap_reads >= cap_reads) {
                    S.ctx = v[i].ctx; cap_anchors = v[i].cap_anchors; cap_reads = v[i].cap_reads;
                    v.erase(v.begin() + (long)i);
                    took = true;
                    break;
                }
            if (took || g_planned[d] <= 0 || g_pool_stop) break;
            g_pool_cv.wait(lk);
        }
    }
    if (!S.ctx) join_early();
    if (S.ctx) {
        if (mm2gb_ctx_set_misc(S.ctx, &misc) != MM2GB_OK) die("setting the chaining parameters");
    } else if (mm2gb_ctx_create_ex(&S.ctx, S.device, cap_anchors, cap_reads, 2 * g_cfg.sub_batches, &misc, MM2GB_CTX_NO_FP_STAGING) != MM2GB_OK) {
        die("cannot create the chaining context");
    }
    S.cap_anchors = cap_anchors;
    S.cap_reads = cap_reads;
    S.misc = misc;
    S.has_misc = true;
    S.t_ctx += now_s() - t0;
    ++S.n_ctx;
    VLOG(2, "context on GPU %d: %zu anchors x %d slots created in %.3f s", S.device, cap_anchors, 2 * g_cfg.sub_batches, now_s() - t0);
}

ThreadState &state_of(int tid)
{
    if (tid < 0) { fprintf(stderr, "[ERROR] mm2gb chaining: thread id %d out of range\n", tid); fatal_exit(); }
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_inited) { fprintf(stderr, "[ERROR] mm2gb chaining: chain_stream_gpu before init_stream_gpu\n"); fatal_exit(); }
    if ((size_t)tid >= g_state.size()) g_state.resize((size_t)tid + 1, nullptr);
    if (!g_state[(size_t)tid]) {
        g_state[(size_t)tid] = new ThreadState();
        g_state[(size_t)tid]->device = g_cfg.gpu_base + tid % g_cfg.n_gpus;
    }
    return *g_state[(size_t)tid];
}

// The chaining parameters of a batch.  The boundary chains every read of a batch with ONE parameter set, built for a single
// segment with qlen_sum = 0 (plchain.cu:497-500; the reference guards that with an assert that release builds compile out).
// That is only right when build_misc does not depend on the read: with the short-read / fragment presets (MM_F_SR, or
// max_frag_len > max_gap without max_gap_ref; map.c:398-406) it does, and chaining would silently use the wrong
// max_dist_x / max_dist_y.  Fail loudly instead.
Misc_abi batch_misc(const mm2gb_idx_t *mi, const mm2gb_mapopt_t *opt)
{
    const Misc_abi misc = build_misc(mi, opt, 0, 1);
    const Misc_abi probe = build_misc(mi, opt, (int64_t)1 << 28, 1);
    if (!same_misc(misc, probe)) {
        fprintf(stderr, "[ERROR] mm2gb chaining: with these options the chaining parameters depend on the read length (short-read / "
                        "fragment mode); --gpu-chain supports single-segment long-read chaining only\n");
        fatal_exit();
    }
    return misc;
}

// finish the batch in flight: sub-batch by sub-batch, wait for the device, publish the chains into the arena, run the
// driver's helper.  Chain extraction + compaction (lchain.c:27-111) ran on the device behind the DP kernels; what arrives
// are the chains u[] and the indices of the chain anchors, and compact_a's gather is done here from the read's own array.
void complete_inflight(ThreadState &S, const mm2gb_idx_t *mi, const mm2gb_mapopt_t *opt, const Misc_abi &misc, void *km)
{
    mm2gb_chain_read_t *reads = S.reads;
    const int n_reads = S.n_reads;
    int done_to = 0;
    auto publish = [&](mm2gb_chain_read_t &rd, int32_t n_u, int32_t n_v, const uint64_t *src_u, const int32_t *src_v) {
        if (n_u > 0) {   // arena traffic stays on the calling thread (kalloc is not thread-safe)
            uint64_t *u = (uint64_t *)kmalloc(km, (size_t)n_u * sizeof(uint64_t));
            memcpy(u, src_u, (size_t)n_u * sizeof(uint64_t));
            mm2gb_anchor_t *b = (mm2gb_anchor_t *)kmalloc(km, (size_t)n_v * sizeof(mm2gb_anchor_t));
            mm2gb_gather_anchors(rd.a, src_v, n_v, b);   // lchain.c:100-105
            kfree(km, rd.a);                             // compact_a frees the oversized input array (lchain.c:107-109)
            rd.a = b; rd.u = u; rd.n_u = n_u;
        } else {                                         // lchain.c:212-215 / plchain.cu:135-143
            kfree(km, rd.a);
            rd.a = nullptr; rd.u = nullptr; rd.n_u = 0;
        }
        post_chaining_helper(mi, opt, &rd, misc, km);
    };
    for (const Sub &sb : S.subs) {
        const uint64_t *const *dev_u = nullptr;
        const int32_t *dev_nu = nullptr, *dev_nv = nullptr;
        const int32_t *const *dev_v = nullptr;
        const double t0 = now_s();
        if (mm2gb_wait_chains(S.ctx, sb.slot, &dev_u, &dev_nu, &dev_v, &dev_nv, nullptr, nullptr) != MM2GB_OK) die("waiting for a chaining batch");
        const double t1 = now_s();
        S.t_wait += t1 - t0;
        long long down = 16LL * (sb.r1 - sb.r0 + 1) + 64;   // per-read counts / positions + the batch counters
        for (int r = sb.r0; r < sb.r1; ++r) {
            publish(reads[r], dev_nu[r - sb.r0], dev_nv[r - sb.r0], dev_u[r - sb.r0], dev_v[r - sb.r0]);
            down += 8LL * dev_nu[r - sb.r0] + 4LL * dev_nv[r - sb.r0];
        }
        g_d2h_bytes += down;
        S.t_publish += now_s() - t1;
        done_to = sb.r1;
    }
    for (int r = done_to; r < n_reads; ++r) publish(reads[r], 0, 0, nullptr, nullptr);   // a batch without anchors
    S.busy = false;
    S.reads = nullptr;
    S.n_reads = 0;
    S.subs.clear();
}

} // namespace

extern "C" void init_stream_gpu(size_t *max_total_n, int *max_reads, int *min_n, char gpu_config_file[], Misc_abi misc)
{
    std::lock_guard<std::mutex> lk(g_mu);
    VLOG(2, "init_stream_gpu: start");
    load_config(gpu_config_file);
    VLOG(2, "init_stream_gpu: config loaded, %d GPU(s), batch limit %zu anchors", g_cfg.n_gpus, g_cfg.max_total_n);
    g_misc = misc;
    g_inited = true;
    for (auto &b : g_dev_batches) b = 0;
    if (max_total_n) *max_total_n = g_cfg.max_total_n;
    if (max_reads) *max_reads = g_cfg.max_read;
    if (min_n) *min_n = g_cfg.min_n;
    // the context pool (see above); MM2GB_POOL=0 switches it off
    const bool pool_on = !(getenv("MM2GB_POOL") && atoi(getenv("MM2GB_POOL")) == 0);
    stop_pool();    // (a second init_stream_gpu without free_stream_gpu in between: start over)
    {
        std::lock_guard<std::mutex> pl(g_pool_mu);
        g_pool_stop = false;
        for (int d = 0; d < kMaxDev; ++d) g_planned[d] = 0;
        if (pool_on)
            for (int k = 0; k < g_cfg.n_gpus; ++k) g_planned[g_cfg.gpu_base + k] = std::min(g_cfg.threads_per_gpu, 64);
    }
    if (pool_on) {
        const Config cfg = g_cfg;
        const Misc_abi m = misc;
        g_pool_thread = new std::thread([cfg, m]() {
            join_early();       // the GPUs come up on the early thread; the driver meanwhile reads and seeds its first reads
            // a slot holds one sub-batch: its share of a full batch + the read that overshoots it (see chain_stream_gpu), with the
            // same head room make_ctx gives a context it sizes from a first batch
            const size_t share = cfg.max_total_n / (size_t)cfg.sub_batches + 1;
            size_t cap = share + share / 4 + cfg.max_total_n / 8 + 4096;
            if (cap > ((size_t)1 << 31) - 2048) cap = ((size_t)1 << 31) - 2048;
            const int reads = cfg.max_read + cfg.max_read / 2 + 1;
            for (int round = 0; round < std::min(cfg.threads_per_gpu, 64); ++round)
                for (int k = 0; k < cfg.n_gpus; ++k) {       // round robin over the GPUs: every GPU gets its first context early
                    const int d = cfg.gpu_base + k;
                    {
                        std::lock_guard<std::mutex> pl(g_pool_mu);
                        if (g_pool_stop) { for (int q = 0; q < kMaxDev; ++q) g_planned[q] = 0; g_pool_cv.notify_all(); return; }
                    }
                    mm2gb_ctx_t *ctx = nullptr;
                    const int rc = mm2gb_ctx_create_ex(&ctx, d, cap, reads, 2 * cfg.sub_batches, &m, MM2GB_CTX_NO_FP_STAGING);
                    VLOG(2, "context pool: context %d for GPU %d ready", round, d);
                    std::lock_guard<std::mutex> pl(g_pool_mu);
                    if (rc == MM2GB_OK) g_ready[d].push_back({ctx, cap, reads});
                    else g_planned[d] = 1;      // out of memory or similar: deliver nothing more for this GPU, threads create their own
                    --g_planned[d];
                    g_pool_cv.notify_all();
                    if (rc != MM2GB_OK) { VLOG(1, "context pool: GPU %d: %s", d, mm2gb_last_error()); }
                }
        });
    }
}

extern "C" void chain_stream_gpu(const mm2gb_idx_t *mi, const mm2gb_mapopt_t *opt, mm2gb_chain_read_t **in_arr_, int *n_read_, int thread_id, void *km)
{
    ThreadState &S = state_of(thread_id);
    if (S.n_batches == 0) {
        VLOG(2, "thread %d: first chain_stream_gpu", thread_id);
        // MM2GB_PIN_THREADS=1: the driver thread moves next to its GPU (several GPUs on a multi-socket box: every gather / publish
        // pass stays on the GPU's node).  Off by default: with fewer GPUs than sockets it would take the far cores away from the
        // driver's own stages (seeding, alignment), which run on the same threads.
        if (getenv("MM2GB_PIN_THREADS") && atoi(getenv("MM2GB_PIN_THREADS")) != 0) {
            const int moved = mm2gb_bind_thread_near_device(S.device);
            VLOG(1, "thread %d: %s the CPUs of GPU %d", thread_id, moved ? "moved to" : "no topology, not moved to", S.device);
        }
    }
    const Misc_abi misc = batch_misc(mi, opt);
    mm2gb_chain_read_t *in = in_arr_ ? *in_arr_ : nullptr;
    const int n_in = (n_read_ && in) ? *n_read_ : 0;

    int64_t total = 0, longest = 0;
    S.ptrs.resize((size_t)n_in);
    S.ns.resize((size_t)n_in);
    for (int r = 0; r < n_in; ++r) {
        if (in[r].n_seg > 1) {
            fprintf(stderr, "[ERROR] mm2gb chaining: read %ld has %d segments; --gpu-chain chains single-segment reads only\n", in[r].seq.i, in[r].n_seg);
            fatal_exit();
        }
        S.ptrs[(size_t)r] = in[r].a;
        S.ns[(size_t)r] = in[r].a ? in[r].n : 0;
        total += S.ns[(size_t)r];
        longest = std::max(longest, S.ns[(size_t)r]);
    }
    // sub-batches: enough anchors each to fill the GPU, at most sub_batches of them
    // ... and enough READS each: the chain-extraction kernels run one CTA per read, and a launch of a dozen long reads leaves the
    // GPU idle for the milliseconds its longest read takes (batches of 100-300 kb reads are not cut at all at the default limit)
    const int n_sub = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(g_cfg.sub_batches, total / g_cfg.sub_min), n_in / g_cfg.sub_min_reads));
    const int64_t target = total / n_sub + 1;
    // (re)size the context: first use, a batch beyond the configured limits, or new chaining parameters.  A slot has to hold
    // one sub-batch: its share of the batch plus the read that overshoots it.
    const size_t need_slot = (size_t)(target + longest);
    const bool need_grow = !S.ctx || need_slot > S.cap_anchors || n_in > S.cap_reads;
    const bool new_misc = S.ctx && !same_misc(S.misc, misc);
    mm2gb_chain_read_t *prev = S.busy ? S.reads : nullptr;
    const int n_prev = S.busy ? S.n_reads : 0;
    const bool had_prev = S.busy;
    if ((need_grow || new_misc) && S.busy) complete_inflight(S, mi, opt, S.misc, km); // old buffers / parameters still in use
    if (need_grow) {
        // sized by the batch at hand, not by the configured limit: the driver fills its batches up to max_total_n, so the first
        // batch of a thread already shows how large they are, and a thread that only ever sees small batches stays small
        size_t cap = need_slot + need_slot / 4 + (size_t)longest + 4096;
        if (cap > ((size_t)1 << 31) - 2048) cap = ((size_t)1 << 31) - 2048;
        if (need_slot > cap) { fprintf(stderr, "[ERROR] mm2gb chaining: a sub-batch of %zu anchors cannot be indexed with 31 bits\n", need_slot); fatal_exit(); }
        make_ctx(S, cap, std::max(g_cfg.max_read, n_in + n_in / 2) + 1, misc);
    } else if (new_misc) {
        if (mm2gb_ctx_set_misc(S.ctx, &misc) != MM2GB_OK) die("updating the chaining parameters");
        S.misc = misc;
    }
    // launch the new batch first, so the device works while this thread finishes the previous one
    const int half = S.busy ? 1 - S.half : 0;
    std::vector<Sub> subs;
    const double t_sub0 = now_s();
    if (in && total > 0) {
        ++S.n_batches;
        S.n_anchors += total;
        if (S.device >= 0 && S.device < kMaxDev) ++g_dev_batches[S.device];
        int r0 = 0;
        for (int k = 0; k < n_sub && r0 < n_in; ++k) {
            int r1 = r0;
            int64_t cnt = 0;
            while (r1 < n_in && (k == n_sub - 1 || cnt < target)) cnt += S.ns[(size_t)r1++];   // every sub-batch but the last reaches the target: the last is not above it
            const int slot = half * g_cfg.sub_batches + k;
            if (mm2gb_submit_gather_chains(S.ctx, slot, S.ptrs.data() + r0, S.ns.data() + r0, r1 - r0) != MM2GB_OK) die("launching a chaining batch");
            g_h2d_bytes += mm2gb_last_upload_bytes(S.ctx, slot) + 8LL * (r1 - r0 + 1) + 4LL * (r1 - r0);   // anchors + offsets + class lists
            subs.push_back({slot, r0, r1});
            r0 = r1;
        }
    }
    S.t_submit += now_s() - t_sub0;
    if (S.busy) complete_inflight(S, mi, opt, misc, km);
    if (in) {
        S.busy = true; S.reads = in; S.n_reads = n_in; S.half = half; S.subs.swap(subs);
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
    const Misc_abi misc = batch_misc(mi, opt);
    complete_inflight(S, mi, opt, misc, km);
    if (reads_) *reads_ = prev;
    if (n_read_) *n_read_ = n_prev;
}

extern "C" void free_stream_gpu(int n_threads)
{
    (void)n_threads;
    std::lock_guard<std::mutex> lk(g_mu);
    const char *verbose = getenv("MM2GB_VERBOSE");
    int tid = 0;
    VLOG(2, "free_stream_gpu: start");
    join_early();
    stop_pool();
    for (ThreadState *&st : g_state) {
        ++tid;
        if (!st) continue;
        const double t0 = now_s();
        if (st->ctx) mm2gb_ctx_destroy(st->ctx);
        if (verbose && atoi(verbose))
            fprintf(stderr, "[mm2gb] thread %d (GPU %d): %lld batches, %lld anchors; context x%lld %.3f s (slot capacity %zu anchors), gather+submit %.3f s, "
                            "wait %.3f s, publish %.3f s, teardown %.3f s\n",
                    tid - 1, st->device, st->n_batches, st->n_anchors, st->n_ctx, st->t_ctx, st->cap_anchors, st->t_submit, st->t_wait, st->t_publish,
                    now_s() - t0);
        delete st;
        st = nullptr;
    }
    g_state.clear();
    g_inited = false;
    VLOG(2, "free_stream_gpu: done");
}

// test hook: parse a gpu config text the way init_stream_gpu does; returns 1 and the value if `key` is a numeric member of the
// top-level object, 0 if it is not there, -1 if the text is not valid JSON
extern "C" int mm2gb_dropin_parse_config_key(const char *text, const char *key, double *out)
{
    JsonReader js;
    if (!js.parse(text ? text : "")) return -1;
    double v;
    if (!js.get(key, &v)) return 0;
    if (out) *out = v;
    return 1;
}

// bytes moved host -> device (out[0]) and device -> host (out[1]) by the boundary since the last reset
extern "C" void mm2gb_dropin_traffic(long long out[2], int reset)
{
    if (out) { out[0] = g_h2d_bytes.load(); out[1] = g_d2h_bytes.load(); }
    if (reset) { g_h2d_bytes = 0; g_d2h_bytes = 0; }
}

// batches launched on GPU `device` since init_stream_gpu (thread_id t drives GPU gpu_base + t % n_gpus)
extern "C" long long mm2gb_dropin_device_batches(int device)
{
    return device >= 0 && device < kMaxDev ? g_dev_batches[device].load() : 0;
}
