// chain_core.cu -- host side of libmm2gb_chain.so: contexts, slots (stream + pinned staging + device buffers),
// kernel launches and the C ABI of include/mm2gb_chain.h.
//
// Replaces, B200-first, what the reference spreads over gpu/plmem.cu (buffer sizing :453-540, pinned/device
// allocation :12-143, 7 H2D + 3 memset per micro-batch :200-236, D2H :324-359) and the launch half of
// gpu/plchain.cu:292-464.  Differences that matter:
//   * the raw 16-byte mm128_t array is uploaded as is -- no AoS->SoA repack on the host (plmem.cu:154-198 is gone);
//   * one flat batch, no micro-batches, no host sync between "short" and "long" phases (plchain.cu:426-452 is gone);
//   * n_slots independent slots per context so upload, kernels and download of consecutive batches overlap.
#include "chain_kernels.cuh"
#include "../../include/mm2gb_chain.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

using namespace mm2gb;

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) return fail(MM2GB_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

extern "C" const char *mm2gb_last_error(void) { return g_err; }

extern "C" int mm2gb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

namespace {

enum { T_RANGE = 0, T_UNITS, T_SCORE, T_LONG, T_H2D, T_D2H };

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    // device
    uint4 *d_a = nullptr;
    long long *d_off = nullptr;
    int *d_st = nullptr, *d_f = nullptr, *d_p = nullptr;
    unsigned *d_selmask = nullptr, *d_clipmask = nullptr;
    int *d_block_cnt = nullptr, *d_block_base = nullptr; unsigned long long *d_block_pairs = nullptr; int *d_unit_start = nullptr, *d_unit_rbase = nullptr, *d_big_order = nullptr;
    int big_cap = 0;
    Counters *d_ctr = nullptr;
    // pinned host
    mm2gb_anchor_t *h_a = nullptr;
    long long *h_off = nullptr;
    int *h_f = nullptr, *h_p = nullptr;
    Counters *h_ctr = nullptr;
    // state
    bool busy = false;
    int n_reads = 0;
    long long n_total = 0;
    // where results of a synchronous chunk go (mm2gb_chain_dp_host)
    int *user_f = nullptr, *user_p = nullptr;
    bool direct_out = false;
};

} // namespace

struct mm2gb_ctx {
    int device = 0, n_sm = 0;
    size_t max_anchors = 0;
    int max_reads = 0, n_slots = 0;
    mm2gb_misc_t misc;
    DevParams prm;
    bool fast = false;
    unsigned char *d_lut = nullptr;
    int ring = 512;
    int score_blocks = 0;
    size_t score_smem = 0;
    int long_min = INT32_MAX;
    Slot slot[4];
    // profiling (slot 0 only)
    bool profile = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev_pending;
    float prof_ms[MM2GB_NTIMERS] = {0};
    int64_t prof_n[MM2GB_NTIMERS] = {0};
    Counters last_dev_ctr;
};

// ---- parameters ---------------------------------------------------------------------------------------------------

// mmpriv.h:118-126 on the host, for the penalty table; this TU is compiled with -ffp-contract=off (nvcc -Xcompiler)
static float host_log2(float x)
{
    union { float f; uint32_t i; } z;
    z.f = x;
    float r = (float)((int)((z.i >> 23) & 255) - 128);
    z.i &= ~(255U << 23);
    z.i += 127U << 23;
    volatile float t1 = -0.34484843f * z.f;
    volatile float t2 = t1 + 2.02466578f;
    volatile float t3 = t2 * z.f;
    volatile float t4 = t3 - 0.67487759f;
    return r + t4;
}

static int setup_params(mm2gb_ctx *c, const mm2gb_misc_t *m)
{
    if (m->bw < 0 || m->max_dist_x < 0 || m->max_dist_y < 0 || m->max_iter < 0)
        return fail(MM2GB_EARG, "negative chaining parameter");
    c->misc = *m;
    DevParams &P = c->prm;
    P.max_iter = m->max_iter;
    P.bw = m->bw;
    P.is_cdna = m->is_cdna;
    P.n_seg = m->n_seg;
    P.max_dist_x = m->max_dist_x < m->bw ? m->bw : m->max_dist_x;                       // lchain.c:160
    P.max_dist_y = (m->max_dist_y < m->bw && !m->is_cdna) ? m->bw : m->max_dist_y;      // lchain.c:161
    P.maxd_q = std::min(P.max_dist_x, P.max_dist_y);
    P.pen_gap = m->chn_pen_gap;
    P.pen_skip = m->chn_pen_skip;
    c->fast = !m->is_cdna && m->n_seg <= 1 && m->chn_pen_skip == 0.0f && m->bw <= kLutMax;
    if (c->fast) {
        std::vector<unsigned char> lut((size_t)2 * m->bw + 1); // symmetric: entry k = pen(|k - bw|)
        for (int dd = 0; dd <= m->bw && c->fast; ++dd) { // lchain.c:128-135 with dg * 0.0f == 0
            volatile float lin = m->chn_pen_gap * (float)dd;
            float lg = dd >= 1 ? host_log2((float)(dd + 1)) : 0.0f;
            volatile float half = .5f * lg;
            volatile float sum = lin + half;
            const int pen = (int)sum;
            if (pen < 0 || pen > 255) c->fast = false; // does not fit a byte table: use the arithmetic path
            lut[(size_t)(m->bw + dd)] = lut[(size_t)(m->bw - dd)] = (unsigned char)pen;
        }
        if (c->fast) CK(cudaMemcpy(c->d_lut, lut.data(), lut.size(), cudaMemcpyHostToDevice));
    }
    P.lut_n = c->fast ? 2 * m->bw + 1 : 0;
    return MM2GB_OK;
}

// ---- kernel plumbing ------------------------------------------------------------------------------------------------

template <int R, bool FAST>
static int config_score(mm2gb_ctx *c, size_t smem)
{
    CK(cudaFuncSetAttribute(k_score_units<R, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_score_units<R, FAST>, kScoreWarps * 32, smem));
    if (nb < 1) return fail(MM2GB_ECUDA, "score kernel does not fit on an SM (smem %zu)", smem);
    c->score_blocks = std::max(c->score_blocks, nb * c->n_sm);
    return MM2GB_OK;
}

template <int R>
static int config_ring(mm2gb_ctx *c)
{
    const size_t smem = (size_t)kScoreWarps * R * sizeof(Rec);
    c->score_smem = smem;
    c->score_blocks = 0;
    int rc = config_score<R, true>(c, smem);
    if (rc) return rc;
    return config_score<R, false>(c, smem);
}

template <int R, bool FAST>
static void launch_score(mm2gb_ctx *c, cudaStream_t s, const uint4 *a, const int *st, const int *us, const int *ur,
                         const unsigned *clip, int *f, int *p, const int *big, int big_cap, Counters *ctr, int run_mode)
{
    const size_t smem = (size_t)kScoreWarps * R * sizeof(Rec);
    k_score_units<R, FAST><<<c->score_blocks, kScoreWarps * 32, smem, s>>>(a, st, us, ur, clip, f, p, big, big_cap, ctr, c->prm,
                                                                          c->d_lut, run_mode, c->long_min);
}

template <bool FAST>
static void launch_score_ring(mm2gb_ctx *c, cudaStream_t s, const uint4 *a, const int *st, const int *us, const int *ur,
                              const unsigned *clip, int *f, int *p, const int *big, int big_cap, Counters *ctr, int run_mode)
{
    switch (c->ring) {
    case 256: launch_score<256, FAST>(c, s, a, st, us, ur, clip, f, p, big, big_cap, ctr, run_mode); break;
    case 1024: launch_score<1024, FAST>(c, s, a, st, us, ur, clip, f, p, big, big_cap, ctr, run_mode); break;
    default: launch_score<512, FAST>(c, s, a, st, us, ur, clip, f, p, big, big_cap, ctr, run_mode); break;
    }
}

struct ProfScope {
    mm2gb_ctx *c; int id; cudaStream_t s; cudaEvent_t e0 = nullptr, e1 = nullptr; bool on;
    ProfScope(mm2gb_ctx *c_, int id_, cudaStream_t s_, bool on_) : c(c_), id(id_), s(s_), on(on_)
    {
        if (!on) return;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, s);
    }
    ~ProfScope()
    {
        if (!on) return;
        cudaEventRecord(e1, s);
        c->ev_pending.push_back({id, {e0, e1}});
    }
};

static void prof_collect(mm2gb_ctx *c)
{
    for (auto &pe : c->ev_pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, pe.second.first, pe.second.second) == cudaSuccess) {
            c->prof_ms[pe.first] += ms;
            c->prof_n[pe.first] += 1;
        }
        cudaEventDestroy(pe.second.first);
        cudaEventDestroy(pe.second.second);
    }
    c->ev_pending.clear();
}

// range -> scan -> units -> score on `s`, all device pointers; scratch from `sl`
static int enqueue_kernels(mm2gb_ctx *c, Slot &sl, cudaStream_t s, const uint4 *d_a, const long long *d_off, int n_reads,
                           long long n_total, int *d_f, int *d_p, bool prof)
{
    CK(cudaMemsetAsync(sl.d_ctr, 0, sizeof(Counters), s));
    if (n_total == 0) return MM2GB_OK;
    const int n = (int)n_total;
    const int n_blocks = (n + kRangeThreads - 1) / kRangeThreads;
    const int n_groups = (n + 31) / 32;
    {
        ProfScope ps(c, T_RANGE, s, prof);
        k_block_reads<<<(n_blocks + 255) / 256, 256, 0, s>>>(d_off, n_reads, n_blocks, sl.d_block_base);   // d_block_base doubles as block_read until k_scan
        k_range<<<n_blocks, kRangeThreads, 0, s>>>(reinterpret_cast<const ulonglong2 *>(d_a), d_off, sl.d_block_base, n, c->prm, sl.d_st,
                                                  sl.d_selmask, sl.d_clipmask, sl.d_block_cnt, sl.d_block_pairs, sl.d_ctr);
    }
    {
        ProfScope ps(c, T_UNITS, s, prof);
        k_scan<<<1, 1024, 0, s>>>(sl.d_block_cnt, sl.d_block_pairs, n_blocks, sl.d_block_base, sl.d_ctr);
        k_units<<<(n_groups + 255) / 256, 256, 0, s>>>(sl.d_selmask, sl.d_block_base, d_off, n_reads, n, n_groups, sl.d_unit_start,
                                                      sl.d_unit_rbase, sl.d_ctr);
        k_order<<<(n_groups + n_reads + 256) / 256, 256, 0, s>>>(sl.d_unit_start, sl.d_big_order, sl.big_cap, sl.d_ctr);
    }
    {
        ProfScope ps(c, T_SCORE, s, prof);
        if (c->fast) {
            launch_score_ring<true>(c, s, d_a, sl.d_st, sl.d_unit_start, sl.d_unit_rbase, sl.d_clipmask, d_f, d_p, sl.d_big_order, sl.big_cap, sl.d_ctr, 1);
            launch_score_ring<false>(c, s, d_a, sl.d_st, sl.d_unit_start, sl.d_unit_rbase, sl.d_clipmask, d_f, d_p, sl.d_big_order, sl.big_cap, sl.d_ctr, 2);
        } else {
            launch_score_ring<false>(c, s, d_a, sl.d_st, sl.d_unit_start, sl.d_unit_rbase, sl.d_clipmask, d_f, d_p, sl.d_big_order, sl.big_cap, sl.d_ctr, 0);
        }
    }
    CK(cudaGetLastError());
    return MM2GB_OK;
}

static void fill_stats(const mm2gb_ctx *c, const Counters &k, long long n_total, mm2gb_stats_t *st)
{
    if (!st) return;
    st->n_anchors = n_total;
    st->n_pairs = (int64_t)k.n_pairs;
    st->n_units = k.n_units;
    st->n_units_exact = k.n_exact;
    st->n_long = k.n_long;
    st->general_path = (!c->fast || k.multi_sid) ? 1 : 0;
}

static void free_slot(Slot &s)
{
    if (s.stream) cudaStreamSynchronize(s.stream);
    cudaFree(s.d_a); cudaFree(s.d_off); cudaFree(s.d_st); cudaFree(s.d_f); cudaFree(s.d_p);
    cudaFree(s.d_selmask); cudaFree(s.d_clipmask); cudaFree(s.d_block_cnt); cudaFree(s.d_block_base); cudaFree(s.d_block_pairs);
    cudaFree(s.d_unit_start); cudaFree(s.d_unit_rbase); cudaFree(s.d_big_order); cudaFree(s.d_ctr);
    cudaFreeHost(s.h_a); cudaFreeHost(s.h_off); cudaFreeHost(s.h_f); cudaFreeHost(s.h_p); cudaFreeHost(s.h_ctr);
    if (s.done) cudaEventDestroy(s.done);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Slot();
}

// ---- C ABI -------------------------------------------------------------------------------------------------------------

extern "C" int mm2gb_ctx_create(mm2gb_ctx_t **out, int device, size_t max_anchors, int max_reads, int n_slots, const mm2gb_misc_t *misc)
{
    if (!out || !misc) return fail(MM2GB_EARG, "null argument");
    *out = nullptr;
    if (n_slots < 1 || n_slots > 4) return fail(MM2GB_EARG, "n_slots must be 1..4");
    if (max_anchors == 0 || max_anchors > (size_t)INT32_MAX - 1024) return fail(MM2GB_EARG, "max_anchors must be in (0, 2^31)");
    if (max_reads < 1) return fail(MM2GB_EARG, "max_reads must be positive");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(MM2GB_EARG, "no CUDA device %d (have %d)", device, ndev);
    CK(cudaSetDevice(device));
    mm2gb_ctx *c = new mm2gb_ctx();
    c->device = device;
    c->max_anchors = max_anchors;
    c->max_reads = max_reads;
    c->n_slots = n_slots;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->n_sm = prop.multiProcessorCount;
    if (const char *e = getenv("MM2GB_RING")) {
        int r = atoi(e);
        if (r == 256 || r == 512 || r == 1024) c->ring = r;
    }
    int rc = MM2GB_OK;
#define CKC(call)                                                                                                  \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) {                                                                                   \
            rc = fail(MM2GB_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));           \
            goto bad;                                                                                              \
        }                                                                                                          \
    } while (0)
    {
        CKC(cudaMalloc(&c->d_lut, (size_t)2 * kLutMax + 16));
        rc = setup_params(c, misc);
        if (rc) goto bad;
        rc = c->ring == 256 ? config_ring<256>(c) : c->ring == 1024 ? config_ring<1024>(c) : config_ring<512>(c);
        if (rc) goto bad;
        const size_t n = max_anchors, n_groups = (n + 31) / 32, n_blocks = (n + kRangeThreads - 1) / kRangeThreads;
        const size_t n_units_cap = n_groups + (size_t)max_reads + 2;
        for (int i = 0; i < n_slots; ++i) {
            Slot &s = c->slot[i];
            CKC(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
            CKC(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
            CKC(cudaMalloc(&s.d_a, n * sizeof(uint4)));
            CKC(cudaMalloc(&s.d_off, ((size_t)max_reads + 1) * sizeof(long long)));
            CKC(cudaMalloc(&s.d_st, n * sizeof(int)));
            CKC(cudaMalloc(&s.d_f, n * sizeof(int)));
            CKC(cudaMalloc(&s.d_p, n * sizeof(int)));
            CKC(cudaMalloc(&s.d_selmask, n_groups * sizeof(unsigned)));
            CKC(cudaMalloc(&s.d_clipmask, n_groups * sizeof(unsigned)));
            CKC(cudaMalloc(&s.d_block_cnt, n_blocks * sizeof(int)));
            CKC(cudaMalloc(&s.d_block_base, n_blocks * sizeof(int)));
            CKC(cudaMalloc(&s.d_block_pairs, n_blocks * sizeof(unsigned long long)));
            CKC(cudaMalloc(&s.d_unit_start, n_units_cap * sizeof(int)));
            CKC(cudaMalloc(&s.d_unit_rbase, n_units_cap * sizeof(int)));
            s.big_cap = (int)(n / kBigMin) + 2;
            CKC(cudaMalloc(&s.d_big_order, (size_t)4 * s.big_cap * sizeof(int)));
            CKC(cudaMalloc(&s.d_ctr, sizeof(Counters)));
            CKC(cudaMallocHost(&s.h_a, n * sizeof(mm2gb_anchor_t)));
            CKC(cudaMallocHost(&s.h_off, ((size_t)max_reads + 1) * sizeof(long long)));
            CKC(cudaMallocHost(&s.h_f, n * sizeof(int)));
            CKC(cudaMallocHost(&s.h_p, n * sizeof(int)));
            CKC(cudaMallocHost(&s.h_ctr, sizeof(Counters)));
        }
    }
    *out = c;
    return MM2GB_OK;
bad:
    mm2gb_ctx_destroy(c);
    return rc;
#undef CKC
}

extern "C" void mm2gb_ctx_destroy(mm2gb_ctx_t *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < 4; ++i) free_slot(c->slot[i]);
    prof_collect(c);
    cudaFree(c->d_lut);
    delete c;
}

extern "C" int mm2gb_ctx_set_misc(mm2gb_ctx_t *c, const mm2gb_misc_t *misc)
{
    if (!c || !misc) return fail(MM2GB_EARG, "null argument");
    CK(cudaSetDevice(c->device));
    for (int i = 0; i < c->n_slots; ++i) CK(cudaStreamSynchronize(c->slot[i].stream));
    int rc = setup_params(c, misc);
    if (rc) return rc;
    return c->ring == 256 ? config_ring<256>(c) : c->ring == 1024 ? config_ring<1024>(c) : config_ring<512>(c);
}

static bool is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// enqueue one batch whose anchors already sit in host memory `src` (pinned: direct DMA; else staged through h_a)
static int submit_impl(mm2gb_ctx *c, int si, const mm2gb_anchor_t *src, bool src_pinned, const long long *off_rel, int n_reads,
                       long long n_total, int *dst_f, int *dst_p, bool dst_pinned)
{
    Slot &s = c->slot[si];
    if (s.busy) return fail(MM2GB_ESTATE, "slot %d is busy", si);
    if ((size_t)n_total > c->max_anchors) return fail(MM2GB_ECAP, "batch of %lld anchors exceeds capacity %zu", n_total, c->max_anchors);
    if (n_reads > c->max_reads) return fail(MM2GB_ECAP, "batch of %d reads exceeds capacity %d", n_reads, c->max_reads);
    CK(cudaSetDevice(c->device));
    const bool prof = c->profile && si == 0;
    memcpy(s.h_off, off_rel, ((size_t)n_reads + 1) * sizeof(long long));
    const mm2gb_anchor_t *h_src = src;
    if (!src_pinned && n_total) { memcpy(s.h_a, src, (size_t)n_total * sizeof(mm2gb_anchor_t)); h_src = s.h_a; }
    {
        ProfScope ps(c, T_H2D, s.stream, prof);
        CK(cudaMemcpyAsync(s.d_off, s.h_off, ((size_t)n_reads + 1) * sizeof(long long), cudaMemcpyHostToDevice, s.stream));
        if (n_total) CK(cudaMemcpyAsync(s.d_a, h_src, (size_t)n_total * sizeof(uint4), cudaMemcpyHostToDevice, s.stream));
    }
    int rc = enqueue_kernels(c, s, s.stream, s.d_a, s.d_off, n_reads, n_total, s.d_f, s.d_p, prof);
    if (rc) return rc;
    s.direct_out = dst_pinned && dst_f && dst_p;
    s.user_f = dst_f; s.user_p = dst_p;
    {
        ProfScope ps(c, T_D2H, s.stream, prof);
        if (n_total) {
            CK(cudaMemcpyAsync(s.direct_out ? dst_f : s.h_f, s.d_f, (size_t)n_total * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
            CK(cudaMemcpyAsync(s.direct_out ? dst_p : s.h_p, s.d_p, (size_t)n_total * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        }
        CK(cudaMemcpyAsync(s.h_ctr, s.d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, s.stream));
    }
    CK(cudaEventRecord(s.done, s.stream));
    s.busy = true;
    s.n_reads = n_reads;
    s.n_total = n_total;
    return MM2GB_OK;
}

static int wait_impl(mm2gb_ctx *c, int si)
{
    Slot &s = c->slot[si];
    if (!s.busy) return fail(MM2GB_ESTATE, "slot %d is idle", si);
    CK(cudaEventSynchronize(s.done));
    s.busy = false;
    if (c->profile && si == 0) prof_collect(c);
    if (!s.direct_out && s.user_f && s.user_p && s.n_total) {
        memcpy(s.user_f, s.h_f, (size_t)s.n_total * sizeof(int));
        memcpy(s.user_p, s.h_p, (size_t)s.n_total * sizeof(int));
    }
    return MM2GB_OK;
}

extern "C" int mm2gb_submit(mm2gb_ctx_t *c, int slot, const mm2gb_anchor_t *a, const int64_t *off, int n_reads)
{
    if (!c || slot < 0 || slot >= c->n_slots || n_reads < 0 || !off) return fail(MM2GB_EARG, "bad argument");
    if (off[0] != 0) return fail(MM2GB_EARG, "off[0] must be 0");
    return submit_impl(c, slot, a, is_pinned(a), (const long long *)off, n_reads, off[n_reads], nullptr, nullptr, false);
}

extern "C" int mm2gb_submit_gather(mm2gb_ctx_t *c, int slot, const mm2gb_anchor_t *const *read_a, const int64_t *read_n, int n_reads)
{
    if (!c || slot < 0 || slot >= c->n_slots || n_reads < 0) return fail(MM2GB_EARG, "bad argument");
    Slot &s = c->slot[slot];
    if (s.busy) return fail(MM2GB_ESTATE, "slot %d is busy", slot);
    if (n_reads > c->max_reads) return fail(MM2GB_ECAP, "batch of %d reads exceeds capacity %d", n_reads, c->max_reads);
    std::vector<long long> off((size_t)n_reads + 1);
    long long tot = 0;
    for (int r = 0; r < n_reads; ++r) { off[(size_t)r] = tot; tot += read_n[r] > 0 ? read_n[r] : 0; }
    off[(size_t)n_reads] = tot;
    if ((size_t)tot > c->max_anchors) return fail(MM2GB_ECAP, "batch of %lld anchors exceeds capacity %zu", tot, c->max_anchors);
    for (int r = 0; r < n_reads; ++r)
        if (read_n[r] > 0) memcpy(s.h_a + off[(size_t)r], read_a[r], (size_t)read_n[r] * sizeof(mm2gb_anchor_t));
    return submit_impl(c, slot, s.h_a, true, off.data(), n_reads, tot, nullptr, nullptr, false);
}

extern "C" int mm2gb_wait(mm2gb_ctx_t *c, int slot, const int32_t **f, const int32_t **p, const int64_t **off, mm2gb_stats_t *stats)
{
    if (!c || slot < 0 || slot >= c->n_slots) return fail(MM2GB_EARG, "bad argument");
    int rc = wait_impl(c, slot);
    if (rc) return rc;
    Slot &s = c->slot[slot];
    if (f) *f = s.h_f;
    if (p) *p = s.h_p;
    if (off) *off = (const int64_t *)s.h_off;
    fill_stats(c, *s.h_ctr, s.n_total, stats);
    return MM2GB_OK;
}

extern "C" int mm2gb_slot_busy(mm2gb_ctx_t *c, int slot)
{
    if (!c || slot < 0 || slot >= c->n_slots) return 0;
    return c->slot[slot].busy ? 1 : 0;
}

// Shared driver of the host-buffer entry points: split the reads into chunks, run them round-robin through the slots
// (upload / kernels / download of consecutive chunks overlap) and call on_done(r0, r1) as each chunk's f/p land.
template <class Done>
static int run_chunked(mm2gb_ctx *c, const mm2gb_anchor_t *a, const int64_t *off, int n_reads, int32_t *f, int32_t *p,
                       mm2gb_stats_t *stats, Done on_done)
{
    for (int i = 0; i < c->n_slots; ++i)
        if (c->slot[i].busy) return fail(MM2GB_ESTATE, "slot %d is busy", i);
    const bool in_pinned = a && is_pinned(a), out_pinned = is_pinned(f) && is_pinned(p);
    const long long total = off[n_reads] - off[0];
    long long target = std::max<long long>(1 << 20, total / (4LL * c->n_slots) + 1);
    target = std::min<long long>(target, (long long)c->max_anchors);
    if (const char *e = getenv("MM2GB_CHUNK")) target = std::min<long long>(std::max(1LL, atoll(e)), (long long)c->max_anchors);
    mm2gb_stats_t acc;
    memset(&acc, 0, sizeof(acc));
    std::vector<long long> rel;
    int slot_r0[4] = {0, 0, 0, 0}, slot_r1[4] = {0, 0, 0, 0};
    int r0 = 0, chunk = 0, rc = MM2GB_OK;
    auto reap = [&](int si) -> int {
        int rc2 = wait_impl(c, si);
        if (rc2) return rc2;
        mm2gb_stats_t st;
        fill_stats(c, *c->slot[si].h_ctr, c->slot[si].n_total, &st);
        acc.n_anchors += st.n_anchors; acc.n_pairs += st.n_pairs; acc.n_units += st.n_units;
        acc.n_units_exact += st.n_units_exact; acc.n_long += st.n_long; acc.general_path |= st.general_path;
        on_done(slot_r0[si], slot_r1[si]);
        return MM2GB_OK;
    };
    while (r0 < n_reads) {
        int r1 = r0;
        long long cnt = 0;
        while (r1 < n_reads && r1 - r0 < c->max_reads) {
            const long long nr = off[r1 + 1] - off[r1];
            if (nr > (long long)c->max_anchors) return fail(MM2GB_ECAP, "read %d has %lld anchors, capacity is %zu", r1, nr, c->max_anchors);
            if (cnt && cnt + nr > target) break;
            cnt += nr; ++r1;
        }
        const int si = chunk % c->n_slots;
        if (c->slot[si].busy && (rc = reap(si))) return rc;
        rel.resize((size_t)(r1 - r0) + 1);
        for (int r = r0; r <= r1; ++r) rel[(size_t)(r - r0)] = off[r] - off[r0];
        rc = submit_impl(c, si, a + off[r0], in_pinned, rel.data(), r1 - r0, cnt, f + off[r0], p + off[r0], out_pinned);
        if (rc) return rc;
        slot_r0[si] = r0; slot_r1[si] = r1;
        r0 = r1; ++chunk;
    }
    // drain in submission order
    for (int k = 0; k < c->n_slots; ++k) {
        const int si = (chunk + k) % c->n_slots;
        if (c->slot[si].busy && (rc = reap(si))) return rc;
    }
    if (stats) *stats = acc;
    return MM2GB_OK;
}

extern "C" int mm2gb_chain_dp_host(mm2gb_ctx_t *c, const mm2gb_anchor_t *a, const int64_t *off, int n_reads, int32_t *f, int32_t *p,
                                   mm2gb_stats_t *stats)
{
    if (!c || !off || n_reads < 0 || (!a && off[n_reads] > 0) || ((!f || !p) && off[n_reads] > 0)) return fail(MM2GB_EARG, "bad argument");
    if (off[0] != 0) return fail(MM2GB_EARG, "off[0] must be 0");
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n_reads == 0) return MM2GB_OK;
    return run_chunked(c, a, off, n_reads, f, p, stats, [](int, int) {});
}

// Whole mg_lchain_dp (lchain.c:148-217) for a batch: device DP, then the host stage (backtracking + compaction) on
// `n_threads` worker threads that start on a chunk's reads as soon as its f/p have landed, while later chunks are
// still on the GPU.  Outputs per read r: u[off[r] .. off[r]+n_u[r]), b[off[r] .. off[r]+n_b[r]).
extern "C" int mm2gb_chain_host(mm2gb_ctx_t *c, const mm2gb_anchor_t *a, const int64_t *off, int n_reads, int32_t *f, int32_t *p,
                                uint64_t *u, int32_t *n_u, mm2gb_anchor_t *b, int64_t *n_b, int n_threads, mm2gb_stats_t *stats)
{
    if (!c || !off || n_reads < 0 || (!a && off[n_reads] > 0) || !n_u || !n_b) return fail(MM2GB_EARG, "bad argument");
    if (off[n_reads] > 0 && (!f || !p || !u || !b)) return fail(MM2GB_EARG, "bad argument");
    if (off[0] != 0) return fail(MM2GB_EARG, "off[0] must be 0");
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n_reads == 0) return MM2GB_OK;
    if (n_threads < 1) n_threads = 1;
    const mm2gb_misc_t m = c->misc;
    const int32_t max_drop = m.is_cdna ? INT32_MAX : m.bw; // lchain.c:151,162
    std::mutex mu;
    std::condition_variable cv;
    int ready = 0;          // reads [0, ready) have f/p on the host
    bool abort_all = false;
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&]() {
            for (;;) {
                const int r = next.fetch_add(1);
                if (r >= n_reads) return;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return ready > r || abort_all; });
                    if (abort_all) return;
                }
                const int64_t s = off[r], n = off[r + 1] - s;
                int64_t nb = 0;
                n_u[r] = mm2gb_backtrack(n, f + s, p + s, a + s, m.min_cnt, m.min_score, max_drop, u + s, b + s, &nb);
                n_b[r] = nb;
            }
        });
    int rc = run_chunked(c, a, off, n_reads, f, p, stats, [&](int, int r1) {
        { std::lock_guard<std::mutex> lk(mu); ready = r1; }
        cv.notify_all();
    });
    {
        std::lock_guard<std::mutex> lk(mu);
        if (rc) abort_all = true; else ready = n_reads;
    }
    cv.notify_all();
    for (auto &t : pool) t.join();
    return rc;
}

// The host stage alone for a batch whose f/p are already in host memory (what mm2gb_chain_host runs behind the device).
extern "C" int mm2gb_backtrack_batch(const mm2gb_misc_t *m, const mm2gb_anchor_t *a, const int64_t *off, int n_reads, const int32_t *f,
                                     const int32_t *p, uint64_t *u, int32_t *n_u, mm2gb_anchor_t *b, int64_t *n_b, int n_threads)
{
    if (!m || !off || n_reads < 0 || !n_u || !n_b) return fail(MM2GB_EARG, "bad argument");
    if (n_reads && off[n_reads] > 0 && (!a || !f || !p || !u || !b)) return fail(MM2GB_EARG, "bad argument");
    if (n_threads < 1) n_threads = 1;
    const int32_t max_drop = m->is_cdna ? INT32_MAX : m->bw; // lchain.c:151,162
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            const int r = next.fetch_add(1);
            if (r >= n_reads) return;
            const int64_t s = off[r], n = off[r + 1] - s;
            int64_t nb = 0;
            n_u[r] = mm2gb_backtrack(n, f + s, p + s, a + s, m->min_cnt, m->min_score, max_drop, u + s, b + s, &nb);
            n_b[r] = nb;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    return MM2GB_OK;
}

extern "C" int mm2gb_chain_dp_device(mm2gb_ctx_t *c, const void *d_a, const void *d_off, int n_reads, int64_t n_total, void *d_f, void *d_p)
{
    if (!c || n_reads < 0 || n_total < 0 || !d_off) return fail(MM2GB_EARG, "bad argument");
    if ((size_t)n_total > c->max_anchors) return fail(MM2GB_ECAP, "batch of %lld anchors exceeds capacity %zu", (long long)n_total, c->max_anchors);
    if (n_reads > c->max_reads) return fail(MM2GB_ECAP, "batch of %d reads exceeds capacity %d", n_reads, c->max_reads);
    CK(cudaSetDevice(c->device));
    Slot &s = c->slot[0];
    int rc = enqueue_kernels(c, s, s.stream, (const uint4 *)d_a, (const long long *)d_off, n_reads, n_total, (int *)d_f, (int *)d_p, c->profile);
    if (rc) return rc;
    CK(cudaMemcpyAsync(s.h_ctr, s.d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, s.stream));
    s.n_total = n_total;
    return MM2GB_OK;
}

extern "C" int mm2gb_sync(mm2gb_ctx_t *c, int slot)
{
    if (!c || slot < 0 || slot >= c->n_slots) return fail(MM2GB_EARG, "bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->slot[slot].stream));
    if (c->profile && slot == 0) prof_collect(c);
    return MM2GB_OK;
}

extern "C" void *mm2gb_stream(mm2gb_ctx_t *c, int slot)
{
    if (!c || slot < 0 || slot >= c->n_slots) return nullptr;
    return (void *)c->slot[slot].stream;
}

extern "C" int mm2gb_device_stats(mm2gb_ctx_t *c, mm2gb_stats_t *stats)
{
    if (!c || !stats) return fail(MM2GB_EARG, "bad argument");
    fill_stats(c, *c->slot[0].h_ctr, c->slot[0].n_total, stats);
    return MM2GB_OK;
}

extern "C" int mm2gb_profile(mm2gb_ctx_t *c, int enable)
{
    if (!c) return fail(MM2GB_EARG, "bad argument");
    c->profile = enable != 0;
    if (enable) { memset(c->prof_ms, 0, sizeof(c->prof_ms)); memset(c->prof_n, 0, sizeof(c->prof_n)); }
    return MM2GB_OK;
}

extern "C" int mm2gb_profile_read(mm2gb_ctx_t *c, float ms[MM2GB_NTIMERS], int64_t launches[MM2GB_NTIMERS])
{
    if (!c) return fail(MM2GB_EARG, "bad argument");
    for (int i = 0; i < MM2GB_NTIMERS; ++i) { if (ms) ms[i] = c->prof_ms[i]; if (launches) launches[i] = c->prof_n[i]; }
    return MM2GB_OK;
}
