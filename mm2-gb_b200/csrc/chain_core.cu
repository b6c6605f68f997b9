// chain_core.cu -- host side of libmm2gb_chain.so: contexts, slots (stream + pinned staging + device buffers),
// kernel launches and the C ABI of include/mm2gb_chain.h.
//
// Replaces, B200-first, what the reference spreads over gpu/plmem.cu (buffer sizing :453-540, pinned/device
// allocation :12-143, 7 H2D + 3 memset per micro-batch :200-236, D2H :324-359) and the launch half of
// gpu/plchain.cu:292-464.  Differences that matter:
//   * anchors in pinned caller memory are DMA'd as they are (no host pass at all); anchors in pageable memory -- the per-read
//     kmalloc'd arrays of the driver -- go through ONE gather pass that writes the 8-byte packed wire format (csrc/wire.h)
//     into pinned staging, and k_expand rebuilds the 16-byte anchors in HBM (plmem.cu:154-198, the AoS->SoA repack, is gone);
//   * only chains and the INDICES of the chain anchors come back (4 bytes per chain anchor): compact_a's gather
//     (lchain.c:100-105) is done by the host from the anchors it still holds;
//   * one flat batch, no micro-batches, no host sync between "short" and "long" phases (plchain.cu:426-452 is gone);
//   * n_slots independent slots per context so upload, kernels and download of consecutive batches overlap.
#include "chain_kernels.cuh"
#include "backtrack_kernels.cuh"
#include "wire.h"
#include "host_place.h"
#include "../../include/mm2gb_chain.h"

#include <sys/mman.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <type_traits>
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
// for the other translation units of the library (seed_core.cu): same per-thread message buffer
extern "C" void mm2gb_internal_set_error(const char *msg) { snprintf(g_err, sizeof(g_err), "%s", msg ? msg : ""); }

extern "C" int mm2gb_device_count(void)
{
    int n = 0;
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);   // see mm2gb_ctx_create_ex
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// total memory of a device WITHOUT creating a context on it (a context costs seconds the first time a GPU is brought up)
extern "C" int mm2gb_device_total_memory(int device, size_t *total_bytes)
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (total_bytes) *total_bytes = prop.totalGlobalMem;
    return MM2GB_OK;
}

extern "C" int mm2gb_bind_thread_near_device(int device)
{
    cpu_set_t near;
    if (!mm2gb::gpu_local_cpus(device, &near)) return 0;
    return sched_setaffinity(0, sizeof(near), &near) == 0 ? 1 : 0;
}

extern "C" int mm2gb_device_memory(int device, size_t *free_bytes, size_t *total_bytes)
{
    int cur = 0;
    CK(cudaGetDevice(&cur));
    CK(cudaSetDevice(device));
    size_t f = 0, t = 0;
    const cudaError_t e = cudaMemGetInfo(&f, &t);
    cudaSetDevice(cur);
    if (e != cudaSuccess) return fail(MM2GB_ECUDA, "cudaMemGetInfo: %s", cudaGetErrorString(e));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return MM2GB_OK;
}

namespace {

// size classes of the chain-extraction kernels: 0..6 shared-memory kernels (k_bt_sort<CAP> / k_bt_walk<CAP>), 7..15 "mid"
// (k_bt_sort_mid / k_bt_walk_mid: keys in global scratch, digits / predecessor links in shared memory), 16 = any size
constexpr int kBtClasses = 17, kBtBig = 16, kBtMid0 = 7, kBtStreams = 24;
// (finer mid classes -- steps of 2^(1/4) -- were tried: better packing of shared memory, but twice the launches, each with its
//  own slowest-read tail; measured slower both device-resident and end to end)
// The mid caps are the largest reads of which k = 12, 10, 8, 6, 5, 4, 3, 2, 1 fit an SM in BOTH kernels: a CTA costs
// cap + 8288 + 1024 B (k_bt_sort_mid) / bt_walk_mid_smem(cap) + 2496 + 1024 B (k_bt_walk_mid) of the SM's 233472 B.  These
// latency-bound kernels are as fast as the number of reads resident per SM, and with power-of-two-ish caps the 49152 class
// missed its fourth and the 65536 class its third resident read by under 2 KB.
const int kBtCaps[kBtBig] = {1024, 1536, 2048, 3072, 4096, 6144, 8192, 10048, 13952, 19776, 29504, 37248, 48640, 65536, 98304, 196608};
// Reads of 8193 .. mid_min anchors go to the global-memory kernels, longer ones (up to 196608) to the mid kernels.  The two
// kinds complement each other: the global-memory kernels need no shared memory, so every read of a batch is resident at once
// but each serial step costs an L2 round trip (fine for the shorter reads); the mid kernels take ~5x less time per read but
// an SM only holds 227 KB / (anchors of the read) of them -- they take the long reads that would otherwise be the tail.
static int bt_mid_min()
{
    const char *e = getenv("MM2GB_BT_MID_MIN");   // read per batch: the tests switch it
    return e ? std::max(8192, atoi(e)) : 8192;
}

enum { T_RANGE = 0, T_UNITS, T_SCORE, T_BACKTRACK, T_H2D, T_D2H };

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    // device
    unsigned char *d_wire = nullptr;         // packed upload (pk | blk_run | runs, csrc/wire.h) before k_expand rebuilds d_a
    uint4 *d_a = nullptr;
    long long *d_off = nullptr;
    int *d_st = nullptr, *d_f = nullptr, *d_p = nullptr;
    unsigned *d_selmask = nullptr, *d_clipmask = nullptr;
    int *d_chunk_tot = nullptr, *d_chunk_base = nullptr; unsigned long long *d_chunk_pairs = nullptr;   // k_scan: one entry per 8192 blocks
    int *d_block_cnt = nullptr, *d_block_base = nullptr; unsigned long long *d_block_pairs = nullptr; int *d_unit_start = nullptr, *d_unit_rbase = nullptr, *d_big_order = nullptr;
    unsigned *d_unit_clip = nullptr;         // per unit: holds a window clipped by max_iter (k_unit_clip)
    int big_cap = 0;
    int *d_exact_list = nullptr;             // clipped units of >= kExactMin anchors, queued by k_score_units for k_score_exact
    Counters *d_ctr = nullptr;
    // device chain extraction (k_bt_sort / k_bt_walk): packed compacted anchors and chains of the batch, scratch, lists
    int *d_vp = nullptr;                     // packed indices of the compacted anchors (positions from Counters::b_cur)
    unsigned long long *d_uscr = nullptr;
    int *d_vs = nullptr, *d_list = nullptr;
    int *d_rinfo = nullptr;                  // per read: n_u | n_b | u_pos | b_pos, four arrays of n_reads + 1 ints
    int *d_nu = nullptr, *d_nb = nullptr, *d_upos = nullptr, *d_bpos = nullptr;   // slices of d_rinfo for the batch in flight
    unsigned long long *d_upack = nullptr;   // packed chains (positions from Counters::u_cur)
    unsigned *d_zs = nullptr;   // sorted (score, index) pairs of every read (k_bt_sort -> k_bt_walk); chain ids of the big kernels
    int *d_nz = nullptr;
    // global-memory scratch of the kernels for reads the shared-memory ones cannot take (k_bt_sort_big / k_bt_walk_big)
    unsigned long long *d_zk = nullptr, *d_zk2 = nullptr;
    unsigned *d_tb = nullptr, *d_pay2 = nullptr;
    int *d_ovf = nullptr;
    size_t u_cap = 0;      // entries of d_upack / h_upack
    cudaEvent_t bt_fork = nullptr, bt_join[kBtClasses] = {nullptr};
    int bt_cnt[kBtClasses] = {0}, bt_base[kBtClasses + 1] = {0};   // reads per size class of the batch in flight, their ranges in d_list
    // pinned host
    mm2gb_anchor_t *h_a = nullptr;
    long long *h_off = nullptr;
    int *h_f = nullptr, *h_p = nullptr;
    Counters *h_ctr = nullptr;
    int *h_vp = nullptr;                     // landing buffer of the packed chain-anchor indices (mapped)
    int *h_rinfo = nullptr, *h_list = nullptr;
    int *h_nu = nullptr, *h_nb = nullptr, *h_upos = nullptr, *h_bpos = nullptr;   // slices of h_rinfo for the batch in flight
    unsigned long long *h_upack = nullptr;
    int *h_vp_dev = nullptr;                 // device views of the mapped pinned result buffers (k_drain writes them)
    unsigned long long *h_upack_dev = nullptr;
    // state
    bool busy = false;
    int n_reads = 0;
    long long n_total = 0;
    // where results of a synchronous chunk go (mm2gb_chain_dp_host)
    int *user_f = nullptr, *user_p = nullptr;
    bool direct_out = false;
    // chains requested for this batch (device backtracking); where the compacted anchors land
    bool chains = false, want_fp = true;
    int *land_v = nullptr;                   // host view of where k_drain puts the packed indices (h_vp or the caller's buffer)
    std::vector<const uint64_t *> u_ptr;     // per read: its chains (in h_upack)
    std::vector<const int32_t *> v_ptr;      // per read: the indices of its compacted anchors (in the packed landing buffer)
    long long b_total = 0;                   // indices in the packed landing buffer
    int v_mis = 0;                           // land_v is this many entries past a 16-byte boundary: the device packs its indices from
                                             // position v_mis on, so that k_drain moves aligned 16-byte pieces on both sides
    int wire_runs = 0;                       // runs of the packed upload of the batch in flight (0: raw upload)
    size_t up_bytes = 0;                     // bytes of anchor data uploaded for the batch in flight
};

} // namespace

constexpr int kMaxSlots = 8;

// per-device state shared by every context of the process on that device (never torn down: the streams live as long as the
// process does)
struct DevicePool {
    std::mutex mu;
    bool ready = false;
    int ring = 0;
    size_t score_smem = 0;
    int score_blocks = 0, long_blocks = 0;
    cudaStream_t bt_stream[kBtStreams] = {nullptr};
    std::atomic<unsigned> rr{0};
};
constexpr int kMaxDevices = 64;
static DevicePool g_pool[kMaxDevices];

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
    int long_classes = 3;       // unit size classes (>= 8192 / 4096 / 2048 anchors) scored by k_score_long; MM2GB_LONG_MIN
    int long_blocks = 0;
    int long_wave = 0;          // the 4096 / 2048 classes go to k_score_long only while the long units fit this many CTAs (0: always)
    bool host_io = true;        // slots own pinned staging + device anchor/f/p buffers (false: device-resident entry points only)
    bool chains_ok = true;      // slots own the chain-extraction buffers (false: DP entry points only)
    bool fp_staging = true;     // slots own pinned f / p staging (false: only the chain entry points download anything)
    // The size classes of the chain-extraction kernels run side by side (each is a partial wave) on a pool of auxiliary
    // streams shared by all slots and handed out round robin, so that neither the classes of one chunk nor the same class of
    // consecutive chunks queue behind each other (the global-memory class runs on the slot's own stream).  With the slots'
    // own streams the pool stays within the hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS, raised to 32 below), so streams
    // do not alias onto one queue and serialise falsely.
    cudaStream_t *bt_stream = nullptr;                // the device's pool of class streams (DevicePool), round robin over all contexts
    std::atomic<unsigned> *bt_rr = nullptr;
    bool pin_registered = false;                      // pin_block is our own huge-page allocation, registered with CUDA
    int drain_blocks = 148;      // CTAs of k_drain (enough 16-byte stores in flight to fill PCIe; MM2GB_DRAIN_BLOCKS)
    bool exact_big = true;       // the clipped units of >= kExactMin anchors that k_score_units queues: k_score_exact (a CTA each); 0: a warp each
    bool range_tma = false;      // k_range_tma (history staged by cp.async.bulk) instead of k_range: MM2GB_RANGE_TMA=1
    int wire_mode = 0;           // 0 auto: pinned sources are DMA'd raw, pageable ones are packed by the gather pass; 1 always raw
                                 // (staged by memcpy); 2 always packed (MM2GB_WIRE=auto|raw|packed)
    size_t stage_bytes = 0;      // size of a slot's pinned staging buffer h_a / device wire buffer
    long long up_bytes_batch = 0; // anchor bytes uploaded by the last pipelined batch (all chunks)
    Slot slot[kMaxSlots];
    void *dev_block = nullptr, *pin_block = nullptr;   // all buffers of all slots (one allocation each)
    size_t dev_bytes = 0, pin_bytes = 0;
    // profiling (slot 0 only)
    bool profile = false;
    bool timeline = false;      // MM2GB_TIMELINE=1: every slot records its stages; dumped (ms since the first event) to stderr
    cudaEvent_t tl_t0 = nullptr;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> tl_pending;   // (slot << 8 | stage, events)
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
    if (const char *e = getenv("MM2GB_SCORE_CTAS")) { const int cap = atoi(e); if (cap >= 1 && cap < nb) nb = cap; }
    c->score_blocks = std::max(c->score_blocks, nb * c->n_sm);
    return MM2GB_OK;
}

template <int R>
static int config_ring(mm2gb_ctx *c)
{
    const size_t smem = (size_t)kScoreWarps * (R * sizeof(Rec) + kStageBytes);
    c->score_smem = smem;
    c->score_blocks = 0;
    int rc = config_score<R, true>(c, smem);
    if (rc) return rc;
    return config_score<R, false>(c, smem);
}

template <int R, bool FAST>
static void launch_score(mm2gb_ctx *c, cudaStream_t s, const uint4 *a, const int *st, const int *us, const int *ur,
                         const unsigned *clip, int *f, int *p, const int *big, int big_cap, Counters *ctr, int run_mode, int *exact_list)
{
    const size_t smem = (size_t)kScoreWarps * (R * sizeof(Rec) + kStageBytes);
    k_score_units<R, FAST><<<c->score_blocks, kScoreWarps * 32, smem, s>>>(a, st, us, ur, clip, f, p, big, big_cap, ctr, c->prm,
                                                                          c->d_lut, run_mode, FAST ? c->long_classes : 0, c->long_wave);
}

template <bool FAST>
static void launch_score_ring(mm2gb_ctx *c, cudaStream_t s, const uint4 *a, const int *st, const int *us, const int *ur,
                              const unsigned *clip, int *f, int *p, const int *big, int big_cap, Counters *ctr, int run_mode, int *exact_list)
{
    switch (c->ring) {
    case 256: launch_score<256, FAST>(c, s, a, st, us, ur, clip, f, p, big, big_cap, ctr, run_mode, exact_list); break;
    case 1024: launch_score<1024, FAST>(c, s, a, st, us, ur, clip, f, p, big, big_cap, ctr, run_mode, exact_list); break;
    default: launch_score<512, FAST>(c, s, a, st, us, ur, clip, f, p, big, big_cap, ctr, run_mode, exact_list); break;
    }
}

static int config_long(mm2gb_ctx *c)
{
    const size_t smem = (size_t)kLongRing * sizeof(RecL);
    CK(cudaFuncSetAttribute(k_score_long<kLongRing, kLongWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_score_exact<kExactRing, kExactWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kExactRing * sizeof(RecL))));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_score_long<kLongRing, kLongWarps>, kLongWarps * 32, smem));
    if (nb < 1) return fail(MM2GB_ECUDA, "long score kernel does not fit on an SM (smem %zu)", smem);
    c->long_blocks = nb * c->n_sm;
    c->long_wave = c->long_blocks;
    if (const char *e = getenv("MM2GB_LONG_WAVE")) c->long_wave = atoi(e);
    return MM2GB_OK;
}

static thread_local int g_tl_slot = 0;   // slot being enqueued (timeline only)

struct ProfScope {
    mm2gb_ctx *c; int id; cudaStream_t s; cudaEvent_t e0 = nullptr, e1 = nullptr; bool on, tl;
    ProfScope(mm2gb_ctx *c_, int id_, cudaStream_t s_, bool on_) : c(c_), id(id_), s(s_), on(on_ && !c_->timeline), tl(c_->timeline)
    {
        if (!on && !tl) return;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        if (tl && !c->tl_t0) { cudaEventCreate(&c->tl_t0); cudaEventRecord(c->tl_t0, s); }
        cudaEventRecord(e0, s);
    }
    ~ProfScope()
    {
        if (!on && !tl) return;
        cudaEventRecord(e1, s);
        if (tl) c->tl_pending.push_back({(g_tl_slot << 8) | id, {e0, e1}});
        else c->ev_pending.push_back({id, {e0, e1}});
    }
};

static void timeline_dump(mm2gb_ctx *c)
{
    static const char *names[] = {"range", "units", "score", "backtrack", "h2d", "d2h"};
    if (!c->tl_t0) return;
    cudaDeviceSynchronize();
    for (auto &pe : c->tl_pending) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, c->tl_t0, pe.second.first);
        cudaEventElapsedTime(&b, c->tl_t0, pe.second.second);
        fprintf(stderr, "TL slot=%d stage=%s t0=%.3f t1=%.3f\n", pe.first >> 8, names[pe.first & 255], a, b);
        cudaEventDestroy(pe.second.first);
        cudaEventDestroy(pe.second.second);
    }
    c->tl_pending.clear();
    cudaEventDestroy(c->tl_t0);
    c->tl_t0 = nullptr;
    fprintf(stderr, "TL end\n");
}

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
        if (c->range_tma)
            k_range_tma<<<n_blocks, kRangeThreads, 0, s>>>(reinterpret_cast<const ulonglong2 *>(d_a), d_off, sl.d_block_base, n, c->prm, sl.d_st,
                                                          sl.d_selmask, sl.d_clipmask, sl.d_block_cnt, sl.d_block_pairs, sl.d_ctr);
        else
            k_range<<<n_blocks, kRangeThreads, 0, s>>>(reinterpret_cast<const ulonglong2 *>(d_a), d_off, sl.d_block_base, n, c->prm, sl.d_st,
                                                      sl.d_selmask, sl.d_clipmask, sl.d_block_cnt, sl.d_block_pairs, sl.d_ctr);
    }
    {
        ProfScope ps(c, T_UNITS, s, prof);
        k_scan<<<(n_blocks + kScanChunk - 1) / kScanChunk, 1024, 0, s>>>(sl.d_block_cnt, sl.d_block_pairs, n_blocks, sl.d_block_base, sl.d_chunk_tot,
                                                                        sl.d_chunk_pairs, sl.d_chunk_base, sl.d_ctr);
        k_units<<<(n_groups + 255) / 256, 256, 0, s>>>(sl.d_selmask, sl.d_block_base, sl.d_chunk_base, d_off, n_reads, n, n_groups, sl.d_unit_start,
                                                      sl.d_unit_rbase, sl.d_unit_clip, sl.d_ctr);
        k_unit_clip<<<(n_groups + 255) / 256, 256, 0, s>>>(sl.d_clipmask, sl.d_unit_start, n_groups, sl.d_unit_clip, sl.d_ctr);
        k_order<<<(n_groups + n_reads + 256) / 256, 256, 0, s>>>(sl.d_unit_start, sl.d_unit_clip, sl.d_big_order, sl.big_cap, sl.d_ctr, c->fast ? 1 : 0);
    }
    {
        ProfScope ps(c, T_SCORE, s, prof);
        if (c->fast) {
            if (c->long_classes > 0)   // long units first: one CTA each, warps pipelined over the tiles
                k_score_long<kLongRing, kLongWarps><<<c->long_blocks, kLongWarps * 32, (size_t)kLongRing * sizeof(RecL), s>>>(
                    d_a, sl.d_st, sl.d_unit_start, sl.d_unit_rbase, sl.d_clipmask, d_f, d_p, sl.d_big_order, sl.big_cap, sl.d_ctr, c->prm, c->d_lut,
                    c->long_classes, c->long_wave);
            int *xl = sl.d_exact_list;
            launch_score_ring<true>(c, s, d_a, sl.d_st, sl.d_unit_start, sl.d_unit_rbase, sl.d_clipmask, d_f, d_p, sl.d_big_order, sl.big_cap, sl.d_ctr, 1, xl);
            // clipped units of >= kExactMin anchors queued by the kernel above: one CTA each (exits at once if there are none)
            if (c->exact_big)
                k_score_exact<kExactRing, kExactWarps><<<c->n_sm, kExactWarps * 32, (size_t)kExactRing * sizeof(RecL), s>>>(
                    d_a, sl.d_st, sl.d_unit_start, sl.d_unit_rbase, d_f, d_p, xl, sl.d_ctr, c->prm, c->d_lut);
            else
                k_score_exact_warp<<<c->n_sm, 256, 0, s>>>(d_a, sl.d_st, sl.d_unit_start, sl.d_unit_rbase, d_f, d_p, xl, sl.d_ctr, c->prm, c->d_lut);
            launch_score_ring<false>(c, s, d_a, sl.d_st, sl.d_unit_start, sl.d_unit_rbase, sl.d_clipmask, d_f, d_p, sl.d_big_order, sl.big_cap, sl.d_ctr, 2, nullptr);
        } else {
            launch_score_ring<false>(c, s, d_a, sl.d_st, sl.d_unit_start, sl.d_unit_rbase, sl.d_clipmask, d_f, d_p, sl.d_big_order, sl.big_cap, sl.d_ctr, 0, nullptr);
        }
    }
    CK(cudaGetLastError());
    return MM2GB_OK;
}


// ---- device chain extraction ------------------------------------------------------------------------------------------

#define MM2GB_BT_CLASSES(X) X(1024) X(1536) X(2048) X(3072) X(4096) X(6144) X(8192)

static int config_backtrack()
{
#define X(CAP)                                                                                                                        \
    CK(cudaFuncSetAttribute(k_bt_sort<CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BtSortSmem<CAP>)));            \
    CK(cudaFuncSetAttribute(k_bt_walk<CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BtWalkSmem<CAP>)));
    MM2GB_BT_CLASSES(X)
#undef X
    CK(cudaFuncSetAttribute(k_bt_sort_mid<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBtCaps[kBtBig - 1] + 16));
    CK(cudaFuncSetAttribute(k_bt_sort_mid<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBtCaps[kBtBig - 1] + 16));
    CK(cudaFuncSetAttribute(k_bt_sort_mid<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBtCaps[kBtBig - 1] + 16));
    CK(cudaFuncSetAttribute(k_bt_walk_mid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bt_walk_mid_smem(kBtCaps[kBtBig - 1])));
    if (const char *e = getenv("MM2GB_VERBOSE")) if (atoi(e) >= 3) {   // reads resident per SM, per mid class
        for (int k = kBtMid0; k < kBtBig; ++k) {
            int ns = 0, nw = 0;
            const int nt = bt_mid_threads(kBtCaps[k]);
            if (nt == 128) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ns, k_bt_sort_mid<128>, nt, (size_t)kBtCaps[k] + 16));
            else if (nt == 256) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ns, k_bt_sort_mid<256>, nt, (size_t)kBtCaps[k] + 16));
            else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ns, k_bt_sort_mid<512>, nt, (size_t)kBtCaps[k] + 16));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nw, k_bt_walk_mid, 32, bt_walk_mid_smem(kBtCaps[k])));
            fprintf(stderr, "[mm2gb] chain-extraction class %6d anchors: %d sorts / %d walks resident per SM\n", kBtCaps[k], ns, nw);
        }
    }
    return MM2GB_OK;
}

template <int CAP>
static void launch_backtrack(cudaStream_t s, const uint4 *d_a, const int *d_f, const int *d_p, const long long *d_off, const int *list, int n_list,
                             const BtParams &bp, Slot &sl)
{
    if (n_list <= 0) return;
    k_bt_sort<CAP><<<n_list, 32, sizeof(BtSortSmem<CAP>), s>>>(d_f, d_off, list, n_list, bp, sl.d_zs, sl.d_nz, sl.d_ovf, sl.d_ctr);
    k_bt_walk<CAP><<<n_list, 32, sizeof(BtWalkSmem<CAP>), s>>>(d_a, d_f, d_p, d_off, list, n_list, bp, sl.d_zs, sl.d_nz, sl.d_st /* dead after scoring: v scratch */,
                                                           sl.d_uscr, sl.d_vs, sl.d_vp, sl.d_upack, (int)sl.u_cap, sl.d_nu, sl.d_nb, sl.d_upos, sl.d_bpos,
                                                           sl.d_ovf, sl.d_ctr);
}

// the global-memory kernels: over the host's list of big reads (ovf == false) or over the device-side overflow list, which
// holds at most the n_list reads the shared-memory kernels were given (CTAs beyond its length exit at once)
static void launch_backtrack_big(cudaStream_t s, const uint4 *d_a, const int *d_f, const int *d_p, const long long *d_off, const int *list, int n_list,
                                 bool ovf, const BtParams &bp, Slot &sl)
{
    const int grid = n_list;
    if (grid <= 0) return;
    const int *ovf_list = ovf ? sl.d_ovf : nullptr;
    k_bt_sort_big<<<grid, 32, 0, s>>>(d_f, d_off, list, n_list, ovf_list, sl.d_ctr, bp, sl.d_zk, sl.d_zk2, sl.d_nz);
    k_bt_walk_big<<<grid, 32, 0, s>>>(d_a, d_f, d_p, d_off, list, n_list, ovf_list, bp, sl.d_zk, sl.d_zk2, sl.d_nz, sl.d_tb, sl.d_zs, sl.d_pay2,
                                      sl.d_st, sl.d_uscr, sl.d_vs, sl.d_vp, sl.d_upack, (int)sl.u_cap, sl.d_nu, sl.d_nb, sl.d_upos, sl.d_bpos, sl.d_ctr);
}

// reads of up to cap anchors (8193 .. 196608): global-memory keys, serial parts in shared memory
static void launch_backtrack_mid(cudaStream_t s, const uint4 *d_a, const int *d_f, const int *d_p, const long long *d_off, const int *list, int n_list,
                                 int cap, const BtParams &bp, Slot &sl)
{
    if (n_list <= 0) return;
    const size_t sm = (size_t)cap + 16;   // + D[m], read ahead by the serial walk
    unsigned *tok = reinterpret_cast<unsigned *>(sl.d_vs);
    switch (bt_mid_threads(cap)) {
    case 128: k_bt_sort_mid<128><<<n_list, 128, sm, s>>>(d_f, d_off, list, n_list, bp, sl.d_zk, sl.d_zk2, sl.d_zs, sl.d_pay2, tok, sl.d_nz, cap); break;
    case 256: k_bt_sort_mid<256><<<n_list, 256, sm, s>>>(d_f, d_off, list, n_list, bp, sl.d_zk, sl.d_zk2, sl.d_zs, sl.d_pay2, tok, sl.d_nz, cap); break;
    default: k_bt_sort_mid<512><<<n_list, 512, sm, s>>>(d_f, d_off, list, n_list, bp, sl.d_zk, sl.d_zk2, sl.d_zs, sl.d_pay2, tok, sl.d_nz, cap); break;
    }
    k_bt_walk_mid<<<n_list, 32, bt_walk_mid_smem(cap), s>>>(d_a, d_f, d_p, d_off, list, n_list, bp, sl.d_zk, sl.d_zk2, sl.d_nz, sl.d_zs, sl.d_pay2,
                                                          sl.d_st, sl.d_uscr, sl.d_vs, sl.d_vp, sl.d_upack, (int)sl.u_cap, sl.d_nu, sl.d_nb, sl.d_upos,
                                                          sl.d_bpos, sl.d_ctr, cap);
}

// the four per-read result arrays of a batch of n_reads reads, device and pinned host side
static void slice_rinfo(Slot &sl, int n_reads)
{
    const size_t rs = (size_t)n_reads + 1;
    sl.d_nu = sl.d_rinfo; sl.d_nb = sl.d_rinfo + rs; sl.d_upos = sl.d_rinfo + 2 * rs; sl.d_bpos = sl.d_rinfo + 3 * rs;
    sl.h_nu = sl.h_rinfo; sl.h_nb = sl.h_rinfo + rs; sl.h_upos = sl.h_rinfo + 2 * rs; sl.h_bpos = sl.h_rinfo + 3 * rs;
}

// Reads are binned by anchor count into the shared-memory classes of k_bt_sort / k_bt_walk (off_rel is the host copy of the
// offsets); reads above kBtMaxAnchors go to the global-memory kernels, and so does -- through a device-side list -- any read a
// shared-memory kernel cannot finish (scores that do not pack into 32 bits, more chains than its key buffer holds).
// step 1 (before the anchors are uploaded, so this small copy does not queue behind them on the copy engine): bin the reads
// by size class and upload the per-class read lists
static int prepare_backtrack(Slot &sl, cudaStream_t s, const long long *off_rel, int n_reads)
{
    int *cnt = sl.bt_cnt, *base = sl.bt_base, fill[kBtClasses];
    const int mid_min = bt_mid_min();
    auto cls = [&](long long n) {
        for (int k = 0; k < kBtMid0; ++k) if (n <= kBtCaps[k]) return k;
        if (n <= mid_min) return kBtBig;
        for (int k = kBtMid0; k < kBtBig; ++k) if (n <= kBtCaps[k]) return k;
        return kBtBig;
    };
    for (int k = 0; k < kBtClasses; ++k) cnt[k] = 0;
    for (int r = 0; r < n_reads; ++r) ++cnt[cls(off_rel[r + 1] - off_rel[r])];
    base[0] = 0;
    for (int k = 0; k < kBtClasses; ++k) { base[k + 1] = base[k] + cnt[k]; fill[k] = base[k]; }
    for (int r = 0; r < n_reads; ++r) sl.h_list[fill[cls(off_rel[r + 1] - off_rel[r])]++] = r;
    slice_rinfo(sl, n_reads);
    if (base[kBtClasses]) CK(cudaMemcpyAsync(sl.d_list, sl.h_list, (size_t)base[kBtClasses] * sizeof(int), cudaMemcpyHostToDevice, s));
    return MM2GB_OK;
}

// step 2: the kernels
static int enqueue_backtrack(mm2gb_ctx *c, Slot &sl, cudaStream_t s, const uint4 *d_a, const long long *d_off, int n_reads, const int *d_f,
                             const int *d_p, bool prof)
{
    const int *cnt = sl.bt_cnt, *base = sl.bt_base;
    const size_t rs = (size_t)n_reads + 1;
    CK(cudaMemsetAsync(sl.d_nu, 0xff, rs * sizeof(int), s));   // -1 = not finished (must not survive the overflow pass)
    CK(cudaMemsetAsync(sl.d_nb, 0, 3 * rs * sizeof(int), s));
    CK(cudaMemsetAsync(&sl.d_ctr->ovf_cnt, 0, 3 * sizeof(int), s));   // ovf_cnt, u_cur, b_cur
    if (sl.v_mis) CK(cudaMemsetAsync(&sl.d_ctr->b_cur, sl.v_mis, 1, s));   // low byte (little endian): b_cur = v_mis
    BtParams bp;
    bp.min_cnt = c->misc.min_cnt;
    // scores are never negative (f[i] >= q_span(i) >= 0, lchain.c:171) and an accepted chain has a positive score, so a negative
    // min_score selects exactly what 0 selects; the packed keys of the sort kernels hold unsigned scores
    bp.min_sc = std::max(0, c->misc.min_score);
    bp.max_drop = c->misc.is_cdna ? INT32_MAX : c->misc.bw;   // lchain.c:151,162
    {
        // fork: one auxiliary stream per non-empty size class, joined back into the slot's stream
        ProfScope ps(c, T_BACKTRACK, s, prof);
        CK(cudaEventRecord(sl.bt_fork, s));
        for (int k = kBtBig - 1; k >= 0; --k) { // longest first
            if (!cnt[k]) continue;
            cudaStream_t bs = c->bt_stream[c->bt_rr->fetch_add(1) % kBtStreams];
            CK(cudaStreamWaitEvent(bs, sl.bt_fork, 0));
            const int *list = sl.d_list + base[k];
            if (k >= kBtMid0) launch_backtrack_mid(bs, d_a, d_f, d_p, d_off, list, cnt[k], kBtCaps[k], bp, sl);
            else switch (k) {
            case 6: launch_backtrack<8192>(bs, d_a, d_f, d_p, d_off, list, cnt[k], bp, sl); break;
            case 5: launch_backtrack<6144>(bs, d_a, d_f, d_p, d_off, list, cnt[k], bp, sl); break;
            case 4: launch_backtrack<4096>(bs, d_a, d_f, d_p, d_off, list, cnt[k], bp, sl); break;
            case 3: launch_backtrack<3072>(bs, d_a, d_f, d_p, d_off, list, cnt[k], bp, sl); break;
            case 2: launch_backtrack<2048>(bs, d_a, d_f, d_p, d_off, list, cnt[k], bp, sl); break;
            case 1: launch_backtrack<1536>(bs, d_a, d_f, d_p, d_off, list, cnt[k], bp, sl); break;
            default: launch_backtrack<1024>(bs, d_a, d_f, d_p, d_off, list, cnt[k], bp, sl); break;
            }
            CK(cudaEventRecord(sl.bt_join[k], bs));
        }
        // the big reads stay on the slot's own stream (next to the forked classes): their kernels are a few long serial warps,
        // and on a stream shared by all slots the chunks of a batch of long reads would queue behind each other
        if (cnt[kBtBig]) launch_backtrack_big(s, d_a, d_f, d_p, d_off, sl.d_list + base[kBtBig], cnt[kBtBig], false, bp, sl);
        for (int k = kBtBig - 1; k >= 0; --k)
            if (cnt[k]) CK(cudaStreamWaitEvent(s, sl.bt_join[k], 0));
        // whatever the shared-memory kernels handed over (normally nothing: CTAs that exit at once)
        if (base[kBtMid0]) launch_backtrack_big(s, d_a, d_f, d_p, d_off, nullptr, base[kBtMid0], true, bp, sl);
    }
    CK(cudaGetLastError());
    return MM2GB_OK;
}

// packed results of the batch -> mapped pinned host memory (dst_v: device view of where the chain-anchor indices land)
static int enqueue_drain(mm2gb_ctx *c, Slot &sl, cudaStream_t s, int *dst_v)
{
    k_drain<<<c->drain_blocks, kDrainThreads, 0, s>>>(sl.d_vp, dst_v - sl.v_mis, sl.v_mis, sl.d_upack, sl.h_upack_dev, sl.d_ctr);
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
    if (s.done) cudaEventDestroy(s.done);
    if (s.bt_fork) cudaEventDestroy(s.bt_fork);
    for (int k = 0; k < kBtClasses; ++k) if (s.bt_join[k]) cudaEventDestroy(s.bt_join[k]);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Slot();     // the buffers live in the context's two blocks
}

// ---- C ABI -------------------------------------------------------------------------------------------------------------

extern "C" int mm2gb_ctx_create(mm2gb_ctx_t **out, int device, size_t max_anchors, int max_reads, int n_slots, const mm2gb_misc_t *misc)
{
    return mm2gb_ctx_create_ex(out, device, max_anchors, max_reads, n_slots, misc, 0);
}

extern "C" int mm2gb_ctx_create_ex(mm2gb_ctx_t **out, int device, size_t max_anchors, int max_reads, int n_slots, const mm2gb_misc_t *misc,
                                   unsigned flags)
{
    if (!out || !misc) return fail(MM2GB_EARG, "null argument");
    *out = nullptr;
    if (n_slots < 1 || n_slots > kMaxSlots) return fail(MM2GB_EARG, "n_slots must be 1..%d", kMaxSlots);
    if (max_anchors == 0 || max_anchors > (size_t)INT32_MAX - 1024) return fail(MM2GB_EARG, "max_anchors must be in (0, 2^31)");
    if (max_reads < 1) return fail(MM2GB_EARG, "max_reads must be positive");
    // one hardware work queue per stream (default is 8 for the whole process); only effective before CUDA is initialised
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev || device >= kMaxDevices) return fail(MM2GB_EARG, "no CUDA device %d (have %d)", device, ndev);
    CK(cudaSetDevice(device));
    mm2gb_ctx *c = new mm2gb_ctx();
    c->device = device;
    c->max_anchors = max_anchors;
    c->max_reads = max_reads;
    c->n_slots = n_slots;
    c->host_io = !(flags & MM2GB_CTX_DEVICE_ONLY);
    c->chains_ok = !(flags & MM2GB_CTX_NO_CHAINS);
    c->fp_staging = !(flags & MM2GB_CTX_NO_FP_STAGING);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->n_sm = prop.multiProcessorCount;
    if (const char *e = getenv("MM2GB_RING")) {
        int r = atoi(e);
        if (r == 256 || r == 512 || r == 1024) c->ring = r;
    }
    if (const char *e = getenv("MM2GB_LONG_MIN")) { // smallest unit scored by k_score_long: 2048 / 4096 / 8192, 0 = never
        const int v = atoi(e);
        c->long_classes = v <= 0 ? 0 : v <= 2048 ? 3 : v <= 4096 ? 2 : 1;
    }
    if (const char *e = getenv("MM2GB_TIMELINE")) c->timeline = atoi(e) != 0;
    if (const char *e = getenv("MM2GB_RANGE_TMA")) c->range_tma = atoi(e) != 0;
    if (const char *e = getenv("MM2GB_EXACT_BIG")) c->exact_big = atoi(e) != 0;
    if (const char *e = getenv("MM2GB_WIRE")) c->wire_mode = !strcmp(e, "raw") ? 1 : !strcmp(e, "packed") ? 2 : 0;
    if (const char *e = getenv("MM2GB_DRAIN_BLOCKS")) {
        int r = atoi(e);
        if (r >= 1 && r <= 4096) c->drain_blocks = r;
    }
    // where the time of a context creation goes (MM2GB_VERBOSE >= 3; driver threads create theirs concurrently and the CUDA
    // driver serialises most of these calls)
    const bool vt = getenv("MM2GB_VERBOSE") && atoi(getenv("MM2GB_VERBOSE")) >= 3;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_phase = now();
    auto phase = [&](const char *what) {
        if (!vt) return;
        const double t = now();
        fprintf(stderr, "[mm2gb] ctx %p: %-28s %8.3f ms\n", (void *)c, what, 1e3 * (t - t_phase));
        t_phase = t;
    };
    phase("device + properties");
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
        phase("penalty table");
        {
            // what is the same for every context of a device is set up once: the kernels' shared-memory attributes and grid sizes,
            // and the pool of class streams.  (Sixteen driver threads creating their contexts at once spent seconds queueing for the
            // driver's lock on ~30 attribute calls and 24 stream creations EACH; profiles/r5d_ctx_probe.txt.)
            DevicePool &dp = g_pool[device];
            std::lock_guard<std::mutex> lk(dp.mu);
            if (!dp.ready || dp.ring != c->ring) {
                rc = c->ring == 256 ? config_ring<256>(c) : c->ring == 1024 ? config_ring<1024>(c) : config_ring<512>(c);
                if (rc) goto bad;
                rc = config_long(c);
                if (rc) goto bad;
                rc = config_backtrack();
                if (rc) goto bad;
                for (int k = 0; k < kBtStreams; ++k)
                    if (!dp.bt_stream[k]) CKC(cudaStreamCreateWithFlags(&dp.bt_stream[k], cudaStreamNonBlocking));
                dp.ring = c->ring; dp.score_smem = c->score_smem; dp.score_blocks = c->score_blocks; dp.long_blocks = c->long_blocks;
                dp.ready = true;
            } else {
                c->score_smem = dp.score_smem; c->score_blocks = dp.score_blocks; c->long_blocks = dp.long_blocks;
            }
            c->long_wave = c->long_blocks;
            if (const char *e = getenv("MM2GB_LONG_WAVE")) c->long_wave = atoi(e);
            c->bt_stream = dp.bt_stream;
            c->bt_rr = &dp.rr;
        }
        phase("kernel attributes + streams");
        phase("24 class streams");
        const size_t n = max_anchors, n_groups = (n + 31) / 32, n_blocks = (n + kRangeThreads - 1) / kRangeThreads;
        const size_t n_units_cap = n_groups + (size_t)max_reads + 2;
        const size_t n_chunks = n_blocks / kScanChunk + 2, n_rd = (size_t)max_reads + 1;
        c->stage_bytes = n * sizeof(uint4);
        // Every buffer of every slot is carved out of ONE device block and ONE pinned block per context: a context is a few
        // dozen buffers per slot, and allocation calls serialise inside the driver -- sixteen driver threads creating their
        // contexts buffer by buffer spent 2-10 s each in cudaMalloc / cudaMallocHost (profiles/r4g_driver_ont_*.json).
        // The same layout function runs twice: once to size the blocks, once to hand out the pointers.
        auto layout = [&](char *dev, char *pin, size_t *dev_bytes, size_t *pin_bytes) {
            size_t d = 0, h = 0;
            auto dtake = [&](auto *&ptr, size_t count) {
                typedef typename std::remove_reference<decltype(*ptr)>::type T;
                d = (d + 255) & ~(size_t)255;
                if (dev) ptr = reinterpret_cast<T *>(dev + d);
                d += count * sizeof(T);
            };
            auto htake = [&](auto *&ptr, size_t count) {
                typedef typename std::remove_reference<decltype(*ptr)>::type T;
                h = (h + 255) & ~(size_t)255;
                if (pin) ptr = reinterpret_cast<T *>(pin + h);
                h += count * sizeof(T);
            };
            for (int i = 0; i < n_slots; ++i) {
                Slot &s = c->slot[i];
                if (c->host_io) dtake(s.d_a, n);
                dtake(s.d_off, n_rd);
                dtake(s.d_st, n);
                if (c->host_io) { dtake(s.d_f, n); dtake(s.d_p, n); }
                dtake(s.d_selmask, n_groups); dtake(s.d_clipmask, n_groups);
                dtake(s.d_block_cnt, n_blocks); dtake(s.d_block_base, n_blocks); dtake(s.d_block_pairs, n_blocks);
                dtake(s.d_chunk_tot, n_chunks); dtake(s.d_chunk_base, n_chunks); dtake(s.d_chunk_pairs, n_chunks);
                dtake(s.d_unit_start, n_units_cap); dtake(s.d_unit_rbase, n_units_cap); dtake(s.d_unit_clip, n_units_cap);
                s.big_cap = (int)(n / kBigMin) + 2;
                dtake(s.d_big_order, (size_t)kBigClasses * s.big_cap + n / kExactMin + 2);   // + the queue of k_score_exact
                if (dev) s.d_exact_list = s.d_big_order + (size_t)kBigClasses * s.big_cap;
                dtake(s.d_ctr, 1);
                // staging: raw anchors (16 B each) or the packed wire format (8 B each + block index + run list), same buffer
                if (c->host_io) htake(s.h_a, n);
                htake(s.h_off, n_rd);
                if (c->host_io && c->fp_staging) { htake(s.h_f, n); htake(s.h_p, n); }
                htake(s.h_ctr, 1);
                if (!c->chains_ok) {
                    if (c->host_io) dtake(s.d_wire, c->stage_bytes);   // no sort scratch to borrow (see below)
                    continue;
                }
                dtake(s.d_vp, n + 4); dtake(s.d_uscr, n); dtake(s.d_vs, n);
                dtake(s.d_rinfo, 4 * n_rd); dtake(s.d_list, n_rd);
                s.u_cap = n;    // a chain has at least one anchor and every anchor is in at most one chain
                dtake(s.d_upack, s.u_cap); dtake(s.d_zs, n); dtake(s.d_nz, n_rd);
                dtake(s.d_zk, 2 * n);
                if (dev) {
                    s.d_zk2 = s.d_zk + n;
                    // on the device the packed upload lands in the sort scratch of the chain extraction, which is dead until the
                    // score kernels of the batch are done
                    if (c->host_io) s.d_wire = reinterpret_cast<unsigned char *>(s.d_zk);
                }
                dtake(s.d_pay2, n); dtake(s.d_tb, n / 32 + 2 * (size_t)max_reads + 8); dtake(s.d_ovf, n_rd);
                if (c->host_io) { htake(s.h_vp, n + 4); htake(s.h_upack, s.u_cap); }
                htake(s.h_rinfo, 4 * n_rd); htake(s.h_list, n_rd);
            }
            *dev_bytes = d + 256;
            *pin_bytes = h + 256;
        };
        size_t dev_bytes = 0, pin_bytes = 0;
        layout(nullptr, nullptr, &dev_bytes, &pin_bytes);
        CKC(cudaMalloc(&c->dev_block, dev_bytes));
        phase("device block");
        {
            // The staging block: transparent-huge-page backed memory of our own, registered with CUDA.  Pinning is paid per page,
            // so 2 MB pages make this ~10x cheaper than cudaMallocHost's 4 KB pages (which took ~100 ms per 230 MB and serialised
            // the driver threads), and the host pass over the staging area takes fewer TLB misses.  MM2GB_PIN=alloc: cudaMallocHost.
            // The pages are first touched / pinned on the NUMA node of the GPU (host_place.h).
            mm2gb::NearGpu near_gpu(device);
            const char *pm = getenv("MM2GB_PIN");
            bool done = false;
            if (!pm || strcmp(pm, "alloc") != 0) {
                const size_t huge = (size_t)2 << 20, bytes = (pin_bytes + huge - 1) & ~(huge - 1);
                void *ptr = nullptr;
                if (posix_memalign(&ptr, huge, bytes) == 0 && ptr) {
                    madvise(ptr, bytes, MADV_HUGEPAGE);
                    for (size_t o = 0; o < bytes; o += 4096) static_cast<volatile char *>(ptr)[o] = 0;   // fault the pages in
                    if (cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) == cudaSuccess) {
                        c->pin_block = ptr; c->pin_registered = true; done = true;
                    } else {
                        cudaGetLastError();
                        free(ptr);
                    }
                }
            }
            if (!done) CKC(cudaMallocHost(&c->pin_block, pin_bytes));
        }
        phase("pinned block");
        layout(static_cast<char *>(c->dev_block), static_cast<char *>(c->pin_block), &dev_bytes, &pin_bytes);
        c->dev_bytes = dev_bytes; c->pin_bytes = pin_bytes;
        for (int i = 0; i < n_slots; ++i) {
            Slot &s = c->slot[i];
            CKC(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
            CKC(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
            CKC(cudaEventCreateWithFlags(&s.bt_fork, cudaEventDisableTiming));
            for (int k = 0; k < kBtClasses; ++k) CKC(cudaEventCreateWithFlags(&s.bt_join[k], cudaEventDisableTiming));
            if (c->chains_ok && c->host_io) {
                CKC(cudaHostGetDevicePointer((void **)&s.h_vp_dev, s.h_vp, 0));
                CKC(cudaHostGetDevicePointer((void **)&s.h_upack_dev, s.h_upack, 0));
            }
        }
    }
    phase("slot streams + events");
    if (vt) fprintf(stderr, "[mm2gb] ctx %p: %.1f MB device, %.1f MB pinned\n", (void *)c, c->dev_bytes / 1e6, c->pin_bytes / 1e6);
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
    for (int i = 0; i < kMaxSlots; ++i) free_slot(c->slot[i]);
    cudaFree(c->dev_block);
    if (c->pin_registered) { if (c->pin_block) { cudaHostUnregister(c->pin_block); free(c->pin_block); } }
    else cudaFreeHost(c->pin_block);
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

// what a batch should produce
struct Want {
    bool fp = true;                 // download f / p
    int *dst_f = nullptr, *dst_p = nullptr;
    bool dst_pinned = false;
    bool chains = false;            // run chain extraction + compaction on the device and download the result
    int *dst_v = nullptr;           // mapped pinned landing area for the packed chain-anchor indices (nullptr: the slot's own) ...
    int *dst_v_dev = nullptr;       // ... and its device view
};

// where the anchors of a batch come from: one flat array, or one array per read (the driver's chain_read_t.a, plutils.h:64)
struct Source {
    const mm2gb_anchor_t *flat = nullptr;
    bool flat_pinned = false;
    const mm2gb_anchor_t *const *read_a = nullptr;
    const int64_t *read_n = nullptr;
};

// Host half of the upload.  Anchors in pinned memory are DMA'd from where they are; everything else goes through one pass
// into the slot's pinned staging buffer, which writes the packed wire format (8 B/anchor, csrc/wire.h) unless the run list
// would not fit (then, or with MM2GB_WIRE=raw, the pass is a plain copy).  *up_src / *up_bytes: what to copy to the device;
// s.wire_runs > 0 says the copy is in the packed format.
static int stage_anchors(mm2gb_ctx *c, Slot &s, const Source &src, const long long *off_rel, int n_reads, long long n_total,
                         const void **up_src, size_t *up_bytes, WireLayout *lay)
{
    s.wire_runs = 0;
    *up_src = nullptr;
    *up_bytes = 0;
    if (n_total == 0) return MM2GB_OK;
    if (src.flat && src.flat_pinned && c->wire_mode != 2) {
        *up_src = src.flat;
        *up_bytes = (size_t)n_total * sizeof(mm2gb_anchor_t);
        return MM2GB_OK;
    }
    if (c->wire_mode != 1) {
        // worth it only while the run list stays small: at most 12 of the raw format's 16 bytes per anchor
        const WireLayout L = wire_layout(n_total, std::min(c->stage_bytes, (size_t)n_total * 12));
        if (L.run_cap >= 1) {
            WirePacker pk;
            pk.begin(s.h_a, L);
            bool ok = true;
            if (src.flat) ok = pk.add(src.flat, n_total);
            else
                for (int r = 0; r < n_reads && ok; ++r)
                    if (src.read_n[r] > 0) ok = pk.add(src.read_a[r], off_rel[r + 1] - off_rel[r]);
            if (ok) {
                *up_bytes = pk.finish();
                *up_src = s.h_a;
                *lay = L;
                s.wire_runs = pk.n_runs();
                return MM2GB_OK;
            }
        }
    }
    if (src.flat) memcpy(s.h_a, src.flat, (size_t)n_total * sizeof(mm2gb_anchor_t));
    else
        for (int r = 0; r < n_reads; ++r)
            if (src.read_n[r] > 0) memcpy(s.h_a + off_rel[r], src.read_a[r], (size_t)(off_rel[r + 1] - off_rel[r]) * sizeof(mm2gb_anchor_t));
    *up_src = s.h_a;
    *up_bytes = (size_t)n_total * sizeof(mm2gb_anchor_t);
    return MM2GB_OK;
}

// enqueue one batch: host staging, H2D (+ k_expand), kernels, D2H on the slot's stream
static int submit_impl(mm2gb_ctx *c, int si, const Source &src, const long long *off_rel, int n_reads, long long n_total, const Want &w)
{
    Slot &s = c->slot[si];
    if (s.busy) return fail(MM2GB_ESTATE, "slot %d is busy", si);
    if (!c->host_io) return fail(MM2GB_ESTATE, "context was created with MM2GB_CTX_DEVICE_ONLY: host-buffer entry points are not available");
    if (w.chains && !c->chains_ok) return fail(MM2GB_ESTATE, "context was created with MM2GB_CTX_NO_CHAINS");
    if (w.fp && !c->fp_staging && !(w.dst_pinned && w.dst_f && w.dst_p)) return fail(MM2GB_ESTATE, "context was created with MM2GB_CTX_NO_FP_STAGING");
    if ((size_t)n_total > c->max_anchors) return fail(MM2GB_ECAP, "batch of %lld anchors exceeds capacity %zu", n_total, c->max_anchors);
    if (n_reads > c->max_reads) return fail(MM2GB_ECAP, "batch of %d reads exceeds capacity %d", n_reads, c->max_reads);
    CK(cudaSetDevice(c->device));
    const bool prof = c->profile && si == 0;
    g_tl_slot = si;
    memcpy(s.h_off, off_rel, ((size_t)n_reads + 1) * sizeof(long long));
    const void *up_src = nullptr;
    size_t up_bytes = 0;
    WireLayout lay{};
    int rc = stage_anchors(c, s, src, s.h_off, n_reads, n_total, &up_src, &up_bytes, &lay);
    if (rc) return rc;
    s.up_bytes = up_bytes;
    {
        ProfScope ps(c, T_H2D, s.stream, prof);
        CK(cudaMemcpyAsync(s.d_off, s.h_off, ((size_t)n_reads + 1) * sizeof(long long), cudaMemcpyHostToDevice, s.stream));
        if (w.chains && n_total) { int rc0 = prepare_backtrack(s, s.stream, s.h_off, n_reads); if (rc0) return rc0; }
        if (n_total && s.wire_runs > 0) {
            CK(cudaMemcpyAsync(s.d_wire, up_src, up_bytes, cudaMemcpyHostToDevice, s.stream));
            k_expand<<<(unsigned)lay.n_blk, kWireBlock, 0, s.stream>>>(reinterpret_cast<const uint2 *>(s.d_wire + lay.pk_off),
                                                                      reinterpret_cast<const int *>(s.d_wire + lay.blk_off),
                                                                      reinterpret_cast<const uint4 *>(s.d_wire + lay.run_off), s.wire_runs,
                                                                      (int)n_total, s.d_a);
        } else if (n_total) {
            CK(cudaMemcpyAsync(s.d_a, up_src, up_bytes, cudaMemcpyHostToDevice, s.stream));
        }
    }
    rc = enqueue_kernels(c, s, s.stream, s.d_a, s.d_off, n_reads, n_total, s.d_f, s.d_p, prof);
    if (rc) return rc;
    s.chains = w.chains;
    s.want_fp = w.fp;
    slice_rinfo(s, n_reads);
    s.land_v = (w.chains && w.dst_v && w.dst_v_dev) ? w.dst_v : s.h_vp;
    s.v_mis = (int)((reinterpret_cast<uintptr_t>(s.land_v) >> 2) & 3);
    if (w.chains && n_total) {
        rc = enqueue_backtrack(c, s, s.stream, s.d_a, s.d_off, n_reads, s.d_f, s.d_p, prof);
        if (rc) return rc;
    }
    s.direct_out = w.fp && w.dst_pinned && w.dst_f && w.dst_p;
    s.user_f = w.dst_f; s.user_p = w.dst_p;
    {
        ProfScope ps(c, T_D2H, s.stream, prof);
        if (n_total && w.fp) {
            CK(cudaMemcpyAsync(s.direct_out ? w.dst_f : s.h_f, s.d_f, (size_t)n_total * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
            CK(cudaMemcpyAsync(s.direct_out ? w.dst_p : s.h_p, s.d_p, (size_t)n_total * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        }
        if (n_total && w.chains) {
            // only what was produced leaves the device: k_drain writes the packed chains / chain-anchor indices into mapped memory
            rc = enqueue_drain(c, s, s.stream, s.land_v == s.h_vp ? s.h_vp_dev : w.dst_v_dev);
            if (rc) return rc;
            CK(cudaMemcpyAsync(s.h_rinfo, s.d_rinfo, 4 * ((size_t)n_reads + 1) * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
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
    if (s.want_fp && !s.direct_out && s.user_f && s.user_p && s.n_total) {
        memcpy(s.user_f, s.h_f, (size_t)s.n_total * sizeof(int));
        memcpy(s.user_p, s.h_p, (size_t)s.n_total * sizeof(int));
    }
    return MM2GB_OK;
}

// after an error in the middle of a pipelined batch: let every slot run dry and mark it idle, so that the context stays usable
// and nothing writes into the caller's buffers after the call has returned
static void drain_slots(mm2gb_ctx *c)
{
    cudaSetDevice(c->device);
    for (int i = 0; i < c->n_slots; ++i) {
        Slot &s = c->slot[i];
        if (!s.busy) continue;
        cudaStreamSynchronize(s.stream);
        s.busy = false;
    }
    cudaGetLastError();
}

static Source flat_source(const mm2gb_anchor_t *a) { Source s; s.flat = a; s.flat_pinned = a && is_pinned(a); return s; }

extern "C" int mm2gb_submit(mm2gb_ctx_t *c, int slot, const mm2gb_anchor_t *a, const int64_t *off, int n_reads)
{
    if (!c || slot < 0 || slot >= c->n_slots || n_reads < 0 || !off) return fail(MM2GB_EARG, "bad argument");
    if (off[0] != 0) return fail(MM2GB_EARG, "off[0] must be 0");
    return submit_impl(c, slot, flat_source(a), (const long long *)off, n_reads, off[n_reads], Want());
}

static int gather_impl(mm2gb_ctx *c, int slot, const mm2gb_anchor_t *const *read_a, const int64_t *read_n, int n_reads, const Want &w)
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
    Source src;
    src.read_a = read_a;
    src.read_n = read_n;
    return submit_impl(c, slot, src, off.data(), n_reads, tot, w);
}

extern "C" int mm2gb_submit_gather(mm2gb_ctx_t *c, int slot, const mm2gb_anchor_t *const *read_a, const int64_t *read_n, int n_reads)
{
    return gather_impl(c, slot, read_a, read_n, n_reads, Want());
}

extern "C" int mm2gb_submit_gather_chains(mm2gb_ctx_t *c, int slot, const mm2gb_anchor_t *const *read_a, const int64_t *read_n, int n_reads)
{
    Want w;
    w.fp = false;
    w.chains = true;
    return gather_impl(c, slot, read_a, read_n, n_reads, w);
}

extern "C" int mm2gb_wait(mm2gb_ctx_t *c, int slot, const int32_t **f, const int32_t **p, const int64_t **off, mm2gb_stats_t *stats)
{
    if (!c || slot < 0 || slot >= c->n_slots) return fail(MM2GB_EARG, "bad argument");
    if (c->slot[slot].busy && !c->slot[slot].want_fp) return fail(MM2GB_ESTATE, "slot %d was submitted for chains: use mm2gb_wait_chains", slot);
    int rc = wait_impl(c, slot);
    if (rc) return rc;
    Slot &s = c->slot[slot];
    if (f) *f = s.h_f;
    if (p) *p = s.h_p;
    if (off) *off = (const int64_t *)s.h_off;
    fill_stats(c, *s.h_ctr, s.n_total, stats);
    return MM2GB_OK;
}

// Chains of a finished batch.  Every read is finished on the device (reads the shared-memory kernels cannot take go to the
// global-memory ones through the device-side list); an unfinished read here is an internal error, not a cue for a CPU path.
// After the call  s.h_nu / s.h_nb  hold the counts,  s.v_ptr[r] / s.u_ptr[r]  point at the read's chain-anchor indices / chains.
static int finish_chains(mm2gb_ctx *c, Slot &s)
{
    (void)c;
    const int n_reads = s.n_reads;
    s.u_ptr.assign((size_t)n_reads, nullptr);
    s.v_ptr.assign((size_t)n_reads, nullptr);
    s.b_total = 0;
    if (!s.n_total) { for (int r = 0; r < n_reads; ++r) s.h_nu[r] = s.h_nb[r] = s.h_bpos[r] = 0; return MM2GB_OK; }
    s.b_total = s.h_ctr->b_cur - s.v_mis;
    for (int r = 0; r < n_reads; ++r) {
        if (s.h_nu[r] < 0) return fail(MM2GB_ECUDA, "device chain extraction left read %d of the batch unfinished", r);
        s.u_ptr[(size_t)r] = reinterpret_cast<const uint64_t *>(s.h_upack) + s.h_upos[r];
        s.h_bpos[r] -= s.v_mis;     // positions relative to land_v
        s.v_ptr[(size_t)r] = s.land_v + s.h_bpos[r];
    }
    return MM2GB_OK;
}

extern "C" int mm2gb_wait_chains(mm2gb_ctx_t *c, int slot, const uint64_t *const **u, const int32_t **n_u, const int32_t *const **v,
                                 const int32_t **n_v, const int64_t **off, mm2gb_stats_t *stats)
{
    if (!c || slot < 0 || slot >= c->n_slots) return fail(MM2GB_EARG, "bad argument");
    if (c->slot[slot].busy && !c->slot[slot].chains) return fail(MM2GB_ESTATE, "slot %d was not submitted for chains", slot);
    int rc = wait_impl(c, slot);
    if (rc) return rc;
    Slot &s = c->slot[slot];
    rc = finish_chains(c, s);
    if (rc) return rc;
    if (u) *u = s.u_ptr.data();
    if (n_u) *n_u = s.h_nu;
    if (v) *v = s.v_ptr.data();
    if (n_v) *n_v = s.h_nb;
    if (off) *off = (const int64_t *)s.h_off;
    fill_stats(c, *s.h_ctr, s.n_total, stats);
    return MM2GB_OK;
}

extern "C" int mm2gb_slot_busy(mm2gb_ctx_t *c, int slot)
{
    if (!c || slot < 0 || slot >= c->n_slots) return 0;
    return c->slot[slot].busy ? 1 : 0;
}

extern "C" void mm2gb_gather_anchors(const mm2gb_anchor_t *a, const int32_t *v, int64_t n, mm2gb_anchor_t *b)
{
    wire_gather(a, v, n, b);
}

// The packed wire format on the host alone (no device involved): pack a batch into `buf` exactly as the upload path does, and
// the inverse by the rule k_expand applies on the device (block index -> bracketed run search).  Returns the bytes to upload,
// or -1 when the run list does not fit `cap` bytes (the upload path then sends the anchors raw).
extern "C" int64_t mm2gb_wire_pack(const mm2gb_anchor_t *a, const int64_t *off, int n_reads, void *buf, size_t cap, int32_t *n_runs)
{
    if (!off || n_reads < 0 || !buf || (reinterpret_cast<uintptr_t>(buf) & 31)) return -1;
    const int64_t n = off[n_reads];
    const WireLayout L = wire_layout(n, cap);
    if (L.run_cap < 1 || L.run_off > cap) return -1;
    WirePacker pk;
    pk.begin(buf, L);
    for (int r = 0; r < n_reads; ++r)
        if (!pk.add(a + off[r], off[r + 1] - off[r])) return -1;
    if (n_runs) *n_runs = pk.n_runs();
    return (int64_t)pk.finish();
}

extern "C" int mm2gb_wire_unpack(const void *buf, size_t cap, int64_t n, int32_t n_runs, mm2gb_anchor_t *out)
{
    if (!buf || n < 0 || (n > 0 && (!out || n_runs < 1))) return MM2GB_EARG;
    const WireLayout L = wire_layout(n, cap);
    const char *base = static_cast<const char *>(buf);
    const uint64_t *pk = reinterpret_cast<const uint64_t *>(base + L.pk_off);
    const int32_t *blk = reinterpret_cast<const int32_t *>(base + L.blk_off);
    const WireRun *runs = reinterpret_cast<const WireRun *>(base + L.run_off);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t b = i / kWireBlock;
        int lo = blk[b], hi = std::min(blk[b + 1], n_runs - 1);
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (runs[mid].start <= i) lo = mid; else hi = mid - 1;
        }
        out[i].x = (uint64_t)(uint32_t)pk[i] | ((uint64_t)runs[lo].x_hi << 32);
        out[i].y = (pk[i] >> 32) | ((uint64_t)runs[lo].y_hi << 32);
    }
    return MM2GB_OK;
}

extern "C" int64_t mm2gb_last_upload_bytes(mm2gb_ctx_t *c, int slot)
{
    if (!c || slot < 0 || slot >= c->n_slots) return -1;
    return (int64_t)c->slot[slot].up_bytes;
}

// Shared driver of the host-buffer entry points: split the reads into chunks, run them round-robin through the slots
// (upload / kernels / download of consecutive chunks overlap) and call on_done(slot, r0, r1) as each chunk's results land.
template <class Done>
static int run_chunked(mm2gb_ctx *c, const mm2gb_anchor_t *a, const int64_t *off, int n_reads, int32_t *f, int32_t *p, int32_t *v,
                       int *v_dev, bool chains, mm2gb_stats_t *stats, Done on_done)
{
    for (int i = 0; i < c->n_slots; ++i)
        if (c->slot[i].busy) return fail(MM2GB_ESTATE, "slot %d is busy", i);
    for (int r = 0; r < n_reads; ++r)   // before anything is in flight
        if (off[r + 1] - off[r] > (long long)c->max_anchors)
            return fail(MM2GB_ECAP, "read %d has %lld anchors, capacity is %zu", r, (long long)(off[r + 1] - off[r]), c->max_anchors);
    const Source src_all = flat_source(a);
    Want w;
    w.fp = f && p;
    w.dst_pinned = w.fp && is_pinned(f) && is_pinned(p);
    w.chains = chains;
    const long long total = off[n_reads] - off[0];
    // chunks: enough of them to overlap upload / kernels / download, few enough that a chunk still fills the GPU
    long long target = std::max<long long>(1 << 20, total / std::max(8, c->n_slots + 2) + 1);
    target = std::min<long long>(target, (long long)c->max_anchors);
    if (const char *e = getenv("MM2GB_CHUNK")) target = std::min<long long>(std::max(1LL, atoll(e)), (long long)c->max_anchors);
    mm2gb_stats_t acc;
    memset(&acc, 0, sizeof(acc));
    c->up_bytes_batch = 0;
    std::vector<long long> rel;
    int slot_r0[kMaxSlots] = {0}, slot_r1[kMaxSlots] = {0};
    int r0 = 0, chunk = 0, rc = MM2GB_OK;
    auto reap = [&](int si) -> int {
        int rc2 = wait_impl(c, si);
        if (rc2) return rc2;
        mm2gb_stats_t st;
        fill_stats(c, *c->slot[si].h_ctr, c->slot[si].n_total, &st);
        acc.n_anchors += st.n_anchors; acc.n_pairs += st.n_pairs; acc.n_units += st.n_units;
        acc.n_units_exact += st.n_units_exact; acc.n_long += st.n_long; acc.general_path |= st.general_path;
        return on_done(si, slot_r0[si], slot_r1[si]);
    };
    while (r0 < n_reads) {
        int r1 = r0;
        long long cnt = 0;
        // small chunks at both ends of the batch shorten the parts of the pipeline that cannot overlap (the first upload,
        // the last kernels + download); full-size chunks in between keep the kernels efficient
        const long long remaining = off[n_reads] - off[r0];
        const long long ramp = (target / 4) << std::min(chunk, 2);
        const long long this_target = std::max<long long>(1 << 18, std::min(std::min(target, ramp), std::max(target / 4, remaining * 2 / 5)));
        while (r1 < n_reads && r1 - r0 < c->max_reads) {
            const long long nr = off[r1 + 1] - off[r1];
            if (cnt && cnt + nr > this_target) break;
            cnt += nr; ++r1;
        }
        const int si = chunk % c->n_slots;
        if (c->slot[si].busy && (rc = reap(si))) break;
        rel.resize((size_t)(r1 - r0) + 1);
        for (int r = r0; r <= r1; ++r) rel[(size_t)(r - r0)] = off[r] - off[r0];
        w.dst_f = w.fp ? f + off[r0] : nullptr;
        w.dst_p = w.fp ? p + off[r0] : nullptr;
        w.dst_v = (v && v_dev) ? v + off[r0] : nullptr;      // a chunk's packed results land where its anchors start
        w.dst_v_dev = (v && v_dev) ? v_dev + off[r0] : nullptr;
        Source src = src_all;
        src.flat = a + off[r0];
        rc = submit_impl(c, si, src, rel.data(), r1 - r0, cnt, w);
        if (rc) break;
        c->up_bytes_batch += (long long)c->slot[si].up_bytes;
        slot_r0[si] = r0; slot_r1[si] = r1;
        r0 = r1; ++chunk;
    }
    // drain in submission order
    for (int k = 0; k < c->n_slots && !rc; ++k) {
        const int si = (chunk + k) % c->n_slots;
        if (c->slot[si].busy) rc = reap(si);
    }
    if (rc) { drain_slots(c); return rc; }
    if (stats) *stats = acc;
    if (c->timeline) timeline_dump(c);
    return MM2GB_OK;
}

extern "C" int mm2gb_chain_dp_host(mm2gb_ctx_t *c, const mm2gb_anchor_t *a, const int64_t *off, int n_reads, int32_t *f, int32_t *p,
                                   mm2gb_stats_t *stats)
{
    if (!c || !off || n_reads < 0 || (!a && off[n_reads] > 0) || ((!f || !p) && off[n_reads] > 0)) return fail(MM2GB_EARG, "bad argument");
    if (off[0] != 0) return fail(MM2GB_EARG, "off[0] must be 0");
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n_reads == 0) return MM2GB_OK;
    return run_chunked(c, a, off, n_reads, f, p, nullptr, nullptr, false, stats, [](int, int, int) { return MM2GB_OK; });
}

// Whole mg_lchain_dp (lchain.c:148-217) for a batch, host-stage variant (a diagnostic: how the reference arranges the work,
// gpu/plchain.cu:99-150): device DP, then backtracking + compaction on `n_threads` host worker threads that start on a chunk's
// reads as soon as its f/p have landed, while later chunks are still on the GPU.
static int chain_host_hoststage(mm2gb_ctx_t *c, const mm2gb_anchor_t *a, const int64_t *off, int n_reads, int32_t *f, int32_t *p,
                                uint64_t *u, int32_t *n_u, mm2gb_anchor_t *b, int64_t *n_b, int n_threads, mm2gb_stats_t *stats)
{
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
    int rc = run_chunked(c, a, off, n_reads, f, p, nullptr, nullptr, false, stats, [&](int, int, int r1) {
        { std::lock_guard<std::mutex> lk(mu); ready = r1; }
        cv.notify_all();
        return MM2GB_OK;
    });
    {
        std::lock_guard<std::mutex> lk(mu);
        if (rc) abort_all = true; else ready = n_reads;
    }
    cv.notify_all();
    for (auto &t : pool) t.join();
    return rc;
}

// Device variant (default): chain extraction + compaction run on the GPU right behind the DP kernels (k_bt_sort / k_bt_walk)
// and only the chains and the indices of their anchors leave the device.  f / p are downloaded only if the caller passes
// buffers for them.  Output layout of mm2gb_chain_host: read r's compacted anchors at b[off[r] ..], gathered here from `a`.
static int chain_host_device(mm2gb_ctx_t *c, const mm2gb_anchor_t *a, const int64_t *off, int n_reads, int32_t *f, int32_t *p,
                             uint64_t *u, int32_t *n_u, mm2gb_anchor_t *b, int64_t *n_b, mm2gb_stats_t *stats)
{
    return run_chunked(c, a, off, n_reads, f, p, nullptr, nullptr, true, stats, [&](int si, int r0, int r1) {
        Slot &s = c->slot[si];
        int rc = finish_chains(c, s);
        if (rc) return rc;
        for (int r = r0; r < r1; ++r) {
            const int k = r - r0;
            n_u[r] = s.h_nu[k];
            n_b[r] = s.h_nb[k];
            if (s.h_nu[k] > 0) memcpy(u + off[r], s.u_ptr[(size_t)k], (size_t)s.h_nu[k] * sizeof(uint64_t));
            if (s.h_nb[k] > 0) wire_gather(a + off[r], s.v_ptr[(size_t)k], s.h_nb[k], b + off[r]);
        }
        return MM2GB_OK;
    });
}

// device view of a caller's buffer if it is mapped pinned memory, else nullptr
static int *mapped_view(int32_t *v)
{
    if (!v || !is_pinned(v)) return nullptr;
    void *d = nullptr;
    if (cudaHostGetDevicePointer(&d, v, 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return reinterpret_cast<int *>(d);
}

// Index output: the compacted anchors of read r are a[off[r] + v[v_pos[r] + k]], k < n_v[r]; the reads of one chunk are packed
// behind each other (in no particular order) starting at index off[first read of the chunk] of `v`.  With `v` in pinned memory
// the device writes them there directly and the host only copies the (few) chains.
extern "C" int mm2gb_chain_host_index(mm2gb_ctx_t *c, const mm2gb_anchor_t *a, const int64_t *off, int n_reads, uint64_t *u, int32_t *n_u,
                                      int32_t *v, int64_t *v_pos, int64_t *n_v, mm2gb_stats_t *stats)
{
    if (!c || !off || n_reads < 0 || (!a && off[n_reads] > 0) || !n_u || !n_v || !v_pos) return fail(MM2GB_EARG, "bad argument");
    if (off[n_reads] > 0 && (!u || !v)) return fail(MM2GB_EARG, "bad argument");
    if (off[0] != 0) return fail(MM2GB_EARG, "off[0] must be 0");
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n_reads == 0) return MM2GB_OK;
    CK(cudaSetDevice(c->device));
    int *v_dev = mapped_view(v);
    return run_chunked(c, a, off, n_reads, nullptr, nullptr, v, v_dev, true, stats, [&](int si, int r0, int r1) {
        Slot &s = c->slot[si];
        int rc = finish_chains(c, s);
        if (rc) return rc;
        int32_t *dst = v + off[r0];
        if (s.land_v != dst && s.b_total > 0) memcpy(dst, s.land_v, (size_t)s.b_total * sizeof(int32_t));
        for (int r = r0; r < r1; ++r) {
            const int k = r - r0;
            n_u[r] = s.h_nu[k];
            n_v[r] = s.h_nb[k];
            v_pos[r] = off[r0] + (s.h_nb[k] > 0 ? s.h_bpos[k] : 0);
            if (s.h_nu[k] > 0) memcpy(u + off[r], s.u_ptr[(size_t)k], (size_t)s.h_nu[k] * sizeof(uint64_t));
        }
        return MM2GB_OK;
    });
}

extern "C" int64_t mm2gb_last_batch_upload_bytes(mm2gb_ctx_t *c) { return c ? (int64_t)c->up_bytes_batch : -1; }

extern "C" int mm2gb_chain_host(mm2gb_ctx_t *c, const mm2gb_anchor_t *a, const int64_t *off, int n_reads, int32_t *f, int32_t *p,
                                uint64_t *u, int32_t *n_u, mm2gb_anchor_t *b, int64_t *n_b, int n_threads, mm2gb_stats_t *stats)
{
    if (!c || !off || n_reads < 0 || (!a && off[n_reads] > 0) || !n_u || !n_b) return fail(MM2GB_EARG, "bad argument");
    if (off[n_reads] > 0 && (!u || !b)) return fail(MM2GB_EARG, "bad argument");
    if (off[0] != 0) return fail(MM2GB_EARG, "off[0] must be 0");
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n_reads == 0) return MM2GB_OK;
    if (n_threads >= 1) { // host-stage variant needs f / p
        if (off[n_reads] > 0 && (!f || !p)) return fail(MM2GB_EARG, "the host-stage variant needs f and p buffers");
        return chain_host_hoststage(c, a, off, n_reads, f, p, u, n_u, b, n_b, n_threads, stats);
    }
    return chain_host_device(c, a, off, n_reads, f, p, u, n_u, b, n_b, stats);
}

// The host stage alone for a batch whose f/p are already in host memory (diagnostic; what the reference runs behind its kernels).
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

// Chain extraction + compaction on the device for caller-supplied f / p (same kernel mm2gb_chain_host runs behind the DP):
// lets the stage be checked on arbitrary score / predecessor arrays.  Synchronous, uses slot 0.
extern "C" int mm2gb_backtrack_device(mm2gb_ctx_t *c, const mm2gb_anchor_t *a, const int64_t *off, int n_reads, const int32_t *f,
                                      const int32_t *p, uint64_t *u, int32_t *n_u, mm2gb_anchor_t *b, int64_t *n_b, int32_t *n_declined)
{
    if (!c || !off || n_reads < 0 || !n_u || !n_b) return fail(MM2GB_EARG, "bad argument");
    if (off[0] != 0) return fail(MM2GB_EARG, "off[0] must be 0");
    const long long n_total = off[n_reads];
    if (n_total > 0 && (!a || !f || !p || !u || !b)) return fail(MM2GB_EARG, "bad argument");
    if ((size_t)n_total > c->max_anchors) return fail(MM2GB_ECAP, "batch of %lld anchors exceeds capacity %zu", n_total, c->max_anchors);
    if (n_reads > c->max_reads) return fail(MM2GB_ECAP, "batch of %d reads exceeds capacity %d", n_reads, c->max_reads);
    Slot &s = c->slot[0];
    if (s.busy) return fail(MM2GB_ESTATE, "slot 0 is busy");
    if (!c->host_io || !c->chains_ok || !c->fp_staging) return fail(MM2GB_ESTATE, "context was created without host staging / chain-extraction buffers");
    CK(cudaSetDevice(c->device));
    if (n_declined) *n_declined = 0;
    memcpy(s.h_off, off, ((size_t)n_reads + 1) * sizeof(long long));
    s.n_reads = n_reads;
    s.n_total = n_total;
    if (n_total) {
        memcpy(s.h_a, a, (size_t)n_total * sizeof(mm2gb_anchor_t));
        memcpy(s.h_f, f, (size_t)n_total * sizeof(int));
        memcpy(s.h_p, p, (size_t)n_total * sizeof(int));
        CK(cudaMemcpyAsync(s.d_off, s.h_off, ((size_t)n_reads + 1) * sizeof(long long), cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemcpyAsync(s.d_a, s.h_a, (size_t)n_total * sizeof(uint4), cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemcpyAsync(s.d_f, s.h_f, (size_t)n_total * sizeof(int), cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemcpyAsync(s.d_p, s.h_p, (size_t)n_total * sizeof(int), cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemsetAsync(s.d_ctr, 0, sizeof(Counters), s.stream));
        s.land_v = s.h_vp;
        s.v_mis = 0;
        int rc = prepare_backtrack(s, s.stream, s.h_off, n_reads);
        if (rc) return rc;
        rc = enqueue_backtrack(c, s, s.stream, s.d_a, s.d_off, n_reads, s.d_f, s.d_p, false);
        if (rc) return rc;
        rc = enqueue_drain(c, s, s.stream, s.h_vp_dev);
        if (rc) return rc;
        CK(cudaMemcpyAsync(s.h_rinfo, s.d_rinfo, 4 * ((size_t)n_reads + 1) * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaMemcpyAsync(s.h_ctr, s.d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        if (n_declined) for (int r = 0; r < n_reads; ++r) *n_declined += s.h_nu[r] < 0 ? 1 : 0;   // must stay 0
    } else {
        slice_rinfo(s, n_reads);
    }
    int rc = finish_chains(c, s);
    if (rc) return rc;
    for (int r = 0; r < n_reads; ++r) {
        n_u[r] = s.h_nu[r];
        n_b[r] = s.h_nb[r];
        if (s.h_nu[r] > 0) memcpy(u + off[r], s.u_ptr[(size_t)r], (size_t)s.h_nu[r] * sizeof(uint64_t));
        if (s.h_nb[r] > 0) wire_gather(a + off[r], s.v_ptr[(size_t)r], s.h_nb[r], b + off[r]);
    }
    return MM2GB_OK;
}

static int chain_dp_device_slot(mm2gb_ctx_t *c, int si, const void *d_a, const void *d_off, int n_reads, int64_t n_total, void *d_f, void *d_p)
{
    if (!c || si < 0 || si >= c->n_slots || n_reads < 0 || n_total < 0 || !d_off) return fail(MM2GB_EARG, "bad argument");
    if ((size_t)n_total > c->max_anchors) return fail(MM2GB_ECAP, "batch of %lld anchors exceeds capacity %zu", (long long)n_total, c->max_anchors);
    if (n_reads > c->max_reads) return fail(MM2GB_ECAP, "batch of %d reads exceeds capacity %d", n_reads, c->max_reads);
    CK(cudaSetDevice(c->device));
    Slot &s = c->slot[si];
    int rc = enqueue_kernels(c, s, s.stream, (const uint4 *)d_a, (const long long *)d_off, n_reads, n_total, (int *)d_f, (int *)d_p, c->profile && si == 0);
    if (rc) return rc;
    CK(cudaMemcpyAsync(s.h_ctr, s.d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, s.stream));
    s.n_total = n_total;
    return MM2GB_OK;
}

extern "C" int mm2gb_chain_dp_device(mm2gb_ctx_t *c, const void *d_a, const void *d_off, int n_reads, int64_t n_total, void *d_f, void *d_p)
{
    return chain_dp_device_slot(c, 0, d_a, d_off, n_reads, n_total, d_f, d_p);
}

// device-resident DP + chain extraction on the stream and scratch of slot `slot`: as mm2gb_chain_dp_device, then k_bt_sort* /
// k_bt_walk* into the slot's output buffers (`off` = host copy of the offsets, needed to bin the reads by size).  Batches
// enqueued on different slots are independent: the latency-bound chain extraction of one overlaps the score kernels of the next.
extern "C" int mm2gb_chain_device_slot(mm2gb_ctx_t *c, int slot, const void *d_a, const void *d_off, const int64_t *off, int n_reads,
                                       int64_t n_total, void *d_f, void *d_p)
{
    if (!off) return fail(MM2GB_EARG, "bad argument");
    if (c && !c->chains_ok) return fail(MM2GB_ESTATE, "context was created with MM2GB_CTX_NO_CHAINS");
    int rc = chain_dp_device_slot(c, slot, d_a, d_off, n_reads, n_total, d_f, d_p);
    if (rc || n_total == 0) return rc;
    Slot &s = c->slot[slot];
    s.v_mis = 0;
    rc = prepare_backtrack(s, s.stream, (const long long *)off, n_reads);
    if (rc) return rc;
    return enqueue_backtrack(c, s, s.stream, (const uint4 *)d_a, (const long long *)d_off, n_reads, (const int *)d_f, (const int *)d_p,
                             c->profile && slot == 0);
}

extern "C" int mm2gb_chain_device(mm2gb_ctx_t *c, const void *d_a, const void *d_off, const int64_t *off, int n_reads, int64_t n_total,
                                  void *d_f, void *d_p)
{
    return mm2gb_chain_device_slot(c, 0, d_a, d_off, off, n_reads, n_total, d_f, d_p);
}

// Results of a batch enqueued by mm2gb_chain_device_slot -> host.  The producer of the anchors lives on the device (mm2gb_seed), so
// the compacted anchors themselves cross PCIe (16 B each, gathered by k_drain_anchors) next to the packed chains; both land in
// caller-supplied pinned memory, the per-read counts / positions in the slot's own.  Works on MM2GB_CTX_DEVICE_ONLY contexts.
extern "C" int mm2gb_chain_device_fetch(mm2gb_ctx_t *c, int slot, const void *d_a, const void *d_off, int n_reads, int64_t n_total,
                                        mm2gb_anchor_t *b_pinned, uint64_t *u_pinned)
{
    if (!c || slot < 0 || slot >= c->n_slots || n_reads < 0 || !b_pinned || !u_pinned) return fail(MM2GB_EARG, "bad argument");
    if (!c->chains_ok) return fail(MM2GB_ESTATE, "context was created with MM2GB_CTX_NO_CHAINS");
    CK(cudaSetDevice(c->device));
    Slot &s = c->slot[slot];
    if (s.busy) return fail(MM2GB_ESTATE, "slot %d is busy", slot);
    void *b_dev = nullptr, *u_dev = nullptr;
    if (cudaHostGetDevicePointer(&b_dev, b_pinned, 0) != cudaSuccess || cudaHostGetDevicePointer(&u_dev, u_pinned, 0) != cudaSuccess) {
        cudaGetLastError();
        return fail(MM2GB_EARG, "result buffers must be pinned (mapped) host memory");
    }
    slice_rinfo(s, n_reads);
    if (n_total) {
        k_drain_anchors<<<c->drain_blocks * 2, 256, 0, s.stream>>>((const uint4 *)d_a, (const long long *)d_off, s.d_vp, s.d_nb, s.d_bpos, n_reads,
                                                                   (uint4 *)b_dev, s.d_upack, (unsigned long long *)u_dev, s.d_ctr);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(s.h_rinfo, s.d_rinfo, 4 * ((size_t)n_reads + 1) * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    }
    CK(cudaMemcpyAsync(s.h_ctr, s.d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaEventRecord(s.done, s.stream));
    s.busy = true;
    s.chains = true; s.want_fp = false; s.direct_out = false; s.user_f = s.user_p = nullptr;
    s.n_reads = n_reads;
    s.n_total = n_total;
    return MM2GB_OK;
}

// wait for mm2gb_chain_device_fetch: per read r  n_u[r] chains at u_pinned[u_pos[r] ..], n_b[r] anchors at b_pinned[b_pos[r] ..]
extern "C" int mm2gb_chain_device_results(mm2gb_ctx_t *c, int slot, const int32_t **n_u, const int32_t **u_pos, const int32_t **n_b,
                                          const int32_t **b_pos, int64_t *n_chains, int64_t *n_chain_anchors, mm2gb_stats_t *stats)
{
    if (!c || slot < 0 || slot >= c->n_slots) return fail(MM2GB_EARG, "bad argument");
    int rc = wait_impl(c, slot);
    if (rc) return rc;
    Slot &s = c->slot[slot];
    if (!s.n_total) {
        for (int r = 0; r < s.n_reads; ++r) s.h_nu[r] = s.h_nb[r] = s.h_upos[r] = s.h_bpos[r] = 0;
    } else {
        for (int r = 0; r < s.n_reads; ++r)
            if (s.h_nu[r] < 0) return fail(MM2GB_ECUDA, "device chain extraction left read %d of the batch unfinished", r);
    }
    if (n_u) *n_u = s.h_nu;
    if (u_pos) *u_pos = s.h_upos;
    if (n_b) *n_b = s.h_nb;
    if (b_pos) *b_pos = s.h_bpos;
    if (n_chains) *n_chains = s.n_total ? s.h_ctr->u_cur : 0;
    if (n_chain_anchors) *n_chain_anchors = s.n_total ? s.h_ctr->b_cur : 0;
    fill_stats(c, *s.h_ctr, s.n_total, stats);
    return MM2GB_OK;
}

// Diagnostic: device -> pinned host of n chain-anchor indices (4 B each) from slot 0's buffers, by k_drain with `blocks` CTAs and
// by the copy engine, each optionally with a host -> device copy of the same number of bytes running on a second stream (PCIe
// is full duplex).  ms[0] = k_drain alone, ms[1] = cudaMemcpyAsync alone, ms[2] = k_drain + H2D, ms[3] = cudaMemcpyAsync + H2D,
// ms[4] = H2D alone.
extern "C" int mm2gb_debug_drain(mm2gb_ctx_t *c, int64_t n, int blocks, float ms[5])
{
    if (!c || !ms || n <= 0 || (size_t)n > c->max_anchors || !c->host_io || !c->chains_ok || blocks < 1) return fail(MM2GB_EARG, "bad argument");
    CK(cudaSetDevice(c->device));
    Slot &s = c->slot[0];
    cudaStream_t s2;
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    Counters hc;
    memset(&hc, 0, sizeof(hc));
    hc.b_cur = (int)n;
    CK(cudaMemcpy(s.d_ctr, &hc, sizeof(hc), cudaMemcpyHostToDevice));
    for (int mode = 0; mode < 5; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, s.stream));
            CK(cudaStreamWaitEvent(s2, e0, 0));
            if (mode == 0 || mode == 2) k_drain<<<blocks, kDrainThreads, 0, s.stream>>>(s.d_vp, s.h_vp_dev, 0, s.d_upack, s.h_upack_dev, s.d_ctr);
            if (mode == 1 || mode == 3) CK(cudaMemcpyAsync(s.h_vp, s.d_vp, (size_t)n * 4, cudaMemcpyDeviceToHost, s.stream));
            if (mode >= 2) CK(cudaMemcpyAsync(s.d_a, s.h_a, (size_t)n * 4, cudaMemcpyHostToDevice, s2));
            CK(cudaEventRecord(e2, s2));
            CK(cudaStreamWaitEvent(s.stream, e2, 0));
            CK(cudaEventRecord(e1, s.stream));
            CK(cudaEventSynchronize(e1));
            float t = 0;
            CK(cudaEventElapsedTime(&t, e0, e1));
            if (rep) best = std::min(best, t);
        }
        ms[mode] = best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    cudaStreamDestroy(s2);
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
