/*
 * mm2gb_chain.h -- C ABI of the B200-native anchor-chaining library (libmm2gb_chain.so).
 *
 * Two layers, both plain C (pointers and sizes only, no CUDA / torch types):
 *
 *  (1) The DROP-IN boundary: the four entry points minimap2's host driver already calls
 *      (reference gpu/plutils.h:98-104; callers main.c:445,466 and map.c:1026,1069).  They are declared in
 *      mm2gb_plchain.h (which needs the reference's minimap.h types) and implemented in
 *      mm2-gb_b200/csrc/plchain_dropin.cpp on top of layer (2).
 *
 *  (2) The CORE chaining API below: a batch of reads' seeded anchors in, chain scores f[] and
 *      predecessors p[] out (reference lchain.c:148-207 at max-chain-skip = infinity), plus the
 *      backtracked chains u[] / compacted anchors (lchain.c:27-111).  It replaces what the reference
 *      spreads over gpu/plchain.cu (orchestration), gpu/plmem.cu (buffers + H2D/D2H), gpu/plrange.cu
 *      (window/segment kernel) and gpu/plscore.cu (score kernels).  Tests and bench.py bind it with ctypes.
 *
 * Every function returns 0 on success and a negative MM2GB_E* code on failure; mm2gb_last_error() gives the
 * message.  Nothing here falls back to the CPU: without a CUDA device the calls fail.
 */
#ifndef MM2GB_CHAIN_H
#define MM2GB_CHAIN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MM2GB_OK 0
#define MM2GB_ECUDA (-1)    /* a CUDA runtime call failed */
#define MM2GB_EARG (-2)     /* bad argument */
#define MM2GB_ECAP (-3)     /* batch larger than the context capacity */
#define MM2GB_ESTATE (-4)   /* slot busy / idle misuse */

/* minimap.h:72 mm128_t.  x = rev<<63 | rid<<32 | rpos ; y = seg_id<<48 | flags<<40 | q_span<<32 | qpos */
typedef struct { uint64_t x, y; } mm2gb_anchor_t;

/* gpu/plutils.h:33-37 Misc -- identical layout (44 bytes), filled by build_misc (map.c:393-426) */
typedef struct {
    int max_iter, max_dist_x, max_dist_y, max_skip, bw, min_cnt, min_score, is_cdna, n_seg;
    float chn_pen_gap, chn_pen_skip;
} mm2gb_misc_t;

typedef struct mm2gb_ctx mm2gb_ctx_t;

/* per-batch counters filled by the device (for the pairs/s metric: pairs = sum_i (i - st_i), lchain.c:177) */
typedef struct {
    int64_t n_anchors;
    int64_t n_pairs;       /* evaluated anchor pairs, equal to the reference's n_iter at max_skip = inf */
    int32_t n_units;       /* independent work units the range kernel cut the batch into */
    int32_t n_units_exact; /* units with a max_iter-clipped window, scored by the exact max_ii path (lchain.c:189-205) */
    int32_t n_long;        /* units scored by the block-cooperative long kernel */
    int32_t general_path;  /* 1 if the float-penalty / multi-segment score path was used for this batch */
} mm2gb_stats_t;

const char *mm2gb_last_error(void);
int mm2gb_device_count(void);
/* free / total memory of a device in bytes (for sizing contexts: ~90 B of device and ~28 B of pinned memory per anchor and slot) */
int mm2gb_device_memory(int device, size_t *free_bytes, size_t *total_bytes);
/* total memory only, without creating a context on the device (bringing a GPU up can take seconds) */
int mm2gb_device_total_memory(int device, size_t *total_bytes);
/* Confines the CALLING thread to the CPUs Linux lists as local to `device` (the NUMA node of its PCI function), so that the
 * host passes of that thread and the pages it touches stay on the GPU's node.  Returns 1 if the affinity was changed, 0 where
 * the box exposes no topology or with MM2GB_NUMA=0.  The staging of a context is placed that way regardless of the caller. */
int mm2gb_bind_thread_near_device(int device);

/* One context per (host thread, GPU).  `max_anchors` / `max_reads` bound one batch; `n_slots` (1..8) is the
 * number of batches that may be in flight (each slot owns a stream, pinned staging and device buffers).
 * Replaces plmem_stream_initialize + plrange/plscore_upload_misc (gpu/plmem.cu:558-624, plchain.cu:470-474). */
int mm2gb_ctx_create(mm2gb_ctx_t **ctx, int device, size_t max_anchors, int max_reads, int n_slots, const mm2gb_misc_t *misc);
/* The same with flags: DEVICE_ONLY = no pinned staging and no slot-owned anchor / f / p buffers (only the device-resident
 * entry points work; for anchor arrays that already live in HBM, e.g. the 500 M-anchor chaining-only benchmark);
 * NO_CHAINS = no chain-extraction buffers (only the DP entry points work); NO_FP_STAGING = no pinned f / p staging. */
#define MM2GB_CTX_DEVICE_ONLY 1u
#define MM2GB_CTX_NO_CHAINS 2u
#define MM2GB_CTX_NO_FP_STAGING 4u   /* no pinned f / p staging: only the chain entry points work (what the drop-in needs) */
int mm2gb_ctx_create_ex(mm2gb_ctx_t **ctx, int device, size_t max_anchors, int max_reads, int n_slots, const mm2gb_misc_t *misc,
                        unsigned flags);
void mm2gb_ctx_destroy(mm2gb_ctx_t *ctx);
int mm2gb_ctx_set_misc(mm2gb_ctx_t *ctx, const mm2gb_misc_t *misc);

/* ---- host-buffer path (what the drop-in uses; copies are part of the call) --------------------------- */

/* Synchronous: chain reads r = 0..n_reads-1 whose anchors are a[off[r] .. off[r+1]) (x-sorted per read, as
 * collect_seed_hits leaves them, map.c:329).  Outputs f[off[n_reads]] and p[off[n_reads]]; p is the index of the
 * predecessor INSIDE the read, -1 for none (lchain.c:202).  stats may be NULL. */
int mm2gb_chain_dp_host(mm2gb_ctx_t *ctx, const mm2gb_anchor_t *a, const int64_t *off, int n_reads, int32_t *f, int32_t *p,
                        mm2gb_stats_t *stats);

/* The whole mg_lchain_dp (lchain.c:148-217) for a batch, pipelined chunk by chunk through the slots.  Per read r: chains
 * u[off[r] .. off[r]+n_u[r]) (score<<32 | count, ordered by chain start) and compacted anchors b[off[r] .. off[r]+n_b[r]);
 * the rest of b[off[r] .. off[r+1]) is scratch.  f/p (size off[n_reads]) receive the DP arrays; they may be NULL when
 * n_threads <= 0.
 *   n_threads <= 0 : chain extraction + compaction (lchain.c:27-111) run on the DEVICE right behind the DP kernels and only
 *                    the chains and the indices of their anchors leave the device (b is gathered from `a` on the host).  Reads
 *                    of any size are handled on the device: up to 8192 anchors in shared memory, longer ones (or scores
 *                    >= 2^19) by the mid / global-memory kernels.  This is the product path.
 *   n_threads >= 1 : DIAGNOSTIC ONLY -- f/p are downloaded and that stage runs on n_threads host threads (what the reference
 *                    does on one thread, gpu/plchain.cu:99-150); used by bench.py to break the end-to-end time down. */
int mm2gb_chain_host(mm2gb_ctx_t *ctx, const mm2gb_anchor_t *a, const int64_t *off, int n_reads, int32_t *f, int32_t *p,
                     uint64_t *u, int32_t *n_u, mm2gb_anchor_t *b, int64_t *n_b, int n_threads, mm2gb_stats_t *stats);

/* The same with the compacted anchors returned as INDICES -- the wire format of the results: compact_a (lchain.c:78-111) is a
 * gather of anchors the caller still holds, so 4 instead of 16 bytes per chain anchor cross PCIe.  Read r's compacted anchors are
 * a[off[r] + v[v_pos[r] + k]], k < n_v[r]; `v` needs room for off[n_reads] entries.  With `v` in pinned (mapped) memory the device
 * writes the indices there itself (k_drain) -- exactly the bytes produced, no host copy of them at all.
 * mm2gb_gather_anchors materialises one read's a'[]: b[k] = a[v[k]]. */
int mm2gb_chain_host_index(mm2gb_ctx_t *ctx, const mm2gb_anchor_t *a, const int64_t *off, int n_reads, uint64_t *u, int32_t *n_u,
                           int32_t *v, int64_t *v_pos, int64_t *n_v, mm2gb_stats_t *stats);
void mm2gb_gather_anchors(const mm2gb_anchor_t *a, const int32_t *v, int64_t n, mm2gb_anchor_t *b);
/* Anchor bytes that crossed PCIe host -> device for the last pipelined batch / for the batch of a slot.  Anchors in pinned caller
 * memory are DMA'd as they are (16 B each, no host pass); anchors in pageable memory go through one gather pass that writes the
 * packed wire format (8 B each + one 16-byte record per run of equal high words, csrc/wire.h) into pinned staging, and k_expand
 * rebuilds the 16-byte anchors on the device.  MM2GB_WIRE=raw|packed|auto overrides the choice. */
int64_t mm2gb_last_batch_upload_bytes(mm2gb_ctx_t *ctx);
/* The packed wire format on the host alone (tests; a producer that wants to write it directly): pack the reads a[off[r] .. off[r+1])
 * into `buf` (32-byte aligned, `cap` bytes) -> bytes to upload, or -1 if the run list does not fit; and the inverse, by the rule
 * k_expand applies on the device. */
int64_t mm2gb_wire_pack(const mm2gb_anchor_t *a, const int64_t *off, int n_reads, void *buf, size_t cap, int32_t *n_runs);
int mm2gb_wire_unpack(const void *buf, size_t cap, int64_t n, int32_t n_runs, mm2gb_anchor_t *out);
int64_t mm2gb_last_upload_bytes(mm2gb_ctx_t *ctx, int slot);

/* Asynchronous pair.  submit: stage anchors into the slot's pinned buffer, enqueue H2D + kernels + D2H on the slot's
 * stream and return.  wait: block until the slot is done and expose the pinned result arrays (valid until the slot is
 * submitted again).  gather variant takes one pointer per read (chain_read_t.a of each read, plutils.h:64). */
int mm2gb_submit(mm2gb_ctx_t *ctx, int slot, const mm2gb_anchor_t *a, const int64_t *off, int n_reads);
int mm2gb_submit_gather(mm2gb_ctx_t *ctx, int slot, const mm2gb_anchor_t *const *read_a, const int64_t *read_n, int n_reads);
int mm2gb_wait(mm2gb_ctx_t *ctx, int slot, const int32_t **f, const int32_t **p, const int64_t **off, mm2gb_stats_t *stats);
/* The same pair with chain extraction on the device (what the drop-in uses; replaces the host loop of
 * gpu/plchain.cu:99-150).  wait_chains: per read r  n_u[r] chains at u[r][0 .. n_u[r]) and the indices (inside the read) of its
 * n_v[r] compacted anchors at v[r][0 .. n_v[r]) -- a'[k] = read_a[r][v[r][k]]; everything lives in the slot's pinned memory
 * until the slot is submitted again. */
int mm2gb_submit_gather_chains(mm2gb_ctx_t *ctx, int slot, const mm2gb_anchor_t *const *read_a, const int64_t *read_n, int n_reads);
int mm2gb_wait_chains(mm2gb_ctx_t *ctx, int slot, const uint64_t *const **u, const int32_t **n_u, const int32_t *const **v,
                      const int32_t **n_v, const int64_t **off, mm2gb_stats_t *stats);
int mm2gb_slot_busy(mm2gb_ctx_t *ctx, int slot);

/* ---- device-resident path (kernel-only timing; inputs already in HBM) ---------------------------------- */

/* d_a: device mm2gb_anchor_t[n_total]; d_off: device int64[n_reads+1]; d_f/d_p: device int32[n_total].
 * Enqueues range + unit + score kernels on slot 0's stream; returns without synchronising. */
int mm2gb_chain_dp_device(mm2gb_ctx_t *ctx, const void *d_a, const void *d_off, int n_reads, int64_t n_total, void *d_f, void *d_p);
/* The same followed by chain extraction + compaction on the device (results stay in the slot's device buffers; used to time
 * the whole device side of mg_lchain_dp).  `off` = host copy of d_off. */
int mm2gb_chain_device(mm2gb_ctx_t *ctx, const void *d_a, const void *d_off, const int64_t *off, int n_reads, int64_t n_total,
                       void *d_f, void *d_p);
/* The same on the stream and scratch of another slot: batches on different slots are independent, so the latency-bound chain
 * extraction of one batch overlaps the score kernels of the next (d_f / d_p must be distinct per batch in flight). */
int mm2gb_chain_device_slot(mm2gb_ctx_t *ctx, int slot, const void *d_a, const void *d_off, const int64_t *off, int n_reads,
                            int64_t n_total, void *d_f, void *d_p);
/* Results of a batch enqueued by mm2gb_chain_device_slot -> host, for producers that create the anchors on the device (mm2gb_seed.h:
 * the host holds no copy of them): the compacted anchors themselves (16 B each) and the packed chains are written into caller-supplied
 * pinned (mapped) host memory with room for n_total entries each; fetch enqueues, results waits.  Per read r: n_u[r] chains at
 * u_pinned[u_pos[r] ..], n_b[r] compacted anchors at b_pinned[b_pos[r] ..] (the per-read arrays live in the slot until its next use). */
int mm2gb_chain_device_fetch(mm2gb_ctx_t *ctx, int slot, const void *d_a, const void *d_off, int n_reads, int64_t n_total,
                             mm2gb_anchor_t *b_pinned, uint64_t *u_pinned);
int mm2gb_chain_device_results(mm2gb_ctx_t *ctx, int slot, const int32_t **n_u, const int32_t **u_pos, const int32_t **n_b,
                               const int32_t **b_pos, int64_t *n_chains, int64_t *n_chain_anchors, mm2gb_stats_t *stats);
int mm2gb_sync(mm2gb_ctx_t *ctx, int slot);
/* the cudaStream_t of a slot, as an opaque pointer (so a caller can record its own events on it) */
void *mm2gb_stream(mm2gb_ctx_t *ctx, int slot);
/* stats of the last batch run through mm2gb_chain_dp_device (after mm2gb_sync) */
int mm2gb_device_stats(mm2gb_ctx_t *ctx, mm2gb_stats_t *stats);

/* Per-kernel device time of slot 0, accumulated with CUDA events while profiling is on.
 * ms[0]=range ms[1]=unit-build ms[2]=score ms[3]=chain extraction (k_bt_sort* + k_bt_walk*) ms[4]=H2D ms[5]=D2H; launches[] likewise. */
#define MM2GB_NTIMERS 6
int mm2gb_profile(mm2gb_ctx_t *ctx, int enable);
int mm2gb_profile_read(mm2gb_ctx_t *ctx, float ms[MM2GB_NTIMERS], int64_t launches[MM2GB_NTIMERS]);

/* Diagnostic (tools/drain_probe.py): device -> pinned host of n chain-anchor indices by k_drain with `blocks` CTAs vs the copy engine,
 * alone and against a concurrent host -> device copy.  ms[0..4]: drain, memcpy, drain+H2D, memcpy+H2D, H2D alone. */
int mm2gb_debug_drain(mm2gb_ctx_t *ctx, int64_t n, int blocks, float ms[5]);

/* ---- host stage: backtracking + compaction of one read (lchain.c:27-111) ---------------------------------
 * u[<=n] (score<<32|count, ordered by chain start), b[<=n] compacted anchors.  Returns n_u (>=0); *n_b anchors kept.
 * max_drop = is_cdna ? INT32_MAX : bw (lchain.c:151,162). */
int32_t mm2gb_backtrack(int64_t n, const int32_t *f, const int32_t *p, const mm2gb_anchor_t *a, int32_t min_cnt, int32_t min_sc,
                        int32_t max_drop, uint64_t *u, mm2gb_anchor_t *b, int64_t *n_b);

/* The device version of that stage (k_bt_sort* / k_bt_walk*; what mm2gb_chain_host with n_threads <= 0 and the drop-in run behind the DP
 * kernels) on caller-supplied f / p: same outputs, layout as in mm2gb_chain_host.  Every read is finished on the device (reads a
 * shared-memory kernel cannot take -- scores >= 2^19, more chains than its key buffer holds -- go to the global-memory kernels
 * through a device-side list); *n_declined counts reads left unfinished and is 0 unless the call fails.  Synchronous; slot 0. */
int mm2gb_backtrack_device(mm2gb_ctx_t *ctx, const mm2gb_anchor_t *a, const int64_t *off, int n_reads, const int32_t *f,
                           const int32_t *p, uint64_t *u, int32_t *n_u, mm2gb_anchor_t *b, int64_t *n_b, int32_t *n_declined);

/* The same for a batch on `n_threads` host threads (layout of u/b/n_u/n_b as in mm2gb_chain_host). */
int mm2gb_backtrack_batch(const mm2gb_misc_t *misc, const mm2gb_anchor_t *a, const int64_t *off, int n_reads, const int32_t *f,
                          const int32_t *p, uint64_t *u, int32_t *n_u, mm2gb_anchor_t *b, int64_t *n_b, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
