/*
 * mm2gb_seed.h -- C ABI of the device seeding stage, part of libmm2gb_chain.so.
 *
 * What it replaces: the reference seeds every read on the host before the GPU sees it -- mm_map_seed (map.c:355-391, called
 * from the driver's worker at map.c:1001) = collect_minimizers -> mm_sketch (sketch.c:77-143), mm_seed_mz_flt (seed.c:5-29),
 * collect_seed_hits (map.c:295-331) -> mm_collect_matches (seed.c:98-131: mm_seed_collect_all :31-53 with mm_idx_get
 * index.c:81-97, mm_seed_select :57-96), anchor construction and radix_sort_128x (ksort.h:98-151).  Here the read SEQUENCES go
 * to the device, minimizers / index lookups / seed filters / anchors / the x-sort (bit-exact, including the tie order the
 * reference's unstable in-place radix sort leaves among anchors of equal x) are computed there, and the anchors go straight
 * into the chaining kernels without crossing PCIe.  Results are identical to the reference's arrays (tests/test_seed*.py).
 *
 * Scope: single-segment reads, minimizers with odd k <= 28 and w <= 32, plain or homopolymer-compressed (map-ont, map-hifi, map-pb,
 * asm* index settings), no sdust masking, and none of the flags that change seed collection (MM_F_NO_DIAG, MM_F_NO_DUAL, MM_F_FOR_ONLY, MM_F_REV_ONLY,
 * MM_F_QSTRAND, MM_F_HEAP_SORT).  Anything else is refused with MM2GB_EARG -- never approximated, never sent to a CPU path.
 *
 * Every function returns 0 on success and a negative MM2GB_E* code (mm2gb_chain.h) on failure; mm2gb_last_error() has the text.
 */
#ifndef MM2GB_SEED_H
#define MM2GB_SEED_H

#include "mm2gb_chain.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mm2gb_index mm2gb_index_t;
typedef struct mm2gb_seeder mm2gb_seeder_t;

/* the fields of mm_mapopt_t / mm_idx_t the seeding stage reads */
typedef struct {
    int32_t mid_occ;       /* opt->mid_occ: the max_occ argument of collect_seed_hits (map.c:381) and q_occ_max of mm_seed_mz_flt */
    int32_t max_max_occ;   /* opt->max_max_occ (4095) */
    int32_t occ_dist;      /* opt->occ_dist (500; 0 switches mm_seed_select off, seed.c:107) */
    float q_occ_frac;      /* opt->q_occ_frac (0.01; <= 0 switches mm_seed_mz_flt off, map.c:379) */
    int64_t flag;          /* opt->flag: refused if it holds a flag that changes seed collection */
    int32_t sdust_thres;   /* opt->sdust_thres: must be 0 */
    int32_t max_qlen;      /* opt->max_qlen: reads longer than this get no anchors (map.c:376); 0 = no limit */
} mm2gb_seed_params_t;

/* ---- index (mm_idx_t as the seeding stage sees it: minimizer -> positions sorted ascending, index.c:213-266) -----------
 * Built on `device` from ASCII sequences: sequence s is seqs[seq_off[s] .. seq_off[s+1]).  The minimizers come from the same
 * sketch kernel the reads go through; they are grouped on the host (set-up, once per index part) and live in an open-addressing
 * hash table in HBM.  bucket_bits is accepted for symmetry with mm_idx_str and ignored (the occurrence order does not depend on
 * it: worker_post sorts every list by position). */
int mm2gb_index_build(mm2gb_index_t **idx, int device, const char *seqs, const int64_t *seq_off, int n_seq, int w, int k, int is_hpc,
                      int bucket_bits);
/* The same index from lists a host index already holds (what enumerating mm_idx_t's buckets gives, index.c:213-266; works for indices
 * loaded from .mmi files and for indices built without their sequences, MM_I_NO_SEQ): key i = minimizer (mm128_t.x >> 8) with its
 * occurrences occ[off[i] .. off[i+1]) in the index's order (ascending positions).  Keys in any order. */
int mm2gb_index_from_lists(mm2gb_index_t **idx, int device, int w, int k, int is_hpc, int64_t n_keys, const uint64_t *keys, const int64_t *off,
                           const uint64_t *occ);
void mm2gb_index_destroy(mm2gb_index_t *idx);
/* mm_idx_cal_max_occ (index.c:186-207): what mm_mapopt_update makes of mid_occ_frac before clamping (options.c:72-77) */
int32_t mm2gb_index_cal_max_occ(const mm2gb_index_t *idx, float frac);
/* mm_idx_get (index.c:81-97) from the host copy of the table: occurrences of `minier`, at most cap copied; returns the count */
int64_t mm2gb_index_get(const mm2gb_index_t *idx, uint64_t minier, uint64_t *out, int64_t cap);
int64_t mm2gb_index_n_keys(const mm2gb_index_t *idx);
int64_t mm2gb_index_n_occ(const mm2gb_index_t *idx);

/* ---- seeder: device buffers for batches of up to max_bases bases / max_reads reads / max_anchors anchors ----------------- */
int mm2gb_seeder_create(mm2gb_seeder_t **sd, const mm2gb_index_t *idx, int64_t max_bases, int max_reads, int64_t max_anchors);
void mm2gb_seeder_destroy(mm2gb_seeder_t *sd);

/* mm_sketch of every sequence of a batch on the device (rid = sequence number if rid_is_seq, else 0); minimizers of sequence s
 * are out_xy[2 * mv_off[s] .. 2 * mv_off[s+1]) as (x, y) pairs.  Test / index-build entry; at most cap pairs are copied. */
int mm2gb_sketch_host(mm2gb_seeder_t *sd, const char *seqs, const int64_t *seq_off, int n_seq, int rid_is_seq, uint64_t *out_xy,
                      int64_t cap, int64_t *mv_off);

/* mm_map_seed of a batch, results on the host (parity entry): anchors of read r are a[a_off[r] .. a_off[r+1]) in the
 * reference's order; rep_len[r]; mini_pos of read r at mini_pos[mp_off[r] .. mp_off[r+1]).  a_cap / mp_cap: room of the outputs.
 * Any output pointer except a_off may be NULL. */
int mm2gb_seed_host(mm2gb_seeder_t *sd, const mm2gb_seed_params_t *prm, const char *seqs, const int64_t *seq_off, int n_reads,
                    mm2gb_anchor_t *a, int64_t a_cap, int64_t *a_off, int32_t *rep_len, uint64_t *mini_pos, int64_t mp_cap,
                    int64_t *mp_off);

/* The fused step: sequences in (host), seeding on the device, the anchors handed to the chaining context `ctx` (same device,
 * created with room for max_anchors / max_reads) without leaving HBM, and only the chains and their compacted anchors come back.
 * Per read r: n_u[r] chains at u[u_pos[r] ..] (score << 32 | count, lchain.c:60-75 order after compact_a), n_b[r] compacted
 * anchors at b[b_pos[r] ..], n_a[r] seeded anchors, rep_len[r].  The u / b arrays live in the seeder's pinned memory until its
 * next call.  seqs may be pageable or pinned host memory. */
typedef struct {
    int n_reads;
    int64_t n_anchors, n_chain_anchors, n_chains;
    const int64_t *a_off;      /* n_reads + 1: seeded anchors per read (prefix sums) */
    const int32_t *rep_len;    /* n_reads */
    const int32_t *n_u, *u_pos, *n_b, *b_pos;   /* n_reads each */
    const uint64_t *u;
    const mm2gb_anchor_t *b;
    mm2gb_stats_t stats;       /* chaining counters of the batch (pairs, units) */
    int64_t h2d_bytes, d2h_bytes;
} mm2gb_seed_chain_result_t;
int mm2gb_seed_chain(mm2gb_seeder_t *sd, mm2gb_ctx_t *ctx, const mm2gb_seed_params_t *prm, const char *seqs, const int64_t *seq_off,
                     int n_reads, mm2gb_seed_chain_result_t *res);

/* mini_pos of the last batch (mm_collect_matches, seed.c:124: q_span << 32 | q_pos >> 1 of every kept seed; what mm_est_err reads later,
 * map.c:608): read r's entries at mini_pos[mp_off[r] .. mp_off[r+1]).  Fetched on demand -- the fused step does not ship them. */
int mm2gb_seed_last_mini_pos(mm2gb_seeder_t *sd, int n_reads, uint64_t *mini_pos, int64_t cap, int64_t *mp_off);

/* The same with the sequences already resident in HBM (device pointer to the concatenated bases) and the results left on the
 * device: kernel-only timing of seeding + chaining.  Enqueues and synchronises once for the anchor counts (the chain-extraction
 * kernels are binned by read size on the host); *n_anchors = anchors seeded. */
int mm2gb_seed_chain_device(mm2gb_seeder_t *sd, mm2gb_ctx_t *ctx, const mm2gb_seed_params_t *prm, const void *d_seqs,
                            const int64_t *seq_off, int n_reads, int64_t *n_anchors);

/* Device time of the stages of the last batch (CUDA events on the seeder's stream), in ms:
 * [0] sketch [1] query-occurrence filter [2] index lookup [3] seed selection + match collection [4] anchor expansion [5] x-sort */
#define MM2GB_SEED_NTIMERS 6
int mm2gb_seed_profile(mm2gb_seeder_t *sd, float ms[MM2GB_SEED_NTIMERS], int64_t *n_minimizers, int64_t *n_seeds);

#ifdef __cplusplus
}
#endif
#endif
