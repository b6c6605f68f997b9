/*
 * mm2gb_plchain.h -- the DROP-IN boundary: the four entry points minimap2's host driver calls for --gpu-chain.
 *
 * These are exactly the symbols the reference declares in gpu/plutils.h:98-104 and implements in gpu/plchain.cu:470-560;
 * the driver (main.c:445-447,466 ; map.c:1026,1069) is NOT modified -- it keeps including its own gpu/plutils.h and
 * simply links libmm2gb_plchain.a + libmm2gb_chain.so instead of the objects of gpu/gpu.mk.
 *
 * This header restates the boundary types layout-for-layout (x86-64, release build, i.e. without DEBUG_CHECK /
 * DEBUG_VERBOSE which grow chain_read_t / seg_t -- plutils.h:54-57,79-82) and pins the layout with static asserts, so a
 * drift in the driver's struct is a compile error here, not silent corruption.
 */
#ifndef MM2GB_PLCHAIN_H
#define MM2GB_PLCHAIN_H

#include <stddef.h>
#include <stdint.h>

#include "mm2gb_chain.h"

#ifdef __cplusplus
extern "C" {
#endif

/* opaque to this layer: only passed through to the driver's own callbacks */
typedef struct mm_idx_s mm2gb_idx_t;       /* mm_idx_t,    minimap.h:94-113 */
typedef struct mm_mapopt_s mm2gb_mapopt_t; /* mm_mapopt_t, minimap.h:121-186 */

typedef mm2gb_misc_t Misc_abi; /* gpu/plutils.h:33-37, passed BY VALUE to init_stream_gpu / post_chaining_helper */

/* gpu/plutils.h:19-31 */
typedef struct {
    long i;
    int seg_id;
    char name[200];
    uint32_t len;
    int n_alt, is_alt;
    int qlen_sum;
} mm2gb_seq_meta_t;

/* gpu/plutils.h:45-73 (release layout) */
typedef struct {
    mm2gb_seq_meta_t seq;
    const char **qseqs;
    int *qlens;
    int n_seg;
    int rep_len;
    int frag_gap;
    uint64_t *mini_pos;
    int n_mini_pos;
    mm2gb_anchor_t *a; /* in: seeded anchors (kmalloc'd from the batch arena); out: compacted chain anchors or NULL */
    int64_t n;         /* in: number of anchors */
    uint64_t *u;       /* out: chains, score<<32 | count (kmalloc'd from the arena passed with the returned batch) */
    int n_u;           /* out: number of chains */
} mm2gb_chain_read_t;

#if defined(__x86_64__) || defined(__aarch64__)
#define MM2GB_SA(c, m) typedef char mm2gb_static_assert_##m[(c) ? 1 : -1]
MM2GB_SA(sizeof(mm2gb_chain_read_t) == 312, chain_read_size);   /* SURVEY.md 8b probe */
MM2GB_SA(offsetof(mm2gb_chain_read_t, rep_len) == 252, chain_read_rep_len);
MM2GB_SA(offsetof(mm2gb_chain_read_t, frag_gap) == 256, chain_read_frag_gap);
MM2GB_SA(offsetof(mm2gb_chain_read_t, a) == 280, chain_read_a);
MM2GB_SA(offsetof(mm2gb_chain_read_t, n) == 288, chain_read_n);
MM2GB_SA(offsetof(mm2gb_chain_read_t, u) == 296, chain_read_u);
MM2GB_SA(offsetof(mm2gb_chain_read_t, n_u) == 304, chain_read_n_u);
MM2GB_SA(sizeof(Misc_abi) == 44, misc_size);
MM2GB_SA(sizeof(mm2gb_anchor_t) == 16, anchor_size);
#endif

/* ---- exported by libmm2gb_plchain.a (replaces gpu/plchain.cu:470-560) -------------------------------------------------- */

/* plutils.h:98-99.  Reads the JSON named by --gpu-cfg (same file format as gpu/gpu_config.json; the tuning keys of the
 * old kernels are accepted and ignored), writes back the batch limits the driver enforces (map.c:1311-1314,887,946).
 * Fatal problems print to stderr and exit(1), as the reference does (gpu/hipify.cuh:47-55). */
void init_stream_gpu(size_t *max_total_n, int *max_reads, int *min_n, char gpu_config_file[], Misc_abi misc);

/* plutils.h:104.  Launch chaining of *in_arr_ (n = *n_read_ reads) asynchronously and hand back, through the same two
 * pointers, the batch launched by the previous call of this thread_id (NULL / 0 the first time).  For every returned
 * read: a = compacted anchors (kmalloc(km)) or NULL, u = chains (kmalloc(km)), n_u, and post_chaining_helper has run.
 * `km` is the arena that owns the RETURNED batch (map.c:1026). */
void chain_stream_gpu(const mm2gb_idx_t *mi, const mm2gb_mapopt_t *opt, mm2gb_chain_read_t **in_arr_, int *n_read_, int thread_id, void *km);

/* plutils.h:100-101.  Drain the in-flight batch of thread_id (NULL / 0 if idle). */
void finish_stream_gpu(const mm2gb_idx_t *mi, const mm2gb_mapopt_t *opt, mm2gb_chain_read_t **reads_, int *n_read_, int thread_id, void *km);

/* plutils.h:102.  Called unconditionally at the end of every index part (main.c:466); no-op when never initialised. */
void free_stream_gpu(int n_threads);

/* ---- callbacks the driver must provide (it already does: kalloc.c, map.c:393-484) ---------------------------------- */
void *kmalloc(void *km, size_t size);
void kfree(void *km, void *ptr);
Misc_abi build_misc(const mm2gb_idx_t *mi, const mm2gb_mapopt_t *opt, const int64_t qlen_sum, const int n_seg);
void post_chaining_helper(const mm2gb_idx_t *mi, const mm2gb_mapopt_t *opt, mm2gb_chain_read_t *read, Misc_abi misc, void *km);

#ifdef __cplusplus
}
#endif
#endif
