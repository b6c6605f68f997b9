/*
 * dump_stub.c -- CPU stand-in for the reference's GPU layer (gpu/plchain.cu), for the oracle driver only.
 *
 * TEST INFRASTRUCTURE ONLY.  oracle/Makefile links this file with the reference's unmodified host
 * sources (main.c, map.c, lchain.c ... where they lie under /root/reference) into
 * oracle/_ref/minimap2_ref.  Without --gpu-chain that binary is plain CPU minimap2 (the PAF ground
 * truth).  With --gpu-chain it exercises the reference's batch hand-off protocol
 * (gpu/plutils.h:98-104, gpu/plchain.cu:292-305,496-546) but chains every read on the host with the
 * reference's own mg_lchain_dp, and -- if MM2GB_DUMP names a file -- writes each read's seeded
 * anchor array there so that oracle/gen_golden.py can build known-answer vectors from real reads.
 *
 * Dump format (little endian): "MM2GBAD1", Misc (44 bytes), then per read:
 *   int64 n, int32 n_seg, int32 qlen_sum, n * mm128_t.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mmpriv.h"
#include "plutils.h" /* reference boundary header, via -I$(REF)/gpu */

mm128_t *mg_lchain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
                      float chn_pen_gap, float chn_pen_skip, int is_cdna, int n_seg, int64_t n, mm128_t *a, int *n_u_,
                      uint64_t **_u, void *km);

#define STUB_MAX_THREADS 256
static struct { chain_read_t *reads; int n; } g_inflight[STUB_MAX_THREADS];
static FILE *g_dump;

void init_stream_gpu(size_t *max_total_n, int *max_reads, int *min_n, char gpu_config_file[], Misc misc)
{
    const char *fn = getenv("MM2GB_DUMP"), *s;
    (void)gpu_config_file;
    *max_total_n = (s = getenv("MM2GB_STUB_MAX_ANCHORS")) ? (size_t)atol(s) : 2000000;
    *max_reads = (s = getenv("MM2GB_STUB_MAX_READS")) ? atoi(s) : 512;
    *min_n = 0;
    memset(g_inflight, 0, sizeof(g_inflight));
    if (fn && !g_dump) {
        g_dump = fopen(fn, "wb");
        if (!g_dump) { perror("MM2GB_DUMP"); exit(1); }
        fwrite("MM2GBAD1", 1, 8, g_dump);
        fwrite(&misc, sizeof(Misc), 1, g_dump);
    }
}

static void stub_chain_batch(const mm_idx_t *mi, const mm_mapopt_t *opt, chain_read_t *reads, int n, void *km)
{
    Misc misc = build_misc(mi, opt, 0, 1);
    int i;
    for (i = 0; i < n; ++i) {
        chain_read_t *r = &reads[i];
        r->a = mg_lchain_dp(misc.max_dist_x, misc.max_dist_y, misc.bw, misc.max_skip, misc.max_iter, misc.min_cnt, misc.min_score,
                            misc.chn_pen_gap, misc.chn_pen_skip, misc.is_cdna, misc.n_seg, r->n, r->a, &r->n_u, &r->u, km);
        post_chaining_helper(mi, opt, r, misc, km);
    }
}

static void stub_dump(const chain_read_t *reads, int n)
{
    int i;
    if (!g_dump) return;
    for (i = 0; i < n; ++i) {
        int64_t na = reads[i].n;
        int32_t meta[2] = { reads[i].n_seg, reads[i].seq.qlen_sum };
        fwrite(&na, 8, 1, g_dump);
        fwrite(meta, 4, 2, g_dump);
        if (na > 0) fwrite(reads[i].a, sizeof(mm128_t), (size_t)na, g_dump);
    }
    fflush(g_dump);
}

/* The arena `km` belongs to the batch being handed BACK (map.c:1026 passes launched_batch.km), so the
 * host-side chaining of a batch runs when it is returned, not when it is submitted. */
void chain_stream_gpu(const mm_idx_t *mi, const mm_mapopt_t *opt, chain_read_t **in_arr_, int *n_read_, int thread_id, void *km)
{
    chain_read_t *prev = g_inflight[thread_id].reads;
    int n_prev = g_inflight[thread_id].n;
    stub_dump(*in_arr_, *n_read_);
    g_inflight[thread_id].reads = *in_arr_;
    g_inflight[thread_id].n = *n_read_;
    if (prev) stub_chain_batch(mi, opt, prev, n_prev, km);
    *in_arr_ = prev;
    *n_read_ = prev ? n_prev : 0;
}

void finish_stream_gpu(const mm_idx_t *mi, const mm_mapopt_t *opt, chain_read_t **reads_, int *n_read_, int t, void *km)
{
    chain_read_t *prev = g_inflight[t].reads;
    int n_prev = g_inflight[t].n;
    g_inflight[t].reads = 0, g_inflight[t].n = 0;
    if (prev) stub_chain_batch(mi, opt, prev, n_prev, km);
    *reads_ = prev;
    *n_read_ = prev ? n_prev : 0;
}

void free_stream_gpu(int n_threads)
{
    (void)n_threads;
    if (g_dump) { fclose(g_dump); g_dump = 0; }
}
