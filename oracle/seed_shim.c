/*
 * seed_shim.c -- thin harness around the UNMODIFIED reference seeding code (row N2 of SURVEY.md 8f).
 *
 * TEST INFRASTRUCTURE ONLY.  oracle/Makefile compiles this file together with the reference's host sources (sketch.c,
 * index.c, seed.c, map.c, lchain.c ... where they lie under /root/reference; nothing is copied) and dump_stub.c (which
 * satisfies the driver's four GPU entry points) into oracle/_ref/libref_seed.so.  It exposes, with plain pointers and
 * sizes so that ctypes can bind them:
 *     mm_idx_str          (index.c)      -> refseed_index_build      index of in-memory sequences
 *     mm_idx_get          (index.c:81)   -> refseed_index_get        occurrence list of one minimizer, in the index's order
 *     mm_idx_cal_max_occ  (index.c)      -> refseed_mid_occ
 *     mm_sketch           (sketch.c:77)  -> refseed_sketch
 *     mm_map_seed         (map.c:355-391)-> refseed_seed / refseed_seed_batch   the whole seeding stage of one read
 *     mm_map_seed + mg_lchain_dp          -> refseed_seed_chain_batch  (CPU baseline of the fused seed + chain step)
 * Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs may load the library.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "mmpriv.h"
#include "kalloc.h"
#include "plutils.h" /* chain_read_t, via -I$(REF)/gpu */

void mm_map_seed(const mm_idx_t *mi, const mm_mapopt_t *opt, chain_read_t *read_, mm_tbuf_t *b, void *km);
mm128_t *mg_lchain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
                      float chn_pen_gap, float chn_pen_skip, int is_cdna, int n_seg, int64_t n, mm128_t *a, int *n_u_,
                      uint64_t **_u, void *km);

void *refseed_index_build(int w, int k, int is_hpc, int bucket_bits, int n, const char **seq, const char **name)
{
    return mm_idx_str(w, k, is_hpc, bucket_bits, n, seq, name);
}

void refseed_index_destroy(void *mi) { if (mi) mm_idx_destroy((mm_idx_t *)mi); }

int refseed_mid_occ(const void *mi, float frac) { return mm_idx_cal_max_occ((const mm_idx_t *)mi, frac); }

/* occurrences of `minier` (= mm128_t.x >> 8 of mm_sketch) in index order; returns the count, copies at most cap entries */
int64_t refseed_index_get(const void *mi, uint64_t minier, uint64_t *out, int64_t cap)
{
    int n = 0, i;
    const uint64_t *cr = mm_idx_get((const mm_idx_t *)mi, minier, &n);
    for (i = 0; i < n && i < cap; ++i) out[i] = cr[i];
    return n;
}

/* minimizers of one sequence as (x, y) pairs; returns the count, copies at most cap pairs */
int64_t refseed_sketch(const char *seq, int len, int w, int k, uint32_t rid, int is_hpc, uint64_t *out_xy, int64_t cap)
{
    mm128_v mv = {0, 0, 0};
    int64_t i, n;
    mm_sketch(0, seq, len, w, k, rid, is_hpc, &mv);
    n = (int64_t)mv.n;
    for (i = 0; i < n && i < cap; ++i) out_xy[2 * i] = mv.a[i].x, out_xy[2 * i + 1] = mv.a[i].y;
    free(mv.a);
    return n;
}

/* map options as the driver sets them up: defaults, then the preset, then mm_mapopt_update against the index (mid_occ) */
void *refseed_opt_new(const char *preset, const void *mi)
{
    mm_idxopt_t io;
    mm_mapopt_t *mo = (mm_mapopt_t *)calloc(1, sizeof(mm_mapopt_t));
    mm_set_opt(0, &io, mo);
    if (preset && *preset && mm_set_opt(preset, &io, mo) < 0) { free(mo); return 0; }
    if (mi) mm_mapopt_update(mo, (const mm_idx_t *)mi);
    return mo;
}
void refseed_opt_free(void *opt) { free(opt); }

/* get / set the few fields the seeding stage reads (what < 0: get) */
double refseed_opt_field(void *opt_, const char *name, int set, double v)
{
    mm_mapopt_t *o = (mm_mapopt_t *)opt_;
#define FIELD(f, T) if (strcmp(name, #f) == 0) { if (set) o->f = (T)v; return (double)o->f; }
    FIELD(mid_occ, int32_t) FIELD(max_occ, int32_t) FIELD(max_max_occ, int32_t) FIELD(occ_dist, int32_t)
    FIELD(q_occ_frac, float) FIELD(flag, int64_t) FIELD(sdust_thres, int) FIELD(max_qlen, int)
    FIELD(max_gap, int) FIELD(max_gap_ref, int) FIELD(bw, int) FIELD(max_chain_skip, int) FIELD(max_chain_iter, int)
    FIELD(min_cnt, int) FIELD(min_chain_score, int) FIELD(chain_gap_scale, float) FIELD(chain_skip_scale, float)
#undef FIELD
    return -1e300;
}

static int64_t seed_one(const mm_idx_t *mi, const mm_mapopt_t *opt, mm_tbuf_t *b, void *km, const char *seq, int len,
                        chain_read_t *rd)
{
    const char *seqs[1];
    int qlens[1];
    memset(rd, 0, sizeof(*rd));
    seqs[0] = seq, qlens[0] = len;
    rd->n_seg = 1, rd->qseqs = seqs, rd->qlens = qlens;
    strcpy(rd->seq.name, "q");
    mm_map_seed(mi, opt, rd, b, km);
    return rd->n;
}

/* mm_map_seed of one read: anchors (x, y pairs, at most cap), rep_len, mini_pos (at most mcap); returns n_a */
int64_t refseed_seed(const void *mi, const void *opt, const char *seq, int len, uint64_t *out_xy, int64_t cap, int *rep_len,
                     uint64_t *mini_pos, int64_t mcap, int *n_mini_pos)
{
    mm_tbuf_t *b = mm_tbuf_init();
    void *km = km_init();
    chain_read_t rd;
    int64_t i, n = seed_one((const mm_idx_t *)mi, (const mm_mapopt_t *)opt, b, km, seq, len, &rd);
    for (i = 0; i < n && i < cap; ++i) out_xy[2 * i] = rd.a[i].x, out_xy[2 * i + 1] = rd.a[i].y;
    if (rep_len) *rep_len = rd.rep_len;
    if (n_mini_pos) *n_mini_pos = rd.n_mini_pos;
    for (i = 0; mini_pos && i < rd.n_mini_pos && i < mcap; ++i) mini_pos[i] = rd.mini_pos[i];
    km_destroy(km);
    mm_tbuf_destroy(b);
    return n;
}

/* ---- batches on host threads (timed CPU legs) ------------------------------------------------------------------- */

typedef struct {
    const mm_idx_t *mi; const mm_mapopt_t *opt;
    const char *seqs; const int64_t *seq_off; int n_reads;
    int chain;                 /* 0: seeding only, 1: seeding + mg_lchain_dp */
    uint64_t *out_xy; const int64_t *out_off;   /* optional: anchors of read r at out_xy + 2 * out_off[r] (room out_off[r+1]-out_off[r]) */
    int64_t *n_a; int32_t *n_u; uint64_t *digest;  /* per read */
    volatile int next;
    Misc misc;
} batch_t;

/* order-sensitive digest of a 64-bit word array, the formula of chain_oracle.c (orc_digest): C (m + 1) + sum_k w[k] (2k + 1) C mod 2^64
 * (vectorisable on the checking side: oracle/pyoracle.py digest) */
static uint64_t word_digest(const uint64_t *w, size_t m)
{
    const uint64_t c = 0x9E3779B97F4A7C15ULL;
    uint64_t h = c * (uint64_t)(m + 1);
    size_t k;
    for (k = 0; k < m; ++k) h += w[k] * ((2 * (uint64_t)k + 1) * c);
    return h;
}

static void *batch_worker(void *arg)
{
    batch_t *t = (batch_t *)arg;
    mm_tbuf_t *b = mm_tbuf_init();
    void *km = km_init();
    for (;;) {
        int r = __sync_fetch_and_add(&t->next, 1);
        chain_read_t rd;
        int64_t n, i;
        if (r >= t->n_reads) break;
        n = seed_one(t->mi, t->opt, b, km, t->seqs + t->seq_off[r], (int)(t->seq_off[r + 1] - t->seq_off[r]), &rd);
        t->n_a[r] = n;
        if (t->out_xy) {
            int64_t room = t->out_off[r + 1] - t->out_off[r];
            uint64_t *o = t->out_xy + 2 * t->out_off[r];
            for (i = 0; i < n && i < room; ++i) o[2 * i] = rd.a[i].x, o[2 * i + 1] = rd.a[i].y;
        }
        if (t->chain) {
            int n_u = 0;
            uint64_t *u = 0;
            mm128_t *a2 = mg_lchain_dp(t->misc.max_dist_x, t->misc.max_dist_y, t->misc.bw, t->misc.max_skip, t->misc.max_iter,
                                       t->misc.min_cnt, t->misc.min_score, t->misc.chn_pen_gap, t->misc.chn_pen_skip,
                                       t->misc.is_cdna, t->misc.n_seg, n, rd.a, &n_u, &u, km);
            int64_t nb = 0;
            for (i = 0; i < n_u; ++i) nb += (int32_t)u[i];
            if (t->n_u) t->n_u[r] = n_u;
            if (t->digest) t->digest[r] = word_digest(u, (size_t)n_u) + 31 * word_digest((const uint64_t *)a2, (size_t)nb * 2);
            kfree(km, a2); kfree(km, u);
        } else {
            if (t->digest) t->digest[r] = word_digest((const uint64_t *)rd.a, (size_t)n * 2);
            kfree(km, rd.a);
        }
        kfree(km, rd.mini_pos);
    }
    km_destroy(km);
    mm_tbuf_destroy(b);
    return 0;
}

/* reads r = 0..n_reads-1 are seqs[seq_off[r] .. seq_off[r+1]); per read n_a, and -- with chain -- n_u and
 * digest(u[]) + 31 * digest(compacted anchors) (without chain: digest of the anchor array; word_digest above).  n_threads host threads. */
int refseed_seed_batch(const void *mi, const void *opt, const char *seqs, const int64_t *seq_off, int n_reads, int chain,
                       int n_threads, uint64_t *out_xy, const int64_t *out_off, int64_t *n_a, int32_t *n_u, uint64_t *digest)
{
    batch_t t;
    pthread_t *th;
    int i;
    memset(&t, 0, sizeof(t));
    t.mi = (const mm_idx_t *)mi, t.opt = (const mm_mapopt_t *)opt, t.seqs = seqs, t.seq_off = seq_off, t.n_reads = n_reads;
    t.chain = chain, t.out_xy = out_xy, t.out_off = out_off, t.n_a = n_a, t.n_u = n_u, t.digest = digest;
    t.misc = build_misc(t.mi, t.opt, 0, 1);
    if (n_threads < 1) n_threads = 1;
    th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (i = 0; i < n_threads; ++i) pthread_create(&th[i], 0, batch_worker, &t);
    for (i = 0; i < n_threads; ++i) pthread_join(th[i], 0);
    free(th);
    return 0;
}

void refseed_misc(const void *mi, const void *opt, void *misc44)
{
    Misc m = build_misc((const mm_idx_t *)mi, (const mm_mapopt_t *)opt, 0, 1);
    memcpy(misc44, &m, sizeof(m));
}

/* radix_sort_128x (ksort.h via misc.c) on (x, y) pairs in place: the tie order among equal x is what the device sort must reproduce */
void refseed_radix_sort_128x(uint64_t *xy, int64_t n) { radix_sort_128x((mm128_t *)xy, (mm128_t *)xy + n); }
