/*
 * mm2gb_seed_glue.c -- the reference-side binding for seeding on the device (row N2).
 *
 * Compiled WITH the reference's headers (minimap.h, mmpriv.h, gpu/plutils.h) and linked into its host driver next to
 * libmm2gb_chain.so; it is the code a maintainer adds, not part of the library.  It turns the driver's per-read
 *     mm_map_seed(mi, opt, read, b, km)              (map.c:1001)      host: sketch, index lookup, seed filters, anchors, sort
 *     mm_map_chain(mi, opt, read, b, km)             (map.c:1061)      host: mg_lchain_dp + post_chaining_helper
 * into one call per batch,
 *     mm2gb_glue_seed_chain_batch(mi, opt, reads, n, tid, km)
 * that ships the read sequences to the GPU, runs include/mm2gb_seed.h's fused step and leaves in every chain_read_t exactly
 * what the two host calls leave there (n, rep_len, mini_pos, n_mini_pos, a = compacted anchors, u, n_u, frag_gap), so
 * mm_map_align (map.c:566-635) continues unchanged.  Switched on with MM2GB_GPU_SEED=1; the two-line change of map.c that calls
 * it is integration/map_gpu_seed.sed (applied to a scratch copy by oracle/Makefile for the test binary minimap2_b200_seed).
 *
 * Batches the device cannot take (even k, sdust, multi-segment reads, all-vs-all flags) end the run with a message:
 * leave MM2GB_GPU_SEED unset for those.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "mmpriv.h"
#include "kalloc.h"
#include "plutils.h"
#include "mm2gb_seed.h"

int64_t mm_idx_export(const mm_idx_t *mi, uint64_t **keys_, int64_t **off_, uint64_t **occ_);   /* integration/index_export.inc, appended to index.c */

#define GLUE_MAX_THREADS 256

static int g_enabled = -1;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
#define GLUE_MAX_GPUS 16
static mm2gb_index_t *g_idx[GLUE_MAX_GPUS];      /* one device index per GPU: worker thread tid drives GPU tid % n_gpus */
static const mm_idx_t *g_idx_of[GLUE_MAX_GPUS];
static int g_n_gpus;

typedef struct {
    mm2gb_seeder_t *sd;
    mm2gb_ctx_t *ctx;
    int device;
    int64_t cap_bases, cap_anchors;
    int cap_reads;
    char *buf; int64_t buf_cap;
    int64_t *off; int off_cap;
    uint64_t *mp; int64_t mp_cap;
    int64_t *mp_off;
    long n_batches, n_reads;
    double t_gpu;
} tstate_t;
static tstate_t g_ts[GLUE_MAX_THREADS];

static void glue_die(const char *what)
{
    fprintf(stderr, "[ERROR] mm2gb seeding: %s: %s\n", what, mm2gb_last_error());
    fflush(0);
    _exit(1);
}

/* CUDA bring-up (driver + context: 1-3 s on a fresh process) in the background from process start, while the driver reads / builds its
 * index; the first batch then finds the device ready */
static void *glue_early(void *arg) { size_t f = 0, t = 0; (void)arg; mm2gb_device_memory(0, &f, &t); return 0; }
__attribute__((constructor)) static void glue_ctor(void)
{
    const char *e = getenv("MM2GB_GPU_SEED");
    if (e && atoi(e)) { pthread_t th; if (pthread_create(&th, 0, glue_early, 0) == 0) pthread_detach(th); }
}

void mm2gb_glue_report(int n_threads);
static void glue_atexit(void) { mm2gb_glue_report(GLUE_MAX_THREADS); }

int mm2gb_glue_enabled(void)
{
    if (g_enabled < 0) {
        const char *e = getenv("MM2GB_GPU_SEED");
        g_enabled = (e && atoi(e)) ? 1 : 0;
        if (g_enabled) {
            atexit(glue_atexit);
            setenv("MM2GB_STAGE_THREADS", "2", 0);     /* several worker threads share the cores: two staging threads per seeder */
        }
    }
    return g_enabled;
}

/* reads per device batch and driver thread (the host path's N_ACCUM = 64, map.c:23, is sized for per-read host chaining) */
int mm2gb_glue_batch_reads(void)
{
    const char *e = getenv("MM2GB_SEED_BATCH_READS");
    int n = e ? atoi(e) : 512;
    return n < 1 ? 1 : n > 65536 ? 65536 : n;
}

/* what mm_map_seed leaves in a read before any seeding happened (map.c:372-376): qlen_sum, nothing else */
void mm2gb_glue_defer_seed(chain_read_t *rd)
{
    int i, sum = 0;
    for (i = 0; i < rd->n_seg; ++i) sum += rd->qlens[i];
    rd->seq.qlen_sum = sum;
    rd->a = 0; rd->n = 0; rd->u = 0; rd->n_u = 0; rd->mini_pos = 0; rd->n_mini_pos = 0; rd->rep_len = 0;
}

static int glue_n_gpus(void)
{
    if (!g_n_gpus) {
        const char *e = getenv("MM2GB_N_GPUS");
        int n = mm2gb_device_count();
        if (n <= 0) { fprintf(stderr, "[ERROR] mm2gb seeding: MM2GB_GPU_SEED=1 needs a CUDA device (no CPU fallback)\n"); fflush(0); _exit(1); }
        if (e && atoi(e) > 0 && atoi(e) < n) n = atoi(e);
        g_n_gpus = n > GLUE_MAX_GPUS ? GLUE_MAX_GPUS : n;
    }
    return g_n_gpus;
}

static mm2gb_index_t *glue_index(const mm_idx_t *mi, int dev)
{
    pthread_mutex_lock(&g_mu);
    if (g_idx_of[dev] != mi) {     /* once per index part and GPU: the part's minimizer lists (integration/index_export.inc) -> device index */
        uint64_t *keys = 0, *occ = 0;
        int64_t *off = 0, n_keys;
        if (g_idx[dev]) mm2gb_index_destroy(g_idx[dev]), g_idx[dev] = 0;
        double t0 = realtime(), t1;
        n_keys = mm_idx_export(mi, &keys, &off, &occ);
        t1 = realtime();
        if (mm2gb_index_from_lists(&g_idx[dev], dev, mi->w, mi->k, mi->flag & MM_I_HPC, n_keys, keys, off, occ)) glue_die("building the device index");
        if (getenv("MM2GB_VERBOSE")) fprintf(stderr, "[mm2gb] device index on GPU %d: %ld keys, export %.3f s, table + upload %.3f s\n", dev, (long)n_keys, t1 - t0, realtime() - t1);
        free(keys); free(off); free(occ);
        g_idx_of[dev] = mi;
    }
    pthread_mutex_unlock(&g_mu);
    return g_idx[dev];
}

static void glue_size(tstate_t *ts, const mm_idx_t *mi, const mm_mapopt_t *opt, mm2gb_index_t *idx, int64_t bases, int n_reads, int64_t anchors)
{
    if (ts->sd && bases <= ts->cap_bases && n_reads <= ts->cap_reads && anchors <= ts->cap_anchors) return;
    if (ts->sd) mm2gb_seeder_destroy(ts->sd), ts->sd = 0;
    if (ts->ctx) mm2gb_ctx_destroy(ts->ctx), ts->ctx = 0;
    if (bases > ts->cap_bases) ts->cap_bases = bases + bases / 4 + (1 << 20);
    if (n_reads > ts->cap_reads) ts->cap_reads = n_reads + 64;
    if (anchors > ts->cap_anchors) ts->cap_anchors = anchors;
    {
        Misc m = build_misc(mi, opt, 0, 1);
        mm2gb_misc_t misc;
        memcpy(&misc, &m, sizeof(misc));
        double t0 = realtime();
        if (mm2gb_seeder_create(&ts->sd, idx, ts->cap_bases, ts->cap_reads, ts->cap_anchors)) glue_die("creating the seeder");
        if (mm2gb_ctx_create_ex(&ts->ctx, ts->device, (size_t)ts->cap_anchors, ts->cap_reads, 1, &misc, MM2GB_CTX_DEVICE_ONLY)) glue_die("creating the chaining context");
        if (getenv("MM2GB_VERBOSE")) fprintf(stderr, "[mm2gb] seeder + context for %ld bases / %ld anchors: %.3f s\n", (long)ts->cap_bases, (long)ts->cap_anchors, realtime() - t0);
    }
}

void mm2gb_glue_seed_chain_batch(const mm_idx_t *mi, const mm_mapopt_t *opt, chain_read_t *reads, int n_reads, int tid, void *km)
{
    tstate_t *ts;
    mm2gb_index_t *idx;
    mm2gb_seed_params_t prm;
    mm2gb_seed_chain_result_t r;
    Misc misc;
    int64_t bases = 0;
    int i, rc;
    double t0;
    if (n_reads <= 0) return;
    if (tid < 0 || tid >= GLUE_MAX_THREADS) { fprintf(stderr, "[ERROR] mm2gb seeding: thread id %d out of range\n", tid); fflush(0); _exit(1); }
    ts = &g_ts[tid];
    ts->device = tid % glue_n_gpus();
    idx = glue_index(mi, ts->device);
    misc = build_misc(mi, opt, 0, 1);
    if (n_reads + 1 > ts->off_cap) {
        ts->off_cap = n_reads + 65;
        ts->off = (int64_t *)realloc(ts->off, (size_t)ts->off_cap * sizeof(int64_t));
        ts->mp_off = (int64_t *)realloc(ts->mp_off, (size_t)ts->off_cap * sizeof(int64_t));
    }
    ts->off[0] = 0;
    for (i = 0; i < n_reads; ++i) {
        int len = reads[i].qlens[0];
        if (reads[i].n_seg != 1) { fprintf(stderr, "[ERROR] mm2gb seeding: read %ld has %d segments; single-segment reads only\n", reads[i].seq.i, reads[i].n_seg); fflush(0); _exit(1); }
        if (opt->max_qlen > 0 && len > opt->max_qlen) len = 0;       /* map.c:376: such reads get no anchors */
        ts->off[i + 1] = ts->off[i] + len;
    }
    bases = ts->off[n_reads];
    if (bases + 1 > ts->buf_cap) { ts->buf_cap = bases + bases / 4 + (1 << 20); ts->buf = (char *)realloc(ts->buf, (size_t)ts->buf_cap); }
    for (i = 0; i < n_reads; ++i) memcpy(ts->buf + ts->off[i], reads[i].qseqs[0], (size_t)(ts->off[i + 1] - ts->off[i]));
    prm.mid_occ = opt->mid_occ; prm.max_max_occ = opt->max_max_occ; prm.occ_dist = opt->occ_dist; prm.q_occ_frac = opt->q_occ_frac;
    prm.flag = opt->flag; prm.sdust_thres = opt->sdust_thres; prm.max_qlen = opt->max_qlen;
    t0 = realtime();
    glue_size(ts, mi, opt, idx, bases, n_reads, ts->cap_anchors ? ts->cap_anchors : (bases / 4 > (1 << 20) ? bases / 4 : (1 << 20)));
    for (;;) {
        rc = mm2gb_seed_chain(ts->sd, ts->ctx, &prm, ts->buf, ts->off, n_reads, &r);
        if (rc == MM2GB_ECAP) { glue_size(ts, mi, opt, idx, bases, n_reads, ts->cap_anchors * 2); continue; }   /* more anchors than planned for: grow, again */
        if (rc) glue_die("seed + chain");
        break;
    }
    if (bases + 16 > ts->mp_cap) { ts->mp_cap = bases + bases / 4 + (1 << 16); ts->mp = (uint64_t *)realloc(ts->mp, (size_t)ts->mp_cap * sizeof(uint64_t)); }
    if (mm2gb_seed_last_mini_pos(ts->sd, n_reads, ts->mp, ts->mp_cap, ts->mp_off)) glue_die("fetching mini_pos");
    ts->t_gpu += realtime() - t0;
    ts->n_batches++; ts->n_reads += n_reads;
    for (i = 0; i < n_reads; ++i) {
        chain_read_t *rd = &reads[i];
        const int64_t nmp = ts->mp_off[i + 1] - ts->mp_off[i];
        rd->n = r.a_off[i + 1] - r.a_off[i];
        rd->rep_len = r.rep_len[i];
        rd->n_mini_pos = (int)nmp;
        rd->mini_pos = (uint64_t *)kmalloc(km, (size_t)(nmp > 0 ? nmp : 1) * sizeof(uint64_t));   /* mm_collect_matches always allocates it (seed.c:103) */
        memcpy(rd->mini_pos, ts->mp + ts->mp_off[i], (size_t)nmp * sizeof(uint64_t));
        if (r.n_u[i] > 0) {
            rd->n_u = r.n_u[i];
            rd->u = (uint64_t *)kmalloc(km, (size_t)r.n_u[i] * sizeof(uint64_t));
            memcpy(rd->u, r.u + r.u_pos[i], (size_t)r.n_u[i] * sizeof(uint64_t));
            rd->a = (mm128_t *)kmalloc(km, (size_t)r.n_b[i] * sizeof(mm128_t));
            memcpy(rd->a, (const char *)r.b + (size_t)r.b_pos[i] * sizeof(mm128_t), (size_t)r.n_b[i] * sizeof(mm128_t));
        } else {                                  /* lchain.c:212-215 */
            rd->a = 0; rd->u = 0; rd->n_u = 0;
        }
        post_chaining_helper(mi, opt, rd, misc, km);
    }
}

void mm2gb_glue_report(int n_threads)
{
    int t;
    if (!mm2gb_glue_enabled() || !getenv("MM2GB_VERBOSE")) return;
    for (t = 0; t < n_threads && t < GLUE_MAX_THREADS; ++t)
        if (g_ts[t].n_batches)
            fprintf(stderr, "[mm2gb] seed+chain thread %d: %ld batches, %ld reads, %.3f s inside the fused call\n", t, g_ts[t].n_batches, g_ts[t].n_reads, g_ts[t].t_gpu);
}
