/* fake_host.c -- stands in for the minimap2 host driver in the drop-in protocol tests (tests/test_dropin.py) and in
 * bench.py's end-to-end leg.
 * Provides the callbacks libmm2gb_plchain.a expects from the driver (include/mm2gb_plchain.h): a counting malloc-backed
 * kmalloc/kfree, build_misc returning a preset Misc, and a post_chaining_helper that records frag_gap like map.c:483;
 * and fake_drive(), which calls the boundary the way `minimap2 -t T --gpu-chain` does (map.c:924-1071, kthread.c:41-57):
 * T worker threads, each accumulating its own batch of seeded reads (per-read kmalloc'd anchor arrays), launching it with
 * chain_stream_gpu(thread_id) and flushing with finish_stream_gpu at the end of every mini-batch. */
#define _GNU_SOURCE
#include <malloc.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mm2gb_plchain.h"

static Misc_abi g_misc;
static long g_live, g_helper_calls;
static int g_misc_depends_on_qlen;

void fake_set_misc(const Misc_abi *m) { g_misc = *m; }
void fake_set_misc_depends_on_qlen(int on) { g_misc_depends_on_qlen = on; }
long fake_live_blocks(void) { return __atomic_load_n(&g_live, __ATOMIC_RELAXED); }
long fake_helper_calls(void) { return __atomic_load_n(&g_helper_calls, __ATOMIC_RELAXED); }

/* km == NULL: plain malloc (the protocol tests count the blocks).  km != NULL: a per-thread arena that, like kalloc
 * (kalloc.c:71-145), hands out pieces of memory it already owns -- no system call and no page fault per request; the driver
 * recycles its arenas once a mini-batch has been written out, fake_drive resets them at the start of every mini-batch. */
typedef struct { char *base; size_t cap, used; } fake_arena_t;

void *kmalloc(void *km, size_t size)
{
    if (size == 0) return 0; /* kalloc.c returns NULL for empty requests */
    __atomic_add_fetch(&g_live, 1, __ATOMIC_RELAXED);
    if (km) {
        fake_arena_t *ar = (fake_arena_t *)km;
        const size_t need = (size + 15) & ~(size_t)15;   /* kalloc's unit is 16 bytes */
        if (ar->used + need <= ar->cap) { void *p = ar->base + ar->used; ar->used += need; return p; }
    }
    return malloc(size);
}
void kfree(void *km, void *ptr)
{
    if (!ptr) return;
    __atomic_sub_fetch(&g_live, 1, __ATOMIC_RELAXED);
    if (km) {
        fake_arena_t *ar = (fake_arena_t *)km;
        if ((char *)ptr >= ar->base && (char *)ptr < ar->base + ar->cap) return;   /* recycled with the arena */
    }
    free(ptr);
}
Misc_abi build_misc(const mm2gb_idx_t *mi, const mm2gb_mapopt_t *opt, const int64_t qlen_sum, const int n_seg)
{
    Misc_abi m = g_misc;
    (void)mi; (void)opt; (void)n_seg;
    if (g_misc_depends_on_qlen && qlen_sum > m.max_dist_y) m.max_dist_y = (int)(qlen_sum > 0x7fffffff ? 0x7fffffff : qlen_sum); /* map.c:398-399 (MM_F_SR) */
    return m;
}
void post_chaining_helper(const mm2gb_idx_t *mi, const mm2gb_mapopt_t *opt, mm2gb_chain_read_t *read, Misc_abi misc, void *km)
{
    (void)mi; (void)opt; (void)km;
    read->frag_gap = misc.max_dist_x; /* map.c:483 */
    __atomic_add_fetch(&g_helper_calls, 1, __ATOMIC_RELAXED);
}
/* allocate a read's anchor array from the fake arena */
mm2gb_anchor_t *fake_alloc_anchors(const mm2gb_anchor_t *src, int64_t n)
{
    mm2gb_anchor_t *a = (mm2gb_anchor_t *)kmalloc(0, (size_t)n * sizeof(*a));
    if (a) memcpy(a, src, (size_t)n * sizeof(*a));
    return a;
}

/* ---- the driver's call pattern, for the end-to-end measurement ------------------------------------------------------ */

/* order-sensitive 64-bit digest of a word array (the same formula in tests / bench.py / the oracle side) */
uint64_t fake_digest(const uint64_t *w, int64_t n)
{
    uint64_t h = 0x9E3779B97F4A7C15ULL * (uint64_t)(n + 1);
    for (int64_t k = 0; k < n; ++k) h += w[k] * ((2 * (uint64_t)k + 1) * 0x9E3779B97F4A7C15ULL);
    return h;
}

typedef struct { mm2gb_chain_read_t *reads; int n, r0; } fake_batch_t;

typedef struct {
    int tid, steps, sync_steps;
    int n_batches;              /* per step */
    fake_batch_t *batches;      /* [steps][n_batches] */
    pthread_barrier_t *bar;
    fake_arena_t arena;         /* where the results of this thread's reads go (u, a') */
} fake_worker_t;

static void *fake_worker(void *arg)
{
    fake_worker_t *w = (fake_worker_t *)arg;
    pthread_barrier_wait(w->bar);       /* start of the timed region */
    for (int s = 0; s < w->steps; ++s) {
        w->arena.used = 0;              /* the previous mini-batch has been written out: its arena is recycled */
        for (int b = 0; b < w->n_batches; ++b) {
            fake_batch_t *fb = &w->batches[s * w->n_batches + b];
            mm2gb_chain_read_t *ptr = fb->reads;
            int n = fb->n;
            chain_stream_gpu(0, 0, &ptr, &n, w->tid, &w->arena);   /* map.c:1026: hands back the batch launched before (results are in its reads) */
        }
        {   /* kt_for's flush call at the end of the mini-batch (kthread.c:52-55 -> map.c:1069) */
            mm2gb_chain_read_t *ptr = 0;
            int n = 0;
            finish_stream_gpu(0, 0, &ptr, &n, w->tid, &w->arena);
        }
        if (w->sync_steps) pthread_barrier_wait(w->bar);   /* kt_for joins its workers before the next mini-batch */
    }
    pthread_barrier_wait(w->bar);       /* end of the timed region */
    return 0;
}

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

/* `steps` mini-batches of the reads a[off[r] .. off[r+1]) through the boundary on n_threads worker threads (thread ids
 * tid0 .. tid0 + n_threads - 1); a thread launches a batch whenever the next read would take it past batch_anchors (the
 * driver's gpu_chain_max_anchors, map.c:886-922).  Everything the timed region needs (the seeded reads of every step, i.e.
 * what mm_map_seed leaves behind) is built before it starts; results are digested and freed after it ends.
 * Per read of the LAST step: n_u, number of chain anchors, digests of u[] and a'[] (any may be NULL).  Returns the seconds of
 * the timed region (first chain_stream_gpu call to the last finish_stream_gpu return, all threads). */
double fake_drive(const mm2gb_anchor_t *a, const int64_t *off, int n_reads, int n_threads, int tid0, int steps, int64_t batch_anchors,
                  int sync_steps, int32_t *out_nu, int64_t *out_nb, uint64_t *out_hu, uint64_t *out_hb)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_reads && n_reads > 0) n_threads = n_reads;
    /* kalloc hands out pieces of arena blocks it keeps (kalloc.c:71-99): no system call, no page fault per read.  Make malloc
     * behave alike -- no mmap / munmap per large array, no trimming -- so that the timed region measures the boundary and not
     * the kernel's page-fault path */
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_TOP_PAD, 64 << 20);
    fake_worker_t *ws = (fake_worker_t *)calloc((size_t)n_threads, sizeof(*ws));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(*th));
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, 0, (unsigned)n_threads + 1);
    /* reads are dealt to the threads in contiguous shares of about equal anchor counts */
    int r = 0;
    const int64_t total = off[n_reads];
    for (int t = 0; t < n_threads; ++t) {
        const int64_t goal = total * (t + 1) / n_threads;
        int r1 = r;
        while (r1 < n_reads && (t == n_threads - 1 || off[r1 + 1] <= goal || r1 == r)) ++r1;
        /* batches of this share */
        int nb = 0, q = r;
        while (q < r1) { int64_t c = 0; int q1 = q; while (q1 < r1 && (q1 == q || c + (off[q1 + 1] - off[q1]) <= batch_anchors)) c += off[q1 + 1] - off[q1], ++q1; ++nb; q = q1; }
        ws[t].tid = tid0 + t; ws[t].steps = steps; ws[t].sync_steps = sync_steps; ws[t].n_batches = nb; ws[t].bar = &bar;
        ws[t].batches = (fake_batch_t *)calloc((size_t)steps * (size_t)(nb > 0 ? nb : 1), sizeof(fake_batch_t));
        {   /* room for the results of one mini-batch of this share (at most every anchor in a chain: 16 B a' + 8 B u), touched once */
            const int64_t share = off[r1] - off[r];
            ws[t].arena.cap = (size_t)share * 24 + (size_t)(r1 - r) * 32 + 4096;
            ws[t].arena.base = (char *)malloc(ws[t].arena.cap);
            memset(ws[t].arena.base, 0, ws[t].arena.cap);
            ws[t].arena.used = 0;
        }
        for (int s = 0; s < steps; ++s) {
            int b = 0;
            q = r;
            while (q < r1) {
                int64_t c = 0; int q1 = q;
                while (q1 < r1 && (q1 == q || c + (off[q1 + 1] - off[q1]) <= batch_anchors)) c += off[q1 + 1] - off[q1], ++q1;
                fake_batch_t *fb = &ws[t].batches[s * nb + b++];
                fb->n = q1 - q; fb->r0 = q;
                fb->reads = (mm2gb_chain_read_t *)calloc((size_t)fb->n, sizeof(mm2gb_chain_read_t));
                for (int k = 0; k < fb->n; ++k) {
                    const int64_t n = off[q + k + 1] - off[q + k];
                    fb->reads[k].n = n; fb->reads[k].n_seg = 1; fb->reads[k].seq.i = q + k;
                    fb->reads[k].a = n > 0 ? fake_alloc_anchors(a + off[q + k], n) : 0;
                }
                q = q1;
            }
        }
        r = r1;
    }
    for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], 0, fake_worker, &ws[t]);
    pthread_barrier_wait(&bar);
    const double t0 = now_s();
    if (sync_steps) for (int s = 0; s < steps; ++s) pthread_barrier_wait(&bar);
    pthread_barrier_wait(&bar);
    const double dt = now_s() - t0;
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], 0);
    for (int t = 0; t < n_threads; ++t) {
        for (int s = 0; s < steps; ++s)
            for (int b = 0; b < ws[t].n_batches; ++b) {
                fake_batch_t *fb = &ws[t].batches[s * ws[t].n_batches + b];
                for (int k = 0; k < fb->n; ++k) {
                    mm2gb_chain_read_t *rd = &fb->reads[k];
                    if (s == steps - 1) {
                        int64_t nb = 0;
                        for (int c = 0; c < rd->n_u; ++c) nb += (int32_t)rd->u[c];
                        if (out_nu) out_nu[fb->r0 + k] = rd->n_u;
                        if (out_nb) out_nb[fb->r0 + k] = nb;
                        if (out_hu) out_hu[fb->r0 + k] = fake_digest(rd->u, rd->n_u);
                        if (out_hb) out_hb[fb->r0 + k] = fake_digest((const uint64_t *)rd->a, 2 * nb);
                    }
                    kfree(&ws[t].arena, rd->a); kfree(&ws[t].arena, rd->u);
                }
                free(fb->reads);
            }
        free(ws[t].batches);
        free(ws[t].arena.base);
    }
    pthread_barrier_destroy(&bar);
    free(ws); free(th);
    return dt;
}
