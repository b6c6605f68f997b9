/*
 * ref_shim.c -- thin harness around the UNMODIFIED reference chaining code.
 *
 * TEST INFRASTRUCTURE ONLY.  oracle/Makefile compiles this file together with
 * /root/reference/lchain.c and /root/reference/misc.c (where they lie; nothing is copied) into
 * oracle/_ref/libref_lchain.so.  The reference's kalloc.c is NOT compiled in: this file supplies
 * the kalloc.h entry points with a recording allocator, because mg_lchain_dp frees its f[]/p[]
 * arrays before returning (lchain.c:211) and the parity tests need them.  The allocator defers
 * every kfree() until ref_release(), so after mg_lchain_dp returns, the first two allocations it
 * made (lchain.c:163-164: p[n] then f[n]) are still readable.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "kalloc.h"  /* reference header, via -I$(REF) */
#include "minimap.h" /* mm128_t */

/* reference symbols (lchain.c, misc.c) */
mm128_t *mg_lchain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
                      float chn_pen_gap, float chn_pen_skip, int is_cdna, int n_seg, int64_t n, mm128_t *a, int *n_u_,
                      uint64_t **_u, void *km);
uint64_t *mg_chain_backtrack(void *km, int64_t n, const int32_t *f, const int64_t *p, int32_t *v, int32_t *t,
                             int32_t min_cnt, int32_t min_sc, int32_t max_drop, int32_t *n_u_, int32_t *n_v_);
mm128_t *compact_a(void *km, int32_t n_u, uint64_t *u, int32_t n_v, int32_t *v, mm128_t *a);
void radix_sort_128x(mm128_t *beg, mm128_t *end);

/* ---- recording allocator ------------------------------------------------------------------ */

typedef struct { void *ptr; size_t size; } rec_t;
static __thread rec_t *g_log;
static __thread size_t g_n, g_m;
static __thread void **g_dead;
static __thread size_t g_nd, g_md;

static void log_alloc(void *p, size_t size)
{
    if (g_n == g_m) {
        g_m = g_m ? g_m * 2 : 64;
        g_log = (rec_t *)realloc(g_log, g_m * sizeof(rec_t));
    }
    g_log[g_n].ptr = p, g_log[g_n++].size = size;
}

void *kmalloc(void *km, size_t size) { void *p = malloc(size ? size : 1); (void)km; log_alloc(p, size); return p; }
void *kcalloc(void *km, size_t count, size_t size) { void *p = calloc(count ? count : 1, size ? size : 1); (void)km; log_alloc(p, count * size); return p; }
void *krealloc(void *km, void *ptr, size_t size)
{ /* never moves a block in place: old block stays readable until ref_release() */
    void *p = malloc(size ? size : 1);
    size_t i, old = 0;
    (void)km;
    for (i = 0; i < g_n; ++i) if (g_log[i].ptr == ptr) old = g_log[i].size;
    if (ptr) { memcpy(p, ptr, old < size ? old : size); kfree(0, ptr); }
    log_alloc(p, size);
    return p;
}
void kfree(void *km, void *ptr)
{
    (void)km;
    if (!ptr) return;
    if (g_nd == g_md) {
        g_md = g_md ? g_md * 2 : 64;
        g_dead = (void **)realloc(g_dead, g_md * sizeof(void *));
    }
    g_dead[g_nd++] = ptr;
}
void *km_init(void) { return 0; }
void *km_init2(void *km_par, size_t min_core_size) { (void)km_par; (void)min_core_size; return 0; }
void km_destroy(void *km) { (void)km; }
void km_stat(const void *km, km_stat_t *s) { (void)km; memset(s, 0, sizeof(*s)); }

static void ref_release(void)
{
    size_t i;
    for (i = 0; i < g_nd; ++i) free(g_dead[i]);
    g_nd = 0, g_n = 0;
}

/* ---- exported harness ------------------------------------------------------------------------ */

/* prm: 9 ints + 2 floats in the order of gpu/plutils.h:33-37 (Misc) */
typedef struct {
    int32_t max_iter, max_dist_x, max_dist_y, max_skip, bw, min_cnt, min_score, is_cdna, n_seg;
    float chn_pen_gap, chn_pen_skip;
} ref_params_t;

/* Runs the reference mg_lchain_dp on a copy of a[n].  Outputs (any may be NULL): u[<=n], b[<=n] compacted
 * anchors, *n_b, f[n], p[n].  Returns n_u. */
int32_t ref_lchain(const ref_params_t *prm, int64_t n, const mm128_t *a, uint64_t *u_out, mm128_t *b_out, int64_t *n_b,
                   int32_t *f_out, int64_t *p_out)
{
    mm128_t *copy, *b;
    uint64_t *u = 0;
    int n_u = 0;
    int64_t i, nb = 0;
    size_t first;
    if (n_b) *n_b = 0;
    if (n <= 0) return 0;
    copy = (mm128_t *)kmalloc(0, (size_t)n * sizeof(mm128_t));
    memcpy(copy, a, (size_t)n * sizeof(mm128_t));
    first = g_n; /* allocations made by mg_lchain_dp start here: p, f, v, t (lchain.c:163-166) */
    b = mg_lchain_dp(prm->max_dist_x, prm->max_dist_y, prm->bw, prm->max_skip, prm->max_iter, prm->min_cnt, prm->min_score,
                     prm->chn_pen_gap, prm->chn_pen_skip, prm->is_cdna, prm->n_seg, n, copy, &n_u, &u, 0);
    if (p_out) memcpy(p_out, g_log[first].ptr, (size_t)n * sizeof(int64_t));
    if (f_out) memcpy(f_out, g_log[first + 1].ptr, (size_t)n * sizeof(int32_t));
    for (i = 0; i < n_u; ++i) nb += (int32_t)u[i];
    if (u_out && n_u > 0) memcpy(u_out, u, (size_t)n_u * sizeof(uint64_t));
    if (b_out && b) memcpy(b_out, b, (size_t)nb * sizeof(mm128_t));
    if (n_b) *n_b = nb;
    if (u) kfree(0, u);
    if (b) kfree(0, b);
    ref_release();
    return n_u;
}

/* the reference's radix_sort_128x in place (misc.c:167-168) */
void ref_radix_sort_128x(mm128_t *a, int64_t n) { radix_sort_128x(a, a + n); }

/* the reference's mg_chain_backtrack + compact_a on caller-supplied f/p (lchain.c:27-111) */
int32_t ref_backtrack_compact(int64_t n, const int32_t *f, const int64_t *p, const mm128_t *a, int32_t min_cnt, int32_t min_sc,
                              int32_t max_drop, uint64_t *u_out, mm128_t *b_out, int64_t *n_b)
{
    int32_t *v, *t, n_u = 0, n_v = 0;
    uint64_t *u;
    int64_t i, nb = 0;
    if (n_b) *n_b = 0;
    if (n <= 0) return 0;
    v = (int32_t *)kmalloc(0, (size_t)n * 4);
    t = (int32_t *)kcalloc(0, (size_t)n, 4);
    u = mg_chain_backtrack(0, n, f, p, v, t, min_cnt, min_sc, max_drop, &n_u, &n_v);
    if (n_u > 0) {
        mm128_t *copy = (mm128_t *)kmalloc(0, (size_t)n * sizeof(mm128_t)), *b;
        memcpy(copy, a, (size_t)n * sizeof(mm128_t));
        b = compact_a(0, n_u, u, n_v, v, copy);
        for (i = 0; i < n_u; ++i) nb += (int32_t)u[i];
        if (u_out) memcpy(u_out, u, (size_t)n_u * sizeof(uint64_t));
        if (b_out) memcpy(b_out, b, (size_t)nb * sizeof(mm128_t));
        kfree(0, b);
    } else kfree(0, v);
    if (u) kfree(0, u);
    kfree(0, t);
    if (n_b) *n_b = nb;
    ref_release();
    return n_u;
}

/* threaded batch driver used by bench.py's CPU-baseline legs (kind "reference") */
#include <pthread.h>
typedef struct { const ref_params_t *prm; const mm128_t *a; const int64_t *off; int64_t r1; int64_t *next; } ref_job_t;
static void *ref_worker(void *arg)
{
    ref_job_t *jb = (ref_job_t *)arg;
    for (;;) {
        int64_t r = __sync_fetch_and_add(jb->next, 1);
        if (r >= jb->r1) break;
        ref_lchain(jb->prm, jb->off[r + 1] - jb->off[r], jb->a + jb->off[r], 0, 0, 0, 0, 0);
    }
    free(g_log); free(g_dead);
    g_log = 0, g_dead = 0, g_m = g_md = 0;
    return 0;
}
void ref_lchain_batch(const ref_params_t *prm, const mm128_t *a, const int64_t *off, int64_t r0, int64_t r1, int n_threads)
{
    int64_t next = r0;
    int i;
    pthread_t *tid;
    ref_job_t jb;
    if (n_threads < 1) n_threads = 1;
    jb.prm = prm, jb.a = a, jb.off = off, jb.r1 = r1, jb.next = &next;
    tid = (pthread_t *)malloc((size_t)n_threads * sizeof(*tid));
    for (i = 0; i < n_threads; ++i) pthread_create(&tid[i], 0, ref_worker, &jb);
    for (i = 0; i < n_threads; ++i) pthread_join(tid[i], 0);
    free(tid);
}

/* per-read digests of the REFERENCE's whole mg_lchain_dp result (same digest as chain_oracle.c: orc_lchain_digest_batch) */
static uint64_t ref_digest(const uint64_t *w, int64_t n)
{
    uint64_t h = 0x9E3779B97F4A7C15ULL * (uint64_t)(n + 1);
    int64_t k;
    for (k = 0; k < n; ++k) h += w[k] * ((2 * (uint64_t)k + 1) * 0x9E3779B97F4A7C15ULL);
    return h;
}
typedef struct {
    const ref_params_t *prm; const mm128_t *a; const int64_t *off, *sel; int64_t n_sel; int64_t *next;
    int32_t *nu; int64_t *nb; uint64_t *hu, *hb;
} ref_djob_t;
static void *ref_digest_worker(void *arg)
{
    ref_djob_t *jb = (ref_djob_t *)arg;
    for (;;) {
        int64_t k = __sync_fetch_and_add(jb->next, 1), r, n, nb = 0;
        int32_t nu = 0;
        uint64_t *u;
        mm128_t *b;
        if (k >= jb->n_sel) break;
        r = jb->sel[k];
        n = jb->off[r + 1] - jb->off[r];
        u = (uint64_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(*u));
        b = (mm128_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(*b));
        if (n > 0) nu = ref_lchain(jb->prm, n, jb->a + jb->off[r], u, b, &nb, 0, 0);
        jb->nu[k] = nu; jb->nb[k] = nb;
        jb->hu[k] = ref_digest(u, nu); jb->hb[k] = ref_digest((const uint64_t *)b, 2 * nb);
        free(u); free(b);
    }
    free(g_log); free(g_dead);
    g_log = 0, g_dead = 0, g_m = g_md = 0;
    return 0;
}
void ref_lchain_digest_batch(const ref_params_t *prm, const mm128_t *a, const int64_t *off, const int64_t *sel, int64_t n_sel,
                             int n_threads, int32_t *nu, int64_t *nb, uint64_t *hu, uint64_t *hb)
{
    int64_t next = 0;
    int i;
    pthread_t *tid;
    ref_djob_t jb;
    if (n_threads < 1) n_threads = 1;
    jb.prm = prm, jb.a = a, jb.off = off, jb.sel = sel, jb.n_sel = n_sel, jb.next = &next;
    jb.nu = nu, jb.nb = nb, jb.hu = hu, jb.hb = hb;
    tid = (pthread_t *)malloc((size_t)n_threads * sizeof(*tid));
    for (i = 0; i < n_threads; ++i) pthread_create(&tid[i], 0, ref_digest_worker, &jb);
    for (i = 0; i < n_threads; ++i) pthread_join(tid[i], 0);
    free(tid);
}
