/*
 * chain_oracle.c -- CPU restatement of minimap2-v2.24 anchor chaining (see chain_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 * Parity status: PINNED against the reference's lchain.c compiled as oracle/_ref/libref_lchain.so
 * (tests/test_oracle.py, tests/golden/).
 *
 * Build: gcc -O3 -ffp-contract=off (no -march, no -ffast-math): the reference is built by plain
 * `gcc -O3` on x86-64, which emits separate mulss/addss and cvttss2si for the float penalty
 * (SURVEY.md Appendix B.4), so this file must not be contracted into FMAs either.
 */
#include "chain_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define ORC_SEG_SHIFT 48 /* mmpriv.h:23 MM_SEED_SEG_SHIFT */
#define ORC_SEG_MASK (0xffULL << ORC_SEG_SHIFT) /* mmpriv.h:24 */

/* mmpriv.h:118-126: mg_log2 -- exponent from the float bits plus a quadratic in the mantissa; valid for x >= 2 */
float orc_log2(float x)
{
    union { float f; uint32_t i; } z;
    float r;
    z.f = x;
    r = (float)((int)((z.i >> 23) & 255) - 128);
    z.i &= ~(255U << 23);
    z.i += 127U << 23;
    r += (-0.34484843f * z.f + 2.02466578f) * z.f - 0.67487759f;
    return r;
}

/* lchain.c:126-135 restricted to the branch taken when !is_cdna and both anchors share a segment id */
int32_t orc_gap_penalty(int32_t dd, int32_t dg, const orc_params_t *prm)
{
    float lin = prm->chn_pen_gap * (float)dd + prm->chn_pen_skip * (float)dg;
    float lg = dd >= 1 ? orc_log2((float)(dd + 1)) : 0.0f;
    return (int32_t)(lin + .5f * lg);
}

/* lchain.c:113-138 */
int32_t orc_pair_score(const orc_anchor_t *ai, const orc_anchor_t *aj, const orc_params_t *prm)
{
    int32_t dq = (int32_t)ai->y - (int32_t)aj->y;
    int32_t seg_i = (int32_t)((ai->y & ORC_SEG_MASK) >> ORC_SEG_SHIFT);
    int32_t seg_j = (int32_t)((aj->y & ORC_SEG_MASK) >> ORC_SEG_SHIFT);
    int same = seg_i == seg_j;
    int32_t dr, dd, dg, span, sc;
    if (dq <= 0 || dq > prm->max_dist_x) return INT32_MIN;              /* :118 */
    dr = (int32_t)(ai->x - aj->x);                                       /* :119 */
    if (same && (dr == 0 || dq > prm->max_dist_y)) return INT32_MIN;     /* :120 */
    dd = dr > dq ? dr - dq : dq - dr;                                    /* :121 */
    if (same && dd > prm->bw) return INT32_MIN;                          /* :122 */
    if (prm->n_seg > 1 && !prm->is_cdna && same && dr > prm->max_dist_y) return INT32_MIN; /* :123 */
    dg = dr < dq ? dr : dq;                                              /* :124 */
    span = (int32_t)(aj->y >> 32 & 0xff);                                /* :125 */
    sc = span < dg ? span : dg;                                          /* :126 */
    if (dd || dg > span) {                                               /* :127 */
        float lin = prm->chn_pen_gap * (float)dd + prm->chn_pen_skip * (float)dg;
        float lg = dd >= 1 ? orc_log2((float)(dd + 1)) : 0.0f;
        if (prm->is_cdna || !same) {                                     /* :131 */
            if (!same && dr == 0) ++sc;
            else if (dr > dq || !same) sc -= (int)(lin < lg ? lin : lg);
            else sc -= (int)(lin + .5f * lg);
        } else sc -= (int)(lin + .5f * lg);                              /* :135 */
    }
    return sc;
}

/* lchain.c:160-207 */
int64_t orc_chain_dp(const orc_params_t *prm_in, int64_t n, const orc_anchor_t *a, int32_t *f, int64_t *p, int32_t *v)
{
    orc_params_t prm = *prm_in;
    int64_t i, j, st = 0, best_prev = -1 /* max_ii */, pairs = 0;
    int32_t *mark; /* t[] of the reference: which anchors were "seen through" a predecessor in this row */
    if (n <= 0) return 0;
    if (prm.max_dist_x < prm.bw) prm.max_dist_x = prm.bw;                        /* :160 */
    if (prm.max_dist_y < prm.bw && !prm.is_cdna) prm.max_dist_y = prm.bw;        /* :161 */
    mark = (int32_t *)calloc((size_t)n, sizeof(int32_t));
    for (i = 0; i < n; ++i) {
        int64_t arg = -1, stop_j;
        int32_t best = (int32_t)(a[i].y >> 32 & 0xff), skipped = 0;           /* :171 */
        while (st < i && (a[i].x >> 32 != a[st].x >> 32 || a[i].x > a[st].x + (uint64_t)prm.max_dist_x)) ++st; /* :172 */
        if (i - st > prm.max_iter) st = i - prm.max_iter;                        /* :173 */
        for (j = i - 1; j >= st; --j) {                                          /* :174 */
            int32_t s = orc_pair_score(&a[i], &a[j], &prm);
            ++pairs;
            if (s == INT32_MIN) continue;
            s += f[j];
            if (s > best) {
                best = s, arg = j;
                if (skipped > 0) --skipped;
            } else if (mark[j] == (int32_t)i) {
                if (++skipped > prm.max_skip) break;
            }
            if (p[j] >= 0) mark[p[j]] = (int32_t)i;
        }
        stop_j = j;                                                              /* :188 */
        if (best_prev < 0 || a[i].x - a[best_prev].x > (uint64_t)(int64_t)prm.max_dist_x) { /* :189, unsigned compare as in C's uint64 vs int64 rule */
            int32_t m = INT32_MIN;
            best_prev = -1;
            for (j = i - 1; j >= st; --j)
                if (m < f[j]) m = f[j], best_prev = j;
        }
        if (best_prev >= 0 && best_prev < stop_j) {                              /* :196 */
            int32_t s = orc_pair_score(&a[i], &a[best_prev], &prm);
            if (s != INT32_MIN && best < s + f[best_prev]) best = s + f[best_prev], arg = best_prev;
        }
        f[i] = best, p[i] = arg;                                                 /* :202 */
        if (v) v[i] = arg >= 0 && v[arg] > best ? v[arg] : best;                 /* :203 */
        if (best_prev < 0 || (a[i].x - a[best_prev].x <= (uint64_t)(int64_t)prm.max_dist_x && f[best_prev] < f[i])) /* :204 */
            best_prev = i;
    }
    free(mark);
    return pairs;
}

/* ---- ksort.h:98-151 (KRADIX_SORT_INIT(128x, mm128_t, .x, 8)) ------------------------------------------- */

#define ORC_RS_SMALL 64 /* ksort.h:98 RS_MIN_SIZE */

/* ksort.h:105-115 */
static void orc_rs_insertion(orc_anchor_t *beg, orc_anchor_t *end)
{
    orc_anchor_t *i;
    for (i = beg + 1; i < end; ++i) {
        if (i->x < (i - 1)->x) {
            orc_anchor_t *j, tmp = *i;
            for (j = i; j > beg && tmp.x < (j - 1)->x; --j) *j = *(j - 1);
            *j = tmp;
        }
    }
}

typedef struct { orc_anchor_t *cur, *end; } orc_bucket_t;

/* ksort.h:116-145: in-place American-flag pass on byte (x >> shift), then recurse on the next lower byte */
static void orc_rs_pass(orc_anchor_t *beg, orc_anchor_t *end, int shift)
{
    orc_bucket_t bk[256];
    orc_anchor_t *it;
    int k;
    for (k = 0; k < 256; ++k) bk[k].cur = bk[k].end = beg;
    for (it = beg; it != end; ++it) ++bk[it->x >> shift & 255].end;
    for (k = 1; k < 256; ++k) {
        bk[k].end += bk[k - 1].end - beg;
        bk[k].cur = bk[k - 1].end;
    }
    for (k = 0; k < 256;) {
        if (bk[k].cur != bk[k].end) {
            int l = (int)(bk[k].cur->x >> shift & 255);
            if (l != k) { /* follow the displacement cycle until an element of bucket k turns up */
                orc_anchor_t hold = *bk[k].cur, sw;
                do {
                    sw = hold;
                    hold = *bk[l].cur;
                    *bk[l].cur++ = sw;
                    l = (int)(hold.x >> shift & 255);
                } while (l != k);
                *bk[k].cur++ = hold;
            } else ++bk[k].cur;
        } else ++k;
    }
    bk[0].cur = beg;
    for (k = 1; k < 256; ++k) bk[k].cur = bk[k - 1].end;
    if (shift) {
        shift = shift > 8 ? shift - 8 : 0;
        for (k = 0; k < 256; ++k) {
            ptrdiff_t sz = bk[k].end - bk[k].cur;
            if (sz > ORC_RS_SMALL) orc_rs_pass(bk[k].cur, bk[k].end, shift);
            else if (sz > 1) orc_rs_insertion(bk[k].cur, bk[k].end);
        }
    }
}

/* ksort.h:146-150 */
void orc_radix_sort_128x(orc_anchor_t *beg, orc_anchor_t *end)
{
    if (end - beg <= ORC_RS_SMALL) orc_rs_insertion(beg, end);
    else orc_rs_pass(beg, end, 56);
}

/* ---- backtracking: lchain.c:9-76 ------------------------------------------------------------------------ */

/* lchain.c:9-25: walk back from chain end z[k]; returns where the chain is cut (exclusive) */
static int64_t orc_chain_cut(int32_t max_drop, const orc_anchor_t *z, const int32_t *f, const int64_t *p, int32_t *t, int64_t k)
{
    int64_t i = (int64_t)z[k].y, stop = -1, cut = i;
    int32_t top = 0;
    if (i < 0 || t[i] != 0) return i;
    do {
        int32_t s;
        t[i] = 2;
        stop = i = p[i];
        s = i < 0 ? (int32_t)z[k].x : (int32_t)z[k].x - f[i];
        if (s > top) top = s, cut = i;
        else if (top - s > max_drop) break;
    } while (i >= 0 && t[i] == 0);
    for (i = (int64_t)z[k].y; i >= 0 && i != stop; i = p[i]) t[i] = 0;
    return cut;
}

/* lchain.c:27-76 (the reference runs the greedy pass twice -- count, then fill; one pass gives the same u/v) */
int32_t orc_backtrack(int64_t n, const int32_t *f, const int64_t *p, int32_t min_cnt, int32_t min_sc, int32_t max_drop,
                      uint64_t *u, int32_t *v, int32_t *t, int32_t *n_v_)
{
    orc_anchor_t *z;
    int64_t i, k, n_z = 0, n_v = 0;
    int32_t n_u = 0;
    *n_v_ = 0;
    for (i = 0; i < n; ++i) if (f[i] >= min_sc) ++n_z;
    if (n_z == 0) return 0;
    z = (orc_anchor_t *)malloc((size_t)n_z * sizeof(*z));
    for (i = 0, k = 0; i < n; ++i)
        if (f[i] >= min_sc) z[k].x = (uint64_t)(int64_t)f[i], z[k++].y = (uint64_t)i; /* :38 (int32 -> uint64 sign-extends) */
    orc_radix_sort_128x(z, z + n_z);
    memset(t, 0, (size_t)n * sizeof(int32_t));
    for (k = n_z - 1; k >= 0; --k) {
        if (t[z[k].y] == 0) {
            int64_t n_v0 = n_v, cut = orc_chain_cut(max_drop, z, f, p, t, k);
            int32_t sc;
            for (i = (int64_t)z[k].y; i != cut; i = p[i]) v[n_v++] = (int32_t)i, t[i] = 1;
            sc = i < 0 ? (int32_t)z[k].x : (int32_t)z[k].x - f[i];
            if (sc >= min_sc && n_v > n_v0 && n_v - n_v0 >= min_cnt) u[n_u++] = (uint64_t)sc << 32 | (uint64_t)(n_v - n_v0);
            else n_v = n_v0;
        }
    }
    free(z);
    *n_v_ = (int32_t)n_v;
    return n_u;
}

/* lchain.c:78-111 */
void orc_compact(int32_t n_u, uint64_t *u, int32_t n_v, const int32_t *v, const orc_anchor_t *a, orc_anchor_t *b)
{
    orc_anchor_t *tmp, *w;
    uint64_t *u2;
    int64_t i, j, k;
    if (n_u <= 0) return;
    tmp = (orc_anchor_t *)malloc((size_t)(n_v > 0 ? n_v : 1) * sizeof(*tmp));
    w = (orc_anchor_t *)malloc((size_t)n_u * sizeof(*w));
    u2 = (uint64_t *)malloc((size_t)n_u * sizeof(*u2));
    for (i = 0, k = 0; i < n_u; ++i) { /* :85-89: each chain was collected end-first; flip it */
        int32_t k0 = (int32_t)k, ni = (int32_t)u[i];
        for (j = 0; j < ni; ++j) tmp[k++] = a[v[k0 + (ni - j - 1)]];
    }
    for (i = k = 0; i < n_u; ++i) {    /* :94-97 */
        w[i].x = tmp[k].x, w[i].y = (uint64_t)k << 32 | (uint64_t)i;
        k += (int32_t)u[i];
    }
    orc_radix_sort_128x(w, w + n_u);   /* :98 */
    for (i = k = 0; i < n_u; ++i) {    /* :100-105 */
        int32_t src = (int32_t)w[i].y, cnt = (int32_t)u[src];
        u2[i] = u[src];
        memcpy(&b[k], &tmp[w[i].y >> 32], (size_t)cnt * sizeof(*b));
        k += cnt;
    }
    memcpy(u, u2, (size_t)n_u * sizeof(*u));
    free(tmp); free(w); free(u2);
}

/* lchain.c:148-217 */
int32_t orc_lchain(const orc_params_t *prm, int64_t n, const orc_anchor_t *a, uint64_t *u, orc_anchor_t *b, int64_t *n_b,
                   int32_t *f_out, int64_t *p_out, int64_t *n_pairs)
{
    int32_t *f, *v, *t, n_u, n_v = 0, max_drop = prm->bw;
    int64_t *p, pairs;
    if (n_b) *n_b = 0;
    if (n_pairs) *n_pairs = 0;
    if (n <= 0 || a == 0) return 0;
    if (prm->is_cdna) max_drop = INT32_MAX;                               /* :162 */
    f = f_out ? f_out : (int32_t *)malloc((size_t)n * sizeof(int32_t));
    p = p_out ? p_out : (int64_t *)malloc((size_t)n * sizeof(int64_t));
    v = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    t = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    pairs = orc_chain_dp(prm, n, a, f, p, v);
    if (n_pairs) *n_pairs = pairs;
    n_u = orc_backtrack(n, f, p, prm->min_cnt, prm->min_score, max_drop, u, v, t, &n_v); /* :209 */
    if (n_u > 0) orc_compact(n_u, u, n_v, v, a, b);
    if (n_b) *n_b = n_u > 0 ? n_v : 0;
    if (!f_out) free(f);
    if (!p_out) free(p);
    free(v); free(t);
    return n_u;
}

/* ---- threaded batch driver (for the CPU-baseline timing only) ------------------------------------------- */

typedef struct {
    const orc_params_t *prm;
    const orc_anchor_t *a;
    const int64_t *off;
    int64_t r1;
    int64_t *next; /* shared cursor */
    int32_t *f;
    int64_t *p;
    int64_t pairs;
} orc_job_t;

static void *orc_worker(void *arg)
{
    orc_job_t *jb = (orc_job_t *)arg;
    for (;;) {
        int64_t r = __sync_fetch_and_add(jb->next, 1), n, np = 0, nb;
        uint64_t *u;
        orc_anchor_t *b;
        if (r >= jb->r1) break;
        n = jb->off[r + 1] - jb->off[r];
        if (n <= 0) continue;
        u = (uint64_t *)malloc((size_t)n * sizeof(*u));
        b = (orc_anchor_t *)malloc((size_t)n * sizeof(*b));
        orc_lchain(jb->prm, n, jb->a + jb->off[r], u, b, &nb, jb->f ? jb->f + jb->off[r] : 0, jb->p ? jb->p + jb->off[r] : 0, &np);
        jb->pairs += np;
        free(u); free(b);
    }
    return 0;
}

int64_t orc_lchain_batch(const orc_params_t *prm, const orc_anchor_t *a, const int64_t *off, int64_t r0, int64_t r1,
                         int n_threads, int32_t *f, int64_t *p)
{
    int64_t next = r0, total = 0;
    int i;
    pthread_t *tid;
    orc_job_t *jobs;
    if (n_threads < 1) n_threads = 1;
    tid = (pthread_t *)malloc((size_t)n_threads * sizeof(*tid));
    jobs = (orc_job_t *)malloc((size_t)n_threads * sizeof(*jobs));
    for (i = 0; i < n_threads; ++i) {
        jobs[i].prm = prm, jobs[i].a = a, jobs[i].off = off, jobs[i].r1 = r1, jobs[i].next = &next;
        jobs[i].f = f, jobs[i].p = p, jobs[i].pairs = 0;
        pthread_create(&tid[i], 0, orc_worker, &jobs[i]);
    }
    for (i = 0; i < n_threads; ++i) {
        pthread_join(tid[i], 0);
        total += jobs[i].pairs;
    }
    free(tid); free(jobs);
    return total;
}

/* ---- per-read digests of the whole mg_lchain_dp result, for parity checks at benchmark scale ------------------------
 * digest(w[0..m)) = C (m + 1) + sum_k w[k] (2k + 1) C  (mod 2^64), C = 0x9E3779B97F4A7C15: order-sensitive, cheap to restate
 * (tests/fake_host.c: fake_digest, bench.py: digest).  For read sel[k]: n_u, number of chain anchors, digest of u[] and of
 * the compacted anchors as 2 n_b words. */
static uint64_t orc_digest(const uint64_t *w, int64_t n)
{
    uint64_t h = 0x9E3779B97F4A7C15ULL * (uint64_t)(n + 1);
    int64_t k;
    for (k = 0; k < n; ++k) h += w[k] * ((2 * (uint64_t)k + 1) * 0x9E3779B97F4A7C15ULL);
    return h;
}

typedef struct {
    const orc_params_t *prm; const orc_anchor_t *a; const int64_t *off, *sel; int64_t n_sel; int64_t *next;
    int32_t *nu; int64_t *nb; uint64_t *hu, *hb;
} orc_djob_t;

static void *orc_digest_worker(void *arg)
{
    orc_djob_t *jb = (orc_djob_t *)arg;
    for (;;) {
        int64_t k = __sync_fetch_and_add(jb->next, 1), r, n, nb = 0, np = 0;
        int32_t nu = 0;
        uint64_t *u;
        orc_anchor_t *b;
        if (k >= jb->n_sel) break;
        r = jb->sel[k];
        n = jb->off[r + 1] - jb->off[r];
        u = (uint64_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(*u));
        b = (orc_anchor_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(*b));
        if (n > 0) nu = orc_lchain(jb->prm, n, jb->a + jb->off[r], u, b, &nb, 0, 0, &np);
        jb->nu[k] = nu; jb->nb[k] = nb;
        jb->hu[k] = orc_digest(u, nu); jb->hb[k] = orc_digest((const uint64_t *)b, 2 * nb);
        free(u); free(b);
    }
    return 0;
}

void orc_lchain_digest_batch(const orc_params_t *prm, const orc_anchor_t *a, const int64_t *off, const int64_t *sel, int64_t n_sel,
                             int n_threads, int32_t *nu, int64_t *nb, uint64_t *hu, uint64_t *hb)
{
    int64_t next = 0;
    int i;
    pthread_t *tid;
    orc_djob_t jb;
    if (n_threads < 1) n_threads = 1;
    jb.prm = prm, jb.a = a, jb.off = off, jb.sel = sel, jb.n_sel = n_sel, jb.next = &next;
    jb.nu = nu, jb.nb = nb, jb.hu = hu, jb.hb = hb;
    tid = (pthread_t *)malloc((size_t)n_threads * sizeof(*tid));
    for (i = 0; i < n_threads; ++i) pthread_create(&tid[i], 0, orc_digest_worker, &jb);
    for (i = 0; i < n_threads; ++i) pthread_join(tid[i], 0);
    free(tid);
}
