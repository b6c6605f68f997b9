/* fake_host.c -- stands in for the minimap2 host driver in the drop-in protocol tests (tests/test_dropin.py).
 * Provides the callbacks libmm2gb_plchain.a expects from the driver (include/mm2gb_plchain.h): a counting malloc-backed
 * kmalloc/kfree, build_misc returning a preset Misc, and a post_chaining_helper that records frag_gap like map.c:483. */
#include <stdlib.h>
#include <string.h>

#include "mm2gb_plchain.h"

static Misc_abi g_misc;
static long g_live, g_helper_calls;

void fake_set_misc(const Misc_abi *m) { g_misc = *m; }
long fake_live_blocks(void) { return g_live; }
long fake_helper_calls(void) { return g_helper_calls; }

void *kmalloc(void *km, size_t size)
{
    (void)km;
    if (size == 0) return 0; /* kalloc.c returns NULL for empty requests */
    ++g_live;
    return malloc(size);
}
void kfree(void *km, void *ptr)
{
    (void)km;
    if (!ptr) return;
    --g_live;
    free(ptr);
}
Misc_abi build_misc(const mm2gb_idx_t *mi, const mm2gb_mapopt_t *opt, const int64_t qlen_sum, const int n_seg)
{
    (void)mi; (void)opt; (void)qlen_sum; (void)n_seg;
    return g_misc;
}
void post_chaining_helper(const mm2gb_idx_t *mi, const mm2gb_mapopt_t *opt, mm2gb_chain_read_t *read, Misc_abi misc, void *km)
{
    (void)mi; (void)opt; (void)km;
    read->frag_gap = misc.max_dist_x; /* map.c:483 */
    ++g_helper_calls;
}
/* allocate a read's anchor array from the fake arena */
mm2gb_anchor_t *fake_alloc_anchors(const mm2gb_anchor_t *src, int64_t n)
{
    mm2gb_anchor_t *a = (mm2gb_anchor_t *)kmalloc(0, (size_t)n * sizeof(*a));
    if (a) memcpy(a, src, (size_t)n * sizeof(*a));
    return a;
}
