# The change of the reference's map.c that routes seeding + chaining through the device (INTEGRATION.md section 5), as a sed script:
#   sed -f integration/map_gpu_seed.sed /root/reference/map.c > map_gpu_seed.c
# 1. seeding (map.c:1001): with MM2GB_GPU_SEED=1 the per-read host seeding is skipped (only qlen_sum is filled in);
# 2. chaining of a full batch on the non-gpu-chain path (map.c:1060-1062): one fused device call for the whole batch instead of
#    mm_map_chain per read.
# 3. reads per batch on that path (map.c:1313): N_ACCUM = 64 is sized for host chaining; the device wants more per launch
#    (MM2GB_SEED_BATCH_READS, default 512).
# Everything else (batch hand-over, mm_map_align, output) is untouched.
s|^\([[:space:]]*\)mm_map_seed(s->p->mi, s->p->opt, read_ptr, b, km);|\1if (mm2gb_glue_enabled()) mm2gb_glue_defer_seed(read_ptr); else mm_map_seed(s->p->mi, s->p->opt, read_ptr, b, km);|
s|^\([[:space:]]*\)for (iread=0; iread<tr->acc_batch.count; iread++) {|\1if (mm2gb_glue_enabled()) mm2gb_glue_seed_chain_batch(s->p->mi, s->p->opt, tr->acc_batch.reads, tr->acc_batch.count, tid, tr->acc_batch.km); else for (iread=0; iread<tr->acc_batch.count; iread++) {|
s|^\([[:space:]]*\)s->batch_max_reads = N_ACCUM;|\1s->batch_max_reads = mm2gb_glue_enabled() ? mm2gb_glue_batch_reads() : N_ACCUM;|
/^#include "ksort.h"/i\
int mm2gb_glue_enabled(void); int mm2gb_glue_batch_reads(void); void mm2gb_glue_defer_seed(chain_read_t *rd); void mm2gb_glue_seed_chain_batch(const mm_idx_t *mi, const mm_mapopt_t *opt, chain_read_t *reads, int n_reads, int tid, void *km);
