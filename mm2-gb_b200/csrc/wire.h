// wire.h -- the compressed upload format of a batch of anchors ("packed wire format"), host side.
//
// An mm128_t anchor (minimap.h:72) is x = rev<<63 | rid<<32 | rpos, y = seg_id<<48 | flags<<40 | q_span<<32 | qpos.  Inside
// a read the anchors are sorted by x (map.c:329), so the high word of x only changes between rid/strand runs, and the high
// word of y (segment id, seed flags, q_span) is the same for almost every seed of a read.  The boundary has to touch every
// anchor once anyway -- the driver hands over per-read kmalloc'd (pageable) arrays that must be gathered into pinned
// staging (the reference does the same pass as an AoS->SoA repack, gpu/plmem.cu:154-198) -- so that pass writes 8 bytes per
// anchor instead of 16:
//     pk[i]      = low32(x_i) | low32(y_i) << 32                       one per anchor
//     runs[r]    = { first anchor of the run, high32(x), high32(y) }   one per maximal run of equal high words
//     blk_run[b] = the run that holds anchor 256 b                     so that the device finds a run without a global search
// laid out back to back in ONE staging buffer (one H2D copy): [pk : 8 n][blk_run : 4 (n_blk + 1), padded to 16][runs : 16 n_runs].
// k_expand (chain_kernels.cuh) rebuilds the 16-byte anchors in HBM.  Lossless for every input; when the high words change so
// often that the run list would not fit (HPC seeds with a q_span per anchor) the packer says so and the batch goes up raw.
#pragma once

#include <stddef.h>
#include <stdint.h>

#include "../../include/mm2gb_chain.h"

namespace mm2gb {

struct WireRun { int32_t start; uint32_t x_hi, y_hi, pad; };   // 16 bytes, read as one uint4 on the device

constexpr int kWireBlock = 256;                                 // anchors per blk_run entry (= threads per k_expand block)

struct WireLayout {
    size_t pk_off, blk_off, run_off;    // byte offsets inside the staging buffer
    int64_t n, n_blk;
    int run_cap;                        // run entries that fit behind blk_run in a buffer of `bytes_cap` bytes
};

// layout of a batch of n anchors inside a staging buffer of bytes_cap bytes; run_cap <= 0 means "does not fit: send raw"
WireLayout wire_layout(int64_t n, size_t bytes_cap);

class WirePacker {
public:
    // base = staging buffer (32-byte aligned), L = wire_layout(n, capacity)
    void begin(void *base, const WireLayout &L);
    // append the anchors of one read (any piece of the flat batch); false = run list full (caller falls back to raw)
    bool add(const mm2gb_anchor_t *a, int64_t n);
    // fill blk_run; returns the number of bytes of the staging buffer to upload
    size_t finish();
    int n_runs() const { return n_runs_; }

private:
    uint64_t *pk_ = nullptr;
    int32_t *blk_ = nullptr;
    WireRun *runs_ = nullptr;
    WireLayout L_{};
    int64_t fill_ = 0;
    int n_runs_ = 0;
    uint32_t cur_x_ = 0, cur_y_ = 0;
};

// b[k] = a[v[k]] for k in [0, n): compact_a's gather (lchain.c:100-105) from a read's own anchor array
void wire_gather(const mm2gb_anchor_t *a, const int32_t *v, int64_t n, mm2gb_anchor_t *b);

} // namespace mm2gb
