// wire.cpp -- host side of the packed upload format (see wire.h): the gather pass of the boundary writes 8 bytes per anchor.
// Plain C++ (compiled by the host compiler through nvcc); an AVX2 body is selected at run time when the CPU has it.
#include "wire.h"

#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace mm2gb {

WireLayout wire_layout(int64_t n, size_t bytes_cap)
{
    WireLayout L;
    L.n = n;
    L.n_blk = (n + kWireBlock - 1) / kWireBlock;
    L.pk_off = 0;
    L.blk_off = (size_t)n * 8;
    L.blk_off = (L.blk_off + 15) & ~(size_t)15;
    L.run_off = L.blk_off + (((size_t)(L.n_blk + 1) * 4 + 15) & ~(size_t)15);
    const size_t left = bytes_cap > L.run_off ? bytes_cap - L.run_off : 0;
    const size_t cap = left / sizeof(WireRun);
    L.run_cap = cap > (size_t)INT32_MAX ? INT32_MAX : (int)cap;
    return L;
}

void WirePacker::begin(void *base, const WireLayout &L)
{
    L_ = L;
    pk_ = reinterpret_cast<uint64_t *>(static_cast<char *>(base) + L.pk_off);
    blk_ = reinterpret_cast<int32_t *>(static_cast<char *>(base) + L.blk_off);
    runs_ = reinterpret_cast<WireRun *>(static_cast<char *>(base) + L.run_off);
    fill_ = 0;
    n_runs_ = 0;
    cur_x_ = cur_y_ = 0;
}

namespace {

struct PackState { uint64_t *pk; WireRun *runs; int64_t fill; int n_runs, run_cap; uint32_t cx, cy; };

inline bool pack_one(PackState &s, const mm2gb_anchor_t &a)
{
    const uint32_t xh = (uint32_t)(a.x >> 32), yh = (uint32_t)(a.y >> 32);
    if (xh != s.cx || yh != s.cy || s.n_runs == 0) {
        if (s.n_runs >= s.run_cap) return false;
        WireRun r;
        r.start = (int32_t)s.fill; r.x_hi = xh; r.y_hi = yh; r.pad = 0;
        s.runs[s.n_runs++] = r;
        s.cx = xh; s.cy = yh;
    }
    s.pk[s.fill++] = (uint64_t)(uint32_t)a.x | (a.y << 32);
    return true;
}

bool pack_scalar(PackState &s, const mm2gb_anchor_t *a, int64_t n)
{
    for (int64_t i = 0; i < n; ++i)
        if (!pack_one(s, a[i])) return false;
    return true;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) bool pack_avx2(PackState &s, const mm2gb_anchor_t *a, int64_t n)
{
    int64_t i = 0;
    // scalar until the destination is 32-byte aligned (non-temporal stores: the staging buffer is read next by the DMA
    // engine, not by this core)
    while (i < n && ((reinterpret_cast<uintptr_t>(s.pk + s.fill) & 31) != 0 || s.n_runs == 0))
        if (!pack_one(s, a[i++])) return false;
    while (i + 4 <= n) {
        const __m256i a0 = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(a + i));
        const __m256i a1 = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(a + i + 2));
        const __m256i cur = _mm256_set_epi32((int)s.cy, (int)s.cx, (int)s.cy, (int)s.cx, (int)s.cy, (int)s.cx, (int)s.cy, (int)s.cx);
        const __m256i h0 = _mm256_shuffle_epi32(a0, 0xDD), h1 = _mm256_shuffle_epi32(a1, 0xDD);   // x.hi y.hi x.hi y.hi per anchor
        const __m256i eq = _mm256_and_si256(_mm256_cmpeq_epi32(h0, cur), _mm256_cmpeq_epi32(h1, cur));
        if (_mm256_movemask_epi8(eq) != -1) {    // a run boundary among these four: take them one by one
            for (int k = 0; k < 4; ++k)
                if (!pack_one(s, a[i + k])) return false;
            i += 4;
            while (i < n && (reinterpret_cast<uintptr_t>(s.pk + s.fill) & 31) != 0)   // (fill moved by 4: still aligned; kept for safety)
                if (!pack_one(s, a[i++])) return false;
            continue;
        }
        const __m256i l0 = _mm256_shuffle_epi32(a0, 0x88), l1 = _mm256_shuffle_epi32(a1, 0x88);   // x.lo y.lo x.lo y.lo per anchor
        __m256i u = _mm256_unpacklo_epi64(l0, l1);       // anchors 0 2 | 1 3
        u = _mm256_permute4x64_epi64(u, 0xd8);           // anchors 0 1 2 3
        _mm256_stream_si256(reinterpret_cast<__m256i *>(s.pk + s.fill), u);
        s.fill += 4;
        i += 4;
    }
    _mm_sfence();
    for (; i < n; ++i)
        if (!pack_one(s, a[i])) return false;
    return true;
}
#endif

bool have_avx2()
{
#if defined(__x86_64__)
    static const bool ok = __builtin_cpu_supports("avx2");
    return ok;
#else
    return false;
#endif
}

} // namespace

bool WirePacker::add(const mm2gb_anchor_t *a, int64_t n)
{
    if (n <= 0) return true;
    if (fill_ + n > L_.n) return false;
    PackState s{pk_, runs_, fill_, n_runs_, L_.run_cap, cur_x_, cur_y_};
    bool ok;
#if defined(__x86_64__)
    ok = have_avx2() ? pack_avx2(s, a, n) : pack_scalar(s, a, n);
#else
    ok = pack_scalar(s, a, n);
#endif
    fill_ = s.fill; n_runs_ = s.n_runs; cur_x_ = s.cx; cur_y_ = s.cy;
    return ok;
}

size_t WirePacker::finish()
{
    // blk_run[b] = last run starting at or before anchor 256 b; one extra entry for the block after the last
    int r = 0;
    for (int64_t b = 0; b <= L_.n_blk; ++b) {
        const int64_t g = b * kWireBlock;
        while (r + 1 < n_runs_ && runs_[r + 1].start <= g) ++r;
        blk_[b] = r;
    }
    return L_.run_off + (size_t)n_runs_ * sizeof(WireRun);
}

void wire_gather(const mm2gb_anchor_t *a, const int32_t *v, int64_t n, mm2gb_anchor_t *b)
{
    int64_t k = 0;
#if defined(__x86_64__)
    if ((reinterpret_cast<uintptr_t>(b) & 15) == 0 && n >= 64) {
        // non-temporal stores: the result array is written once here and read much later by the driver, so it should not pull
        // its own cache lines in first (read-for-ownership) nor push the source lines out
        for (; k + 4 <= n; k += 4) {   // four independent 16-byte moves in flight
            const __m128i t0 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(a + v[k]));
            const __m128i t1 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(a + v[k + 1]));
            const __m128i t2 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(a + v[k + 2]));
            const __m128i t3 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(a + v[k + 3]));
            _mm_stream_si128(reinterpret_cast<__m128i *>(b + k), t0);
            _mm_stream_si128(reinterpret_cast<__m128i *>(b + k + 1), t1);
            _mm_stream_si128(reinterpret_cast<__m128i *>(b + k + 2), t2);
            _mm_stream_si128(reinterpret_cast<__m128i *>(b + k + 3), t3);
        }
        _mm_sfence();
    }
    for (; k + 4 <= n; k += 4) {
        const __m128i t0 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(a + v[k]));
        const __m128i t1 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(a + v[k + 1]));
        const __m128i t2 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(a + v[k + 2]));
        const __m128i t3 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(a + v[k + 3]));
        _mm_storeu_si128(reinterpret_cast<__m128i *>(b + k), t0);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(b + k + 1), t1);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(b + k + 2), t2);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(b + k + 3), t3);
    }
#endif
    for (; k < n; ++k) b[k] = a[v[k]];
}

} // namespace mm2gb
