/* hostbw_probe.c -- how fast can T host threads move anchors?  Decides where the host passes of the boundary may live.
 *   copy : memcpy of 16-byte anchors (what a gather into pinned staging costs)
 *   pack : 16 B -> 8 B (low words of x and y, non-temporal stores), the compressed upload format
 *   gath : b[k] = a[v[k]] with mostly sequential v (what compact_a's gather costs on the result side)
 * build: gcc -O3 -mavx2 -pthread tools/hostbw_probe.c -o /tmp/hostbw_probe ; run: /tmp/hostbw_probe [MB]
 */
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { uint64_t x, y; } anchor_t;
static anchor_t *A, *B;
static uint64_t *P;
static int32_t *V;
static size_t N;
static int T, MODE;

static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static void *work(void *arg)
{
    const int t = (int)(intptr_t)arg;
    const size_t lo = N * t / T, hi = N * (t + 1) / T;
    if (MODE == 0) memcpy(B + lo, A + lo, (hi - lo) * sizeof(anchor_t));
    else if (MODE == 1) {
        size_t i = lo;
        for (; i + 4 <= hi; i += 4) {
            __m256i a0 = _mm256_loadu_si256((const __m256i *)(A + i)), a1 = _mm256_loadu_si256((const __m256i *)(A + i + 2));
            /* low dwords of the four qwords of each */
            __m256i s0 = _mm256_shuffle_epi32(a0, 0x88), s1 = _mm256_shuffle_epi32(a1, 0x88);
            __m256i u = _mm256_unpacklo_epi64(s0, s1);           /* lanes: [a0.lo128 | a1.lo128], [a0.hi128 | a1.hi128] */
            u = _mm256_permute4x64_epi64(u, 0xd8);
            _mm256_stream_si256((__m256i *)(P + i), u);
        }
        for (; i < hi; ++i) P[i] = (A[i].x & 0xffffffffu) | (A[i].y << 32);
        _mm_sfence();
    } else {
        for (size_t i = lo; i < hi; ++i) B[i] = A[V[i]];
    }
    return 0;
}

int main(int argc, char **argv)
{
    const size_t mb = argc > 1 ? (size_t)atol(argv[1]) : 512;
    N = mb * 1024 * 1024 / 16;
    A = aligned_alloc(64, N * 16); B = aligned_alloc(64, N * 16); P = aligned_alloc(64, N * 8); V = aligned_alloc(64, N * 4);
    for (size_t i = 0; i < N; ++i) { A[i].x = i * 3; A[i].y = i * 7; B[i].x = 0; P[i] = 0; V[i] = (int32_t)(i < 3 ? i : i - (i % 7 == 0 ? 3 : 0)); }
    static const char *names[] = {"copy16", "pack16to8", "gather16"};
    for (MODE = 0; MODE < 3; ++MODE)
        for (T = 1; T <= 32; T *= 2) {
            double best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                pthread_t th[64];
                const double t0 = now();
                for (int t = 0; t < T; ++t) pthread_create(&th[t], 0, work, (void *)(intptr_t)t);
                for (int t = 0; t < T; ++t) pthread_join(th[t], 0);
                const double dt = now() - t0;
                if (dt < best) best = dt;
            }
            printf("{\"mode\": \"%s\", \"threads\": %d, \"anchors\": %zu, \"ms\": %.3f, \"input_GBps\": %.2f}\n", names[MODE], T, N, best * 1e3,
                   N * 16 / best / 1e9);
            fflush(stdout);
        }
    return 0;
}
