// int_peak.cu -- measured integer / issue ceilings of one B200 for the chaining score kernel's roofline (bench.py reads
// profiles/int_peak.json, which is this program's output).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/int_peak tools/int_peak.cu && /tmp/int_peak
//
// Three kernels, each at the score kernel's own residency (6 CTAs of 4 warps per SM, persistent, 512-record ring per warp):
//   mid_mix   the MID loop of score_unit_packed itself (packed_walk<R, MODE_MID, false> from chain_kernels.cuh: warp-uniform
//             LDS.64 of the record, IADD3, ISETP, predicated LDS.U8 of the penalty, IMAD, predicated VIMNMX) on registers and
//             shared memory that stay resident -- the speed of light of a score kernel that did nothing but MID pairs;
//   alu_mix   independent IADD3 / LOP3 / IMNMX chains only (the alu pipe alone);
//   issue_mix independent IADD3 + IMAD chains, half and half (alu + fma pipes together: the issue ceiling).
// Output: thread-instructions/s for each, pairs/s and instructions per pair for mid_mix, SM clock sampled by the caller.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mm2-gb_b200/csrc/chain_kernels.cuh"

using namespace mm2gb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int R = 512;

__global__ void __launch_bounds__(kScoreWarps * 32, 6)
k_mid_mix(int iters, int bw, int maxd_q, const unsigned char *__restrict__ lut_g, int lut_n, int *__restrict__ out)
{
    __shared__ __align__(16) unsigned char lut[2 * kLutMax + 16];
    extern __shared__ int4 smem_raw[];
    RecP *ring = reinterpret_cast<RecP *>(smem_raw) + (threadIdx.x >> 5) * R;
    const unsigned lut_s = (unsigned)__cvta_generic_to_shared(lut);
    const int lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < lut_n; k += blockDim.x) lut[k] = lut_g[k];
    for (int k = lane; k < R; k += 32) {   // records on a drifting diagonal: most pairs are inside the band, as in a real chain
        RecP r;
        r.e = 1000 + (k * 37) % 300 - 150; r.g = (k * 15) << kSlotBits | k; r.y = 10 * k; r.q = 15;
        ring[k] = r;
    }
    __syncthreads();
    int thr = 0, pen = 0;
    const int D0 = 1000 + bw + lane * 3;
    for (int it = 0; it < iters; ++it)
        packed_walk<R, MODE_MID, false>(thr, pen, 0, R, 0, ring, D0 + (it & 63), 0, maxd_q, (unsigned)bw, 2u * (unsigned)bw, lut_s, 0);
    if (thr == 0x7fffffff) out[blockIdx.x * blockDim.x + threadIdx.x] = thr + pen;
}

template <int MODE>
__global__ void __launch_bounds__(kScoreWarps * 32, 6)
k_chains(int iters, int seed, int *__restrict__ out)
{
    int a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = seed + q * 7 + threadIdx.x;
    int m = seed | 1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (MODE == 0) {            // alu pipe only: add / xor / max in rotation
                    if ((q + rep) % 3 == 0) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[q]) : "r"(m));
                    else if ((q + rep) % 3 == 1) asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[q]) : "r"(m));
                    else asm volatile("max.s32 %0, %0, %1;" : "+r"(a[q]) : "r"(m));
                } else {                    // alu + fma pipes: add and mad alternate
                    if (q & 1) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(a[q]) : "r"(m));
                    else asm volatile("add.s32 %0, %0, %1;" : "+r"(a[q]) : "r"(m));
                }
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s ^= a[q];
    if (s == 0x12345678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float host_log2(float x)
{
    union { float f; unsigned i; } z;
    z.f = x;
    float r = (float)((int)((z.i >> 23) & 255) - 128);
    z.i &= ~(255U << 23);
    z.i += 127U << 23;
    return r + ((-0.34484843f * z.f + 2.02466578f) * z.f - 0.67487759f);
}

int main()
{
    int dev = 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int n_sm = prop.multiProcessorCount;
    const int bw = 500, maxd_q = 5000;
    std::vector<unsigned char> lut(2 * bw + 1);
    for (int dd = 0; dd <= bw; ++dd) {
        const int pen = (int)(0.12f * (float)dd + .5f * (dd >= 1 ? host_log2((float)(dd + 1)) : 0.f));
        lut[bw + dd] = lut[bw - dd] = (unsigned char)pen;
    }
    unsigned char *d_lut;
    int *d_out;
    CK(cudaMalloc(&d_lut, lut.size()));
    CK(cudaMemcpy(d_lut, lut.data(), lut.size(), cudaMemcpyHostToDevice));
    const size_t smem = (size_t)kScoreWarps * R * sizeof(RecP);
    CK(cudaFuncSetAttribute(k_mid_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_mid_mix, kScoreWarps * 32, smem));
    const int grid = nb * n_sm, threads = kScoreWarps * 32;
    CK(cudaMalloc(&d_out, (size_t)grid * threads * sizeof(int)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](auto launch) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        return best;
    };
    const int it_mid = 400, it_ch = 4000;
    const float ms_mid = timeit([&] { k_mid_mix<<<grid, threads, smem>>>(it_mid, bw, maxd_q, d_lut, (int)lut.size(), d_out); });
    const float ms_alu = timeit([&] { k_chains<0><<<grid, threads>>>(it_ch, 3, d_out); });
    const float ms_iss = timeit([&] { k_chains<1><<<grid, threads>>>(it_ch, 3, d_out); });
    const double pairs = (double)grid * threads * it_mid * R;
    const double mid_instr_per_pair = 6.6;   // LDS.64, IADD3 / IMAD.IADD, ISETP, LDS.U8, IMAD, VIMNMX + 5 instructions of loop overhead per 8 pairs (cuobjdump -sass)
    const double ch_instr = (double)grid * threads * it_ch * 64.0;
    printf("{\"gpu\": \"%s\", \"n_sm\": %d, \"resident_ctas_per_sm\": %d, \"warps_per_sm\": %d, "
           "\"mid_mix\": {\"pairs_per_s\": %.4e, \"ms\": %.3f, \"instr_per_pair\": %.1f, \"thread_instr_per_s\": %.4e}, "
           "\"alu_mix\": {\"thread_instr_per_s\": %.4e, \"ms\": %.3f}, \"issue_mix\": {\"thread_instr_per_s\": %.4e, \"ms\": %.3f}, "
           "\"nominal_issue_per_s_at_1965MHz\": %.4e}\n",
           prop.name, n_sm, nb, nb * kScoreWarps, pairs / (ms_mid * 1e-3), ms_mid, mid_instr_per_pair, mid_instr_per_pair * pairs / (ms_mid * 1e-3),
           ch_instr / (ms_alu * 1e-3), ms_alu, ch_instr / (ms_iss * 1e-3), ms_iss, (double)n_sm * 128 * 1.965e9);
    return 0;
}
