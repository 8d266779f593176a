// Microbenchmark: cycles per repetition of a PATTERN of tcgen05.mma instructions (kind::f16, M = 128,
// cta_group::1, no-swizzle K-major operands as slab_conv / thinz_conv use them), one CTA per SM.  Each pattern
// entry is (N, accumulator column, A tile): the question is how much column overlap, accumulator switching and
// N cost when MMAs follow each other, i.e. which issue order a kernel should use.  Operands are zeros.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I timed_design_b200/csrc tools/mma_pattern_probe.cu -o tools/mma_pattern_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "conv_umma.cuh"
#include "thin_conv.cuh"

using namespace tb;

struct Pattern {
    int len;
    int n[8], col[8], a_tile[8], b_tile[8];
    const char* name;
};

template <int LEN>
__global__ void __launch_bounds__(320, 1) pattern_kernel(Pattern c, int reps, int iters, int unroll, long long* out_cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t bar2;
    __shared__ __align__(8) uint64_t bar3;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc_512(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const bool poll_mode = unroll >= 20;
    if (poll_mode) unroll = 3;
    if (poll_mode && warp >= 2) {
        if (blockDim.x > 128) mbar_wait(&bar3, 0);
    } else if (warp == 1 || (warp == 2 && unroll == 4)) {
        const bool leader = elect_one();
        uint64_t* my_bar = warp == 1 ? &bar : &bar2;
        const int half = warp - 1;
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t a_base = (smem_u32(smem) & 0x3FFFFu) >> 4;
        const uint32_t b_base = a_base + (96u * 1024u >> 4);
        uint32_t idesc[8], a[8], b[8], d[8];
        for (int i = 0; i < 8; ++i) {
            idesc[i] = umma_idesc_bf16_m128(c.n[i] ? c.n[i] : 16);
            a[i] = (a_base + static_cast<uint32_t>(c.a_tile[i]) * 128u) | (570u << 16);     // slab-like: K halves 570 rows apart
            b[i] = (b_base + static_cast<uint32_t>(c.b_tile[i]) * 512u) | (256u << 16);
            d[i] = tmem + static_cast<uint32_t>(c.col[i]);
        }
        uint32_t ph = 0;
        long long t0 = clock64();
        if (leader) {
            for (int it = 0; it < iters; ++it) {
                if (unroll == 1) {
                    for (int r = 0; r < reps; ++r) {
#pragma unroll
                        for (int i = 0; i < LEN; ++i) umma_bf16_desc(true, d[i], a[i], desc_hi, b[i], desc_hi, idesc[i], 1u);
                    }
                } else if (unroll == 2) {
                    for (int r = 0; r < reps; r += 2) {
#pragma unroll
                        for (int u = 0; u < 2; ++u)
#pragma unroll
                            for (int i = 0; i < LEN; ++i) umma_bf16_desc(true, d[i], a[i], desc_hi, b[i], desc_hi, idesc[i], 1u);
                    }
                } else if (unroll == 4) {
                    uint32_t adv = 0;
                    for (int r = 0; r < reps; ++r, adv = (adv + 2u) & 63u) {
#pragma unroll
                        for (int i = 0; i < LEN; ++i)
                            if ((i * 2) / LEN == half) umma_bf16_desc(true, d[i], a[i] + adv, desc_hi, b[i] + adv, desc_hi, idesc[i], 1u);
                    }
                } else if (unroll >= 10) {
                    // as 3, plus a tcgen05.commit to an unwatched barrier every (unroll - 10 + 1) reps
                    uint32_t adv = 0;
                    const int every = unroll - 9;
                    int cnt = 0;
                    for (int r = 0; r < reps; ++r, adv = (adv + 2u) & 63u) {
#pragma unroll
                        for (int i = 0; i < LEN; ++i) umma_bf16_desc(true, d[i], a[i] + adv, desc_hi, b[i] + adv, desc_hi, idesc[i], 1u);
                        if (++cnt == every) { cnt = 0; umma_commit(&bar2); }
                    }
                } else {
                    // descriptors advance every rep as in a real K loop (address arithmetic between the MMAs)
                    uint32_t adv = 0;
                    for (int r = 0; r < reps; ++r, adv = (adv + 2u) & 63u) {
#pragma unroll
                        for (int i = 0; i < LEN; ++i) umma_bf16_desc(true, d[i], a[i] + adv, desc_hi, b[i] + adv, desc_hi, idesc[i], 1u);
                    }
                }
                umma_commit(my_bar);
                mbar_wait(my_bar, ph);
                ph ^= 1u;
            }
        }
        __syncwarp();
        if (poll_mode && leader) mbar_arrive(&bar3);
        long long t1 = clock64();
        if (blockIdx.x == 0 && leader && warp == 1) out_cycles[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc_512(tmem); }
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 8);
    cudaFuncSetAttribute(pattern_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(pattern_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(pattern_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(pattern_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    std::vector<Pattern> ps = {
        {2, {128, 64}, {0, 64}, {0, 1}, {0, 0}, "slab mt=1: main N128 @0, corr N64 @64"},
        {4, {128, 64, 128, 64}, {0, 64, 128, 192}, {0, 1, 2, 3}, {0, 0, 0, 0}, "slab mt=2 (current order)"},
        {4, {128, 128, 64, 64}, {0, 128, 64, 192}, {0, 2, 1, 3}, {0, 0, 0, 0}, "slab mt=2, mains then corrections"},
        {4, {128, 64, 128, 64}, {0, 256, 128, 320}, {0, 1, 2, 3}, {0, 0, 0, 0}, "mt=2, corrections in disjoint columns"},
        {2, {128, 64}, {0, 256}, {0, 1}, {0, 0}, "mt=1, correction in disjoint columns"},
        {1, {128}, {0}, {0}, {0}, "N128 same accumulator"},
        {2, {128, 128}, {0, 128}, {0, 1}, {0, 0}, "N128 two accumulators"},
        {1, {64}, {0}, {0}, {0}, "N64 same accumulator"},
        {2, {64, 64}, {0, 64}, {0, 1}, {0, 0}, "N64 two accumulators"},
        {4, {64, 64, 64, 64}, {0, 64, 128, 192}, {0, 1, 2, 3}, {0, 0, 0, 0}, "N64 four accumulators"},
        {1, {192}, {0}, {0}, {0}, "N192 same accumulator"},
        {2, {192, 192}, {0, 256}, {0, 1}, {0, 0}, "N192 two accumulators"},
        {3, {192, 192, 192}, {0, 192, 192}, {0, 0, 1}, {0, 1, 0}, "kw-in-N: main N192 @0, two corrections N192 @192"},
        {1, {256}, {0}, {0}, {0}, "N256 same accumulator"},
        {2, {256, 256}, {0, 256}, {0, 1}, {0, 0}, "N256 two accumulators"},
        {2, {192, 160}, {0, 32}, {0, 1}, {0, 1}, "thinz conv1: N192 @0 then N160 @32"},
        {2, {192, 160}, {0, 256}, {0, 1}, {0, 1}, "thinz-like, disjoint columns"},
        {2, {128, 64}, {0, 64}, {0, 0}, {0, 0}, "mt=1 pattern, same A tile"},
        {4, {128, 64, 128, 64}, {0, 64, 0, 64}, {0, 1, 2, 3}, {0, 0, 1, 1}, "mt=1, two K steps"},
    };
    printf("modes: u1 one lane, constant descriptors; u3 one lane, descriptors advance every rep; u4 two warps, each issuing half\n"
           "of the pattern; u10 as u3 plus a tcgen05.commit per rep; u20 as u3 with 8 more warps polling an mbarrier.\n"
           "model: cycles(MMA) = max(N/2 [math], 32 + N/4 [128 B/cycle of shared-memory operand reads: 4 KB of A + 32 N bytes of B])\n");
    printf("%-58s | %10s %10s %10s\n", "pattern", "cyc/rep", "cyc/MMA", "model");
    for (int unroll : {1, 3, 4, 10, 20})
    for (const Pattern& c : ps) {
        const int reps = 54, iters = 100;
        for (int rep = 0; rep < 2; ++rep) {
            switch (c.len) {
                case 1: pattern_kernel<1><<<148, unroll >= 20 ? 320 : 128, 200 * 1024>>>(c, reps, iters, unroll, d_out); break;
                case 2: pattern_kernel<2><<<148, unroll >= 20 ? 320 : 128, 200 * 1024>>>(c, reps, iters, unroll, d_out); break;
                case 3: pattern_kernel<3><<<148, unroll >= 20 ? 320 : 128, 200 * 1024>>>(c, reps, iters, unroll, d_out); break;
                default: pattern_kernel<4><<<148, unroll >= 20 ? 320 : 128, 200 * 1024>>>(c, reps, iters, unroll, d_out); break;
            }
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        }
        long long cyc = 0;
        cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
        const double per = static_cast<double>(cyc) / iters / reps;
        double model = 0;
        for (int i = 0; i < c.len; ++i) model += c.n[i] / 2.0 > 32 + c.n[i] / 4.0 ? c.n[i] / 2.0 : 32 + c.n[i] / 4.0;
        printf("u%-2d %-55s | %10.1f %10.1f %10.1f\n", unroll, c.name, per, per / c.len, model);
    }
    return 0;
}
