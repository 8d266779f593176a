// Microbenchmark: issue rate / execution time of tcgen05.mma (kind::f16, M=128, cta_group::1) for a given
// N and shared-memory operand layout, one CTA per SM.  Operands are whatever is in shared memory (zeros);
// only timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I timed_design_b200/csrc tools/mma_probe.cu -o tools/mma_probe
//   tools/mma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "conv_umma.cuh"

using namespace tb;

struct ProbeCase {
    int n;            // UMMA N
    int layout;       // 2 = SW128, 4 = SW64, 6 = SW32, 0 = no swizzle (LBO = 16 B aliasing as thin_conv)
    int n2;           // number of accumulators used round-robin (1, 2, 4, 8)
    int reps;         // MMA (pairs) per commit
    int a_rotate;     // unused
    int a_off16;      // extra A start offset in 16-byte units (alignment of the core matrices)
    int lbo16;        // A leading-dimension offset in 16-byte units (no-swizzle layout only)
};

__global__ void __launch_bounds__(128, 1) mma_probe_kernel(ProbeCase c, int iters, long long* out_cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc_512(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        const bool leader = elect_one();
        const uint32_t row_bytes = c.layout == 2 ? 128 : c.layout == 4 ? 64 : c.layout == 6 ? 32 : 16;
        const uint32_t desc_hi = c.layout ? (((row_bytes * 8u) >> 4) | (1u << 14) | (static_cast<uint32_t>(c.layout) << 29))
                                          : ((128u >> 4) | (1u << 14));
        const uint32_t lbo = 1u << 16;
        const uint32_t a_lbo = static_cast<uint32_t>(c.lbo16 ? c.lbo16 : 1) << 16;
        const uint32_t a_tile16 = (128u * (c.layout ? row_bytes : 16u) + 1024u) >> 4;      // distinct A tiles
        const uint32_t a_base = (smem_u32(smem) & 0x3FFFFu) >> 4;
        const uint32_t b_base = a_base + (96u * 1024u >> 4);
        const uint32_t idesc = umma_idesc_bf16_m128(c.n);
        const uint32_t idesc2 = umma_idesc_bf16_m128(c.n2 ? c.n2 : 16);
        uint32_t ph = 0;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int r0 = 0; r0 < c.reps; r0 += 8) {
                if (leader) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const uint32_t a = (a_base + static_cast<uint32_t>(r & 3) * a_tile16 + static_cast<uint32_t>(c.a_off16)) | a_lbo;   // 4 distinct A tiles
                    const uint32_t b = b_base | lbo;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                        "setp.ne.b32 p, %5, 0;\n\t"
                        "mov.b64 da, {%1, %3};\n\t"
                        "mov.b64 db, {%2, %3};\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem + static_cast<uint32_t>(r % 8 % c.n2) * static_cast<uint32_t>(c.n)),
                        "r"(a), "r"(b), "r"(desc_hi), "r"(idesc), "r"(1u) : "memory");
                }
                }
                __syncwarp();
            }
            if (leader) umma_commit(&bar);
            __syncwarp();
            mbar_wait(&bar, ph);
            ph ^= 1u;
        }
        long long t1 = clock64();
        if (blockIdx.x == 0 && leader) out_cycles[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc_512(tmem); }
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 8);
    cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    std::vector<ProbeCase> cases;
    for (int n : {128, 64})
        for (int off : {0, 1, 2, 4, 7})
            for (int lbo : {1, 570}) cases.push_back({n, 0, 2, 64, 4, off, lbo});
    for (int n : {128, 64}) for (int off : {0, 2, 4}) cases.push_back({n, 2, 2, 64, 4, off * 8, 0});   // SW128: shift whole rows
    printf("%6s %6s %6s %6s %6s | %12s %14s\n", "N", "layout", "N2", "reps", "rot", "cyc/commit", "cyc/MMA(pair)");
    for (const ProbeCase& c : cases) {
        const int iters = 200;
        for (int rep = 0; rep < 2; ++rep) {
            mma_probe_kernel<<<148, 128, 200 * 1024>>>(c, iters, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        }
        long long cyc = 0;
        cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
        printf("%6d %6d %6d %6d off %3d lbo %4d | %12.1f %14.1f\n", c.n, c.layout, c.n2, c.reps, c.a_off16, c.lbo16,
               static_cast<double>(cyc) / iters, static_cast<double>(cyc) / iters / c.reps);
    }
    return 0;
}
