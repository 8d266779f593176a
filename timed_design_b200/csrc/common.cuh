// Shared device/host helpers for libtimed_b200.so (sm_100a only).
// Thin inline-PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor) and tcgen05 (UMMA/TMEM).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace tb {

// ----------------------------------------------------------------------------- errors (host)
void set_error(const std::string& msg);          // defined in api.cu (thread-local)
#define TB_CHECK_CUDA(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            tb::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e));    \
            return -2;                                                                   \
        }                                                                                \
    } while (0)
#define TB_REQUIRE(cond, msg)                                                            \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            tb::set_error(std::string("invalid argument: ") + (msg));                    \
            return -1;                                                                   \
        }                                                                                \
    } while (0)

// Role-timing switches (skip copies / MMA issue / epilogue / stores) are compiled in only for bring-up builds
// (-DTIMED_B200_DEBUG, tools/role_timing.sh); in the release library the tests below are constant false.
#ifdef TIMED_B200_DEBUG
#define TB_DBG(mask, bit) (((mask) & (bit)) != 0)
#else
#define TB_DBG(mask, bit) (false)
#endif

// ----------------------------------------------------------------------------- device PTX
#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a protocol bug must surface as a CUDA error (trap), never as a hung GPU.
#ifndef TB_MBAR_SPIN_LIMIT
#define TB_MBAR_SPIN_LIMIT (1u << 24)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > TB_MBAR_SPIN_LIMIT) {
            printf("timed_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

// Same for waits that are expected to be long (an epilogue warp waiting for a whole tile's mainloop): sleep between
// polls so the spinning warps do not burn issue slots and power next to a power-capped tensor pipe.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(spins < 8 ? 32 : 256);
        if (++spins > TB_MBAR_SPIN_LIMIT) {
            printf("timed_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

__device__ __forceinline__ int32_t ld_acquire_gpu(const int32_t* p) {
    int32_t v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// tiled 2-D load: coordinates {c0 (fastest), c1}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// im2col 5-D load: base pixel {c, w, h, d, n} (already offset by the lower corner),
// filter-tap offsets {w, h, d}.
__device__ __forceinline__ void tma_load_im2col_5d(void* smem_dst, const CUtensorMap* m,
                                                   uint64_t* bar, int32_t c, int32_t w, int32_t h,
                                                   int32_t d, int32_t n, uint16_t ow, uint16_t oh,
                                                   uint16_t od) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], {%8, %9, %10};" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(d),
        "r"(n), "h"(ow), "h"(oh), "h"(od)
        : "memory");
}

// tiled 2-D load multicast to the CTAs of `mask` in this cluster: data and the mbarrier's
// complete_tx land at the same CTA-relative offsets in every destination CTA
__device__ __forceinline__ void tma_load_2d_mcast(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                                  int32_t c0, int32_t c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc_512(uint32_t* smem_dst) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                     smem_u32(smem_dst))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t taddr) {     // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (bf16 operands, fp32 accumulate), one issuing thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives once all tcgen05 ops previously issued by this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
            smem_u32(bar))
        : "memory");
}
// same, arriving on the barrier at this CTA-relative offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(mask)
        : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 16 consecutive columns (fp32 bit patterns).
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory matrix descriptor, K-major operand tile whose rows are `row_bytes`
// (= swizzle span: 128/64/32) wide; 8-row groups are `row_bytes*8` apart (SBO).
// layout_type: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B (cute::UMMA::LayoutType).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t row_bytes,
                                                   uint32_t layout_type) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);            // start address  [0,14)
    d |= static_cast<uint64_t>(1) << 16;                               // LBO (ignored for swizzled K-major)
    d |= static_cast<uint64_t>((row_bytes * 8) >> 4) << 32;            // SBO            [32,46)
    d |= static_cast<uint64_t>(1) << 46;                               // descriptor version (sm_100)
    d |= static_cast<uint64_t>(layout_type) << 61;                     // swizzle mode   [61,64)
    return d;
}
// Instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n (multiple of 16, <=256).
__device__ __host__ __forceinline__ uint32_t umma_idesc_bf16_m128(uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// fp32 -> (hi, lo) bf16 pair with hi + lo ~= v to ~2^-17 relative.
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return static_cast<uint32_t>(__bfloat16_as_ushort(a)) |
           (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
}

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_ELU = 2, ACT_SIGMOID = 3, ACT_TANH = 4 };

__device__ __forceinline__ float apply_act(float x, int act, float alpha) {
    switch (act) {
        case ACT_RELU: return fmaxf(x, 0.0f);
        case ACT_ELU: return x > 0.0f ? x : alpha * expm1f(x);
        case ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
        case ACT_TANH: return tanhf(x);
        default: return x;
    }
}
#endif  // __CUDACC__

}  // namespace tb
