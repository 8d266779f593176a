// BatchNorm -> ReLU -> 1x1x1 Conv3D in ONE kernel (DenseNet pre-activation bottlenecks: DenseCPD's BN-ReLU-Conv(1x1x1)).
//
//   out[m, n] = act2( scale[n] * act1( sum_c relu(in_scale[c] * X[m, c] + in_shift[c]) * Wt[n, c] + bias[n] ) + shift[n] )
//
// Every dense layer normalises ALL the channels concatenated so far with its own BatchNorm, so as separate launches the
// pre-activation is an HBM pass per layer (read fp32, write bf16 hi/lo planes) that the 1x1 conv then reads back: three
// times the bytes of the concatenated tensor per layer.  Here the fp32 tensor is read ONCE: TMA stages 128 rows x 32
// channels of fp32 in shared memory, eight transform warps (thread = row x 16 channels) apply the affine + ReLU, split it into
// bf16 hi / lo and write the two K-major SWIZZLE_64B operand tiles the UMMA descriptors read, make them visible to the async
// proxy (fence.proxy.async) and hand the stage to the MMA thread.  Same arithmetic per element as affine_act_vec8_kernel
// (fmaxf(fmaf(x, s, t), 0), split_bf16) and the same MMA sequence as conv_umma_kernel's N-folded mode, so the results are
// bit-identical to the two-launch route.
//
// Warp roles: 0 = TMA producer (fp32 X tile + W hi/lo tiles per k-block of 32 channels), 1 = MMA issuer (one lane) and
// TMEM allocation, 2..9 = transform, 10..17 = epilogue (two warps per TMEM lane quadrant, as conv_umma_kernel).
// (Four transform warps -- one per SM sub-partition, a whole 32-channel row per thread -- left the kernel paced by the
// transform's dependent instruction chains: ~2100 cycles per k-block against ~800 of shared-memory traffic.)
#pragma once
#include "conv_umma.cuh"

namespace tb {

constexpr int kXfTransformWarps = 8;            // two per SM sub-partition: thread = (row, 16-channel half of the k-block)
constexpr int kXfThreads = 32 * (2 + kXfTransformWarps + kConvEpilogueWarps);
constexpr int kXfKc = 32;                       // channels per k-block: fp32 row = 128 B, bf16 row = 64 B
constexpr int kXfMaxStages = 6;                 // operand ring (A hi/lo written by the transform + W hi/lo by TMA)
constexpr int kXfMaxXStages = 10;               // fp32 ring (TMA from HBM -> transform)
constexpr int kXfMaxCin = 512;                  // in_scale / in_shift staged in shared memory

struct XformParams {
    const float* in_scale;    // [c_in] BatchNorm scale of the pre-activation
    const float* in_shift;    // [c_in]
    int32_t c_in;             // multiple of kXfKc
    // The fp32 tiles come from HBM (~2 us away) and are the only HBM reads of the kernel; their ring is separate from (and
    // deeper than) the operand ring so that a buffer is free again as soon as the transform has read it, not when the MMAs
    // of its k-block retire: x_stages * 16 KB in flight per SM instead of ~2.5 stages' worth.
    int32_t x_stages;
    // Split-plane outputs whose width is a multiple of 64 channels leave through shared memory: the epilogue warps write
    // the tile's bf16 hi / lo planes into a staging buffer (128-byte swizzled rows, conflict-free 16-byte stores) and one
    // thread issues cp.async.bulk.tensor stores (map_ohi / map_olo).  Stored from registers, every lane's 32 bytes land in
    // a different 128-byte line -- 32 wavefronts of the LSU data pipe per instruction -- and that pipe, shared with the
    // transform's shared-memory traffic, was the busiest unit of the kernel (76 %, profiles/r2v_bnrelu_ncu.csv).
    int32_t tstore;
};

#if defined(__CUDACC__)

__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// tiled 2-D store shared -> global (bulk async group of the issuing thread); rows / columns outside the tensor are dropped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// two values at once: hi = bf16_rn(v), lo = bf16_rn(v - hi) as split_bf16, packed (first value in the low half)
__device__ __forceinline__ void split_bf16x2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - h0, v1 - h1);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ConvKernelParams fields used: m_total, n_ctile_m (128-row tiles), n_tile (== N, multiple of 16, <= 128), acc_cols
// (2 * n_tile rounded up to 32), n_kblocks (= c_in / 32), stages, w_lo_rows, w_sub_bytes (n_tile * 64), the epilogue block.
template <int ACT1, int ACT2, int FMT>
__global__ void __launch_bounds__(kXfThreads, 1)
bnrelu_conv1x1_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_ohi, const __grid_constant__ CUtensorMap map_olo,
                      const ConvKernelParams p, const XformParams xp) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

    __shared__ __align__(8) uint64_t xfull_bar[kXfMaxXStages];   // TMA: the fp32 tile has landed
    __shared__ __align__(8) uint64_t xempty_bar[kXfMaxXStages];  // transform: the fp32 tile has been read
    __shared__ __align__(8) uint64_t full_bar[kXfMaxStages];     // TMA: W of the stage has landed (=> its A buffers are free)
    __shared__ __align__(8) uint64_t ready_bar[kXfMaxStages];    // transform: A hi/lo of the stage are written
    __shared__ __align__(8) uint64_t empty_bar[kXfMaxStages];    // MMA: the stage's operands have been consumed
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_in[2][kXfMaxCin];
    __shared__ __align__(16) float s_epi[3][128];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < xp.x_stages; ++s) {
            mbar_init(&xfull_bar[s], 1);
            mbar_init(&xempty_bar[s], 32 * kXfTransformWarps);
        }
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&ready_bar[s], 32 * kXfTransformWarps);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], kConvEpilogueWarps);
        }
        mbar_fence_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
        if (xp.tstore) { tma_prefetch_desc(&map_ohi); tma_prefetch_desc(&map_olo); }
    }
    if (warp == 1) tmem_alloc_512(&tmem_base_slot);
    for (int i = threadIdx.x; i < xp.c_in; i += blockDim.x) {
        s_in[0][i] = xp.in_scale[i];
        s_in[1][i] = xp.in_shift[i];
    }
    for (int i = threadIdx.x; i < p.n_tile; i += blockDim.x) {
        s_epi[0][i] = p.bias[i];
        s_epi[1][i] = p.scale[i];
        s_epi[2][i] = p.shift[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    const int total_tiles = p.n_ctile_m;
    const int tile_first = static_cast<int>(blockIdx.x), tile_step = static_cast<int>(gridDim.x);
    // shared memory: [x_stages x (X fp32 128 x 128 B)] [stages x ([A_hi 128 x 64 B][A_lo][W_hi n_tile x 64 B][W_lo])]
    constexpr uint32_t kXBytes = 128u * kXfKc * 4u, kASub = 128u * kXfKc * 2u;
    const uint32_t stage_bytes = 2u * kASub + 2u * p.w_sub_bytes;
    uint8_t* const smem_aw = smem + static_cast<size_t>(xp.x_stages) * kXBytes;
    uint8_t* const smem_out = smem_aw + static_cast<size_t>(p.stages) * stage_bytes;   // tstore: 2 planes x n_tile/64 boxes x 16 KB
    const int n_kb = p.n_kblocks;

    if (warp == 0) {
        // =============================================================== TMA producer
        // One thread feeds both rings over the flattened (tile, k-block) sequence of this CTA; the fp32 tile of item
        // i + lead is requested before the W tiles of item i (lead = x_stages - stages, so the request can only wait for
        // a transform that the W tiles already issued allow to run).
        const bool leader = elect_one();
        const int my_tiles = tile_first < total_tiles ? (total_tiles - tile_first + tile_step - 1) / tile_step : 0;
        const int64_t total = static_cast<int64_t>(my_tiles) * n_kb;
        const int lead = xp.x_stages - p.stages;
        int xt = tile_first, xk = 0, sx = 0, wk = 0, sa = 0;
        uint32_t phx = 0, pha = 0;
        for (int64_t i = -lead; i < total; ++i) {
            if (i + lead < total) {
                mbar_wait(&xempty_bar[sx], phx ^ 1u);
                if (leader) {
                    mbar_expect_tx(&xfull_bar[sx], kXBytes);
                    tma_load_2d(smem + static_cast<size_t>(sx) * kXBytes, &map_x, &xfull_bar[sx], xk * kXfKc, xt * 128);
                }
                __syncwarp();
                if (++xk == n_kb) { xk = 0; xt += tile_step; }
                if (++sx == xp.x_stages) { sx = 0; phx ^= 1u; }
            }
            if (i >= 0) {
                mbar_wait(&empty_bar[sa], pha ^ 1u);
                if (leader) {
                    uint8_t* wb = smem_aw + static_cast<size_t>(sa) * stage_bytes + 2u * kASub;
                    mbar_expect_tx(&full_bar[sa], 2u * p.w_sub_bytes);
                    tma_load_2d(wb, &map_w, &full_bar[sa], wk * kXfKc, 0);
                    tma_load_2d(wb + p.w_sub_bytes, &map_w, &full_bar[sa], wk * kXfKc, p.w_lo_rows);
                }
                __syncwarp();
                if (++wk == n_kb) wk = 0;
                if (++sa == p.stages) { sa = 0; pha ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // =============================================================== MMA issuer (one lane runs the role)
        const bool leader = elect_one();
        const uint32_t idesc = umma_idesc_bf16_m128(static_cast<uint32_t>(p.n_tile));
        const uint32_t idesc2 = umma_idesc_bf16_m128(static_cast<uint32_t>(2 * p.n_tile));
        const uint32_t desc_hi = ((64u * 8u) >> 4) | (1u << 14) | (4u << 29);      // 64-byte rows, SWIZZLE_64B
        const uint32_t lo_flags = 1u << 16;
        const uint32_t smem_base16 = (smem_u32(smem_aw) & 0x3FFFFu) >> 4;
        const uint32_t a_lo_off16 = kASub >> 4, w_off16 = (2u * kASub) >> 4;
        int s = 0;
        uint32_t ph = 0;
        int acc = 0;
        uint32_t acc_ph = 0;
        if (leader)
        for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
            mbar_wait(&tempty_bar[acc], acc_ph ^ 1u);
            tc_fence_after();
            const uint32_t d = tmem_base + static_cast<uint32_t>(acc * p.acc_cols);
            uint32_t accumulate = 0;
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(&full_bar[s], ph);         // W tiles (async proxy)
                mbar_wait(&ready_bar[s], ph);        // A tiles (generic proxy writes, fenced by the writers)
                tc_fence_after();
                const uint32_t base16 = (smem_base16 + static_cast<uint32_t>(s) * (stage_bytes >> 4)) | lo_flags;
                umma_issue_stage<1>(base16, 1, 0u, kXfKc / 16, w_off16, p.w_sub_bytes >> 4, a_lo_off16, d,
                                    d + static_cast<uint32_t>(p.n_tile), desc_hi, idesc, idesc2, accumulate);
                accumulate = 1u;
                umma_commit(&empty_bar[s]);
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
            umma_commit(&tfull_bar[acc]);
            if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
        }
        __syncwarp();
    } else if (warp < 2 + kXfTransformWarps) {
        // =============================================================== transform: thread = (row of the tile, channel half)
        const int row = ((warp - 2) & 3) * 32 + lane;
        const int chalf = (warp - 2) >> 2;                          // channels [16*chalf, 16*chalf + 16) of the k-block
        const uint32_t x_row = static_cast<uint32_t>(row) * 128u, x_swz = static_cast<uint32_t>(row & 7);
        const uint32_t a_row = static_cast<uint32_t>(row) * 64u, a_swz = static_cast<uint32_t>((row >> 1) & 3);
        int s = 0, sx = 0;
        uint32_t ph = 0, phx = 0;
        for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(&xfull_bar[sx], phx);
                mbar_wait(&full_bar[s], ph);          // the producer requested this stage's W tiles: its A buffers are free
                const uint8_t* xs = smem + static_cast<size_t>(sx) * kXBytes;
                uint8_t* st = smem_aw + static_cast<size_t>(s) * stage_bytes;
                const float* sc = &s_in[0][kb * kXfKc];
                const float* sh = &s_in[1][kb * kXfKc];
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {                    // 8 channels -> one 16-byte chunk of each bf16 plane
                    const int j = 2 * chalf + jj;
                    const float4 x0 = *reinterpret_cast<const float4*>(xs + x_row + (((2u * j) ^ x_swz) << 4));
                    const float4 x1 = *reinterpret_cast<const float4*>(xs + x_row + (((2u * j + 1u) ^ x_swz) << 4));
                    const float4 s0 = *reinterpret_cast<const float4*>(sc + 8 * j), s1 = *reinterpret_cast<const float4*>(sc + 8 * j + 4);
                    const float4 h0 = *reinterpret_cast<const float4*>(sh + 8 * j), h1 = *reinterpret_cast<const float4*>(sh + 8 * j + 4);
                    const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                    const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                    const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float v0 = fmaxf(fmaf(xv[2 * e], sv[2 * e], hv[2 * e]), 0.0f);
                        const float v1 = fmaxf(fmaf(xv[2 * e + 1], sv[2 * e + 1], hv[2 * e + 1]), 0.0f);
                        split_bf16x2(v0, v1, hi[e], lo[e]);
                    }
                    const uint32_t a_off = a_row + ((static_cast<uint32_t>(j) ^ a_swz) << 4);
                    *reinterpret_cast<uint4*>(st + a_off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(st + kASub + a_off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                mbar_arrive(&xempty_bar[sx]);         // the fp32 tile is in registers / consumed: TMA may refill it
                fence_proxy_async_smem();             // generic-proxy stores -> visible to the tensor core's async-proxy reads
                mbar_arrive(&ready_bar[s]);
                if (++s == p.stages) { s = 0; ph ^= 1u; }
                if (++sx == xp.x_stages) { sx = 0; phx ^= 1u; }
            }
        }
    } else {
        // =============================================================== epilogue (the last eight warps)
        const int quad = warp & 3;
        const int half = (warp - 2 - kXfTransformWarps) >> 2;
        const int row_in_tile = quad * 32 + lane;
        const int chunks = p.n_tile / 16;
        const bool issuer = warp == 2 + kXfTransformWarps && lane == 0;      // issues (and waits for) the bulk stores
        int acc = 0;
        uint32_t acc_ph = 0;
        for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
            mbar_wait_relaxed(&tfull_bar[acc], acc_ph);
            tc_fence_after();
            const int64_t m = static_cast<int64_t>(tile) * 128 + row_in_tile;
            const bool row_ok = m < p.m_total;
            const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * p.acc_cols);
            if (xp.tstore) {
                // the previous tile's bulk stores have finished reading the staging buffer
                if (issuer) tma_store_wait_read();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kConvEpilogueWarps) : "memory");
            }
            for (int c = half; c < chunks; c += 2) {
                uint32_t r[16], rc[16];
                __syncwarp();
                tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(c * 16), r);
                tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(p.n_tile + c * 16), rc);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    r[i] = __float_as_uint(fmaf(__uint_as_float(r[i]), p.acc_comp, __uint_as_float(rc[i])));
                const int n0 = c * 16;
                if (n0 >= p.c_store) continue;
                if (!xp.tstore) {
                    epilogue_chunk<ACT1, ACT2, FMT>(p, r, n0, m, row_ok, s_epi[0], s_epi[1], s_epi[2]);
                    continue;
                }
                float v[16];
                epilogue_math16<ACT1, ACT2>(p, r, n0, s_epi[0], s_epi[1], s_epi[2], v);
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) split_bf16x2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
                // box = 64 channels x 128 rows, 128-byte rows, 16-byte unit u of row r at u ^ (r & 7)
                const uint32_t box = static_cast<uint32_t>(c >> 2), u0 = static_cast<uint32_t>((c & 3) * 2);
                const uint32_t rsw = static_cast<uint32_t>(row_in_tile & 7);
                uint8_t* bh = smem_out + box * 16384u + static_cast<uint32_t>(row_in_tile) * 128u;
                uint8_t* bl = bh + static_cast<uint32_t>(p.n_tile >> 6) * 16384u;
                *reinterpret_cast<uint4*>(bh + ((u0 ^ rsw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(bh + (((u0 + 1u) ^ rsw) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                *reinterpret_cast<uint4*>(bl + ((u0 ^ rsw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<uint4*>(bl + (((u0 + 1u) ^ rsw) << 4)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (xp.tstore) {
                fence_proxy_async_smem();             // this thread's staging stores -> visible to the bulk-copy engine
                asm volatile("bar.sync 2, %0;" ::"n"(32 * kConvEpilogueWarps) : "memory");
                if (issuer) {
                    const int boxes = p.n_tile >> 6;
                    for (int b = 0; b < boxes; ++b) {
                        tma_store_2d(&map_ohi, smem_out + static_cast<uint32_t>(b) * 16384u, b * 64, tile * 128);
                        tma_store_2d(&map_olo, smem_out + static_cast<uint32_t>(boxes + b) * 16384u, b * 64, tile * 128);
                    }
                    tma_store_commit();
                }
            }
            if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
        }
        if (xp.tstore && issuer) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_512(tmem_base);
    }
}

#endif  // __CUDACC__

}  // namespace tb
