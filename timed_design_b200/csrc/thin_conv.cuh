// First-layer ("thin input") Conv3D on tcgen05: C_in <= 8.
//
// The generic kernel gathers one 128-pixel im2col column per filter tap; for a 6-channel input that
// is 27 (or, W-folded, 9 four-times-overlapping) TMA boxes of 32-64 byte rows per tile, and the layer
// ends up bound by L2->SM bandwidth and TMA row rate with the tensor pipe ~12 % busy
// (profiles/r1_summary.md).  Here the input is stored as a zero-padded volume with 8 channels (16
// bytes) per pixel, so for a run of 128 consecutive output positions of one (frame, z) plane the
// operand of filter row (kd, kh) is ONE contiguous span of 128+kw-1 stored pixels: a plain 1-D bulk
// copy, every input byte fetched once per (kd, kh).  The kw taps are not materialised at all: a
// no-swizzle K-major UMMA descriptor with a leading-dimension byte offset of 16 bytes makes the
// second K-half of row r alias the first K-half of row r+1, i.e. one K=16 MMA step consumes pixels
// (r, r+1) x 8 channels straight from the span.  A left-over odd tap is paired with the same tap of
// the next filter row (LBO = distance between the two spans).  Weights (a few tens of KB) stay
// resident in shared memory for the whole kernel.  Rows that fall into the W margin of the padded
// volume (Wp - Wo of every Wp positions) are computed and dropped by the epilogue.
//
// Two MMAs per K step instead of three: the resident weights hold, per K chunk, the n_tile hi rows
// followed by the n_tile lo rows, so A_hi x [W_hi | W_lo] is ONE MMA of N = 2*n_tile writing the main
// product to TMEM columns [0,n) and the A_hi*W_lo correction to [n,2n); A_lo x W_hi accumulates into
// [n,2n) as well (a thin-N MMA costs about the same whatever N is).  Keeping the corrections out of the
// main accumulator also removes two thirds of the accumulator truncation error; the epilogue adds them.
#pragma once
#include "common.cuh"
#include "conv_umma.cuh"

namespace tb {

constexpr int kThinMaxSteps = 64;

struct ThinConvParams {
    // ---- tiling: tiles of 128 consecutive positions t = p*Wp + wp inside one (frame, z) plane
    int32_t n_tiles_total;    // n_frames * Do * tiles_per_plane
    int32_t tiles_per_plane;  // ceil(Ho*Wp / 128)
    int32_t Do, Ho, Wo, Wp;
    // ---- input addressing (bytes); hi plane at in_hi, lo plane lo_plane_off further
    const uint8_t* in_hi;
    int64_t lo_plane_off;
    int64_t frame_bytes;      // Dp*Hp*Wp*16
    int64_t dplane_bytes;     // Hp*Wp*16
    int32_t row_bytes;        // Wp*16
    int32_t off_d, off_h, off_w;   // stored-coordinate offset of this conv's window origin
    int32_t kd, kh, kw;
    int32_t span_bytes;       // bytes copied per (kd,kh) span (multiple of 16)
    int32_t span_stride;      // smem distance between consecutive spans
    // ---- K=16 steps (the kernel regenerates the order of thin_plan_create with two affine loops)
    int32_t n_steps;
    // ---- weights: [2*n_steps K-chunks][2*n_tile rows: hi then lo][8] bf16, resident in smem
    const uint8_t* w_packed;
    uint32_t w_bytes;
    int32_t n_tile;           // output channels per tile (UMMA N = 2*n_tile for the folded MMA)
    int32_t acc_cols;         // TMEM columns per accumulator stage (2*n_tile rounded up to 32)
    int32_t acc_stages;
    int32_t stages;
    // ---- epilogue (as ConvKernelParams)
    const float* bias;
    const float* scale;
    const float* shift;
    int32_t act1, act2;
    float alpha1, alpha2;
    float* out_f32;
    __nv_bfloat16* out_hi;
    __nv_bfloat16* out_lo;
    int64_t ldc;
    int32_t c_store;
    int32_t dbg;              // TIMED_B200_DBG role timing: 1 skip copies, 2 skip MMAs, 4 skip epilogue
};

#if defined(__CUDACC__)

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// generic UMMA issue with explicit per-operand descriptor halves
__device__ __forceinline__ void umma_bf16_desc(bool leader, uint32_t d_tmem, uint32_t a_lo32, uint32_t a_hi32,
                                               uint32_t b_lo32, uint32_t b_hi32, uint32_t idesc,
                                               uint32_t accumulate) {
    if (leader) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "mov.b64 da, {%1, %2};\n\t"
            "mov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
            "r"(a_lo32), "r"(a_hi32), "r"(b_lo32), "r"(b_hi32), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}

template <int ACT1, int ACT2, int FMT>
__global__ void __launch_bounds__(kConvThreads, 1)
thin_conv_kernel(const __grid_constant__ ThinConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));

    __shared__ __align__(8) uint64_t full_bar[kConvMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kConvMaxStages];
    __shared__ __align__(8) uint64_t tfull_bar[4];
    __shared__ __align__(8) uint64_t tempty_bar[4];
    __shared__ __align__(8) uint64_t w_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_epi[3][256];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 4; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], kConvEpilogueWarps);
        }
        mbar_init(&w_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_512(&tmem_base_slot);
    for (int i = threadIdx.x; i < p.n_tile; i += blockDim.x) {
        s_epi[0][i] = p.bias[i];
        s_epi[1][i] = p.scale[i];
        s_epi[2][i] = p.shift[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    const int taps = p.kd * p.kh;
    const uint32_t w_bytes = p.w_bytes;
    uint8_t* w_smem = smem;                                            // resident weights
    uint8_t* stage0 = smem + ((w_bytes + 127u) & ~127u);
    const uint32_t plane_region = static_cast<uint32_t>(taps) * p.span_stride;   // hi spans, then lo spans
    const uint32_t stage_bytes = 2u * plane_region;

    if (warp == 0) {
        // =============================================================== bulk-copy producer
        const bool leader = elect_one();
        if (leader) {
            mbar_expect_tx(&w_bar, w_bytes);
            for (uint32_t off = 0; off < w_bytes; off += 16384u)
                bulk_load_1d(w_smem + off, p.w_packed + off, min(16384u, w_bytes - off), &w_bar);
        }
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles_total; tile += gridDim.x) {
            const int plane = tile / p.tiles_per_plane;
            const int t0 = (tile - plane * p.tiles_per_plane) * 128;
            const int nf = plane / p.Do;
            const int z = plane - nf * p.Do;
            mbar_wait(&empty_bar[s], ph ^ 1u);
            if (leader && TB_DBG(p.dbg, 1)) {
                mbar_arrive(&full_bar[s]);
            } else if (leader) {
                mbar_expect_tx(&full_bar[s], 2u * static_cast<uint32_t>(taps) * p.span_bytes);
                uint8_t* st = stage0 + static_cast<size_t>(s) * stage_bytes;
                const uint8_t* src0 = p.in_hi + nf * p.frame_bytes + static_cast<int64_t>(z + p.off_d) * p.dplane_bytes +
                                      static_cast<int64_t>(p.off_h) * p.row_bytes +
                                      static_cast<int64_t>(t0 + p.off_w) * 16;
                int tap = 0;
                for (int a = 0; a < p.kd; ++a)
                    for (int b = 0; b < p.kh; ++b, ++tap) {
                        const uint8_t* src = src0 + a * p.dplane_bytes + static_cast<int64_t>(b) * p.row_bytes;
                        bulk_load_1d(st + tap * p.span_stride, src, p.span_bytes, &full_bar[s]);
                        bulk_load_1d(st + plane_region + tap * p.span_stride, src + p.lo_plane_off, p.span_bytes,
                                     &full_bar[s]);
                    }
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        // =============================================================== MMA issuer
        const bool leader = elect_one();
        const uint32_t idesc = umma_idesc_bf16_m128(static_cast<uint32_t>(p.n_tile));
        const uint32_t idesc2 = umma_idesc_bf16_m128(static_cast<uint32_t>(2 * p.n_tile));
        // no-swizzle K-major descriptors: hi word = SBO>>4 | version<<14 (layout type 0)
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t w_lbo16 = (static_cast<uint32_t>(2 * p.n_tile) * 16u) >> 4;   // K-chunk stride of W
        const uint32_t w_step16 = 2u * w_lbo16;
        const uint32_t w_base16 = (smem_u32(w_smem) & 0x3FFFFu) >> 4;
        const uint32_t stage0_16 = (smem_u32(stage0) & 0x3FFFFu) >> 4;
        const uint32_t plane16 = plane_region >> 4;
        const uint32_t ss16 = static_cast<uint32_t>(p.span_stride) >> 4;
        const int pairs = p.kw >> 1;
        mbar_wait(&w_bar, 0);
        int s = 0;
        uint32_t ph = 0;
        int acc = 0;
        uint32_t acc_ph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles_total; tile += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_ph ^ 1u);
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t d_main = tmem_base + static_cast<uint32_t>(acc * p.acc_cols);
            const uint32_t d_corr = d_main + static_cast<uint32_t>(p.n_tile);
            const uint32_t a_base16 = stage0_16 + static_cast<uint32_t>(s) * (stage_bytes >> 4);
            // K steps in the packing order of thin_plan_create, as two affine loops (no descriptor table:
            // an indexed constant load per step put ~110 cycles of dependent latency between MMAs):
            //   (a) filter row r, in-row pixel pair (2jp, 2jp+1): start = span r + 2jp pixels, LBO = one pixel
            //   (b) left-over odd tap of rows (r, r+1): start = span r + (kw-1) pixels, LBO = span stride
            if (leader && !TB_DBG(p.dbg, 2)) {
                uint32_t b = w_base16 | (w_lbo16 << 16);
                uint32_t accumulate = 0;
                uint32_t a_row = a_base16;
                for (int r = 0; r < taps; ++r, a_row += ss16) {
#pragma unroll 2
                    for (int jp = 0; jp < pairs; ++jp) {
                        const uint32_t a_hi = (a_row + 2u * static_cast<uint32_t>(jp)) | (1u << 16);
                        umma_bf16_desc(true, d_main, a_hi, desc_hi, b, desc_hi, idesc2, accumulate);
                        umma_bf16_desc(true, d_corr, a_hi + plane16, desc_hi, b, desc_hi, idesc, 1u);
                        accumulate = 1u;
                        b += w_step16;
                    }
                }
                if (p.kw & 1) {
                    a_row = a_base16 + static_cast<uint32_t>(p.kw - 1);
                    for (int r = 0; r < taps; r += 2, a_row += 2u * ss16) {
                        const uint32_t a_hi = a_row | ((r + 1 < taps ? ss16 : 1u) << 16);
                        umma_bf16_desc(true, d_main, a_hi, desc_hi, b, desc_hi, idesc2, accumulate);
                        umma_bf16_desc(true, d_corr, a_hi + plane16, desc_hi, b, desc_hi, idesc, 1u);
                        accumulate = 1u;
                        b += w_step16;
                    }
                }
            }
            if (leader) {
                umma_commit(&empty_bar[s]);
                umma_commit(&tfull_bar[acc]);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1u; }
            if (++acc == p.acc_stages) { acc = 0; acc_ph ^= 1u; }
        }
    } else {
        // =============================================================== epilogue (warps 2..9)
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int chunks = p.n_tile / 16;
        const int plane_positions = p.Ho * p.Wp;
        int acc = 0;
        uint32_t acc_ph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles_total; tile += gridDim.x) {
            const int plane = tile / p.tiles_per_plane;
            const int t = (tile - plane * p.tiles_per_plane) * 128 + quad * 32 + lane;
            const int prow = t / p.Wp;
            const int wp = t - prow * p.Wp;
            const bool row_ok = t < plane_positions && wp < p.Wo;
            const int64_t m = (static_cast<int64_t>(plane) * p.Ho + prow) * p.Wo + wp;
            mbar_wait(&tfull_bar[acc], acc_ph);
            tc_fence_after();
            const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                   static_cast<uint32_t>(acc * p.acc_cols);
            for (int c = half; c < chunks && !TB_DBG(p.dbg, 4); c += 2) {
                uint32_t r[16], rc[16];
                __syncwarp();
                tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(c * 16), r);
                tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(p.n_tile + c * 16), rc);
                tmem_ld_wait();
                const int n0 = c * 16;
                if (n0 >= p.c_store) continue;
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float x = (__uint_as_float(r[i]) + __uint_as_float(rc[i])) + s_epi[0][n0 + i];
                    x = act_ct<ACT1>(x, p.act1, p.alpha1);
                    x = fmaf(x, s_epi[1][n0 + i], s_epi[2][n0 + i]);
                    v[i] = act_ct<ACT2>(x, p.act2, p.alpha2);
                }
                if (row_ok && FMT == FMT_SPLIT) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        __nv_bfloat16 h0, l0, h1, l1;
                        split_bf16(v[2 * i], h0, l0);
                        split_bf16(v[2 * i + 1], h1, l1);
                        hi[i] = pack_bf16x2(h0, h1);
                        lo[i] = pack_bf16x2(l0, l1);
                    }
                    uint4* dh = reinterpret_cast<uint4*>(p.out_hi + m * p.ldc + n0);
                    uint4* dl = reinterpret_cast<uint4*>(p.out_lo + m * p.ldc + n0);
                    dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                    dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                } else if (row_ok) {
                    float* dst = p.out_f32 + m * p.ldc + n0;
                    if (n0 + 16 <= p.c_store && (p.ldc & 3) == 0) {
                        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (n0 + i < p.c_store) dst[i] = v[i];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == p.acc_stages) { acc = 0; acc_ph ^= 1u; }
        }
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
