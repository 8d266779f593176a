// First-layer ("thin input", C_in <= 8) Conv3D with the kd taps folded into the MMA N dimension.
//
// thin_conv_kernel issues, per 128 output positions, kd*kh*kw/2 K steps of two MMAs with N = 2*n_tile and
// N = n_tile (64 and 32 for TIMED's first block).  A tcgen05.mma of M = 128 costs ~70 cycles however small N
// is (tools/mma_probe.cu), so that kernel sits at its MMA floor with the tensor pipe mostly idle.  Here one
// CTA tile covers the same 128 in-plane positions of `zt` consecutive OUTPUT planes.  An input plane i then
// contributes to the output planes i, i-1, ..., i-(kd-1) through the filter slices kd' = 0, 1, ..., kd-1 --
// with the SAME A operand (the plane's (kh, kw) window).  Stacking those filter slices along N,
//     B1 = [ W_hi(kd-1) | W_lo(kd-1) | ... | W_hi(0) | W_lo(0) ]      (kd * 2*n_tile rows)
// turns them into ONE MMA of N = kd*2*n_tile whose destination columns are the (main | correction) accumulators
// of kd adjacent output planes.  A second MMA adds A_lo x [ W_hi(kd-1) | 0 | ... | W_hi(0) ] into the correction
// columns.  Per input plane that is kh*kw/2 K steps of two ~100-cycle MMAs serving kd output planes, and every
// input plane is copied into shared memory once per (zt + kd - 1)/zt output planes instead of kd times.
// Input layout, K-step construction (pixel pairs aliased by a 16-byte LBO, left-over odd taps paired across
// filter rows) and epilogue are those of thin_conv.cuh.
#pragma once
#include "common.cuh"
#include "conv_umma.cuh"
#include "thin_conv.cuh"

namespace tb {

struct ThinZParams {
    // ---- tiling: tile = (frame, z group of zt output planes, window of 128 in-plane positions u = p*Wp + q)
    int32_t n_tiles_total;
    int32_t z_groups;         // ceil(Do / zt)
    int32_t windows;          // ceil(Ho*Wp / 128)
    int32_t zt;
    int32_t Do, Ho, Wo, Wp;
    // ---- input (padded volume, 16 bytes per stored pixel; hi plane then lo plane)
    const uint8_t* in_hi;
    int64_t lo_plane_off;
    int64_t frame_bytes;
    int64_t dplane_bytes;
    int32_t off_d;            // stored plane of (output plane 0, kd 0)
    int32_t off_hw;           // stored in-plane position of (output row 0, col 0, kh 0, kw 0)
    int32_t kd, kh, kw;
    int32_t span_bytes;       // bytes copied per input plane (multiple of 16)
    int32_t span_stride;
    // ---- resident weights: per K step [B1: 2 K-chunks x b1_rows x 16 B][B2: 2 K-chunks x b2_rows x 16 B]
    const uint8_t* w_packed;
    uint32_t w_bytes;
    int32_t n_steps;          // K steps per input plane
    int32_t b1_rows, b2_rows; // kd*2*n_tile, (2*kd-1)*n_tile
    int32_t n_tile;
    int32_t acc_cols;         // TMEM columns per accumulator stage (zt*2*n_tile rounded up to 32)
    int32_t acc_stages;
    int32_t stages;
    ConvKernelParams epi;     // epilogue fields
    int32_t dbg;
};

#if defined(__CUDACC__)

template <int ACT1, int ACT2, int FMT>
__global__ void __launch_bounds__(kConvThreads, 1)
thinz_conv_kernel(const __grid_constant__ ThinZParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));

    __shared__ __align__(8) uint64_t full_bar[kConvMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kConvMaxStages];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ __align__(8) uint64_t w_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_epi[3][128];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], kConvEpilogueWarps);
        }
        mbar_init(&w_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_512(&tmem_base_slot);
    for (int i = threadIdx.x; i < p.n_tile; i += blockDim.x) {
        s_epi[0][i] = p.epi.bias[i];
        s_epi[1][i] = p.epi.scale[i];
        s_epi[2][i] = p.epi.shift[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    uint8_t* w_smem = smem;
    uint8_t* stage0 = smem + ((p.w_bytes + 127u) & ~127u);
    const int max_planes = p.zt + p.kd - 1;
    const uint32_t plane_region = static_cast<uint32_t>(max_planes) * p.span_stride;   // hi spans, then lo spans
    const uint32_t stage_bytes = 2u * plane_region;
    const int tiles_per_frame = p.z_groups * p.windows;

    if (warp == 0) {
        // =============================================================== bulk-copy producer
        const bool leader = elect_one();
        if (leader) {
            mbar_expect_tx(&w_bar, p.w_bytes);
            for (uint32_t off = 0; off < p.w_bytes; off += 16384u)
                bulk_load_1d(w_smem + off, p.w_packed + off, min(16384u, p.w_bytes - off), &w_bar);
        }
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles_total; tile += gridDim.x) {
            const int nf = tile / tiles_per_frame;
            const int r = tile - nf * tiles_per_frame;
            const int zg = r / p.windows;
            const int win = r - zg * p.windows;
            const int z0 = zg * p.zt;
            const int n_planes = min(p.zt, p.Do - z0) + p.kd - 1;
            mbar_wait(&empty_bar[s], ph ^ 1u);
            if (leader && (p.dbg & 1)) {
                mbar_arrive(&full_bar[s]);
            } else if (leader) {
                mbar_expect_tx(&full_bar[s], 2u * static_cast<uint32_t>(n_planes) * p.span_bytes);
                uint8_t* st = stage0 + static_cast<size_t>(s) * stage_bytes;
                const uint8_t* src = p.in_hi + nf * p.frame_bytes + static_cast<int64_t>(z0 + p.off_d) * p.dplane_bytes +
                                     static_cast<int64_t>(p.off_hw + win * 128) * 16;
                for (int i = 0; i < n_planes; ++i, src += p.dplane_bytes) {
                    bulk_load_1d(st + i * p.span_stride, src, p.span_bytes, &full_bar[s]);
                    bulk_load_1d(st + plane_region + i * p.span_stride, src + p.lo_plane_off, p.span_bytes, &full_bar[s]);
                }
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        // =============================================================== MMA issuer
        const bool leader = elect_one();
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);            // no swizzle, SBO = 128 B
        const uint32_t n2 = static_cast<uint32_t>(2 * p.n_tile);
        const uint32_t w_base16 = (smem_u32(w_smem) & 0x3FFFFu) >> 4;
        const uint32_t stage0_16 = (smem_u32(stage0) & 0x3FFFFu) >> 4;
        const uint32_t plane16 = plane_region >> 4;
        const uint32_t ss16 = static_cast<uint32_t>(p.span_stride) >> 4;
        const uint32_t b1_lbo = static_cast<uint32_t>(p.b1_rows), b2_lbo = static_cast<uint32_t>(p.b2_rows);   // rows*16 B >> 4
        const uint32_t step16 = 2u * (b1_lbo + b2_lbo);
        const int pairs = p.kw >> 1;
        mbar_wait(&w_bar, 0);
        int s = 0, acc = 0;
        uint32_t ph = 0, acc_ph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles_total; tile += gridDim.x) {
            const int r = tile % tiles_per_frame;
            const int z0 = (r / p.windows) * p.zt;
            const int zt_eff = min(p.zt, p.Do - z0);
            const int n_planes = zt_eff + p.kd - 1;
            mbar_wait(&tempty_bar[acc], acc_ph ^ 1u);
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (leader && !(p.dbg & 2)) {
                const uint32_t d_tile = tmem_base + static_cast<uint32_t>(acc * p.acc_cols);
                uint32_t a_plane = stage0_16 + static_cast<uint32_t>(s) * (stage_bytes >> 4);
                for (int i = 0; i < n_planes; ++i, a_plane += ss16) {
                    // output planes fed by input plane i: j_lo..j_hi through filter slices kd' = i - j
                    const int j_hi = min(zt_eff - 1, i);
                    const int j_lo = max(0, i - (p.kd - 1));
                    const uint32_t cnt = static_cast<uint32_t>(j_hi - j_lo + 1);
                    const uint32_t row0 = static_cast<uint32_t>(p.kd - 1 - (i - j_lo)) * n2;     // first B row (both blocks)
                    const uint32_t d_lo = d_tile + static_cast<uint32_t>(j_lo) * n2;
                    const bool fresh = i < zt_eff;              // output plane i gets its first contribution (kd' = 0)
                    const uint32_t idesc1 = umma_idesc_bf16_m128(cnt * n2);
                    const uint32_t idesc1b = umma_idesc_bf16_m128(cnt > 1 ? (cnt - 1) * n2 : n2);
                    const uint32_t idesc_f = umma_idesc_bf16_m128(n2);
                    const uint32_t idesc2 = umma_idesc_bf16_m128(cnt * n2 - static_cast<uint32_t>(p.n_tile));
                    uint32_t b = w_base16;
                    bool first = true;
                    auto issue = [&](uint32_t a_hi) {
                        const uint32_t b1 = (b + row0) | (b1_lbo << 16);
                        const uint32_t b2 = (b + 2u * b1_lbo + row0) | (b2_lbo << 16);
                        if (first && fresh) {
                            // the kd' = 0 block initialises output plane i's (main | correction) columns
                            const uint32_t bf = (b + static_cast<uint32_t>(p.kd - 1) * n2) | (b1_lbo << 16);
                            umma_bf16_desc(true, d_tile + static_cast<uint32_t>(i) * n2, a_hi, desc_hi, bf, desc_hi, idesc_f, 0u);
                            if (cnt > 1) umma_bf16_desc(true, d_lo, a_hi, desc_hi, b1, desc_hi, idesc1b, 1u);
                        } else {
                            umma_bf16_desc(true, d_lo, a_hi, desc_hi, b1, desc_hi, idesc1, 1u);
                        }
                        umma_bf16_desc(true, d_lo + p.n_tile, a_hi + plane16, desc_hi, b2, desc_hi, idesc2, 1u);
                        first = false;
                        b += step16;
                    };
                    uint32_t a_row = a_plane;
                    for (int rr = 0; rr < p.kh; ++rr, a_row += static_cast<uint32_t>(p.Wp))
                        for (int jp = 0; jp < pairs; ++jp)
                            issue((a_row + 2u * static_cast<uint32_t>(jp)) | (1u << 16));
                    if (p.kw & 1) {
                        a_row = a_plane + static_cast<uint32_t>(p.kw - 1);
                        for (int rr = 0; rr < p.kh; rr += 2, a_row += 2u * static_cast<uint32_t>(p.Wp))
                            issue(a_row | ((rr + 1 < p.kh ? static_cast<uint32_t>(p.Wp) : 1u) << 16));
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
            const int nf = tile / tiles_per_frame;
            const int r = tile - nf * tiles_per_frame;
            const int zg = r / p.windows;
            const int win = r - zg * p.windows;
            const int z0 = zg * p.zt;
            const int zt_eff = min(p.zt, p.Do - z0);
            const int u = win * 128 + quad * 32 + lane;
            const int prow = u / p.Wp;
            const int q = u - prow * p.Wp;
            const bool row_ok = u < plane_positions && q < p.Wo;
            mbar_wait(&tfull_bar[acc], acc_ph);
            tc_fence_after();
            for (int j = 0; j < zt_eff && !(p.dbg & 4); ++j) {
                const int64_t m = ((static_cast<int64_t>(nf) * p.Do + z0 + j) * p.Ho + prow) * p.Wo + q;
                const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                       static_cast<uint32_t>(acc * p.acc_cols + j * 2 * p.n_tile);
                for (int c = half; c < chunks; c += 2) {
                    uint32_t rv[16], rc[16];
                    __syncwarp();
                    tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(c * 16), rv);
                    tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(p.n_tile + c * 16), rc);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) rv[i] = __float_as_uint(__uint_as_float(rv[i]) + __uint_as_float(rc[i]));
                    const int n0 = c * 16;
                    if (n0 >= p.epi.c_store) continue;
                    epilogue_chunk<ACT1, ACT2, FMT>(p.epi, rv, n0, m, row_ok, s_epi[0], s_epi[1], s_epi[2]);
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
