// "Slab" Conv3D on tcgen05 for narrow layers (16..64 input channels, <= 128 output channels) on large
// volumes -- TIMED's second block (32 -> 64 channels at 11^3).
//
// The generic kernel gathers one 128-pixel im2col box per filter tap.  With 32 channels that is 54 boxes of
// 64-byte rows per 128 output pixels, and the layer sits on the TMA row rate: 3.7 ms of copies against a
// 1.2 ms MMA floor (profiles/r1_summary.md, role timing).  Here the activation tensor is stored as a
// "chunk-plane padded volume" (CPV):
//     plane (hi | lo)  x  chunk of 8 channels  x  position t  x  8 bf16 (16 bytes)
// where t runs over ONE linearisation of all frames with shared zero margins: t = lead + ((f*Dp + z)*Hp +
// p)*Wp + q, Dp = D + pad etc.; the zero pixel after a row is also the one before the next row, the zero row
// after a plane is the one before the next plane, and so on.  In that numbering every filter tap is a constant
// position offset, so for a run of 128*mt consecutive output positions ALL taps read from one contiguous run
// of positions per chunk: the CTA stages that "slab" once (2 * n_chunks one-dimensional bulk copies) and
// feeds the tensor core by moving the start address of a no-swizzle K-major UMMA descriptor (16 bytes per
// position; the second K half of a K=16 step is the next chunk plane: LBO = slab stride).  Activations cross
// L2->SM about (1 + halo/tile) times instead of 27 times; weights (too large to keep resident) stream per
// tap through a small ring.  Positions that fall into the margins are computed and dropped by the
// epilogue (utilisation (D*H*W)/((D+1)(H+1)(W+1)): 77 % at 11^3).
//
// MMA issue per K step and M tile as in thin_conv.cuh: A_hi x [W_hi | W_lo] (N = 2*n_tile, main and
// correction columns) then A_lo x W_hi into the correction columns.
#pragma once
#include "common.cuh"
#include "conv_umma.cuh"
#include "thin_conv.cuh"

namespace tb {

constexpr int kSlabWStages = 8;
constexpr int kSlabThreads = kConvThreads + 32;   // + a second MMA-issuing warp (warp 10)
constexpr int kSlabMaxTaps = 343;            // 7 x 7 x 7, the planner's limit

struct SlabConvParams {
    // ---- tiling over the padded linearisation of the input
    int64_t t_first;          // first position a tile may start at (= lead of the CPV tensor)
    int64_t t_count;          // n_frames * Dp*Hp*Wp
    int32_t n_tiles_total;    // ceil(t_count / (128*mt))
    int32_t mt;               // 128-position M tiles per CTA tile
    int32_t Dp, Hp, Wp;
    int32_t Do, Ho, Wo;
    // ---- input (CPV)
    const uint8_t* in_hi;
    int64_t lo_plane_off;     // bytes from the hi plane to the lo plane
    int64_t chunk_stride;     // bytes between chunk planes (= T*16)
    int32_t n_chunks;         // stored channels / 8 (even)
    int32_t neg_halo, pos_halo;
    int32_t slab_pix;         // 128*mt + neg_halo + pos_halo
    int32_t slab_stride;      // smem bytes between chunk spans
    // ---- filter
    int32_t kd, kh, kw, pd, ph, pw;   // extents and pad-before
    // ---- weights: [tap][chunk][2*n_tile rows: hi then lo][8] bf16, one bulk copy per tap
    const uint8_t* w_packed;
    uint32_t w_tap_bytes;
    int32_t n_tile;
    int32_t acc_cols;         // TMEM columns per M tile (2*n_tile rounded up to 32)
    int32_t acc_stages;
    int32_t w_stages;
    int32_t w_group;          // taps per ring stage (one commit per stage: a tcgen05.commit costs 100-400 tensor cycles)
    ConvKernelParams epi;     // epilogue fields (bias/scale/shift, activations, output pointers, ldc, c_store)
    int32_t dbg;
};

#if defined(__CUDACC__)

template <int ACT1, int ACT2, int FMT>
__global__ void __launch_bounds__(kSlabThreads, 1)
slab_conv_kernel(const __grid_constant__ SlabConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));

    __shared__ __align__(8) uint64_t slab_full[2];
    __shared__ __align__(8) uint64_t slab_empty[2];
    __shared__ __align__(8) uint64_t w_full[kSlabWStages];
    __shared__ __align__(8) uint64_t w_empty[kSlabWStages];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_epi[3][128];
    __shared__ int32_t s_tap_off[kSlabMaxTaps];      // position offset of every filter tap

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        // one issuing thread per M tile (p.mt of them): each commits once to the barriers that release operands / publish a tile
        for (int b = 0; b < 2; ++b) {
            mbar_init(&slab_full[b], 1);
            mbar_init(&slab_empty[b], p.mt);
            mbar_init(&tfull_bar[b], p.mt);
            mbar_init(&tempty_bar[b], kConvEpilogueWarps);
        }
        for (int s = 0; s < p.w_stages; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&w_empty[s], p.mt);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_512(&tmem_base_slot);
    for (int i = threadIdx.x; i < p.kd * p.kh * p.kw; i += blockDim.x) {
        const int a = i / (p.kh * p.kw), b = (i / p.kw) % p.kh, c = i % p.kw;
        s_tap_off[i] = (a - p.pd) * p.Hp * p.Wp + (b - p.ph) * p.Wp + (c - p.pw);
    }
    for (int i = threadIdx.x; i < p.n_tile; i += blockDim.x) {
        s_epi[0][i] = p.epi.bias[i];
        s_epi[1][i] = p.epi.scale[i];
        s_epi[2][i] = p.epi.shift[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    const uint32_t plane_region = static_cast<uint32_t>(p.n_chunks) * static_cast<uint32_t>(p.slab_stride);
    const uint32_t slab_bytes = 2u * plane_region;                       // hi spans, then lo spans
    uint8_t* w_ring = smem + 2u * slab_bytes;
    const uint32_t w_stage_bytes = (static_cast<uint32_t>(p.w_group) * p.w_tap_bytes + 127u) & ~127u;
    const int n_taps = p.kd * p.kh * p.kw;
    const int tile_pos = 128 * p.mt;

    if (warp == 0) {
        // =============================================================== bulk-copy producer
        const bool leader = elect_one();
        const uint32_t span_bytes = static_cast<uint32_t>(p.slab_pix) * 16u;
        uint32_t slab_uses[2] = {0, 0};
        auto load_slab = [&](int tile, int b) {
            mbar_wait(&slab_empty[b], (slab_uses[b] & 1u) ^ 1u);
            ++slab_uses[b];
            if (leader) {
                if (TB_DBG(p.dbg, 1)) {
                    mbar_arrive(&slab_full[b]);
                } else {
                    mbar_expect_tx(&slab_full[b], 2u * static_cast<uint32_t>(p.n_chunks) * span_bytes);
                    const int64_t t0 = p.t_first + static_cast<int64_t>(tile) * tile_pos - p.neg_halo;
                    const uint8_t* src = p.in_hi + t0 * 16;
                    uint8_t* dst = smem + static_cast<size_t>(b) * slab_bytes;
                    for (int c = 0; c < p.n_chunks; ++c) {
                        bulk_load_1d(dst + c * p.slab_stride, src + c * p.chunk_stride, span_bytes, &slab_full[b]);
                        bulk_load_1d(dst + plane_region + c * p.slab_stride, src + p.lo_plane_off + c * p.chunk_stride,
                                     span_bytes, &slab_full[b]);
                    }
                }
            }
            __syncwarp();
        };
        int sb = 0, ws = 0;
        uint32_t wph = 0;
        const int prefetch_tap = min(2, n_taps - 1);
        bool first = true;
        for (int tile = blockIdx.x; tile < p.n_tiles_total; tile += gridDim.x) {
            if (first) { load_slab(tile, sb); first = false; }
            for (int tap = 0; tap < n_taps; tap += p.w_group) {
                // the next tile's slab is requested early: its buffer was released when the previous tile finished
                if (tap <= prefetch_tap && prefetch_tap < tap + p.w_group && tile + static_cast<int>(gridDim.x) < p.n_tiles_total)
                    load_slab(tile + gridDim.x, sb ^ 1);
                mbar_wait(&w_empty[ws], wph ^ 1u);
                if (leader) {
                    if (TB_DBG(p.dbg, 1)) {
                        mbar_arrive(&w_full[ws]);
                    } else {
                        const uint32_t bytes = static_cast<uint32_t>(min(p.w_group, n_taps - tap)) * p.w_tap_bytes;
                        mbar_expect_tx(&w_full[ws], bytes);
                        for (uint32_t off = 0; off < bytes; off += 16384u)
                            bulk_load_1d(w_ring + static_cast<size_t>(ws) * w_stage_bytes + off,
                                         p.w_packed + static_cast<size_t>(tap) * p.w_tap_bytes + off,
                                         min(16384u, bytes - off), &w_full[ws]);
                    }
                }
                __syncwarp();
                if (++ws == p.w_stages) { ws = 0; wph ^= 1u; }
            }
            sb ^= 1;
        }
    } else if (warp == 1 || warp == 10) {
        // =============================================================== MMA issuers: warp 1 -> M tile 0, warp 10 -> M tile 1
        const bool leader = elect_one();
        const uint32_t idesc = umma_idesc_bf16_m128(static_cast<uint32_t>(p.n_tile));
        const uint32_t idesc2 = umma_idesc_bf16_m128(static_cast<uint32_t>(2 * p.n_tile));
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);            // no swizzle, SBO = 128 B
        const uint32_t stride16 = static_cast<uint32_t>(p.slab_stride) >> 4;
        const uint32_t plane16 = plane_region >> 4;
        const uint32_t w_lbo16 = static_cast<uint32_t>(2 * p.n_tile);   // (2*n_tile rows * 16 B) >> 4
        const uint32_t smem16 = (smem_u32(smem) & 0x3FFFFu) >> 4;
        const uint32_t w_ring16 = (smem_u32(w_ring) & 0x3FFFFu) >> 4;
        const int k_steps = p.n_chunks >> 1;
        int sb = 0, ws = 0, acc = 0;
        uint32_t sph[2] = {0, 0}, wph = 0, acc_ph = 0;
        // ONE lane runs the whole role, and the per-tap work between MMAs is a table load and a few adds: the tensor
        // pipe does not run ahead of an issuing thread by more than an MMA or two, so every cycle the issue loop spends
        // on address arithmetic, branches or warp reconvergence is a cycle the pipe idles.  Even a bare issue loop
        // leaves ~8 cycles between two MMAs of one thread; two threads issuing to DISJOINT accumulators (one M tile
        // each, so the accumulation order of every output stays fixed) close that gap: tools/mma_pattern_probe.cu
        // measures 252 -> 224 cycles per K step for this kernel's pattern, which is the shared-memory operand bound.
        const int q = warp == 1 ? 0 : 1;
        const uint32_t n_tile = static_cast<uint32_t>(p.n_tile), acc_cols = static_cast<uint32_t>(p.acc_cols);
        const uint32_t stage16 = w_stage_bytes >> 4;
        const int w_group = p.w_group, w_stages = p.w_stages, acc_stages = p.acc_stages;
        const bool skip = TB_DBG(p.dbg, 2);
        const uint32_t a_step = 2u * stride16, b_step = 2u * w_lbo16;
        if (leader && q < p.mt)
        for (int tile = blockIdx.x; tile < p.n_tiles_total; tile += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_ph ^ 1u);
            mbar_wait(&slab_full[sb], sph[sb]);
            sph[sb] ^= 1u;
            tc_fence_after();
            const uint32_t d_main = tmem_base + static_cast<uint32_t>(acc * p.mt + q) * acc_cols;
            const uint32_t d_corr = d_main + n_tile;
            // first row of this issuer's M tile for a tap with position offset 0
            const uint32_t a_tile = (smem16 + static_cast<uint32_t>(sb) * (slab_bytes >> 4) + static_cast<uint32_t>(p.neg_halo) +
                                     static_cast<uint32_t>(q) * 128u) | (stride16 << 16);
            uint32_t accumulate = 0;
            int in_group = 0;
            uint32_t b_k = 0;
            for (int tap = 0; tap < n_taps; ++tap) {
                if (in_group == 0) {
                    mbar_wait(&w_full[ws], wph);
                    tc_fence_after();
                    b_k = (w_ring16 + static_cast<uint32_t>(ws) * stage16) | (w_lbo16 << 16);
                }
                uint32_t a_k = a_tile + static_cast<uint32_t>(s_tap_off[tap]);
                if (!skip) {
                    for (int ks = 0; ks < k_steps; ++ks, a_k += a_step, b_k += b_step) {
                        umma_bf16_desc(true, d_main, a_k, desc_hi, b_k, desc_hi, idesc2, accumulate);
                        umma_bf16_desc(true, d_corr, a_k + plane16, desc_hi, b_k, desc_hi, idesc, 1u);
                        accumulate = 1u;
                    }
                } else {
                    b_k += static_cast<uint32_t>(k_steps) * b_step;
                }
                if (++in_group == w_group || tap + 1 == n_taps) {
                    in_group = 0;
                    umma_commit(&w_empty[ws]);
                    if (++ws == w_stages) { ws = 0; wph ^= 1u; }
                }
            }
            umma_commit(&slab_empty[sb]);
            umma_commit(&tfull_bar[acc]);
            sb ^= 1;
            if (++acc == acc_stages) { acc = 0; acc_ph ^= 1u; }
        }
        __syncwarp();
    } else {
        // =============================================================== epilogue (warps 2..9)
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int chunks = p.n_tile / 16;
        const int64_t fpos = static_cast<int64_t>(p.Dp) * p.Hp * p.Wp;
        const int hw = p.Hp * p.Wp;
        int acc = 0;
        uint32_t acc_ph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles_total; tile += gridDim.x) {
            mbar_wait(&tfull_bar[acc], acc_ph);
            tc_fence_after();
            for (int mi = 0; mi < p.mt && !TB_DBG(p.dbg, 4); ++mi) {
                const int64_t u = static_cast<int64_t>(tile) * tile_pos + mi * 128 + quad * 32 + lane;   // t - t_first
                const int64_t f = u / fpos;
                int rem = static_cast<int>(u - f * fpos);
                const int z = rem / hw;
                rem -= z * hw;
                const int pr = rem / p.Wp;
                const int q = rem - pr * p.Wp;
                const bool row_ok = u < p.t_count && z < p.Do && pr < p.Ho && q < p.Wo && !TB_DBG(p.dbg, 8);
                const int64_t m = ((f * p.Do + z) * p.Ho + pr) * p.Wo + q;
                const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                       static_cast<uint32_t>((acc * p.mt + mi) * p.acc_cols);
                for (int c = half; c < chunks; c += 2) {
                    uint32_t r[16], rc[16];
                    __syncwarp();
                    tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(c * 16), r);
                    tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(p.n_tile + c * 16), rc);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(fmaf(__uint_as_float(r[i]), p.epi.acc_comp, __uint_as_float(rc[i])));
                    const int n0 = c * 16;
                    if (n0 >= p.epi.c_store) continue;
                    epilogue_chunk<ACT1, ACT2, FMT>(p.epi, r, n0, m, row_ok, s_epi[0], s_epi[1], s_epi[2]);
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
