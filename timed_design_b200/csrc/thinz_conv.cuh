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
#include "kernels.cuh"
#include "thin_conv.cuh"

namespace tb {

// Epilogue warps: kThinzSubs per TMEM lane quadrant.  Measured: 16 warps are no faster than 8 for the unfused epilogue
// (profiles/r1_summary.md), nor are 12 for the fused max-pool one (round 2: 2.91 ms either way).  Role timing of the fused
// epilogue (bring-up build, TIMED_B200_DBG 2/18/34/50): barriers + loads 0.71 ms, phase 1 +0.67 ms (128 KB of TMEM per
// tile at ~128 B/cycle), phase 2 +0.69 ms (a short dependent chain per pooled pixel between two CTA-wide barriers);
// neither phase responds to fewer ALU instructions, more warps or wider TMEM loads.
constexpr int kThinzEpiWarps = 8;
constexpr int kThinzSubs = kThinzEpiWarps / 4;
constexpr int kThinzThreads = 64 + 32 * kThinzEpiWarps + 32;   // producer, MMA issuer, epilogue warps, second MMA issuer (the last warp)

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
    int32_t win_stride;       // in-plane positions between consecutive windows (128)
    // ---- fused MaxPool(2,2,2; stride 2) of the conv output (POOL = 1): see the epilogue.  Pooled pixels are written in the
    // consumer's layout (chunk-plane padded volume, or a plain NDHWC view).
    int32_t issuers;          // MMA-issuing threads (1 or 2; 2 needs two accumulator stages)
    int32_t pool_same;        // TF 'same' (partial windows at the far edge are kept) or 'valid'
    int32_t Zo, Po, Qo;       // pooled extents
    int32_t pool_cpv;         // 1: chunk-plane padded volume (out_hi4/out_lo4 + geometry below), 0: plain NDHWC view
    int64_t cpv_T, cpv_lead;
    int32_t cpv_Dp, cpv_Hp, cpv_Wp;
    uint4* out_hi4;
    uint4* out_lo4;
    TView pool_out;
};

#if defined(__CUDACC__)

// Tile walked by this CTA at iteration `it`: CTAs take whole (frame, z group) units round-robin and walk a unit's
// windows consecutively (tile = unit * windows + window), which the fused max-pool needs -- a pooled pixel's 2x2 in-plane
// partners may lie in the NEXT window -- and which keeps a unit's input planes hot in L2.  -1 when the CTA is done.
__device__ __forceinline__ int thinz_tile(int it, int windows, int n_tiles_total) {
    const int unit = static_cast<int>(blockIdx.x) + (it / windows) * static_cast<int>(gridDim.x);
    const int tile = unit * windows + it % windows;
    return tile < n_tiles_total ? tile : -1;
}

// FMT: format of the conv output (POOL = 0, 2) or of the plain pooled output (POOL = 1, ignored for CPV).
// POOL: 0 none, 1 whole MaxPool(2,2,2;2) in the epilogue, 2 its z direction only.
template <int ACT1, int ACT2, int FMT, int POOL>
__global__ void __launch_bounds__(kThinzThreads, 1)
thinz_conv_kernel(const __grid_constant__ ThinZParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));

    __shared__ __align__(8) uint64_t full_bar[kConvMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kConvMaxStages];
    __shared__ __align__(8) uint64_t tfull_bar[4];
    __shared__ __align__(8) uint64_t tempty_bar[4];
    __shared__ __align__(8) uint64_t w_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_epi[3][128];
    __shared__ uint32_t s_step_a[64];                // per K step: A descriptor low word relative to the plane's span
    __shared__ __align__(16) float s_sign[128];      // POOL: +1 where the epilogue map is increasing in the accumulator, -1 where decreasing

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 4; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], kThinzEpiWarps);
        }
        mbar_init(&w_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_512(&tmem_base_slot);
    if (threadIdx.x < 64) {
        // K steps of one input plane: (kh row, kw pair) -- the pair's pixels are adjacent, LBO = 1 -- then the odd kw tap
        // of two consecutive rows, LBO = Wp (or 1: an aliased, zero-weighted neighbour, for the last odd row)
        const int st = threadIdx.x, pairs = p.kw >> 1, n_pair_steps = p.kh * pairs;
        uint32_t v = 0;
        if (st < n_pair_steps) {
            v = (static_cast<uint32_t>((st / pairs) * p.Wp + 2 * (st % pairs))) | (1u << 16);
        } else {
            const int rr = 2 * (st - n_pair_steps);
            if (rr < p.kh)
                v = static_cast<uint32_t>(p.kw - 1 + rr * p.Wp) | ((rr + 1 < p.kh ? static_cast<uint32_t>(p.Wp) : 1u) << 16);
        }
        s_step_a[st] = v;
    }
    for (int i = threadIdx.x; i < p.n_tile; i += blockDim.x) {
        s_epi[0][i] = p.epi.bias[i];
        s_epi[1][i] = p.epi.scale[i];
        s_epi[2][i] = p.epi.shift[i];
        s_sign[i] = p.epi.scale[i] >= 0.f ? 1.f : -1.f;
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
    float* pool_stage = reinterpret_cast<float*>(stage0 + static_cast<size_t>(p.stages) * stage_bytes);   // POOL 1: zt/2 rings of [256][n_tile] fp32

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
        for (int it = 0;; ++it) {
            const int tile = thinz_tile(it, p.windows, p.n_tiles_total);
            if (tile < 0) break;
            const int nf = tile / tiles_per_frame;
            const int r = tile - nf * tiles_per_frame;
            const int zg = r / p.windows;
            const int win = r - zg * p.windows;
            const int z0 = zg * p.zt;
            const int n_planes = min(p.zt, p.Do - z0) + p.kd - 1;
            mbar_wait(&empty_bar[s], ph ^ 1u);
            if (leader && TB_DBG(p.dbg, 1)) {
                mbar_arrive(&full_bar[s]);
            } else if (leader) {
                mbar_expect_tx(&full_bar[s], 2u * static_cast<uint32_t>(n_planes) * p.span_bytes);
                uint8_t* st = stage0 + static_cast<size_t>(s) * stage_bytes;
                const uint8_t* src = p.in_hi + nf * p.frame_bytes + static_cast<int64_t>(z0 + p.off_d) * p.dplane_bytes +
                                     static_cast<int64_t>(p.off_hw + win * p.win_stride) * 16;
                for (int i = 0; i < n_planes; ++i, src += p.dplane_bytes) {
                    bulk_load_1d(st + i * p.span_stride, src, p.span_bytes, &full_bar[s]);
                    bulk_load_1d(st + plane_region + i * p.span_stride, src + p.lo_plane_off, p.span_bytes, &full_bar[s]);
                }
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1 || warp == 2 + kThinzEpiWarps) {
        // =============================================================== MMA issuers
        const bool leader = elect_one();
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);            // no swizzle, SBO = 128 B
        const uint32_t n2 = static_cast<uint32_t>(2 * p.n_tile);
        const uint32_t w_base16 = (smem_u32(w_smem) & 0x3FFFFu) >> 4;
        const uint32_t stage0_16 = (smem_u32(stage0) & 0x3FFFFu) >> 4;
        const uint32_t plane16 = plane_region >> 4;
        const uint32_t ss16 = static_cast<uint32_t>(p.span_stride) >> 4;
        const uint32_t b1_lbo = static_cast<uint32_t>(p.b1_rows), b2_lbo = static_cast<uint32_t>(p.b2_rows);   // rows*16 B >> 4
        const uint32_t step16 = 2u * (b1_lbo + b2_lbo);
        const int n_steps = p.kh * (p.kw >> 1) + ((p.kw & 1) ? (p.kh + 1) / 2 : 0);
        const uint32_t n_tile = static_cast<uint32_t>(p.n_tile);
        const int zt = p.zt, kd = p.kd, Do = p.Do, windows = p.windows, n_tiles_total = p.n_tiles_total;
        const int stages = p.stages, acc_stages = p.acc_stages;
        const uint32_t acc_cols = static_cast<uint32_t>(p.acc_cols), stage16 = stage_bytes >> 4;
        const bool skip = TB_DBG(p.dbg, 2);
        const uint32_t idesc_f = umma_idesc_bf16_m128(n2);
        const uint32_t bf = (w_base16 + static_cast<uint32_t>(kd - 1) * n2) | (b1_lbo << 16);
        // ONE lane runs the whole role (see slab_conv.cuh: issue-side work is exposed, the tensor pipe does not run ahead),
        // with per-plane descriptors hoisted and the per-step A offsets in a table.  With two accumulator stages there
        // are two issuers: warp 1 takes the even tiles of this CTA's sequence (stage 0), the last warp the odd ones (stage 1),
        // so the two streams never touch the same accumulator or input stage -- every output keeps a fixed accumulation
        // order -- and one fills the other's issue gaps (tools/mma_pattern_probe.cu: 193 -> 176 cycles per step pair).
        const int n_iss = (acc_stages >= 2 && (acc_stages & 1) == 0 && p.issuers == 2) ? 2 : 1;
        const int q = warp == 1 ? 0 : 1;
        int s = q, acc = q;                                   // stages >= 2
        uint32_t ph = 0, acc_ph = 0;
        if (leader && q < n_iss) {
        mbar_wait(&w_bar, 0);
        for (int it = q;; it += n_iss) {
            const int tile = thinz_tile(it, windows, n_tiles_total);
            if (tile < 0) break;
            const int r = tile % tiles_per_frame;
            const int z0 = (r / windows) * zt;
            const int zt_eff = min(zt, Do - z0);
            const int n_planes = zt_eff + kd - 1;
            mbar_wait(&tempty_bar[acc], acc_ph ^ 1u);
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (!skip) {
                const uint32_t d_tile = tmem_base + static_cast<uint32_t>(acc) * acc_cols;
                uint32_t a_plane = stage0_16 + static_cast<uint32_t>(s) * stage16;
                for (int i = 0; i < n_planes; ++i, a_plane += ss16) {
                    // output planes fed by input plane i: j_lo..j_hi through filter slices kd' = i - j
                    const int j_hi = min(zt_eff - 1, i);
                    const int j_lo = max(0, i - (kd - 1));
                    const uint32_t cnt = static_cast<uint32_t>(j_hi - j_lo + 1);
                    const uint32_t row0 = static_cast<uint32_t>(kd - 1 - (i - j_lo)) * n2;     // first B row (both blocks)
                    const uint32_t d_lo = d_tile + static_cast<uint32_t>(j_lo) * n2;
                    const uint32_t d_lo2 = d_lo + n_tile;
                    const uint32_t idesc1 = umma_idesc_bf16_m128(cnt * n2);
                    const uint32_t idesc2 = umma_idesc_bf16_m128(cnt * n2 - n_tile);
                    if constexpr (POOL == 1) {
                        // corrections into the main columns (see thinz_geometry): per step A_hi x W_hi, A_hi x W_lo, A_lo x W_hi,
                        // each N = cnt * n_tile on the same accumulator columns; blocks [W_hi][W_lo] of kd*n_tile rows each
                        const uint32_t rowm = static_cast<uint32_t>(kd - 1 - (i - j_lo)) * n_tile;
                        const uint32_t dm = d_tile + static_cast<uint32_t>(j_lo) * n_tile;
                        const uint32_t idm = umma_idesc_bf16_m128(cnt * n_tile);
                        uint32_t bh = (w_base16 + rowm) | (b1_lbo << 16);
                        uint32_t bl = (w_base16 + 2u * b1_lbo + rowm) | (b1_lbo << 16);
                        const uint32_t stepm16 = 4u * b1_lbo;
                        int st = 0;
                        if (i < zt_eff) {
                            // output plane i gets its first contribution (kd' = 0, the last W_hi block)
                            const uint32_t a_hi = a_plane + s_step_a[0];
                            const uint32_t bfm = (w_base16 + static_cast<uint32_t>(kd - 1) * n_tile) | (b1_lbo << 16);
                            umma_bf16_desc(true, d_tile + static_cast<uint32_t>(i) * n_tile, a_hi, desc_hi, bfm, desc_hi,
                                           umma_idesc_bf16_m128(n_tile), 0u);
                            if (cnt > 1)
                                umma_bf16_desc(true, dm, a_hi, desc_hi, bh, desc_hi, umma_idesc_bf16_m128((cnt - 1) * n_tile), 1u);
                            umma_bf16_desc(true, dm, a_hi, desc_hi, bl, desc_hi, idm, 1u);
                            umma_bf16_desc(true, dm, a_hi + plane16, desc_hi, bh, desc_hi, idm, 1u);
                            bh += stepm16;
                            bl += stepm16;
                            st = 1;
                        }
                        for (; st < n_steps; ++st, bh += stepm16, bl += stepm16) {
                            const uint32_t a_hi = a_plane + s_step_a[st];
                            umma_bf16_desc(true, dm, a_hi, desc_hi, bh, desc_hi, idm, 1u);
                            umma_bf16_desc(true, dm, a_hi, desc_hi, bl, desc_hi, idm, 1u);
                            umma_bf16_desc(true, dm, a_hi + plane16, desc_hi, bh, desc_hi, idm, 1u);
                        }
                        continue;
                    }
                    uint32_t b1 = (w_base16 + row0) | (b1_lbo << 16);
                    uint32_t b2 = (w_base16 + 2u * b1_lbo + row0) | (b2_lbo << 16);
                    int st = 0;
                    if (i < zt_eff) {
                        // output plane i gets its first contribution (kd' = 0): that block initialises the plane's
                        // (main | correction) columns
                        const uint32_t a_hi = a_plane + s_step_a[0];
                        umma_bf16_desc(true, d_tile + static_cast<uint32_t>(i) * n2, a_hi, desc_hi, bf, desc_hi, idesc_f, 0u);
                        if (cnt > 1)
                            umma_bf16_desc(true, d_lo, a_hi, desc_hi, b1, desc_hi, umma_idesc_bf16_m128((cnt - 1) * n2), 1u);
                        umma_bf16_desc(true, d_lo2, a_hi + plane16, desc_hi, b2, desc_hi, idesc2, 1u);
                        b1 += step16;
                        b2 += step16;
                        st = 1;
                    }
                    for (; st < n_steps; ++st, b1 += step16, b2 += step16) {
                        const uint32_t a_hi = a_plane + s_step_a[st];
                        umma_bf16_desc(true, d_lo, a_hi, desc_hi, b1, desc_hi, idesc1, 1u);
                        umma_bf16_desc(true, d_lo2, a_hi + plane16, desc_hi, b2, desc_hi, idesc2, 1u);
                    }
                }
            }
            umma_commit(&empty_bar[s]);
            umma_commit(&tfull_bar[acc]);
            s += n_iss;
            if (s >= stages) { s -= stages; ph ^= 1u; }
            acc += n_iss;                                      // this issuer's accumulator stages: q, q + n_iss, ...
            if (acc >= acc_stages) { acc -= acc_stages; acc_ph ^= 1u; }
        }
        }
        __syncwarp();
    } else {
        // =============================================================== epilogue (warps 2 .. 2 + kThinzEpiWarps - 1)
        const int quad = warp & 3;                    // TMEM lane quadrant this warp may read
        const int sub = (warp - 2) >> 2;              // which of the quadrant's kThinzSubs warps
        const int chunks = p.n_tile / 16;
        const int plane_positions = p.Ho * p.Wp;
        // POOL 1: the shared-memory pipe is the MMAs' (operand reads take ~86 % of its cycles), so the epilogue keeps its
        // per-channel constants in registers: a bit mask of the decreasing channels, and -- a thread's channel group in
        // phase 2 never changes -- that group's bias / scale / shift
        uint32_t neg_mask[4] = {0u, 0u, 0u, 0u};
        float e_bias[8], e_scale[8], e_shift[8];
        uint32_t e_flip[8];
        const int pool_g = (threadIdx.x - 64) % (p.n_tile >> 3);
        if constexpr (POOL == 1) {
#pragma unroll
            for (int w = 0; w < 4; ++w)
                for (int b = 0; b < 32 && w * 32 + b < p.n_tile; ++b)
                    if (s_sign[w * 32 + b] < 0.f) neg_mask[w] |= 1u << b;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                e_bias[e] = s_epi[0][pool_g * 8 + e];
                e_scale[e] = s_epi[1][pool_g * 8 + e];
                e_shift[e] = s_epi[2][pool_g * 8 + e];
                e_flip[e] = s_sign[pool_g * 8 + e] < 0.f ? 0x80000000u : 0u;
            }
        }
        auto neg_bits8 = [&](int unit) -> uint32_t {          // the 8 mask bits of channels unit*8 .. unit*8+7
            const int w = unit >> 2;
            const uint32_t word = w == 0 ? neg_mask[0] : w == 1 ? neg_mask[1] : w == 2 ? neg_mask[2] : neg_mask[3];
            return (word >> ((unit & 3) * 8)) & 0xFFu;
        };
        int acc = 0;
        uint32_t acc_ph = 0;
        for (int it = 0;; ++it) {
            const int tile = thinz_tile(it, p.windows, p.n_tiles_total);
            if (tile < 0) break;
            const int nf = tile / tiles_per_frame;
            const int r = tile - nf * tiles_per_frame;
            const int zg = r / p.windows;
            const int win = r - zg * p.windows;
            const int z0 = zg * p.zt;
            const int zt_eff = min(p.zt, p.Do - z0);
            const int u0 = win * p.win_stride;
            const int u = u0 + quad * 32 + lane;
            const int prow = u / p.Wp;
            const int q = u - prow * p.Wp;
            const bool row_ok = u < plane_positions && q < p.Wo && !TB_DBG(p.dbg, 8);   // dbg 8: compute, do not store
            mbar_wait(&tfull_bar[acc], acc_ph);
            tc_fence_after();
            if constexpr (POOL == 0) {
                for (int item = sub; item < zt_eff * chunks && !TB_DBG(p.dbg, 4); item += kThinzSubs) {
                    const int j = item / chunks;
                    const int64_t m = ((static_cast<int64_t>(nf) * p.Do + z0 + j) * p.Ho + prow) * p.Wo + q;
                    const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                           static_cast<uint32_t>(acc * p.acc_cols + j * 2 * p.n_tile);
                    {
                        const int c = item - j * chunks;
                        uint32_t rv[16], rc[16];
                        __syncwarp();
                        tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(c * 16), rv);
                        tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(p.n_tile + c * 16), rc);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) rv[i] = __float_as_uint(__uint_as_float(rv[i]) + __uint_as_float(rc[i]));
                        const int n0 = c * 16;
                        if (n0 < p.epi.c_store)
                            epilogue_chunk<ACT1, ACT2, FMT>(p.epi, rv, n0, m, row_ok, s_epi[0], s_epi[1], s_epi[2]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            } else if constexpr (POOL == 2) {
                // max over the z pairs only: halves what this kernel writes and what the (then 1x2x2) pooling pass reads;
                // no staging, no barrier, no overlapping windows.  bias -> activation -> BatchNorm -> activation is a
                // monotone map of the accumulator (increasing where the folded BN scale is >= 0, decreasing where it is
                // negative), so the max over the pair is taken on the RAW sums -- max, or min for a negative scale -- and
                // the epilogue math runs once per pooled value instead of once per plane (the epilogue is ALU-bound).
                const int n_zp = (zt_eff + 1) >> 1;
                for (int item = sub; item < n_zp * chunks && !TB_DBG(p.dbg, 4); item += kThinzSubs) {
                    const int zp = item / chunks;
                    const int c = item - zp * chunks;
                    const int Z = (z0 >> 1) + zp;
                    const bool two = 2 * zp + 1 < zt_eff;
                    const bool ok = row_ok && Z < p.Zo && (two || p.pool_same);
                    uint32_t raw[16];
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl) {
                        if (pl == 1 && !two) break;
                        const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                               static_cast<uint32_t>(acc * p.acc_cols + (2 * zp + pl) * 2 * p.n_tile);
                        uint32_t rv[16], rc[16];
                        __syncwarp();
                        tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(c * 16), rv);
                        tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(p.n_tile + c * 16), rc);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float x = __uint_as_float(rv[i]) + __uint_as_float(rc[i]);
                            const float prev = __uint_as_float(raw[i]);
                            // s_sign[ch] > 0: increasing map -> keep the larger raw sum; < 0: the smaller one
                            const float pick = s_sign[c * 16 + i] >= 0.f ? fmaxf(prev, x) : fminf(prev, x);
                            raw[i] = __float_as_uint(pl == 0 ? x : pick);
                        }
                    }
                    float v[16];
                    epilogue_math16<ACT1, ACT2>(p.epi, raw, c * 16, s_epi[0], s_epi[1], s_epi[2], v);
                    const int64_t m = ((static_cast<int64_t>(nf) * p.Zo + Z) * p.Ho + prow) * p.Wo + q;
                    if (c * 16 < p.epi.c_store) epilogue_store16<FMT>(p.epi, v, c * 16, m, ok);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            } else {
                // Whole MaxPool(2,2,2; stride 2) in the epilogue.  The map accumulator -> output (bias, activation, BatchNorm,
                // activation) is monotone per channel, so the 2x2x2 maximum is taken on RAW sums (max where the folded BN scale
                // is >= 0, min where it is negative) and the epilogue math runs once per POOLED value: 8x less ALU work than
                // activating every conv output (the unfused epilogue is ALU-bound).
                //   phase 1: z pairs straight from TMEM (planes 2zp, 2zp+1 of the tile), the raw extremum staged in shared
                //            memory: ring of two windows (256 positions) per z pair, [position][channel] fp32, 16-byte slots
                //            XOR-swizzled by the position;
                //   phase 2: in-plane 2x2 of the pooled pixels whose four partners are now staged -- anchors u = 2P*Wp + 2Q in
                //            [128w - Wp - 1, 128(w+1) - Wp - 1), the tail of the plane with its last window -- then epilogue
                //            math, split, store in the consumer's layout.  Windows of a unit are consecutive tiles of this CTA
                //            (thinz_tile), so the ring always holds the previous window.
                const int et = threadIdx.x - 64;                         // index among the epilogue threads
                const int n_zp = (zt_eff + 1) >> 1;
                const int row_f4 = p.n_tile >> 2;                        // float4 slots per staged position
                const int ring_f4 = 256 * row_f4;                        // one z pair's ring
                float4* ring = reinterpret_cast<float4*>(pool_stage);
                if (!TB_DBG(p.dbg, 4) && !TB_DBG(p.dbg, 32)) {         // bring-up build: 32 skips phase 1, 16 phase 2
                    // ---- phase 1
                    for (int item = sub; item < n_zp * (p.n_tile >> 3); item += kThinzSubs) {   // (z pair, 8 channels)
                        const int zp = item / (p.n_tile >> 3);
                        const int unit = item - zp * (p.n_tile >> 3);
                        const bool two = 2 * zp + 1 < zt_eff;
                        // decreasing channels are staged NEGATED (sign-bit flip, exact), so that every later reduction is a
                        // plain maximum; phase 2 flips them back
                        const uint32_t neg8 = neg_bits8(unit);
                        uint32_t xm[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) xm[i] = ((neg8 >> i) & 1u) << 31;
                        float v[8];
                        {
                            // both planes' loads in flight before the one wait (the phase is bound by TMEM round trips)
                            // (this instantiation accumulates the corrections in the main columns: plane j at column j * n_tile)
                            const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                                   static_cast<uint32_t>(acc * p.acc_cols + 2 * zp * p.n_tile + unit * 8);
                            uint32_t rv0[8], rv1[8];
                            __syncwarp();
                            tmem_ld_32x32b_x8(tbase, rv0);
                            if (two) tmem_ld_32x32b_x8(tbase + static_cast<uint32_t>(p.n_tile), rv1);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                // acc_comp > 0 undoes the accumulator's truncation shrink (conv_umma.cuh); a positive factor
                                // commutes with the maximum
                                const float x0 = __uint_as_float(rv0[i] ^ xm[i]) * p.epi.acc_comp;
                                v[i] = two ? fmaxf(x0, __uint_as_float(rv1[i] ^ xm[i]) * p.epi.acc_comp) : x0;
                            }
                        }
                        const int srow = (u0 + quad * 32 + lane) & 255;
                        float4* dst = ring + zp * ring_f4 + srow * row_f4;
                        const int sw = srow & 7 & (row_f4 - 1);
                        dst[(unit * 2) ^ sw] = make_float4(v[0], v[1], v[2], v[3]);
                        dst[(unit * 2 + 1) ^ sw] = make_float4(v[4], v[5], v[6], v[7]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);            // every accumulator of this stage has been read
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kThinzEpiWarps) : "memory");
                // ---- phase 2
                const bool last_win = win + 1 == p.windows;
                const int a_lo = max(0, u0 - p.Wp - 1);
                const int a_hi = last_win ? plane_positions : u0 + 128 - p.Wp - 1;
                // dense enumeration of the pooled pixels anchored in [a_lo, a_hi): even rows pa = 2P, columns Q < Qo; a thread
                // owns one (channel group, z pair) combination and strides over the pixels (no per-item division cascade, no
                // lanes idling on odd anchors)
                const int groups = p.n_tile >> 3;
                const int combos = groups * n_zp;
                const int P0 = (a_lo / p.Wp) >> 1;                       // pooled row of the range's first position
                const int P_hi = min(p.Po - 1, ((a_hi - 1) / p.Wp) >> 1);
                const int n_cand = a_hi > a_lo ? max(0, P_hi - P0 + 1) * p.Qo : 0;
                const int my_combo = et % combos;
                const int g = pool_g;                                    // == my_combo % groups
                const int zp = my_combo / groups;
                const int Z = (z0 >> 1) + zp;
                const bool two = 2 * zp + 1 < zt_eff;
                const bool z_ok = Z < p.Zo && (two || p.pool_same);
                const int k_step = (32 * kThinzEpiWarps) / combos;
                const int k0 = et / combos;
                const int dP = k_step / p.Qo, dQ = k_step - dP * p.Qo;      // pixel stride as (rows, columns): no division per pixel
                int P = P0 + k0 / p.Qo, Q = k0 % p.Qo;
                for (int k = k0; k < n_cand && z_ok && et < k_step * combos && !TB_DBG(p.dbg, 4) && !TB_DBG(p.dbg, 16);
                     k += k_step, P += dP + (Q + dQ >= p.Qo ? 1 : 0), Q = Q + dQ >= p.Qo ? Q + dQ - p.Qo : Q + dQ) {
                    const int pa = 2 * P, qa = 2 * Q;
                    const int ua = pa * p.Wp + qa;
                    if (ua < a_lo || ua >= a_hi || pa >= p.Ho || qa >= p.Wo) continue;
                    const bool has_q = qa + 1 < p.Wo, has_p = pa + 1 < p.Ho;
                    if (!p.pool_same && (!has_q || !has_p)) continue;
                    const float4* zr = ring + zp * ring_f4;
                    float m8[8];
                    auto take = [&](int uu, bool first) {
                        const int row = uu & 255;
                        const int sw = row & 7 & (row_f4 - 1);
                        const float4 a = zr[row * row_f4 + ((2 * g) ^ sw)];
                        const float4 b = zr[row * row_f4 + ((2 * g + 1) ^ sw)];
                        const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            m8[e] = first ? x[e] : fmaxf(m8[e], x[e]);
                    };
                    take(ua, true);
                    if (has_q) take(ua + 1, false);
                    if (has_p) take(ua + p.Wp, false);
                    if (has_q && has_p) take(ua + p.Wp + 1, false);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float x = __uint_as_float(__float_as_uint(m8[e]) ^ e_flip[e]) + e_bias[e];
                        x = act_ct<ACT1>(x, p.epi.act1, p.epi.alpha1);
                        x = fmaf(x, e_scale[e], e_shift[e]);
                        m8[e] = act_ct<ACT2>(x, p.epi.act2, p.epi.alpha2);
                    }
                    if (p.pool_cpv) {
                        const int64_t t = p.cpv_lead + ((static_cast<int64_t>(nf) * p.cpv_Dp + Z) * p.cpv_Hp + P) * p.cpv_Wp + Q;
                        cpv_store(p.out_hi4, p.out_lo4, g * p.cpv_T + t, m8);
                    } else {
                        const int64_t pix = ((static_cast<int64_t>(nf) * p.Zo + Z) * p.Po + P) * p.Qo + Q;
                        if (g * 8 < (FMT == FMT_SPLIT ? p.pool_out.c_pad : p.pool_out.c))
                            store8<FMT>(p.pool_out, pix * p.pool_out.ld + g * 8, m8);
                    }
                }
                // the next window's phase 1 overwrites the ring slots of the window before this one: every thread must be
                // done reading them
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kThinzEpiWarps) : "memory");
            }
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
