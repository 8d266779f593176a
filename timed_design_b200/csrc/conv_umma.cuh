// Conv3D / Dense as an implicit GEMM on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
//   out[m, n] = act2( scale[n] * act1( sum_k A[m, k] * Wt[n, k] + bias[n] ) + shift[n] )
//
//   m = flattened output pixel (frame, z, p, q)          M = n_frames * Do*Ho*Wo
//   k = (tap, input channel)                              K = kd*kh*kw * C_in_pad
//   n = output channel                                    N = C_out
//
// Operands are bf16 "split planes": every fp32 value v is stored as hi = bf16(v) and
// lo = bf16(v - hi); the product is accumulated in fp32 TMEM as  A_hi*W_hi + A_lo*W_hi +
// A_hi*W_lo  (three tcgen05.mma per K-step), which carries ~16 mantissa bits per operand --
// what the 1e-4 probability / bit-exact-argmax contract needs (DESIGN.md "Numerics";
// plain bf16 or fp16/tf32 single-pass misses it by 25-100x on the TIMED stand-in).
//
// A tiles (128 output pixels x kc channels of one filter tap) are gathered straight from the
// NDHWC activation tensor by TMA in IM2COL mode (zero-fill supplies the 'same' padding); W tiles
// by tiled TMA.  Both land in shared memory in the canonical K-major swizzled layout that the
// UMMA descriptors read.  Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM
// allocation), warps 2..9 = epilogue (TMEM -> registers -> bias/act/BN-affine -> HBM; two warps per
// TMEM lane quadrant, alternating 16-column chunks).  The epilogue is specialised at compile
// time on (act1, act2, output format): a runtime switch per element made it as long as the
// mainloop (profiles/r1_summary.md, r1a).
#pragma once
#include "common.cuh"

namespace tb {

constexpr int kConvMaxStages = 8;
constexpr int kConvEpilogueWarps = 8;                       // two per TMEM lane quadrant
constexpr int kConvThreads = 64 + 32 * kConvEpilogueWarps;  // + TMA producer warp + MMA warp
constexpr int kUmmaThreads = kConvThreads + 32;            // conv_umma_kernel: + a second MMA-issuing warp (warp 10)

enum OutFmt : int { FMT_F32 = 0, FMT_SPLIT = 1 };

struct ConvKernelParams {
    // ---- GEMM tiling
    int32_t m_total;     // valid output pixels (rows)
    int32_t n_ctile_m;   // CTA tiles along M (each covers mt*128 rows)
    int32_t n_tiles;     // CTA tiles along N
    int32_t n_tile;      // columns per CTA tile (multiple of 16, <= 256) == UMMA N
    int32_t mt;          // 128-row sub-tiles per CTA tile (1 or 2)
    int32_t acc_cols;    // TMEM columns reserved per accumulator (n_tile, or 2*n_tile when nfold, rounded up to 32)
    int32_t acc_stages;  // accumulator ring depth (1 or 2)
    // N-folded issue for n_tile <= 128: the W_hi and W_lo sub-tiles are adjacent in shared memory, so ONE
    // MMA of N = 2*n_tile computes A_hi*[W_hi | W_lo] into columns [0,n) (main) and [n,2n) (correction)
    // and a second MMA adds A_lo*W_hi into [n,2n): two MMAs per K-step instead of three (a thin-N MMA costs
    // ~the same whatever N is), and the corrections no longer truncate against the main accumulator.
    int32_t nfold;
    // 2-CTA cluster mode for n_tile = 256 (two such accumulators fill TMEM): the pair shares every W tile
    // -- each CTA loads half of it and TMA-multicasts it to both -- so a 128 x 256 tile per CTA costs
    // the same L2->SM traffic as 256 x 256 per CTA did, and the freed TMEM half holds a separate
    // accumulator for the correction MMAs (3x less accumulator truncation).
    // Requires mt = 1 and a (2,1,1) cluster launch.  corr_off = TMEM column offset of that accumulator.
    int32_t cluster2;
    int32_t corr_off;
    // ---- K loop
    int32_t kh, kw;      // filter extents (kd implied by n_taps)
    int32_t n_taps;      // kd*kh*kw
    int32_t cin_pad;     // padded input channels (K stride of one tap in the packed weights)
    int32_t cin_blocks;  // cin_pad / kc
    int32_t kc;          // channels per k-block: 16 / 32 / 64  (row = 32 / 64 / 128 bytes)
    int32_t kg;          // k-blocks per pipeline stage
    int32_t n_kblocks;   // n_taps * cin_blocks
    int32_t stages;      // smem ring depth
    // ---- geometry of the output volume and the im2col lower corner (= -pad_before)
    int32_t Do, Ho, Wo;
    int32_t lc_d, lc_h, lc_w;
    int32_t lo_plane_frames;  // frame offset of the lo plane inside the activation tensor map
    int32_t w_lo_rows;        // row offset of the lo plane inside the weight tensor map
    // ---- smem layout
    uint32_t a_sub_bytes;  // 128 * kc * 2
    uint32_t w_sub_bytes;  // n_tile * kc * 2
    uint32_t row_bytes;    // kc * 2
    uint32_t layout_type;  // UMMA swizzle code (2 / 4 / 6)
    // ---- epilogue
    const float* bias;     // [n_tiles*n_tile], zero padded (never NULL)
    const float* scale;    // idem (ones when absent)
    const float* shift;    // idem (zeros when absent)
    int32_t act1, act2;
    float alpha1, alpha2;
    int32_t out_fmt;       // FMT_F32 / FMT_SPLIT
    float* out_f32;        // [m_total][ldc] (FMT_F32)
    __nv_bfloat16* out_hi; // [m_total][ldc] (FMT_SPLIT)
    __nv_bfloat16* out_lo;
    int64_t ldc;           // row pitch of the output in elements
    int32_t c_store;       // columns < c_store are stored (FMT_SPLIT: multiple of 16)
    // Accumulator-truncation compensation: the tensor core adds every MMA's partial sums into the
    // fp32 TMEM accumulator with truncation toward zero, a relative shrink of ~1.2e-8 per accumulated MMA that is
    // COHERENT across the outputs of a layer (tools/accum_error.py).  The epilogue multiplies the main accumulator by
    // acc_comp = 1 + c(n)*n, n = MMAs with non-zero operands per output (host: accum_comp()).  1.0f switches it off.
    float acc_comp;
    // ---- voxel-stationary tiles (conv_pair_kernel only, stride 1): the 256 rows of a pair-tile are 256 consecutive FRAMES at
    // ONE output voxel, so a filter tap that falls into the zero 'same' padding is invalid for the whole tile and is
    // skipped -- at 6^3 with a 3^3 filter that is 30 % of the MMAs (and of the operand traffic).  Skipped taps would have
    // added exact zeros, and the remaining ones keep their order: results are bit-identical to the im2col tiling.
    // n_ctile_m = frame blocks x output voxels (tile -> m_ct -> (frame block, voxel)).
    int32_t vox;
    int32_t vox_frames;       // frames in this launch
    // Throttle for the static round-robin over tiles of UNEQUAL cost (8 .. 27 taps): no CTA pair starts tile t before
    // `*progress` >= t - window tiles have been completed (any pair's).  A pair that runs ahead would otherwise idle at the end
    // of the kernel; idling earlier costs nothing and keeps the tiles in flight within ~3 rounds, i.e. within the input planes
    // the L2 can hold.  The pair that owns the lowest unfinished tile is never held, so this cannot
    // deadlock.  nullptr: off.
    int32_t* progress;
    int32_t window;
    int32_t Di, Hi, Wi;       // input extents (validity of a tap: 0 <= out + lc + tap < in)
    // ---- col2im over kw in the epilogue (conv_umma_kernel, tap-to-N with the kw taps in N, kw = 3, 'same', stride 1):
    // a tile is c2i_rows = (128 / Wo) * Wo consecutive output pixels, i.e. WHOLE rows of the volume, so the kw shift never
    // leaves the tile: out[x] = Z[x-1][kw=0] + Z[x][kw=1] + Z[x+1][kw=2] is a lane shuffle of the accumulator chunks
    // (plus one shared-memory exchange at the three lane-quadrant seams), and the layer's real epilogue (bias / act / BN /
    // split) runs on the sum.  The (rows, kw*C_out) fp32 Z matrix never goes to HBM and the col2im kernel is not launched;
    // same operations in the same order, so the results are bit-identical to GEMM + col2im_kernel.  mt = 1, C_out = 16 / 32.
    int32_t c2i;
    int32_t c2i_rows;
    int32_t c2i_w;            // Wo
    int32_t c2i_cout;
    // ---- bring-up / tuning only (env TIMED_B200_DBG): 1 = skip TMA loads, 2 = skip MMA issue,
    // 4 = skip epilogue math+stores.  Results are garbage; used to time each role in isolation.
    int32_t dbg;
};

#if defined(__CUDACC__)

constexpr int kEpiSmemN = 512;   // conv_pair_kernel: epilogue vectors are staged in smem when n_alloc <= this
// conv_umma_kernel stages fewer: its static shared memory must stay under 2 KB so that FOUR 56 KB operand stages (DenseCPD's
// growth convs: 128 x 64 A hi/lo + 96 x 64 W hi/lo) fit beside it -- with three, the 112 KB in flight covered ~55 % of the
// L2 round trip and the tensor pipe idled the rest (profiles/r2c_densecpd_ncu_kernels.csv)
constexpr int kUmmaEpiSmemN = 128;

// ACT = -1 selects the runtime-dispatched activation (sigmoid/tanh/mixed cases).
// ELU is branch-free: a divergent expm1f call per element made the epilogue as long as the
// mainloop.  exp(x)-1 uses the SFU exp (abs. error ~1e-7) away from zero and a 4-term series
// within (-1/32, 0] (error < 3e-10) -- both far below the 2^-17 operand-split resolution.
template <int ACT>
__device__ __forceinline__ float act_ct(float x, int act_rt, float alpha) {
    if constexpr (ACT == ACT_NONE) return x;
    else if constexpr (ACT == ACT_RELU) return fmaxf(x, 0.0f);
    else if constexpr (ACT == ACT_ELU) {
        const float series = x * fmaf(x, fmaf(x, fmaf(x, 0.041666668f, 0.16666667f), 0.5f), 1.0f);
        const float em1 = x > -0.03125f ? series : __expf(x) - 1.0f;
        return x > 0.0f ? x : alpha * em1;
    } else return apply_act(x, act_rt, alpha);
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// D[tmem] (+)= A * B with descriptors given as (lo, shared hi) 32-bit halves; issued by the
// elected lane only (pred).  Keeping everything else warp-uniform lets ptxas do the address
// arithmetic on the uniform datapath instead of R2UR-ing per instruction.
__device__ __forceinline__ void umma_bf16_lohi(bool leader, uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo,
                                               uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
    if (leader) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "setp.ne.b32 p, %5, 0;\n\t"
            "mov.b64 da, {%1, %3};\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem),
            "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}

// 256-bit global store (STG.256, sm_100): `dst` must be 32-byte aligned
__device__ __forceinline__ void st_global_256(void* dst, const uint32_t (&r)[8]) {
    // evict-first in L2: a conv output is GBs per step and is not read again before it has been evicted anyway, while the
    // input planes and weights the other tiles still need should stay
    asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// One 16-column chunk of one accumulator row, in two steps: bias -> act1 -> folded BatchNorm -> act2 (`r` holds the raw
// fp32 accumulator bits of columns [n0, n0+16)), then the store (bf16 hi/lo split planes or fp32).
template <int ACT1, int ACT2>
__device__ __forceinline__ void epilogue_math16(const ConvKernelParams& p, const uint32_t (&r)[16], int n0,
                                                const float* bias_v, const float* scale_v, const float* shift_v,
                                                float (&v)[16]) {
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
        const float4 b = *(reinterpret_cast<const float4*>(bias_v + n0) + i4);
        const float4 sc = *(reinterpret_cast<const float4*>(scale_v + n0) + i4);
        const float4 sh = *(reinterpret_cast<const float4*>(shift_v + n0) + i4);
        const float bb[4] = {b.x, b.y, b.z, b.w};
        const float ss[4] = {sc.x, sc.y, sc.z, sc.w};
        const float hh[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float x = __uint_as_float(r[i4 * 4 + i]) + bb[i];
            x = act_ct<ACT1>(x, p.act1, p.alpha1);
            x = fmaf(x, ss[i], hh[i]);
            v[i4 * 4 + i] = act_ct<ACT2>(x, p.act2, p.alpha2);
        }
    }
}

template <int FMT>
__device__ __forceinline__ void epilogue_store16(const ConvKernelParams& p, const float (&v)[16], int n0, int64_t m,
                                                 bool row_ok) {
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
        // one 256-bit store per plane: a lane's 32 bytes are exactly one L2 sector (two 16-byte stores would each
        // write half a sector); rows are 32-byte aligned because ldc and n0 are multiples of 16 bf16
        st_global_256(p.out_hi + m * p.ldc + n0, hi);
        st_global_256(p.out_lo + m * p.ldc + n0, lo);
    } else if (row_ok) {
        float* dst = p.out_f32 + m * p.ldc + n0;
        if (n0 + 16 <= p.c_store && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i] = __float_as_uint(v[i]);
            st_global_256(dst, *reinterpret_cast<const uint32_t(*)[8]>(&w[0]));
            st_global_256(dst + 8, *reinterpret_cast<const uint32_t(*)[8]>(&w[8]));
        } else if (n0 + 16 <= p.c_store && (p.ldc & 3) == 0) {
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

template <int ACT1, int ACT2, int FMT>
__device__ __forceinline__ void epilogue_chunk(const ConvKernelParams& p, const uint32_t (&r)[16], int n0,
                                               int64_t m, bool row_ok, const float* bias_v,
                                               const float* scale_v, const float* shift_v) {
    float v[16];
    epilogue_math16<ACT1, ACT2>(p, r, n0, bias_v, scale_v, shift_v, v);
    epilogue_store16<FMT>(p, v, n0, m, row_ok);
}

// Early-release epilogue for tiles whose main and correction accumulators together fill TMEM (one accumulator stage,
// DESIGN.md section 4): the warp first drains its share of BOTH accumulators into registers (summing them), releases the
// TMEM stage -- the MMA warp starts the next tile's mainloop -- and only then runs bias / activation / BatchNorm / split /
// stores on the register copy.  The exposed part of the epilogue shrinks from the whole tile (~7 % of conv5) to the
// TMEM reads.  A warp owns the chunks half, half+2, ... of its lane quadrant: at most 8 for n_tile <= 256.
template <int ACT1, int ACT2, int FMT, typename Release>
__device__ __forceinline__ void epilogue_drained(const ConvKernelParams& p, uint32_t tbase, int half, int chunks,
                                                 int n_base, int64_t m, bool row_ok, const float* bias_v,
                                                 const float* scale_v, const float* shift_v, Release release) {
    uint32_t acc[8][16];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        const int c0 = half + 2 * i, c1 = c0 + 2;             // warp-uniform
        uint32_t rc0[16], rc1[16];
        __syncwarp();                                          // tcgen05.ld is .sync.aligned
        if (c0 < chunks) {
            tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(c0 * 16), acc[i]);
            tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(p.corr_off + c0 * 16), rc0);
        }
        if (c1 < chunks) {
            tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(c1 * 16), acc[i + 1]);
            tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(p.corr_off + c1 * 16), rc1);
        }
        tmem_ld_wait();
        if (c0 < chunks) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
                acc[i][e] = __float_as_uint(fmaf(__uint_as_float(acc[i][e]), p.acc_comp, __uint_as_float(rc0[e])));
        }
        if (c1 < chunks) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
                acc[i + 1][e] = __float_as_uint(fmaf(__uint_as_float(acc[i + 1][e]), p.acc_comp, __uint_as_float(rc1[e])));
        }
    }
    tc_fence_before();
    __syncwarp();
    release();                                                 // every tcgen05.ld of this warp has completed
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = half + 2 * i;
        if (c < chunks) {
            const int n0 = n_base + c * 16;
            if (n0 < p.c_store) epilogue_chunk<ACT1, ACT2, FMT>(p, acc[i], n0, m, row_ok, bias_v, scale_v, shift_v);
        }
    }
}

// Voxel-stationary tile (ConvKernelParams::vox): m_ct -> (frame block, output voxel) and the range of filter taps that
// fall inside the input
struct VoxTile {
    int fb, vox, z, y, x;
    int a0, a1, b0, b1, c0, c1;      // valid tap ranges along d, h, w (inclusive)
    int n_kb;                        // k-blocks of the tile: valid taps x cin_blocks
};
__device__ __forceinline__ VoxTile vox_decode(const ConvKernelParams& p, int m_ct) {
    VoxTile v;
    const int n_vox = p.Do * p.Ho * p.Wo;
    v.fb = m_ct / n_vox;
    v.vox = m_ct - v.fb * n_vox;
    v.x = v.vox % p.Wo;
    const int t = v.vox / p.Wo;
    v.y = t % p.Ho;
    v.z = t / p.Ho;
    const int kd = p.n_taps / (p.kh * p.kw);
    v.a0 = max(0, -(v.z + p.lc_d)); v.a1 = min(kd - 1, p.Di - 1 - (v.z + p.lc_d));
    v.b0 = max(0, -(v.y + p.lc_h)); v.b1 = min(p.kh - 1, p.Hi - 1 - (v.y + p.lc_h));
    v.c0 = max(0, -(v.x + p.lc_w)); v.c1 = min(p.kw - 1, p.Wi - 1 - (v.x + p.lc_w));
    v.n_kb = max(0, v.a1 - v.a0 + 1) * max(0, v.b1 - v.b0 + 1) * max(0, v.c1 - v.c0 + 1) * p.cin_blocks;
    return v;
}
// tiled (not im2col) 5-D box: {kc channels} x {128 frames} at one input voxel
// (L2 evict-last: a voxel's frames are re-read by up to 27 taps x the N tiles of neighbouring voxel tiles)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c, int32_t w,
                                            int32_t h, int32_t d, int32_t n, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n), "l"(policy)
        : "memory");
}

// One pipeline stage's MMAs for ONE 128-row sub-tile, issued by one thread as a straight run of UTCHMMAs with affine
// descriptor updates.  MODE 0: separate correction accumulator (d_corr); 1: N-folded (d_corr = d + n_tile);
// 3: corrections into the main accumulator.
template <int MODE>
__device__ __forceinline__ void umma_issue_stage(uint32_t base16, int nkb, uint32_t kb16, int k16_steps, uint32_t w_off16,
                                                 uint32_t w_sub16, uint32_t a_lo_off16, uint32_t d, uint32_t d_corr,
                                                 uint32_t desc_hi, uint32_t idesc, uint32_t idesc2, uint32_t accumulate) {
    for (int j = 0; j < nkb; ++j, base16 += kb16) {
        for (int kk = 0; kk < k16_steps; ++kk) {
            const uint32_t a0 = base16 + static_cast<uint32_t>(kk) * 2u;   // +32 B per K=16
            const uint32_t w_hi = a0 + w_off16;
            if constexpr (MODE == 0) {
                umma_bf16_lohi(true, d, a0, w_hi, desc_hi, idesc, accumulate);
                umma_bf16_lohi(true, d_corr, a0 + a_lo_off16, w_hi, desc_hi, idesc, accumulate);
                umma_bf16_lohi(true, d_corr, a0, w_hi + w_sub16, desc_hi, idesc, 1u);
            } else if constexpr (MODE == 1) {
                umma_bf16_lohi(true, d, a0, w_hi, desc_hi, idesc2, accumulate);
                umma_bf16_lohi(true, d_corr, a0 + a_lo_off16, w_hi, desc_hi, idesc, 1u);
            } else {
                umma_bf16_lohi(true, d, a0, w_hi, desc_hi, idesc, accumulate);
                umma_bf16_lohi(true, d, a0 + a_lo_off16, w_hi, desc_hi, idesc, 1u);
                umma_bf16_lohi(true, d, a0, w_hi + w_sub16, desc_hi, idesc, 1u);
            }
            accumulate = 1u;
        }
    }
}

template <int ACT1, int ACT2, int FMT>
__global__ void __launch_bounds__(kUmmaThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap map_a,
                 const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_v,
                 const ConvKernelParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

    __shared__ __align__(8) uint64_t full_bar[kConvMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kConvMaxStages];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_epi[3][kUmmaEpiSmemN];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_iss = (p.mt == 2 && !p.cluster2) ? 2 : 1;      // MMA-issuing threads: one per 128-row sub-tile

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            // cluster mode: both CTAs' MMA warps release a stage; two sub-tiles: one issuing thread each
            mbar_init(&empty_bar[s], p.cluster2 ? 2 : n_iss);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], n_iss);
            mbar_init(&tempty_bar[a], kConvEpilogueWarps);
        }
        mbar_fence_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_v);
    }
    const uint32_t cta_rank = p.cluster2 ? cluster_ctarank() : 0u;
    if (warp == 1) tmem_alloc_512(&tmem_base_slot);
    const int n_alloc = p.n_tiles * p.n_tile;
    const bool epi_in_smem = n_alloc <= kUmmaEpiSmemN;
    if (epi_in_smem) {
        for (int i = threadIdx.x; i < (p.c2i ? p.c2i_cout : n_alloc); i += blockDim.x) {
            s_epi[0][i] = p.bias[i];
            s_epi[1][i] = p.scale[i];
            s_epi[2][i] = p.shift[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.cluster2) cluster_sync_all();   // peer barriers are initialised before any remote arrive / multicast
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    // cluster mode: the pair walks the same sequence of (256-row pair-tile, n-tile); this CTA owns the
    // rank-th 128-row half.  tile_first / tile_step are in pair-tiles.
    const int total_tiles = p.n_ctile_m * p.n_tiles;
    const int tile_first = p.cluster2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int tile_step = p.cluster2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const uint32_t kb_bytes = static_cast<uint32_t>(p.mt) * 2u * p.a_sub_bytes + 2u * p.w_sub_bytes;
    const uint32_t stage_bytes = kb_bytes * static_cast<uint32_t>(p.kg);

    if (warp == 0) {
        // =============================================================== TMA producer
        // The whole warp walks the loops (uniform control flow); one elected lane issues.
        const bool leader = elect_one();
        const uint64_t pol_a = l2_policy_evict_last();
        int s = 0;
        uint32_t ph = 0;
        for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
            const int m_ct = tile / p.n_tiles;
            const int n_idx = tile - m_ct * p.n_tiles;
            int32_t bw[2], bh[2], bd[2], bn[2];
            int n_kb = p.n_kblocks;
            int a0 = 0, b0 = 0, b1 = p.kh - 1, c0 = 0, c1 = p.kw - 1;       // tap ranges the tile walks
            if (p.vox) {
                const VoxTile v = vox_decode(p, m_ct);
                for (int mi = 0; mi < 2; ++mi) {
                    bw[mi] = v.x + p.lc_w; bh[mi] = v.y + p.lc_h; bd[mi] = v.z + p.lc_d;
                    bn[mi] = (v.fb * p.mt + mi) * 128;
                }
                a0 = v.a0; b0 = v.b0; b1 = v.b1; c0 = v.c0; c1 = v.c1;
                n_kb = v.n_kb;
            } else {
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) {
                    int m0 = p.cluster2 ? (m_ct * 2 + static_cast<int>(cta_rank)) * 128
                           : p.c2i      ? m_ct * p.c2i_rows + mi * 128        // whole rows of the volume per tile
                                        : (m_ct * p.mt + mi) * 128;
                    if (m0 >= p.m_total) m0 = 0;       // dummy sub-tile: rows are discarded later
                    const int q = m0 % p.Wo;
                    int t = m0 / p.Wo;
                    const int pp = t % p.Ho;
                    t /= p.Ho;
                    const int z = t % p.Do;
                    const int nf = t / p.Do;
                    bw[mi] = q + p.lc_w;
                    bh[mi] = pp + p.lc_h;
                    bd[mi] = z + p.lc_d;
                    bn[mi] = nf;
                }
            }
            const int n_groups = (n_kb + p.kg - 1) / p.kg;
            if (p.progress) {                      // do not run more than `window` tiles ahead of the completed count
                if (leader) {
                    uint32_t spins = 0;
                    while (ld_acquire_gpu(p.progress) < tile - p.window) {
                        __nanosleep(256);
                        if (++spins > (1u << 24)) { printf("timed_b200: tile throttle timed out (block %d)\n", blockIdx.x); __trap(); }
                    }
                }
                __syncwarp();
            }
            int ta = a0, tb = b0, tc = c0, cb = 0;                 // current tap (d, h, w) and channel block
            for (int g = 0; g < n_groups; ++g) {
                mbar_wait(&empty_bar[s], ph ^ 1u);
                const int nkb = min(p.kg, n_kb - g * p.kg);
                if (leader) {
                    if (TB_DBG(p.dbg, 1)) {
                        mbar_arrive(&full_bar[s]);
                    } else {
                        mbar_expect_tx(&full_bar[s], static_cast<uint32_t>(nkb) * kb_bytes);
                        uint8_t* st = smem + static_cast<size_t>(s) * stage_bytes;
                        for (int j = 0; j < nkb; ++j) {
                            const int tap = (ta * p.kh + tb) * p.kw + tc;
                            uint8_t* base = st + static_cast<size_t>(j) * kb_bytes;
                            for (int mi = 0; mi < p.mt; ++mi) {
                                if (p.vox) {
                                    tma_load_5d(base + mi * p.a_sub_bytes, &map_v, &full_bar[s], cb * p.kc, bw[mi] + tc,
                                                bh[mi] + tb, bd[mi] + ta, bn[mi], pol_a);
                                    tma_load_5d(base + (p.mt + mi) * p.a_sub_bytes, &map_v, &full_bar[s], cb * p.kc, bw[mi] + tc,
                                                bh[mi] + tb, bd[mi] + ta, bn[mi] + p.lo_plane_frames, pol_a);
                                } else {
                                    tma_load_im2col_5d(base + mi * p.a_sub_bytes, &map_a, &full_bar[s],
                                                       cb * p.kc, bw[mi], bh[mi], bd[mi], bn[mi],
                                                       static_cast<uint16_t>(tc), static_cast<uint16_t>(tb),
                                                       static_cast<uint16_t>(ta));
                                    tma_load_im2col_5d(base + (p.mt + mi) * p.a_sub_bytes, &map_a, &full_bar[s],
                                                       cb * p.kc, bw[mi], bh[mi], bd[mi],
                                                       bn[mi] + p.lo_plane_frames, static_cast<uint16_t>(tc),
                                                       static_cast<uint16_t>(tb), static_cast<uint16_t>(ta));
                                }
                            }
                            uint8_t* wb = base + 2 * p.mt * p.a_sub_bytes;
                            const int kcoord = tap * p.cin_pad + cb * p.kc;
                            if (p.cluster2) {
                                // this CTA fetches its half of the W rows and multicasts it to the pair
                                const int half_rows = p.n_tile >> 1;
                                const uint32_t half_bytes = p.w_sub_bytes >> 1;
                                const int r0 = n_idx * p.n_tile + static_cast<int>(cta_rank) * half_rows;
                                tma_load_2d_mcast(wb + cta_rank * half_bytes, &map_w, &full_bar[s], kcoord, r0, 3);
                                tma_load_2d_mcast(wb + p.w_sub_bytes + cta_rank * half_bytes, &map_w, &full_bar[s],
                                                  kcoord, p.w_lo_rows + r0, 3);
                            } else {
                                tma_load_2d(wb, &map_w, &full_bar[s], kcoord, n_idx * p.n_tile);
                                tma_load_2d(wb + p.w_sub_bytes, &map_w, &full_bar[s], kcoord,
                                            p.w_lo_rows + n_idx * p.n_tile);
                            }
                            if (++cb == p.cin_blocks) {
                                cb = 0;
                                if (++tc > c1) { tc = c0; if (++tb > b1) { tb = b0; ++ta; } }
                            }
                        }
                    }
                }
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1 || warp == 10) {
        // =============================================================== MMA issuers (warp 1: sub-tile 0, warp 10: sub-tile 1)
        // ONE lane of each runs the whole role.  The tensor pipe does not run ahead of an issuing thread by more than
        // an MMA or two, so issue-side work (address arithmetic, branches, warp reconvergence) is exposed, and even a
        // bare loop leaves ~8 cycles between two MMAs of one thread; with two sub-tiles a second thread issues the
        // second one's MMAs (disjoint accumulators: every output's accumulation order stays fixed), which closes the
        // gap (tools/mma_pattern_probe.cu, profiles/r2i_mma_pattern_probe.txt).
        const bool leader = elect_one();
        const int q = warp == 1 ? 0 : 1;
        const uint32_t idesc = umma_idesc_bf16_m128(static_cast<uint32_t>(p.n_tile));
        const uint32_t idesc2 = umma_idesc_bf16_m128(static_cast<uint32_t>(2 * p.n_tile));
        // descriptor = {lo: start address >> 4 | LBO(=1) << 16, hi: SBO >> 4 | version << 14 | swizzle << 29}
        const uint32_t desc_hi = ((p.row_bytes * 8u) >> 4) | (1u << 14) | (p.layout_type << 29);
        const uint32_t lo_flags = 1u << 16;
        const int k16_steps = p.kc / 16;
        const uint32_t a_sub16 = p.a_sub_bytes >> 4, w_sub16 = p.w_sub_bytes >> 4, kb16 = kb_bytes >> 4;
        const uint32_t a_lo_off16 = static_cast<uint32_t>(p.mt) * a_sub16;       // hi plane -> lo plane
        const uint32_t w_off16 = 2u * static_cast<uint32_t>(p.mt) * a_sub16;     // A block -> W block
        const uint32_t smem_base16 = (smem_u32(smem) & 0x3FFFFu) >> 4;
        const int mode = (p.corr_off && !p.nfold) ? 0 : p.nfold ? 1 : 3;
        const uint32_t corr_off = mode == 0 ? static_cast<uint32_t>(p.corr_off) : static_cast<uint32_t>(p.n_tile);
        const bool skip = TB_DBG(p.dbg, 2);
        const int kg = p.kg, n_kblocks = p.n_kblocks, stages = p.stages, acc_stages = p.acc_stages, mt = p.mt;
        const bool mcast = p.cluster2 != 0;
        int s = 0;
        uint32_t ph = 0;
        int acc = 0;
        uint32_t acc_ph = 0;
        if (leader && q < n_iss)
        for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
            const int n_kb = p.vox ? vox_decode(p, tile / p.n_tiles).n_kb : n_kblocks;
            const int n_groups = (n_kb + kg - 1) / kg;
            mbar_wait(&tempty_bar[acc], acc_ph ^ 1u);
            tc_fence_after();
            uint32_t accumulate = 0;
            for (int g = 0; g < n_groups; ++g) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const int nkb = min(kg, n_kb - g * kg);
                const uint32_t stage16 = (smem_base16 + static_cast<uint32_t>(s) * (stage_bytes >> 4)) | lo_flags;
                // this thread's sub-tiles: its own when each has an issuer, otherwise all of them
                for (int sub = q; sub < mt && !skip; sub += n_iss) {
                    const uint32_t base16 = stage16 + static_cast<uint32_t>(sub) * a_sub16;
                    const uint32_t w_rel = w_off16 - static_cast<uint32_t>(sub) * a_sub16;
                    const uint32_t a_lo_rel = a_lo_off16;
                    const uint32_t d = tmem_base + static_cast<uint32_t>(acc * mt + sub) * static_cast<uint32_t>(p.acc_cols);
                    if (mode == 0)
                        umma_issue_stage<0>(base16, nkb, kb16, k16_steps, w_rel, w_sub16, a_lo_rel, d, d + corr_off, desc_hi, idesc, idesc2, accumulate);
                    else if (mode == 1)
                        umma_issue_stage<1>(base16, nkb, kb16, k16_steps, w_rel, w_sub16, a_lo_rel, d, d + corr_off, desc_hi, idesc, idesc2, accumulate);
                    else
                        umma_issue_stage<3>(base16, nkb, kb16, k16_steps, w_rel, w_sub16, a_lo_rel, d, d + corr_off, desc_hi, idesc, idesc2, accumulate);
                }
                accumulate = 1u;
                // frees the smem stage when the MMAs retire
                if (mcast) umma_commit_mcast(&empty_bar[s], 3);
                else umma_commit(&empty_bar[s]);
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
            umma_commit(&tfull_bar[acc]);     // accumulator complete -> epilogue
            if (++acc == acc_stages) { acc = 0; acc_ph ^= 1u; }
        }
        __syncwarp();
    } else {
        // =============================================================== epilogue (warps 2..9)
        const int quad = warp & 3;                    // TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;             // which of the quadrant's two warps
        const int row_in_tile = quad * 32 + lane;
        const int chunks = p.n_tile / 16;
        const float* bias_v = epi_in_smem ? s_epi[0] : p.bias;
        const float* scale_v = epi_in_smem ? s_epi[1] : p.scale;
        const float* shift_v = epi_in_smem ? s_epi[2] : p.shift;
        int acc = 0;
        uint32_t acc_ph = 0;
        for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
            const int m_ct = tile / p.n_tiles;
            const int n_idx = tile - m_ct * p.n_tiles;
            mbar_wait(&tfull_bar[acc], acc_ph);
            tc_fence_after();
            if (p.c2i) {
                // ---- col2im over kw on the accumulator (ConvKernelParams::c2i).  This warp: rows quad*32 .. +31, channels
                // [16*half, 16*half + 16) of each of the three kw column groups.
                // seam exchange buffers [seam between quadrants s and s+1][column half][channel], aliased onto the tail of the
                // staged epilogue vectors (C_out <= 32 of their 128 entries are in use; the static shared memory has no room
                // left beside the operand ring)
                float (*s_left)[2][16] = reinterpret_cast<float (*)[2][16]>(&s_epi[0][32]);     // last row of quadrant s, kw = 0
                float (*s_right)[2][16] = reinterpret_cast<float (*)[2][16]>(&s_epi[1][32]);    // first row of quadrant s+1, kw = 2
                const bool active = half * 16 < p.c2i_cout;                // warp-uniform
                const int64_t m = static_cast<int64_t>(m_ct) * p.c2i_rows + row_in_tile;
                const bool row_ok = row_in_tile < p.c2i_rows && m < p.m_total && !TB_DBG(p.dbg, 8);
                const int x = row_in_tile % p.c2i_w;                       // tiles start at x = 0
                const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                       static_cast<uint32_t>(acc * p.acc_cols);
                const int corr = p.nfold ? p.n_tile : p.corr_off;
                uint32_t z[3][16];
                if (active) {
                    __syncwarp();
#pragma unroll
                    for (int t = 0; t < 3; ++t)
                        tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(t * p.c2i_cout + half * 16), z[t]);
                    if (corr) {
                        uint32_t zc[3][16];
#pragma unroll
                        for (int t = 0; t < 3; ++t)
                            tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(corr + t * p.c2i_cout + half * 16), zc[t]);
                        tmem_ld_wait();
#pragma unroll
                        for (int t = 0; t < 3; ++t)
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                z[t][i] = __float_as_uint(fmaf(__uint_as_float(z[t][i]), p.acc_comp, __uint_as_float(zc[t][i])));
                    } else {
                        tmem_ld_wait();
#pragma unroll
                        for (int t = 0; t < 3; ++t)
#pragma unroll
                            for (int i = 0; i < 16; ++i) z[t][i] = __float_as_uint(__uint_as_float(z[t][i]) * p.acc_comp);
                    }
                }
                // the accumulator stage is free again: the next tile's mainloop overlaps the exchange, math and stores
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                if (active) {
                    if (lane == 31 && quad < 3) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) s_left[quad][half][i] = __uint_as_float(z[0][i]);
                    }
                    if (lane == 0 && quad > 0) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) s_right[quad - 1][half][i] = __uint_as_float(z[2][i]);
                    }
                }
                // all eight epilogue warps: seams written -> read ...
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kConvEpilogueWarps) : "memory");
                uint32_t r[16];
                if (active) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float left = __shfl_up_sync(0xffffffffu, __uint_as_float(z[0][i]), 1);
                        float right = __shfl_down_sync(0xffffffffu, __uint_as_float(z[2][i]), 1);
                        if (lane == 0 && quad > 0) left = s_left[quad - 1][half][i];
                        if (lane == 31 && quad < 3) right = s_right[quad][half][i];
                        float a = 0.0f;                                    // col2im_kernel's order: kw = 0, 1, 2
                        if (x > 0) a += left;
                        a += __uint_as_float(z[1][i]);
                        if (x < p.c2i_w - 1) a += right;
                        r[i] = __float_as_uint(a);
                    }
                }
                // ... and read -> the next tile's writes
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kConvEpilogueWarps) : "memory");
                if (active && !TB_DBG(p.dbg, 4))
                    epilogue_chunk<ACT1, ACT2, FMT>(p, r, half * 16, m, row_ok, bias_v, scale_v, shift_v);
                if (++acc == p.acc_stages) { acc = 0; acc_ph ^= 1u; }
                continue;
            }
            for (int mi = 0; mi < p.mt && !TB_DBG(p.dbg, 4); ++mi) {
                int64_t m = (p.cluster2 ? static_cast<int64_t>(m_ct * 2 + static_cast<int>(cta_rank))
                                        : static_cast<int64_t>(m_ct * p.mt + mi)) * 128 + row_in_tile;
                bool row_ok = m < p.m_total && !TB_DBG(p.dbg, 8);
                if (p.vox) {                           // row = frame, at the tile's voxel
                    const int n_vox = p.Do * p.Ho * p.Wo;
                    const int fb = m_ct / n_vox;
                    const int frame = (fb * p.mt + mi) * 128 + row_in_tile;
                    m = static_cast<int64_t>(frame) * n_vox + (m_ct - fb * n_vox);
                    row_ok = frame < p.vox_frames && !TB_DBG(p.dbg, 8);
                }
                const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                       static_cast<uint32_t>((acc * p.mt + mi) * p.acc_cols);
                for (int c = half; c < chunks; c += 2) {
                    uint32_t r[16];
                    __syncwarp();                      // tcgen05.ld is .sync.aligned
                    tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(c * 16), r);
                    if (p.nfold || p.corr_off) {       // warp-uniform: add the correction accumulator
                        uint32_t rc[16];
                        tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>((p.nfold ? p.n_tile : p.corr_off) + c * 16), rc);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            r[i] = __float_as_uint(fmaf(__uint_as_float(r[i]), p.acc_comp, __uint_as_float(rc[i])));
                    } else {
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * p.acc_comp);
                    }
                    const int n0 = n_idx * p.n_tile + c * 16;
                    if (n0 >= p.c_store) continue;     // warp-uniform
                    epilogue_chunk<ACT1, ACT2, FMT>(p, r, n0, m, row_ok, bias_v, scale_v, shift_v);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (p.progress && warp == 2 && lane == 0) atomicAdd(p.progress, 1);
            if (++acc == p.acc_stages) { acc = 0; acc_ph ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (p.cluster2) cluster_sync_all();   // the peer may still multicast into / arrive on this CTA's smem
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_512(tmem_base);
    }
}

#endif  // __CUDACC__

}  // namespace tb
