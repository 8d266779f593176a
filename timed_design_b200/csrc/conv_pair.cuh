// Conv3D implicit GEMM on a CTA PAIR: tcgen05.mma.cta_group::2 (UMMA M = 256 across two SMs of a TPC).
//
// Same contraction, operand layout (bf16 hi/lo split planes, TMA im2col A tiles, K-major swizzled
// W tiles) and epilogue as conv_umma_kernel, for the wide layers (n_tile = 256) whose 256 x 256 CTA
// tile filled TMEM in the single-CTA kernel.  There the epilogue could not overlap the next tile's
// mainloop, and every 128 x 256 x 16 MMA re-read 12 KB of shared memory per 128 tensor cycles while TMA
// wrote another 42 B/cycle into it -- above the 128 B/cycle the SM's shared memory delivers
// (profiles/r1_summary.md, role timing).  With the pair:
//   * each CTA owns 128 rows of a 256-row pair-tile and HALF of the W tile (n_tile/2 rows); the tensor
//     cores of both SMs read the two W halves through the pair's shared-memory window, so per CTA an MMA
//     reads 4 KB (A) + 4 KB (W half) instead of 4 + 8 KB, and L2->SM weight traffic per MAC is what the
//     256 x 256 tile paid;
//   * the accumulator of a CTA is 128 lanes x n_tile columns (<= 256), so TMEM holds TWO stages and
//     the epilogue of tile i runs under the mainloop of tile i+1.
// Protocol (all barriers live at identical offsets in both CTAs):
//   full[s]   leader's copy only: 1 arrival (leader's expect_tx for BOTH CTAs' bytes); every TMA of
//             either CTA signals it (.cta_group::2 loads, peer bit of the barrier address cleared);
//   empty[s]  own copy: tcgen05.commit.cta_group::2 multicast to both CTAs when the stage's MMAs retire;
//   tfull[a]  own copy: commit multicast when a tile's accumulator is complete;
//   tempty[a] leader's copy only: 8 epilogue warps x 2 CTAs arrive (the peer's remotely).
// Only the leader CTA (cluster rank 0) issues MMAs and commits.
#pragma once
#include "common.cuh"
#include "conv_umma.cuh"

namespace tb {

#if defined(__CUDACC__)

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the even CTA

__device__ __forceinline__ void tmem_alloc_512_pair(uint32_t* smem_dst) {   // whole warp, both CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(smem_dst))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512_pair(uint32_t taddr) {     // whole warp, both CTAs
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
// TMA loads executed by either CTA of the pair; the transaction bytes are credited to the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                                 int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_im2col_5d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                                        int32_t c, int32_t w, int32_t h, int32_t d, int32_t n,
                                                        uint16_t ow, uint16_t oh, uint16_t od) {
    asm volatile(
        "cp.async.bulk.tensor.5d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], {%8, %9, %10};" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c), "r"(w), "r"(h),
        "r"(d), "r"(n), "h"(ow), "h"(oh), "h"(od)
        : "memory");
}
// tiled (not im2col) 5-D box: voxel-stationary tiles fetch {kc channels} x {128 frames} at one input voxel
__device__ __forceinline__ void tma_load_5d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                                 int32_t c, int32_t w, int32_t h, int32_t d, int32_t n, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c), "r"(w), "r"(h),
        "r"(d), "r"(n), "l"(policy)
        : "memory");
}
// D[tmem, both CTAs] (+)= A * B, M = 256 over the pair; descriptors as (lo, shared hi) halves
__device__ __forceinline__ void umma_bf16_pair(bool leader, uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo,
                                               uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
    if (leader) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "setp.ne.b32 p, %5, 0;\n\t"
            "mov.b64 da, {%1, %3};\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem),
            "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {   // arrives in both CTAs of the pair
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}
// arrive on the barrier at this offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// Instruction descriptor: D=f32, A=B=bf16, both K-major, M=256 (cta_group::2), N=n (multiple of 16, <=256).
__device__ __host__ __forceinline__ uint32_t umma_idesc_bf16_m256(uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}

// Warp roles: warpgroup 0 = {TMA producer, MMA issuer, two idle warps}, warpgroups 1-2 = eight epilogue warps.  With the
// correction accumulator next to the main one a tile fills TMEM, so an epilogue warp drains its 8 chunks (128 registers)
// before releasing the stage (epilogue_drained): the epilogue warpgroups take 224 registers per thread and warpgroup 0
// gives back all but 56 (setmaxnreg; 128*56 + 256*224 = the 384*168 the CTA is launched with).
constexpr int kPairEpilogueWarp0 = 4;
constexpr int kPairThreads = 32 * (kPairEpilogueWarp0 + kConvEpilogueWarps);

// Launch: cluster (2,1,1), grid = 2 * min(pair_tiles, 74), kPairThreads threads.
// Uses ConvKernelParams with mt = 1, nfold = 0; n_ctile_m counts 256-row pair-tiles; w_sub_bytes is the
// per-CTA HALF tile ((n_tile/2) * kc * 2); map_w's box has n_tile/2 rows.
template <int ACT1, int ACT2, int FMT>
__global__ void __launch_bounds__(kPairThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap map_a,
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
    __shared__ __align__(16) float s_epi[3][kEpiSmemN];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader_cta = cta_rank == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 2 * kConvEpilogueWarps);
        }
        mbar_fence_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_v);
    }
    if (warp == 1) tmem_alloc_512_pair(&tmem_base_slot);
    const int n_alloc = p.n_tiles * p.n_tile;
    const bool epi_in_smem = n_alloc <= kEpiSmemN;
    if (epi_in_smem) {
        for (int i = threadIdx.x; i < n_alloc; i += blockDim.x) {
            s_epi[0][i] = p.bias[i];
            s_epi[1][i] = p.scale[i];
            s_epi[2][i] = p.shift[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // both CTAs' barriers are initialised before any remote arrive / peer TMA signal
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    const int total_tiles = p.n_ctile_m * p.n_tiles;            // pair-tiles
    const int tile_first = static_cast<int>(blockIdx.x >> 1);
    const int tile_step = static_cast<int>(gridDim.x >> 1);
    const uint32_t kb_bytes = 2u * p.a_sub_bytes + 2u * p.w_sub_bytes;   // per CTA
    const uint32_t stage_bytes = kb_bytes * static_cast<uint32_t>(p.kg);
    const int half_rows = p.n_tile >> 1;

    if (warp < kPairEpilogueWarp0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // =============================================================== TMA producer (both CTAs)
        const bool leader = elect_one();
        const uint64_t pol_a = l2_policy_evict_last();
        int s = 0;
        uint32_t ph = 0;
        for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
            const int m_ct = tile / p.n_tiles;
            const int n_idx = tile - m_ct * p.n_tiles;
            // base coordinates of the tile's A boxes, the tap ranges it walks and its k-block count
            int bw, bh, bd, nf, n_kb = p.n_kblocks;
            int a0 = 0, b0 = 0, b1 = p.kh - 1, c0 = 0, c1 = p.kw - 1;
            if (p.vox) {
                const VoxTile v = vox_decode(p, m_ct);
                bw = v.x + p.lc_w; bh = v.y + p.lc_h; bd = v.z + p.lc_d;
                nf = (v.fb * 2 + static_cast<int>(cta_rank)) * 128;
                a0 = v.a0; b0 = v.b0; b1 = v.b1; c0 = v.c0; c1 = v.c1;
                n_kb = v.n_kb;
            } else {
                int m0 = (m_ct * 2 + static_cast<int>(cta_rank)) * 128;
                if (m0 >= p.m_total) m0 = 0;           // dummy half: rows are discarded by the epilogue
                const int q = m0 % p.Wo;
                int t = m0 / p.Wo;
                const int pp = t % p.Ho;
                t /= p.Ho;
                const int z = t % p.Do;
                nf = t / p.Do;
                bw = q + p.lc_w; bh = pp + p.lc_h; bd = z + p.lc_d;
            }
            const int n_groups = (n_kb + p.kg - 1) / p.kg;
            const int w_row0 = n_idx * p.n_tile + static_cast<int>(cta_rank) * half_rows;
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
                        if (leader_cta) mbar_arrive(&full_bar[s]);
                    } else {
                        if (leader_cta) mbar_expect_tx(&full_bar[s], 2u * static_cast<uint32_t>(nkb) * kb_bytes);
                        uint8_t* st = smem + static_cast<size_t>(s) * stage_bytes;
                        for (int j = 0; j < nkb; ++j) {
                            const int tap = (ta * p.kh + tb) * p.kw + tc;
                            uint8_t* base = st + static_cast<size_t>(j) * kb_bytes;
                            if (p.vox) {
                                tma_load_5d_pair(base, &map_v, &full_bar[s], cb * p.kc, bw + tc, bh + tb, bd + ta, nf, pol_a);
                                tma_load_5d_pair(base + p.a_sub_bytes, &map_v, &full_bar[s], cb * p.kc, bw + tc, bh + tb, bd + ta,
                                                 nf + p.lo_plane_frames, pol_a);
                            } else {
                                tma_load_im2col_5d_pair(base, &map_a, &full_bar[s], cb * p.kc, bw, bh, bd, nf,
                                                        static_cast<uint16_t>(tc), static_cast<uint16_t>(tb),
                                                        static_cast<uint16_t>(ta));
                                tma_load_im2col_5d_pair(base + p.a_sub_bytes, &map_a, &full_bar[s], cb * p.kc, bw, bh,
                                                        bd, nf + p.lo_plane_frames, static_cast<uint16_t>(tc),
                                                        static_cast<uint16_t>(tb), static_cast<uint16_t>(ta));
                            }
                            uint8_t* wb = base + 2 * p.a_sub_bytes;
                            const int kcoord = tap * p.cin_pad + cb * p.kc;
                            tma_load_2d_pair(wb, &map_w, &full_bar[s], kcoord, w_row0);
                            tma_load_2d_pair(wb + p.w_sub_bytes, &map_w, &full_bar[s], kcoord, p.w_lo_rows + w_row0);
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
    } else if (warp == 1) {
        // =============================================================== MMA issuer (leader CTA only)
        if (leader_cta) {
            const bool leader = elect_one();
            const uint32_t idesc = umma_idesc_bf16_m256(static_cast<uint32_t>(p.n_tile));
            const uint32_t desc_hi = ((p.row_bytes * 8u) >> 4) | (1u << 14) | (p.layout_type << 29);
            const uint32_t lo_flags = 1u << 16;
            const int k16_steps = p.kc / 16;
            const uint32_t a_sub16 = p.a_sub_bytes >> 4, w_sub16 = p.w_sub_bytes >> 4, kb16 = kb_bytes >> 4;
            const uint32_t smem_base16 = (smem_u32(smem) & 0x3FFFFu) >> 4;
            int s = 0;
            uint32_t ph = 0;
            int acc = 0;
            uint32_t acc_ph = 0;
            for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
                const int n_kb = p.vox ? vox_decode(p, tile / p.n_tiles).n_kb : p.n_kblocks;
                const int n_groups = (n_kb + p.kg - 1) / p.kg;
                mbar_wait(&tempty_bar[acc], acc_ph ^ 1u);
                tc_fence_after();
                uint32_t accumulate = 0;
                const uint32_t d0 = tmem_base + static_cast<uint32_t>(acc * p.acc_cols);
                for (int g = 0; g < n_groups; ++g) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const int nkb = min(p.kg, n_kb - g * p.kg);
                    uint32_t base16 = (smem_base16 + static_cast<uint32_t>(s) * (stage_bytes >> 4)) | lo_flags;
                    if (!TB_DBG(p.dbg, 2)) {
                        for (int j = 0; j < nkb; ++j, base16 += kb16) {
                            for (int kk = 0; kk < k16_steps; ++kk) {
                                const uint32_t a_hi = base16 + static_cast<uint32_t>(kk) * 2u;   // +32 B per K=16
                                const uint32_t a_lo = a_hi + a_sub16;
                                const uint32_t w_hi = a_hi + 2u * a_sub16;
                                const uint32_t w_lo = w_hi + w_sub16;
                                // precise graphs (corr_off != 0): the two correction products go to their own accumulator
                                // (3x less accumulator truncation); TMEM then holds one stage only
                                const uint32_t dc = d0 + static_cast<uint32_t>(p.corr_off);
                                umma_bf16_pair(leader, d0, a_hi, w_hi, desc_hi, idesc, accumulate);
                                umma_bf16_pair(leader, dc, a_lo, w_hi, desc_hi, idesc, p.corr_off ? accumulate : 1u);
                                umma_bf16_pair(leader, dc, a_hi, w_lo, desc_hi, idesc, 1u);
                                accumulate = 1u;
                            }
                        }
                    }
                    if (leader) umma_commit_pair(&empty_bar[s]);      // frees the stage in both CTAs
                    __syncwarp();
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
                if (leader) umma_commit_pair(&tfull_bar[acc]);        // accumulators complete in both CTAs
                __syncwarp();
                if (++acc == p.acc_stages) { acc = 0; acc_ph ^= 1u; }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        // =============================================================== epilogue (warps 4..11, both CTAs)
        const int quad = warp & 3;
        const int half = (warp - kPairEpilogueWarp0) >> 2;
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
            mbar_wait_relaxed(&tfull_bar[acc], acc_ph);   // a whole mainloop away: sleep between polls (0.7 % on conv5)
            tc_fence_after();
            int64_t m = static_cast<int64_t>(m_ct * 2 + static_cast<int>(cta_rank)) * 128 + row_in_tile;
            bool row_ok = m < p.m_total && !TB_DBG(p.dbg, 8);
            if (p.vox) {                               // row = frame, at the tile's voxel
                const int n_vox = p.Do * p.Ho * p.Wo;
                const int fb = m_ct / n_vox;
                const int frame = (fb * 2 + static_cast<int>(cta_rank)) * 128 + row_in_tile;
                m = static_cast<int64_t>(frame) * n_vox + (m_ct - fb * n_vox);
                row_ok = frame < p.vox_frames && !TB_DBG(p.dbg, 8);
            }
            const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                   static_cast<uint32_t>(acc * p.acc_cols);
            if (p.corr_off && !TB_DBG(p.dbg, 4)) {
                // separate correction accumulator (the default): main + correction fill TMEM, so the stage is drained
                // into registers and released before the epilogue math runs (conv_umma.cuh, epilogue_drained)
                uint64_t* bar = &tempty_bar[acc];
                epilogue_drained<ACT1, ACT2, FMT>(p, tbase, half, chunks, n_idx * p.n_tile, m, row_ok, bias_v, scale_v,
                                                  shift_v, [&] { if (lane == 0) mbar_arrive_cluster(bar, 0); });
                if (p.progress && leader_cta && warp == kPairEpilogueWarp0 && lane == 0) atomicAdd(p.progress, 1);
                if (++acc == p.acc_stages) { acc = 0; acc_ph ^= 1u; }
                continue;
            }
            for (int c = half; c < chunks && !TB_DBG(p.dbg, 4); c += 2) {
                uint32_t r[16];
                __syncwarp();                      // tcgen05.ld is .sync.aligned
                tmem_ld_32x32b_x16(tbase + static_cast<uint32_t>(c * 16), r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * p.acc_comp);
                const int n0 = n_idx * p.n_tile + c * 16;
                if (n0 >= p.c_store) continue;     // warp-uniform
                epilogue_chunk<ACT1, ACT2, FMT>(p, r, n0, m, row_ok, bias_v, scale_v, shift_v);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);
            if (p.progress && leader_cta && warp == kPairEpilogueWarp0 && lane == 0) atomicAdd(p.progress, 1);
            if (++acc == p.acc_stages) { acc = 0; acc_ph ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // the leader's MMAs read the peer's shared memory and write its TMEM until here
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_512_pair(tmem_base);
    }
}

#endif  // __CUDACC__

}  // namespace tb
