// libtimed_b200.so -- C-ABI (include/timed_b200.h): inference-graph executor + sampler entry
// points.  Host side: shape inference, workspace planning, weight packing (bf16 hi/lo planes,
// K-major), TMA tensor-map construction, kernel launches.  No CPU compute fallback exists: every
// entry point either launches CUDA kernels or fails with an error code.
#include "../../include/timed_b200.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <atomic>
#include <zlib.h>
#include <vector>

#include "common.cuh"
#include "conv_umma.cuh"
#include "conv_pair.cuh"
#include "kernels.cuh"
#include "thin_conv.cuh"
#include "slab_conv.cuh"
#include "thinz_conv.cuh"
#include "xform_conv.cuh"
#include "pdb_parse.cuh"
#include "inflate.cuh"
#include "hdf5_index.cuh"
#include "voxelise.cuh"

namespace tb {

static thread_local std::string g_last_error;

// Role-timing switches (skip copies / MMA issue / epilogue) exist only in bring-up builds (-DTIMED_B200_DEBUG): the
// release library ignores TIMED_B200_DBG, so no environment variable can make a kernel return garbage.
static int debug_mask() {
#ifdef TIMED_B200_DEBUG
    static const int dbg = [] { const char* e = getenv("TIMED_B200_DBG"); return e ? atoi(e) : 0; }();
    return dbg;
#else
    return 0;
#endif
}
void set_error(const std::string& msg) { g_last_error = msg; }

// ----------------------------------------------------------------------------- driver entry points
// libcuda is NOT linked: the tensor-map encoders are fetched through the runtime so that the
// library loads (and exports its symbols) on a machine without a GPU driver.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_driver_version = 0;

static int load_driver_fns() {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (g_encode_tiled && g_encode_im2col) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    TB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    TB_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
    g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    fn = nullptr;
    TB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres));
    TB_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeIm2col unavailable");
    g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
    TB_CHECK_CUDA(cudaDriverGetVersion(&g_driver_version));
    return 0;
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// TF/Keras 'same' padding for (n, k, s): out = ceil(n/s), total = max((out-1)*s+k-n, 0),
// before = total/2.
static void same_pads(int n, int k, int s, int* out, int* before, int* after) {
    *out = ceil_div(n, s);
    const int total = std::max((*out - 1) * s + k - n, 0);
    *before = total / 2;
    *after = total - total / 2;
}

static int grid_for(int64_t work_items, int threads) {
    int64_t b = (work_items + threads - 1) / threads;
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(b, 148 * 32)));
}

// ----------------------------------------------------------------------------- tensors
struct TensorInfo {
    int D = 1, H = 1, W = 1, C = 0;
    int c_pad = 0;        // stored channels per pixel (SPLIT: round_up(C,16); F32: C)
    int fmt = FMT_F32;
    int slack_pix = 0;    // extra readable pixels past the end required by im2col consumers
    int last_use = -1;    // index of the last op reading this tensor
    // Zero-copy Concatenate: an fp32 tensor that is only ever read as a channel slice of a later Concatenate's output is a
    // VIEW of that output's buffer (producers write their slice with the buffer's row pitch; the Concatenate copies nothing).
    // Views nest (DenseNet blocks: concat_k is the first slice of concat_{k+1}); view_root / view_c0 are resolved.
    int view_of = -1, view_off = 0;
    int view_root = -1, view_c0 = 0;
    // "W-folded" layout of a thin (C <= 8) graph input: 8 stored channels per pixel and W-rows of
    // wf_pitch pixels (wf_lm zero pixels, the W real ones, zero pixels up to the pitch).  A conv
    // then reads kwin consecutive pixels (kwin*8 contiguous elements) as ONE im2col "pixel", so the
    // kw taps of a row collapse into the K dimension: 3x fewer TMA rows and MMAs for TIMED's
    // first layer (DESIGN.md "First layer").
    bool wfold = false;
    int wf_lm = 0, wf_pitch = 0;
    // "padded volume" layout of a thin graph input read by thin_conv_kernel: 8 stored channels, zero
    // margins materialised in all three dimensions (pv_*0 before, up to the padded extents after).
    bool padvol = false;
    int pv_d0 = 0, pv_h0 = 0, pv_w0 = 0, pv_Dp = 0, pv_Hp = 0, pv_Wp = 0;
    // "chunk-plane padded volume" read by slab_conv_kernel (slab_conv.cuh): plane x 8-channel chunk x position
    // x 8 bf16, positions = lead + one linearisation of all frames with shared zero margins + tail.
    bool cpv = false;
    int cpv_Dp = 0, cpv_Hp = 0, cpv_Wp = 0;
    int64_t cpv_lead = 0, cpv_tail = 0;
    int64_t cpv_T(int64_t n) const { return cpv_lead + n * cpv_Dp * cpv_Hp * cpv_Wp + cpv_tail; }
    int64_t pix_per_frame() const { return static_cast<int64_t>(D) * H * W; }
    int64_t stored_pix_per_frame() const {
        if (cpv) return static_cast<int64_t>(cpv_Dp) * cpv_Hp * cpv_Wp;
        if (padvol) return static_cast<int64_t>(pv_Dp) * pv_Hp * pv_Wp;
        return static_cast<int64_t>(D) * H * (wfold ? wf_pitch : W);
    }
    // frames allocated per plane for n frames (so that a 128-pixel im2col column starting at
    // any valid pixel stays inside the allocation)
    int64_t frames_alloc(int64_t n) const {
        return n + (slack_pix + pix_per_frame() - 1) / pix_per_frame();
    }
    size_t bytes(int64_t n) const {
        if (cpv) return static_cast<size_t>(round_up64(2 * (c_pad / 8) * cpv_T(n) * 16, 1024));
        const int64_t elems = frames_alloc(n) * stored_pix_per_frame() * c_pad;
        return static_cast<size_t>(round_up64(fmt == FMT_SPLIT ? elems * 2 * 2 : elems * 4, 1024));
    }
};

// ----------------------------------------------------------------------------- conv plan
struct ConvPlan {
    // static (graph_create)
    int kd = 1, kh = 1, kw = 1;
    int cin = 0, cin_pad = 0, cout = 0;
    int pad0[3] = {0, 0, 0}, pad1[3] = {0, 0, 0};
    int Di = 1, Hi = 1, Wi = 1, Do = 1, Ho = 1, Wo = 1;
    int n_tile = 0, n_tiles = 0, n_alloc = 0;
    int k_total = 0;
    // W-folded input (see TensorInfo): kwin pixels x 8 channels form one K-block of a (kd,kh) tap
    bool wfold = false;
    int kwin = 0, in_lm = 0, in_pitch = 0;
    // "tap-to-N" formulation for convs with few output channels (TIMED's 20-class head): instead of
    // taps x (C_in/16) thin MMAs of N = C_out per tile, run ONE 1x1x1 GEMM  Z[pixel, (tap, co)] =
    // sum_c X[pixel, c] * W[tap, c, co]  with N = taps*C_out (wide MMAs, every activation read once)
    // followed by a col2im gather  out[p, co] = sum_tap Z[p + tap - pad, tap, co]  (+ bias/act/BN).
    // thin-input path (thin_conv.cuh): padded-volume input, kw taps aliased by the UMMA descriptor
    bool precise = true;             // graph option (default on): wide tiles keep the correction products in their own TMEM accumulator
    bool slab = false;               // chunk-plane padded-volume input, slab_conv_kernel
    SlabConvParams slab_params;      // static part, completed per launch
    uint8_t* d_slab_w = nullptr;
    int64_t slab_lead = 0, slab_tail = 0;
    bool thin = false;
    bool fuse_zpool = false;         // thinz only: the z direction of the MaxPool(2,2,2;2) that follows runs in the epilogue
    bool fuse_pool = false;          // thinz only: the MaxPool(2,2,2;2) that follows runs in the conv epilogue
    int pool_same = 0, pool_Zo = 0, pool_Po = 0, pool_Qo = 0;
    bool thinz = false;              // kd taps folded into N (thinz_conv.cuh)
    ThinZParams thinz_params;
    size_t thinz_merged_off = 0, thinz_merged_bytes = 0;   // the fused-pool instantiation's weight layout inside d_thin_w
    ThinConvParams thin_params;      // static part, completed per launch
    uint8_t* d_thin_w = nullptr;
    bool tap2n = false;
    int z_cols = 0, z_ld = 0;
    float* d_c2i_bias = nullptr;     // [cout] epilogue vectors applied by the col2im kernel
    float* d_c2i_scale = nullptr;
    float* d_c2i_shift = nullptr;
    // tap-to-N with the kw taps kept in K ("t2n_kw"): Z[pixel, (kd,kh), co] = sum_{kw,c} X[pixel + kw - pad_w, c] * W, an
    // im2col along W only.  Same MMA count and activation traffic as folding all taps into N (three N tiles there, three
    // K taps here) but a Z matrix kw times smaller to write and to gather from.
    bool t2n_kw = false;
    // tap-to-N with the kw taps in N and the (kd,kh) taps in K ("t2n_w"): Z[(d,h) output row, w input column, kw, co] =
    // sum_{kd,kh,c} X * W, an im2col over D and H only; col2im then sums the kw shifted copies along W.  N = kw*cout columns
    // (96 for DenseCPD's 128 -> 32 growth convs: two N-folded MMAs of N = 192 / 96 per K step instead of N = 64 / 32 -- the
    // thin-N MMAs cost the same ~72 cycles -- for a Z matrix of only 384 bytes per pixel).
    bool t2n_w = false;
    // t2n_w with kw = 3, 'same', C_out = 16 / 32: the col2im over kw runs in the GEMM's epilogue on tiles of whole volume
    // rows (ConvKernelParams::c2i) -- no Z matrix in HBM, no col2im launch.  Not when the conv is the fused network head.
    // 1x1x1 conv whose input is a BatchNorm -> ReLU of an fp32 tensor read by nothing else (DenseNet pre-activation): the
    // affine + ReLU + bf16 split run in the conv's operand path (xform_conv.cuh), the AFFINE op is skipped and the conv reads
    // tensor `xf_src` (the AFFINE's input).  xf_scale / xf_shift are owned by the AFFINE node.
    bool xform = false;
    int xf_src = -1;
    const float* xf_scale = nullptr;
    const float* xf_shift = nullptr;
    bool c2i = false;
    int c2i_rows = 0;
    bool c2i_active() const { return c2i && !fuse_head; }
    // network head: this tap-to-N conv is read only by GlobalPooling -> Softmax (the graph output): the col2im gather, the
    // pooling and the softmax run as ONE launch after the GEMM (head_col2im_pool_softmax_kernel) straight into `probs`
    bool fuse_head = false;
    int head_is_avg = 1;
    // network head, linear conv -> GlobalAveragePooling -> (Softmax): the average commutes with the conv, so the launch is a
    // box sum of the input per filter tap (gap_boxsum_kernel), ONE dense GEMM over K = taps*cin (`dense`, a 1x1x1 plan over
    // the same DHWIO weights) and the softmax -- D*H*W times fewer MMAs than convolving every voxel
    bool gap_collapse = false;
    bool gap_softmax = false;        // the Softmax that follows runs in this launch too (else: pooled logits are the output)
    ConvPlan* dense = nullptr;       // owned
    int32_t* d_progress = nullptr;   // completed-tile counter of the voxel-stationary pair kernel (ConvKernelParams::progress)
    int taps_eff() const { return tap2n ? (t2n_kw ? kw : t2n_w ? kd * kh : 1) : (wfold ? kd * kh : kd * kh * kw); }
    // geometry of the GEMM rows (output pixels, or input pixels for tap-to-N)
    int Mo_d() const { return tap2n && !t2n_w ? Di : Do; }
    int Mo_h() const { return tap2n && !t2n_w ? Hi : Ho; }
    int Mo_w() const { return tap2n ? (t2n_kw ? Wo : Wi) : Wo; }
    int gemm_n() const { return tap2n ? z_cols : cout; }
    __nv_bfloat16* d_w = nullptr;   // [2][n_alloc][k_total]
    float* d_bias = nullptr;        // [n_alloc]
    float* d_scale = nullptr;
    float* d_shift = nullptr;
    int act1 = 0, act2 = 0;
    float alpha1 = 1.f, alpha2 = 1.f;
    // tile configuration (depends on the frame count -> chosen at launch)
    struct Config {
        int kc, mt, kg, stages, acc_stages, acc_cols, nfold, cluster2, corr_off, pair;
        uint32_t swizzle_code;       // UMMA layout type
        CUtensorMapSwizzle tma_swz;
        size_t smem_bytes;
    };
    std::map<int, CUtensorMap> w_maps;   // keyed by kc * 1024 + box rows
    double flops_per_frame() const {
        return 2.0 * Do * Ho * Wo * kd * kh * kw * static_cast<double>(cin) * cout;
    }
};

static void free_conv_plan(ConvPlan& p) {
    cudaFree(p.d_progress);
    p.d_progress = nullptr;
    if (p.dense) {
        free_conv_plan(*p.dense);
        delete p.dense;
        p.dense = nullptr;
    }
    cudaFree(p.d_w);
    cudaFree(p.d_bias);
    cudaFree(p.d_scale);
    cudaFree(p.d_shift);
    cudaFree(p.d_thin_w);
    p.d_thin_w = nullptr;
    cudaFree(p.d_slab_w);
    p.d_slab_w = nullptr;
    cudaFree(p.d_c2i_bias);
    cudaFree(p.d_c2i_scale);
    cudaFree(p.d_c2i_shift);
    p.d_w = nullptr;
    p.d_bias = p.d_scale = p.d_shift = nullptr;
    p.d_c2i_bias = p.d_c2i_scale = p.d_c2i_shift = nullptr;
}

// Accumulator-truncation compensation factor for a main accumulator that receives `n_nominal` MMAs per output of which
// the fraction `valid` has non-zero operands (zero 'same' padding adds exact zeros: no truncation).  Measured on the
// B200 (tools/accum_error.py, profiles/r2_accum_error.jsonl): operands exactly representable in bf16, fp64 reference,
// relative shrink of the result = c(n)*n with c rising slowly from 1.27e-8 (n = 32) to 1.71e-8 (n = 864) per
// EFFECTIVE MMA, identical for positive and negative results (truncation toward zero).
static float accum_comp(double n_nominal, double valid) {
    if (getenv("TIMED_B200_NO_ACC_COMP")) return 1.0f;
    const double n = n_nominal * valid;
    if (n < 1.0) return 1.0f;
    const double c = 1.30e-8 + 0.09e-8 * std::log2(std::max(n, 32.0) / 32.0);
    return static_cast<float>(1.0 + c * n);
}

// mean over the output positions of one axis of (taps that fall inside the input) / (taps)
static double valid_tap_fraction(int in, int out, int k, int pad0) {
    double s = 0.0;
    for (int o = 0; o < out; ++o)
        for (int t = 0; t < k; ++t) s += (o + t - pad0 >= 0 && o + t - pad0 < in) ? 1.0 : 0.0;
    return s / (static_cast<double>(out) * k);
}

// 227 KB per CTA is the limit for static + dynamic shared memory; the kernel keeps ~6.3 KB static
// (barriers + staged epilogue vectors), and 1 KB of the dynamic part is alignment slack.
static constexpr size_t kSmemDynamicMax = 220 * 1024;
static constexpr size_t kHeadSmemMax = 96 * 1024;      // fused head: one frame's (pixels, classes) activation in shared memory
static constexpr size_t kSmemBudget = kSmemDynamicMax - 1024;
// conv_umma_kernel keeps its static shared memory under 2 KB (kUmmaEpiSmemN) and may use 225 KB of operand ring: exactly four
// stages of the 56 KB k-blocks of DenseCPD's growth convs, which three stages left latency-bound
static constexpr size_t kUmmaSmemDynamicMax = 225 * 1024;
static constexpr size_t kUmmaSmemBudget = kUmmaSmemDynamicMax - 1024;

// Pick (kc, mt, kg, stages) for a conv given the number of output rows.
static int choose_config(const ConvPlan& p, int64_t m_total, ConvPlan::Config* cfg) {
    const int m_tiles = static_cast<int>((m_total + 127) / 128);
    std::vector<int> kcs;
    if (p.wfold) {
        kcs.push_back(p.cin_pad);            // one K-block per (kd,kh) tap: kwin*8 = 32 or 64
    } else {
        for (int kc : {64, 32, 16})
            if (p.cin_pad % kc == 0) kcs.push_back(kc);
    }
    TB_REQUIRE(!kcs.empty(), "conv: padded input channels must be a multiple of 16");
    cfg->cluster2 = 0;
    cfg->corr_off = 0;
    cfg->pair = 0;
    // "precise" graphs, full-width tiles (n_tile = 256) with plenty of rows: 2-CTA cluster, W tiles
    // multicast, one 128-row sub-tile per CTA with separate main / correction accumulators.  Same L2
    // traffic as mt = 2 and 3x less accumulator truncation, but every CTA now stages the whole W tile for
    // half the rows, so shared memory covers ~1/3 less latency: measured 1.4x slower on the two big TIMED
    // layers (profiles/r1_summary.md) -- hence opt-in.
    if (p.precise && p.n_tile == 256 && p.cin_pad % 32 == 0 && getenv("TIMED_B200_NO_PAIR") &&
        static_cast<int64_t>(ceil_div(m_tiles, 2)) * p.n_tiles >= 2 * 74) {
        const int kc = 32;
        const size_t kb = 2 * (128u * kc * 2u) + 2 * (static_cast<size_t>(p.n_tile) * kc * 2u);
        cfg->kc = kc;
        cfg->mt = 1;
        cfg->kg = 1;
        cfg->stages = static_cast<int>(std::min<size_t>(kConvMaxStages, kSmemBudget / kb));
        cfg->acc_cols = 256;
        cfg->acc_stages = 1;
        cfg->nfold = 0;
        cfg->cluster2 = 1;
        cfg->corr_off = 256;
        cfg->swizzle_code = 4u;
        cfg->tma_swz = CU_TENSOR_MAP_SWIZZLE_64B;
        cfg->smem_bytes = kb * cfg->stages + 1024;
        return 0;
    }
    // Wide tiles (n_tile > 128) with plenty of rows: CTA pair, tcgen05.mma.cta_group::2 (conv_pair.cuh).
    // Each CTA stages 128 rows of A and half of the W tile, so TMEM holds two accumulator stages (the
    // epilogue overlaps the next mainloop) and an MMA reads 8 KB instead of 12 KB of shared memory.
    if (p.n_tile > 128 && p.cin_pad % 32 == 0 && !getenv("TIMED_B200_NO_PAIR") &&
        (static_cast<int64_t>(ceil_div(m_tiles, 2)) * p.n_tiles >= 2 * 74 || getenv("TIMED_B200_FORCE_PAIR"))) {
        int kc = p.cin_pad % 64 == 0 ? 64 : 32;
        if (const char* e = getenv("TIMED_B200_PAIR_KC")) {
            const int v = atoi(e);
            if ((v == 32 || v == 64) && p.cin_pad % v == 0) kc = v;
        }
        const size_t kb = 2 * (128u * kc * 2u) + 2 * (static_cast<size_t>(p.n_tile / 2) * kc * 2u);
        const int n_kblocks = p.taps_eff() * (p.cin_pad / kc);
        cfg->kc = kc;
        cfg->mt = 1;
        cfg->kg = 1;
        cfg->stages = static_cast<int>(std::min<size_t>(kConvMaxStages, kSmemBudget / kb));
        cfg->stages = std::max(2, std::min(cfg->stages, std::max(2, n_kblocks)));
        cfg->acc_cols = round_up(p.n_tile, 32);
        cfg->acc_stages = std::min(2, 512 / cfg->acc_cols);
        if (p.precise) {             // separate correction accumulator next to the main one: one stage fills TMEM
            cfg->corr_off = cfg->acc_cols;
            cfg->acc_stages = 1;
        }
        cfg->nfold = 0;
        cfg->pair = 1;
        cfg->swizzle_code = kc == 64 ? 2u : 4u;
        cfg->tma_swz = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
        cfg->smem_bytes = kb * cfg->stages + 1024;
        return 0;
    }
    const bool nfold = p.n_tile <= 128 && !getenv("TIMED_B200_NO_NFOLD");
    const int acc_cols = round_up(nfold ? 2 * p.n_tile : p.n_tile, 32);
    // two M sub-tiles per CTA halve the weight traffic per MAC; only worth it when there are
    // enough tiles to still fill the machine twice over
    // wide tiles (no N-fold) of a precise graph: one M sub-tile, main and correction accumulators side by side
    const bool sep_corr = p.precise && !nfold && 2 * acc_cols <= 512;
    std::vector<int> mts;
    if (!sep_corr && !p.c2i_active() && 2 * acc_cols <= 512 && static_cast<int64_t>(ceil_div(m_tiles, 2)) * p.n_tiles >= 2 * 148)
        mts.push_back(2);
    mts.push_back(1);
    int best_score = INT_MIN;
    for (int mt : mts) {
        for (int kc : kcs) {
            const size_t a_sub = 128u * kc * 2u, w_sub = static_cast<size_t>(p.n_tile) * kc * 2u;
            const size_t kb = mt * 2 * a_sub + 2 * w_sub;
            const int n_kblocks = p.taps_eff() * (p.cin_pad / kc);
            int kg = std::max(1, 64 / kc);
            kg = std::min(kg, n_kblocks);
            while (kg > 1 && kb * kg * 2 > kUmmaSmemBudget) --kg;
            if (kb * kg * 2 > kUmmaSmemBudget) continue;   // cannot even double-buffer
            int stages = static_cast<int>(std::min<size_t>(kConvMaxStages, kUmmaSmemBudget / (kb * kg)));
            // score: prefer >=3 stages, then mt=2, then wider kc; a single accumulator stage
            // serialises the epilogue with the mainloop, which only a long K loop amortises
            const int acc_stages_c = sep_corr ? 1 : std::min(2, 512 / (mt * acc_cols));
            const int mainloop_mmas = n_kblocks * (kc / 16) * 3 * mt;
            const int score = (stages >= 3 ? 100 : 0) + (mt == 2 ? 10 : 0) + kc / 16 -
                              ((acc_stages_c == 1 && mainloop_mmas < 1500) ? 20 : 0);
            if (score > best_score) {
                best_score = score;
                cfg->kc = kc;
                cfg->mt = mt;
                cfg->kg = kg;
                cfg->stages = stages;
                cfg->acc_cols = acc_cols;
                cfg->nfold = nfold ? 1 : 0;
                cfg->acc_stages = acc_stages_c;
                cfg->corr_off = sep_corr ? acc_cols : 0;
                cfg->swizzle_code = kc == 64 ? 2u : (kc == 32 ? 4u : 6u);
                cfg->tma_swz = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                        : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                    : CU_TENSOR_MAP_SWIZZLE_32B);
                cfg->smem_bytes = kb * kg * stages + 1024;
            }
        }
    }
    TB_REQUIRE(best_score != INT_MIN, "conv: no tile configuration fits shared memory");
    return 0;
}

static int encode_w_map(ConvPlan& p, int kc, CUtensorMapSwizzle swz, CUtensorMap* out, int box_rows = 0) {
    if (box_rows <= 0) box_rows = p.n_tile;
    const int key = kc * 1024 + box_rows;
    auto it = p.w_maps.find(key);
    if (it != p.w_maps.end()) { *out = it->second; return 0; }
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(p.k_total), static_cast<cuuint64_t>(2 * p.n_alloc)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(p.k_total) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kc), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = g_encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.d_w, dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(weights) failed, CUresult=" + std::to_string(r));
        return TB_ERR_CUDA;
    }
    p.w_maps[key] = m;
    *out = m;
    return 0;
}

// Activation tensor map (im2col).  Tensor = (2*frames_alloc, Di, Hi, Wi, cin_pad) bf16, the lo
// plane stacked after the hi plane along the frame axis.
static int encode_a_map(const ConvPlan& p, const ConvPlan::Config& cfg, void* base,
                        int64_t frames_alloc, int64_t ld_channels, CUtensorMap* out) {
    cuuint64_t dims[5] = {static_cast<cuuint64_t>(p.cin_pad), static_cast<cuuint64_t>(p.Wi),
                          static_cast<cuuint64_t>(p.Hi), static_cast<cuuint64_t>(p.Di),
                          static_cast<cuuint64_t>(2 * frames_alloc)};
    const cuuint64_t px = static_cast<cuuint64_t>(ld_channels) * 2;
    cuuint64_t strides[4] = {px, px * p.Wi, px * p.Wi * p.Hi, px * p.Wi * p.Hi * p.Di};
    int lower[3] = {-p.pad0[2], -p.pad0[1], -p.pad0[0]};
    int upper[3] = {p.pad1[2] - (p.kw - 1), p.pad1[1] - (p.kh - 1), p.pad1[0] - (p.kd - 1)};
    if (p.tap2n) {      // index 0 = W, 1 = H, 2 = D: t2n_kw keeps the W window, t2n_w the D and H windows
        for (int i = p.t2n_kw ? 1 : 0; i < (p.t2n_w ? 1 : 3); ++i) lower[i] = upper[i] = 0;
    }
    if (p.wfold) {
        // "pixel" = kwin consecutive stored pixels starting pad_w0 to the left of the output
        // column; consecutive "pixels" overlap in memory (stride = one stored pixel = 16 bytes).
        // The W extent of the map is the number of output columns; no W padding/filter extent left.
        dims[0] = static_cast<cuuint64_t>(p.kwin * 8);
        dims[1] = static_cast<cuuint64_t>(p.Wo);
        const cuuint64_t row = static_cast<cuuint64_t>(p.in_pitch) * 16;
        strides[0] = 16;
        strides[1] = row;
        strides[2] = row * p.Hi;
        strides[3] = row * p.Hi * p.Di;
        lower[0] = 0;
        upper[0] = 0;
        base = static_cast<uint8_t*>(base) + static_cast<size_t>(p.in_lm - p.pad0[2]) * 16;
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUtensorMap m;
    CUresult r = g_encode_im2col(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, lower,
                                 upper, static_cast<cuuint32_t>(cfg.kc), 128u, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, cfg.tma_swz,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeIm2col(activations) failed, CUresult=" + std::to_string(r));
        return TB_ERR_CUDA;
    }
    // Same driver workaround CUTLASS applies (cute/atom/copy_traits_sm90_im2col.hpp): drivers up
    // to 13.1 mis-set a descriptor bit for im2col maps over tensors smaller than 128 KiB.
    const uint64_t tensor_bytes = strides[3] * dims[4];   // (wfold: same formula, padded rows)
    if (g_driver_version <= 13010 && tensor_bytes < 131072)
        reinterpret_cast<uint64_t*>(&m)[1] &= ~(1ull << 21);
    *out = m;
    return 0;
}

// Tiled map over the same activation tensor for voxel-stationary tiles (conv_pair.cuh): box = kc channels of ONE
// voxel x 128 consecutive frames.
static int encode_a_vox_map(const ConvPlan& p, const ConvPlan::Config& cfg, void* base, int64_t frames_alloc,
                            CUtensorMap* out) {
    cuuint64_t dims[5] = {static_cast<cuuint64_t>(p.cin_pad), static_cast<cuuint64_t>(p.Wi),
                          static_cast<cuuint64_t>(p.Hi), static_cast<cuuint64_t>(p.Di),
                          static_cast<cuuint64_t>(2 * frames_alloc)};
    const cuuint64_t px = static_cast<cuuint64_t>(p.cin_pad) * 2;
    cuuint64_t strides[4] = {px, px * p.Wi, px * p.Wi * p.Hi, px * p.Wi * p.Hi * p.Di};
    cuuint32_t box[5] = {static_cast<cuuint32_t>(cfg.kc), 1, 1, 1, 128};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, cfg.tma_swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(activations, voxel-stationary) failed, CUresult=" + std::to_string(r));
        return TB_ERR_CUDA;
    }
    return 0;
}

// ----------------------------------------------------------------------------- thin-input conv plan
// Number of K=16 MMA steps thin_conv_kernel needs for a (kd,kh,kw) filter: kw/2 in-row pixel pairs per
// filter row plus the left-over odd taps paired across consecutive filter rows.
static int thin_step_count(int kd, int kh, int kw) {
    const int rows = kd * kh;
    return rows * (kw / 2) + ((kw & 1) ? (rows + 1) / 2 : 0);
}

// thin_conv_kernel applies when the K-step count is bounded (kThinMaxSteps), the output channels fit one N tile and the
// resident weights plus two pipeline stages fit shared memory.
static bool thin_fits(int kd, int kh, int kw, int cout) {
    const int steps = thin_step_count(kd, kh, kw);
    const int n_tile = round_up(cout, 16);
    if (steps > kThinMaxSteps || n_tile > 128 || kd < 1 || kh < 1 || kw < 1) return false;   // folded MMA: N = 2*n_tile <= 256
    const size_t w_smem = (static_cast<size_t>(2 * steps) * n_tile * 8 * 2 * 2 + 127) & ~static_cast<size_t>(127);
    const size_t span = static_cast<size_t>((128 + kw + 2) & ~1) * 16;
    return w_smem + 2 * (2u * kd * kh * span) + 128 <= kSmemDynamicMax;
}

static int thin_plan_create(ConvPlan& p, const tb_op_desc& d, const TensorInfo& tin) {
    p.thin = true;
    const int rows = p.kd * p.kh;
    const int n_tile = round_up(p.cout, 16);
    TB_REQUIRE(n_tile <= 128, "thin conv: too many output channels");
    p.n_tiles = 1;
    p.n_tile = p.n_alloc = n_tile;
    // stored pixels per span: row r of a K step reads pixels r+j and r+j+1 (j <= kw-1, the partner of a
    // partner-less left-over tap aliases the next pixel), so 128 + kw + 1 pixels, rounded to even.  Every
    // aliased read must stay inside the stage: an out-of-allocation smem read by the tensor core silently
    // corrupted the last pipeline stage.
    const int span_pix = (128 + p.kw + 1 + 1) & ~1;
    const int span_bytes = span_pix * 16;
    const int span_stride = round_up(span_bytes, 16);
    ThinConvParams& t = p.thin_params;
    std::memset(&t, 0, sizeof(t));
    // ---- K=16 steps and the matching weight K order
    struct Half { int row, kwi; };                                  // (filter row, kw index) or row=-1: zero
    std::vector<std::pair<Half, Half>> steps;
    for (int r = 0; r < rows; ++r)
        for (int j = 0; j + 1 < p.kw; j += 2) steps.push_back({{r, j}, {r, j + 1}});
    if (p.kw & 1)
        for (int r = 0; r < rows; r += 2)
            steps.push_back({{r, p.kw - 1}, {r + 1 < rows ? r + 1 : -1, p.kw - 1}});
    TB_REQUIRE(static_cast<int>(steps.size()) <= kThinMaxSteps, "thin conv: too many K steps");
    t.n_steps = static_cast<int>(steps.size());
    // ---- weights: [2*n_steps chunks][2*n_tile rows: hi then lo][8]
    const size_t w_elems = static_cast<size_t>(2 * t.n_steps) * 2 * n_tile * 8;
    std::vector<__nv_bfloat16> w(w_elems, __float2bfloat16(0.0f));
    for (int k = 0; k < t.n_steps; ++k)
        for (int hsel = 0; hsel < 2; ++hsel) {
            const Half h = hsel ? steps[k].second : steps[k].first;
            if (h.row < 0) continue;
            const int tap = h.row * p.kw + h.kwi;                   // (kd,kh,kw) flattened, kw fastest
            for (int c = 0; c < p.cin; ++c)
                for (int n = 0; n < p.cout; ++n) {
                    const float v = d.kernel_w[(static_cast<size_t>(tap) * p.cin + c) * p.cout + n];
                    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
                    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
                    const size_t chunk = static_cast<size_t>(2 * k + hsel) * 2 * n_tile;
                    w[(chunk + n) * 8 + c] = hi;
                    w[(chunk + n_tile + n) * 8 + c] = lo;
                }
        }
    TB_CHECK_CUDA(cudaMalloc(&p.d_thin_w, w_elems * sizeof(__nv_bfloat16)));
    TB_CHECK_CUDA(cudaMemcpy(p.d_thin_w, w.data(), w_elems * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
    t.w_packed = p.d_thin_w;
    t.w_bytes = static_cast<uint32_t>(w_elems * 2);
    // ---- geometry
    t.Do = p.Do; t.Ho = p.Ho; t.Wo = p.Wo; t.Wp = tin.pv_Wp;
    t.tiles_per_plane = ceil_div(p.Ho * tin.pv_Wp, 128);
    t.frame_bytes = static_cast<int64_t>(tin.pv_Dp) * tin.pv_Hp * tin.pv_Wp * 16;
    t.dplane_bytes = static_cast<int64_t>(tin.pv_Hp) * tin.pv_Wp * 16;
    t.row_bytes = tin.pv_Wp * 16;
    t.off_d = tin.pv_d0 - p.pad0[0];
    t.off_h = tin.pv_h0 - p.pad0[1];
    t.off_w = tin.pv_w0 - p.pad0[2];
    t.kd = p.kd; t.kh = p.kh; t.kw = p.kw;
    t.span_bytes = span_bytes;
    t.span_stride = span_stride;
    t.n_tile = n_tile;
    t.acc_cols = round_up(2 * n_tile, 32);
    t.acc_stages = std::max(1, std::min(4, 512 / t.acc_cols));
    const size_t w_smem = (w_elems * 2 + 127) & ~static_cast<size_t>(127);
    const size_t stage = 2u * rows * span_stride;
    TB_REQUIRE(w_smem + 2 * stage + 128 <= kSmemDynamicMax, "thin conv: weights + two stages exceed shared memory");
    t.stages = static_cast<int>(std::min<size_t>(kConvMaxStages, (kSmemDynamicMax - 128 - w_smem) / stage));
    if (const char* e = getenv("TIMED_B200_THIN_STAGES")) t.stages = std::max(2, std::min(t.stages, atoi(e)));
    t.act1 = p.act1; t.act2 = p.act2; t.alpha1 = p.alpha1; t.alpha2 = p.alpha2;
    return 0;
}

template <int A1, int A2, int F>
static int launch_thin_instance(const ThinConvParams& k, int grid, size_t smem_bytes, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        TB_CHECK_CUDA(cudaFuncSetAttribute(thin_conv_kernel<A1, A2, F>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(kSmemDynamicMax)));
        attr_set = true;
    }
    thin_conv_kernel<A1, A2, F><<<grid, kConvThreads, smem_bytes, stream>>>(k);
    return 0;
}

// ----------------------------------------------------------------------------- slab conv plan
struct SlabGeom {
    int pd, ph, pw;          // shared zero margins per dimension
    int Dp, Hp, Wp;
    int neg, pos;            // halo positions before / after a tile
    int mt, w_stages, w_group, acc_stages, acc_cols, slab_pix, slab_stride;
    size_t smem_bytes;
};

// Can a stride-1 conv over a (D,H,W,cin) tensor run on slab_conv_kernel?  Fills the geometry when it can.
static bool slab_geometry(int D, int H, int W, int cin, const tb_op_desc& c, SlabGeom* out, int min_pd = 0,
                          int min_ph = 0, int min_pw = 0) {
    if (getenv("TIMED_B200_NO_SLAB")) return false;
    if (c.op != TB_OP_CONV3D || c.stride[0] != 1 || c.stride[1] != 1 || c.stride[2] != 1) return false;
    const int c_pad = round_up(cin, 16), n_tile = round_up(c.c_out, 16);
    if (cin <= 8 || c_pad > 64 || n_tile > 128) return false;
    const int ks[3] = {c.kernel[0], c.kernel[1], c.kernel[2]}, in[3] = {D, H, W};
    int pb[3] = {0, 0, 0}, pa[3] = {0, 0, 0};
    for (int a = 0; a < 3; ++a) {
        if (ks[a] < 1 || ks[a] > 7) return false;
        if (c.pad_same) {
            int o;
            same_pads(in[a], ks[a], 1, &o, &pb[a], &pa[a]);
        } else if (in[a] < ks[a]) {
            return false;
        }
    }
    SlabGeom g;
    g.pd = std::max({pb[0], pa[0], min_pd}); g.ph = std::max({pb[1], pa[1], min_ph}); g.pw = std::max({pb[2], pa[2], min_pw});
    g.Dp = D + g.pd; g.Hp = H + g.ph; g.Wp = W + g.pw;
    // margin positions are computed and dropped: require decent utilisation of the MMA rows
    const double util = static_cast<double>(D) * H * W / (static_cast<double>(g.Dp) * g.Hp * g.Wp);
    if (util < 0.70 && !getenv("TIMED_B200_FORCE_SLAB")) return false;
    g.neg = pb[0] * g.Hp * g.Wp + pb[1] * g.Wp + pb[2];
    g.pos = (ks[0] - 1 - pb[0]) * g.Hp * g.Wp + (ks[1] - 1 - pb[1]) * g.Wp + (ks[2] - 1 - pb[2]);
    const int n_chunks = c_pad / 8;
    g.acc_cols = round_up(2 * n_tile, 32);
    const size_t w_tap = static_cast<size_t>(n_chunks) * 2 * n_tile * 16;
    const size_t w_stage = (w_tap + 127) & ~static_cast<size_t>(127);
    for (int mt : {2, 1}) {
        if (mt * g.acc_cols > 512) continue;
        if (mt == 2 && getenv("TIMED_B200_SLAB_MT1")) continue;
        g.mt = mt;
        g.acc_stages = std::min(2, 512 / (mt * g.acc_cols));
        g.slab_pix = 128 * mt + g.neg + g.pos;
        g.slab_stride = g.slab_pix * 16;
        const size_t slabs = 2u * 2u * n_chunks * g.slab_stride;
        if (slabs + 2 * w_stage + 128 > kSmemDynamicMax) continue;
        // taps per ring stage: one tcgen05.commit per stage (each costs 100-400 tensor cycles), >= 3 stages kept
        const size_t avail = kSmemDynamicMax - 128 - slabs;
        int wg = 2;
        if (const char* e = getenv("TIMED_B200_SLAB_WG")) wg = std::max(1, atoi(e));
        while (wg > 1 && avail / (((wg * w_tap) + 127) & ~static_cast<size_t>(127)) < 3) --wg;
        g.w_group = wg;
        const size_t w_stage_g = ((wg * w_tap) + 127) & ~static_cast<size_t>(127);
        g.w_stages = static_cast<int>(std::min<size_t>(kSlabWStages, avail / w_stage_g));
        g.smem_bytes = 128 + slabs + g.w_stages * w_stage_g;
        *out = g;
        return true;
    }
    return false;
}

static int slab_plan_create(ConvPlan& p, const tb_op_desc& d, const TensorInfo& tin) {
    SlabGeom g;
    TB_REQUIRE(slab_geometry(p.Di, p.Hi, p.Wi, p.cin, d, &g, tin.cpv_Dp - p.Di, tin.cpv_Hp - p.Hi, tin.cpv_Wp - p.Wi),
               "internal: slab conv not applicable");
    TB_REQUIRE(g.Dp == tin.cpv_Dp && g.Hp == tin.cpv_Hp && g.Wp == tin.cpv_Wp, "internal: CPV margins mismatch");
    p.slab = true;
    const int n_tile = round_up(p.cout, 16);
    p.n_tiles = 1;
    p.n_tile = p.n_alloc = n_tile;
    const int n_chunks = p.cin_pad / 8;
    const int taps = p.kd * p.kh * p.kw;
    // halo in the TENSOR's linearisation (its margins may be wider than this conv needs)
    const int Hp = tin.cpv_Hp, Wp = tin.cpv_Wp;
    SlabConvParams& t = p.slab_params;
    std::memset(&t, 0, sizeof(t));
    t.mt = g.mt;
    t.Dp = tin.cpv_Dp; t.Hp = Hp; t.Wp = Wp;
    t.Do = p.Do; t.Ho = p.Ho; t.Wo = p.Wo;
    t.n_chunks = n_chunks;
    t.kd = p.kd; t.kh = p.kh; t.kw = p.kw;
    t.pd = p.pad0[0]; t.ph = p.pad0[1]; t.pw = p.pad0[2];
    t.neg_halo = p.pad0[0] * Hp * Wp + p.pad0[1] * Wp + p.pad0[2];
    t.pos_halo = (p.kd - 1 - p.pad0[0]) * Hp * Wp + (p.kh - 1 - p.pad0[1]) * Wp + (p.kw - 1 - p.pad0[2]);
    t.slab_pix = 128 * g.mt + t.neg_halo + t.pos_halo;
    t.slab_stride = t.slab_pix * 16;
    t.n_tile = n_tile;
    t.acc_cols = g.acc_cols;
    t.acc_stages = g.acc_stages;
    t.w_tap_bytes = static_cast<uint32_t>(n_chunks) * 2u * n_tile * 16u;
    t.w_group = g.w_group;
    t.w_stages = g.w_stages;
    p.slab_lead = tin.cpv_lead;
    p.slab_tail = tin.cpv_tail;
    TB_REQUIRE(tin.cpv_lead >= t.neg_halo && tin.cpv_tail >= t.pos_halo + 128 * g.mt, "internal: CPV lead/tail too small");
    // ---- weights: [tap][chunk][2*n_tile rows: hi then lo][8]
    const size_t w_elems = static_cast<size_t>(taps) * n_chunks * 2 * n_tile * 8;
    std::vector<__nv_bfloat16> w(w_elems, __float2bfloat16(0.0f));
    for (int tap = 0; tap < taps; ++tap)
        for (int c = 0; c < p.cin; ++c)
            for (int n = 0; n < p.cout; ++n) {
                const float v = d.kernel_w[(static_cast<size_t>(tap) * p.cin + c) * p.cout + n];
                const __nv_bfloat16 hi = __float2bfloat16_rn(v);
                const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
                const size_t blk = (static_cast<size_t>(tap) * n_chunks + c / 8) * 2 * n_tile;
                w[(blk + n) * 8 + c % 8] = hi;
                w[(blk + n_tile + n) * 8 + c % 8] = lo;
            }
    TB_CHECK_CUDA(cudaMalloc(&p.d_slab_w, w_elems * sizeof(__nv_bfloat16)));
    TB_CHECK_CUDA(cudaMemcpy(p.d_slab_w, w.data(), w_elems * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
    t.w_packed = p.d_slab_w;
    return 0;
}

template <int A1, int A2, int F>
static int launch_slab_instance(const SlabConvParams& k, int grid, size_t smem_bytes, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        TB_CHECK_CUDA(cudaFuncSetAttribute(slab_conv_kernel<A1, A2, F>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(kSmemDynamicMax)));
        attr_set = true;
    }
    slab_conv_kernel<A1, A2, F><<<grid, kSlabThreads, smem_bytes, stream>>>(k);
    return 0;
}

static int slab_launch(ConvPlan& p, void* in_base, int64_t n_frames, const TView& out, cudaStream_t stream) {
    SlabConvParams k = p.slab_params;
    const int64_t fpos = static_cast<int64_t>(k.Dp) * k.Hp * k.Wp;
    const int64_t T = p.slab_lead + n_frames * fpos + p.slab_tail;
    k.t_first = p.slab_lead;
    k.t_count = n_frames * fpos;
    const int64_t tiles = (k.t_count + 128 * k.mt - 1) / (128 * k.mt);
    TB_REQUIRE(tiles > 0 && tiles < (1ll << 31), "slab conv: too many tiles per launch");
    k.n_tiles_total = static_cast<int32_t>(tiles);
    k.in_hi = static_cast<const uint8_t*>(in_base);
    k.chunk_stride = T * 16;
    k.lo_plane_off = static_cast<int64_t>(k.n_chunks) * T * 16;
    ConvKernelParams& e = k.epi;
    e.bias = p.d_bias; e.scale = p.d_scale; e.shift = p.d_shift;
    e.act1 = p.act1; e.act2 = p.act2; e.alpha1 = p.alpha1; e.alpha2 = p.alpha2;
    e.out_fmt = out.fmt;
    e.out_f32 = out.f32; e.out_hi = out.hi; e.out_lo = out.lo;
    e.ldc = out.ld;
    e.c_store = out.fmt == FMT_SPLIT ? out.c_pad : out.c;
    k.dbg = debug_mask();
    e.acc_comp = accum_comp(static_cast<double>(p.kd) * p.kh * p.kw * ceil_div(p.cin, 16),
                            valid_tap_fraction(p.Di, p.Do, p.kd, p.pad0[0]) * valid_tap_fraction(p.Hi, p.Ho, p.kh, p.pad0[1]) *
                                valid_tap_fraction(p.Wi, p.Wo, p.kw, p.pad0[2]));
    TB_REQUIRE(out.fmt != FMT_SPLIT || (out.c_pad % 16 == 0 && out.c_pad <= p.n_alloc),
               "slab conv: split output channel padding mismatch");
    const size_t w_stage = (static_cast<size_t>(k.w_group) * k.w_tap_bytes + 127) & ~static_cast<size_t>(127);
    const size_t smem_bytes = 128 + 2u * 2u * k.n_chunks * static_cast<size_t>(k.slab_stride) + k.w_stages * w_stage;
    const int grid = static_cast<int>(std::min<int64_t>(tiles, 148));
    int rc = 0;
    bool launched = false;
#define TB_SLAB_CASE(A1, A2, F)                                                              \
    if (!launched && e.act1 == (A1) && e.act2 == (A2) && out.fmt == (F)) {                   \
        rc = launch_slab_instance<A1, A2, F>(k, grid, smem_bytes, stream);                   \
        launched = true;                                                                     \
    }
    TB_SLAB_CASE(ACT_ELU, ACT_NONE, FMT_F32)
    TB_SLAB_CASE(ACT_ELU, ACT_NONE, FMT_SPLIT)
    TB_SLAB_CASE(ACT_RELU, ACT_NONE, FMT_F32)
    TB_SLAB_CASE(ACT_RELU, ACT_NONE, FMT_SPLIT)
    TB_SLAB_CASE(ACT_NONE, ACT_NONE, FMT_F32)
    TB_SLAB_CASE(ACT_NONE, ACT_NONE, FMT_SPLIT)
#undef TB_SLAB_CASE
    if (!launched)
        rc = out.fmt == FMT_SPLIT ? launch_slab_instance<-1, -1, FMT_SPLIT>(k, grid, smem_bytes, stream)
                                  : launch_slab_instance<-1, -1, FMT_F32>(k, grid, smem_bytes, stream);
    if (rc) return rc;
    TB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ----------------------------------------------------------------------------- thin-input conv, kd folded into N
struct ThinZGeom { int zt, n_steps, b1_rows, b2_rows, span_bytes, span_stride, stages, acc_cols, acc_stages; size_t w_bytes; };

static bool thinz_geometry(int kd, int kh, int kw, int cout, int Wp, ThinZGeom* out, bool pool = false) {
    if (getenv("TIMED_B200_NO_ZFOLD")) return false;
    const int n_tile = round_up(cout, 16);
    if (kd < 2 || kd * 2 * n_tile > 256) return false;
    ThinZGeom g;
    g.zt = std::max(1, 256 / (2 * n_tile));                       // two accumulator stages of zt*2*n_tile columns
    g.acc_cols = round_up(g.zt * 2 * n_tile, 32);
    g.acc_stages = std::min(2, 512 / g.acc_cols);
    g.n_steps = kh * (kw / 2) + ((kw & 1) ? (kh + 1) / 2 : 0);
    g.b1_rows = kd * 2 * n_tile;
    g.b2_rows = (2 * kd - 1) * n_tile;
    if (pool) {
        // fused max-pool instantiation: corrections accumulate in the MAIN columns (three MMAs of N = cnt*n_tile per step
        // instead of N-folded two; a conv this thin has <= ~45 MMAs per output, so the truncation they add is < 1e-6): the
        // epilogue -- which paces this kernel -- reads half the TMEM, and the accumulators of a tile take half the columns,
        // so there are four stages instead of two between the MMA issuers and the epilogue
        g.acc_cols = round_up(g.zt * n_tile, 32);
        const int st = 512 / g.acc_cols;
        g.acc_stages = st >= 4 ? 4 : st >= 2 ? 2 : 1;
        g.b1_rows = g.b2_rows = kd * n_tile;                      // [W_hi slices] and [W_lo slices]
    }
    g.w_bytes = static_cast<size_t>(g.n_steps) * 2 * 16 * (g.b1_rows + g.b2_rows);
    const int span_pix = (128 + (kh - 1) * Wp + kw + 1 + 1) & ~1;   // window + filter extent + the aliased next pixel
    g.span_bytes = span_pix * 16;
    g.span_stride = g.span_bytes;
    const size_t w_smem = (g.w_bytes + 127) & ~static_cast<size_t>(127);
    const size_t stage = 2u * (g.zt + kd - 1) * g.span_stride;
    // fused max-pool: per z pair of a tile a ring of two windows (256 positions) of raw fp32 sums
    const size_t pool_buf = pool ? static_cast<size_t>(std::max(1, g.zt / 2)) * 256u * n_tile * sizeof(float) : 0;
    if (w_smem + 2 * stage + pool_buf + 128 > kSmemDynamicMax) return false;
    if (pool && (g.zt & 1)) return false;                                   // z pairs must not straddle tiles
    if (pool && Wp + 2 > 128) return false;                                 // a 2x2 partner lies at most one window ahead
    g.stages = static_cast<int>(std::min<size_t>(kConvMaxStages, (kSmemDynamicMax - 128 - w_smem - pool_buf) / stage));
    *out = g;
    return true;
}

static int thinz_plan_create(ConvPlan& p, const tb_op_desc& d, const TensorInfo& tin) {
    ThinZGeom g;
    TB_REQUIRE(thinz_geometry(p.kd, p.kh, p.kw, p.cout, tin.pv_Wp, &g), "internal: thinz conv not applicable");
    p.thin = true;
    p.thinz = true;
    const int n_tile = round_up(p.cout, 16);
    p.n_tiles = 1;
    p.n_tile = p.n_alloc = n_tile;
    ThinZParams& t = p.thinz_params;
    std::memset(&t, 0, sizeof(t));
    // ---- K steps of one input plane: (kh row, kw pair), then odd kw taps paired across rows
    struct Half { int row, kwi; };
    std::vector<std::pair<Half, Half>> steps;
    for (int r = 0; r < p.kh; ++r)
        for (int j = 0; j + 1 < p.kw; j += 2) steps.push_back({{r, j}, {r, j + 1}});
    if (p.kw & 1)
        for (int r = 0; r < p.kh; r += 2) steps.push_back({{r, p.kw - 1}, {r + 1 < p.kh ? r + 1 : -1, p.kw - 1}});
    TB_REQUIRE(static_cast<int>(steps.size()) == g.n_steps, "internal: thinz step count");
    // ---- weights: per step [B1: 2 K-chunks][b1_rows][8] then [B2: 2 K-chunks][b2_rows][8]
    const size_t step_elems = static_cast<size_t>(2) * 8 * (g.b1_rows + g.b2_rows);
    std::vector<__nv_bfloat16> w(step_elems * g.n_steps, __float2bfloat16(0.0f));
    for (int sidx = 0; sidx < g.n_steps; ++sidx)
        for (int hsel = 0; hsel < 2; ++hsel) {
            const Half h = hsel ? steps[sidx].second : steps[sidx].first;
            if (h.row < 0) continue;
            __nv_bfloat16* b1 = w.data() + sidx * step_elems + static_cast<size_t>(hsel) * g.b1_rows * 8;
            __nv_bfloat16* b2 = w.data() + sidx * step_elems + static_cast<size_t>(2) * g.b1_rows * 8 +
                                static_cast<size_t>(hsel) * g.b2_rows * 8;
            for (int blk = 0; blk < p.kd; ++blk) {                 // block blk holds filter slice kd' = kd-1-blk
                const int tap = ((p.kd - 1 - blk) * p.kh + h.row) * p.kw + h.kwi;
                for (int c = 0; c < p.cin; ++c)
                    for (int n = 0; n < p.cout; ++n) {
                        const float v = d.kernel_w[(static_cast<size_t>(tap) * p.cin + c) * p.cout + n];
                        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
                        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
                        b1[(static_cast<size_t>(blk) * 2 * n_tile + n) * 8 + c] = hi;
                        b1[(static_cast<size_t>(blk) * 2 * n_tile + n_tile + n) * 8 + c] = lo;
                        b2[(static_cast<size_t>(blk) * 2 * n_tile + n) * 8 + c] = hi;
                    }
            }
        }
    // the fused max-pool instantiation's layout (thinz_geometry(pool)): per step [W_hi: 2 K-chunks][kd*n_tile rows][8] then
    // [W_lo: the same]; appended to the same allocation, selected at launch
    const size_t m_rows = static_cast<size_t>(p.kd) * n_tile;
    const size_t m_step = 2 * 2 * m_rows * 8;
    const size_t m_off = (w.size() + 63) & ~static_cast<size_t>(63);           // 128-byte aligned
    w.resize(m_off + m_step * g.n_steps, __float2bfloat16(0.0f));
    for (int sidx = 0; sidx < g.n_steps; ++sidx)
        for (int hsel = 0; hsel < 2; ++hsel) {
            const Half h = hsel ? steps[sidx].second : steps[sidx].first;
            if (h.row < 0) continue;
            __nv_bfloat16* bh = w.data() + m_off + sidx * m_step + static_cast<size_t>(hsel) * m_rows * 8;
            __nv_bfloat16* bl = bh + 2 * m_rows * 8;
            for (int blk = 0; blk < p.kd; ++blk) {
                const int tap = ((p.kd - 1 - blk) * p.kh + h.row) * p.kw + h.kwi;
                for (int c = 0; c < p.cin; ++c)
                    for (int n = 0; n < p.cout; ++n) {
                        const float v = d.kernel_w[(static_cast<size_t>(tap) * p.cin + c) * p.cout + n];
                        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
                        bh[(static_cast<size_t>(blk) * n_tile + n) * 8 + c] = hi;
                        bl[(static_cast<size_t>(blk) * n_tile + n) * 8 + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
                    }
            }
        }
    TB_CHECK_CUDA(cudaMalloc(&p.d_thin_w, w.size() * sizeof(__nv_bfloat16)));
    TB_CHECK_CUDA(cudaMemcpy(p.d_thin_w, w.data(), w.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
    t.w_packed = p.d_thin_w;
    t.w_bytes = static_cast<uint32_t>(m_off * 2);
    p.thinz_merged_off = m_off * 2;
    p.thinz_merged_bytes = m_step * g.n_steps * 2;
    t.n_steps = g.n_steps;
    t.b1_rows = g.b1_rows; t.b2_rows = g.b2_rows;
    t.zt = g.zt;
    t.z_groups = ceil_div(p.Do, g.zt);
    t.windows = ceil_div(p.Ho * tin.pv_Wp, 128);
    t.win_stride = 128;
    t.Do = p.Do; t.Ho = p.Ho; t.Wo = p.Wo; t.Wp = tin.pv_Wp;
    t.frame_bytes = static_cast<int64_t>(tin.pv_Dp) * tin.pv_Hp * tin.pv_Wp * 16;
    t.dplane_bytes = static_cast<int64_t>(tin.pv_Hp) * tin.pv_Wp * 16;
    t.off_d = tin.pv_d0 - p.pad0[0];
    t.off_hw = (tin.pv_h0 - p.pad0[1]) * tin.pv_Wp + (tin.pv_w0 - p.pad0[2]);
    t.kd = p.kd; t.kh = p.kh; t.kw = p.kw;
    t.span_bytes = g.span_bytes;
    t.span_stride = g.span_stride;
    t.n_tile = n_tile;
    t.acc_cols = g.acc_cols;
    t.acc_stages = g.acc_stages;
    t.stages = g.stages;
    t.issuers = getenv("TIMED_B200_THINZ_ISSUERS") ? std::max(1, std::min(2, atoi(getenv("TIMED_B200_THINZ_ISSUERS")))) : 2;
    return 0;
}

template <int A1, int A2, int F, int POOL>
static int launch_thinz_instance(const ThinZParams& k, int grid, size_t smem_bytes, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        TB_CHECK_CUDA(cudaFuncSetAttribute(thinz_conv_kernel<A1, A2, F, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(kSmemDynamicMax)));
        attr_set = true;
    }
    thinz_conv_kernel<A1, A2, F, POOL><<<grid, kThinzThreads, smem_bytes, stream>>>(k);
    return 0;
}

static int thinz_launch(ConvPlan& p, void* in_base, int64_t in_frames_alloc, int64_t n_frames, const TView& out,
                        cudaStream_t stream, const TensorInfo* out_info = nullptr) {
    ThinZParams k = p.thinz_params;
    if (p.fuse_zpool) {
        k.pool_same = p.pool_same;
        k.Zo = p.pool_Zo;
    }
    if (p.fuse_pool) {
        TB_REQUIRE(out_info != nullptr, "internal: fused pool needs the output tensor description");
        ThinZGeom zg;
        TB_REQUIRE(thinz_geometry(k.kd, k.kh, k.kw, p.cout, k.Wp, &zg, true), "internal: fused pool does not fit");
        k.stages = zg.stages;
        k.acc_cols = zg.acc_cols;
        k.acc_stages = zg.acc_stages;
        k.b1_rows = zg.b1_rows; k.b2_rows = zg.b2_rows;
        k.w_packed = p.d_thin_w + p.thinz_merged_off;
        k.w_bytes = static_cast<uint32_t>(p.thinz_merged_bytes);
        TB_REQUIRE(zg.w_bytes == p.thinz_merged_bytes, "internal: thinz merged weight size");
        k.pool_same = p.pool_same;
        k.Zo = p.pool_Zo; k.Po = p.pool_Po; k.Qo = p.pool_Qo;
        k.pool_cpv = out_info->cpv ? 1 : 0;
        if (out_info->cpv) {
            k.cpv_T = out_info->cpv_T(n_frames);
            k.cpv_lead = out_info->cpv_lead;
            k.cpv_Dp = out_info->cpv_Dp; k.cpv_Hp = out_info->cpv_Hp; k.cpv_Wp = out_info->cpv_Wp;
            k.out_hi4 = reinterpret_cast<uint4*>(out.hi);
            k.out_lo4 = reinterpret_cast<uint4*>(out.lo);
        }
        k.pool_out = out;
    }
    const int64_t tiles = n_frames * k.z_groups * k.windows;
    TB_REQUIRE(tiles > 0 && tiles < (1ll << 31), "thinz conv: too many tiles per launch");
    k.n_tiles_total = static_cast<int32_t>(tiles);
    k.in_hi = static_cast<const uint8_t*>(in_base);
    k.lo_plane_off = in_frames_alloc * k.frame_bytes;
    ConvKernelParams& e = k.epi;
    e.bias = p.d_bias; e.scale = p.d_scale; e.shift = p.d_shift;
    e.act1 = p.act1; e.act2 = p.act2; e.alpha1 = p.alpha1; e.alpha2 = p.alpha2;
    e.out_fmt = out.fmt;
    e.out_f32 = out.f32; e.out_hi = out.hi; e.out_lo = out.lo;
    e.ldc = out.ld;
    e.c_store = out.fmt == FMT_SPLIT ? out.c_pad : out.c;
    e.acc_comp = 1.0f;                     // <= 42 MMAs per accumulator: the shrink is below 1e-6
    if (p.fuse_pool)                       // corrections share the main columns: 3 MMAs per step and filter slice
        e.acc_comp = accum_comp(3.0 * p.kd * k.n_steps,
                                valid_tap_fraction(p.Di, p.Do, p.kd, p.pad0[0]) * valid_tap_fraction(p.Hi, p.Ho, p.kh, p.pad0[1]) *
                                    valid_tap_fraction(p.Wi, p.Wo, p.kw, p.pad0[2]));
    k.dbg = debug_mask();
    TB_REQUIRE(out.fmt != FMT_SPLIT || (out.c_pad % 16 == 0 && out.c_pad <= p.n_alloc),
               "thinz conv: split output channel padding mismatch");
    const size_t w_smem = (static_cast<size_t>(k.w_bytes) + 127) & ~static_cast<size_t>(127);
    const size_t smem_bytes = 128 + w_smem + static_cast<size_t>(k.stages) * 2u * (k.zt + k.kd - 1) * k.span_stride +
                              (p.fuse_pool ? static_cast<size_t>(std::max(1, k.zt / 2)) * 256u * k.n_tile * sizeof(float) : 0);
    const int grid = static_cast<int>(std::min<int64_t>(tiles, 148));
    int rc = 0;
    bool launched = false;
#define TB_THINZ_CASE(A1, A2, F)                                                             \
    if (!launched && e.act1 == (A1) && e.act2 == (A2) && out.fmt == (F)) {                   \
        rc = p.fuse_pool    ? launch_thinz_instance<A1, A2, F, 1>(k, grid, smem_bytes, stream)  \
             : p.fuse_zpool ? launch_thinz_instance<A1, A2, F, 2>(k, grid, smem_bytes, stream)  \
                            : launch_thinz_instance<A1, A2, F, 0>(k, grid, smem_bytes, stream); \
        launched = true;                                                                     \
    }
    TB_THINZ_CASE(ACT_ELU, ACT_NONE, FMT_F32)
    TB_THINZ_CASE(ACT_ELU, ACT_NONE, FMT_SPLIT)
    TB_THINZ_CASE(ACT_RELU, ACT_NONE, FMT_F32)
    TB_THINZ_CASE(ACT_RELU, ACT_NONE, FMT_SPLIT)
    TB_THINZ_CASE(ACT_NONE, ACT_NONE, FMT_F32)
    TB_THINZ_CASE(ACT_NONE, ACT_NONE, FMT_SPLIT)
#undef TB_THINZ_CASE
    if (!launched && p.fuse_pool)
        rc = out.fmt == FMT_SPLIT ? launch_thinz_instance<-1, -1, FMT_SPLIT, 1>(k, grid, smem_bytes, stream)
                                  : launch_thinz_instance<-1, -1, FMT_F32, 1>(k, grid, smem_bytes, stream);
    else if (!launched && p.fuse_zpool)
        rc = out.fmt == FMT_SPLIT ? launch_thinz_instance<-1, -1, FMT_SPLIT, 2>(k, grid, smem_bytes, stream)
                                  : launch_thinz_instance<-1, -1, FMT_F32, 2>(k, grid, smem_bytes, stream);
    else if (!launched)
        rc = out.fmt == FMT_SPLIT ? launch_thinz_instance<-1, -1, FMT_SPLIT, 0>(k, grid, smem_bytes, stream)
                                  : launch_thinz_instance<-1, -1, FMT_F32, 0>(k, grid, smem_bytes, stream);
    if (rc) return rc;
    TB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int thin_launch(ConvPlan& p, void* in_base, int64_t in_frames_alloc, int64_t n_frames, const TView& out,
                       cudaStream_t stream) {
    ThinConvParams k = p.thin_params;
    const int64_t tiles = n_frames * k.Do * k.tiles_per_plane;
    TB_REQUIRE(tiles > 0 && tiles < (1ll << 31), "thin conv: too many tiles per launch");
    k.n_tiles_total = static_cast<int32_t>(tiles);
    k.in_hi = static_cast<const uint8_t*>(in_base);
    k.lo_plane_off = in_frames_alloc * k.frame_bytes;
    k.bias = p.d_bias; k.scale = p.d_scale; k.shift = p.d_shift;
    k.out_f32 = out.f32; k.out_hi = out.hi; k.out_lo = out.lo;
    k.ldc = out.ld;
    k.c_store = out.fmt == FMT_SPLIT ? out.c_pad : out.c;
    k.dbg = debug_mask();
    TB_REQUIRE(out.fmt != FMT_SPLIT || (out.c_pad % 16 == 0 && out.c_pad <= p.n_alloc),
               "thin conv: split output channel padding mismatch");
    const size_t w_smem = (static_cast<size_t>(k.w_bytes) + 127) & ~static_cast<size_t>(127);
    const size_t smem_bytes = 128 + w_smem + static_cast<size_t>(k.stages) * 2u * (k.kd * k.kh) * k.span_stride;
    const int grid = static_cast<int>(std::min<int64_t>(tiles, 148));
    int rc = 0;
    bool launched = false;
#define TB_THIN_CASE(A1, A2, F)                                                              \
    if (!launched && k.act1 == (A1) && k.act2 == (A2) && out.fmt == (F)) {                   \
        rc = launch_thin_instance<A1, A2, F>(k, grid, smem_bytes, stream);                   \
        launched = true;                                                                     \
    }
    TB_THIN_CASE(ACT_ELU, ACT_NONE, FMT_F32)
    TB_THIN_CASE(ACT_ELU, ACT_NONE, FMT_SPLIT)
    TB_THIN_CASE(ACT_RELU, ACT_NONE, FMT_F32)
    TB_THIN_CASE(ACT_RELU, ACT_NONE, FMT_SPLIT)
    TB_THIN_CASE(ACT_NONE, ACT_NONE, FMT_F32)
    TB_THIN_CASE(ACT_NONE, ACT_NONE, FMT_SPLIT)
#undef TB_THIN_CASE
    if (!launched)
        rc = out.fmt == FMT_SPLIT ? launch_thin_instance<-1, -1, FMT_SPLIT>(k, grid, smem_bytes, stream)
                                  : launch_thin_instance<-1, -1, FMT_F32>(k, grid, smem_bytes, stream);
    if (rc) return rc;
    TB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Build the static part of a conv plan: geometry, packed weights, epilogue vectors.
static int conv_plan_create(ConvPlan& p, const tb_op_desc& d, int Di, int Hi, int Wi, int cin,
                            int cin_pad, const TensorInfo* wfold_in = nullptr) {
    p.kd = d.kernel[0]; p.kh = d.kernel[1]; p.kw = d.kernel[2];
    TB_REQUIRE(p.kd >= 1 && p.kh >= 1 && p.kw >= 1 && p.kd <= 16 && p.kh <= 16 && p.kw <= 16,
               "conv: kernel extents must be in [1,16]");
    TB_REQUIRE(d.stride[0] == 1 && d.stride[1] == 1 && d.stride[2] == 1, "conv: only stride 1");
    TB_REQUIRE(d.kernel_w != nullptr, "conv: kernel weights missing");
    TB_REQUIRE(d.c_out >= 1, "conv: c_out must be positive");
    p.Di = Di; p.Hi = Hi; p.Wi = Wi; p.cin = cin; p.cin_pad = cin_pad; p.cout = d.c_out;
    if (wfold_in && wfold_in->wfold) {
        p.wfold = true;
        p.kwin = p.kw <= 4 ? 4 : 8;
        p.in_lm = wfold_in->wf_lm;
        p.in_pitch = wfold_in->wf_pitch;
        p.cin_pad = cin_pad = p.kwin * 8;
    }
    const int ks[3] = {p.kd, p.kh, p.kw}, in[3] = {Di, Hi, Wi};
    int out[3];
    for (int i = 0; i < 3; ++i) {
        if (d.pad_same) {
            same_pads(in[i], ks[i], 1, &out[i], &p.pad0[i], &p.pad1[i]);
        } else {
            out[i] = in[i] - ks[i] + 1;
            p.pad0[i] = p.pad1[i] = 0;
        }
        TB_REQUIRE(out[i] >= 1, "conv: kernel larger than input with 'valid' padding");
    }
    p.Do = out[0]; p.Ho = out[1]; p.Wo = out[2];
    if (wfold_in && (wfold_in->padvol || wfold_in->cpv)) {
        p.act1 = d.act1; p.act2 = d.act2; p.alpha1 = d.alpha1; p.alpha2 = d.alpha2;
        ThinZGeom zg;
        int rc = wfold_in->cpv ? slab_plan_create(p, d, *wfold_in)
                 : thinz_geometry(p.kd, p.kh, p.kw, p.cout, wfold_in->pv_Wp, &zg) ? thinz_plan_create(p, d, *wfold_in)
                                                                                 : thin_plan_create(p, d, *wfold_in);
        if (rc) return rc;
        std::vector<float> b(p.n_alloc, 0.f), sc(p.n_alloc, 1.f), sh(p.n_alloc, 0.f);
        for (int n = 0; n < p.cout; ++n) {
            if (d.bias) b[n] = d.bias[n];
            if (d.scale) sc[n] = d.scale[n];
            if (d.shift) sh[n] = d.shift[n];
        }
        TB_CHECK_CUDA(cudaMalloc(&p.d_bias, p.n_alloc * sizeof(float)));
        TB_CHECK_CUDA(cudaMalloc(&p.d_scale, p.n_alloc * sizeof(float)));
        TB_CHECK_CUDA(cudaMalloc(&p.d_shift, p.n_alloc * sizeof(float)));
        TB_CHECK_CUDA(cudaMemcpy(p.d_bias, b.data(), p.n_alloc * sizeof(float), cudaMemcpyHostToDevice));
        TB_CHECK_CUDA(cudaMemcpy(p.d_scale, sc.data(), p.n_alloc * sizeof(float), cudaMemcpyHostToDevice));
        TB_CHECK_CUDA(cudaMemcpy(p.d_shift, sh.data(), p.n_alloc * sizeof(float), cudaMemcpyHostToDevice));
        return 0;
    }
    {
        const int taps_all = p.kd * p.kh * p.kw;
        // cost model (cycles per 128-row tile, from profiles/r1_summary.md): a thin-N MMA costs ~65
        // cycles whatever N is, a wide one ~110; the Z matrix is written and read once through HBM at
        // ~23 B/cycle/SM.  Take tap-to-N only when it wins clearly (TIMED's head: 68k vs 168k cycles;
        // DenseCPD's 128->32 growth convs: 68k vs 42k, so they stay on the direct path).
        // cost model, cycles per 128-row tile (tools/mma_probe.cu: an M = 128 MMA costs ~72 cycles up to N = 64, 135 at
        // N = 256; N-folded issue = MMAs of 2N and N per K step for N <= 128, three of N otherwise; the Z matrix is written
        // and read once through HBM at ~23 B/cycle/SM).  TIMED's 512 -> 20 head: direct 124k, (kd,kh)-in-N 45k, kw-in-N 52k;
        // DenseCPD's 128 -> 32 growth convs: direct 31k, (kd,kh)-in-N 33k, kw-in-N 20k.
        const double k16 = cin_pad / 16.0;
        auto mma = [](double n) { return n <= 64 ? 72.0 : 72.0 + (n - 64.0) * (135.0 - 72.0) / 192.0; };
        auto gemm = [&](double cols, double k16_total) {
            const int tiles = ceil_div(round_up(static_cast<int>(cols), 16), 256);
            const double nt = round_up(ceil_div(round_up(static_cast<int>(cols), 16), tiles), 16);
            return k16_total * tiles * (nt <= 128 ? mma(2 * nt) + mma(nt) : 3 * mma(nt));
        };
        auto zcost = [](double cols) { return 8.0 * cols * 128 / 23.0 * 1.5; };
        const double direct_cost = gemm(p.cout, taps_all * k16);
        const double full_cost = gemm(static_cast<double>(taps_all) * p.cout, k16) + zcost(static_cast<double>(taps_all) * p.cout);
        const double kwk_cost = p.kw > 1 ? gemm(static_cast<double>(p.kd) * p.kh * p.cout, p.kw * k16) + zcost(static_cast<double>(p.kd) * p.kh * p.cout) : 1e30;
        const double wn_cost = (p.kd * p.kh > 1 && p.kw > 1) ? gemm(static_cast<double>(p.kw) * p.cout, p.kd * p.kh * k16) + zcost(static_cast<double>(p.kw) * p.cout) : 1e30;
        int variant = 0;                                     // 0 all taps in N, 1 kw in K, 2 kw in N
        double t2n_cost = full_cost;
        if (getenv("TIMED_B200_TAP2N_FULL")) variant = 0;
        else if (getenv("TIMED_B200_TAP2N_W") && wn_cost < 1e29) { variant = 2; t2n_cost = wn_cost; }
        else {
            if (kwk_cost < t2n_cost) { variant = 1; t2n_cost = kwk_cost; }
            if (wn_cost < t2n_cost && !getenv("TIMED_B200_NO_TAP2N_W")) { variant = 2; t2n_cost = wn_cost; }
        }
        const double margin = getenv("TIMED_B200_TAP2N_MARGIN") ? atof(getenv("TIMED_B200_TAP2N_MARGIN")) : 0.7;
        const int zc = (variant == 1 ? p.kd * p.kh : variant == 2 ? p.kw : taps_all) * p.cout;
        p.tap2n = !p.wfold && taps_all > 1 && p.cout <= 64 && zc <= 1024 && cin_pad >= 64 &&
                  t2n_cost < margin * direct_cost && !getenv("TIMED_B200_NO_TAP2N");
        if (p.tap2n) {
            p.t2n_kw = variant == 1;
            p.t2n_w = variant == 2;
            p.z_cols = zc;
            p.z_ld = round_up(p.z_cols, 4);
            if (p.t2n_w && p.kw == 3 && p.pad0[2] == 1 && p.Wo == p.Wi && (p.cout == 16 || p.cout == 32) && p.Wo <= 128 &&
                !getenv("TIMED_B200_NO_C2I_FUSE")) {
                const int rows = (128 / p.Wo) * p.Wo;        // whole rows of the volume per 128-row tile
                if (rows >= 104) { p.c2i = true; p.c2i_rows = rows; }
            }
        }
    }
    const int n_pad = round_up(p.gemm_n(), 16);
    p.n_tiles = ceil_div(n_pad, 256);
    p.n_tile = round_up(ceil_div(n_pad, p.n_tiles), 16);
    p.n_alloc = p.n_tile * p.n_tiles;
    p.k_total = p.taps_eff() * cin_pad;
    p.act1 = d.act1; p.act2 = d.act2; p.alpha1 = d.alpha1; p.alpha2 = d.alpha2;

    // ---- pack weights: Keras DHWIO fp32 -> [plane][n][tap*cin_pad + c] bf16 hi/lo
    const size_t plane = static_cast<size_t>(p.n_alloc) * p.k_total;
    std::vector<__nv_bfloat16> w(2 * plane, __float2bfloat16(0.0f));
    const int taps = p.kd * p.kh * p.kw;
    for (int t = 0; t < taps; ++t)
        for (int c = 0; c < cin; ++c) {
            const float* src = d.kernel_w + (static_cast<size_t>(t) * cin + c) * p.cout;
            for (int n = 0; n < p.cout; ++n) {
                const float v = src[n];
                const __nv_bfloat16 hi = __float2bfloat16_rn(v);
                const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
                // K index: dense = tap*cin_pad + c; W-folded = (kd,kh)-tap * (kwin*8) + kw*8 + c
                const size_t kidx = p.tap2n ? (p.t2n_kw ? static_cast<size_t>(t % p.kw) * cin_pad + c
                                               : p.t2n_w ? static_cast<size_t>(t / p.kw) * cin_pad + c : static_cast<size_t>(c))
                                  : p.wfold ? static_cast<size_t>(t / p.kw) * cin_pad + static_cast<size_t>(t % p.kw) * 8 + c
                                            : static_cast<size_t>(t) * cin_pad + c;
                // GEMM column: the output channel, or (tap, channel) for tap-to-N
                const size_t ncol = p.tap2n ? static_cast<size_t>(p.t2n_kw ? t / p.kw : p.t2n_w ? t % p.kw : t) * p.cout + n
                                            : static_cast<size_t>(n);
                const size_t o = ncol * p.k_total + kidx;
                w[o] = hi;
                w[plane + o] = lo;
            }
        }
    TB_CHECK_CUDA(cudaMalloc(&p.d_w, 2 * plane * sizeof(__nv_bfloat16)));
    TB_CHECK_CUDA(cudaMemcpy(p.d_w, w.data(), 2 * plane * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
    std::vector<float> b(p.n_alloc, 0.f), sc(p.n_alloc, 1.f), sh(p.n_alloc, 0.f);
    if (p.tap2n) {      // the GEMM epilogue is the identity; the col2im kernel applies bias/act/BN
        // n_alloc entries (identity past cout): the fused-col2im GEMM epilogue stages n_alloc of them
        const int nv = std::max(p.cout, p.n_alloc);
        std::vector<float> cb(nv, 0.f), cs(nv, 1.f), ch(nv, 0.f);
        for (int n = 0; n < p.cout; ++n) {
            if (d.bias) cb[n] = d.bias[n];
            if (d.scale) cs[n] = d.scale[n];
            if (d.shift) ch[n] = d.shift[n];
        }
        TB_CHECK_CUDA(cudaMalloc(&p.d_c2i_bias, nv * sizeof(float)));
        TB_CHECK_CUDA(cudaMalloc(&p.d_c2i_scale, nv * sizeof(float)));
        TB_CHECK_CUDA(cudaMalloc(&p.d_c2i_shift, nv * sizeof(float)));
        TB_CHECK_CUDA(cudaMemcpy(p.d_c2i_bias, cb.data(), nv * sizeof(float), cudaMemcpyHostToDevice));
        TB_CHECK_CUDA(cudaMemcpy(p.d_c2i_scale, cs.data(), nv * sizeof(float), cudaMemcpyHostToDevice));
        TB_CHECK_CUDA(cudaMemcpy(p.d_c2i_shift, ch.data(), nv * sizeof(float), cudaMemcpyHostToDevice));
    } else
    for (int n = 0; n < p.cout; ++n) {
        if (d.bias) b[n] = d.bias[n];
        if (d.scale) sc[n] = d.scale[n];
        if (d.shift) sh[n] = d.shift[n];
    }
    // padded output channels must come out as exact zeros: act(0)*1 + 0
    TB_CHECK_CUDA(cudaMalloc(&p.d_bias, p.n_alloc * sizeof(float)));
    TB_CHECK_CUDA(cudaMalloc(&p.d_scale, p.n_alloc * sizeof(float)));
    TB_CHECK_CUDA(cudaMalloc(&p.d_shift, p.n_alloc * sizeof(float)));
    TB_CHECK_CUDA(cudaMemcpy(p.d_bias, b.data(), p.n_alloc * sizeof(float), cudaMemcpyHostToDevice));
    TB_CHECK_CUDA(cudaMemcpy(p.d_scale, sc.data(), p.n_alloc * sizeof(float), cudaMemcpyHostToDevice));
    TB_CHECK_CUDA(cudaMemcpy(p.d_shift, sh.data(), p.n_alloc * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

template <int A1, int A2, int F>
static int launch_conv_instance(const CUtensorMap& map_a, const CUtensorMap& map_w, const CUtensorMap& map_v,
                                const ConvKernelParams& k,
                                int grid, size_t smem_bytes, cudaStream_t stream) {
    static bool attr_set = false;       // per instantiation
    if (!attr_set) {
        TB_CHECK_CUDA(cudaFuncSetAttribute(conv_umma_kernel<A1, A2, F>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(kUmmaSmemDynamicMax)));
        attr_set = true;
    }
    if (k.cluster2) {
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3(grid);
        lc.blockDim = dim3(kUmmaThreads);
        lc.dynamicSmemBytes = smem_bytes;
        lc.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        lc.attrs = attr;
        lc.numAttrs = 1;
        TB_CHECK_CUDA(cudaLaunchKernelEx(&lc, conv_umma_kernel<A1, A2, F>, map_a, map_w, map_v, k));
        return 0;
    }
    conv_umma_kernel<A1, A2, F><<<grid, kUmmaThreads, smem_bytes, stream>>>(map_a, map_w, map_v, k);
    return 0;
}

template <int A1, int A2, int F>
static int launch_pair_instance(const CUtensorMap& map_a, const CUtensorMap& map_w, const CUtensorMap& map_v,
                                const ConvKernelParams& k, int grid, size_t smem_bytes, cudaStream_t stream) {
    static bool attr_set = false;       // per instantiation
    if (!attr_set) {
        TB_CHECK_CUDA(cudaFuncSetAttribute(conv_pair_kernel<A1, A2, F>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(kSmemDynamicMax)));
        attr_set = true;
    }
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(kPairThreads);
    lc.dynamicSmemBytes = smem_bytes;
    lc.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    lc.attrs = attr;
    lc.numAttrs = 1;
    TB_CHECK_CUDA(cudaLaunchKernelEx(&lc, conv_pair_kernel<A1, A2, F>, map_a, map_w, map_v, k));
    return 0;
}

// Voxel-stationary tiles (ConvKernelParams::vox) for the im2col kernels: worth it when the taps skipped in the zero
// padding outweigh the frames a partial last block of 256 (128 for single-sub-tile configurations) wastes.
// `valid_out`: fraction of the filter taps that fall inside the volume, averaged over the output voxels.
static bool vox_tiles(const ConvPlan& p, const ConvPlan::Config& cfg, int64_t n_frames, double* valid_out) {
    if (cfg.cluster2 || p.tap2n || p.wfold || getenv("TIMED_B200_NO_VOX")) return false;
    const int vox_rows = cfg.pair ? 256 : 128 * cfg.mt;
    const double valid = valid_tap_fraction(p.Di, p.Do, p.kd, p.pad0[0]) * valid_tap_fraction(p.Hi, p.Ho, p.kh, p.pad0[1]) *
                         valid_tap_fraction(p.Wi, p.Wo, p.kw, p.pad0[2]);
    if (valid_out) *valid_out = valid;
    const int64_t fblocks = (n_frames + vox_rows - 1) / vox_rows;
    const double rows_ratio = static_cast<double>(fblocks * vox_rows) / static_cast<double>(n_frames);
    // every output voxel keeps at least its centre tap
    const bool centre_ok = p.pad0[0] < p.kd && p.pad0[1] < p.kh && p.pad0[2] < p.kw && p.Do <= p.Di && p.Ho <= p.Hi &&
                           p.Wo <= p.Wi;
    // A tile's operand boxes are frame-strided, so their reuse across neighbouring voxels has to come from the L2: the
    // kd input planes x vox_rows frames a z-plane of tiles touches (hi + lo) must fit (measured up to 57 MB: TIMED-338's
    // 512-channel head at 6^3); large volumes stay on the im2col tiling, whose rows are contiguous in memory
    const double live_bytes = static_cast<double>(p.kd) * p.Hi * p.Wi * vox_rows * p.cin_pad * 4.0;
    const bool l2_ok = live_bytes <= 64e6 || getenv("TIMED_B200_FORCE_VOX");
    return centre_ok && l2_ok && (valid * rows_ratio < 0.93 || getenv("TIMED_B200_FORCE_VOX")) &&
           fblocks * p.Do * p.Ho * p.Wo * p.n_tiles < (1ll << 30);
}

// Launch one conv over `n_frames` frames.  `in_base`: split tensor base (hi plane first);
// `in_frames_alloc`: frames per plane in that allocation.
// bytes of fp32 scratch (the Z matrix) a tap-to-N conv needs for n_frames
static size_t conv_scratch_bytes(const ConvPlan& p, int64_t n_frames) {
    if (p.gap_collapse)       // box sums (split planes) + pooled logits
        return static_cast<size_t>(round_up64(2 * n_frames * p.dense->cin_pad * 2, 1024) +
                                   round_up64(n_frames * round_up(p.cout, 4) * 4, 1024));
    if (!p.tap2n || p.c2i_active()) return 0;
    return static_cast<size_t>(round_up64(n_frames * p.Mo_d() * p.Mo_h() * p.Mo_w() * static_cast<int64_t>(p.z_ld) * 4, 1024));
}

static int conv_launch(ConvPlan& p, void* in_base, int64_t in_frames_alloc, int64_t n_frames,
                       const TView& final_out, cudaStream_t stream, void* scratch = nullptr,
                       size_t scratch_bytes = 0, const TensorInfo* out_info = nullptr);

// Linear head conv -> GlobalAveragePooling (-> Softmax), collapsed (ConvPlan::gap_collapse).  `in_base`: the conv's input,
// split NDHWC planes; `final_out`: (frames, classes) fp32 -- probabilities, or pooled logits without the softmax.
static int gap_head_launch(ConvPlan& p, void* in_base, int64_t in_frames_alloc, int64_t n_frames, const TView& final_out,
                           cudaStream_t stream, void* scratch, size_t scratch_bytes) {
    TB_REQUIRE(scratch && scratch_bytes >= conv_scratch_bytes(p, n_frames), "head: scratch too small");
    ConvPlan& d = *p.dense;
    const int64_t s_plane = n_frames * d.cin_pad;                          // elements per split plane of the box sums
    __nv_bfloat16* S = static_cast<__nv_bfloat16*>(scratch);
    float* logits = reinterpret_cast<float*>(static_cast<uint8_t*>(scratch) + round_up64(2 * s_plane * 2, 1024));
    const int taps = p.kd * p.kh * p.kw;
    if (d.cin_pad != taps * p.cin) TB_CHECK_CUDA(cudaMemsetAsync(S, 0, static_cast<size_t>(2 * s_plane) * 2, stream));
    BoxSumParams b;
    b.D = p.Di; b.H = p.Hi; b.W = p.Wi; b.c_pad = p.cin_pad; b.cin = p.cin;
    b.kd = p.kd; b.kh = p.kh; b.kw = p.kw; b.pd = p.pad0[0]; b.ph = p.pad0[1]; b.pw = p.pad0[2];
    b.Do = p.Do; b.Ho = p.Ho; b.Wo = p.Wo;
    b.k_pad = d.cin_pad;
    b.inv_n = 1.0f / static_cast<float>(static_cast<int64_t>(p.Do) * p.Ho * p.Wo);
    b.in_lo_off = in_frames_alloc * p.Di * p.Hi * p.Wi * static_cast<int64_t>(p.cin_pad);
    b.out_lo_off = s_plane;
    const int64_t work = n_frames * ((p.cin + 1) / 2);
    gap_boxsum_kernel<<<grid_for(work, 128), 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(in_base), S, n_frames, b);
    TB_CHECK_CUDA(cudaGetLastError());
    TView lv{};
    lv.fmt = FMT_F32;
    lv.f32 = p.gap_softmax ? logits : final_out.f32;
    lv.ld = p.gap_softmax ? round_up(p.cout, 4) : final_out.ld;
    lv.c = lv.c_pad = p.cout;
    int rc = conv_launch(d, S, n_frames, n_frames, lv, stream);
    if (rc) return rc;
    if (p.gap_softmax) {
        softmax_kernel<<<static_cast<unsigned>((n_frames * 32 + 255) / 256), 256, 0, stream>>>(lv, final_out, n_frames);
        TB_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

template <int A1, int A2, int F>
static int launch_xform_instance(const CUtensorMap& map_x, const CUtensorMap& map_w, const CUtensorMap& map_ohi,
                                 const CUtensorMap& map_olo, const ConvKernelParams& k,
                                 const XformParams& xp, int grid, size_t smem_bytes, cudaStream_t stream) {
    static bool attr_set = false;       // per instantiation
    if (!attr_set) {
        TB_CHECK_CUDA(cudaFuncSetAttribute(bnrelu_conv1x1_kernel<A1, A2, F>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(kSmemDynamicMax)));
        attr_set = true;
    }
    bnrelu_conv1x1_kernel<A1, A2, F><<<grid, kXfThreads, smem_bytes, stream>>>(map_x, map_w, map_ohi, map_olo, k, xp);
    return 0;
}

// Can `c` (already planned) take its BatchNorm -> ReLU pre-activation in its own operand path?
static bool xform_eligible(const ConvPlan& c, int c_in) {
    return c.kd * c.kh * c.kw == 1 && !c.tap2n && !c.slab && !c.thin && !c.thinz && !c.wfold && !c.gap_collapse &&
           c.n_tiles == 1 && c.n_tile % 16 == 0 && c.n_tile <= 128 && c.cin == c_in && c.cin_pad == c_in &&
           c_in % kXfKc == 0 && c_in <= kXfMaxCin && !getenv("TIMED_B200_NO_NFOLD");
}

// BatchNorm -> ReLU -> 1x1x1 conv in one launch (ConvPlan::xform).  `x`: the fp32 tensor the BatchNorm reads (a channel-slice
// view of a Concatenate buffer as a rule: x.ld >= x.c).
static int xform_launch(ConvPlan& p, const TView& x, int64_t n_frames, const TView& out, cudaStream_t stream) {
    TB_REQUIRE(x.fmt == FMT_F32 && x.c == p.cin && x.ld % 4 == 0 && (reinterpret_cast<uintptr_t>(x.f32) & 15) == 0,
               "fused pre-activation: input must be a 16-byte aligned fp32 tensor");
    const int64_t m_total64 = n_frames * p.Do * p.Ho * p.Wo;
    TB_REQUIRE(m_total64 > 0 && m_total64 < (1ll << 31) - 512, "conv: too many output pixels per launch");
    CUtensorMap map_w, map_x;
    int rc = encode_w_map(p, kXfKc, CU_TENSOR_MAP_SWIZZLE_64B, &map_w, p.n_tile);
    if (rc) return rc;
    {
        cuuint64_t dims[2] = {static_cast<cuuint64_t>(p.cin), static_cast<cuuint64_t>(m_total64)};
        cuuint64_t strides[1] = {static_cast<cuuint64_t>(x.ld) * 4};
        cuuint32_t box[2] = {static_cast<cuuint32_t>(kXfKc), 128};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = g_encode_tiled(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x.f32, dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(fp32 activations) failed, CUresult=" + std::to_string(r));
            return TB_ERR_CUDA;
        }
    }
    ConvKernelParams k;
    std::memset(&k, 0, sizeof(k));
    k.m_total = static_cast<int32_t>(m_total64);
    k.n_ctile_m = static_cast<int32_t>((m_total64 + 127) / 128);
    k.mt = 1;
    k.n_tiles = 1;
    k.n_tile = p.n_tile;
    k.acc_cols = round_up(2 * p.n_tile, 32);
    k.acc_stages = 2;
    k.nfold = 1;
    k.n_kblocks = p.cin / kXfKc;
    k.w_lo_rows = p.n_alloc;
    k.w_sub_bytes = static_cast<uint32_t>(p.n_tile) * kXfKc * 2u;
    // split-plane output through a shared-memory staging tile + bulk tensor stores (XformParams::tstore)
    const bool tstore = out.fmt == FMT_SPLIT && p.n_tile % 64 == 0 && out.c_pad == p.n_tile && out.ld % 8 == 0 &&
                        (reinterpret_cast<uintptr_t>(out.hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(out.lo) & 15) == 0 &&
                        !getenv("TIMED_B200_NO_TSTORE");
    CUtensorMap map_ohi = map_x, map_olo = map_x;
    if (tstore) {
        cuuint64_t dims[2] = {static_cast<cuuint64_t>(out.c_pad), static_cast<cuuint64_t>(m_total64)};
        cuuint64_t strides[1] = {static_cast<cuuint64_t>(out.ld) * 2};
        cuuint32_t box[2] = {64, 128};
        cuuint32_t estr[2] = {1, 1};
        for (int pl = 0; pl < 2; ++pl) {
            CUresult r = g_encode_tiled(pl ? &map_olo : &map_ohi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl ? out.lo : out.hi, dims,
                                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                set_error("cuTensorMapEncodeTiled(split output) failed, CUresult=" + std::to_string(r));
                return TB_ERR_CUDA;
            }
        }
    }
    // operand ring: 3 stages of (A hi/lo + W hi/lo); output staging tile; fp32 ring: whatever is left
    const size_t stage = 2u * 128u * kXfKc * 2u + 2u * k.w_sub_bytes, x_stage = 128u * kXfKc * 4u;
    const size_t out_stage = tstore ? 2u * static_cast<size_t>(p.n_tile / 64) * 16384u : 0u;
    k.stages = 3;
    if (const char* e = getenv("TIMED_B200_XFORM_AWSTAGES")) k.stages = std::max(2, std::min<int>(kXfMaxStages, atoi(e)));     // A/B
    const size_t budget = kSmemDynamicMax - 8 * 1024 - 1024 - out_stage;
    int x_stages = static_cast<int>(std::min<size_t>(kXfMaxXStages, (budget - stage * k.stages) / x_stage));
    if (const char* e = getenv("TIMED_B200_XFORM_XSTAGES")) x_stages = std::max(k.stages, std::min(x_stages, atoi(e)));   // A/B
    TB_REQUIRE(x_stages >= k.stages, "fused pre-activation: shared memory too small");
    k.bias = p.d_bias; k.scale = p.d_scale; k.shift = p.d_shift;
    k.act1 = p.act1; k.act2 = p.act2; k.alpha1 = p.alpha1; k.alpha2 = p.alpha2;
    k.out_fmt = out.fmt;
    k.out_f32 = out.f32; k.out_hi = out.hi; k.out_lo = out.lo;
    k.ldc = out.ld;
    k.c_store = out.fmt == FMT_SPLIT ? out.c_pad : out.c;
    k.acc_comp = accum_comp(static_cast<double>(ceil_div(p.cin, 16)), 1.0);
    TB_REQUIRE(out.fmt != FMT_SPLIT || (out.c_pad % 16 == 0 && out.c_pad <= p.n_alloc),
               "conv: split output channel padding mismatch");
    XformParams xp;
    xp.in_scale = p.xf_scale;
    xp.in_shift = p.xf_shift;
    xp.c_in = p.cin;
    xp.x_stages = x_stages;
    xp.tstore = tstore ? 1 : 0;
    const size_t smem_bytes = stage * k.stages + x_stage * x_stages + out_stage + 1024;
    const int grid = std::min(k.n_ctile_m, 148);
    bool launched = false;
#define TB_XF_CASE(A1, A2, F)                                                                  \
    if (!launched && k.act1 == (A1) && k.act2 == (A2) && k.out_fmt == (F)) {                   \
        rc = launch_xform_instance<A1, A2, F>(map_x, map_w, map_ohi, map_olo, k, xp, grid, smem_bytes, stream);  \
        launched = true;                                                                       \
    }
    TB_XF_CASE(ACT_NONE, ACT_RELU, FMT_SPLIT)
    TB_XF_CASE(ACT_NONE, ACT_NONE, FMT_SPLIT)
    TB_XF_CASE(ACT_RELU, ACT_NONE, FMT_SPLIT)
#undef TB_XF_CASE
    if (!launched)
        rc = k.out_fmt == FMT_SPLIT ? launch_xform_instance<-1, -1, FMT_SPLIT>(map_x, map_w, map_ohi, map_olo, k, xp, grid, smem_bytes, stream)
                                    : launch_xform_instance<-1, -1, FMT_F32>(map_x, map_w, map_ohi, map_olo, k, xp, grid, smem_bytes, stream);
    if (rc) return rc;
    TB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int conv_launch(ConvPlan& p, void* in_base, int64_t in_frames_alloc, int64_t n_frames,
                       const TView& final_out, cudaStream_t stream, void* scratch,
                       size_t scratch_bytes, const TensorInfo* out_info) {
    if (p.gap_collapse) return gap_head_launch(p, in_base, in_frames_alloc, n_frames, final_out, stream, scratch, scratch_bytes);
    if (p.slab) return slab_launch(p, in_base, n_frames, final_out, stream);
    if (p.thinz) return thinz_launch(p, in_base, in_frames_alloc, n_frames, final_out, stream, out_info);
    if (p.thin) return thin_launch(p, in_base, in_frames_alloc, n_frames, final_out, stream);
    TView out = final_out;
    const bool c2i = p.c2i_active();
    if (p.tap2n && !c2i) {
        TB_REQUIRE(scratch && scratch_bytes >= conv_scratch_bytes(p, n_frames), "conv: scratch too small");
        out = TView{};
        out.fmt = FMT_F32;
        out.f32 = static_cast<float*>(scratch);
        out.ld = p.z_ld;
        out.c = out.c_pad = p.z_cols;
    }
    const int64_t m_total64 = n_frames * p.Mo_d() * p.Mo_h() * p.Mo_w();
    TB_REQUIRE(m_total64 > 0 && m_total64 < (1ll << 31) - 512, "conv: too many output pixels per launch");
    ConvPlan::Config cfg;
    int rc = choose_config(p, m_total64, &cfg);
    if (rc) return rc;
    CUtensorMap map_w, map_a;
    rc = encode_w_map(p, cfg.kc, cfg.tma_swz, &map_w, (cfg.cluster2 || cfg.pair) ? p.n_tile / 2 : p.n_tile);
    if (rc) return rc;
    rc = encode_a_map(p, cfg, in_base, in_frames_alloc, p.cin_pad, &map_a);
    if (rc) return rc;
    CUtensorMap map_v = map_a;
    const int vox_rows = cfg.pair ? 256 : 128 * cfg.mt;          // frames per voxel-stationary tile
    const bool vox = vox_tiles(p, cfg, n_frames, nullptr);
    if (vox) {
        rc = encode_a_vox_map(p, cfg, in_base, in_frames_alloc, &map_v);
        if (rc) return rc;
    }

    ConvKernelParams k;
    std::memset(&k, 0, sizeof(k));
    k.m_total = static_cast<int32_t>(m_total64);
    const int m_tiles = static_cast<int>((m_total64 + 127) / 128);
    k.mt = cfg.mt;
    k.n_ctile_m = ceil_div(m_tiles, (cfg.cluster2 || cfg.pair) ? 2 : cfg.mt);   // cluster / pair mode: tiles are 256-row pair-tiles
    if (c2i) {
        TB_REQUIRE(cfg.mt == 1 && !cfg.cluster2 && !cfg.pair && p.n_tiles == 1 && p.n_tile == 3 * p.cout && p.c2i_rows > 0,
                   "conv: fused col2im needs one 128-row sub-tile and N = 3 * C_out");
        k.c2i = 1;
        k.c2i_rows = p.c2i_rows;
        k.c2i_w = p.Wo;
        k.c2i_cout = p.cout;
        k.n_ctile_m = static_cast<int32_t>((m_total64 + p.c2i_rows - 1) / p.c2i_rows);
    }
    if (vox) {
        k.vox = 1;
        k.vox_frames = static_cast<int32_t>(n_frames);
        // two rounds of the persistent grid: 74 CTA pairs, or 148 CTAs
        const int window = (getenv("TIMED_B200_TILE_WINDOW") ? atoi(getenv("TIMED_B200_TILE_WINDOW")) : 148) * (cfg.pair ? 1 : 2);
        if (window > 0) {
            if (!p.d_progress) TB_CHECK_CUDA(cudaMalloc(&p.d_progress, sizeof(int32_t)));
            TB_CHECK_CUDA(cudaMemsetAsync(p.d_progress, 0, sizeof(int32_t), stream));
            k.progress = p.d_progress;
            k.window = window;
        }
        k.Di = p.Di; k.Hi = p.Hi; k.Wi = p.Wi;
        k.n_ctile_m = static_cast<int32_t>(((n_frames + vox_rows - 1) / vox_rows) * p.Do * p.Ho * p.Wo);
    }
    k.cluster2 = cfg.cluster2;
    k.corr_off = cfg.corr_off;
    k.n_tiles = p.n_tiles;
    k.n_tile = p.n_tile;
    k.acc_cols = cfg.acc_cols;
    k.acc_stages = cfg.acc_stages;
    k.nfold = cfg.nfold;
    k.kh = (p.tap2n && !p.t2n_w) ? 1 : p.kh; k.kw = (p.wfold || (p.tap2n && !p.t2n_kw)) ? 1 : p.kw;
    k.n_taps = p.taps_eff();
    k.cin_pad = p.cin_pad;
    k.kc = cfg.kc;
    k.cin_blocks = p.cin_pad / cfg.kc;
    k.kg = cfg.kg;
    k.n_kblocks = k.n_taps * k.cin_blocks;
    k.stages = cfg.stages;
    k.Do = p.Mo_d(); k.Ho = p.Mo_h(); k.Wo = p.Mo_w();
    k.lc_d = (p.tap2n && !p.t2n_w) ? 0 : -p.pad0[0];
    k.lc_h = (p.tap2n && !p.t2n_w) ? 0 : -p.pad0[1];
    k.lc_w = (p.wfold || (p.tap2n && !p.t2n_kw)) ? 0 : -p.pad0[2];
    k.lo_plane_frames = static_cast<int32_t>(in_frames_alloc);
    k.w_lo_rows = p.n_alloc;
    k.a_sub_bytes = 128u * cfg.kc * 2u;
    k.w_sub_bytes = static_cast<uint32_t>(cfg.pair ? p.n_tile / 2 : p.n_tile) * cfg.kc * 2u;   // pair: per-CTA half tile
    k.row_bytes = cfg.kc * 2u;
    k.layout_type = cfg.swizzle_code;
    k.bias = p.d_bias; k.scale = p.d_scale; k.shift = p.d_shift;
    if (c2i) { k.bias = p.d_c2i_bias; k.scale = p.d_c2i_scale; k.shift = p.d_c2i_shift; }   // the layer's own epilogue
    k.act1 = p.act1; k.act2 = p.act2; k.alpha1 = p.alpha1; k.alpha2 = p.alpha2;
    if (p.tap2n && !c2i) k.act1 = k.act2 = ACT_NONE;     // applied by the col2im kernel
    k.out_fmt = out.fmt;
    k.out_f32 = out.f32; k.out_hi = out.hi; k.out_lo = out.lo;
    k.ldc = out.ld;
    k.c_store = out.fmt == FMT_SPLIT ? out.c_pad : out.c;
    k.dbg = debug_mask();
    {
        // MMAs with real channels per tap (zero-padded K=16 blocks add exact zeros) x taps, x the fraction of taps that
        // fall inside the volume; three times as many when the corrections share the main accumulator
        const double per_tap = p.wfold ? p.kwin * 8 / 16.0 : ceil_div(p.cin, 16);
        double valid = 1.0;
        if (!p.tap2n && !p.wfold) {
            valid = valid_tap_fraction(p.Di, p.Do, p.kd, p.pad0[0]) * valid_tap_fraction(p.Hi, p.Ho, p.kh, p.pad0[1]) *
                    valid_tap_fraction(p.Wi, p.Wo, p.kw, p.pad0[2]);
        } else if (p.wfold) {
            valid = valid_tap_fraction(p.Di, p.Do, p.kd, p.pad0[0]) * valid_tap_fraction(p.Hi, p.Ho, p.kh, p.pad0[1]);
        } else if (p.t2n_kw) {
            valid = valid_tap_fraction(p.Wi, p.Wo, p.kw, p.pad0[2]);
        } else if (p.t2n_w) {
            valid = valid_tap_fraction(p.Di, p.Do, p.kd, p.pad0[0]) * valid_tap_fraction(p.Hi, p.Ho, p.kh, p.pad0[1]);
        }
        const bool shared_acc = !cfg.nfold && !cfg.corr_off;
        k.acc_comp = accum_comp(per_tap * k.n_taps * (shared_acc ? 3.0 : 1.0), valid);
    }
    TB_REQUIRE(out.fmt != FMT_SPLIT || (out.c_pad % 16 == 0 && out.c_pad <= p.n_alloc),
               "conv: split output channel padding mismatch");

    const int total_tiles = k.n_ctile_m * k.n_tiles;
    const int grid = (cfg.cluster2 || cfg.pair) ? 2 * std::min(total_tiles, 74) : std::min(total_tiles, 148);
    // compile-time specialised epilogues for the activation pairs Keras graphs actually produce;
    // everything else goes through the runtime-dispatched instance
    bool launched = false;
#define TB_CONV_CASE(A1, A2, F)                                                                         \
    if (!launched && k.act1 == (A1) && k.act2 == (A2) && k.out_fmt == (F)) {                            \
        rc = cfg.pair ? launch_pair_instance<A1, A2, F>(map_a, map_w, map_v, k, grid, cfg.smem_bytes, stream)  \
                      : launch_conv_instance<A1, A2, F>(map_a, map_w, map_v, k, grid, cfg.smem_bytes, stream); \
        launched = true;                                                                                \
    }
    TB_CONV_CASE(ACT_ELU, ACT_NONE, FMT_SPLIT)
    TB_CONV_CASE(ACT_ELU, ACT_NONE, FMT_F32)
    TB_CONV_CASE(ACT_NONE, ACT_NONE, FMT_SPLIT)
    TB_CONV_CASE(ACT_NONE, ACT_NONE, FMT_F32)
    TB_CONV_CASE(ACT_RELU, ACT_NONE, FMT_SPLIT)
    TB_CONV_CASE(ACT_RELU, ACT_NONE, FMT_F32)
    TB_CONV_CASE(ACT_NONE, ACT_RELU, FMT_SPLIT)
    TB_CONV_CASE(ACT_NONE, ACT_RELU, FMT_F32)
    TB_CONV_CASE(ACT_NONE, ACT_ELU, FMT_SPLIT)
    TB_CONV_CASE(ACT_NONE, ACT_ELU, FMT_F32)
#undef TB_CONV_CASE
    if (!launched && cfg.pair)
        rc = k.out_fmt == FMT_SPLIT ? launch_pair_instance<-1, -1, FMT_SPLIT>(map_a, map_w, map_v, k, grid, cfg.smem_bytes, stream)
                                    : launch_pair_instance<-1, -1, FMT_F32>(map_a, map_w, map_v, k, grid, cfg.smem_bytes, stream);
    else if (!launched)
        rc = k.out_fmt == FMT_SPLIT ? launch_conv_instance<-1, -1, FMT_SPLIT>(map_a, map_w, map_v, k, grid, cfg.smem_bytes, stream)
                                    : launch_conv_instance<-1, -1, FMT_F32>(map_a, map_w, map_v, k, grid, cfg.smem_bytes, stream);
    if (rc) return rc;
    TB_CHECK_CUDA(cudaGetLastError());
    if (p.tap2n && !c2i) {
        Col2imParams cp;
        cp.Di = p.Mo_d(); cp.Hi = p.Mo_h(); cp.Wi = p.Mo_w(); cp.Do = p.Do; cp.Ho = p.Ho; cp.Wo = p.Wo;
        cp.kd = p.t2n_w ? 1 : p.kd; cp.kh = p.t2n_w ? 1 : p.kh;              // t2n_w: the GEMM already summed over (kd,kh)
        cp.kw = p.t2n_kw ? 1 : p.kw;                                         // t2n_kw: ... over kw
        cp.pd = p.t2n_w ? 0 : p.pad0[0]; cp.ph = p.t2n_w ? 0 : p.pad0[1]; cp.pw = p.t2n_kw ? 0 : p.pad0[2];
        cp.cout = p.cout;
        cp.z_ld = p.z_ld;
        cp.act1 = p.act1; cp.act2 = p.act2; cp.alpha1 = p.alpha1; cp.alpha2 = p.alpha2;
        if (p.fuse_head) {
            // final_out is the (frames, classes) probability matrix
            const size_t smem = (static_cast<size_t>(p.Do) * p.Ho * p.Wo * p.cout + p.cout) * sizeof(float);
            static bool attr_set = false;
            if (!attr_set) {
                TB_CHECK_CUDA(cudaFuncSetAttribute(head_col2im_pool_softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(kHeadSmemMax)));
                attr_set = true;
            }
            head_col2im_pool_softmax_kernel<<<static_cast<unsigned>(n_frames), 256, smem, stream>>>(
                static_cast<const float*>(scratch), cp, p.d_c2i_bias, p.d_c2i_scale, p.d_c2i_shift, p.head_is_avg, final_out.f32);
            TB_CHECK_CUDA(cudaGetLastError());
            return 0;
        }
        const int cw = final_out.fmt == FMT_SPLIT ? final_out.c_pad : final_out.c;
        const int64_t work = n_frames * p.Do * p.Ho * p.Wo * cw;
        if (final_out.fmt == FMT_F32 && p.cout % 4 == 0 && p.z_ld % 4 == 0 && final_out.ld % 4 == 0 &&
            (reinterpret_cast<uintptr_t>(final_out.f32) & 15) == 0 && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0)
            col2im_vec4_kernel<<<grid_for(work / 4, 256), 256, 0, stream>>>(static_cast<const float*>(scratch), final_out,
                                                                            n_frames, cp, p.d_c2i_bias, p.d_c2i_scale,
                                                                            p.d_c2i_shift);
        else
            col2im_kernel<<<grid_for(work, 256), 256, 0, stream>>>(static_cast<const float*>(scratch), final_out,
                                                                   n_frames, cp, p.d_c2i_bias, p.d_c2i_scale,
                                                                   p.d_c2i_shift);
        TB_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

// ----------------------------------------------------------------------------- graph
struct OpNode {
    int alias_of = -1;          // pooling op fused into the conv that feeds it: shares that op's output tensor
    bool skip = false;          // runs inside another op's launch (fused network head)
    bool pool_softmax = false;  // GPOOL whose only reader is the final SOFTMAX: one launch writes the probabilities
    tb_op_desc d;
    ConvPlan conv;
    float* d_scale = nullptr;   // AFFINE
    float* d_shift = nullptr;
    PoolParams pool;
};

struct Layout {
    std::vector<size_t> offset;
    size_t scratch_off = 0, scratch_bytes = 0;   // shared fp32 scratch of the tap-to-N convs
    size_t total = 0;
};

}  // namespace tb

struct tb_graph {
    int device = 0;
    std::vector<tb::OpNode> ops;
    std::vector<tb::TensorInfo> tensors;
    int n_classes = 0;
    double flops = 0.0;
    int launches = 0;
    std::map<int64_t, tb::Layout> layouts;
    // host-path staging (predict_host)
    void* d_stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    void* d_ws = nullptr;
    size_t ws_bytes = 0;
    float* d_probs = nullptr;
    size_t probs_bytes = 0;
    int64_t last_passes = 0, last_max_pass = 0;          // how the last predict_host call was chunked
    cudaStream_t s_copy = nullptr, s_compute = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    // per-op device timing (timed_b200_graph_set_timing): one event set per recorded forward
    bool timing = false;
    std::vector<std::vector<cudaEvent_t>> timing_sets;   // [forward][n_ops + 1]
    size_t timing_used = 0;
    static constexpr size_t kMaxTimedForwards = 512;
};

namespace tb {

static TView make_view(const TensorInfo& t, uint8_t* base, int64_t n_frames) {
    TView v;
    v.fmt = t.fmt;
    v.c = t.C;
    v.c_pad = t.c_pad;
    v.ld = t.c_pad;
    v.f32 = nullptr;
    v.hi = v.lo = nullptr;
    if (t.fmt == FMT_F32) {
        v.f32 = reinterpret_cast<float*>(base);      // (for a view the caller passes the root's base and fixes ld / offset)
    } else if (t.cpv) {
        v.hi = reinterpret_cast<__nv_bfloat16*>(base);
        v.lo = v.hi + (t.c_pad / 8) * t.cpv_T(n_frames) * 8;
    } else {
        v.hi = reinterpret_cast<__nv_bfloat16*>(base);
        v.lo = v.hi + t.frames_alloc(n_frames) * t.stored_pix_per_frame() * t.c_pad;
    }
    return v;
}

// view of tensor `idx` inside the workspace: a zero-copy Concatenate slice points into its root's buffer
static TView tensor_view(const tb_graph* g, const Layout& L, int idx, uint8_t* base, int64_t n_frames) {
    const TensorInfo& t = g->tensors[idx];
    TView v = make_view(t, base + L.offset[idx], n_frames);
    if (t.view_root >= 0) {
        v.f32 += t.view_c0;
        v.ld = g->tensors[t.view_root].C;
    }
    return v;
}

static CpvGeom make_cpv_geom(const TensorInfo& t, int64_t n_frames) {
    CpvGeom g;
    g.T = t.cpv_T(n_frames);
    g.lead = t.cpv_lead;
    g.n_pos = n_frames * t.cpv_Dp * t.cpv_Hp * t.cpv_Wp;
    g.D = t.D; g.H = t.H; g.W = t.W;
    g.Dp = t.cpv_Dp; g.Hp = t.cpv_Hp; g.Wp = t.cpv_Wp;
    g.n_chunks = t.c_pad / 8;
    g.c = t.C;
    return g;
}

// Liveness-based first-fit workspace layout for n_frames.
static const Layout& get_layout(tb_graph* g, int64_t n_frames) {
    auto it = g->layouts.find(n_frames);
    if (it != g->layouts.end()) return it->second;
    Layout L;
    const int n = static_cast<int>(g->tensors.size());
    L.offset.assign(n, 0);
    struct Block { size_t off, size; };
    std::vector<Block> free_list;
    size_t top = 0;
    std::vector<std::pair<size_t, size_t>> live(n, {0, 0});
    std::vector<bool> placed_flag(n, false);
    for (int i = 0; i < n; ++i) {
        // release tensors whose last reader ran before op i
        for (int j = 0; j < i; ++j)
            if (live[j].second && g->tensors[j].last_use < i) {
                free_list.push_back({live[j].first, live[j].second});
                live[j].second = 0;
            }
        // coalesce
        std::sort(free_list.begin(), free_list.end(), [](const Block& a, const Block& b) { return a.off < b.off; });
        for (size_t k = 0; k + 1 < free_list.size();) {
            if (free_list[k].off + free_list[k].size == free_list[k + 1].off) {
                free_list[k].size += free_list[k + 1].size;
                free_list.erase(free_list.begin() + k + 1);
            } else {
                ++k;
            }
        }
        if (g->ops[i].alias_of >= 0) {                 // fused pooling op: same storage as the conv output
            L.offset[i] = L.offset[g->ops[i].alias_of];
            live[i] = {L.offset[i], 0};
            continue;
        }
        if (g->ops[i].skip && g->ops[i].d.op == TB_OP_AFFINE) {     // applied inside the conv that reads it: never materialised
            L.offset[i] = 0;
            live[i] = {0, 0};
            continue;
        }
        // zero-copy Concatenate: the first view to be produced places the root buffer; views themselves take no storage
        int place = i;
        if (g->tensors[i].view_root >= 0) place = g->tensors[i].view_root;
        if (placed_flag[place]) {
            if (place != i) { L.offset[i] = L.offset[place]; live[i] = {L.offset[i], 0}; }
            continue;
        }
        placed_flag[place] = true;
        const size_t need = g->tensors[place].bytes(n_frames);
        bool placed = false;
        for (size_t k = 0; k < free_list.size(); ++k)
            if (free_list[k].size >= need) {
                L.offset[place] = free_list[k].off;
                free_list[k].off += need;
                free_list[k].size -= need;
                if (!free_list[k].size) free_list.erase(free_list.begin() + k);
                placed = true;
                break;
            }
        if (!placed) {
            // grow from the top; merge with a trailing free block if there is one
            if (!free_list.empty() && free_list.back().off + free_list.back().size == top) {
                L.offset[place] = free_list.back().off;
                top = free_list.back().off + need;
                free_list.pop_back();
            } else {
                L.offset[place] = top;
                top += need;
            }
        }
        live[place] = {L.offset[place], need};
        if (place != i) { L.offset[i] = L.offset[place]; live[i] = {L.offset[i], 0}; }
    }
    L.scratch_off = static_cast<size_t>(round_up64(static_cast<int64_t>(top), 1024));
    for (const auto& op : g->ops)
        L.scratch_bytes = std::max(L.scratch_bytes, conv_scratch_bytes(op.conv, n_frames));
    L.total = L.scratch_off + L.scratch_bytes;
    return g->layouts.emplace(n_frames, std::move(L)).first->second;
}

static int upload_vec(const float* src, int n, float fill, float** dst) {
    std::vector<float> v(n, fill);
    if (src) std::copy(src, src + n, v.begin());
    TB_CHECK_CUDA(cudaMalloc(dst, n * sizeof(float)));
    TB_CHECK_CUDA(cudaMemcpy(*dst, v.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

// Give tensor `ti` (dims known, produced by the input op or a pooling op) the CPV layout when every reader is a
// conv that slab_conv_kernel can run; margins are the union over the readers.
static void decide_cpv(TensorInfo& t, int ti, const tb_op_desc* ops, int n_ops) {   // ti: op whose READERS decide
    int pd = 0, ph = 0, pw = 0, n_cons = 0;
    for (int pass = 0; pass < 2; ++pass)          // second pass: every reader must also fit with the union margins
        for (int i = ti + 1; i < n_ops; ++i)
            for (int k = 0; k < ops[i].n_inputs; ++k)
                if (ops[i].inputs[k] == ti) {
                    SlabGeom g;
                    if (!slab_geometry(t.D, t.H, t.W, t.C, ops[i], &g, pd, ph, pw)) return;
                    pd = g.pd; ph = g.ph; pw = g.pw;
                    if (pass == 0) ++n_cons;
                }
    if (!n_cons) return;
    int neg = 0, pos = 0, mt = 1;
    for (int i = ti + 1; i < n_ops; ++i)
        for (int k = 0; k < ops[i].n_inputs; ++k)
            if (ops[i].inputs[k] == ti) {
                SlabGeom g;
                slab_geometry(t.D, t.H, t.W, t.C, ops[i], &g, pd, ph, pw);
                neg = std::max(neg, g.neg); pos = std::max(pos, g.pos); mt = std::max(mt, g.mt);
            }
    t.cpv = true;
    t.cpv_Dp = t.D + pd; t.cpv_Hp = t.H + ph; t.cpv_Wp = t.W + pw;
    t.cpv_lead = round_up(neg, 8);
    t.cpv_tail = round_up(pos + 128 * mt, 8);
}

static int graph_build(tb_graph* g, const tb_op_desc* ops, int n_ops) {
    TB_REQUIRE(n_ops >= 2, "graph needs an input op and at least one more op");
    TB_REQUIRE(ops[0].op == TB_OP_INPUT, "ops[0] must be TB_OP_INPUT");
    g->ops.resize(n_ops);
    g->tensors.resize(n_ops);
    // consumers decide the storage format: anything a contraction reads lives as bf16 split planes
    std::vector<bool> feeds_conv(n_ops, false);
    for (int i = 0; i < n_ops; ++i) {
        const tb_op_desc& d = ops[i];
        TB_REQUIRE(d.n_inputs >= 0 && d.n_inputs <= TB_MAX_INPUTS, "bad n_inputs");
        for (int k = 0; k < d.n_inputs; ++k) {
            TB_REQUIRE(d.inputs[k] >= 0 && d.inputs[k] < i, "op inputs must reference earlier ops");
            if (d.op == TB_OP_CONV3D) feeds_conv[d.inputs[k]] = true;
            g->tensors[d.inputs[k]].last_use = i;
        }
    }
    g->tensors[n_ops - 1].last_use = n_ops;   // graph output stays live
    // padded-volume input layout + thin_conv_kernel: thin input (C <= 8) read only by stride-1 convs
    // whose K=16 step count is bounded and whose output channels fit one N tile
    {
        bool ok = ops[0].c_out <= 8 && !getenv("TIMED_B200_NO_THIN");
        int d0 = 0, d1 = 0, h0 = 0, h1 = 0, w0 = 0, w1 = 0, n_cons = 0;
        for (int i = 1; i < n_ops && ok; ++i)
            for (int k = 0; k < ops[i].n_inputs; ++k)
                if (ops[i].inputs[k] == 0) {
                    ++n_cons;
                    const tb_op_desc& c = ops[i];
                    if (c.op != TB_OP_CONV3D || !thin_fits(c.kernel[0], c.kernel[1], c.kernel[2], c.c_out)) {
                        ok = false;
                        break;
                    }
                    int pb[3] = {0, 0, 0}, pa[3] = {0, 0, 0};
                    for (int a = 0; a < 3 && c.pad_same; ++a) {
                        int o;
                        same_pads(ops[0].kernel[a], c.kernel[a], 1, &o, &pb[a], &pa[a]);
                    }
                    d0 = std::max(d0, pb[0]); d1 = std::max(d1, pa[0]);
                    h0 = std::max(h0, pb[1]); h1 = std::max(h1, pa[1]);
                    w0 = std::max(w0, pb[2]); w1 = std::max(w1, pa[2]);
                }
        if (ok && n_cons > 0) {
            TensorInfo& t0 = g->tensors[0];
            t0.padvol = true;
            t0.pv_d0 = d0; t0.pv_h0 = h0; t0.pv_w0 = w0;
            t0.pv_Dp = d0 + ops[0].kernel[0] + d1;
            t0.pv_Hp = h0 + ops[0].kernel[1] + h1;
            t0.pv_Wp = w0 + ops[0].kernel[2] + w1;
        }
    }
    // W-folded input layout: thin input (C <= 8) read only by stride-1 convs with kw <= 8
    if (!g->tensors[0].padvol) {
        bool ok = ops[0].c_out <= 8 && !getenv("TIMED_B200_NO_WFOLD");
        int lm = 0, rm = 0, n_cons = 0;
        for (int i = 1; i < n_ops && ok; ++i)
            for (int k = 0; k < ops[i].n_inputs; ++k)
                if (ops[i].inputs[k] == 0) {
                    ++n_cons;
                    if (ops[i].op != TB_OP_CONV3D || ops[i].kernel[2] > 8 || ops[i].kernel[2] < 1) { ok = false; break; }
                    const int kw = ops[i].kernel[2];
                    const int pad0 = ops[i].pad_same ? (kw - 1) / 2 : 0;
                    const int kwin = kw <= 4 ? 4 : 8;
                    lm = std::max(lm, pad0);
                    rm = std::max(rm, kwin - 1 - pad0);
                }
        if (ok && n_cons > 0) {
            TensorInfo& t0 = g->tensors[0];
            t0.wfold = true;
            t0.wf_lm = lm;
            t0.wf_pitch = lm + ops[0].kernel[2] + rm;
        }
    }
    g->flops = 0.0;
    g->launches = 0;
    for (int i = 0; i < n_ops; ++i) {
        OpNode& node = g->ops[i];
        node.d = ops[i];
        const tb_op_desc& d = ops[i];
        TensorInfo& t = g->tensors[i];
        t.fmt = feeds_conv[i] ? FMT_SPLIT : FMT_F32;
        const TensorInfo* in0 = d.n_inputs > 0 ? &g->tensors[d.inputs[0]] : nullptr;
        switch (d.op) {
            case TB_OP_INPUT:
                TB_REQUIRE(i == 0, "only ops[0] may be TB_OP_INPUT");
                t.D = d.kernel[0]; t.H = d.kernel[1]; t.W = d.kernel[2]; t.C = d.c_out;
                TB_REQUIRE(t.D > 0 && t.H > 0 && t.W > 0 && t.C > 0, "input dims must be positive");
                if (!t.padvol && !t.wfold) decide_cpv(t, i, ops, n_ops);
                g->launches += 1;
                break;
            case TB_OP_CONV3D: {
                TB_REQUIRE(d.n_inputs == 1, "conv takes one input");
                int rc = conv_plan_create(node.conv, d, in0->D, in0->H, in0->W, in0->C, in0->c_pad, in0);
                if (rc) return rc;
                t.D = node.conv.Do; t.H = node.conv.Ho; t.W = node.conv.Wo; t.C = d.c_out;
                g->flops += node.conv.flops_per_frame();
                g->launches += (node.conv.tap2n && !node.conv.c2i) ? 2 : 1;     // GEMM + col2im, unless the col2im is in the epilogue
                // thinz conv read only by a MaxPool(2,2,2; stride 2): the pooling runs in the conv epilogue and this
                // op's tensor IS the pooled tensor (the pooling op becomes an alias of it)
                // Round 1 staged ACTIVATED outputs over overlapping windows (+25 % MMA work) and measured neutral; round 2 pools
                // the raw sums (monotone epilogue) over a rolling two-window ring: no extra MMA work, 8x less epilogue math.
                if (node.conv.thinz && !getenv("TIMED_B200_NO_POOLFUSE") &&
                    (d.act1 != TB_ACT_ELU || d.alpha1 >= 0.f) && (d.act2 != TB_ACT_ELU || d.alpha2 >= 0.f)) {
                    int j = -1, n_readers = 0;
                    for (int k = i + 1; k < n_ops; ++k)
                        for (int q = 0; q < ops[k].n_inputs; ++q)
                            if (ops[k].inputs[q] == i) { ++n_readers; j = k; }
                    bool ok = n_readers == 1 && ops[j].op == TB_OP_POOL3D && ops[j].pool_kind == 0 && ops[j].n_inputs == 1;
                    for (int a = 0; a < 3 && ok; ++a)
                        ok = ops[j].kernel[a] == 2 && (ops[j].stride[a] == 2 || ops[j].stride[a] <= 0);
                    ThinZGeom zg;
                    ok = ok && thinz_geometry(node.conv.kd, node.conv.kh, node.conv.kw, node.conv.cout, in0->pv_Wp, &zg, true);
                    const bool same = ok && ops[j].pad_same;
                    const int dims[3] = {t.D, t.H, t.W};
                    int po[3] = {0, 0, 0};
                    for (int a = 0; a < 3 && ok; ++a) {
                        po[a] = same ? (dims[a] + 1) / 2 : dims[a] / 2;
                        ok = po[a] >= 1;
                    }
                    // a plain fp32 pooled tensor is written 8 channels at a time
                    ok = ok && (feeds_conv[j] || t.C % 8 == 0);
                    if (ok) {
                        node.conv.fuse_pool = true;
                        node.conv.pool_same = same ? 1 : 0;
                        node.conv.pool_Zo = po[0]; node.conv.pool_Po = po[1]; node.conv.pool_Qo = po[2];
                        g->ops[j].alias_of = i;
                        t.D = po[0]; t.H = po[1]; t.W = po[2];
                        t.fmt = feeds_conv[j] ? FMT_SPLIT : FMT_F32;
                        t.last_use = std::max(t.last_use, g->tensors[j].last_use);
                        decide_cpv(t, j, ops, n_ops);
                        if (t.cpv) g->launches += 1;          // margin zeroing
                    }
                }
                // Default: only the z direction of that max-pool runs in the epilogue (max over accumulator pairs: no
                // staging, no barrier); the pooling op that follows becomes 1x2x2 over half the data.
                if (node.conv.thinz && !node.conv.fuse_pool && !getenv("TIMED_B200_NO_ZPOOL")) {
                    int j = -1, n_readers = 0;
                    for (int k = i + 1; k < n_ops; ++k)
                        for (int q = 0; q < ops[k].n_inputs; ++q)
                            if (ops[k].inputs[q] == i) { ++n_readers; j = k; }
                    bool ok = n_readers == 1 && ops[j].op == TB_OP_POOL3D && ops[j].pool_kind == 0 && ops[j].n_inputs == 1 &&
                              ops[j].kernel[0] == 2 && (ops[j].stride[0] == 2 || ops[j].stride[0] <= 0) &&
                              node.conv.thinz_params.zt % 2 == 0 &&
                              // the epilogue pools the RAW sums: its map must be monotone (ELU with a negative alpha is not)
                              (d.act1 != TB_ACT_ELU || d.alpha1 >= 0.f) && (d.act2 != TB_ACT_ELU || d.alpha2 >= 0.f);
                    const int zo = ok ? (ops[j].pad_same ? (t.D + 1) / 2 : t.D / 2) : 0;
                    if (ok && zo >= 1) {
                        node.conv.fuse_zpool = true;
                        node.conv.pool_same = ops[j].pad_same ? 1 : 0;
                        node.conv.pool_Zo = zo;
                        t.D = zo;
                    }
                }
                break;
            }
            case TB_OP_POOL3D: {
                TB_REQUIRE(d.n_inputs == 1, "pool takes one input");
                if (node.alias_of >= 0) {                     // fused into the producing conv's epilogue
                    const int lu = t.last_use;
                    t = g->tensors[node.alias_of];
                    t.last_use = lu;
                    break;
                }
                PoolParams& pp = node.pool;
                pp.D = in0->D; pp.H = in0->H; pp.W = in0->W;
                const int in[3] = {in0->D, in0->H, in0->W};
                const bool z_done = g->ops[d.inputs[0]].d.op == TB_OP_CONV3D && g->ops[d.inputs[0]].conv.fuse_zpool;
                int out[3];
                for (int a = 0; a < 3; ++a) {
                    pp.k[a] = d.kernel[a];
                    pp.s[a] = d.stride[a] > 0 ? d.stride[a] : d.kernel[a];
                    if (a == 0 && z_done) pp.k[a] = pp.s[a] = 1;     // the producing conv's epilogue pooled z already
                    TB_REQUIRE(pp.k[a] >= 1, "pool size must be positive");
                    if (d.pad_same) {
                        int after;
                        same_pads(in[a], pp.k[a], pp.s[a], &out[a], &pp.pad0[a], &after);
                    } else {
                        out[a] = (in[a] - pp.k[a]) / pp.s[a] + 1;
                        pp.pad0[a] = 0;
                    }
                    TB_REQUIRE(out[a] >= 1, "pool window larger than input");
                }
                pp.Do = out[0]; pp.Ho = out[1]; pp.Wo = out[2];
                pp.is_avg = d.pool_kind;
                t.D = out[0]; t.H = out[1]; t.W = out[2]; t.C = in0->C;
                // the CPV writer reads 8-channel vectors: the pool input must store >= round_up(C,16) channels
                if (!in0->cpv && !in0->padvol && !in0->wfold && (in0->fmt == FMT_SPLIT || in0->C % 16 == 0))
                    decide_cpv(t, i, ops, n_ops);
                g->launches += 1;
                break;
            }
            case TB_OP_AFFINE: {
                TB_REQUIRE(d.n_inputs == 1, "affine takes one input");
                t.D = in0->D; t.H = in0->H; t.W = in0->W; t.C = in0->C;
                int rc = upload_vec(d.scale, t.C, 1.f, &node.d_scale);
                if (rc) return rc;
                rc = upload_vec(d.shift, t.C, 0.f, &node.d_shift);
                if (rc) return rc;
                g->launches += 1;
                break;
            }
            case TB_OP_GPOOL:
                TB_REQUIRE(d.n_inputs == 1, "global pool takes one input");
                t.D = t.H = t.W = 1; t.C = in0->C;
                g->launches += 1;
                break;
            case TB_OP_SOFTMAX:
                TB_REQUIRE(d.n_inputs == 1, "softmax takes one input");
                TB_REQUIRE(in0->pix_per_frame() == 1, "softmax is supported on (n, C) tensors only");
                t.D = t.H = t.W = 1; t.C = in0->C;
                g->launches += 1;
                break;
            case TB_OP_CONCAT: {
                TB_REQUIRE(d.n_inputs >= 1, "concat needs inputs");
                t.D = in0->D; t.H = in0->H; t.W = in0->W; t.C = 0;
                for (int k = 0; k < d.n_inputs; ++k) {
                    const TensorInfo& s = g->tensors[d.inputs[k]];
                    TB_REQUIRE(s.D == t.D && s.H == t.H && s.W == t.W, "concat: spatial dims differ");
                    t.C += s.C;
                }
                g->launches += d.n_inputs + 1;
                break;
            }
            case TB_OP_ADD: {
                TB_REQUIRE(d.n_inputs == 2, "add takes two inputs");
                const TensorInfo& s = g->tensors[d.inputs[1]];
                TB_REQUIRE(s.D == in0->D && s.H == in0->H && s.W == in0->W && s.C == in0->C,
                           "add: shapes differ");
                t.D = in0->D; t.H = in0->H; t.W = in0->W; t.C = in0->C;
                g->launches += 1;
                break;
            }
            default:
                TB_REQUIRE(false, "unknown op kind");
        }
        t.c_pad = t.fmt == FMT_SPLIT ? round_up(t.C, 16) : t.C;
        if (t.wfold || t.padvol) { t.fmt = FMT_SPLIT; t.c_pad = 8; }
        if (t.cpv) { t.fmt = FMT_SPLIT; t.c_pad = round_up(t.C, 16); }
        // pointers are only valid during graph_create
        node.d.kernel_w = node.d.bias = node.d.scale = node.d.shift = nullptr;
    }
    // Zero-copy Concatenate: fp32 inputs of a Concatenate become channel-slice views of its output buffer
    if (!getenv("TIMED_B200_NO_ZEROCOPY_CONCAT")) {
        auto readers_ok = [&](int tix) {
            for (int a = 0; a < n_ops; ++a)
                for (int b = 0; b < ops[a].n_inputs; ++b)
                    if (ops[a].inputs[b] == tix && ops[a].op == TB_OP_CONV3D) return false;
            return true;
        };
        for (int k = 1; k < n_ops; ++k) {
            if (ops[k].op != TB_OP_CONCAT || g->tensors[k].fmt != FMT_F32 || g->tensors[k].C % 8 != 0) continue;
            int off = 0;
            for (int b = 0; b < ops[k].n_inputs; ++b) {
                const int a = ops[k].inputs[b];
                TensorInfo& ta = g->tensors[a];
                const int kind = ops[a].op;
                const bool producer_ok = kind == TB_OP_CONV3D || kind == TB_OP_POOL3D || kind == TB_OP_AFFINE ||
                                         kind == TB_OP_CONCAT || kind == TB_OP_ADD;
                bool dup = false;                                  // the same tensor listed twice cannot live at two offsets
                for (int b2 = 0; b2 < b; ++b2) dup = dup || ops[k].inputs[b2] == a;
                if (producer_ok && !dup && ta.fmt == FMT_F32 && ta.view_of < 0 && !ta.cpv && !ta.padvol && !ta.wfold &&
                    g->ops[a].alias_of < 0 && off % 8 == 0 && ta.C % 8 == 0 && readers_ok(a) &&
                    !(kind == TB_OP_CONV3D && (g->ops[a].conv.fuse_pool || g->ops[a].conv.fuse_zpool))) {
                    ta.view_of = k;
                    ta.view_off = off;
                }
                off += ta.C;
            }
        }
        for (int i = 0; i < n_ops; ++i) {
            TensorInfo& t = g->tensors[i];
            if (t.view_of < 0) continue;
            int root = i, c0 = 0;
            while (g->tensors[root].view_of >= 0) { c0 += g->tensors[root].view_off; root = g->tensors[root].view_of; }
            t.view_root = root;
            t.view_c0 = c0;
            g->tensors[root].last_use = std::max(g->tensors[root].last_use, t.last_use);
        }
        for (int k = 1; k < n_ops; ++k) {                       // launches saved: one copy per in-place slice
            if (ops[k].op != TB_OP_CONCAT) continue;
            for (int b = 0; b < ops[k].n_inputs; ++b)
                if (g->tensors[ops[k].inputs[b]].view_of == k) g->launches -= 1;
            if (g->tensors[k].fmt != FMT_SPLIT) g->launches -= 1;   // (the +1 counted the channel-pad zeroing of split outputs)
        }
    }
    // Fused network head: ... -> GlobalPooling (j) -> Softmax (k = graph output), each read once.  The pooling op then
    // writes the probabilities itself (head_pool_softmax_kernel); when it reads a tap-to-N conv (TIMED's 20-class head)
    // that conv's col2im launch does gather + epilogue + pooling + softmax and both later ops are skipped.
    if (n_ops >= 3 && ops[n_ops - 1].op == TB_OP_SOFTMAX && !getenv("TIMED_B200_NO_HEADFUSE")) {
        const int k = n_ops - 1, j = ops[k].inputs[0];
        auto readers = [&](int t) {
            int n = 0;
            for (int a = 0; a < n_ops; ++a)
                for (int b = 0; b < ops[a].n_inputs; ++b) n += ops[a].inputs[b] == t;
            return n;
        };
        if (ops[j].op == TB_OP_GPOOL && readers(j) == 1 && g->tensors[j].fmt == FMT_F32) {
            g->ops[j].pool_softmax = true;
            g->ops[k].skip = true;
            g->launches -= 1;
            const int i = ops[j].inputs[0];
            ConvPlan& c = g->ops[i].conv;
            const size_t smem = (static_cast<size_t>(c.Do) * c.Ho * c.Wo * c.cout + c.cout) * sizeof(float);
            // linear conv -> average pooling: collapse to box sums + one dense GEMM (ConvPlan::gap_collapse)
            const TensorInfo& hin = g->tensors[ops[i].inputs[0]];
            if (getenv("TIMED_B200_VERBOSE"))
                fprintf(stderr, "head: op %d conv=%d readers=%d act=%d/%d pool_kind=%d k=%d,%d,%d slab=%d thin=%d thinz=%d wfold=%d "
                        "in fmt=%d cpv=%d padvol=%d wfold=%d view_of=%d\n", i, ops[i].op == TB_OP_CONV3D, readers(i), ops[i].act1,
                        ops[i].act2, ops[j].pool_kind, c.kd, c.kh, c.kw, c.slab, c.thin, c.thinz, c.wfold, hin.fmt, hin.cpv,
                        hin.padvol, hin.wfold, hin.view_of);
            if (ops[i].op == TB_OP_CONV3D && readers(i) == 1 && ops[i].act1 == ACT_NONE && ops[i].act2 == ACT_NONE &&
                ops[j].pool_kind == 1 && c.kd <= 3 && c.kh <= 3 && c.kw <= 3 && c.kd * c.kh * c.kw > 1 && !c.slab && !c.thin &&
                !c.thinz && !c.wfold && hin.fmt == FMT_SPLIT && !hin.cpv && !hin.padvol && !hin.wfold &&
                hin.view_of < 0 && !getenv("TIMED_B200_NO_GAPFOLD")) {
                tb_op_desc d2 = ops[i];
                d2.kernel[0] = d2.kernel[1] = d2.kernel[2] = 1;
                d2.pad_same = 0;
                const int k_dense = c.kd * c.kh * c.kw * c.cin;             // DHWIO read as (taps * cin, cout)
                ConvPlan* dn = new ConvPlan;
                dn->precise = c.precise;
                const int rc = conv_plan_create(*dn, d2, 1, 1, 1, k_dense, round_up(k_dense, 16));
                if (rc) { delete dn; return rc; }
                c.dense = dn;
                c.gap_collapse = c.gap_softmax = true;
                cudaFree(c.d_w);                                            // the per-voxel conv's packed weights are not used
                c.d_w = nullptr;
                g->launches += 3 - ((c.tap2n && !c.c2i) ? 2 : 1);                       // box sums, dense GEMM, softmax
                g->ops[j].pool_softmax = false;
                g->ops[j].skip = true;
                g->launches -= 1;
            } else
            if (ops[i].op == TB_OP_CONV3D && c.tap2n && readers(i) == 1 && c.cout % 4 == 0 && c.z_ld % 4 == 0 &&
                smem <= kHeadSmemMax && g->tensors[i].fmt == FMT_F32) {
                if (c.c2i) g->launches += 1;                                // the fused head takes the Z matrix route
                c.fuse_head = true;
                c.head_is_avg = ops[j].pool_kind;
                g->ops[j].pool_softmax = false;
                g->ops[j].skip = true;
                g->launches -= 1;
            }
        }
    }
    // DenseNet pre-activation: AFFINE(+ReLU) (a) of an fp32 tensor, read only by a 1x1x1 conv (j): the conv applies it in
    // its operand path (ConvPlan::xform, xform_conv.cuh), the AFFINE launch and its split-plane output disappear
    if (!getenv("TIMED_B200_NO_XFORM")) {
        for (int j = 1; j < n_ops; ++j) {
            if (ops[j].op != TB_OP_CONV3D || g->ops[j].skip) continue;
            const int a = ops[j].inputs[0];
            if (ops[a].op != TB_OP_AFFINE || g->ops[a].skip || g->ops[a].alias_of >= 0) continue;
            const tb_op_desc& ad = g->ops[a].d;
            if (ad.act1 != ACT_NONE || ad.act2 != ACT_RELU) continue;
            int n_readers = 0;
            for (int q = 0; q < n_ops; ++q)
                for (int b = 0; b < ops[q].n_inputs; ++b) n_readers += ops[q].inputs[b] == a;
            const int src = ops[a].inputs[0];
            const TensorInfo& ts = g->tensors[src];
            ConvPlan& c = g->ops[j].conv;
            if (n_readers != 1 || a == n_ops - 1 || ts.fmt != FMT_F32 || ts.cpv || ts.padvol || ts.wfold ||
                g->tensors[a].view_of >= 0 || !xform_eligible(c, ts.C))
                continue;
            const int root = ts.view_of >= 0 ? ts.view_root : src;
            if ((ts.view_of >= 0 ? g->tensors[root].C : ts.C) % 4 != 0 || (ts.view_of >= 0 && ts.view_c0 % 4 != 0)) continue;
            c.xform = true;
            c.xf_src = src;
            c.xf_scale = g->ops[a].d_scale;
            c.xf_shift = g->ops[a].d_shift;
            g->ops[a].skip = true;
            g->launches -= 1;
            // the conv now reads the AFFINE's input: keep it (and the buffer it is a slice of) alive until then
            g->tensors[src].last_use = std::max(g->tensors[src].last_use, j);
            g->tensors[root].last_use = std::max(g->tensors[root].last_use, j);
        }
    }
    // conv inputs need 127 readable pixels past the last valid one (im2col column of 128)
    for (int i = 0; i < n_ops; ++i)
        if (ops[i].op == TB_OP_CONV3D) {
            TensorInfo& s = g->tensors[ops[i].inputs[0]];
            const ConvPlan& c = g->ops[i].conv;
            if (s.cpv) continue;
            const int owner = g->ops[ops[i].inputs[0]].alias_of;      // fused pooling op: the storage belongs to the conv
            // base pixels advance through Do*Ho*Wo per frame: express the slack in input pixels
            const int64_t out_ppf = static_cast<int64_t>(c.Do) * c.Ho * c.Wo;
            const int64_t frames = c.thin ? 1 : (128 + out_ppf - 1) / out_ppf;   // thin: spans overrun < 1 frame
            s.slack_pix = std::max<int64_t>(s.slack_pix, frames * s.pix_per_frame());
            if (owner >= 0) g->tensors[owner].slack_pix = std::max(g->tensors[owner].slack_pix, s.slack_pix);
        }
    const TensorInfo& last = g->tensors[n_ops - 1];
    TB_REQUIRE(last.pix_per_frame() == 1, "graph output must be (n, classes)");
    TB_REQUIRE(last.fmt == FMT_F32, "internal: output tensor must be fp32");
    g->n_classes = last.C;
    return 0;
}

template <typename T>
static void launch_input_convert_wfold(const void* x, const TensorInfo& t, int64_t n_frames, const TView& out,
                                       cudaStream_t s) {
    const int64_t rows = n_frames * t.D * t.H;
    input_convert_wfold_kernel<T><<<grid_for(rows * t.wf_pitch, 256), 256, 0, s>>>(
        static_cast<const T*>(x), rows, t.W, t.C, t.wf_lm, t.wf_pitch, out.hi, out.lo);
}

template <typename T>
static void launch_input_convert_padvol(const void* x, const TensorInfo& t, int64_t n_frames, const TView& out,
                                        cudaStream_t s) {
    const int64_t total = n_frames * t.stored_pix_per_frame();
    input_convert_padvol_kernel<T><<<grid_for(total, 256), 256, 0, s>>>(
        static_cast<const T*>(x), n_frames, t.D, t.H, t.W, t.C, t.pv_d0, t.pv_h0, t.pv_w0, t.pv_Dp, t.pv_Hp,
        t.pv_Wp, out.hi, out.lo);
    // The thin kernels pair a partner-less odd filter tap with the NEXT stored pixel under a zero weight: for the last
    // valid output of the last frame that pixel is the first one of the slack frame behind the volume, which nobody
    // writes -- and 0 x NaN is NaN, so a workspace that happens to hold NaN bit patterns there would poison one row.
    // Zero the head of the slack frame in both planes (the allocation holds at least one slack frame, TensorInfo::frames_alloc).
    const size_t head = static_cast<size_t>(std::min<int64_t>(64, t.stored_pix_per_frame())) * t.c_pad * sizeof(__nv_bfloat16);
    cudaMemsetAsync(out.hi + total * t.c_pad, 0, head, s);
    cudaMemsetAsync(out.lo + total * t.c_pad, 0, head, s);
}

template <typename T>
static void launch_input_convert(const void* x, int64_t n_pix, const TView& out, cudaStream_t s) {
    const int cw = out.fmt == FMT_SPLIT ? out.c_pad : out.c;
    input_convert_kernel<T><<<grid_for(n_pix * cw, 256), 256, 0, s>>>(static_cast<const T*>(x), n_pix, out);
}

static int graph_forward(tb_graph* g, const void* d_frames, int dtype, int64_t n_frames, void* ws,
                         size_t ws_bytes, float* d_probs, cudaStream_t s) {
    TB_REQUIRE(n_frames > 0, "n_frames must be positive");
    const Layout& L = get_layout(g, n_frames);
    TB_REQUIRE(ws_bytes >= L.total, "workspace too small (see timed_b200_graph_workspace_bytes)");
    TB_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "workspace must be 256-byte aligned");
    uint8_t* base = static_cast<uint8_t*>(ws);
    const int n_ops = static_cast<int>(g->ops.size());
    std::vector<cudaEvent_t>* tset = nullptr;
    if (g->timing && g->timing_used < tb_graph::kMaxTimedForwards) {
        if (g->timing_used == g->timing_sets.size()) {
            std::vector<cudaEvent_t> evs(n_ops + 1);
            for (auto& e : evs) TB_CHECK_CUDA(cudaEventCreate(&e));
            g->timing_sets.push_back(std::move(evs));
        }
        tset = &g->timing_sets[g->timing_used++];
        TB_CHECK_CUDA(cudaEventRecord((*tset)[0], s));
    }
    for (int i = 0; i < n_ops; ++i) {
        OpNode& node = g->ops[i];
        const tb_op_desc& d = node.d;
        const TensorInfo& t = g->tensors[i];
        TView out = tensor_view(g, L, i, base, n_frames);
        if (i == n_ops - 1 || node.pool_softmax || (d.op == TB_OP_CONV3D && (node.conv.fuse_head || node.conv.gap_collapse))) {
            // the graph output (or the op that computes it in a fused head) goes straight to the caller's buffer
            out.f32 = d_probs;
            out.ld = g->n_classes;
        }
        if (node.skip) {
            if (tset) TB_CHECK_CUDA(cudaEventRecord((*tset)[i + 1], s));
            continue;
        }
        const int64_t out_pix = n_frames * t.pix_per_frame();
        TView in0{};
        const TensorInfo* ti0 = nullptr;
        if (d.n_inputs > 0) {
            ti0 = &g->tensors[d.inputs[0]];
            in0 = tensor_view(g, L, d.inputs[0], base, n_frames);
        }
        const int cw = out.fmt == FMT_SPLIT ? out.c_pad : out.c;
        switch (d.op) {
            case TB_OP_INPUT:
                if (t.cpv) {
                    const CpvGeom cg = make_cpv_geom(t, n_frames);
                    const int grid = grid_for(cg.T * cg.n_chunks, 256);
                    uint4* oh = reinterpret_cast<uint4*>(out.hi);
                    uint4* ol = reinterpret_cast<uint4*>(out.lo);
                    if (dtype == TB_DTYPE_F32) input_convert_cpv_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(d_frames), n_frames, cg, oh, ol);
                    else if (dtype == TB_DTYPE_F64) input_convert_cpv_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double*>(d_frames), n_frames, cg, oh, ol);
                    else if (dtype == TB_DTYPE_U8) input_convert_cpv_kernel<uint8_t><<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(d_frames), n_frames, cg, oh, ol);
                    else if (dtype == TB_DTYPE_F16) input_convert_cpv_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half*>(d_frames), n_frames, cg, oh, ol);
                    else TB_REQUIRE(false, "unknown frames dtype");
                    break;
                }
                if (t.padvol) {
                    if (dtype == TB_DTYPE_F32) launch_input_convert_padvol<float>(d_frames, t, n_frames, out, s);
                    else if (dtype == TB_DTYPE_F64) launch_input_convert_padvol<double>(d_frames, t, n_frames, out, s);
                    else if (dtype == TB_DTYPE_U8) launch_input_convert_padvol<uint8_t>(d_frames, t, n_frames, out, s);
                    else if (dtype == TB_DTYPE_F16) launch_input_convert_padvol<__half>(d_frames, t, n_frames, out, s);
                    else TB_REQUIRE(false, "unknown frames dtype");
                    break;
                }
                if (t.wfold) {
                    if (dtype == TB_DTYPE_F32) launch_input_convert_wfold<float>(d_frames, t, n_frames, out, s);
                    else if (dtype == TB_DTYPE_F64) launch_input_convert_wfold<double>(d_frames, t, n_frames, out, s);
                    else if (dtype == TB_DTYPE_U8) launch_input_convert_wfold<uint8_t>(d_frames, t, n_frames, out, s);
                    else if (dtype == TB_DTYPE_F16) launch_input_convert_wfold<__half>(d_frames, t, n_frames, out, s);
                    else TB_REQUIRE(false, "unknown frames dtype");
                    break;
                }
                if (dtype == TB_DTYPE_F32) launch_input_convert<float>(d_frames, out_pix, out, s);
                else if (dtype == TB_DTYPE_F64) launch_input_convert<double>(d_frames, out_pix, out, s);
                else if (dtype == TB_DTYPE_U8) launch_input_convert<uint8_t>(d_frames, out_pix, out, s);
                else if (dtype == TB_DTYPE_F16) launch_input_convert<__half>(d_frames, out_pix, out, s);
                else TB_REQUIRE(false, "unknown frames dtype");
                break;
            case TB_OP_CONV3D: {
                if (node.conv.xform) {
                    const TView x = tensor_view(g, L, node.conv.xf_src, base, n_frames);
                    int rc = xform_launch(node.conv, x, n_frames, out, s);
                    if (rc) return rc;
                    break;
                }
                TB_REQUIRE(ti0->fmt == FMT_SPLIT, "internal: conv input must be split planes");
                if (node.conv.fuse_pool && t.cpv) {
                    const CpvGeom cg = make_cpv_geom(t, n_frames);
                    cpv_zero_margins_kernel<<<grid_for(cg.T * cg.n_chunks, 256), 256, 0, s>>>(
                        reinterpret_cast<uint4*>(out.hi), reinterpret_cast<uint4*>(out.lo), cg);
                }
                int rc = conv_launch(node.conv, base + L.offset[d.inputs[0]], ti0->frames_alloc(n_frames),
                                     n_frames, out, s, base + L.scratch_off, L.scratch_bytes, &t);
                if (rc) return rc;
                break;
            }
            case TB_OP_POOL3D: {
                if (node.alias_of >= 0) break;                 // ran inside the producing conv's epilogue
                if (t.cpv) {
                    const CpvGeom cg = make_cpv_geom(t, n_frames);
                    const int grid = grid_for(cg.T * cg.n_chunks, 256);
                    if (in0.fmt == FMT_SPLIT)
                        pool3d_cpv_kernel<FMT_SPLIT><<<grid, 256, 0, s>>>(
                            in0, reinterpret_cast<uint4*>(out.hi), reinterpret_cast<uint4*>(out.lo), cg, node.pool, n_frames);
                    else
                        pool3d_cpv_kernel<FMT_F32><<<grid, 256, 0, s>>>(
                            in0, reinterpret_cast<uint4*>(out.hi), reinterpret_cast<uint4*>(out.lo), cg, node.pool, n_frames);
                    break;
                }
                // 16-byte vector path when both sides store a multiple of 8 channels per pixel
                const int in_cw = in0.fmt == FMT_SPLIT ? in0.c_pad : in0.c;
                const bool vec = (cw % 8 == 0) && (in_cw % 8 == 0) && (in0.ld % 8 == 0) && (out.ld % 8 == 0) &&
                                 in_cw >= cw;
                if (vec) {
                    const int grid = grid_for(out_pix * (cw / 8), 256);
                    if (in0.fmt == FMT_SPLIT && out.fmt == FMT_SPLIT)
                        pool3d_vec8_kernel<FMT_SPLIT, FMT_SPLIT><<<grid, 256, 0, s>>>(in0, out, node.pool, n_frames);
                    else if (in0.fmt == FMT_F32 && out.fmt == FMT_SPLIT)
                        pool3d_vec8_kernel<FMT_F32, FMT_SPLIT><<<grid, 256, 0, s>>>(in0, out, node.pool, n_frames);
                    else if (in0.fmt == FMT_SPLIT && out.fmt == FMT_F32)
                        pool3d_vec8_kernel<FMT_SPLIT, FMT_F32><<<grid, 256, 0, s>>>(in0, out, node.pool, n_frames);
                    else
                        pool3d_vec8_kernel<FMT_F32, FMT_F32><<<grid, 256, 0, s>>>(in0, out, node.pool, n_frames);
                } else {
                    pool3d_kernel<<<grid_for(out_pix * cw, 256), 256, 0, s>>>(in0, out, node.pool, n_frames);
                }
                break;
            }
            case TB_OP_AFFINE: {
                const int in_cw = in0.fmt == FMT_SPLIT ? in0.c_pad : in0.c;
                const bool vec = cw % 8 == 0 && in_cw % 8 == 0 && in0.ld % 8 == 0 && out.ld % 8 == 0 &&
                                 in_cw >= round_up(out.c, 8);
                if (vec) {
                    const int grid = grid_for(out_pix * (cw / 8), 256);
#define TB_AFFINE_VEC(FI, FO)                                                                                     \
    affine_act_vec8_kernel<FI, FO><<<grid, 256, 0, s>>>(in0, out, out_pix, node.d_scale, node.d_shift, d.act1,     \
                                                        d.alpha1, d.act2, d.alpha2)
                    if (in0.fmt == FMT_SPLIT && out.fmt == FMT_SPLIT) TB_AFFINE_VEC(FMT_SPLIT, FMT_SPLIT);
                    else if (in0.fmt == FMT_F32 && out.fmt == FMT_SPLIT) TB_AFFINE_VEC(FMT_F32, FMT_SPLIT);
                    else if (in0.fmt == FMT_SPLIT && out.fmt == FMT_F32) TB_AFFINE_VEC(FMT_SPLIT, FMT_F32);
                    else TB_AFFINE_VEC(FMT_F32, FMT_F32);
#undef TB_AFFINE_VEC
                } else {
                    affine_act_kernel<<<grid_for(out_pix * cw, 256), 256, 0, s>>>(
                        in0, out, out_pix, node.d_scale, node.d_shift, d.act1, d.alpha1, d.act2, d.alpha2);
                }
                break;
            }
            case TB_OP_GPOOL:
                if (node.pool_softmax)
                    head_pool_softmax_kernel<<<static_cast<unsigned>(n_frames), 128, in0.c * sizeof(float), s>>>(
                        in0, static_cast<int>(ti0->pix_per_frame()), d.pool_kind, d_probs);
                else
                    gpool_kernel<<<static_cast<unsigned>(n_frames), 128, 0, s>>>(
                        in0, out, static_cast<int>(ti0->pix_per_frame()), d.pool_kind);
                break;
            case TB_OP_SOFTMAX:
                softmax_kernel<<<static_cast<unsigned>((n_frames * 32 + 255) / 256), 256, 0, s>>>(in0, out, n_frames);
                break;
            case TB_OP_CONCAT: {
                int c_off = 0;
                for (int k = 0; k < d.n_inputs; ++k) {
                    const TensorInfo& ts = g->tensors[d.inputs[k]];
                    if (ts.view_of == i) {                     // zero-copy: the producer wrote this slice in place
                        c_off += ts.C;
                        continue;
                    }
                    TView src = tensor_view(g, L, d.inputs[k], base, n_frames);
                    if (src.c % 8 == 0 && c_off % 8 == 0 && src.ld % 8 == 0 && out.ld % 8 == 0) {
                        const int grid = grid_for(out_pix * (src.c / 8), 256);
                        if (src.fmt == FMT_SPLIT && out.fmt == FMT_SPLIT)
                            copy_channels_vec8_kernel<FMT_SPLIT, FMT_SPLIT><<<grid, 256, 0, s>>>(src, out, out_pix, c_off);
                        else if (src.fmt == FMT_F32 && out.fmt == FMT_SPLIT)
                            copy_channels_vec8_kernel<FMT_F32, FMT_SPLIT><<<grid, 256, 0, s>>>(src, out, out_pix, c_off);
                        else if (src.fmt == FMT_SPLIT && out.fmt == FMT_F32)
                            copy_channels_vec8_kernel<FMT_SPLIT, FMT_F32><<<grid, 256, 0, s>>>(src, out, out_pix, c_off);
                        else
                            copy_channels_vec8_kernel<FMT_F32, FMT_F32><<<grid, 256, 0, s>>>(src, out, out_pix, c_off);
                    } else {
                        copy_channels_kernel<<<grid_for(out_pix * src.c, 256), 256, 0, s>>>(src, out, out_pix, c_off);
                    }
                    c_off += ts.C;
                }
                if (out.fmt == FMT_SPLIT && out.c_pad > out.c)
                    zero_pad_channels_kernel<<<grid_for(out_pix * (out.c_pad - out.c), 256), 256, 0, s>>>(out, out_pix);
                break;
            }
            case TB_OP_ADD: {
                TView in1 = tensor_view(g, L, d.inputs[1], base, n_frames);
                add_kernel<<<grid_for(out_pix * cw, 256), 256, 0, s>>>(in0, in1, out, out_pix);
                break;
            }
            default:
                TB_REQUIRE(false, "unknown op kind");
        }
        TB_CHECK_CUDA(cudaGetLastError());
        if (tset) TB_CHECK_CUDA(cudaEventRecord((*tset)[i + 1], s));
    }
    return 0;
}

static size_t dtype_size(int dtype) {
    return dtype == TB_DTYPE_F64 ? 8 : (dtype == TB_DTYPE_F32 ? 4 : (dtype == TB_DTYPE_F16 ? 2 : 1));
}

// Largest frame count one pass may take: every tensor's pixel count must stay below 2^31 (the
// kernels index GEMM rows with 32-bit integers).  Larger requests are run as several passes.
static int64_t max_frames_per_pass(const tb_graph* g) {
    int64_t max_ppf = 1;
    for (const auto& t : g->tensors) max_ppf = std::max<int64_t>(max_ppf, t.stored_pix_per_frame());
    return std::max<int64_t>(1, ((1ll << 31) - (1 << 20)) / max_ppf);
}

static void graph_free(tb_graph* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    for (auto& set : g->timing_sets)
        for (auto& e : set) cudaEventDestroy(e);
    for (auto& n : g->ops) {
        free_conv_plan(n.conv);
        cudaFree(n.d_scale);
        cudaFree(n.d_shift);
    }
    for (int i = 0; i < 2; ++i) {
        cudaFree(g->d_stage[i]);
        if (g->ev_copied[i]) cudaEventDestroy(g->ev_copied[i]);
        if (g->ev_consumed[i]) cudaEventDestroy(g->ev_consumed[i]);
    }
    cudaFree(g->d_ws);
    cudaFree(g->d_probs);
    if (g->s_copy) cudaStreamDestroy(g->s_copy);
    if (g->s_compute) cudaStreamDestroy(g->s_compute);
    delete g;
}


}  // namespace tb

// ================================================================================= C ABI
using namespace tb;

extern "C" {

int timed_b200_abi_version(void) { return TIMED_B200_ABI_VERSION; }

const char* timed_b200_last_error(void) { return g_last_error.c_str(); }

int timed_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int timed_b200_graph_create(const tb_op_desc* ops, int32_t n_ops, int32_t device, tb_graph** out) {
    TB_REQUIRE(ops && out, "null argument");
    *out = nullptr;
    if (timed_b200_device_count() <= device) {
        set_error("no CUDA device " + std::to_string(device) + " (libtimed_b200 has no CPU path)");
        return TB_ERR_NO_DEVICE;
    }
    TB_CHECK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    TB_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error(std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                  "; libtimed_b200 is built for sm_100a (B200) only");
        return TB_ERR_NO_DEVICE;
    }
    int rc = load_driver_fns();
    if (rc) return rc;
    tb_graph* g = new tb_graph();
    g->device = device;
    rc = graph_build(g, ops, n_ops);
    if (rc) {
        graph_free(g);
        return rc;
    }
    *out = g;
    return TB_OK;
}

void timed_b200_graph_destroy(tb_graph* g) { graph_free(g); }

int timed_b200_graph_info(const tb_graph* g, int32_t* n_classes, double* flops_per_frame,
                          int32_t* n_kernel_launches_per_forward) {
    TB_REQUIRE(g, "null graph");
    if (n_classes) *n_classes = g->n_classes;
    if (flops_per_frame) *flops_per_frame = g->flops;
    if (n_kernel_launches_per_forward) *n_kernel_launches_per_forward = g->launches;
    return TB_OK;
}

int timed_b200_graph_op_count(const tb_graph* g, int32_t* n_ops) {
    TB_REQUIRE(g && n_ops, "null argument");
    *n_ops = static_cast<int32_t>(g->ops.size());
    return TB_OK;
}

int timed_b200_graph_op_kernel(const tb_graph* g, int32_t op, int64_t n_frames, char* buf, int32_t buflen) {
    TB_REQUIRE(g && buf && buflen > 0, "null argument");
    TB_REQUIRE(op >= 0 && op < static_cast<int32_t>(g->ops.size()) && n_frames > 0, "op index / frame count out of range");
    const OpNode& node = g->ops[op];
    std::string name;
    switch (node.d.op) {
        case TB_OP_CONV3D: {
            const ConvPlan& c = node.conv;
            if (c.gap_collapse) name = "gap_boxsum_kernel+conv_umma_kernel(dense over taps*cin)+softmax_kernel";
            else if (c.slab) name = "slab_conv_kernel";
            else if (c.thinz) name = c.fuse_pool ? "thinz_conv_kernel(+maxpool)" : c.fuse_zpool ? "thinz_conv_kernel(+z-maxpool)" : "thinz_conv_kernel";
            else if (c.thin) name = "thin_conv_kernel";
            else if (c.xform) name = "bnrelu_conv1x1_kernel(BatchNorm-ReLU in the operand path)";
            else {
                ConvPlan::Config cfg;
                const int rc = choose_config(c, n_frames * c.Mo_d() * c.Mo_h() * c.Mo_w(), &cfg);
                if (rc) return rc;
                name = cfg.pair ? "conv_pair_kernel" : (cfg.cluster2 ? "conv_umma_kernel(cluster2)" : "conv_umma_kernel");
                double valid = 1.0;
                if (vox_tiles(c, cfg, n_frames, &valid)) {
                    char buf[96];
                    snprintf(buf, sizeof(buf), "(voxel-stationary tiles, valid taps %.3f)", valid);
                    name += buf;
                }
                if (c.tap2n) name += c.fuse_head ? "+head_col2im_pool_softmax_kernel" : c.c2i_active() ? "(col2im over kw in the epilogue)" : "+col2im_kernel";
            }
            break;
        }
        case TB_OP_INPUT: name = g->tensors[op].cpv ? "input_convert_cpv_kernel" : g->tensors[op].padvol ? "input_convert_padvol_kernel" : "input_convert_kernel"; break;
        case TB_OP_POOL3D: name = node.alias_of >= 0 ? "(fused into the producing conv)" : g->tensors[op].cpv ? "pool3d_cpv_kernel" : "pool3d_vec8_kernel"; break;
        case TB_OP_AFFINE: name = node.skip ? "(applied in the operand path of the 1x1 conv that reads it)" : "affine_act_kernel"; break;
        case TB_OP_GPOOL: name = node.skip ? "(fused into the head conv's launch)" : node.pool_softmax ? "head_pool_softmax_kernel" : "gpool_kernel"; break;
        case TB_OP_SOFTMAX: name = node.skip ? "(fused into the pooling launch)" : "softmax_kernel"; break;
        case TB_OP_CONCAT: {
            bool all_views = true;
            for (int b = 0; b < node.d.n_inputs; ++b) all_views = all_views && g->tensors[node.d.inputs[b]].view_of == op;
            name = all_views ? "(zero-copy: producers write the channel slices)" : "copy_channels_kernel";
            break;
        }
        case TB_OP_ADD: name = "add_kernel"; break;
        default: name = "?";
    }
    std::snprintf(buf, static_cast<size_t>(buflen), "%s", name.c_str());
    return TB_OK;
}

int timed_b200_graph_set_precise(tb_graph* g, int32_t precise) {
    TB_REQUIRE(g, "null graph");
    for (auto& n : g->ops) n.conv.precise = precise != 0;
    g->layouts.clear();        // (tile configurations are chosen per launch; nothing else is cached per mode)
    return TB_OK;
}

int timed_b200_graph_set_timing(tb_graph* g, int32_t enabled) {
    TB_REQUIRE(g, "null graph");
    g->timing = enabled != 0;
    g->timing_used = 0;
    return TB_OK;
}

int timed_b200_graph_read_op_times(tb_graph* g, float* ms_per_op, int32_t* op_kinds,
                                   double* op_flops_per_frame, int32_t* n_forwards) {
    TB_REQUIRE(g && ms_per_op && n_forwards, "null argument");
    TB_CHECK_CUDA(cudaSetDevice(g->device));
    const int n_ops = static_cast<int>(g->ops.size());
    for (int i = 0; i < n_ops; ++i) {
        ms_per_op[i] = 0.f;
        if (op_kinds) op_kinds[i] = g->ops[i].d.op;
        if (op_flops_per_frame)
            op_flops_per_frame[i] = g->ops[i].d.op == TB_OP_CONV3D ? g->ops[i].conv.flops_per_frame() : 0.0;
    }
    for (size_t f = 0; f < g->timing_used; ++f) {
        auto& set = g->timing_sets[f];
        TB_CHECK_CUDA(cudaEventSynchronize(set[n_ops]));
        for (int i = 0; i < n_ops; ++i) {
            float ms = 0.f;
            TB_CHECK_CUDA(cudaEventElapsedTime(&ms, set[i], set[i + 1]));
            ms_per_op[i] += ms;
        }
    }
    *n_forwards = static_cast<int32_t>(g->timing_used);
    g->timing_used = 0;
    return TB_OK;
}

int timed_b200_graph_workspace_bytes(const tb_graph* g, int64_t n_frames, size_t* out) {
    TB_REQUIRE(g && out, "null argument");
    TB_REQUIRE(n_frames > 0, "n_frames must be positive");
    *out = get_layout(const_cast<tb_graph*>(g), std::min(n_frames, max_frames_per_pass(g))).total;
    return TB_OK;
}

int timed_b200_graph_forward(tb_graph* g, const void* d_frames, int32_t frames_dtype, int64_t n_frames,
                             void* d_workspace, size_t workspace_bytes, float* d_probs,
                             void* cuda_stream) {
    TB_REQUIRE(g && d_frames && d_workspace && d_probs, "null argument");
    TB_REQUIRE(n_frames > 0, "n_frames must be positive");
    TB_CHECK_CUDA(cudaSetDevice(g->device));
    const int64_t pass = max_frames_per_pass(g);
    const TensorInfo& tin = g->tensors[0];
    const size_t frame_bytes = static_cast<size_t>(tin.pix_per_frame()) * tin.C * dtype_size(frames_dtype);
    for (int64_t f0 = 0; f0 < n_frames; f0 += pass) {
        int rc = graph_forward(g, static_cast<const uint8_t*>(d_frames) + f0 * frame_bytes, frames_dtype,
                               std::min(pass, n_frames - f0), d_workspace, workspace_bytes,
                               d_probs + f0 * g->n_classes, static_cast<cudaStream_t>(cuda_stream));
        if (rc) return rc;
    }
    return TB_OK;
}

int timed_b200_graph_predict_host(tb_graph* g, const void* h_frames, int32_t frames_dtype,
                                  int64_t n_frames, float* h_probs, int64_t max_chunk_frames) {
    TB_REQUIRE(g && h_frames && h_probs, "null argument");
    TB_REQUIRE(n_frames > 0, "n_frames must be positive");
    TB_CHECK_CUDA(cudaSetDevice(g->device));
    const int64_t chunk = std::min<int64_t>(std::min<int64_t>(n_frames, max_chunk_frames > 0 ? max_chunk_frames : 1024),
                                            max_frames_per_pass(g));
    const TensorInfo& tin = g->tensors[0];
    const size_t frame_bytes = static_cast<size_t>(tin.pix_per_frame()) * tin.C * dtype_size(frames_dtype);
    // (re)size library-owned staging
    if (!g->s_copy) {
        TB_CHECK_CUDA(cudaStreamCreateWithFlags(&g->s_copy, cudaStreamNonBlocking));
        TB_CHECK_CUDA(cudaStreamCreateWithFlags(&g->s_compute, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            TB_CHECK_CUDA(cudaEventCreateWithFlags(&g->ev_copied[i], cudaEventDisableTiming));
            TB_CHECK_CUDA(cudaEventCreateWithFlags(&g->ev_consumed[i], cudaEventDisableTiming));
        }
    }
    if (g->stage_bytes < chunk * frame_bytes) {
        for (int i = 0; i < 2; ++i) {
            cudaFree(g->d_stage[i]);
            g->d_stage[i] = nullptr;
            TB_CHECK_CUDA(cudaMalloc(&g->d_stage[i], chunk * frame_bytes));
        }
        g->stage_bytes = chunk * frame_bytes;
    }
    const size_t ws_need = get_layout(g, chunk).total;
    if (g->ws_bytes < ws_need) {
        cudaFree(g->d_ws);
        g->d_ws = nullptr;
        TB_CHECK_CUDA(cudaMalloc(&g->d_ws, ws_need));
        g->ws_bytes = ws_need;
    }
    const size_t probs_need = static_cast<size_t>(n_frames) * g->n_classes * sizeof(float);
    if (g->probs_bytes < probs_need) {
        cudaFree(g->d_probs);
        g->d_probs = nullptr;
        TB_CHECK_CUDA(cudaMalloc(&g->d_probs, probs_need));
        g->probs_bytes = probs_need;
    }
    // double-buffered: H2D of chunk i+1 overlaps the forward of chunk i
    const uint8_t* src = static_cast<const uint8_t*>(h_frames);
    int it = 0;
    // Uniform passes of min(chunk, 512) frames.  Exposed are the first pass's copy and the last pass's forward, so passes
    // should be small; 512 = two 256-frame blocks of the voxel-stationary conv tiles keeps the kernels efficient.  Measured
    // with TIMED-20 at 214 k frames/s on the device (tools/gpu_e2e_chunks.sh, float32 / float16 / uint8 host frames):
    // uniform 512: 191.5 / 197.9 / 204.4 k frames/s; uniform 256: 190.0 / 196.1 / 199.1; uniform 1024: 179.6 / 195.0 / 201.8;
    // the 1.5x ramp this replaces (128 -> 1024, tuned when a forward took 1.45x the copy of its frames): 183.9 / 196.4 / 195.1.
    // A pass never exceeds `chunk`: the staging buffers and the workspace are sized for it, and the caller's batch_size
    // bound holds for every pass.
    int64_t cur = std::min<int64_t>(chunk, 512);
    if (const char* e = getenv("TIMED_B200_PASS_FRAMES")) cur = std::max<int64_t>(1, std::min<int64_t>(chunk, atoll(e)));   // A/B
    for (int64_t f0 = 0, nf = 0; f0 < n_frames; f0 += nf, ++it) {
        const int b = it & 1;
        nf = std::min(cur, n_frames - f0);
        TB_REQUIRE(nf > 0 && nf <= chunk, "internal: predict_host chunk exceeds the staged size");
        g->last_passes = it + 1;
        g->last_max_pass = it == 0 ? nf : std::max(g->last_max_pass, nf);
        if (it >= 2) TB_CHECK_CUDA(cudaStreamWaitEvent(g->s_copy, g->ev_consumed[b], 0));
        TB_CHECK_CUDA(cudaMemcpyAsync(g->d_stage[b], src + f0 * frame_bytes, nf * frame_bytes,
                                      cudaMemcpyHostToDevice, g->s_copy));
        TB_CHECK_CUDA(cudaEventRecord(g->ev_copied[b], g->s_copy));
        TB_CHECK_CUDA(cudaStreamWaitEvent(g->s_compute, g->ev_copied[b], 0));
        // a shorter tail chunk uses its own (smaller) layout inside the same workspace
        int rc = graph_forward(g, g->d_stage[b], frames_dtype, nf, g->d_ws, g->ws_bytes,
                               g->d_probs + f0 * g->n_classes, g->s_compute);
        if (rc) return rc;
        TB_CHECK_CUDA(cudaEventRecord(g->ev_consumed[b], g->s_compute));
    }
    TB_CHECK_CUDA(cudaMemcpyAsync(h_probs, g->d_probs, probs_need, cudaMemcpyDeviceToHost, g->s_compute));
    TB_CHECK_CUDA(cudaStreamSynchronize(g->s_compute));
    return TB_OK;
}

int timed_b200_graph_predict_stats(const tb_graph* g, int64_t* n_passes, int64_t* max_pass_frames) {
    TB_REQUIRE(g, "null graph");
    if (n_passes) *n_passes = g->last_passes;
    if (max_pass_frames) *max_pass_frames = g->last_max_pass;
    return TB_OK;
}

int timed_b200_conv3d_fwd(const float* d_x, int64_t n, int32_t D, int32_t H, int32_t W, int32_t c_in,
                          const tb_op_desc* conv, int32_t device, float* d_y) {
    TB_REQUIRE(d_x && conv && d_y, "null argument");
    TB_REQUIRE(conv->op == TB_OP_CONV3D, "conv descriptor expected");
    tb_op_desc ops[2];
    std::memset(ops, 0, sizeof(ops));
    ops[0].op = TB_OP_INPUT;
    ops[0].kernel[0] = D; ops[0].kernel[1] = H; ops[0].kernel[2] = W;
    ops[0].c_out = c_in;
    ops[1] = *conv;
    ops[1].n_inputs = 1;
    ops[1].inputs[0] = 0;
    // graph_build wants an (n, classes) output; run the two ops by hand instead
    if (timed_b200_device_count() <= device) {
        set_error("no CUDA device (libtimed_b200 has no CPU path)");
        return TB_ERR_NO_DEVICE;
    }
    TB_CHECK_CUDA(cudaSetDevice(device));
    int rc = load_driver_fns();
    if (rc) return rc;
    TensorInfo tin;
    tin.D = D; tin.H = H; tin.W = W; tin.C = c_in;
    tin.fmt = FMT_SPLIT;
    tin.c_pad = round_up(c_in, 16);
    if (c_in <= 8 && !getenv("TIMED_B200_NO_THIN") &&
        thin_fits(conv->kernel[0], conv->kernel[1], conv->kernel[2], conv->c_out)) {
        const int in[3] = {D, H, W};
        int pb[3] = {0, 0, 0}, pa[3] = {0, 0, 0};
        for (int a = 0; a < 3 && conv->pad_same; ++a) {
            int o;
            same_pads(in[a], conv->kernel[a], 1, &o, &pb[a], &pa[a]);
        }
        tin.padvol = true;
        tin.pv_d0 = pb[0]; tin.pv_h0 = pb[1]; tin.pv_w0 = pb[2];
        tin.pv_Dp = pb[0] + D + pa[0]; tin.pv_Hp = pb[1] + H + pa[1]; tin.pv_Wp = pb[2] + W + pa[2];
        tin.c_pad = 8;
    } else if (c_in > 8) {
        tb_op_desc two[2] = {ops[0], ops[1]};
        tin.cpv = false;
        decide_cpv(tin, 0, two, 2);
    } else if (c_in <= 8 && conv->kernel[2] <= 8 && !getenv("TIMED_B200_NO_WFOLD")) {
        const int kw = conv->kernel[2], pad0 = conv->pad_same ? (kw - 1) / 2 : 0, kwin = kw <= 4 ? 4 : 8;
        tin.wfold = true;
        tin.wf_lm = pad0;
        tin.wf_pitch = pad0 + W + (kwin - 1 - pad0);
        tin.c_pad = 8;
    }
    ConvPlan plan;
    plan.precise = getenv("TIMED_B200_FAST_ACCUM") == nullptr;   // test hook: corrections into the main accumulator
    rc = conv_plan_create(plan, ops[1], D, H, W, c_in, tin.c_pad, &tin);
    if (rc) { free_conv_plan(plan); return rc; }
    const int64_t out_ppf = static_cast<int64_t>(plan.Do) * plan.Ho * plan.Wo;
    tin.slack_pix = static_cast<int>((plan.thin ? 1 : (128 + out_ppf - 1) / out_ppf) * tin.pix_per_frame());
    void* d_in = nullptr;
    cudaError_t e = cudaMalloc(&d_in, tin.bytes(n));
    if (e != cudaSuccess) { free_conv_plan(plan); TB_CHECK_CUDA(e); }
    cudaMemset(d_in, 0, tin.bytes(n));
    TView vin = make_view(tin, static_cast<uint8_t*>(d_in), n);
    if (tin.cpv) {
        const CpvGeom cg = make_cpv_geom(tin, n);
        input_convert_cpv_kernel<float><<<grid_for(cg.T * cg.n_chunks, 256), 256>>>(
            d_x, n, cg, reinterpret_cast<uint4*>(vin.hi), reinterpret_cast<uint4*>(vin.lo));
    } else if (tin.padvol) launch_input_convert_padvol<float>(d_x, tin, n, vin, nullptr);
    else if (tin.wfold) launch_input_convert_wfold<float>(d_x, tin, n, vin, nullptr);
    else launch_input_convert<float>(d_x, n * tin.pix_per_frame(), vin, nullptr);
    TView vout{};
    vout.fmt = FMT_F32;
    vout.f32 = d_y;
    vout.c = plan.cout;
    vout.c_pad = plan.cout;
    vout.ld = plan.cout;
    void* d_scratch = nullptr;
    const size_t sb = conv_scratch_bytes(plan, n);
    if (sb && cudaMalloc(&d_scratch, sb) != cudaSuccess) {
        cudaFree(d_in);
        free_conv_plan(plan);
        set_error("cudaMalloc(scratch) failed");
        return TB_ERR_CUDA;
    }
    rc = conv_launch(plan, d_in, tin.frames_alloc(n), n, vout, nullptr, d_scratch, sb);
    cudaError_t se = cudaDeviceSynchronize();
    cudaFree(d_scratch);
    cudaFree(d_in);
    free_conv_plan(plan);
    if (rc) return rc;
    TB_CHECK_CUDA(se);
    return TB_OK;
}

// ---------------------------------------------------------------------------------- sampler
int timed_b200_apply_temperature(const double* d_probs_in, int64_t n_rows, int32_t n_cls, double t,
                                 double* d_probs_out, void* cuda_stream) {
    TB_REQUIRE(d_probs_in && d_probs_out, "null argument");
    TB_REQUIRE(n_rows > 0 && n_cls > 0, "empty probability matrix");
    TB_REQUIRE(t != 0.0, "temperature must be non-zero");
    const double inv_t = 1.0 / t;
    if (n_cls > 32 && n_cls <= 1024) {       // wide rows: a warp per row (kernels.cuh)
        const int wpb = 4;
        const unsigned grid = static_cast<unsigned>(std::min<int64_t>((n_rows + wpb - 1) / wpb, 148 * 16));
        temperature_wide_kernel<<<grid, 32 * wpb, static_cast<size_t>(wpb) * n_cls * sizeof(double),
                                  static_cast<cudaStream_t>(cuda_stream)>>>(d_probs_in, n_rows, n_cls, inv_t, d_probs_out);
        TB_CHECK_CUDA(cudaGetLastError());
        return TB_OK;
    }
    temperature_kernel<<<static_cast<unsigned>((n_rows + 127) / 128), 128, 0,
                         static_cast<cudaStream_t>(cuda_stream)>>>(d_probs_in, n_rows, n_cls, inv_t, d_probs_out);
    TB_CHECK_CUDA(cudaGetLastError());
    return TB_OK;
}

int timed_b200_cumsum_rows(const double* d_probs, int64_t n_rows, int32_t n_cls, double* d_cdf,
                           void* cuda_stream) {
    TB_REQUIRE(d_probs && d_cdf, "null argument");
    TB_REQUIRE(n_rows > 0 && n_cls > 0, "empty probability matrix");
    if (n_cls > 32 && n_cls <= 1024) {
        const int wpb = 4;
        const unsigned grid = static_cast<unsigned>(std::min<int64_t>((n_rows + wpb - 1) / wpb, 148 * 16));
        cumsum_rows_wide_kernel<<<grid, 32 * wpb, static_cast<size_t>(wpb) * n_cls * sizeof(double),
                                  static_cast<cudaStream_t>(cuda_stream)>>>(d_probs, n_rows, n_cls, d_cdf);
        TB_CHECK_CUDA(cudaGetLastError());
        return TB_OK;
    }
    cumsum_rows_kernel<<<static_cast<unsigned>((n_rows + 127) / 128), 128, 0,
                         static_cast<cudaStream_t>(cuda_stream)>>>(d_probs, n_rows, n_cls, d_cdf);
    TB_CHECK_CUDA(cudaGetLastError());
    return TB_OK;
}

int timed_b200_sample(const double* d_cdf, int64_t n_res, int32_t n_cls, int64_t n_samples,
                      int64_t first_sample, uint64_t seed, uint64_t stream_id, const double* d_uniforms,
                      const uint8_t* d_cls_to_letter, uint8_t* d_seqs, int32_t* d_idx, void* cuda_stream) {
    TB_REQUIRE(d_cdf && d_cls_to_letter && d_seqs, "null argument");
    TB_REQUIRE(n_res > 0 && n_samples > 0, "nothing to sample");
    TB_REQUIRE(n_cls > 0 && n_cls <= 512, "n_cls must be in [1,512]");
    TB_REQUIRE((reinterpret_cast<uintptr_t>(d_seqs) & 3) == 0, "d_seqs must be 4-byte aligned");
    TB_REQUIRE(!d_idx || (reinterpret_cast<uintptr_t>(d_idx) & 15) == 0, "d_idx must be 16-byte aligned");
    const int64_t words = (n_res * n_samples + 3) / 4;
    sample_kernel<<<grid_for(words, 256), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        d_cdf, n_res, n_cls, n_samples, first_sample, seed, stream_id, d_uniforms, d_cls_to_letter, d_seqs,
        d_idx);
    TB_CHECK_CUDA(cudaGetLastError());
    return TB_OK;
}

int timed_b200_sample_chains(const double* d_cdf, const int64_t* d_row_off, const int64_t* d_seq_off, int32_t n_chains,
                             int64_t total_seq_bytes, int32_t n_cls, int64_t n_samples, int64_t first_sample, uint64_t seed,
                             uint64_t stream_id0, const uint8_t* d_cls_to_letter, uint8_t* d_seqs, void* cuda_stream) {
    TB_REQUIRE(d_cdf && d_row_off && d_seq_off && d_cls_to_letter && d_seqs, "null argument");
    TB_REQUIRE(n_chains > 0 && n_samples > 0 && total_seq_bytes > 0, "nothing to sample");
    TB_REQUIRE(n_cls > 0 && n_cls <= 512, "n_cls must be in [1,512]");
    TB_REQUIRE((reinterpret_cast<uintptr_t>(d_seqs) & 3) == 0 && (total_seq_bytes & 3) == 0,
               "d_seqs and the chain offsets must be 4-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    if (!getenv("TIMED_B200_SAMPLER_GATHER")) {
        // tiled kernel (kernels.cuh): CDF rows of a residue tile in shared memory, lanes = samples of one residue
        const int tr = n_cls <= 64 ? 32 : 8;
        const size_t smem = static_cast<size_t>(tr) * n_cls * sizeof(double) + static_cast<size_t>(kSampTileS) * (tr / 4 + 1) * 4;
        // rows <= letters / samples; every chain adds at most one partial tile
        const int64_t row_tiles_ub = total_seq_bytes / n_samples / tr + n_chains + 1;
        const int64_t s_tiles = (n_samples + kSampTileS - 1) / kSampTileS;
        TB_REQUIRE(row_tiles_ub < (1ll << 31) && s_tiles < 65536, "too many sampler tiles per launch");
        int32_t* d_tile_off = nullptr;
        TB_CHECK_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_tile_off), (static_cast<size_t>(n_chains) + 1) * sizeof(int32_t), s));
        sample_tile_prefix_kernel<<<1, 32, 0, s>>>(d_row_off, n_chains, tr, d_tile_off);
        const dim3 grid(static_cast<unsigned>(row_tiles_ub), static_cast<unsigned>(s_tiles));
        if (tr == 32) {
            static bool attr = false;
            if (!attr) { TB_CHECK_CUDA(cudaFuncSetAttribute(sample_tiled_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); attr = true; }
            sample_tiled_kernel<32><<<grid, kSampTileS, smem, s>>>(d_cdf, d_row_off, d_seq_off, d_tile_off, n_chains, n_cls, n_samples,
                                                                   first_sample, seed, stream_id0, d_cls_to_letter, d_seqs);
        } else {
            static bool attr = false;
            if (!attr) { TB_CHECK_CUDA(cudaFuncSetAttribute(sample_tiled_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); attr = true; }
            sample_tiled_kernel<8><<<grid, kSampTileS, smem, s>>>(d_cdf, d_row_off, d_seq_off, d_tile_off, n_chains, n_cls, n_samples,
                                                                  first_sample, seed, stream_id0, d_cls_to_letter, d_seqs);
        }
        TB_CHECK_CUDA(cudaGetLastError());
        TB_CHECK_CUDA(cudaFreeAsync(d_tile_off, s));
        return TB_OK;
    }
    sample_chains_kernel<<<grid_for(total_seq_bytes / 4, 256), 256, 0, s>>>(
        d_cdf, d_row_off, d_seq_off, n_chains, n_cls, n_samples, first_sample, seed, stream_id0, d_cls_to_letter,
        d_seqs);
    TB_CHECK_CUDA(cudaGetLastError());
    return TB_OK;
}

int timed_b200_sample_uniforms(int64_t n_res, int64_t n_samples, int64_t first_sample, uint64_t seed,
                               uint64_t stream_id, double* d_out, void* cuda_stream) {
    TB_REQUIRE(d_out, "null argument");
    TB_REQUIRE(n_res > 0 && n_samples > 0, "nothing to draw");
    sample_uniforms_kernel<<<grid_for(n_res * n_samples, 256), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        n_res, n_samples, first_sample, seed, stream_id, d_out);
    TB_CHECK_CUDA(cudaGetLastError());
    return TB_OK;
}

int timed_b200_argmax_fp16(const float* d_probs, int64_t n, int32_t n_cls, int32_t* d_idx, void* cuda_stream) {
    TB_REQUIRE(d_probs && d_idx, "null argument");
    TB_REQUIRE(n > 0 && n_cls > 0, "empty matrix");
    argmax_fp16_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0,
                         static_cast<cudaStream_t>(cuda_stream)>>>(d_probs, n, n_cls, d_idx);
    TB_CHECK_CUDA(cudaGetLastError());
    return TB_OK;
}

int timed_b200_consensus_fp16(const float* d_probs, int64_t n_rows, int32_t n_cls, const int64_t* d_first_row,
                              const int32_t* d_n_states, const int32_t* d_n_res, const int64_t* d_out_row0,
                              int32_t n_groups, int64_t n_out_rows, uint16_t* d_consensus_fp16, int32_t* d_idx,
                              void* cuda_stream) {
    TB_REQUIRE(d_probs && d_first_row && d_n_states && d_n_res && d_out_row0 && d_consensus_fp16 && d_idx, "null argument");
    TB_REQUIRE(n_rows > 0 && n_cls > 0 && n_groups > 0 && n_out_rows > 0, "empty input");
    consensus_fp16_kernel<<<static_cast<unsigned>((n_out_rows * 32 + 255) / 256), 256, 0,
                            static_cast<cudaStream_t>(cuda_stream)>>>(
        d_probs, d_first_row, d_n_states, d_n_res, d_out_row0, n_groups, n_out_rows, n_cls,
        reinterpret_cast<__half*>(d_consensus_fp16), d_idx);
    TB_CHECK_CUDA(cudaGetLastError());
    return TB_OK;
}

int timed_b200_seq_metrics(const uint8_t* d_seqs, int64_t n_seqs, int64_t n_res, const int8_t* d_letter_lut,
                           const double* d_tables, int32_t n_table_doubles, double* d_out, void* cuda_stream) {
    TB_REQUIRE(d_seqs && d_letter_lut && d_tables && d_out, "null argument");
    TB_REQUIRE(n_seqs > 0 && n_res > 0, "empty input");
    TB_REQUIRE(n_table_doubles >= 63, "metric table too short");
    seq_metrics_kernel<<<static_cast<unsigned>((n_seqs + 7) / 8), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        d_seqs, n_seqs, n_res, d_letter_lut, d_tables, d_out);
    TB_CHECK_CUDA(cudaGetLastError());
    return TB_OK;
}

// ---------------------------------------------------------------------------------- voxeliser (8(f)-1)
int timed_b200_voxelise(const float* d_atoms_xyzs, const int32_t* d_atom_channel, const int32_t* d_atom_residue,
                        const int32_t* d_atom_is_cb, int64_t n_atoms, const float* d_res_frame, const float* d_res_property,
                        const int32_t* d_res_index, const int32_t* d_res_atom_range, int64_t res_first, int64_t n_res,
                        int32_t voxels_per_side, float voxel_edge, int32_t n_channels,
                        int32_t as_gaussian, int32_t encode_cb, const float* ideal_cb_xyz_sigma, int32_t cb_channel,
                        int32_t property_channel, int32_t* d_scratch, void* d_frames, int32_t frames_dtype, void* cuda_stream) {
    TB_REQUIRE(d_atoms_xyzs && d_atom_channel && d_atom_residue && d_atom_is_cb && d_res_frame && d_scratch && d_frames,
               "null argument");
    TB_REQUIRE(n_atoms > 0 && n_res > 0 && res_first >= 0, "nothing to voxelise");
    TB_REQUIRE(voxels_per_side > 0 && (voxels_per_side & 1) && voxel_edge > 0.f, "voxels per side must be odd, edge positive");
    TB_REQUIRE(n_channels > 0 && cb_channel < n_channels && property_channel < n_channels, "channel index out of range");
    TB_REQUIRE(!encode_cb || (ideal_cb_xyz_sigma && cb_channel >= 0), "encode_cb needs the ideal C-beta and its channel");
    TB_REQUIRE(frames_dtype == TB_DTYPE_F32 || frames_dtype == TB_DTYPE_F16 || (frames_dtype == TB_DTYPE_U8 && !as_gaussian),
               "frames must be float32 / float16 (or uint8 for boolean voxels)");
    TB_REQUIRE((reinterpret_cast<uintptr_t>(d_atoms_xyzs) & 15) == 0, "atom table must be 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const int64_t per_frame = static_cast<int64_t>(voxels_per_side) * voxels_per_side * voxels_per_side * n_channels;
    const int64_t n = n_res * per_frame;
    TB_CHECK_CUDA(cudaMemsetAsync(d_scratch, 0, static_cast<size_t>(n) * sizeof(int32_t), s));
    VoxeliseParams p;
    p.atoms = reinterpret_cast<const float4*>(d_atoms_xyzs);
    p.atom_channel = d_atom_channel; p.atom_residue = d_atom_residue; p.atom_is_cb = d_atom_is_cb;
    p.n_atoms = n_atoms;
    p.res_frame = d_res_frame; p.res_property = d_res_property; p.res_index = d_res_index; p.res_first = res_first;
    p.res_atom_range = d_res_atom_range;
    p.V = voxels_per_side; p.inv_edge = 1.0f / voxel_edge; p.C = n_channels;
    p.gaussian = as_gaussian; p.encode_cb = encode_cb;
    p.cb_x = encode_cb ? ideal_cb_xyz_sigma[0] : 0.f; p.cb_y = encode_cb ? ideal_cb_xyz_sigma[1] : 0.f;
    p.cb_z = encode_cb ? ideal_cb_xyz_sigma[2] : 0.f; p.cb_sigma = encode_cb ? ideal_cb_xyz_sigma[3] : 1.f;
    p.cb_channel = cb_channel; p.property_channel = property_channel;
    p.scratch = d_scratch;
    voxelise_kernel<<<static_cast<unsigned>(n_res), 256, 0, s>>>(p);
    TB_CHECK_CUDA(cudaGetLastError());
    const int grid = grid_for(n, 256);
    if (frames_dtype == TB_DTYPE_F32) voxelise_finalize_kernel<float><<<grid, 256, 0, s>>>(d_scratch, n, !as_gaussian, static_cast<float*>(d_frames));
    else if (frames_dtype == TB_DTYPE_F16) voxelise_finalize_kernel<__half><<<grid, 256, 0, s>>>(d_scratch, n, !as_gaussian, static_cast<__half*>(d_frames));
    else voxelise_finalize_kernel<uint8_t><<<grid, 256, 0, s>>>(d_scratch, n, 1, static_cast<uint8_t*>(d_frames));
    TB_CHECK_CUDA(cudaGetLastError());
    return TB_OK;
}

// ---- host-side index of many HDF5 frame datasets (hdf5_index.cuh): no device work
int timed_b200_hdf5_frame_index(const uint8_t* file_base, int64_t file_len, int64_t base_addr, int64_t n,
                                const int64_t* obj_addr, const uint8_t* tmpl_space, int32_t space_len,
                                const uint8_t* tmpl_type, int32_t type_len, const uint8_t* tmpl_filters, int32_t filters_len,
                                const uint8_t* tmpl_attr_hdr, int32_t attr_hdr_len, int32_t attr_data_len, int32_t rank,
                                int64_t* chunk_off, int64_t* chunk_size, uint8_t* attr_data, int32_t* status,
                                int32_t n_threads) {
    TB_REQUIRE(file_base && obj_addr && tmpl_space && tmpl_type && tmpl_attr_hdr && chunk_off && chunk_size && attr_data && status,
               "null argument");
    TB_REQUIRE(file_len > 0 && n >= 0 && space_len > 0 && type_len > 0 && filters_len >= 0 && (filters_len == 0 || tmpl_filters) &&
               attr_hdr_len > 0 && attr_data_len > 0 && rank >= 1 && rank <= 8, "bad template / sizes");
    H5Template t{tmpl_space, space_len, tmpl_type, type_len, tmpl_filters, filters_len, tmpl_attr_hdr, attr_hdr_len, attr_data_len, rank};
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const int64_t i0 = next.fetch_add(256);
            if (i0 >= n) return;
            for (int64_t i = i0; i < std::min(n, i0 + 256); ++i)
                status[i] = h5_index_one(file_base, file_len, base_addr, obj_addr[i], t, chunk_off + i, chunk_size + i,
                                         attr_data + i * attr_data_len);
        }
    };
    const int nt = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(n_threads, (n + 255) / 256)));
    if (nt == 1) {
        worker();
    } else {
        std::vector<std::thread> th;
        for (int k = 0; k < nt; ++k) th.emplace_back(worker);
        for (auto& x : th) x.join();
    }
    return TB_OK;
}

// ---- device-side inflate of stored HDF5 frame chunks (inflate.cuh)
int timed_b200_inflate_device(const uint8_t* d_comp, int64_t n_streams, const int64_t* d_off, const int64_t* d_size,
                              int64_t out_bytes, void* d_out, int32_t* d_status, void* cuda_stream) {
    TB_REQUIRE(d_comp && d_off && d_size && d_out && d_status, "null argument");
    TB_REQUIRE(n_streams >= 0 && out_bytes > 0, "bad stream count / output size");
    if (n_streams == 0) return TB_OK;
    if (timed_b200_device_count() <= 0) {
        set_error("no CUDA device (libtimed_b200 has no CPU path)");
        return TB_ERR_NO_DEVICE;
    }
    const int64_t blocks = (n_streams + kInflateWarps - 1) / kInflateWarps;
    TB_REQUIRE(blocks < (1ll << 31), "too many streams per launch");
    inflate_streams_kernel<<<static_cast<unsigned>(blocks), 32 * kInflateWarps, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        d_comp, d_off, d_size, n_streams, out_bytes, static_cast<uint8_t*>(d_out), d_status);
    TB_CHECK_CUDA(cudaGetLastError());
    return TB_OK;
}

// ---- host-side structure files (pdb_parse.cuh): no device work
int timed_b200_pdb_parse(const char* const* paths, int32_t n_paths, int32_t all_states, int32_t n_threads, tb_pdb_batch** out) {
    TB_REQUIRE(paths && out && n_paths >= 0, "null argument");
    tb_pdb_batch* b = new tb_pdb_batch;
    b->files.resize(static_cast<size_t>(n_paths));
    std::atomic<int32_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const int32_t i = next.fetch_add(1);
            if (i >= n_paths) return;
            pdb_parse_one(paths[i], i, all_states != 0, &b->files[i]);
        }
    };
    const int nt = std::max(1, std::min<int>(n_threads, n_paths));
    if (nt == 1) {
        worker();
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    *out = b;
    return TB_OK;
}

int timed_b200_pdb_sizes(const tb_pdb_batch* b, int64_t* n_states, int64_t* n_res, int64_t* n_atoms) {
    TB_REQUIRE(b && n_states && n_res && n_atoms, "null argument");
    int64_t s = 0, r = 0, a = 0;
    for (const auto& f : b->files)
        for (const auto& st : f.states) { ++s; r += static_cast<int64_t>(st.res.size()); a += static_cast<int64_t>(st.atoms.size()); }
    *n_states = s; *n_res = r; *n_atoms = a;
    return TB_OK;
}

int timed_b200_pdb_export(const tb_pdb_batch* b, int32_t* file_status, int32_t* state_file, int64_t* state_res_off,
                          int64_t* state_atom_off, int32_t* state_dup, char* res_chain, char* res_id, char* res_label,
                          uint8_t* res_has_bb, double* res_bb, double* atom_xyz, int32_t* atom_name, int32_t* atom_res) {
    TB_REQUIRE(b && file_status && state_file && state_res_off && state_atom_off && state_dup && res_chain && res_id &&
               res_label && res_has_bb && res_bb && atom_xyz && atom_name && atom_res, "null argument");
    int64_t s = 0, r = 0, a = 0;
    for (size_t fi = 0; fi < b->files.size(); ++fi) {
        const auto& f = b->files[fi];
        file_status[fi] = f.status;
        for (const auto& st : f.states) {
            state_file[s] = static_cast<int32_t>(fi);
            state_res_off[s] = r;
            state_atom_off[s] = a;
            state_dup[s] = st.n_dup;
            for (const auto& pr : st.res) {
                res_chain[r] = pr.chain;
                std::memcpy(res_id + 4 * r, pr.res_id, 4);
                std::memcpy(res_label + 3 * r, pr.label, 3);
                res_has_bb[r] = pr.has_bb;
                std::memcpy(res_bb + 9 * r, pr.bb, sizeof(pr.bb));
                ++r;
            }
            for (const auto& pa : st.atoms) {
                std::memcpy(atom_xyz + 3 * a, pa.xyz, sizeof(pa.xyz));
                atom_name[a] = pa.name;
                atom_res[a] = pa.res;
                ++a;
            }
            ++s;
        }
    }
    state_res_off[s] = r;
    state_atom_off[s] = a;
    return TB_OK;
}

void timed_b200_pdb_free(tb_pdb_batch* b) { delete b; }

// ---- host-side frame I/O helper: no device work, no Python: inflate + unshuffle + scatter + cast on host threads
int timed_b200_inflate_chunks(const uint8_t* file_base, int64_t n_chunks, const int64_t* src_off, const int64_t* src_size,
                              const int64_t* dst_frame, const int32_t* origin, int32_t rank, const int32_t* chunk_dims,
                              const int32_t* frame_dims, int32_t deflate, int32_t shuffle_elem_size, int32_t src_dtype,
                              int32_t dst_dtype, void* dst, int32_t n_threads) {
    TB_REQUIRE(file_base && src_off && src_size && dst_frame && origin && chunk_dims && frame_dims && dst, "null argument");
    TB_REQUIRE(rank >= 1 && rank <= 8 && n_chunks >= 0, "rank must be in [1,8]");
    TB_REQUIRE(src_dtype == TB_DTYPE_F32 || src_dtype == TB_DTYPE_F64 || src_dtype == TB_DTYPE_U8, "unknown source dtype");
    TB_REQUIRE(dst_dtype == TB_DTYPE_F32 || (dst_dtype == TB_DTYPE_U8 && src_dtype == TB_DTYPE_U8),
               "destination must be float32 (or uint8 for uint8 sources)");
    const size_t esz = dtype_size(src_dtype), dsz = dtype_size(dst_dtype);
    TB_REQUIRE(shuffle_elem_size == 0 || shuffle_elem_size == static_cast<int32_t>(esz), "shuffle element size mismatch");
    int64_t chunk_elems = 1, frame_elems = 1;
    for (int d = 0; d < rank; ++d) {
        TB_REQUIRE(chunk_dims[d] > 0 && frame_dims[d] > 0, "dims must be positive");
        chunk_elems *= chunk_dims[d];
        frame_elems *= frame_dims[d];
    }
    const size_t raw_bytes = static_cast<size_t>(chunk_elems) * esz;
    std::atomic<int64_t> next{0};
    std::atomic<int> failed{0};
    auto worker = [&]() {
        std::vector<uint8_t> raw(raw_bytes), tmp(shuffle_elem_size ? raw_bytes : 0);
        for (;;) {
            const int64_t c = next.fetch_add(1);
            if (c >= n_chunks || failed.load()) return;
            const uint8_t* src = file_base + src_off[c];
            if (deflate) {
                uLongf got = static_cast<uLongf>(raw_bytes);
                if (uncompress(raw.data(), &got, src, static_cast<uLong>(src_size[c])) != Z_OK || got != raw_bytes) {
                    failed.store(1);
                    return;
                }
            } else {
                if (static_cast<size_t>(src_size[c]) != raw_bytes) { failed.store(2); return; }
                std::memcpy(raw.data(), src, raw_bytes);
            }
            const uint8_t* data = raw.data();
            if (shuffle_elem_size) {              // HDF5 shuffle: byte k of every element stored contiguously
                for (size_t b = 0; b < esz; ++b)
                    for (int64_t e = 0; e < chunk_elems; ++e) tmp[e * esz + b] = raw[b * chunk_elems + e];
                data = tmp.data();
            }
            // scatter the chunk into its frame: runs along the last dimension, odometer over the others
            const int32_t* org = origin + c * rank;
            uint8_t* out = static_cast<uint8_t*>(dst) + static_cast<size_t>(dst_frame[c]) * frame_elems * dsz;
            const int last = rank - 1;
            const int64_t run = std::min<int64_t>(chunk_dims[last], static_cast<int64_t>(frame_dims[last]) - org[last]);
            if (run <= 0) continue;
            int32_t idx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (;;) {
                bool inside = true;
                int64_t src_e = 0, dst_e = 0;
                for (int d = 0; d < last; ++d) {
                    const int64_t g = static_cast<int64_t>(org[d]) + idx[d];
                    if (g >= frame_dims[d]) inside = false;
                    src_e = src_e * chunk_dims[d] + idx[d];
                    dst_e = dst_e * frame_dims[d] + g;
                }
                if (inside) {
                    src_e = src_e * chunk_dims[last];
                    dst_e = dst_e * frame_dims[last] + org[last];
                    if (src_dtype == TB_DTYPE_F64) {
                        const double* sp = reinterpret_cast<const double*>(data) + src_e;
                        float* dp = reinterpret_cast<float*>(out) + dst_e;
                        for (int64_t e = 0; e < run; ++e) dp[e] = static_cast<float>(sp[e]);
                    } else if (src_dtype == TB_DTYPE_F32) {
                        std::memcpy(reinterpret_cast<float*>(out) + dst_e, reinterpret_cast<const float*>(data) + src_e,
                                    static_cast<size_t>(run) * 4);
                    } else if (dst_dtype == TB_DTYPE_U8) {
                        std::memcpy(out + dst_e, data + src_e, static_cast<size_t>(run));
                    } else {
                        float* dp = reinterpret_cast<float*>(out) + dst_e;
                        for (int64_t e = 0; e < run; ++e) dp[e] = static_cast<float>(data[src_e + e]);
                    }
                }
                int d = last - 1;
                for (; d >= 0; --d) {
                    if (++idx[d] < chunk_dims[d]) break;
                    idx[d] = 0;
                }
                if (d < 0) break;
            }
        }
    };
    const int nt = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(n_threads > 0 ? n_threads : 1, n_chunks)));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    TB_REQUIRE(failed.load() != 1, "zlib failed to inflate a chunk (corrupt data or wrong chunk size)");
    TB_REQUIRE(failed.load() != 2, "stored chunk size does not match the chunk dimensions");
    return TB_OK;
}

// ---- host-side text writer: rows of "%.18e" numbers (numpy's savetxt default), formatted on host threads
int timed_b200_format_csv_e18(const void* data, int32_t dtype, int64_t rows, int64_t cols, char* out, int64_t out_cap,
                              int64_t* written, int32_t n_threads) {
    TB_REQUIRE(data && out && written, "null argument");
    TB_REQUIRE(dtype == TB_DTYPE_F32 || dtype == TB_DTYPE_F64, "float32 or float64 data expected");
    TB_REQUIRE(rows >= 0 && cols > 0, "bad shape");
    TB_REQUIRE(out_cap >= rows * cols * 26, "output buffer must hold 26 bytes per number");
    const int nt = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(n_threads > 0 ? n_threads : 1, rows)));
    std::vector<int64_t> used(nt, 0);
    std::vector<int64_t> first(nt + 1, 0);
    std::atomic<int> overflow{0};
    for (int t = 0; t <= nt; ++t) first[t] = rows * t / nt;
    // every thread formats its block of rows at the worst-case offset of that block, then the blocks are compacted
    auto worker = [&](int t) {
        char* p = out + first[t] * cols * 26;
        char* p0 = p;
        for (int64_t r = first[t]; r < first[t + 1]; ++r)
            for (int64_t c = 0; c < cols; ++c) {
                const double v = dtype == TB_DTYPE_F32 ? static_cast<double>(static_cast<const float*>(data)[r * cols + c])
                                                       : static_cast<const double*>(data)[r * cols + c];
                // "%.18e" is at most 27 characters (sign, 20 digits + point, e, sign, 3-digit exponent); 26 bytes per number
                // INCLUDING the separator are guaranteed by the caller only for |exponent| < 100 and the float32 / float16
                // data this writer is used for -- format into a local buffer and copy what fits the slot
                char tmp[40];
                int len = std::snprintf(tmp, sizeof(tmp), "%.18e", v);
                if (len > 25) { overflow.store(1); len = 25; }
                std::memcpy(p, tmp, static_cast<size_t>(len));
                p += len;
                *p++ = c + 1 == cols ? '\n' : ',';
            }
        used[t] = p - p0;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto& th : pool) th.join();
    int64_t total = used[0];
    for (int t = 1; t < nt; ++t) {
        std::memmove(out + total, out + first[t] * cols * 26, static_cast<size_t>(used[t]));
        total += used[t];
    }
    *written = total;
    TB_REQUIRE(!overflow.load(), "a value needs more than 25 characters in %.18e (negative with a 3-digit exponent)");
    return TB_OK;
}

// ---- host-side text reader: a comma-separated matrix of numbers (the {model}.csv that sample.py reads back,
// sample.py:32-34) parsed with strtod on host threads.  Call with out == NULL to get the shape first.
int timed_b200_parse_csv(const char* text, int64_t len, double* out, int64_t out_cap, int64_t* rows, int64_t* cols,
                         int32_t n_threads) {
    TB_REQUIRE(text && rows && cols && len >= 0, "null argument");
    std::vector<int64_t> starts;                       // non-empty lines
    for (int64_t p = 0; p < len;) {
        const char* nl = static_cast<const char*>(std::memchr(text + p, '\n', static_cast<size_t>(len - p)));
        const int64_t end = nl ? nl - text : len;
        int64_t e = end;
        while (e > p && (text[e - 1] == '\r' || text[e - 1] == ' ')) --e;
        if (e > p) starts.push_back(p);
        p = end + 1;
    }
    *rows = static_cast<int64_t>(starts.size());
    *cols = 0;
    if (starts.empty()) return TB_OK;
    {
        const char* nl = static_cast<const char*>(std::memchr(text + starts[0], '\n', static_cast<size_t>(len - starts[0])));
        const int64_t end = nl ? nl - text : len;
        int64_t c = 1;
        for (int64_t q = starts[0]; q < end; ++q) c += text[q] == ',';
        *cols = c;
    }
    if (!out) return TB_OK;
    TB_REQUIRE(out_cap >= *rows * *cols, "output buffer too small");
    const int64_t n_rows = *rows, n_cols = *cols;
    std::atomic<int> bad{0};
    const int nt = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(n_threads > 0 ? n_threads : 1, n_rows)));
    auto worker = [&](int t) {
        std::string line;
        for (int64_t r = n_rows * t / nt; r < n_rows * (t + 1) / nt && !bad.load(); ++r) {
            const char* nl = static_cast<const char*>(std::memchr(text + starts[r], '\n', static_cast<size_t>(len - starts[r])));
            const int64_t end = nl ? nl - text : len;
            line.assign(text + starts[r], static_cast<size_t>(end - starts[r]));     // NUL-terminated copy for strtod
            const char* p = line.c_str();
            for (int64_t c = 0; c < n_cols; ++c) {
                char* q = nullptr;
                const double v = std::strtod(p, &q);
                if (q == p) { bad.store(1); return; }
                out[r * n_cols + c] = v;
                p = q;
                while (*p == ' ' || *p == '\r') ++p;
                if (c + 1 < n_cols) {
                    if (*p != ',') { bad.store(1); return; }
                    ++p;
                } else if (*p != '\0') {
                    bad.store(1);
                    return;
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto& th : pool) th.join();
    TB_REQUIRE(!bad.load(), "not a rectangular comma-separated matrix of numbers");
    return TB_OK;
}

}  // extern "C"
