// Bandwidth-side kernels of the inference graph (everything that is not a contraction) and
// the Monte-Carlo sampler kernels.  All HBM-bound: coalesced channel-fastest access, 16-byte
// vectors on the bf16 split planes, grids sized from the element count.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "conv_umma.cuh"

namespace tb {

// NDHWC view: `pix` rows of `c` logical channels, row pitch `ld` elements.
// FMT_SPLIT rows hold c_pad (multiple of 16) stored channels; channels >= c are zero.
struct TView {
    int32_t fmt;
    float* f32;
    __nv_bfloat16* hi;
    __nv_bfloat16* lo;
    int64_t ld;
    int32_t c;
    int32_t c_pad;
};

#if defined(__CUDACC__)

__device__ __forceinline__ float tv_load(const TView& t, int64_t pix, int ch) {
    const int64_t o = pix * t.ld + ch;
    if (t.fmt == FMT_F32) return t.f32[o];
    return __bfloat162float(t.hi[o]) + __bfloat162float(t.lo[o]);
}
__device__ __forceinline__ void tv_store(const TView& t, int64_t pix, int ch, float v) {
    const int64_t o = pix * t.ld + ch;
    if (t.fmt == FMT_F32) {
        t.f32[o] = v;
    } else {
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        t.hi[o] = h;
        t.lo[o] = l;
    }
}

// ------------------------------------------------------------------ input frames -> tensor
// x: (n_pix, c) of T (float/double/uint8) -> out view; the value is first cast to float32
// (Keras casts X to the InputLayer dtype), pad channels are zeroed.
template <typename T>
__global__ void input_convert_kernel(const T* __restrict__ x, int64_t n_pix, TView out) {
    const int cw = out.fmt == FMT_SPLIT ? out.c_pad : out.c;
    const int64_t total = n_pix * cw;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t pix = i / cw;
        const int ch = static_cast<int>(i - pix * cw);
        const float v = ch < out.c ? static_cast<float>(x[pix * out.c + ch]) : 0.0f;
        tv_store(out, pix, ch, v);
    }
}

// W-folded input layout (api.cu TensorInfo::wfold): rows = (frame, d, h); each row stores `pitch`
// pixels of 8 channels: `lm` zero pixels, the W real pixels (channels >= c zero), zeros to the pitch.
// One thread per stored pixel: 16-byte store per plane.
template <typename T>
__global__ void input_convert_wfold_kernel(const T* __restrict__ x, int64_t n_rows, int W, int c, int lm,
                                           int pitch, __nv_bfloat16* __restrict__ hi,
                                           __nv_bfloat16* __restrict__ lo) {
    const int64_t total = n_rows * pitch;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t row = i / pitch;
        const int w = static_cast<int>(i - row * pitch) - lm;
        uint32_t h4[4] = {0, 0, 0, 0}, l4[4] = {0, 0, 0, 0};
        if (w >= 0 && w < W) {
            const T* src = x + (row * W + w) * c;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = k < c ? static_cast<float>(src[k]) : 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                __nv_bfloat16 h0, l0, h1, l1;
                split_bf16(v[2 * k], h0, l0);
                split_bf16(v[2 * k + 1], h1, l1);
                h4[k] = pack_bf16x2(h0, h1);
                l4[k] = pack_bf16x2(l0, l1);
            }
        }
        reinterpret_cast<uint4*>(hi)[i] = make_uint4(h4[0], h4[1], h4[2], h4[3]);
        reinterpret_cast<uint4*>(lo)[i] = make_uint4(l4[0], l4[1], l4[2], l4[3]);
    }
}

// Padded-volume input layout (api.cu TensorInfo::padvol): per frame Dp x Hp x Wp stored pixels of 8
// channels, the real D x H x W block at offset (d0, h0, w0), everything else zero.  One thread per
// stored pixel: one 16-byte store per plane.
template <typename T>
__global__ void input_convert_padvol_kernel(const T* __restrict__ x, int64_t n_frames, int D, int H, int W, int c,
                                            int d0, int h0, int w0, int Dp, int Hp, int Wp,
                                            __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const int64_t total = n_frames * Dp * Hp * Wp;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        int64_t t = i;
        const int w = static_cast<int>(t % Wp) - w0; t /= Wp;
        const int h = static_cast<int>(t % Hp) - h0; t /= Hp;
        const int d = static_cast<int>(t % Dp) - d0;
        const int64_t nf = t / Dp;
        uint32_t h4[4] = {0, 0, 0, 0}, l4[4] = {0, 0, 0, 0};
        if (w >= 0 && w < W && h >= 0 && h < H && d >= 0 && d < D) {
            const T* src = x + ((((nf * D + d) * H + h) * W + w) * static_cast<int64_t>(c));
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = k < c ? static_cast<float>(src[k]) : 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                __nv_bfloat16 a0, b0, a1, b1;
                split_bf16(v[2 * k], a0, b0);
                split_bf16(v[2 * k + 1], a1, b1);
                h4[k] = pack_bf16x2(a0, a1);
                l4[k] = pack_bf16x2(b0, b1);
            }
        }
        reinterpret_cast<uint4*>(hi)[i] = make_uint4(h4[0], h4[1], h4[2], h4[3]);
        reinterpret_cast<uint4*>(lo)[i] = make_uint4(l4[0], l4[1], l4[2], l4[3]);
    }
}

// ------------------------------------------------------------------ pooling (TF semantics)
struct PoolParams {
    int32_t D, H, W, Do, Ho, Wo;
    int32_t k[3], s[3], pad0[3];
    int32_t is_avg;
};

// scalar path: one thread per (out pixel, channel)
__global__ void pool3d_kernel(TView in, TView out, PoolParams pp, int64_t n_frames) {
    const int cw = out.fmt == FMT_SPLIT ? out.c_pad : out.c;
    const int64_t opix = static_cast<int64_t>(pp.Do) * pp.Ho * pp.Wo;
    const int64_t total = n_frames * opix * cw;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int ch = static_cast<int>(i % cw);
        int64_t t = i / cw;
        const int q = static_cast<int>(t % pp.Wo); t /= pp.Wo;
        const int p = static_cast<int>(t % pp.Ho); t /= pp.Ho;
        const int z = static_cast<int>(t % pp.Do);
        const int64_t nf = t / pp.Do;
        float acc = pp.is_avg ? 0.0f : -INFINITY;
        int cnt = 0;
        if (ch < out.c) {
            for (int a = 0; a < pp.k[0]; ++a) {
                const int d = z * pp.s[0] - pp.pad0[0] + a;
                if (d < 0 || d >= pp.D) continue;
                for (int b = 0; b < pp.k[1]; ++b) {
                    const int h = p * pp.s[1] - pp.pad0[1] + b;
                    if (h < 0 || h >= pp.H) continue;
                    for (int c = 0; c < pp.k[2]; ++c) {
                        const int w = q * pp.s[2] - pp.pad0[2] + c;
                        if (w < 0 || w >= pp.W) continue;
                        const int64_t ipix = ((nf * pp.D + d) * pp.H + h) * pp.W + w;
                        const float v = tv_load(in, ipix, ch);
                        acc = pp.is_avg ? acc + v : fmaxf(acc, v);
                        ++cnt;
                    }
                }
            }
            if (pp.is_avg) acc = acc / static_cast<float>(cnt);
        } else {
            acc = 0.0f;
        }
        const int64_t op = ((nf * pp.Do + z) * pp.Ho + p) * pp.Wo + q;
        tv_store(out, op, ch, acc);
    }
}

// vector path: one thread per (out pixel, 8 channels), 16-byte loads/stores; input and output
// formats are compile-time (fp32 rows or bf16 split planes).  Requires 8 | channels stored.
template <int IN_FMT>
__device__ __forceinline__ void load8(const TView& t, int64_t elem_off, float (&v)[8]) {
    if constexpr (IN_FMT == FMT_F32) {
        const float4 a = *reinterpret_cast<const float4*>(t.f32 + elem_off);
        const float4 b = *reinterpret_cast<const float4*>(t.f32 + elem_off + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        const uint4 vh = *reinterpret_cast<const uint4*>(t.hi + elem_off);
        const uint4 vl = *reinterpret_cast<const uint4*>(t.lo + elem_off);
        const uint32_t hw[4] = {vh.x, vh.y, vh.z, vh.w};
        const uint32_t lw[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {      // bf16 -> fp32 is a 16-bit shift
            v[2 * e] = __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
            v[2 * e + 1] = __uint_as_float(hw[e] & 0xFFFF0000u) + __uint_as_float(lw[e] & 0xFFFF0000u);
        }
    }
}
template <int OUT_FMT>
__device__ __forceinline__ void store8(const TView& t, int64_t elem_off, const float (&v)[8]) {
    if constexpr (OUT_FMT == FMT_F32) {
        *reinterpret_cast<float4*>(t.f32 + elem_off) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(t.f32 + elem_off + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        uint32_t ho[4], lo_[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(v[2 * e], h0, l0);
            split_bf16(v[2 * e + 1], h1, l1);
            ho[e] = pack_bf16x2(h0, h1);
            lo_[e] = pack_bf16x2(l0, l1);
        }
        *reinterpret_cast<uint4*>(t.hi + elem_off) = make_uint4(ho[0], ho[1], ho[2], ho[3]);
        *reinterpret_cast<uint4*>(t.lo + elem_off) = make_uint4(lo_[0], lo_[1], lo_[2], lo_[3]);
    }
}

template <int IN_FMT, int OUT_FMT>
__global__ void pool3d_vec8_kernel(TView in, TView out, PoolParams pp, int64_t n_frames) {
    const int groups = (OUT_FMT == FMT_SPLIT ? out.c_pad : out.c) / 8;
    const int64_t opix = static_cast<int64_t>(pp.Do) * pp.Ho * pp.Wo;
    const int64_t total = n_frames * opix * groups;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % groups);
        int64_t t = i / groups;
        const int q = static_cast<int>(t % pp.Wo); t /= pp.Wo;
        const int p = static_cast<int>(t % pp.Ho); t /= pp.Ho;
        const int z = static_cast<int>(t % pp.Do);
        const int64_t nf = t / pp.Do;
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = pp.is_avg ? 0.0f : -INFINITY;
        int cnt = 0;
        for (int a = 0; a < pp.k[0]; ++a) {
            const int d = z * pp.s[0] - pp.pad0[0] + a;
            if (d < 0 || d >= pp.D) continue;
            for (int b = 0; b < pp.k[1]; ++b) {
                const int h = p * pp.s[1] - pp.pad0[1] + b;
                if (h < 0 || h >= pp.H) continue;
                for (int c = 0; c < pp.k[2]; ++c) {
                    const int w = q * pp.s[2] - pp.pad0[2] + c;
                    if (w < 0 || w >= pp.W) continue;
                    float v[8];
                    load8<IN_FMT>(in, (((nf * pp.D + d) * pp.H + h) * pp.W + w) * in.ld + g * 8, v);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] = pp.is_avg ? acc[e] + v[e] : fmaxf(acc[e], v[e]);
                    ++cnt;
                }
            }
        }
        if (pp.is_avg) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] /= static_cast<float>(cnt);
        }
        store8<OUT_FMT>(out, ((((nf * pp.Do + z) * pp.Ho + p) * pp.Wo + q)) * out.ld + g * 8, acc);
    }
}

// ------------------------------------------------------------------ chunk-plane padded volume (CPV) writers
// Layout read by slab_conv_kernel: plane (hi | lo) x chunk of 8 channels x position x 8 bf16, positions =
// lead zeros, then every frame as a (Dp, Hp, Wp) box whose trailing margins are zero, then tail zeros.
struct CpvGeom {
    int64_t T, lead, n_pos;          // stored positions; leading zeros; n_frames * Dp*Hp*Wp
    int32_t D, H, W, Dp, Hp, Wp;
    int32_t n_chunks, c;
};

__device__ __forceinline__ void cpv_store(uint4* hi, uint4* lo, int64_t i, const float (&v)[8]) {
    uint32_t h4[4], l4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        __nv_bfloat16 a0, b0, a1, b1;
        split_bf16(v[2 * k], a0, b0);
        split_bf16(v[2 * k + 1], a1, b1);
        h4[k] = pack_bf16x2(a0, a1);
        l4[k] = pack_bf16x2(b0, b1);
    }
    hi[i] = make_uint4(h4[0], h4[1], h4[2], h4[3]);
    lo[i] = make_uint4(l4[0], l4[1], l4[2], l4[3]);
}

// position -> (frame, z, p, q); false for lead / tail / margin positions
__device__ __forceinline__ bool cpv_decode(const CpvGeom& g, int64_t t, int64_t& nf, int& z, int& p, int& q) {
    const int64_t u = t - g.lead;
    if (u < 0 || u >= g.n_pos) return false;
    const int64_t fpos = static_cast<int64_t>(g.Dp) * g.Hp * g.Wp;
    nf = u / fpos;
    int rem = static_cast<int>(u - nf * fpos);
    z = rem / (g.Hp * g.Wp);
    rem -= z * g.Hp * g.Wp;
    p = rem / g.Wp;
    q = rem - p * g.Wp;
    return z < g.D && p < g.H && q < g.W;
}

template <typename T>
__global__ void input_convert_cpv_kernel(const T* __restrict__ x, int64_t n_frames, CpvGeom g, uint4* __restrict__ hi,
                                         uint4* __restrict__ lo) {
    const int64_t total = g.T * g.n_chunks;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t t = i / g.n_chunks;            // chunk fastest: a pixel's chunks are read by adjacent threads
        const int chunk = static_cast<int>(i - t * g.n_chunks);
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int64_t nf;
        int z, p, q;
        if (cpv_decode(g, t, nf, z, p, q)) {
            const T* src = x + (((nf * g.D + z) * g.H + p) * g.W + q) * static_cast<int64_t>(g.c) + chunk * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (chunk * 8 + k < g.c) v[k] = static_cast<float>(src[k]);
        }
        cpv_store(hi, lo, chunk * g.T + t, v);
    }
}

// zero the lead / tail / margin positions of a CPV tensor whose interior is written by a fused conv+pool epilogue
__global__ void cpv_zero_margins_kernel(uint4* __restrict__ hi, uint4* __restrict__ lo, CpvGeom g) {
    const int64_t total = g.T * g.n_chunks;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int chunk = static_cast<int>(i / g.T);
        const int64_t t = i - chunk * g.T;
        int64_t nf;
        int z, p, q;
        if (!cpv_decode(g, t, nf, z, p, q)) {
            hi[i] = make_uint4(0, 0, 0, 0);
            lo[i] = make_uint4(0, 0, 0, 0);
        }
    }
}

// pooling straight into the CPV layout (input stores >= n_chunks*8 channels per pixel)
template <int IN_FMT>
__global__ void pool3d_cpv_kernel(TView in, uint4* __restrict__ hi, uint4* __restrict__ lo, CpvGeom g, PoolParams pp,
                                  int64_t n_frames) {
    const int64_t total = g.T * g.n_chunks;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t t = i / g.n_chunks;            // chunk fastest: a pixel's chunks are read by adjacent threads
        const int chunk = static_cast<int>(i - t * g.n_chunks);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int64_t nf;
        int z, p, q;
        if (cpv_decode(g, t, nf, z, p, q)) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = pp.is_avg ? 0.0f : -INFINITY;
            int cnt = 0;
            for (int a = 0; a < pp.k[0]; ++a) {
                const int d = z * pp.s[0] - pp.pad0[0] + a;
                if (d < 0 || d >= pp.D) continue;
                for (int b = 0; b < pp.k[1]; ++b) {
                    const int h = p * pp.s[1] - pp.pad0[1] + b;
                    if (h < 0 || h >= pp.H) continue;
                    for (int c = 0; c < pp.k[2]; ++c) {
                        const int w = q * pp.s[2] - pp.pad0[2] + c;
                        if (w < 0 || w >= pp.W) continue;
                        float v[8];
                        load8<IN_FMT>(in, (((nf * pp.D + d) * pp.H + h) * pp.W + w) * in.ld + chunk * 8, v);
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[e] = pp.is_avg ? acc[e] + v[e] : fmaxf(acc[e], v[e]);
                        ++cnt;
                    }
                }
            }
            if (pp.is_avg) {
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] /= static_cast<float>(cnt);
            }
        }
        cpv_store(hi, lo, chunk * g.T + t, acc);
    }
}

// ------------------------------------------------------------------ standalone BN / activation
__global__ void affine_act_kernel(TView in, TView out, int64_t n_pix, const float* __restrict__ scale,
                                  const float* __restrict__ shift, int act1, float alpha1, int act2,
                                  float alpha2) {
    const int cw = out.fmt == FMT_SPLIT ? out.c_pad : out.c;
    const int64_t total = n_pix * cw;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t pix = i / cw;
        const int ch = static_cast<int>(i - pix * cw);
        float v = 0.0f;
        if (ch < out.c) {
            v = apply_act(tv_load(in, pix, ch), act1, alpha1);
            v = fmaf(v, scale[ch], shift[ch]);
            v = apply_act(v, act2, alpha2);
        }
        tv_store(out, pix, ch, v);
    }
}

// 8 channels per thread (16/32-byte vectors): both sides store a multiple of 8 channels per pixel
template <int IN_FMT, int OUT_FMT>
__global__ void affine_act_vec8_kernel(TView in, TView out, int64_t n_pix, const float* __restrict__ scale,
                                       const float* __restrict__ shift, int act1, float alpha1, int act2,
                                       float alpha2) {
    const int groups = (OUT_FMT == FMT_SPLIT ? out.c_pad : out.c) / 8;
    const int64_t total = n_pix * groups;
    // (pixel, channel group) of this thread's items advance by a constant stride: one 64-bit division up front instead of
    // one per item (the DenseNet pre-activation passes run ~10^9 items per launch and were division-bound)
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t i0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    int64_t pix = i0 / groups;
    int g = static_cast<int>(i0 - pix * groups);
    const int64_t d_pix = stride / groups;
    const int d_g = static_cast<int>(stride - d_pix * groups);
    const bool plain_relu = act1 == ACT_NONE && act2 == ACT_RELU;        // BatchNorm -> ReLU, the DenseNet / ResNet case
    for (int64_t i = i0; i < total; i += stride) {
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (g * 8 < out.c) {
            load8<IN_FMT>(in, pix * in.ld + g * 8, v);
            if (plain_relu && g * 8 + 8 <= out.c) {
                const float4 s0 = *reinterpret_cast<const float4*>(scale + g * 8), s1 = *reinterpret_cast<const float4*>(scale + g * 8 + 4);
                const float4 h0 = *reinterpret_cast<const float4*>(shift + g * 8), h1 = *reinterpret_cast<const float4*>(shift + g * 8 + 4);
                const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaxf(fmaf(v[e], sc[e], sh[e]), 0.0f);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int ch = g * 8 + e;
                    if (ch < out.c) {
                        float x = apply_act(v[e], act1, alpha1);
                        x = fmaf(x, scale[ch], shift[ch]);
                        v[e] = apply_act(x, act2, alpha2);
                    } else {
                        v[e] = 0.0f;
                    }
                }
            }
        }
        store8<OUT_FMT>(out, pix * out.ld + g * 8, v);
        pix += d_pix;
        g += d_g;
        if (g >= groups) { g -= groups; ++pix; }
    }
}

template <int IN_FMT, int OUT_FMT>
__global__ void copy_channels_vec8_kernel(TView in, TView out, int64_t n_pix, int c_off) {
    const int groups = in.c / 8;
    const int64_t total = n_pix * groups;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t pix = i / groups;
        const int g = static_cast<int>(i - pix * groups);
        float v[8];
        load8<IN_FMT>(in, pix * in.ld + g * 8, v);
        store8<OUT_FMT>(out, pix * out.ld + c_off + g * 8, v);
    }
}

// ------------------------------------------------------------------ col2im of the tap-to-N convs
// Z: (input pixels, z_ld) fp32 with Z[p', tap*cout + co] = sum_c X[p', c] * W[tap, c, co].
// out[p, co] = act2(scale * act1(bias + sum_tap Z[p + tap - pad, tap, co]) + shift), taps summed in
// (kd, kh, kw) order (deterministic), out-of-volume taps skipped (= zero padding).
struct Col2imParams {
    int32_t Di, Hi, Wi, Do, Ho, Wo, kd, kh, kw, pd, ph, pw, cout, z_ld;
    int32_t act1, act2;
    float alpha1, alpha2;
};
__global__ void col2im_kernel(const float* __restrict__ Z, TView out, int64_t n_frames, Col2imParams cp,
                              const float* __restrict__ bias, const float* __restrict__ scale,
                              const float* __restrict__ shift) {
    const int cw = out.fmt == FMT_SPLIT ? out.c_pad : out.c;
    const int64_t total = n_frames * cp.Do * cp.Ho * cp.Wo * cw;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int co = static_cast<int>(i % cw);
        int64_t t = i / cw;
        const int q = static_cast<int>(t % cp.Wo); t /= cp.Wo;
        const int p = static_cast<int>(t % cp.Ho); t /= cp.Ho;
        const int z = static_cast<int>(t % cp.Do);
        const int64_t nf = t / cp.Do;
        float v = 0.0f;
        if (co < cp.cout) {
            float acc = 0.0f;
            int tap = 0;
            for (int a = 0; a < cp.kd; ++a) {
                const int d = z + a - cp.pd;
                for (int b = 0; b < cp.kh; ++b) {
                    const int h = p + b - cp.ph;
                    for (int c = 0; c < cp.kw; ++c, ++tap) {
                        const int w = q + c - cp.pw;
                        if (d < 0 || d >= cp.Di || h < 0 || h >= cp.Hi || w < 0 || w >= cp.Wi) continue;
                        const int64_t ip = ((nf * cp.Di + d) * cp.Hi + h) * cp.Wi + w;
                        acc += __ldg(Z + ip * cp.z_ld + tap * cp.cout + co);
                    }
                }
            }
            v = apply_act(acc + bias[co], cp.act1, cp.alpha1);
            v = fmaf(v, scale[co], shift[co]);
            v = apply_act(v, cp.act2, cp.alpha2);
        }
        tv_store(out, ((nf * cp.Do + z) * cp.Ho + p) * cp.Wo + q, co, v);
    }
}

// float4 variant: cout % 4 == 0, z_ld % 4 == 0, fp32 output with ld % 4 == 0; one thread per (pixel, 4 channels)
__global__ void col2im_vec4_kernel(const float* __restrict__ Z, TView out, int64_t n_frames, Col2imParams cp,
                                   const float* __restrict__ bias, const float* __restrict__ scale,
                                   const float* __restrict__ shift) {
    const int groups = cp.cout / 4;
    const int64_t total = n_frames * cp.Do * cp.Ho * cp.Wo * groups;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int co = static_cast<int>(i % groups) * 4;
        int64_t t = i / groups;
        const int q = static_cast<int>(t % cp.Wo); t /= cp.Wo;
        const int p = static_cast<int>(t % cp.Ho); t /= cp.Ho;
        const int z = static_cast<int>(t % cp.Do);
        const int64_t nf = t / cp.Do;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int tap = 0;
        for (int a = 0; a < cp.kd; ++a) {
            const int d = z + a - cp.pd;
            for (int b = 0; b < cp.kh; ++b) {
                const int h = p + b - cp.ph;
                for (int c = 0; c < cp.kw; ++c, ++tap) {
                    const int w = q + c - cp.pw;
                    if (d < 0 || d >= cp.Di || h < 0 || h >= cp.Hi || w < 0 || w >= cp.Wi) continue;
                    const int64_t ip = ((nf * cp.Di + d) * cp.Hi + h) * cp.Wi + w;
                    const float4 v = __ldg(reinterpret_cast<const float4*>(Z + ip * cp.z_ld + tap * cp.cout + co));
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;      // same (kd,kh,kw) order as the scalar path
                }
            }
        }
        float r[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v = apply_act(r[e] + bias[co + e], cp.act1, cp.alpha1);
            v = fmaf(v, scale[co + e], shift[co + e]);
            r[e] = apply_act(v, cp.act2, cp.alpha2);
        }
        const int64_t pix = ((nf * cp.Do + z) * cp.Ho + p) * cp.Wo + q;
        *reinterpret_cast<float4*>(out.f32 + pix * out.ld + co) = make_float4(r[0], r[1], r[2], r[3]);
    }
}

// ------------------------------------------------------------------ channel-slice copy / add
__global__ void copy_channels_kernel(TView in, TView out, int64_t n_pix, int c_off) {
    const int64_t total = n_pix * in.c;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t pix = i / in.c;
        const int ch = static_cast<int>(i - pix * in.c);
        tv_store(out, pix, c_off + ch, tv_load(in, pix, ch));
    }
}
__global__ void zero_pad_channels_kernel(TView out, int64_t n_pix) {
    const int npad = out.c_pad - out.c;
    if (out.fmt != FMT_SPLIT || npad <= 0) return;
    const int64_t total = n_pix * npad;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t pix = i / npad;
        const int ch = out.c + static_cast<int>(i - pix * npad);
        tv_store(out, pix, ch, 0.0f);
    }
}
__global__ void add_kernel(TView a, TView b, TView out, int64_t n_pix) {
    const int cw = out.fmt == FMT_SPLIT ? out.c_pad : out.c;
    const int64_t total = n_pix * cw;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t pix = i / cw;
        const int ch = static_cast<int>(i - pix * cw);
        const float v = ch < out.c ? tv_load(a, pix, ch) + tv_load(b, pix, ch) : 0.0f;
        tv_store(out, pix, ch, v);
    }
}

// ------------------------------------------------------------------ global pooling
// one block per frame; thread t owns channels t, t+blockDim, ...; sums in fp32 over the
// frame's pixels in index order (deterministic).
__global__ void gpool_kernel(TView in, TView out, int pix_per_frame, int is_avg) {
    const int64_t nf = blockIdx.x;
    const int cw = out.fmt == FMT_SPLIT ? out.c_pad : out.c;
    for (int ch = threadIdx.x; ch < cw; ch += blockDim.x) {
        float acc = is_avg ? 0.0f : -INFINITY;
        if (ch < out.c) {
            for (int px = 0; px < pix_per_frame; ++px) {
                const float v = tv_load(in, nf * pix_per_frame + px, ch);
                acc = is_avg ? acc + v : fmaxf(acc, v);
            }
            if (is_avg) acc = acc / static_cast<float>(pix_per_frame);
        } else {
            acc = 0.0f;
        }
        tv_store(out, nf, ch, acc);
    }
}

// ------------------------------------------------------------------ fused network head
// GlobalAveragePooling3D -> Softmax of one frame by one block, optionally preceded by the col2im gather of a tap-to-N
// head conv (+ bias / activation / BatchNorm): the TIMED head is then GEMM + ONE launch instead of GEMM + col2im + pool +
// softmax, and the (frames, pixels, classes) activation never goes to HBM.  Every sum runs in the order of the unfused
// kernels (col2im_vec4_kernel's tap order, gpool_kernel's pixel order, softmax_kernel's warp tree), so the fused and
// unfused paths agree bit for bit.  Dynamic shared memory: pix_per_frame * classes floats + classes floats.
__device__ __forceinline__ void head_pool_softmax(const float* act, float* logits, int n_pix, int c, int is_avg,
                                                  float* __restrict__ probs_row) {
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float acc = is_avg ? 0.0f : -INFINITY;
        for (int px = 0; px < n_pix; ++px) {
            const float v = act[px * c + ch];
            acc = is_avg ? acc + v : fmaxf(acc, v);
        }
        logits[ch] = is_avg ? acc / static_cast<float>(n_pix) : acc;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        float mx = -INFINITY;
        for (int ch = lane; ch < c; ch += 32) mx = fmaxf(mx, logits[ch]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.0f;
        for (int ch = lane; ch < c; ch += 32) sum += expf(logits[ch] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        for (int ch = lane; ch < c; ch += 32) probs_row[ch] = expf(logits[ch] - mx) / sum;
    }
}

// col2im (+ epilogue) -> global pool -> softmax; requires cout % 4 == 0 and z_ld % 4 == 0 (the float4 gather)
__global__ void head_col2im_pool_softmax_kernel(const float* __restrict__ Z, Col2imParams cp, const float* __restrict__ bias,
                                                const float* __restrict__ scale, const float* __restrict__ shift,
                                                int is_avg, float* __restrict__ probs) {
    extern __shared__ float s_head[];
    const int n_pix = cp.Do * cp.Ho * cp.Wo;
    float* act = s_head;
    float* logits = s_head + n_pix * cp.cout;
    const int64_t nf = blockIdx.x;
    const int groups = cp.cout / 4;
    for (int i = threadIdx.x; i < n_pix * groups; i += blockDim.x) {
        const int co = (i % groups) * 4;
        int t = i / groups;
        const int q = t % cp.Wo; t /= cp.Wo;
        const int pr = t % cp.Ho;
        const int z = t / cp.Ho;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int tap = 0;
        for (int a = 0; a < cp.kd; ++a) {
            const int d = z + a - cp.pd;
            for (int b = 0; b < cp.kh; ++b) {
                const int h = pr + b - cp.ph;
                for (int c = 0; c < cp.kw; ++c, ++tap) {
                    const int w = q + c - cp.pw;
                    if (d < 0 || d >= cp.Di || h < 0 || h >= cp.Hi || w < 0 || w >= cp.Wi) continue;
                    const int64_t ip = ((nf * cp.Di + d) * cp.Hi + h) * cp.Wi + w;
                    const float4 v = __ldg(reinterpret_cast<const float4*>(Z + ip * cp.z_ld + tap * cp.cout + co));
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            }
        }
        float r[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v = apply_act(r[e] + bias[co + e], cp.act1, cp.alpha1);
            v = fmaf(v, scale[co + e], shift[co + e]);
            r[e] = apply_act(v, cp.act2, cp.alpha2);
        }
        *reinterpret_cast<float4*>(act + ((z * cp.Ho + pr) * cp.Wo + q) * cp.cout + co) = make_float4(r[0], r[1], r[2], r[3]);
    }
    __syncthreads();
    head_pool_softmax(act, logits, n_pix, cp.cout, is_avg, probs + nf * cp.cout);
}

// global pool -> softmax of an fp32 (frames, pixels, classes) tensor: the pool sums straight from global memory
__global__ void head_pool_softmax_kernel(TView in, int n_pix, int is_avg, float* __restrict__ probs) {
    extern __shared__ float s_head[];
    float* logits = s_head;
    const int64_t nf = blockIdx.x;
    const int c = in.c;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float acc = is_avg ? 0.0f : -INFINITY;
        for (int px = 0; px < n_pix; ++px) {
            const float v = tv_load(in, nf * n_pix + px, ch);
            acc = is_avg ? acc + v : fmaxf(acc, v);
        }
        logits[ch] = is_avg ? acc / static_cast<float>(n_pix) : acc;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        float mx = -INFINITY;
        for (int ch = lane; ch < c; ch += 32) mx = fmaxf(mx, logits[ch]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.0f;
        for (int ch = lane; ch < c; ch += 32) sum += expf(logits[ch] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        for (int ch = lane; ch < c; ch += 32) probs[nf * c + ch] = expf(logits[ch] - mx) / sum;
    }
}

// ------------------------------------------------------------------ linear head conv + GlobalAveragePooling, collapsed
// mean_p conv(x)[p] = bias + sum_tap W_tap . S_tap,  S_tap = (1/P) sum_{p : p + tap - pad inside the volume} x[p + tap - pad]:
// the average over the output voxels commutes with a conv that has no activation, so the head needs one box sum of the
// input per filter tap (this kernel) and ONE (frames x taps*C_in) x (taps*C_in x classes) GEMM instead of a conv over every
// voxel -- D*H*W times fewer MMAs.  A thread owns two channels of one frame; the box sums are separable: row sums over x for
// the kw ranges, folded over y into the (kh, kw) ranges of a plane, folded over z.  Filter extents <= 3.
// Output: split bf16 planes, row = frame, column = tap * cin + c (the K order of the DHWIO kernel read as a dense layer).
struct BoxSumParams {
    int32_t D, H, W, c_pad, cin;
    int32_t kd, kh, kw, pd, ph, pw;      // filter extents, padding before
    int32_t Do, Ho, Wo;
    int32_t k_pad;                       // columns per output row (>= taps * cin, zero padded by the caller)
    float inv_n;                         // 1 / (Do * Ho * Wo)
    int64_t in_lo_off, out_lo_off;       // element offset of the lo plane
};
__global__ void gap_boxsum_kernel(const __nv_bfloat16* __restrict__ in_hi, __nv_bfloat16* __restrict__ out_hi,
                                  int64_t n_frames, BoxSumParams p) {
    const int pairs = (p.cin + 1) >> 1;
    const int64_t total = n_frames * pairs;
    // first / last input coordinate each tap touches, per axis
    int lo_d[3], hi_d[3], lo_h[3], hi_h[3], lo_w[3], hi_w[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        lo_d[t] = max(0, t - p.pd); hi_d[t] = min(p.D - 1, p.Do - 1 + t - p.pd);
        lo_h[t] = max(0, t - p.ph); hi_h[t] = min(p.H - 1, p.Ho - 1 + t - p.ph);
        lo_w[t] = max(0, t - p.pw); hi_w[t] = min(p.W - 1, p.Wo - 1 + t - p.pw);
    }
    for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t f = idx / pairs;
        const int c = 2 * static_cast<int>(idx - f * pairs);
        const bool two = c + 1 < p.cin;
        float acc[3][3][3][2];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int w = 0; w < 3; ++w) acc[a][b][w][0] = acc[a][b][w][1] = 0.f;
        const __nv_bfloat16* src = in_hi + f * p.D * p.H * p.W * static_cast<int64_t>(p.c_pad) + c;
        for (int z = 0; z < p.D; ++z) {
            float py[3][3][2];
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int w = 0; w < 3; ++w) py[b][w][0] = py[b][w][1] = 0.f;
            for (int y = 0; y < p.H; ++y) {
                float r[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
                const __nv_bfloat16* row = src + (static_cast<int64_t>(z) * p.H + y) * p.W * p.c_pad;
                for (int x = 0; x < p.W; ++x, row += p.c_pad) {
                    // c is even and c_pad a multiple of 16: the pair is 4-byte aligned (the odd tail channel of an odd cin
                    // reads a zero pad channel)
                    const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(row);
                    const __nv_bfloat162 l2 = *reinterpret_cast<const __nv_bfloat162*>(row + p.in_lo_off);
                    const float v0 = __low2float(h2) + __low2float(l2), v1 = __high2float(h2) + __high2float(l2);
#pragma unroll
                    for (int w = 0; w < 3; ++w)
                        if (w < p.kw && x >= lo_w[w] && x <= hi_w[w]) { r[w][0] += v0; r[w][1] += v1; }
                }
#pragma unroll
                for (int b = 0; b < 3; ++b)
                    if (b < p.kh && y >= lo_h[b] && y <= hi_h[b]) {
#pragma unroll
                        for (int w = 0; w < 3; ++w) { py[b][w][0] += r[w][0]; py[b][w][1] += r[w][1]; }
                    }
            }
#pragma unroll
            for (int a = 0; a < 3; ++a)
                if (a < p.kd && z >= lo_d[a] && z <= hi_d[a]) {
#pragma unroll
                    for (int b = 0; b < 3; ++b)
#pragma unroll
                        for (int w = 0; w < 3; ++w) { acc[a][b][w][0] += py[b][w][0]; acc[a][b][w][1] += py[b][w][1]; }
                }
        }
        __nv_bfloat16* dst = out_hi + f * static_cast<int64_t>(p.k_pad) + c;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int w = 0; w < 3; ++w) {
                    if (a >= p.kd || b >= p.kh || w >= p.kw) continue;
                    const int tap = (a * p.kh + b) * p.kw + w;
                    __nv_bfloat16 h0, l0, h1, l1;
                    split_bf16(acc[a][b][w][0] * p.inv_n, h0, l0);
                    split_bf16(acc[a][b][w][1] * p.inv_n, h1, l1);
                    __nv_bfloat16* o = dst + static_cast<int64_t>(tap) * p.cin;
                    o[0] = h0;
                    o[p.out_lo_off] = l0;
                    if (two) { o[1] = h1; o[1 + p.out_lo_off] = l1; }
                }
    }
}

// ------------------------------------------------------------------ softmax over channels
// one warp per row; max-subtracted, fp32 (Keras Softmax / activation='softmax').
__global__ void softmax_kernel(TView in, TView out, int64_t n_rows) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const int c = in.c;
    float mx = -INFINITY;
    for (int ch = lane; ch < c; ch += 32) mx = fmaxf(mx, tv_load(in, row, ch));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.0f;
    for (int ch = lane; ch < c; ch += 32) sum += expf(tv_load(in, row, ch) - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    for (int ch = lane; ch < c; ch += 32) tv_store(out, row, ch, expf(tv_load(in, row, ch) - mx) / sum);
}

// ------------------------------------------------------------------ argmax of fp16-rounded probs
__global__ void argmax_fp16_kernel(const float* __restrict__ probs, int64_t n, int n_cls,
                                   int32_t* __restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    if (row >= n) return;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int ch = lane; ch < n_cls; ch += 32) {
        // np.float16 cast (RNE), then compare as the reference's np.argmax does
        const float v = __half2float(__float2half_rn(probs[row * n_cls + ch]));
        if (v > best || (v == best && ch < bi) || bi == 0x7fffffff) { best = v; bi = ch; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) {
            best = ob;
            bi = oi;
        }
    }
    if (lane == 0) idx[row] = bi;
}

// ------------------------------------------------------------------ NMR consensus (utils.py:694-713)
// One warp per consensus row.  States of one structure are equally long blocks of rows; the reference folds them
// with a RUNNING PAIRWISE mean in float16 arithmetic, c <- fp16(fp16(c + p_s) / 2) (numpy float16 ops round after
// every operation), then takes the first-index argmax.  probs are fp32 and rounded to fp16 on the way in, as
// save_outputs_to_file / predict.py:163 do.
__global__ void consensus_fp16_kernel(const float* __restrict__ probs, const int64_t* __restrict__ first_row,
                                      const int32_t* __restrict__ n_states, const int32_t* __restrict__ n_res,
                                      const int64_t* __restrict__ out_row0, int n_groups, int64_t n_out_rows, int n_cls,
                                      __half* __restrict__ cons, int32_t* __restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    if (row >= n_out_rows) return;
    int lo = 0, hi = n_groups - 1;                    // last group whose first output row is <= row
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (out_row0[mid] <= row) lo = mid; else hi = mid - 1;
    }
    const int g = lo;
    const int64_t r = row - out_row0[g];
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int ch = lane; ch < n_cls; ch += 32) {
        __half c = __float2half_rn(probs[(first_row[g] + r) * n_cls + ch]);
        for (int s = 1; s < n_states[g]; ++s) {
            const __half v = __float2half_rn(probs[(first_row[g] + static_cast<int64_t>(s) * n_res[g] + r) * n_cls + ch]);
            const __half sum = __float2half_rn(__half2float(c) + __half2float(v));
            c = __float2half_rn(__half2float(sum) * 0.5f);
        }
        cons[row * n_cls + ch] = c;
        const float v = __half2float(c);
        if (v > best || (v == best && ch < bi) || bi == 0x7fffffff) { best = v; bi = ch; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) {
            best = ob;
            bi = oi;
        }
    }
    if (lane == 0) idx[row] = bi;
}

// ------------------------------------------------------------------ sequence metrics (analyse_utils.py:351-371)
// One warp per sampled sequence: 20-bin residue histogram, then composition-only closed forms from host-built
// tables T = [mw(20) | ext280(20) | q(pH_ref)(20) | term(pH_ref) | water | n_grid | grid(n_grid) | term(grid)(n_grid) |
// q(grid)(n_grid x 20)]: net charge at the reference pH, isoelectric point = first grid pH minimising |charge|,
// molecular weight, molar extinction at 280 nm.  out: (n_seqs, 4) float64; NaN row when a letter is not in `lut`.
__global__ void seq_metrics_kernel(const uint8_t* __restrict__ seqs, int64_t n_seqs, int64_t n_res,
                                   const int8_t* __restrict__ lut, const double* __restrict__ T,
                                   double* __restrict__ out) {
    __shared__ int s_cnt[8][21];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const int64_t seq = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + w;
    if (seq >= n_seqs) return;
    if (lane < 21) s_cnt[w][lane] = 0;
    __syncwarp();
    const uint8_t* srow = seqs + seq * n_res;
    for (int64_t i = lane; i < n_res; i += 32) {
        const int k = lut[srow[i]];
        atomicAdd(&s_cnt[w][k < 0 ? 20 : k], 1);
    }
    __syncwarp();
    const int n_grid = static_cast<int>(T[62]);
    const double* grid = T + 63;
    const double* tgrid = grid + n_grid;
    const double* qgrid = tgrid + n_grid;
    double best = INFINITY;
    int bi = 0x7fffffff;
    for (int gp = lane; gp < n_grid; gp += 32) {
        double c = 0.0;
        for (int i = 0; i < 20; ++i) c += static_cast<double>(s_cnt[w][i]) * qgrid[gp * 20 + i];
        c = fabs(c + tgrid[gp]);
        if (c < best) { best = c; bi = gp; }       // ascending gp per lane: first minimum kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
        double q = 0.0, mw = 0.0, ext = 0.0;
        for (int i = 0; i < 20; ++i) {
            const double n = static_cast<double>(s_cnt[w][i]);
            mw += n * T[i];
            ext += n * T[20 + i];
            q += n * T[40 + i];
        }
        const bool bad = s_cnt[w][20] != 0;
        double* o = out + seq * 4;
        o[0] = bad ? NAN : q + T[60];
        o[1] = bad ? NAN : grid[bi];
        o[2] = bad ? NAN : mw + T[61];
        o[3] = bad ? NAN : ext;
    }
}

// =================================================================== Monte-Carlo sampler
// Philox4x32-10 (Salmon et al. 2011), counter = (sample_lo, sample_hi, res_lo, res_hi).
__device__ __host__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c[0];
        const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c[2];
        const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = static_cast<uint32_t>(p1);
        const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = static_cast<uint32_t>(p0);
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
// 53-bit uniform in [0,1) from two 32-bit words, the same construction numpy's legacy
// random_sample uses: (a>>5, b>>6) -> (a*2^26 + b) / 2^53.
// One Philox block serves TWO residues: counter = (sample, residue >> 1); the even residue takes words 0-1, the odd one
// words 2-3 (round 1 threw half of every block away).
__device__ __host__ __forceinline__ void philox_block(uint64_t sample, uint64_t res_pair, uint64_t seed, uint64_t stream_id,
                                                      uint32_t (&c)[4]) {
    const uint64_t mix = stream_id * 0x9E3779B97F4A7C15ull;
    c[0] = static_cast<uint32_t>(sample); c[1] = static_cast<uint32_t>(sample >> 32);
    c[2] = static_cast<uint32_t>(res_pair); c[3] = static_cast<uint32_t>(res_pair >> 32);
    philox4x32_10(c, static_cast<uint32_t>(seed) ^ static_cast<uint32_t>(mix >> 32),
                  static_cast<uint32_t>(seed >> 32) ^ static_cast<uint32_t>(mix));
}
__device__ __host__ __forceinline__ double philox_words_to_uniform(uint32_t w0, uint32_t w1) {
    const uint32_t a = w0 >> 5, b = w1 >> 6;
    return (static_cast<double>(a) * 67108864.0 + static_cast<double>(b)) * (1.0 / 9007199254740992.0);
}
__device__ __host__ __forceinline__ double philox_uniform(uint64_t sample, uint64_t res,
                                                          uint64_t seed, uint64_t stream_id) {
    uint32_t c[4];
    philox_block(sample, res >> 1, seed, stream_id, c);
    return (res & 1) ? philox_words_to_uniform(c[2], c[3]) : philox_words_to_uniform(c[0], c[1]);
}

__global__ void sample_uniforms_kernel(int64_t n_res, int64_t n_samples, int64_t first_sample,
                                       uint64_t seed, uint64_t stream_id, double* __restrict__ out) {
    const int64_t total = n_res * n_samples;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t s = i / n_res;
        out[i] = philox_uniform(static_cast<uint64_t>(first_sample + s),
                                static_cast<uint64_t>(i - s * n_res), seed, stream_id);
    }
}

// numpy's pairwise summation (numpy/core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum),
// restated so that the row sums of apply_temp_to_probs match np.sum(axis=1) add-for-add.
__device__ double np_pairwise_sum(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

// apply_temp_to_probs (sampling_utils.py:159-161): p ** (1/t), row sum, divide.  One thread
// per row (rows are short: 20 or 338); fp64 throughout.
// `in` and `out` may be the SAME buffer (the sampler rescales in place): no __restrict__, no read-only loads; every
// element is read into a register before the store to the same address.
__global__ void temperature_kernel(const double* in, int64_t n_rows, int n_cls, double inv_t, double* out) {
    const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (row >= n_rows) return;
    const double* src = in + row * n_cls;
    double* dst = out + row * n_cls;
    for (int j = 0; j < n_cls; ++j) dst[j] = pow(src[j], inv_t);
    const double s = np_pairwise_sum(dst, n_cls);
    for (int j = 0; j < n_cls; ++j) dst[j] = dst[j] / s;
}

// Wide rows (338 rotamer categories): one WARP per row.  The thread-per-row kernels walk rows 2.7 KB apart (uncoalesced) with
// 13 k threads in flight; here the lanes load / pow / store a row together through shared memory and lane 0 alone does the
// order-sensitive part (numpy's pairwise sum, the sequential cumsum) on the staged row: same operations per element and the
// same summation order => bit-identical results.  Dynamic shared memory: warps_per_block * n_cls doubles.
__global__ void temperature_wide_kernel(const double* in, int64_t n_rows, int n_cls, double inv_t, double* out) {
    extern __shared__ double s_rows[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    double* buf = s_rows + static_cast<size_t>(warp) * n_cls;
    for (int64_t row = blockIdx.x * static_cast<int64_t>(wpb) + warp; row < n_rows; row += static_cast<int64_t>(gridDim.x) * wpb) {
        const double* src = in + row * n_cls;
        for (int j = lane; j < n_cls; j += 32) buf[j] = pow(src[j], inv_t);
        __syncwarp();
        double ssum = 0.0;
        if (lane == 0) ssum = np_pairwise_sum(buf, n_cls);
        ssum = __shfl_sync(0xffffffffu, ssum, 0);
        double* dst = out + row * n_cls;
        for (int j = lane; j < n_cls; j += 32) dst[j] = buf[j] / ssum;
        __syncwarp();
    }
}
__global__ void cumsum_rows_wide_kernel(const double* __restrict__ in, int64_t n_rows, int n_cls, double* __restrict__ out) {
    extern __shared__ double s_rows[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    double* buf = s_rows + static_cast<size_t>(warp) * n_cls;
    for (int64_t row = blockIdx.x * static_cast<int64_t>(wpb) + warp; row < n_rows; row += static_cast<int64_t>(gridDim.x) * wpb) {
        for (int j = lane; j < n_cls; j += 32) buf[j] = in[row * n_cls + j];
        __syncwarp();
        if (lane == 0) {
            double acc = buf[0];
            for (int j = 1; j < n_cls; ++j) {
                acc = __dadd_rn(acc, buf[j]);
                buf[j] = acc;
            }
        }
        __syncwarp();
        for (int j = lane; j < n_cls; j += 32) out[row * n_cls + j] = buf[j];
        __syncwarp();
    }
}

// probs.cumsum(axis=1): strictly sequential fp64 adds, one thread per row (bit-identical to numpy).
__global__ void cumsum_rows_kernel(const double* __restrict__ in, int64_t n_rows, int n_cls,
                                   double* __restrict__ out) {
    const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (row >= n_rows) return;
    double acc = 0.0;
    for (int j = 0; j < n_cls; ++j) {
        acc = j == 0 ? in[row * n_cls] : __dadd_rn(acc, in[row * n_cls + j]);
        out[row * n_cls + j] = acc;
    }
}

// Inverse-CDF draw.  Each thread owns 4 consecutive (sample, residue) cells of the flat
// (n_samples, n_res) output and writes them as one 32-bit word (and one int4 of indices).
// idx = first j with cdf[j] > r; none -> 0  ((cumsum > r).argmax(), sampling_utils.py:82).
// The CDF is nondecreasing (cumsum of non-negative terms), so "first j with cdf[j] > r" ==
// "number of j with cdf[j] <= r" -- a branch-free count for short rows, a binary search for
// long ones.  NaN-safe in the same way as the reference: comparisons with NaN are false.
// index drawn for one cell: first j with cdf_row[j] > r, none -> 0.  A CDF built by a sequential cumsum of non-negative
// probabilities is nondecreasing, so "first j with cdf[j] > r" is a lower-bound search (5 loads for 20 classes, 9 for
// 338); a row holding a NaN (then its last entry is NaN, cumsum propagates it) is not ordered and takes the literal
// first-true scan, which is what numpy's (cumsum > r).argmax() evaluates.
__device__ __forceinline__ int sample_index(const double* __restrict__ row, int n_cls, double r) {
    const double last = __ldg(row + n_cls - 1);
    int j;
    if (last != last) {
        j = n_cls;
        for (int k = 0; k < n_cls; ++k)
            if (__ldg(row + k) > r) { j = k; break; }
    } else {
        int lo = 0, hi = n_cls;          // first index with cdf > r in [lo, hi]
        if (!(last > r)) lo = n_cls;     // nothing exceeds r
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(row + mid) > r) hi = mid; else lo = mid + 1;
        }
        j = lo;
    }
    return j >= n_cls ? 0 : j;           // no entry exceeds r -> argmax of all-False is 0
}

// four consecutive cells (word `wi`) of one chain's flat (n_samples, n_res) block
__device__ __forceinline__ void sample_word(const double* __restrict__ cdf, int64_t n_res, int n_cls, int64_t total,
                                            int64_t first_sample, uint64_t seed, uint64_t stream_id,
                                            const double* __restrict__ uniforms, const uint8_t* s_letters,
                                            uint8_t* __restrict__ seqs, int32_t* __restrict__ idx_out, int64_t wi) {
    uint32_t packed = 0;
    int32_t id4[4] = {0, 0, 0, 0};
    // one division per word; (sample, residue) then advance incrementally.  A Philox block is reused by the two residues
    // of a pair when both fall into this word (they do unless the pair straddles the word or a row end).
    int64_t s = (wi * 4) / n_res;
    int64_t res = wi * 4 - s * n_res;
    uint32_t c[4];
    int64_t have_s = -1, have_pair = -1;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int64_t flat = wi * 4 + e;
        if (flat >= total) break;
        double r;
        if (uniforms) {
            r = uniforms[flat];
        } else {
            if (s != have_s || (res >> 1) != have_pair) {
                philox_block(static_cast<uint64_t>(first_sample + s), static_cast<uint64_t>(res >> 1), seed, stream_id, c);
                have_s = s;
                have_pair = res >> 1;
            }
            r = (res & 1) ? philox_words_to_uniform(c[2], c[3]) : philox_words_to_uniform(c[0], c[1]);
        }
        const int j = sample_index(cdf + res * n_cls, n_cls, r);
        id4[e] = j;
        packed |= static_cast<uint32_t>(s_letters[j]) << (8 * e);
        if (++res == n_res) { res = 0; ++s; }
    }
    if (wi * 4 + 3 < total) {
        reinterpret_cast<uint32_t*>(seqs)[wi] = packed;
        if (idx_out) reinterpret_cast<int4*>(idx_out)[wi] = make_int4(id4[0], id4[1], id4[2], id4[3]);
    } else {
        for (int e = 0; e < 4 && wi * 4 + e < total; ++e) {
            seqs[wi * 4 + e] = static_cast<uint8_t>(packed >> (8 * e));
            if (idx_out) idx_out[wi * 4 + e] = id4[e];
        }
    }
}

__global__ void sample_kernel(const double* __restrict__ cdf, int64_t n_res, int n_cls,
                              int64_t n_samples, int64_t first_sample, uint64_t seed,
                              uint64_t stream_id, const double* __restrict__ uniforms,
                              const uint8_t* __restrict__ letters, uint8_t* __restrict__ seqs,
                              int32_t* __restrict__ idx_out) {
    __shared__ uint8_t s_letters[512];
    for (int i = threadIdx.x; i < n_cls && i < 512; i += blockDim.x) s_letters[i] = letters[i];
    __syncthreads();
    const int64_t total = n_res * n_samples;
    const int64_t n_words = (total + 3) / 4;
    for (int64_t wi = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; wi < n_words;
         wi += static_cast<int64_t>(gridDim.x) * blockDim.x)
        sample_word(cdf, n_res, n_cls, total, first_sample, seed, stream_id, uniforms, s_letters, seqs, idx_out, wi);
}

// All chains of a structure set in ONE launch.  Chain c owns rows [row_off[c], row_off[c+1]) of the concatenated CDF
// and the byte range [seq_off[c], seq_off[c] + n_samples*n_res_c) of `seqs` (seq_off multiples of 4, ascending, last
// entry = total bytes); its draws are keyed (seed, stream_id0 + c) exactly as a per-chain timed_b200_sample call with
// that stream id, so the letters are byte-identical to per-chain launches.
__global__ void sample_chains_kernel(const double* __restrict__ cdf, const int64_t* __restrict__ row_off,
                                     const int64_t* __restrict__ seq_off, int n_chains, int n_cls, int64_t n_samples,
                                     int64_t first_sample, uint64_t seed, uint64_t stream_id0,
                                     const uint8_t* __restrict__ letters, uint8_t* __restrict__ seqs) {
    __shared__ uint8_t s_letters[512];
    for (int i = threadIdx.x; i < n_cls && i < 512; i += blockDim.x) s_letters[i] = letters[i];
    __syncthreads();
    const int64_t n_words = seq_off[n_chains] / 4;
    for (int64_t w = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; w < n_words;
         w += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        int lo = 0, hi = n_chains - 1;               // last chain whose block starts at or before this word
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (seq_off[mid] <= 4 * w) lo = mid; else hi = mid - 1;
        }
        const int64_t n_res = row_off[lo + 1] - row_off[lo];
        const int64_t total = n_res * n_samples;
        const int64_t wl = w - seq_off[lo] / 4;
        if (wl * 4 >= total) continue;               // alignment padding after the chain's block
        sample_word(cdf + row_off[lo] * n_cls, n_res, n_cls, total, first_sample, seed, stream_id0 + lo, nullptr,
                    s_letters, seqs + seq_off[lo], nullptr, wl);
    }
}

// ------------------------------------------------------------------ tiled sampler
// ncu on sample_chains_kernel (profiles/r2c_sampler_ncu.json): l1tex throughput 94 % of peak, 21.5 sectors per load request --
// every lane of a warp searches a DIFFERENT CDF row, so each of the ~6 loads per residue touches 32 cache lines and the
// kernel is bound by L1 line throughput, not by bytes or arithmetic.  Here a CTA owns a tile of TR consecutive residues of
// one chain x 256 consecutive samples: the tile's CDF rows are staged in shared memory once, the lanes of a warp are 32
// SAMPLES of the same residue (every search reads one shared-memory row), letters are packed through a shared tile and
// leave as contiguous runs of TR bytes.  Same generator keys and counters as sample_word => byte-identical output.
constexpr int kSampTileS = 256;

// tiles of chain c start at tile_off[c]: one thread, chains are few (tens to thousands)
__global__ void sample_tile_prefix_kernel(const int64_t* __restrict__ row_off, int n_chains, int tr, int32_t* __restrict__ tile_off) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int32_t acc = 0;
        for (int c = 0; c < n_chains; ++c) {
            tile_off[c] = acc;
            acc += static_cast<int32_t>((row_off[c + 1] - row_off[c] + tr - 1) / tr);
        }
        tile_off[n_chains] = acc;
    }
}

template <int TR>
__global__ void __launch_bounds__(kSampTileS)
sample_tiled_kernel(const double* __restrict__ cdf, const int64_t* __restrict__ row_off, const int64_t* __restrict__ seq_off,
                    const int32_t* __restrict__ tile_off, int n_chains, int n_cls, int64_t n_samples, int64_t first_sample,
                    uint64_t seed, uint64_t stream_id0, const uint8_t* __restrict__ letters, uint8_t* __restrict__ seqs) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    double* s_cdf = reinterpret_cast<double*>(s_raw);                                   // [TR][n_cls]
    uint32_t* s_out = reinterpret_cast<uint32_t*>(s_cdf + static_cast<size_t>(TR) * n_cls);   // [256][TR/4 + 1]
    __shared__ uint8_t s_letters[512];
    constexpr int PITCH = TR / 4 + 1;
    for (int i = threadIdx.x; i < n_cls && i < 512; i += blockDim.x) s_letters[i] = letters[i];
    const int rt = blockIdx.x;
    if (rt >= tile_off[n_chains]) return;
    int lo = 0, hi = n_chains - 1;                       // chain owning row tile rt
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tile_off[mid] <= rt) lo = mid; else hi = mid - 1;
    }
    const int chain = lo;
    const int64_t n_res = row_off[chain + 1] - row_off[chain];
    const int64_t r0 = static_cast<int64_t>(rt - tile_off[chain]) * TR;
    const int nr = static_cast<int>(min(static_cast<int64_t>(TR), n_res - r0));
    const double* rows = cdf + (row_off[chain] + r0) * n_cls;
    for (int i = threadIdx.x; i < nr * n_cls; i += blockDim.x) s_cdf[i] = rows[i];
    __syncthreads();
    const int64_t s_local = static_cast<int64_t>(blockIdx.y) * kSampTileS + threadIdx.x;
    const bool active = s_local < n_samples;
    const uint64_t stream = stream_id0 + static_cast<uint64_t>(chain);
    if (active) {
        uint32_t c[4];
        int64_t have_pair = -1;
        uint32_t packed = 0;
        for (int r = 0; r < nr; ++r) {
            const int64_t res = r0 + r;
            if ((res >> 1) != have_pair) {
                philox_block(static_cast<uint64_t>(first_sample + s_local), static_cast<uint64_t>(res >> 1), seed, stream, c);
                have_pair = res >> 1;
            }
            const double u = (res & 1) ? philox_words_to_uniform(c[2], c[3]) : philox_words_to_uniform(c[0], c[1]);
            const double* row = s_cdf + r * n_cls;
            const double last = row[n_cls - 1];
            int j;
            if (last != last) {                          // NaN row: the literal first-true scan (see sample_index)
                j = n_cls;
                for (int k = 0; k < n_cls; ++k)
                    if (row[k] > u) { j = k; break; }
            } else {
                int a = 0, b = n_cls;
                if (!(last > u)) a = n_cls;
                while (a < b) {
                    const int mid = (a + b) >> 1;
                    if (row[mid] > u) b = mid; else a = mid + 1;
                }
                j = a;
            }
            if (j >= n_cls) j = 0;
            packed |= static_cast<uint32_t>(s_letters[j]) << (8 * (r & 3));
            if ((r & 3) == 3 || r + 1 == nr) {
                s_out[threadIdx.x * PITCH + (r >> 2)] = packed;
                packed = 0;
            }
        }
    }
    __syncthreads();
    // write-out: a warp stores one sample row of the tile (nr contiguous bytes) per iteration
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* out = seqs + seq_off[chain];
    const int64_t s_base = static_cast<int64_t>(blockIdx.y) * kSampTileS;
    const int rows_here = static_cast<int>(min(static_cast<int64_t>(kSampTileS), n_samples - s_base));
    for (int sr = warp; sr < rows_here; sr += kSampTileS / 32) {
        for (int b = lane; b < nr; b += 32) {
            const uint32_t w = s_out[sr * PITCH + (b >> 2)];
            out[(s_base + sr) * n_res + r0 + b] = static_cast<uint8_t>(w >> (8 * (b & 3)));
        }
    }
}

#endif  // __CUDACC__

}  // namespace tb
