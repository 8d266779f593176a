// zlib / DEFLATE (RFC 1950 / 1951) decoder on the device: ONE WARP PER STREAM, thousands of independent streams per launch.
//
// The frames of an aposteriori dataset are stored as one gzip-filtered HDF5 chunk each (18 KB for 222 KB of voxels on real
// structures, design_utils/utils.py:514-529 reads them one by one through h5py).  Inflating them on the host costs ~0.5 ms of
// a core per frame and then ships 222 KB per frame over PCIe; here the STORED bytes go to the device and every frame is
// inflated by its own warp straight into the batch tensor the network reads:
//   * lane 0 walks the bit stream: block headers, code-length decoding, canonical Huffman tables in the warp's slice of
//     shared memory (a 9-bit direct lookup for literal/length codes, the count/offset walk of RFC 1951 3.2.2 for the rest),
//     literals stored as it goes;
//   * a match (length, distance) is broadcast and copied by all 32 lanes, out[pos + i] = out[pos - dist + i % dist] -- the
//     source bytes precede pos, so overlapping matches (runs of zeros: dist 4, length 258) need no ordering.
// Integer / byte work only; results are the bytes zlib produces (tests compare with zlib on stored, fixed and dynamic
// blocks).  A malformed or truncated stream, or one whose output is not exactly `out_bytes`, sets status[stream] != 0 and
// the caller inflates that batch on the host instead.
#pragma once
#include "common.cuh"

namespace tb {

constexpr int kInflateWarps = 8;                 // warps (streams) per block
struct InflateTables {
    uint16_t lcount[16], lsym[288];              // literal/length code: symbols per bit length, symbols in canonical order
    uint16_t dcount[16], dsym[32];               // distance code
    uint16_t lfast[512];                         // 9-bit direct lookup: (symbol << 4) | length, 0 = longer code
    uint8_t lens[320];                           // code lengths while a dynamic header is read
};

#if defined(__CUDACC__)

struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t buf;
    int cnt;
    bool over;                                   // read past the end of the stream
    __device__ __forceinline__ void refill() {
        while (cnt <= 56) {
            uint64_t b = 0;
            if (p < end) b = *p; else if (p >= end + 8) over = true;     // a few zero bytes of look-ahead are legitimate
            ++p;
            buf |= b << cnt;
            cnt += 8;
        }
    }
    __device__ __forceinline__ uint32_t peek(int n) const { return static_cast<uint32_t>(buf) & ((1u << n) - 1u); }
    __device__ __forceinline__ void skip(int n) { buf >>= n; cnt -= n; }
    __device__ __forceinline__ uint32_t bits(int n) {
        if (cnt < n) refill();
        const uint32_t v = n ? peek(n) : 0u;
        skip(n);
        return v;
    }
};

// canonical Huffman tables from code lengths (RFC 1951 3.2.2); returns false for an over-subscribed code
__device__ inline bool inflate_build(const uint8_t* lens, int n, uint16_t* count, uint16_t* sym, uint16_t* fast) {
    for (int i = 0; i < 16; ++i) count[i] = 0;
    for (int i = 0; i < n; ++i) ++count[lens[i]];
    int left = 1;
    for (int l = 1; l < 16; ++l) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return false;
    }
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + count[l];
    for (int i = 0; i < n; ++i)
        if (lens[i]) sym[offs[lens[i]]++] = static_cast<uint16_t>(i);
    if (fast) {
        for (int i = 0; i < 512; ++i) fast[i] = 0;
        // canonical codes, bit-reversed (DEFLATE packs Huffman codes most significant bit first into an LSB-first stream)
        int code = 0, idx = 0;
        for (int l = 1; l <= 9; ++l) {
            for (int k = 0; k < count[l]; ++k, ++idx, ++code) {
                uint32_t rev = __brev(static_cast<uint32_t>(code)) >> (32 - l);
                const uint16_t e = static_cast<uint16_t>((sym[idx] << 4) | l);
                for (uint32_t f = rev; f < 512u; f += 1u << l) fast[f] = e;
            }
            code <<= 1;
        }
    }
    count[0] = 0;
    return true;
}

// one symbol by the count / first-code walk; -1 = invalid code
__device__ __forceinline__ int inflate_decode_slow(BitReader& br, const uint16_t* count, const uint16_t* sym) {
    if (br.cnt < 15) br.refill();
    int code = 0, first = 0, index = 0;
    uint32_t b = static_cast<uint32_t>(br.buf);
    for (int l = 1; l <= 15; ++l) {
        code |= static_cast<int>(b & 1u);
        b >>= 1;
        const int c = count[l];
        if (code - c < first) {
            br.skip(l);
            return sym[index + (code - first)];
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

__device__ __constant__ uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__device__ __constant__ uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__device__ __constant__ uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__device__ __constant__ uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__device__ __constant__ uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// status: 0 ok, 1 bad zlib header, 2 bad block / code, 3 output overrun, 4 stream truncated, 5 output shorter than expected
__global__ void __launch_bounds__(32 * kInflateWarps)
inflate_streams_kernel(const uint8_t* __restrict__ comp, const int64_t* __restrict__ off, const int64_t* __restrict__ size,
                       int64_t n_streams, int64_t out_bytes, uint8_t* out, int32_t* status) {
    __shared__ InflateTables s_tab[kInflateWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t stream = static_cast<int64_t>(blockIdx.x) * kInflateWarps + warp;
    if (stream >= n_streams) return;
    InflateTables& T = s_tab[warp];
    uint8_t* dst = out + stream * out_bytes;
    BitReader br;
    br.p = comp + off[stream];
    br.end = br.p + size[stream];
    br.buf = 0; br.cnt = 0; br.over = false;
    int64_t pos = 0;
    int err = 0;
    int state = 0;            // lane 0: 0 = need a block header, 1 = inside a Huffman block, 2 = final block done
    bool last = false;
    if (lane == 0) {
        const uint32_t cmf = br.bits(8), flg = br.bits(8);
        if ((cmf & 15u) != 8u || ((cmf << 8) | flg) % 31u != 0u || (flg & 32u)) err = 1;
    }
    for (;;) {
        // ---- lane 0 decodes up to the next match (or the end); everything it stores itself are literals
        int op = 0, mlen = 0, mdist = 0;       // op 0 = match to copy, 1 = finished
        if (lane == 0 && !err) {
            for (;;) {
                if (state == 0) {
                    if (last) { op = 1; break; }
                    last = br.bits(1) != 0;
                    const uint32_t type = br.bits(2);
                    if (type == 0) {                                        // stored
                        br.skip(br.cnt & 7);
                        const uint32_t len = br.bits(16), nlen = br.bits(16);
                        if ((len ^ 0xffffu) != nlen) { err = 2; break; }
                        if (pos + len > out_bytes) { err = 3; break; }
                        for (uint32_t i = 0; i < len; ++i) dst[pos + i] = static_cast<uint8_t>(br.bits(8));
                        pos += len;
                        if (br.over) { err = 4; break; }
                        continue;
                    }
                    if (type == 1) {                                        // fixed codes
                        for (int i = 0; i < 144; ++i) T.lens[i] = 8;
                        for (int i = 144; i < 256; ++i) T.lens[i] = 9;
                        for (int i = 256; i < 280; ++i) T.lens[i] = 7;
                        for (int i = 280; i < 288; ++i) T.lens[i] = 8;
                        inflate_build(T.lens, 288, T.lcount, T.lsym, T.lfast);
                        for (int i = 0; i < 30; ++i) T.lens[i] = 5;
                        inflate_build(T.lens, 30, T.dcount, T.dsym, nullptr);
                    } else if (type == 2) {                                 // dynamic codes
                        const int nlen = static_cast<int>(br.bits(5)) + 257, ndist = static_cast<int>(br.bits(5)) + 1;
                        const int ncode = static_cast<int>(br.bits(4)) + 4;
                        if (nlen > 286 || ndist > 30) { err = 2; break; }
                        uint8_t cl[19];
                        for (int i = 0; i < 19; ++i) cl[i] = 0;
                        for (int i = 0; i < ncode; ++i) cl[kClOrder[i]] = static_cast<uint8_t>(br.bits(3));
                        // the code-length code lives in the distance slots until the real distance code replaces it
                        if (!inflate_build(cl, 19, T.dcount, T.dsym, nullptr)) { err = 2; break; }
                        int idx = 0;
                        while (idx < nlen + ndist) {
                            const int s = inflate_decode_slow(br, T.dcount, T.dsym);
                            if (s < 0) { err = 2; break; }
                            if (s < 16) { T.lens[idx++] = static_cast<uint8_t>(s); continue; }
                            int rep, val = 0;
                            if (s == 16) {
                                if (idx == 0) { err = 2; break; }
                                val = T.lens[idx - 1];
                                rep = 3 + static_cast<int>(br.bits(2));
                            } else if (s == 17) rep = 3 + static_cast<int>(br.bits(3));
                            else rep = 11 + static_cast<int>(br.bits(7));
                            if (idx + rep > nlen + ndist) { err = 2; break; }
                            while (rep--) T.lens[idx++] = static_cast<uint8_t>(val);
                        }
                        if (err) break;
                        if (T.lens[256] == 0) { err = 2; break; }
                        uint8_t dl[30];
                        for (int i = 0; i < ndist; ++i) dl[i] = T.lens[nlen + i];
                        if (!inflate_build(T.lens, nlen, T.lcount, T.lsym, T.lfast)) { err = 2; break; }
                        if (!inflate_build(dl, ndist, T.dcount, T.dsym, nullptr)) { err = 2; break; }
                    } else { err = 2; break; }
                    state = 1;
                }
                // ---- symbols of the current Huffman block
                if (br.cnt < 48) br.refill();
                int s;
                const uint16_t e = T.lfast[br.peek(9)];
                if (e) { s = e >> 4; br.skip(e & 15); }
                else {
                    s = inflate_decode_slow(br, T.lcount, T.lsym);
                    if (s < 0) { err = 2; break; }
                }
                if (s < 256) {
                    if (pos >= out_bytes) { err = 3; break; }
                    dst[pos++] = static_cast<uint8_t>(s);
                    continue;
                }
                if (s == 256) {
                    state = 0;
                    if (br.over) { err = 4; break; }
                    continue;
                }
                s -= 257;
                if (s >= 29) { err = 2; break; }
                mlen = kLenBase[s] + static_cast<int>(br.bits(kLenExtra[s]));
                const int d = inflate_decode_slow(br, T.dcount, T.dsym);
                if (d < 0 || d >= 30) { err = 2; break; }
                mdist = kDistBase[d] + static_cast<int>(br.bits(kDistExtra[d]));
                if (mdist > pos) { err = 2; break; }
                if (pos + mlen > out_bytes) { err = 3; break; }
                op = 0;
                break;
            }
        }
        err = __shfl_sync(0xffffffffu, err, 0);
        if (err) break;
        op = __shfl_sync(0xffffffffu, op, 0);
        if (op == 1) break;
        mlen = __shfl_sync(0xffffffffu, mlen, 0);
        mdist = __shfl_sync(0xffffffffu, mdist, 0);
        const int64_t p0 = __shfl_sync(0xffffffffu, pos, 0);
        __syncwarp();                                               // lane 0's literal stores are visible to the copying lanes
        const uint8_t* src = dst + p0 - mdist;
        for (int i = lane; i < mlen; i += 32) dst[p0 + i] = src[i % mdist];
        __syncwarp();
        pos = p0 + mlen;
    }
    if (lane == 0) {
        if (!err && pos != out_bytes) err = 5;
        status[stream] = err;
    }
}

#endif  // __CUDACC__

}  // namespace tb
