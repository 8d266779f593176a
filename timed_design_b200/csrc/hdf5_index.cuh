// Host-side index of many HDF5 frame datasets at once (the per-frame h5py lookups of
// design_utils/utils.py:514-529): for every object-header address, the file offset / stored size of the frame's single
// deflate chunk and the bytes of its `encoded_residue` attribute -- what frames.load_batch_device needs to ship the stored
// chunks to the GPU.  The Python reader parses ONE frame of the batch in full; every other frame must carry byte-identical
// dataspace / datatype / filter-pipeline messages and an identical attribute header, which is what a dataset written in one
// go looks like.  Version-1 object headers (with continuation blocks), version-3 chunked layout with a one-leaf version-1
// B-tree, exactly one unmasked chunk at the origin: anything else sets a per-frame status and the caller walks that batch
// with the Python reader.  Every read is bounds-checked against the mapped length.  No device work.
#pragma once
#include <cstdint>
#include <cstring>

namespace tb {

struct H5Template {
    const uint8_t* space; int32_t space_len;
    const uint8_t* type; int32_t type_len;
    const uint8_t* filters; int32_t filters_len;
    const uint8_t* attr_hdr; int32_t attr_hdr_len;
    int32_t attr_data_len;
    int32_t rank;
};

static inline uint64_t h5_u64(const uint8_t* p) { uint64_t v; std::memcpy(&v, p, 8); return v; }
static inline uint32_t h5_u32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
static inline uint16_t h5_u16(const uint8_t* p) { uint16_t v; std::memcpy(&v, p, 2); return v; }

// status: 0 ok, 1 out of bounds, 2 not a version-1 object header, 3 shared / unexpected message, 4 message differs from the
// template or is missing, 5 layout is not one unmasked chunk at the origin in a one-leaf B-tree
static int h5_index_one(const uint8_t* base, int64_t flen, int64_t base_addr, int64_t obj, const H5Template& t,
                        int64_t* chunk_off, int64_t* chunk_size, uint8_t* attr_out) {
    auto ok = [&](int64_t off, int64_t len) { return off >= 0 && len >= 0 && off <= flen && len <= flen - off; };   // (no overflow)
    int64_t a = obj + base_addr;
    if (!ok(a, 16)) return 1;
    if (base[a] != 1) return 2;
    const int n_msgs = h5_u16(base + a + 2);
    int64_t blk_off[16], blk_len[16];
    int n_blk = 1, seen = 0;
    blk_off[0] = a + 16;
    blk_len[0] = h5_u32(base + a + 8);
    bool got_space = false, got_type = false, got_filters = t.filters_len == 0, got_attr = false, got_layout = false;
    uint64_t btree = 0;
    for (int b = 0; b < n_blk; ++b) {
        int64_t q = blk_off[b];
        const int64_t end = q + blk_len[b];
        if (!ok(q, blk_len[b])) return 1;
        while (q + 8 <= end && seen < n_msgs + 64) {
            const int mtype = h5_u16(base + q), msize = h5_u16(base + q + 2), mflags = base[q + 4];
            const uint8_t* d = base + q + 8;
            if (q + 8 + msize > end) return 1;
            q += 8 + msize;
            ++seen;
            if (mtype == 0) continue;
            if (mtype == 0x10) {
                if (msize < 16 || n_blk >= 16) return 3;
                blk_off[n_blk] = static_cast<int64_t>(h5_u64(d)) + base_addr;
                blk_len[n_blk] = static_cast<int64_t>(h5_u64(d + 8));
                ++n_blk;
                continue;
            }
            if (mflags & 0x02) {                                        // shared message: the Python reader's business
                if (mtype == 0x01 || mtype == 0x03 || mtype == 0x0B || mtype == 0x08 || mtype == 0x0C) return 3;
                continue;
            }
            if (mtype == 0x01) {
                if (msize != t.space_len || std::memcmp(d, t.space, msize)) return 4;
                got_space = true;
            } else if (mtype == 0x03) {
                if (msize != t.type_len || std::memcmp(d, t.type, msize)) return 4;
                got_type = true;
            } else if (mtype == 0x0B) {
                if (msize != t.filters_len || std::memcmp(d, t.filters, msize)) return 4;
                got_filters = true;
            } else if (mtype == 0x08) {
                if (msize < 3 + 8 + 4 * (t.rank + 1) || d[0] != 3 || d[1] != 2 || d[2] != t.rank + 1) return 5;
                btree = h5_u64(d + 3);
                got_layout = true;
            } else if (mtype == 0x0C) {
                if (msize == t.attr_hdr_len + t.attr_data_len && !std::memcmp(d, t.attr_hdr, t.attr_hdr_len)) {
                    std::memcpy(attr_out, d + t.attr_hdr_len, t.attr_data_len);
                    got_attr = true;
                }
            }
        }
    }
    if (!(got_space && got_type && got_filters && got_attr && got_layout)) return 4;
    if (btree == ~0ull) return 5;
    const int64_t n = static_cast<int64_t>(btree) + base_addr;
    const int key = 8 + 8 * (t.rank + 1);
    if (!ok(n, 24 + key + 8)) return 1;
    if (std::memcmp(base + n, "TREE", 4) || base[n + 4] != 1 || base[n + 5] != 0 || h5_u16(base + n + 6) != 1) return 5;
    const uint8_t* k = base + n + 24;
    if (h5_u32(k + 4) != 0) return 5;                                    // filter mask: a filter was skipped for this chunk
    for (int i = 0; i <= t.rank; ++i)
        if (h5_u64(k + 8 + 8 * i) != 0) return 5;
    const int64_t off = static_cast<int64_t>(h5_u64(k + key)) + base_addr, size = h5_u32(k);
    if (!ok(off, size) || size == 0) return 1;
    *chunk_off = off;
    *chunk_size = size;
    return 0;
}

}  // namespace tb
