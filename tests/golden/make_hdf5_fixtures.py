"""Spec-derived HDF5 byte fixtures for the reader (tests/test_hdf5_fixtures_cpu.py).

    python tests/golden/make_hdf5_fixtures.py        # rewrites tests/golden/spec_*.h5 + spec_expected.json

This script does NOT import timed_design_b200.hdf5 (neither the reader under test nor its writer): every structure is
assembled here, field by field, from the HDF5 File Format Specification v3.0 (section numbers in the comments), so that
the reader's branches which its own writer never produces get an input: version-2 superblock, version-2 object headers
with a continuation chunk, new-style groups with compact link messages, version-3 attribute messages, filter pipeline
v2 with shuffle + deflate + fletcher32, a two-level chunk B-tree, variable-length strings in a global heap (including a
> 32 KB model_config), h5py-style boolean enums -- and, in the second file, the version-0 superblock / symbol-table
layout with a group B-tree that spans several symbol-table nodes and a version-1 object header continuation.
No h5py / libhdf5 exists offline; these files are what the specification says such a library would write.
"""
from __future__ import annotations

import json
import struct
import zlib
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
UNDEF = 0xFFFFFFFFFFFFFFFF


# ----------------------------------------------------------------------------- checksums (spec: "Checksum" fields)
def lookup3(data: bytes, init: int = 0) -> int:
    """Bob Jenkins' lookup3 hashlittle, the metadata checksum of version-2 structures (H5_checksum_lookup3)."""
    M = 0xFFFFFFFF
    rot = lambda x, k: ((x << k) | (x >> (32 - k))) & M
    a = b = c = (0xDEADBEEF + len(data) + init) & M
    p, n = 0, len(data)
    while n > 12:
        a = (a + int.from_bytes(data[p:p + 4], "little")) & M
        b = (b + int.from_bytes(data[p + 4:p + 8], "little")) & M
        c = (c + int.from_bytes(data[p + 8:p + 12], "little")) & M
        a = (a - c) & M; a ^= rot(c, 4); c = (c + b) & M
        b = (b - a) & M; b ^= rot(a, 6); a = (a + c) & M
        c = (c - b) & M; c ^= rot(b, 8); b = (b + a) & M
        a = (a - c) & M; a ^= rot(c, 16); c = (c + b) & M
        b = (b - a) & M; b ^= rot(a, 19); a = (a + c) & M
        c = (c - b) & M; c ^= rot(b, 4); b = (b + a) & M
        p += 12
        n -= 12
    tail = data[p:] + b"\x00" * 12
    if n == 0:
        return c
    a = (a + int.from_bytes(tail[0:4], "little")) & M if n > 0 else a
    b = (b + int.from_bytes(tail[4:8], "little")) & M if n > 4 else b
    c = (c + int.from_bytes(tail[8:12], "little")) & M if n > 8 else c
    c ^= b; c = (c - rot(b, 14)) & M
    a ^= c; a = (a - rot(c, 11)) & M
    b ^= a; b = (b - rot(a, 25)) & M
    c ^= b; c = (c - rot(b, 16)) & M
    a ^= c; a = (a - rot(c, 4)) & M
    b ^= a; b = (b - rot(a, 14)) & M
    c ^= b; c = (c - rot(b, 24)) & M
    return c


def fletcher32(data: bytes) -> int:
    """H5_checksum_fletcher32: 16-bit big-endian words, sums modulo 65535, (sum2 << 16) | sum1."""
    s1 = s2 = 0
    n = len(data) // 2
    for i in range(n):
        s1 = (s1 + ((data[2 * i] << 8) | data[2 * i + 1])) % 65535
        s2 = (s2 + s1) % 65535
    if len(data) % 2:
        s1 = (s1 + (data[-1] << 8)) % 65535
        s2 = (s2 + s1) % 65535
    return (s2 << 16) | s1


# ----------------------------------------------------------------------------- messages shared by both layouts
def dt_float(size: int) -> bytes:
    """IV.A.2.d class 1, little-endian IEEE (bit field: mantissa normalisation 2 = implied msb; sign location)."""
    if size == 4:
        return struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    return struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)


def dt_int(size: int, signed: bool = True) -> bytes:
    return struct.pack("<BBBBI", 0x10, 0x08 if signed else 0, 0, 0, size) + struct.pack("<HH", 0, 8 * size)


def dt_fixed_string(size: int) -> bytes:
    return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, size)          # null-terminated ASCII


def dt_vlen_utf8() -> bytes:
    """class 9, type = string (1), padding null-terminate, charset UTF-8; base type = 1-byte UTF-8 string."""
    return struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + struct.pack("<BBBBI", 0x13, 0x10, 0, 0, 1)


def dt_bool_enum(version: int = 1) -> bytes:
    """h5py's numpy-bool: ENUM over int8 with members FALSE=0, TRUE=1 (datatype version 1 pads names to 8 bytes, 3 does not)."""
    base = dt_int(1)
    names = b""
    for nm in (b"FALSE", b"TRUE"):
        s = nm + b"\x00"
        if version < 3:
            s += b"\x00" * (-len(s) % 8)
        names += s
    return struct.pack("<BBBBI", (version << 4) | 8, 2, 0, 0, 1) + base + names + bytes([0, 1])


def ds_simple_v1(shape) -> bytes:
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def ds_v2(shape) -> bytes:
    """dataspace message version 2: type 0 scalar / 1 simple."""
    if shape == ():
        return struct.pack("<BBBB", 2, 0, 0, 0)
    return struct.pack("<BBBB", 2, len(shape), 0, 1) + b"".join(struct.pack("<Q", s) for s in shape)


class Blob:
    """Append-only file image with 8-byte aligned allocation."""

    def __init__(self, reserve: int):
        self.b = bytearray(reserve)

    def alloc(self, data: bytes, align: int = 8) -> int:
        self.b += b"\x00" * (-len(self.b) % align)
        off = len(self.b)
        self.b += data
        return off

    def patch(self, off: int, data: bytes):
        self.b[off:off + len(data)] = data


class GlobalHeap:
    """III.E: 'GCOL' collections of (index, refcount, size, data) objects; one collection per heap here."""

    def __init__(self, blob: Blob):
        self.blob = blob

    def put(self, objs) -> list:
        body = b""
        for i, o in enumerate(objs, start=1):
            body += struct.pack("<HHIQ", i, 1, 0, len(o)) + o + b"\x00" * (-len(o) % 8)
        size = 16 + len(body) + 16
        body += struct.pack("<HHIQ", 0, 0, 0, 16)                 # free-space object 0
        addr = self.blob.alloc(b"GCOL" + struct.pack("<B3xQ", 1, size) + body)
        return [(addr, i) for i in range(1, len(objs) + 1)]


def vlen_elements(blob: Blob, strings) -> bytes:
    """16-byte vlen descriptors (length, global heap address, object index) for UTF-8 strings."""
    raws = [s.encode("utf-8") for s in strings]
    ids = GlobalHeap(blob).put(raws)
    return b"".join(struct.pack("<IQI", len(r), a, i) for r, (a, i) in zip(raws, ids))


# ----------------------------------------------------------------------------- file 1: "latest" layout
def attr_v3(name: str, dtype: bytes, space: bytes, data: bytes) -> bytes:
    """IV.A.2.m version 3: version, flags, name size, datatype size, dataspace size, name charset; no padding."""
    nm = name.encode() + b"\x00"
    return struct.pack("<BBHHHB", 3, 0, len(nm), len(dtype), len(space), 0) + nm + dtype + space + data


def ohdr_v2(messages, split_after: int | None, blob: Blob) -> int:
    """IV.A.1.b version-2 object header.  messages: [(type, body)].  With split_after = k the messages after the k-th go
    to an 'OCHK' continuation chunk referenced by a continuation message (0x10) in the first chunk."""
    def pack_msgs(ms):
        return b"".join(struct.pack("<BHB", t, len(b), 0) + b for t, b in ms)
    first = messages if split_after is None else messages[:split_after]
    rest = [] if split_after is None else messages[split_after:]
    cont_addr_patch = None
    if rest:
        body = pack_msgs(rest)
        ochk = b"OCHK" + body
        ochk += struct.pack("<I", lookup3(ochk))
        cont = blob.alloc(ochk)
        first = first + [(0x10, struct.pack("<QQ", cont, len(ochk)))]
    body = pack_msgs(first)
    # flags: bits 0-1 = size of the chunk#0 length field (here 2 -> 4 bytes); bit 5: times stored
    head = b"OHDR" + struct.pack("<BB", 2, 0x22) + struct.pack("<IIII", 1700000000, 1700000000, 1700000000, 1700000000)
    head += struct.pack("<I", len(body))
    raw = head + body
    raw += struct.pack("<I", lookup3(raw))
    return blob.alloc(raw)


def link_msg(name: str, addr: int) -> bytes:
    """IV.A.2.g link message, version 1, hard link, 1-byte name length."""
    nm = name.encode()
    return struct.pack("<BB", 1, 0x00) + struct.pack("<B", len(nm)) + nm + struct.pack("<Q", addr)


def link_info_compact() -> bytes:
    return struct.pack("<BB", 0, 0) + struct.pack("<QQ", UNDEF, UNDEF)      # fractal heap / name index: undefined = compact


def chunk_btree(blob: Blob, rank: int, entries, dims_elem_size: int, shape, fanout: int = 3) -> int:
    """III.A.1 version-1 B-tree, node type 1 (raw data chunks).  entries: [(origin tuple, stored size, address)] in
    row-major chunk order.  More than `fanout` entries -> a level-1 root over level-0 leaves."""
    def key(origin, size):
        return struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", o) for o in origin) + struct.pack("<Q", 0)

    def node(level, ents, next_key_origin):
        raw = b"TREE" + struct.pack("<BBH", 1, level, len(ents)) + struct.pack("<QQ", UNDEF, UNDEF)
        for origin, size, addr in ents:
            raw += key(origin, size) + struct.pack("<Q", addr)
        raw += key(next_key_origin, 0)
        return blob.alloc(raw)
    past_end = tuple(shape)
    if len(entries) <= fanout:
        return node(0, entries, past_end)
    leaves = []
    for i in range(0, len(entries), fanout):
        part = entries[i:i + fanout]
        nxt = entries[i + fanout][0] if i + fanout < len(entries) else past_end
        leaves.append((part[0][0], part[0][1], node(0, part, nxt)))
    return node(1, leaves, past_end)


def filtered_chunk(raw: bytes, esize: int, shuffle: bool, deflate: bool, fletcher: bool) -> bytes:
    if shuffle:
        n = len(raw) // esize
        raw = bytes(np.frombuffer(raw[:n * esize], np.uint8).reshape(n, esize).T.tobytes()) + raw[n * esize:]
    if deflate:
        raw = zlib.compress(raw, 4)
    if fletcher:
        raw += struct.pack("<I", fletcher32(raw))
    return raw


def chunked_dataset_v2(blob: Blob, arr: np.ndarray, chunk, shuffle, deflate, fletcher, attrs, fanout=3) -> int:
    esize = arr.dtype.itemsize
    rank = arr.ndim
    entries = []
    grid = [range(0, s, c) for s, c in zip(arr.shape, chunk)]
    for origin in np.ndindex(*[len(g) for g in grid]):
        org = tuple(grid[d][o] for d, o in enumerate(origin))
        block = np.zeros(chunk, dtype=arr.dtype)
        sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(org, chunk, arr.shape))
        block[tuple(slice(0, s.stop - s.start) for s in sl)] = arr[sl]
        data = filtered_chunk(block.tobytes(), esize, shuffle, deflate, fletcher)
        entries.append((org, len(data), blob.alloc(data)))
    btree = chunk_btree(blob, rank, entries, esize, arr.shape, fanout)
    # IV.A.2.i data layout version 3, class 2 (chunked): dimensionality = rank+1, B-tree address, chunk dims + element size
    layout = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", btree) + b"".join(struct.pack("<I", c) for c in chunk) + struct.pack("<I", esize)
    # IV.A.2.l filter pipeline version 2: (id, flags, n client values, values), no name for ids < 256
    filt = []
    if shuffle:
        filt.append(struct.pack("<HHH", 2, 0, 1) + struct.pack("<I", esize))
    if deflate:
        filt.append(struct.pack("<HHH", 1, 1, 1) + struct.pack("<I", 4))
    if fletcher:
        filt.append(struct.pack("<HHH", 3, 0, 0))
    dtype = dt_float(esize) if arr.dtype.kind == "f" else (dt_bool_enum(3) if arr.dtype == np.bool_ else dt_int(esize, arr.dtype.kind == "i"))
    msgs = [(0x01, ds_v2(arr.shape)), (0x03, dtype),
            (0x05, struct.pack("<BBBB", 3, 0x09, 0, 0)[:2]),             # fill value v3, flags: alloc time early, undefined fill
            (0x0B, struct.pack("<BB", 2, len(filt)) + b"".join(filt)), (0x08, layout)] + [(0x0C, a) for a in attrs]
    return ohdr_v2(msgs, split_after=4, blob=blob)


def contiguous_dataset_v2(blob: Blob, arr: np.ndarray, attrs=()) -> int:
    data = blob.alloc(arr.tobytes())
    layout = struct.pack("<BB", 3, 1) + struct.pack("<QQ", data, arr.nbytes)
    msgs = [(0x01, ds_v2(arr.shape)), (0x03, dt_float(arr.dtype.itemsize)), (0x08, layout)] + [(0x0C, a) for a in attrs]
    return ohdr_v2(msgs, None, blob)


def compact_dataset_v2(blob: Blob, arr: np.ndarray) -> int:
    raw = arr.tobytes()
    layout = struct.pack("<BB", 3, 0) + struct.pack("<H", len(raw)) + raw
    return ohdr_v2([(0x01, ds_v2(arr.shape)), (0x03, dt_float(arr.dtype.itemsize)), (0x08, layout)], None, blob)


def group_v2(blob: Blob, links, attrs=(), split_after=None) -> int:
    msgs = [(0x02, link_info_compact()), (0x0A, struct.pack("<BB", 0, 0))] + [(0x06, link_msg(n, a)) for n, a in links] + [(0x0C, a) for a in attrs]
    return ohdr_v2(msgs, split_after, blob)


def build_latest(rng) -> tuple:
    """Keras-model-shaped file + one frame-dataset-shaped branch, 'libver=latest' structures throughout."""
    blob = Blob(48)                                        # superblock v2 is 48 bytes with 8-byte offsets
    expected = {}
    # ---- model_weights/conv3d/{kernel:0 (chunked, shuffle+deflate+fletcher32, 2-level B-tree), bias:0 (contiguous)}
    kernel = rng.standard_normal((3, 3, 3, 6, 8)).astype(np.float32)
    bias = rng.standard_normal(8).astype(np.float32)
    gamma = rng.uniform(0.5, 1.5, 8).astype(np.float32)
    k_addr = chunked_dataset_v2(blob, kernel, (2, 2, 3, 6, 8), True, True, True, [], fanout=3)     # 2*2*1 = 4 chunks > fanout
    b_addr = contiguous_dataset_v2(blob, bias)
    g_addr = compact_dataset_v2(blob, gamma)
    wn = [b"conv3d/kernel:0", b"conv3d/bias:0"]
    width = max(len(x) for x in wn)
    wn_attr = attr_v3("weight_names", dt_fixed_string(width), ds_v2((2,)), b"".join(x.ljust(width, b"\x00") for x in wn))
    inner = group_v2(blob, [("bias:0", b_addr), ("kernel:0", k_addr)])
    conv_g = group_v2(blob, [("conv3d", inner)], [wn_attr])
    bn_inner = group_v2(blob, [("gamma:0", g_addr)])
    bn_names = attr_v3("weight_names", dt_fixed_string(29), ds_v2((1,)), b"batch_normalization/gamma:0".ljust(29, b"\x00"))
    bn_g = group_v2(blob, [("batch_normalization", bn_inner)], [bn_names])
    ln = [b"conv3d", b"batch_normalization"]
    lw = max(len(x) for x in ln)
    ln_attr = attr_v3("layer_names", dt_fixed_string(lw), ds_v2((2,)), b"".join(x.ljust(lw, b"\x00") for x in ln))
    # vlen string attributes (what h5py writes for python str): backend, keras_version through the global heap
    mw = group_v2(blob, [("batch_normalization", bn_g), ("conv3d", conv_g)],
                  [ln_attr, attr_v3("backend", dt_vlen_utf8(), ds_v2(()), vlen_elements(blob, ["tensorflow"])),
                   attr_v3("keras_version", dt_vlen_utf8(), ds_v2(()), vlen_elements(blob, ["2.13.1"]))], split_after=3)
    # ---- a frame-dataset branch: /1abc/A/{7, 12}: bool voxels (enum) and float frames with label + encoded_residue attrs
    frame = (rng.random((5, 5, 5, 6)) > 0.8)
    fr_attrs = [attr_v3("label", dt_vlen_utf8(), ds_v2(()), vlen_elements(blob, ["GLY"])),
                attr_v3("encoded_residue", dt_float(8), ds_v2((20,)), np.eye(20)[5].astype("<f8").tobytes())]
    f7 = chunked_dataset_v2(blob, frame, (3, 5, 5, 6), False, True, False, fr_attrs)
    gfr = rng.random((5, 5, 5, 6))
    f12 = chunked_dataset_v2(blob, gfr.astype("<f8"), (5, 5, 5, 6), True, True, False,
                             [attr_v3("label", dt_vlen_utf8(), ds_v2(()), vlen_elements(blob, ["TRP"])),
                              attr_v3("encoded_residue", dt_float(8), ds_v2((20,)), np.eye(20)[18].astype("<f8").tobytes())])
    chain = group_v2(blob, [("12", f12), ("7", f7)])
    pdb = group_v2(blob, [("A", chain)])
    # ---- root: a > 32 KB model_config (vlen UTF-8 in the global heap), numeric / boolean / array attributes
    cfg = {"class_name": "Functional", "config": {"name": "spec_fixture", "layers": [
        {"class_name": "InputLayer", "name": f"input_{i}", "config": {"name": f"input_{i}", "note": "x" * 300}, "inbound_nodes": []}
        for i in range(100)]}, "keras_version": "2.13.1", "backend": "tensorflow"}
    cfg_s = json.dumps(cfg)
    assert len(cfg_s) > 32768
    root_attrs = [attr_v3("model_config", dt_vlen_utf8(), ds_v2(()), vlen_elements(blob, [cfg_s])),
                  attr_v3("keras_version", dt_vlen_utf8(), ds_v2(()), vlen_elements(blob, ["2.13.1"])),
                  attr_v3("frame_dims", dt_int(8), ds_v2((4,)), np.array([5, 5, 5, 6], "<i8").tobytes()),
                  attr_v3("voxels_as_gaussian", dt_bool_enum(3), ds_v2(()), bytes([1])),
                  attr_v3("frame_edge_length", dt_float(8), ds_v2(()), struct.pack("<d", 21.0)),
                  attr_v3("atom_encoder", dt_vlen_utf8(), ds_v2((3,)), vlen_elements(blob, ["C", "N", "Cα"]))]
    root = group_v2(blob, [("1abc", pdb), ("model_weights", mw)], root_attrs, split_after=5)
    # ---- superblock version 2 (II.A): signature, version, offset/length sizes, flags, base, extension, EOF, root, checksum
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", 0, UNDEF, len(blob.b), root)
    sb += struct.pack("<I", lookup3(sb))
    blob.patch(0, sb)
    expected.update({"kernel": kernel.tolist(), "bias": bias.tolist(), "gamma": gamma.tolist(),
                     "model_config_len": len(cfg_s), "model_config_layers": 100, "frame7": frame.astype(int).tolist(),
                     "frame12": gfr.tolist()})
    return bytes(blob.b), expected


# ----------------------------------------------------------------------------- file 2: "earliest" layout
def attr_v1(name: str, dtype: bytes, space: bytes, data: bytes) -> bytes:
    """version-1 attribute message: name, datatype and dataspace each padded to 8 bytes."""
    pad = lambda b: b + b"\x00" * (-len(b) % 8)
    nm = name.encode() + b"\x00"
    return struct.pack("<BxHHH", 1, len(nm), len(dtype), len(space)) + pad(nm) + pad(dtype) + pad(space) + data


def ohdr_v1(messages, blob: Blob, split_after=None) -> int:
    """IV.A.1.a version-1 object header: 16-byte prefix, messages 8-byte aligned (type u16, size u16, flags u8, 3 reserved)."""
    def pack_msgs(ms):
        out = b""
        for t, b in ms:
            b = b + b"\x00" * (-len(b) % 8)
            out += struct.pack("<HHB3x", t, len(b), 0) + b
        return out
    first = messages if split_after is None else messages[:split_after]
    rest = [] if split_after is None else messages[split_after:]
    n_msgs = len(messages)
    if rest:
        body = pack_msgs(rest)
        cont = blob.alloc(body)
        first = first + [(0x10, struct.pack("<QQ", cont, len(body)))]
        n_msgs += 1
    body = pack_msgs(first)
    return blob.alloc(struct.pack("<BxHII4x", 1, n_msgs, 1, len(body)) + body)


def symbol_table_group(blob: Blob, children, attrs=(), per_node: int = 4) -> int:
    """Old-style group: local heap with the names, symbol-table nodes ('SNOD') of at most `per_node` entries in name
    order, one version-1 B-tree node (type 0) over them, symbol-table message (0x11) in the object header."""
    children = sorted(children, key=lambda kv: kv[0].encode())
    heap_data = bytearray(b"\x00" * 8)                         # offset 0: the empty string
    name_off = {}
    for n, _ in children:
        name_off[n] = len(heap_data)
        s = n.encode() + b"\x00"
        heap_data += s + b"\x00" * (-len(s) % 8)
    data_addr = blob.alloc(bytes(heap_data))
    heap = blob.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, data_addr))
    snods = []
    for i in range(0, len(children), per_node):
        part = children[i:i + per_node]
        raw = b"SNOD" + struct.pack("<BxH", 1, len(part))
        for n, a in part:
            raw += struct.pack("<QQI4x16x", name_off[n], a, 0)
        raw += b"\x00" * (40 * (per_node - len(part)))
        snods.append((name_off[part[-1][0]], blob.alloc(raw)))
    tree = b"TREE" + struct.pack("<BBH", 0, 0, len(snods)) + struct.pack("<QQ", UNDEF, UNDEF) + struct.pack("<Q", 0)
    for last_name, addr in snods:
        tree += struct.pack("<QQ", addr, last_name)
    btree = blob.alloc(tree)
    return ohdr_v1([(0x11, struct.pack("<QQ", btree, heap))] + [(0x0C, a) for a in attrs], blob, split_after=1 if len(attrs) > 1 else None), btree, heap


def chunked_dataset_v1(blob: Blob, arr: np.ndarray, chunk, attrs) -> int:
    esize = arr.dtype.itemsize
    entries = []
    grid = [range(0, s, c) for s, c in zip(arr.shape, chunk)]
    for origin in np.ndindex(*[len(g) for g in grid]):
        org = tuple(grid[d][o] for d, o in enumerate(origin))
        block = np.zeros(chunk, dtype=arr.dtype)
        sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(org, chunk, arr.shape))
        block[tuple(slice(0, s.stop - s.start) for s in sl)] = arr[sl]
        data = filtered_chunk(block.tobytes(), esize, True, True, False)
        entries.append((org, len(data), blob.alloc(data)))
    btree = chunk_btree(blob, arr.ndim, entries, esize, arr.shape, fanout=64)
    layout = struct.pack("<BBB", 3, 2, arr.ndim + 1) + struct.pack("<Q", btree) + b"".join(struct.pack("<I", c) for c in chunk) + struct.pack("<I", esize)
    # filter pipeline version 1: 8-byte header, each filter (id, name length, flags, n values, name padded to 8, values padded)
    def f1(fid, name, vals):
        nm = name + b"\x00"
        nm += b"\x00" * (-len(nm) % 8)
        raw = struct.pack("<HHHH", fid, len(nm), 1, len(vals)) + nm + b"".join(struct.pack("<I", v) for v in vals)
        return raw + (b"\x00" * 4 if len(vals) % 2 else b"")
    filt = struct.pack("<BB6x", 1, 2) + f1(2, b"shuffle", [esize]) + f1(1, b"deflate", [4])
    msgs = [(0x01, ds_simple_v1(arr.shape)), (0x03, dt_float(esize)), (0x0B, filt), (0x08, layout)] + [(0x0C, a) for a in attrs]
    return ohdr_v1(msgs, blob, split_after=3)


def build_earliest(rng) -> tuple:
    """aposteriori-shaped frame dataset with h5py's default ('earliest') structures: superblock 0, symbol tables."""
    blob = Blob(96)                                            # superblock v0 with 8-byte offsets: 56 + 40-byte root entry
    res_ids = [str(i) for i in (1, 2, 3, 10, 11, 20, 21, 22, 100, 101)]          # name order != integer order
    frames = {}
    children = []
    for r in res_ids:
        fr = rng.random((4, 4, 4, 2))
        frames[r] = fr
        lab = "ALA" if int(r) % 2 else "SER"
        attrs = [attr_v1("label", dt_fixed_string(4), ds_simple_v1(()), lab.encode() + b"\x00"),
                 attr_v1("encoded_residue", dt_float(8), ds_simple_v1((20,)), np.eye(20)[0 if lab == "ALA" else 15].astype("<f8").tobytes())]
        children.append((r, chunked_dataset_v1(blob, fr.astype("<f8"), (2, 4, 4, 2), attrs)))
    chain, _, _ = symbol_table_group(blob, children, per_node=4)                 # 10 entries -> 3 symbol-table nodes
    pdb, _, _ = symbol_table_group(blob, [("A", chain)])
    root_attrs = [attr_v1("make_frame_dataset_ver", dt_fixed_string(6), ds_simple_v1(()), b"2.0.0\x00"),
                  attr_v1("frame_dims", dt_int(8), ds_simple_v1((4,)), np.array([4, 4, 4, 2], "<i8").tobytes()),
                  attr_v1("voxels_as_gaussian", dt_bool_enum(1), ds_simple_v1(()), bytes([1]))]
    root, btree, heap = symbol_table_group(blob, [("2xyz", pdb)], root_attrs)
    # II.A superblock version 0: versions, sizes, group leaf/internal K, flags, base, free-space, EOF, driver, root symbol-table entry
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, 0, UNDEF)                # EOF patched below
    sb += struct.pack("<QQI4x", 0, root, 1) + struct.pack("<QQ", btree, heap)   # cache type 1: B-tree + heap addresses
    sb = bytearray(sb)
    sb[40:48] = struct.pack("<Q", len(blob.b))
    blob.patch(0, bytes(sb))
    return bytes(blob.b), {"res_ids": res_ids, "frames": {k: v.tolist() for k, v in frames.items()}}


def main():
    rng = np.random.default_rng(20261017)
    latest, e1 = build_latest(rng)
    earliest, e2 = build_earliest(rng)
    (HERE / "spec_latest.h5").write_bytes(latest)
    (HERE / "spec_earliest.h5").write_bytes(earliest)
    (HERE / "spec_expected.json").write_text(json.dumps({"latest": e1, "earliest": e2}))
    print(len(latest), len(earliest))


if __name__ == "__main__":
    main()
