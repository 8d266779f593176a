"""Minimal pure-Python HDF5 writer -- produces the classic on-disk structures h5py writes with
default settings (superblock v0, v1 object headers, symbol-table groups with v1 B-tree + local
heap, contiguous or gzip-chunked datasets with a v1 B-tree chunk index, version-1 attribute
messages, variable-length strings in a global heap).  Used to build synthetic Keras ``.h5``
models and aposteriori-style frame datasets (no h5py / libhdf5 in this image, no real files
offline).  Written from the HDF5 File Format Specification v3.0.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K = 4          # group leaf node K  -> up to 8 symbols per SNOD
INTERNAL_K = 16     # group/chunk internal node K -> up to 32 children per TREE node


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


class VLenStr(str):
    """Marks an attribute value to be stored as a variable-length UTF-8 string (h5py's str)."""


class _Node:
    def __init__(self):
        self.attrs: Dict[str, object] = {}


class GroupSpec(_Node):
    def __init__(self):
        super().__init__()
        self.children: Dict[str, Union["GroupSpec", "DatasetSpec"]] = {}

    def group(self, name: str) -> "GroupSpec":
        node = self
        for part in [p for p in name.split("/") if p]:
            nxt = node.children.get(part)
            if nxt is None:
                nxt = node.children[part] = GroupSpec()
            node = nxt
        return node

    def dataset(self, name: str, data, compression: Optional[str] = None, chunks=None) -> "DatasetSpec":
        parts = [p for p in name.split("/") if p]
        parent = self.group("/".join(parts[:-1])) if len(parts) > 1 else self
        ds = parent.children[parts[-1]] = DatasetSpec(np.asarray(data), compression, chunks)
        return ds


class DatasetSpec(_Node):
    def __init__(self, data: np.ndarray, compression, chunks):
        super().__init__()
        self.data, self.compression, self.chunks = data, compression, chunks


class Writer:
    """Build a tree with ``root.group()/dataset()`` + ``.attrs`` then ``save(path)``."""

    def __init__(self):
        self.root = GroupSpec()
        self.buf = bytearray()
        self._gcol_items: List[bytes] = []

    # ------------------------------------------------------------------ allocation
    def _alloc(self, data: bytes, align: int = 8) -> int:
        self.buf += b"\x00" * (-len(self.buf) % align)
        addr = len(self.buf)
        self.buf += data
        return addr

    # ------------------------------------------------------------------ datatype / dataspace
    @staticmethod
    def _dtype_msg(dt: np.dtype) -> bytes:
        dt = np.dtype(dt)
        if dt == np.bool_:
            # h5py: ENUM {FALSE=0, TRUE=1} over int8 (datatype version 1: names padded to 8)
            base = Writer._dtype_msg(np.dtype("i1"))
            body = base + _pad8(b"FALSE\x00") + _pad8(b"TRUE\x00") + bytes([0, 1])
            return struct.pack("<BBBBI", 0x10 | 8, 2, 0, 0, 1) + body
        if dt.kind in "iu":
            flags = 0x08 if dt.kind == "i" else 0
            return struct.pack("<BBBBIHH", 0x10 | 0, flags, 0, 0, dt.itemsize, 0, dt.itemsize * 8)
        if dt.kind == "f":
            if dt.itemsize == 4:
                props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
                b1 = 31
            elif dt.itemsize == 8:
                props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
                b1 = 63
            elif dt.itemsize == 2:
                props = struct.pack("<HHBBBBI", 0, 16, 10, 5, 0, 10, 15)
                b1 = 15
            else:
                raise ValueError(dt)
            # bit field: byte order LE, padding 0, mantissa normalisation = implied (2<<4), sign location
            return struct.pack("<BBBBI", 0x10 | 1, 0x20, b1, 0, dt.itemsize) + props
        if dt.kind == "S":
            return struct.pack("<BBBBI", 0x10 | 3, 0x01, 0, 0, dt.itemsize)      # null-padded ASCII
        raise ValueError(f"unsupported dtype {dt}")

    @staticmethod
    def _vlen_str_dtype_msg() -> bytes:
        base = struct.pack("<BBBBI", 0x10 | 3, 0x10, 0, 0, 1)      # 1-byte UTF-8 string base
        # class 9, type=string(1), padding null-term(0), charset UTF-8(1)
        return struct.pack("<BBBBI", 0x10 | 9, 0x01, 0x01, 0, 16) + base

    @staticmethod
    def _space_msg(shape: Tuple[int, ...]) -> bytes:
        rank = len(shape)
        return struct.pack("<BBB5x", 1, rank, 0) + b"".join(struct.pack("<Q", s) for s in shape)

    # ------------------------------------------------------------------ global heap (vlen strings)
    def _gcol_add(self, data: bytes) -> int:
        self._gcol_items.append(data)
        return len(self._gcol_items)          # 1-based object index

    def _flush_gcol(self) -> None:
        body = bytearray()
        for i, d in enumerate(self._gcol_items, start=1):
            body += struct.pack("<HHIQ", i, 1, 0, len(d)) + _pad8(d)
        size = 16 + len(body) + 16
        size = max(size, 4096)
        free = size - 16 - len(body)
        blob = b"GCOL" + struct.pack("<B3xQ", 1, size) + bytes(body) + struct.pack("<HHIQ", 0, 0, 0, free)
        blob += b"\x00" * (size - len(blob))
        assert len(blob) == size
        # patch: the collection must live at the address the references already point to
        self.buf[self._gcol_addr:self._gcol_addr + size] = blob

    # ------------------------------------------------------------------ attributes
    def _attr_msg(self, name: str, value) -> bytes:
        nm = name.encode("utf-8") + b"\x00"
        if isinstance(value, VLenStr) or (isinstance(value, str)):
            idx = self._gcol_add(str(value).encode("utf-8"))
            dt = self._vlen_str_dtype_msg()
            sp = self._space_msg(())
            data = struct.pack("<IQI", len(str(value).encode("utf-8")), self._gcol_addr, idx)
        else:
            arr = np.asarray(value)
            if arr.dtype.kind == "U":
                arr = np.char.encode(arr, "utf-8")
            if arr.dtype == object:
                raise ValueError(f"attribute {name}: object arrays are not supported")
            if arr.dtype.kind == "f" and arr.dtype.itemsize not in (2, 4, 8):
                arr = arr.astype(np.float64)
            dt = self._dtype_msg(arr.dtype)
            sp = self._space_msg(arr.shape)
            data = np.ascontiguousarray(arr).astype(arr.dtype.newbyteorder("<")).tobytes() \
                if arr.dtype != np.bool_ else np.ascontiguousarray(arr).astype("i1").tobytes()
        return struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + data

    # ------------------------------------------------------------------ object headers
    def _object_header(self, messages: List[Tuple[int, bytes]]) -> int:
        body = bytearray()
        for mtype, data in messages:
            if len(data) > 0xFFF8:
                raise ValueError("object header message larger than 64 KiB (HDF5 limit; h5py fails too)")
            d = _pad8(data)
            body += struct.pack("<HHB3x", mtype, len(d), 0) + d
        hdr = struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body))
        return self._alloc(hdr + bytes(body))

    # ------------------------------------------------------------------ groups
    def _write_group(self, g: GroupSpec) -> int:
        names = sorted(g.children)                     # B-tree order = name order
        child_addr = {}
        for n in names:
            c = g.children[n]
            child_addr[n] = self._write_group(c) if isinstance(c, GroupSpec) else self._write_dataset(c)
        # local heap: offset 0 holds the empty string, names follow
        heap = bytearray(b"\x00" * 8)
        name_off = {}
        for n in names:
            name_off[n] = len(heap)
            heap += _pad8(n.encode("utf-8") + b"\x00")
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 16)              # free block: next=1 (none), size 16
        heap_data_addr = self._alloc(bytes(heap))
        heap_addr = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data_addr))
        # symbol nodes (leaves) of up to 2*LEAF_K entries
        per = 2 * LEAF_K
        leaves = []
        for i in range(0, max(len(names), 1), per):
            part = names[i:i + per]
            snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(part)))
            for n in part:
                snod += struct.pack("<QQII16x", name_off[n], child_addr[n], 0, 0)
            snod += b"\x00" * (40 * (per - len(part)))
            leaves.append((self._alloc(bytes(snod)), name_off[part[-1]] if part else 0))

        def build(level: int, items: List[Tuple[int, int]]) -> Tuple[int, int]:
            """items: (child address, heap offset of the largest name below) -> (node addr, max key)."""
            node = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, level, len(items), UNDEF, UNDEF))
            node += struct.pack("<Q", 0)               # key 0: empty string (smaller than everything)
            for addr, key in items:
                node += struct.pack("<QQ", addr, key)
            node += b"\x00" * (16 * (2 * INTERNAL_K - len(items)))
            return self._alloc(bytes(node)), items[-1][1]

        level = 0
        nodes = leaves
        while True:
            groups = [nodes[i:i + 2 * INTERNAL_K] for i in range(0, len(nodes), 2 * INTERNAL_K)]
            nodes = [build(level, grp) for grp in groups]
            if len(nodes) == 1:
                break
            level += 1
        btree_addr = nodes[0][0]
        msgs = [(0x0011, struct.pack("<QQ", btree_addr, heap_addr))]
        msgs += [(0x000C, self._attr_msg(k, v)) for k, v in g.attrs.items()]
        return self._object_header(msgs)

    # ------------------------------------------------------------------ datasets
    def _write_dataset(self, ds: DatasetSpec) -> int:
        arr = np.ascontiguousarray(ds.data)
        store = arr.astype("i1") if arr.dtype == np.bool_ else arr.astype(arr.dtype.newbyteorder("<"))
        msgs = [(0x0001, self._space_msg(arr.shape)), (0x0003, self._dtype_msg(arr.dtype)),
                (0x0005, struct.pack("<BBBB", 2, 2, 2, 0))]          # fill value v2: undefined
        if ds.compression is None:
            addr = self._alloc(store.tobytes()) if store.size else UNDEF
            msgs.append((0x0008, struct.pack("<BBQQ", 3, 1, addr, store.nbytes)))
        else:
            if ds.compression != "gzip":
                raise ValueError("only gzip compression is supported")
            rank = arr.ndim
            chunks = tuple(ds.chunks) if ds.chunks else tuple(arr.shape)
            entries = []
            grid = [range(0, arr.shape[d], chunks[d]) for d in range(rank)]
            for offs in np.ndindex(*[len(r) for r in grid]):
                start = tuple(grid[d][offs[d]] for d in range(rank))
                block = np.zeros(chunks, dtype=store.dtype)
                sl = tuple(slice(s, min(s + c, n)) for s, c, n in zip(start, chunks, arr.shape))
                block[tuple(slice(0, s.stop - s.start) for s in sl)] = store[sl]
                comp = zlib.compress(block.tobytes(), 4)
                entries.append((start, len(comp), self._alloc(comp)))
            if len(entries) > 2 * INTERNAL_K:
                raise ValueError("too many chunks for a single-level chunk B-tree in this writer")
            node = bytearray(b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), UNDEF, UNDEF))
            for start, csize, caddr in entries:
                node += struct.pack("<II", csize, 0) + b"".join(struct.pack("<Q", s) for s in start) + \
                    struct.pack("<Q", 0) + struct.pack("<Q", caddr)
            # final key: one past the last chunk
            node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", n) for n in arr.shape) + struct.pack("<Q", 0)
            key_size = 8 + 8 * (rank + 1)
            node += b"\x00" * ((2 * INTERNAL_K - len(entries)) * (key_size + 8))
            btree = self._alloc(bytes(node))
            msgs.append((0x000B, struct.pack("<BB6x", 1, 1) +
                         struct.pack("<HHHH", 1, 8, 1, 1) + _pad8(b"deflate\x00") + struct.pack("<II", 4, 0)))
            lay = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", btree) + \
                b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", store.dtype.itemsize)
            msgs.append((0x0008, lay))
        msgs += [(0x000C, self._attr_msg(k, v)) for k, v in ds.attrs.items()]
        return self._object_header(msgs)

    # ------------------------------------------------------------------ file
    def save(self, path) -> None:
        self.buf = bytearray(b"\x00" * 96)             # superblock v0 placeholder
        # reserve the global heap collection up front so vlen references know its address
        total = sum(16 + len(_pad8(str(v).encode("utf-8"))) for v in self._iter_str_attrs(self.root))
        size = max(4096, 16 + total + 16 + 64)
        self._gcol_addr = self._alloc(b"\x00" * size)
        self._gcol_size = size
        self._gcol_items = []
        root_addr = self._write_group(self.root)
        body = bytearray()
        for i, d in enumerate(self._gcol_items, start=1):
            body += struct.pack("<HHIQ", i, 1, 0, len(d)) + _pad8(d)
        free = size - 16 - len(body)
        assert free >= 16
        blob = b"GCOL" + struct.pack("<B3xQ", 1, size) + bytes(body) + struct.pack("<HHIQ", 0, 0, 0, free)
        blob += b"\x00" * (size - len(blob))
        self.buf[self._gcol_addr:self._gcol_addr + size] = blob
        eof = len(self.buf)
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII16x", 0, root_addr, 0, 0)         # root symbol table entry
        assert len(sb) == 96
        self.buf[0:96] = sb
        with open(path, "wb") as f:
            f.write(bytes(self.buf))

    def _iter_str_attrs(self, node):
        for v in node.attrs.values():
            if isinstance(v, str):
                yield v
        if isinstance(node, GroupSpec):
            for c in node.children.values():
                yield from self._iter_str_attrs(c)
