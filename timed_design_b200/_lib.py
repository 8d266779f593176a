"""ctypes binding of libtimed_b200.so (C ABI declared in include/timed_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` /
``python -m timed_design_b200.build``.  There is no CPU fallback: if the library is missing,
or no sm_100 device is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libtimed_b200.so"

TB_MAX_INPUTS = 8
DTYPE_F32, DTYPE_F64, DTYPE_U8, DTYPE_F16 = 0, 1, 2, 3
ABI_VERSION = 4


class TimedB200Error(RuntimeError):
    pass


class tb_op_desc(C.Structure):
    _fields_ = [
        ("op", C.c_int32),
        ("n_inputs", C.c_int32),
        ("inputs", C.c_int32 * TB_MAX_INPUTS),
        ("kernel", C.c_int32 * 3),
        ("stride", C.c_int32 * 3),
        ("pad_same", C.c_int32),
        ("c_out", C.c_int32),
        ("pool_kind", C.c_int32),
        ("act1", C.c_int32),
        ("act2", C.c_int32),
        ("alpha1", C.c_float),
        ("alpha2", C.c_float),
        ("kernel_w", C.POINTER(C.c_float)),
        ("bias", C.POINTER(C.c_float)),
        ("scale", C.POINTER(C.c_float)),
        ("shift", C.POINTER(C.c_float)),
    ]


# name -> (restype, argtypes): every symbol include/timed_b200.h declares
SYMBOLS = {
    "timed_b200_abi_version": (C.c_int, []),
    "timed_b200_last_error": (C.c_char_p, []),
    "timed_b200_device_count": (C.c_int, []),
    "timed_b200_graph_create": (C.c_int, [C.POINTER(tb_op_desc), C.c_int32, C.c_int32,
                                          C.POINTER(C.c_void_p)]),
    "timed_b200_graph_destroy": (None, [C.c_void_p]),
    "timed_b200_graph_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                        C.POINTER(C.c_int32)]),
    "timed_b200_graph_set_precise": (C.c_int, [C.c_void_p, C.c_int32]),
    "timed_b200_graph_op_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "timed_b200_graph_op_kernel": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_char_p, C.c_int32]),
    "timed_b200_graph_set_timing": (C.c_int, [C.c_void_p, C.c_int32]),
    "timed_b200_graph_read_op_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32),
                                                 C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "timed_b200_graph_workspace_bytes": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_size_t)]),
    "timed_b200_graph_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                                           C.c_size_t, C.c_void_p, C.c_void_p]),
    "timed_b200_graph_predict_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                                C.c_void_p, C.c_int64]),
    "timed_b200_graph_predict_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "timed_b200_voxelise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_int32,
                                      C.POINTER(C.c_float), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                      C.c_void_p]),
    "timed_b200_conv3d_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.POINTER(tb_op_desc), C.c_int32, C.c_void_p]),
    "timed_b200_apply_temperature": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_double,
                                               C.c_void_p, C.c_void_p]),
    "timed_b200_cumsum_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "timed_b200_sample": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int64,
                                    C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "timed_b200_consensus_fp16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "timed_b200_seq_metrics": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_void_p]),
    "timed_b200_sample_chains": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int64,
                                           C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "timed_b200_inflate_chunks": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                            C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                            C.c_int32]),
    "timed_b200_format_csv_e18": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                            C.POINTER(C.c_int64), C.c_int32]),
    "timed_b200_parse_csv": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int64), C.c_int32]),
    "timed_b200_sample_uniforms": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64,
                                             C.c_void_p, C.c_void_p]),
    "timed_b200_argmax_fp16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "timed_b200_hdf5_frame_index": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_char_p, C.c_int32,
                                              C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32,
                                              C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "timed_b200_inflate_device": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "timed_b200_pdb_parse": (C.c_int, [C.POINTER(C.c_char_p), C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "timed_b200_pdb_sizes": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "timed_b200_pdb_export": (C.c_int, [C.c_void_p] + [C.c_void_p] * 13),
    "timed_b200_pdb_free": (None, [C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and attach prototypes.  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("TIMED_B200_LIB", LIB_PATH))
    if not path.exists():
        raise TimedB200Error(
            f"{path} not found: build it with `python -m timed_design_b200.build` "
            "(there is no CPU fallback for this path)")
    lib = C.CDLL(str(path))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.timed_b200_abi_version() != ABI_VERSION:
        raise TimedB200Error("libtimed_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().timed_b200_last_error().decode("utf-8", "replace")
        raise TimedB200Error(f"libtimed_b200 error {rc}: {msg}")


def require_device() -> int:
    n = load().timed_b200_device_count()
    if n <= 0:
        raise TimedB200Error("no CUDA device visible: the timed-design B200 path has no CPU fallback")
    return n


def np_dtype_code(arr: np.ndarray) -> int:
    if arr.dtype == np.float32:
        return DTYPE_F32
    if arr.dtype == np.float64:
        return DTYPE_F64
    if arr.dtype in (np.bool_, np.uint8):
        return DTYPE_U8
    if arr.dtype == np.float16:
        return DTYPE_F16
    raise TypeError(f"frames dtype {arr.dtype} not supported (float32, float64, float16, bool, uint8)")


def fptr(arr):
    return None if arr is None else arr.ctypes.data_as(C.POINTER(C.c_float))


def make_op_array(graph):
    """keras_graph.Graph -> (ctypes array of tb_op_desc, keep-alive list of numpy buffers)."""
    n = len(graph.ops)
    arr = (tb_op_desc * n)()
    keep = []
    for i, op in enumerate(graph.ops):
        d = arr[i]
        d.op = op.kind
        d.n_inputs = len(op.inputs)
        for k, v in enumerate(op.inputs):
            d.inputs[k] = v
        for k in range(3):
            d.kernel[k] = int(op.kernel[k])
            d.stride[k] = int(op.stride[k])
        d.pad_same = int(op.pad_same)
        d.c_out = int(op.c_out)
        d.pool_kind = int(op.pool_kind)
        d.act1, d.act2 = int(op.act1), int(op.act2)
        d.alpha1, d.alpha2 = float(op.alpha1), float(op.alpha2)
        for field in ("kernel_w", "bias", "scale", "shift"):
            a = getattr(op, field)
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float32)
                keep.append(a)
                setattr(d, field, fptr(a))
    return arr, keep
