"""``load_model(path)`` / ``Model.predict(X)`` -- the drop-in for the two Keras calls on the
reference's hot path (``tf.keras.models.load_model`` at predict.py:121 and
``frame_model.predict(X_batch)`` at predict.py:142).

Same contract as the reference call sites: ``predict`` is synchronous, takes a host numpy
array ``(B, D, H, W, C)`` of float64 / float32 / float16 / bool and returns a host ``float32``
``(B, n_classes)`` array; exceptions propagate.  All arithmetic runs in libtimed_b200.so on the
B200 -- there is no CPU path; without the library or a device the calls raise.
"""
from __future__ import annotations

import ctypes as C
import json
from pathlib import Path
from typing import Dict, Optional

import numpy as np

from . import _lib
from .keras_graph import Graph, parse_model_config


class Model:
    """An inference graph resident on one GPU."""

    def __init__(self, model_config: dict, weights: Dict[str, Dict[str, np.ndarray]],
                 device: int = 0, max_chunk_frames: int = 1024, precise: bool = True):
        self.model_config = model_config
        self.graph: Graph = parse_model_config(model_config, weights)
        self.device = device
        self.max_chunk_frames = int(max_chunk_frames)
        self.name = self.graph.name
        self.input_shape = self.graph.input_shape
        self.n_classes = self.graph.n_classes
        self._h = C.c_void_p()
        lib = _lib.load()
        _lib.require_device()
        ops, keep = _lib.make_op_array(self.graph)
        _lib.check(lib.timed_b200_graph_create(ops, len(self.graph.ops), device, C.byref(self._h)))
        del keep
        ncls, flops, launches = C.c_int32(), C.c_double(), C.c_int32()
        _lib.check(lib.timed_b200_graph_info(self._h, C.byref(ncls), C.byref(flops), C.byref(launches)))
        assert ncls.value == self.n_classes
        self.flops_per_frame = flops.value
        self.launches_per_forward = launches.value
        self.precise = True                     # the library's default
        if not precise:
            self.set_precise(False)

    # ------------------------------------------------------------------ host path (drop-in)
    def predict(self, X: np.ndarray, batch_size: Optional[int] = None, verbose: int = 0) -> np.ndarray:
        """Keras-compatible signature; ``batch_size`` bounds the frames per device chunk."""
        X = np.asarray(X)
        if X.ndim != 5 or tuple(X.shape[1:]) != tuple(self.input_shape):
            raise ValueError(f"expected input of shape (B, {', '.join(map(str, self.input_shape))}), "
                             f"got {X.shape}")
        if X.dtype not in (np.float32, np.float64, np.float16, np.bool_, np.uint8):
            X = X.astype(np.float32)
        X = np.ascontiguousarray(X)
        n = X.shape[0]
        out = np.empty((n, self.n_classes), dtype=np.float32)
        if n == 0:
            return out
        chunk = self.max_chunk_frames if not batch_size else min(int(batch_size), self.max_chunk_frames)
        _lib.check(_lib.load().timed_b200_graph_predict_host(
            self._h, X.ctypes.data_as(C.c_void_p), _lib.np_dtype_code(X), n,
            out.ctypes.data_as(C.c_void_p), chunk))
        return out

    __call__ = predict

    def predict_stats(self):
        """(device passes, largest pass in frames) of the last ``predict`` call."""
        a, b = C.c_int64(), C.c_int64()
        _lib.check(_lib.load().timed_b200_graph_predict_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ------------------------------------------------------------------ device-resident path
    def workspace_bytes(self, n_frames: int) -> int:
        out = C.c_size_t()
        _lib.check(_lib.load().timed_b200_graph_workspace_bytes(self._h, int(n_frames), C.byref(out)))
        return out.value

    def forward_device(self, frames, probs, workspace, stream: int = 0) -> None:
        """Enqueue one forward on device buffers (torch tensors used as containers only).
        ``frames``: (n,D,H,W,C) float32/float64/uint8 CUDA tensor; ``probs``: (n,classes) float32;
        ``workspace``: uint8 CUDA tensor of at least ``workspace_bytes(n)`` bytes."""
        import torch
        code = {torch.float32: _lib.DTYPE_F32, torch.float64: _lib.DTYPE_F64, torch.float16: _lib.DTYPE_F16,
                torch.uint8: _lib.DTYPE_U8, torch.bool: _lib.DTYPE_U8}[frames.dtype]
        n = frames.shape[0]
        assert frames.is_contiguous() and probs.is_contiguous() and probs.dtype == torch.float32
        assert tuple(frames.shape[1:]) == tuple(self.input_shape) and tuple(probs.shape) == (n, self.n_classes)
        _lib.check(_lib.load().timed_b200_graph_forward(
            self._h, C.c_void_p(frames.data_ptr()), code, n, C.c_void_p(workspace.data_ptr()),
            workspace.numel() * workspace.element_size(), C.c_void_p(probs.data_ptr()),
            C.c_void_p(stream)))

    def set_precise(self, precise: bool) -> None:
        """Default True: the wide conv layers accumulate the bf16-split correction products in their own TMEM accumulator
        (3x less accumulator truncation: max |dp| 2.2e-5 instead of 5.5e-5 on the TIMED-20 stand-in, so the 1e-4 contract
        holds with margin on sharper networks); their epilogue drains TMEM into registers and releases it before the
        activation math, so the single accumulator stage costs < 1 %.  False: corrections go into the main accumulator
        (two accumulator stages) -- an A/B switch, not a production mode."""
        _lib.check(_lib.load().timed_b200_graph_set_precise(self._h, int(bool(precise))))
        self.precise = bool(precise)

    def set_timing(self, enabled: bool) -> None:
        """Record CUDA events around every fused op of subsequent forwards (bench roofline)."""
        _lib.check(_lib.load().timed_b200_graph_set_timing(self._h, int(bool(enabled))))

    def read_op_times(self):
        """-> (list of dicts per op: kind, name, ms (summed), flops_per_frame), n_forwards)."""
        n = len(self.graph.ops)
        ms = (C.c_float * n)()
        kinds = (C.c_int32 * n)()
        flops = (C.c_double * n)()
        nf = C.c_int32()
        _lib.check(_lib.load().timed_b200_graph_read_op_times(self._h, ms, kinds, flops, C.byref(nf)))
        ops = [{"index": i, "kind": int(kinds[i]), "name": self.graph.ops[i].name, "ms": float(ms[i]),
                "flops_per_frame": float(flops[i])} for i in range(n)]
        return ops, nf.value

    def op_kernel(self, op_index: int, n_frames: int) -> str:
        """Name of the CUDA kernel(s) fused op `op_index` launches at this batch size."""
        buf = C.create_string_buffer(160)
        _lib.check(_lib.load().timed_b200_graph_op_kernel(self._h, int(op_index), int(n_frames), buf, 160))
        return buf.value.decode()

    def close(self) -> None:
        if self._h:
            _lib.load().timed_b200_graph_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def save_npz(path, model_config: dict, weights: Dict[str, Dict[str, np.ndarray]]) -> None:
    """numpy-native container (escape hatch where h5py exists to convert a real .h5)."""
    arrays = {"__model_config__": np.frombuffer(json.dumps(model_config).encode(), dtype=np.uint8)}
    for layer, ws in weights.items():
        for k, v in ws.items():
            arrays[f"{layer}/{k}"] = np.asarray(v)
    np.savez(path, **arrays)


def _load_npz(path):
    z = np.load(path)
    cfg = json.loads(bytes(z["__model_config__"]).decode())
    weights: Dict[str, Dict[str, np.ndarray]] = {}
    for key in z.files:
        if key == "__model_config__":
            continue
        layer, wname = key.split("/", 1)
        weights.setdefault(layer, {})[wname] = z[key]
    return cfg, weights


def read_model_file(path):
    """(model_config, weights) from a Keras ``.h5`` (own HDF5 reader) or an ``.npz`` container."""
    path = Path(path)
    if path.suffix == ".npz":
        return _load_npz(path)
    from .hdf5 import read_keras_h5
    return read_keras_h5(path)


def load_model(path, custom_objects=None, compile: bool = False, device: int = 0, **kw) -> Model:
    """Drop-in for ``tf.keras.models.load_model(Path(m))`` (predict.py:121).  ``custom_objects``
    / ``compile`` are accepted and ignored: the reference only registers ``top_3_cat_acc``
    (predict.py:24-25,88) so that Keras can deserialise a training metric."""
    cfg, weights = read_model_file(path)
    m = Model(cfg, weights, device=device, **kw)
    m.name = Path(path).stem
    return m
