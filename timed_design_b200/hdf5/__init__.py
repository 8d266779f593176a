"""Pure-Python HDF5 for the two containers on the path: Keras-2 ``.h5`` model files
(``tf.keras.models.load_model``, predict.py:121) and aposteriori frame datasets
(design_utils/utils.py:238-251).  ``reader.File`` mimics the small part of h5py's API the
reference uses (``f[pdb][chain][res][()]``, ``.attrs[...]``, ``.keys()``, iteration)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, Tuple

import numpy as np

from .reader import File, Hdf5FormatError  # noqa: F401
from .writer import Writer

_CHUNK_ATTR_BYTES = 64512      # Keras' HDF5_OBJECT_HEADER_LIMIT: longer name lists are split


def _as_str(x) -> str:
    if isinstance(x, bytes):
        return x.decode("utf-8")
    if isinstance(x, np.bytes_):
        return bytes(x).decode("utf-8")
    return str(x)


def _load_attr_list(attrs, name: str):
    """Keras ``load_attributes_from_hdf5_group``: ``name`` or the chunks ``name0, name1, ...``."""
    if name in attrs:
        return [_as_str(n) for n in np.atleast_1d(attrs[name])]
    out, i = [], 0
    while f"{name}{i}" in attrs:
        out.extend(_as_str(n) for n in np.atleast_1d(attrs[f"{name}{i}"]))
        i += 1
    return out


def read_keras_h5(path) -> Tuple[dict, Dict[str, Dict[str, np.ndarray]]]:
    """(model_config dict, {layer: {weight_name: float32 array}}) from a Keras 2.x HDF5 model
    (SURVEY.md App. B): root attr ``model_config`` (JSON), group ``model_weights`` with attr
    ``layer_names`` and one sub-group per layer carrying ``weight_names``."""
    f = File(path)
    if "model_config" not in f.attrs:
        raise Hdf5FormatError(f"{path}: no 'model_config' attribute (weights-only file? Keras SavedModel?)")
    cfg = json.loads(_as_str(f.attrs["model_config"]))
    mw = f["model_weights"] if "model_weights" in f else f.root
    weights: Dict[str, Dict[str, np.ndarray]] = {}
    for layer in _load_attr_list(mw.attrs, "layer_names"):
        g = mw[layer]
        names = _load_attr_list(g.attrs, "weight_names")
        if not names:
            continue
        weights[layer] = {n: np.asarray(g[n][()], dtype=np.float32) for n in names}
    return cfg, weights


def write_keras_h5(path, model_config: dict, weights: Dict[str, Dict[str, np.ndarray]],
                   keras_version: str = "2.13.1") -> None:
    """Write ``(model_config, weights)`` in the layout Keras 2.x ``model.save('x.h5')`` uses."""
    w = Writer()
    w.root.attrs["keras_version"] = keras_version
    w.root.attrs["backend"] = "tensorflow"
    w.root.attrs["model_config"] = json.dumps(model_config)
    mw = w.root.group("model_weights")
    layer_names = [l["config"]["name"] if "name" not in l else l["name"]
                   for l in model_config["config"]["layers"]]
    mw.attrs["layer_names"] = np.array([n.encode() for n in layer_names])
    mw.attrs["backend"] = "tensorflow"
    mw.attrs["keras_version"] = keras_version
    for layer in layer_names:
        g = mw.group(layer)
        ws = weights.get(layer, {})
        full = [k if "/" in k else f"{layer}/{k}" for k in ws]
        g.attrs["weight_names"] = np.array([n.encode() for n in full]) if full else np.zeros((0,), "S1")
        for name, arr in zip(full, ws.values()):
            g.dataset(name, np.asarray(arr, dtype=np.float32))
    w.save(path)


def write_frame_dataset(path, frames: Dict[str, Dict[str, Dict[str, Tuple[np.ndarray, str]]]],
                        frame_dims, voxels_as_gaussian: bool = True, atom_encoder=("C", "N", "O", "CB", "CA", "Q"),
                        residue_encoder=None, frame_edge_length: float = 21.0, compression="gzip",
                        version: str = "2.0.0") -> None:
    """aposteriori-style frame dataset: ``/{pdb}/{chain}/{res_id}`` datasets of
    shape ``frame_dims`` with attrs ``label`` (three-letter code) and ``encoded_residue``
    (20-dim one-hot); root attrs as listed at design_utils/utils.py:238-251."""
    from ..postprocess import standard_amino_acids
    order = list(standard_amino_acids.values()) if residue_encoder is None else list(residue_encoder)
    w = Writer()
    w.root.attrs["make_frame_dataset_ver"] = version
    w.root.attrs["frame_dims"] = np.array(frame_dims, dtype=np.int64)
    w.root.attrs["atom_encoder"] = np.array([a.encode() for a in atom_encoder])
    w.root.attrs["encode_cb"] = np.bool_(True)
    w.root.attrs["atom_filter_fn"] = "keep_sidechain_cb_atom_filter"
    w.root.attrs["residue_encoder"] = np.array([r.encode() for r in order])
    w.root.attrs["frame_edge_length"] = np.float64(frame_edge_length)
    w.root.attrs["voxels_as_gaussian"] = np.bool_(voxels_as_gaussian)
    for pdb, chains in frames.items():
        for chain, residues in chains.items():
            for res_id, (arr, label) in residues.items():
                data = np.asarray(arr, dtype=np.float32 if voxels_as_gaussian else np.bool_)
                ds = w.root.dataset(f"{pdb}/{chain}/{res_id}", data, compression=compression)
                ds.attrs["label"] = label
                ds.attrs["encoded_residue"] = np.eye(20, dtype=np.float64)[order.index(label)]
    w.save(path)
