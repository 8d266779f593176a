"""Device-side post-processing of the two "next" rows of SURVEY.md 8(f) that sit directly on the hot path's outputs:

* ``nmr_consensus``   -- the running pairwise float16 mean over the states of an NMR ensemble and its argmax
  (``/root/reference/design_utils/utils.py:694-713``), one warp per consensus row (``timed_b200_consensus_fp16``);
* ``seq_metrics``     -- charge / isoelectric point / molecular weight / molar extinction of every sampled sequence
  (``/root/reference/design_utils/analyse_utils.py:351-371``, called inside the reference's Monte-Carlo loop at
  ``sampling_utils.py:132``), one warp per sequence straight from the sampler's letter block in HBM
  (``timed_b200_seq_metrics``).

torch tensors are device-memory containers only.  The numpy restatements these are checked against live in
``postprocess.extract_sequence_from_pred_matrix`` (pinned by golden vectors from the reference's own function) and
``seq_metrics.metrics_from_composition`` (tables unverified against ampal -- SURVEY.md App. G)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, seq_metrics
from .postprocess import _letters_for


def _torch():
    import torch
    _lib.require_device()
    return torch


def _ptr(t):
    return C.c_void_p(t.data_ptr())


# ----------------------------------------------------------------------------- NMR consensus
def consensus_groups(keys, lengths):
    """Group consecutive chain keys by ``key.split('_')[0]`` exactly as utils.py:696-705 walks the dict:
    -> list of (structure, first_index_in_keys, n_states).  States of one structure must be equally long."""
    groups = []
    last = None
    for i, k in enumerate(keys):
        cur = k.split("_")[0]
        if cur != last:
            groups.append([cur, i, 1])
            last = cur
        else:
            if lengths[i] != lengths[groups[-1][1]]:
                raise ValueError(f"NMR states of {cur} differ in length ({lengths[i]} vs {lengths[groups[-1][1]]})")
            groups[-1][2] += 1
    return [tuple(g) for g in groups]


def nmr_consensus(pdb_to_probability: dict, rotamers_categories=None):
    """-> (pdb_to_consensus {structure: sequence}, pdb_to_consensus_prob {structure: (n_res, C) float16}).

    ``pdb_to_probability``: ordered {chain key: (n_res, C) probabilities}; values are rounded to float16 first (they
    already are float16 on the reference's path, predict.py:163)."""
    torch = _torch()
    keys = list(pdb_to_probability.keys())
    if not keys:
        return {}, {}
    mats = [np.asarray(pdb_to_probability[k]) for k in keys]
    lengths = [len(m) for m in mats]
    groups = consensus_groups(keys, lengths)
    # a structure that re-appears later in the dict overwrites its earlier consensus in the reference (dict
    # assignment); keep that by letting later groups win
    n_cls = mats[0].shape[1]
    starts = np.concatenate([[0], np.cumsum(lengths)])
    probs = np.concatenate([m.astype(np.float16).astype(np.float32) for m in mats], axis=0)
    first_row = np.array([starts[g[1]] for g in groups], dtype=np.int64)
    n_states = np.array([g[2] for g in groups], dtype=np.int32)
    n_res = np.array([lengths[g[1]] for g in groups], dtype=np.int32)
    out_row0 = np.concatenate([[0], np.cumsum(n_res)[:-1]]).astype(np.int64)
    n_out = int(n_res.sum())
    if n_out == 0:
        return {g[0]: "" for g in groups}, {g[0]: np.zeros((0, n_cls), np.float16) for g in groups}
    keep = n_res > 0                                   # empty groups own no output rows
    d = {name: torch.from_numpy(np.ascontiguousarray(a[keep])).cuda()
         for name, a in (("first", first_row), ("states", n_states), ("res", n_res), ("out0", out_row0))}
    d_probs = torch.from_numpy(np.ascontiguousarray(probs)).cuda()
    d_cons = torch.empty((n_out, n_cls), dtype=torch.float16, device="cuda")
    d_idx = torch.empty(n_out, dtype=torch.int32, device="cuda")
    _lib.check(_lib.load().timed_b200_consensus_fp16(
        _ptr(d_probs), probs.shape[0], n_cls, _ptr(d["first"]), _ptr(d["states"]), _ptr(d["res"]), _ptr(d["out0"]),
        int(keep.sum()), n_out, _ptr(d_cons), _ptr(d_idx), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    cons = d_cons.cpu().numpy()
    idx = d_idx.cpu().numpy()
    letters = _letters_for(rotamers_categories)
    pdb_to_consensus, pdb_to_consensus_prob = {}, {}
    for g, r0, n in zip(groups, out_row0, n_res):
        pdb_to_consensus_prob[g[0]] = cons[r0:r0 + n]
        pdb_to_consensus[g[0]] = "".join(letters[idx[r0:r0 + n]])
    return pdb_to_consensus, pdb_to_consensus_prob


# ----------------------------------------------------------------------------- sequence metrics
_tables_cache = {}


def _metric_tables(torch):
    if "t" not in _tables_cache:
        _tables_cache["t"] = torch.from_numpy(seq_metrics.device_tables()).cuda()
        lut = seq_metrics._LUT.astype(np.int8)
        _tables_cache["lut"] = torch.from_numpy(lut).cuda()
    return _tables_cache["t"], _tables_cache["lut"]


def seq_metrics_device(d_seqs, n_seqs: int, n_res: int) -> np.ndarray:
    """d_seqs: CUDA uint8 tensor holding (n_seqs, n_res) ASCII letters -> (n_seqs, 4) float64 host array
    [charge, isoelectric point, molecular weight, molar extinction]."""
    torch = _torch()
    from . import seq_metrics
    seq_metrics.warn_unverified()
    tables, lut = _metric_tables(torch)
    out = torch.empty((n_seqs, 4), dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().timed_b200_seq_metrics(_ptr(d_seqs), int(n_seqs), int(n_res), _ptr(lut), _ptr(tables),
                                                  int(tables.numel()), _ptr(out),
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    m = out.cpu().numpy()
    if np.isnan(m).any():
        raise ValueError("sequence contains a non-standard residue letter")
    return m


def seq_metrics_of(seqs_u8: np.ndarray) -> np.ndarray:
    """Host (n_seqs, n_res) uint8 letters -> (n_seqs, 4) metrics through the device kernel."""
    torch = _torch()
    a = np.ascontiguousarray(seqs_u8, dtype=np.uint8)
    if a.size == 0:
        return np.zeros((a.shape[0], 4))
    return seq_metrics_device(torch.from_numpy(a).cuda(), a.shape[0], a.shape[1])
