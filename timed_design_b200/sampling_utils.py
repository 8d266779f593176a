"""Monte-Carlo sequence sampler on the GPU -- same function names and argument meaning as
``/root/reference/design_utils/sampling_utils.py`` (lines cited per function).

What changes underneath: instead of ``sample_n`` Python iterations per chain, each recomputing
``probs.cumsum`` and drawing ``n_res`` uniforms from the global legacy numpy RNG inside a forked
``multiprocessing.Pool``, one chain's whole (sample_n, n_res) block is drawn by a single kernel
launch through libtimed_b200 (``timed_b200_sample``): the float64 CDF is built once on the
device by a strictly sequential per-row cumsum (bit-identical to numpy), uniforms come from
Philox4x32-10 keyed (seed, chain) with counter (sample, residue), and the letters land as one
(sample_n, n_res) uint8 block.  torch tensors are used only as device-memory containers.

Kept quirks (parity): a row whose cumsum never exceeds r yields class 0 ('A'); temperature 1
skips renormalisation (that lives in sample.py).  Documented deviations: ``--seed`` is effective
(the reference discards it, sample.py:21); ``workers`` is accepted and ignored.
There is no CPU path: without libtimed_b200.so and a B200 these functions raise.
"""
from __future__ import annotations

import ctypes as C
import json
import typing as t

import numpy as np

from . import _lib, seq_metrics
from .postprocess import standard_amino_acids

_state = {"seed": 42}


def set_seed(seed: int) -> None:
    """Seed of the counter-based generator used by all subsequent draws."""
    _state["seed"] = int(seed)


def _torch():
    import torch
    _lib.require_device()
    return torch


def _ptr(tensor):
    return C.c_void_p(tensor.data_ptr()) if tensor is not None else None


def _stream(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _letters_u8(rotamer_categories) -> np.ndarray:
    cats = list(rotamer_categories) if rotamer_categories is not None and len(rotamer_categories) else \
        list(standard_amino_acids.keys())
    if any(len(c) != 1 for c in cats):
        raise ValueError("categories must be one-letter codes (sample.py:50 maps rotamers to letters)")
    return np.frombuffer("".join(cats).encode("ascii"), dtype=np.uint8).copy()


# ----------------------------------------------------------------------------- temperature
def apply_temp_to_probs(probs: np.ndarray, t: float = 1.0) -> np.ndarray:
    """sampling_utils.py:139-161: ``p ** (1/t)`` renormalised per row, float64, on the device
    (row sums in numpy's pairwise order).  Agrees with numpy to the last few ulps (CUDA ``pow``
    is <= 2 ulp)."""
    torch = _torch()
    p = np.ascontiguousarray(np.array(probs, dtype=np.float64))
    if p.ndim != 2:
        raise ValueError("probs must be 2-D (n_residues, n_categories)")
    if p.size == 0:
        return p.copy()
    d = torch.from_numpy(p).cuda()
    out = torch.empty_like(d)
    _lib.check(_lib.load().timed_b200_apply_temperature(_ptr(d), p.shape[0], p.shape[1], float(t),
                                                        _ptr(out), _stream(torch)))
    return out.cpu().numpy()


# ----------------------------------------------------------------------------- draws
def sample_block(probs: np.ndarray, sample_n: int, rotamer_categories=None, *,
                 uniforms: t.Optional[np.ndarray] = None, seed: t.Optional[int] = None,
                 stream_id: int = 0, first_sample: int = 0, return_idx: bool = False,
                 temperature: t.Optional[float] = None, return_metrics: bool = False):
    """Draw ``sample_n`` sequences for one chain in one launch.

    probs: (n_res, C) float64 probabilities.  ``uniforms`` (sample_n, n_res) injects the random
    numbers (parity hook for ``np.random.rand``); otherwise Philox(seed, stream_id) is used and
    ``first_sample`` offsets the sample counter so that shards of one chain drawn on different
    GPUs concatenate to exactly what a single GPU would draw.
    Returns (letters uint8 (sample_n, n_res), idx int32 (sample_n, n_res) or None); with ``return_metrics`` a third
    item: (sample_n, 4) float64 [charge, pI, molecular weight, extinction], computed on the device from the letter
    block before it is copied back (device_post.seq_metrics_device)."""
    torch = _torch()
    lib = _lib.load()
    p = np.ascontiguousarray(np.array(probs, dtype=np.float64))
    if p.ndim != 2:
        raise ValueError("probs must be 2-D (n_residues, n_categories)")
    n_res, n_cls = p.shape
    letters = _letters_u8(rotamer_categories)
    if len(letters) != n_cls:
        raise ValueError(f"{n_cls} probability columns but {len(letters)} categories")
    if n_res == 0 or sample_n == 0:
        empty = (np.zeros((sample_n, n_res), np.uint8), (np.zeros((sample_n, n_res), np.int32) if return_idx else None))
        return empty + (np.zeros((sample_n, 4)),) if return_metrics else empty
    st = _stream(torch)
    d_p = torch.from_numpy(p).cuda()
    if temperature is not None and temperature != 1:
        _lib.check(lib.timed_b200_apply_temperature(_ptr(d_p), n_res, n_cls, float(temperature), _ptr(d_p), st))
    d_cdf = torch.empty_like(d_p)
    _lib.check(lib.timed_b200_cumsum_rows(_ptr(d_p), n_res, n_cls, _ptr(d_cdf), st))
    d_u = None
    if uniforms is not None:
        u = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64))
        if u.shape != (sample_n, n_res):
            raise ValueError(f"uniforms must have shape {(sample_n, n_res)}, got {u.shape}")
        d_u = torch.from_numpy(u).cuda()
    d_letters = torch.from_numpy(letters).cuda()
    n_cells = sample_n * n_res
    d_seq = torch.empty((n_cells + 3) // 4 * 4, dtype=torch.uint8, device="cuda")
    d_idx = torch.empty((n_cells + 3) // 4 * 4, dtype=torch.int32, device="cuda") if return_idx else None
    _lib.check(lib.timed_b200_sample(_ptr(d_cdf), n_res, n_cls, sample_n, int(first_sample),
                                     int(_state["seed"] if seed is None else seed) & (2 ** 64 - 1),
                                     int(stream_id) & (2 ** 64 - 1), _ptr(d_u), _ptr(d_letters),
                                     _ptr(d_seq), _ptr(d_idx), st))
    metrics = None
    if return_metrics:
        from . import device_post
        metrics = device_post.seq_metrics_device(d_seq, sample_n, n_res)
    seqs = d_seq[:n_cells].cpu().numpy().reshape(sample_n, n_res)
    idx = d_idx[:n_cells].cpu().numpy().reshape(sample_n, n_res) if return_idx else None
    return (seqs, idx, metrics) if return_metrics else (seqs, idx)


def sample_chains(prob_list: t.Sequence[np.ndarray], sample_n: int, rotamer_categories=None, *,
                  seed: t.Optional[int] = None, stream_id0: int = 0, first_sample: int = 0,
                  temperature: t.Optional[float] = None, return_metrics: bool = False):
    """Draw ``sample_n`` sequences for EVERY chain in ``prob_list`` with one cumsum launch and one sampling launch
    (``timed_b200_sample_chains``).  Chain ``i`` is keyed ``(seed, stream_id0 + i)``: the letters are byte-identical to
    ``sample_block(prob_list[i], ..., stream_id=stream_id0 + i)``.
    Returns a list of (sample_n, n_res_i) uint8 arrays (and, with ``return_metrics``, a list of (sample_n, 4) arrays)."""
    torch = _torch()
    lib = _lib.load()
    mats = [np.ascontiguousarray(np.array(p, dtype=np.float64)) for p in prob_list]
    if not mats:
        return ([], []) if return_metrics else []
    n_cls = mats[0].shape[1]
    if any(m.ndim != 2 or m.shape[1] != n_cls for m in mats):
        raise ValueError("every chain needs a 2-D (n_residues, n_categories) matrix with the same category count")
    letters = _letters_u8(rotamer_categories)
    if len(letters) != n_cls:
        raise ValueError(f"{n_cls} probability columns but {len(letters)} categories")
    lengths = np.array([m.shape[0] for m in mats], dtype=np.int64)
    row_off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    blocks = (lengths * int(sample_n) + 3) // 4 * 4                      # per-chain byte blocks, 4-byte aligned
    seq_off = np.concatenate([[0], np.cumsum(blocks)]).astype(np.int64)
    total_rows, total_bytes = int(row_off[-1]), int(seq_off[-1])
    if total_rows == 0 or sample_n == 0:
        seqs = [np.zeros((sample_n, int(n)), np.uint8) for n in lengths]
        return (seqs, [np.zeros((sample_n, 4)) for _ in lengths]) if return_metrics else seqs
    st = _stream(torch)
    d_p = torch.from_numpy(np.concatenate(mats, axis=0)).cuda()
    if temperature is not None and temperature != 1:
        _lib.check(lib.timed_b200_apply_temperature(_ptr(d_p), total_rows, n_cls, float(temperature), _ptr(d_p), st))
    d_cdf = torch.empty_like(d_p)
    _lib.check(lib.timed_b200_cumsum_rows(_ptr(d_p), total_rows, n_cls, _ptr(d_cdf), st))
    d_row = torch.from_numpy(row_off).cuda()
    d_off = torch.from_numpy(seq_off).cuda()
    d_letters = torch.from_numpy(letters).cuda()
    d_seq = torch.empty(total_bytes, dtype=torch.uint8, device="cuda")
    _lib.check(lib.timed_b200_sample_chains(_ptr(d_cdf), _ptr(d_row), _ptr(d_off), len(mats), total_bytes, n_cls,
                                            int(sample_n), int(first_sample),
                                            int(_state["seed"] if seed is None else seed) & (2 ** 64 - 1),
                                            int(stream_id0) & (2 ** 64 - 1), _ptr(d_letters), _ptr(d_seq), st))
    metrics = None
    if return_metrics:
        from . import device_post
        metrics = [device_post.seq_metrics_device(d_seq[int(seq_off[i]):], int(sample_n), int(n)) if n else
                   np.zeros((sample_n, 4)) for i, n in enumerate(lengths)]
    host, pin = _pinned_host(torch, total_bytes)      # pinned: the letters are the bulk of the traffic (1 B per residue)
    (pin[:total_bytes] if pin is not None else torch.from_numpy(host[:total_bytes])).copy_(d_seq, non_blocking=True)
    del pin
    torch.cuda.current_stream().synchronize()
    seqs = [host[int(seq_off[i]):int(seq_off[i]) + int(sample_n) * int(n)].reshape(int(sample_n), int(n))
            for i, n in enumerate(lengths)]
    return (seqs, metrics) if return_metrics else seqs


_pin_pool: list = []


def _pinned_host(torch, n_bytes: int):
    """A page-locked uint8 host buffer of at least ``n_bytes`` from a small pool.  The arrays handed back to the caller
    are numpy views of it; a buffer is reused only once no such view is alive any more (its reference count is back to the
    pool's own), so results of earlier calls are never overwritten.  Page-locking 100+ MB per call would cost more than the
    copy it speeds up."""
    import sys
    for entry in _pin_pool:
        if entry["host"].size >= n_bytes and sys.getrefcount(entry["host"]) <= entry["base_refs"]:
            return entry["host"], entry["tensor"]
    if len(_pin_pool) >= 4:                               # all leased: drop the oldest lease-free entry, or fall back
        for i, entry in enumerate(_pin_pool):
            if sys.getrefcount(entry["host"]) <= entry["base_refs"]:
                del _pin_pool[i]
                break
        else:
            return np.empty(n_bytes, dtype=np.uint8), None      # pageable (slower copy, still correct)
    t = torch.empty(max(n_bytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
    host = t.numpy()
    entry = {"tensor": t, "host": host}
    _pin_pool.append(entry)
    entry["base_refs"] = sys.getrefcount(host) - 1        # minus this frame's local name
    return host, t


_draw_counter = {"n": 0}
# Philox stream reserved for random_choice_prob_index: its draws never coincide with a chain's (chains are keyed by
# their index in sample_with_multiprocessing, or by a hash of their name in sample_from_sequences)
_RCPI_STREAM = (1 << 63) + 0x52435049


def _chain_stream_id(pdb: str) -> int:
    """Stable 63-bit stream id of a chain key (independent of PYTHONHASHSEED): a caller that loops
    ``sample_from_sequences`` over chains -- what the reference's starmap does -- gets an independent uniform block per
    chain instead of the same one for every chain."""
    import hashlib
    return int.from_bytes(hashlib.blake2b(str(pdb).encode(), digest_size=8).digest(), "little") >> 1


def random_choice_prob_index(probs: np.ndarray, axis: int = 1, return_seq: bool = True,
                             rotamer_categories: t.Optional[t.Sequence[str]] = None, *,
                             uniforms: t.Optional[np.ndarray] = None) -> np.ndarray:
    """sampling_utils.py:53-90: one categorical draw per row (``axis=1``) of ``probs``; returns
    the letters (``return_seq``) or the indices.  Successive calls advance the sample counter,
    as successive ``np.random.rand`` calls advance the reference's generator."""
    p = np.asarray(probs, dtype=np.float64)
    if axis == 0:
        p = p.T
    elif axis != 1:
        raise ValueError("axis must be 0 or 1")
    u = None if uniforms is None else np.asarray(uniforms, dtype=np.float64).reshape(1, -1)
    cats = rotamer_categories if (return_seq and rotamer_categories) else None
    n = _draw_counter["n"]
    _draw_counter["n"] += 1
    if not return_seq and p.shape[1] != 20:
        cats = ["A"] * p.shape[1]            # letters unused: indices requested
    seqs, idx = sample_block(p, 1, cats, uniforms=u, first_sample=n, return_idx=not return_seq,
                             stream_id=_RCPI_STREAM)
    if return_seq:
        return np.array(list(seqs[0].tobytes().decode("ascii")))
    return idx[0].astype(np.int64)


def _rows_to_tuples(seqs_u8: np.ndarray, metrics: t.Optional[np.ndarray] = None) -> list:
    """letters (+ the device-computed (n, 4) metric block) -> the reference's per-sample tuples."""
    strings = [row.tobytes().decode("ascii") for row in seqs_u8]
    if metrics is None or not strings:
        return [(s,) for s in strings]
    return [(s, float(m[0]), float(m[1]), float(m[2]), float(m[3])) for s, m in zip(strings, metrics)]


def sample_from_sequences(pdb: str, sample_n: int, pdb_to_probability: dict,
                          rotamer_categories: t.Optional[t.Sequence[str]], *,
                          stream_id: t.Optional[int] = None) -> dict:
    """sampling_utils.py:93-136: ``{pdb: [(sequence, charge, pI, mw, ext280)] * sample_n}``.  Without an explicit
    ``stream_id`` the chain draws from the Philox stream keyed by a stable hash of ``pdb``."""
    probs = np.array(pdb_to_probability[pdb], dtype=np.float64)
    if stream_id is None:
        stream_id = _chain_stream_id(pdb)
    seqs, _, metrics = sample_block(probs, int(sample_n), rotamer_categories, stream_id=stream_id, return_metrics=True)
    return {pdb: _rows_to_tuples(seqs, metrics)}


def sample_with_multiprocessing(workers, pdb_codes, sample_n, pdb_to_probability, flat_categories) -> dict:
    """sampling_utils.py:164-197.  ``workers`` is accepted for compatibility and ignored: the fan-out over chains is ONE
    sampling launch over the concatenated chains (``sample_chains``), each chain keyed by its index so the draws do not
    depend on how chains are distributed (and equal ``sample_from_sequences(..., stream_id=index)``)."""
    pdb_codes = list(pdb_codes)
    probs = [pdb_to_probability[p] for p in pdb_codes]
    from .predict import _dist_context
    rank, world, local_rank = _dist_context()
    if world <= 1:
        seqs, metrics = sample_chains(probs, int(sample_n), flat_categories, return_metrics=True)
        return {pdb: _rows_to_tuples(s, m) for pdb, s, m in zip(pdb_codes, seqs, metrics)}
    # torchrun: every rank draws a contiguous block of the sample index for all chains (the counter-based generator makes
    # the union identical to one device's draw), one all-gather per output reassembles the blocks on every rank
    import torch
    import torch.distributed as dist
    from .dist import gather_rows, shard_samples
    if dist.get_backend() == "nccl":
        torch.cuda.set_device(local_rank)
    first, count = shard_samples(int(sample_n), rank, world)
    lengths = [len(p) for p in probs]
    seqs, metrics = sample_chains(probs, count, flat_categories, first_sample=first, return_metrics=True)
    local_s = np.concatenate([s.reshape(count, n) for s, n in zip(seqs, lengths)], axis=1) if lengths else np.zeros((count, 0), np.uint8)
    local_m = np.concatenate([np.asarray(m, dtype=np.float64).reshape(count, 4) for m in metrics], axis=1) if lengths else np.zeros((count, 0))

    def gather(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        if dist.get_backend() == "nccl":
            t = t.cuda(local_rank)
        return gather_rows(t, int(sample_n)).cpu().numpy()

    all_s, all_m = gather(local_s), gather(local_m)
    out, col = {}, 0
    for i, (pdb, n) in enumerate(zip(pdb_codes, lengths)):
        out[pdb] = _rows_to_tuples(all_s[:, col:col + n], all_m[:, 4 * i:4 * i + 4])
        col += n
    return out


def save_as(pdb_to_sampled: dict, filename: str, mode: str) -> t.List[str]:
    """sampling_utils.py:12-50: ``.json`` / ``.fasta`` (mode) and always ``_metrics.csv``, written
    relative to the CWD like the reference."""
    output_paths = []
    print(f"Saving sampled sequences in mode {mode}")
    if mode != "fasta":
        path = f"{filename}.json"
        output_paths.append(path)
        with open(path, "w") as f:
            json.dump(pdb_to_sampled, f)
    if mode != "json":
        path = f"{filename}.fasta"
        output_paths.append(path)
        with open(path, "w") as f:
            for pdb, rows in pdb_to_sampled.items():
                f.writelines(f">{pdb}_{i}\n{row[0]}\n" for i, row in enumerate(rows))
    print("Saving Metrics")
    path = f"{filename}_metrics.csv"
    output_paths.append(path)
    with open(path, "w") as f:
        f.write("pdb,sequence,charge,isoelectric_point,molecular_weight,molar_extinction\n")
        for pdb, rows in pdb_to_sampled.items():
            f.writelines(f"{pdb},{r[0]},{r[1]},{r[2]},{r[3]},{r[4]}\n" for r in rows)
    return output_paths
