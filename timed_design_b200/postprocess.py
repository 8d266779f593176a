"""Host side of the prediction path: argmax -> sequences, per-chain grouping, NMR consensus,
the 338 <-> 20 rotamer codec and every file the reference's ``predict.py`` leaves behind.

Same function names, arguments and return conventions as
``/root/reference/design_utils/utils.py`` (lines cited per function) so that ``predict.py``,
``sample.py`` and ``ui.py`` written against the reference keep working; the bodies are
vectorised over the (N, C) probability matrix instead of looping per row in Python, because at
1 M frames the reference's loops would dwarf the GPU time.  Output files are byte-identical
(tests/golden/files.json).

Deliberate, documented deviations:
  * ``pdb_to_probability`` values are 2-D ndarrays (n_res, C) rather than lists of lists of numpy
    scalars -- ``np.array(v)``, ``len(v)``, iteration and ``np.savetxt`` behave the same.
"""
from __future__ import annotations

import io
import typing as t
from itertools import product
from pathlib import Path

import numpy as np

# ampal.amino_acids.standard_amino_acids (one-letter -> three-letter), in ampal's order.  The
# order is pinned in-repo by the rotamer offsets quoted at design_utils/utils.py:425.
standard_amino_acids = {
    "A": "ALA", "C": "CYS", "D": "ASP", "E": "GLU", "F": "PHE", "G": "GLY", "H": "HIS", "I": "ILE",
    "K": "LYS", "L": "LEU", "M": "MET", "N": "ASN", "P": "PRO", "Q": "GLN", "R": "ARG", "S": "SER",
    "T": "THR", "V": "VAL", "W": "TRP", "Y": "TYR"}
# number of side-chain chi angles (len(ampal.amino_acids.side_chain_dihedrals[res]))
side_chain_chi_count = {
    "ARG": 4, "ASN": 2, "ASP": 2, "CYS": 1, "GLN": 3, "GLU": 3, "HIS": 2, "ILE": 2, "LEU": 2, "LYS": 4,
    "MET": 3, "PHE": 2, "PRO": 2, "SER": 1, "THR": 1, "TRP": 2, "TYR": 2, "VAL": 1}
_THREE_TO_ONE = {v: k for k, v in standard_amino_acids.items()}
LETTERS20 = np.array(list(standard_amino_acids.keys()))
SRB_HEADER = "ignore_uncommon False\ninclude_pdbs\n##########\n"


# ----------------------------------------------------------------------------- rotamer codec
def get_rotamer_codec(return_reduction_guide: bool = False):
    """utils.py:410-465.  338 classes = residues in standard order x 3**n_chi rotamers
    (``product([1,2,3], repeat=n_chi)`` order); labels ``RES_chi...`` (``ALA_0``, ``CYS_1`` ...).
    Returns ({class: one-hot(20)}, labels[, first class of each residue])."""
    labels: t.List[str] = []
    owner: t.List[int] = []
    guide: t.List[int] = []
    for i, res in enumerate(standard_amino_acids.values()):
        guide.append(len(labels))
        n_chi = side_chain_chi_count.get(res, 0)
        if n_chi:
            for combo in product("123", repeat=n_chi):
                labels.append(f"{res}_{''.join(combo)}")
                owner.append(i)
        else:
            labels.append(f"{res}_0")
            owner.append(i)
    eye = np.eye(20, dtype=int)
    rot_to_20res = {cls: eye[o].copy() for cls, o in enumerate(owner)}
    if return_reduction_guide:
        return rot_to_20res, labels, guide
    return rot_to_20res, labels


def rotamer_class_to_residue() -> np.ndarray:
    """(338,) residue index of every rotamer class (vector form of the codec)."""
    codec, _ = get_rotamer_codec()
    return np.array([int(np.argmax(codec[i])) for i in range(len(codec))])


def compress_rotamer_predictions_to_20(prediction_matrix: np.ndarray) -> np.ndarray:
    """utils.py:468-484: sum each residue's rotamer block, (n,338) -> (n,20)."""
    _, _, guide = get_rotamer_codec(return_reduction_guide=True)
    return np.add.reduceat(prediction_matrix, guide, axis=1)


def _letters_for(rotamers_categories) -> np.ndarray:
    """Class index -> one-letter code (utils.py:650-657)."""
    if rotamers_categories is not None and len(rotamers_categories) > 0:
        cats = list(rotamers_categories)
        if len(cats[0]) == 1:
            return np.array(cats)
        return np.array([_THREE_TO_ONE[c.split("_")[0]] for c in cats])
    return LETTERS20


# ----------------------------------------------------------------------------- sequences
def extract_sequence_from_pred_matrix(
    flat_dataset_map,
    prediction_matrix: np.ndarray,
    rotamers_categories: t.Optional[t.List[str]],
    old_datasetmap: bool = False,
    is_consensus: bool = False,
) -> t.Tuple[dict, dict, dict, t.Optional[dict], t.Optional[dict]]:
    """utils.py:616-723.  argmax (first index on ties, on whatever dtype the caller rounded the
    matrix to -- the reference hands in float16, predict.py:163) -> letters; rows grouped per
    chain key ``pdb+chain`` for the old 4-column map or by running ``count`` offsets for the
    2-column ``{model}.txt`` map; optional NMR consensus = running pairwise mean over states."""
    prediction_matrix = np.asarray(prediction_matrix)
    letters = _letters_for(rotamers_categories)
    flat_dataset_map = np.asarray(flat_dataset_map)
    if flat_dataset_map.ndim == 1:
        flat_dataset_map = flat_dataset_map[None, :]
    picked = letters[np.argmax(prediction_matrix, axis=1)] if len(prediction_matrix) else np.array([], "<U1")
    is_old = flat_dataset_map.shape[1] == 4           # decided by the data, as utils.py:662 does

    pdb_to_sequence: dict = {}
    pdb_to_probability: dict = {}
    pdb_to_real_sequence: dict = {}
    if is_old:
        keys = np.char.add(flat_dataset_map[:, 0], flat_dataset_map[:, 1])
        uniq, first, inverse = np.unique(keys, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")       # dict insertion order = first appearance
        real_letters = np.array([_THREE_TO_ONE[r] for r in flat_dataset_map[:, 3]])
        by_key = np.argsort(inverse, kind="stable")
        bounds = np.concatenate([[0], np.cumsum(np.bincount(inverse, minlength=len(uniq)))])
        for u in order:
            rows = by_key[bounds[u]:bounds[u + 1]]
            k = str(uniq[u])
            pdb_to_sequence[k] = "".join(picked[rows])
            pdb_to_real_sequence[k] = "".join(real_letters[rows])
            pdb_to_probability[k] = prediction_matrix[rows]
    else:
        start = 0
        for key, count in flat_dataset_map:
            k, n = str(key), int(count)
            rows = slice(start, start + n)
            seq = "".join(picked[rows])
            if k in pdb_to_sequence:
                pdb_to_sequence[k] += seq
                pdb_to_probability[k] = np.concatenate([pdb_to_probability[k], prediction_matrix[rows]])
            else:
                pdb_to_sequence[k] = seq
                pdb_to_real_sequence[k] = ""         # new-style maps carry no true sequence
                pdb_to_probability[k] = prediction_matrix[rows]
            start += n

    if not is_consensus:
        return pdb_to_sequence, pdb_to_probability, pdb_to_real_sequence, None, None

    pdb_to_consensus_prob: dict = {}
    last = ""
    for key in pdb_to_sequence:
        cur = key.split("_")[0]
        p = np.array(pdb_to_probability[key])
        if cur != last:
            pdb_to_consensus_prob[cur] = p
            last = cur
        else:                                         # (prev + cur) / 2, not a true mean: kept
            pdb_to_consensus_prob[cur] = (pdb_to_consensus_prob[cur] + p) / 2
    pdb_to_consensus = {k: "".join(letters[np.argmax(v, axis=1)]) for k, v in pdb_to_consensus_prob.items()}
    return pdb_to_sequence, pdb_to_probability, pdb_to_real_sequence, pdb_to_consensus, pdb_to_consensus_prob


# ----------------------------------------------------------------------------- dataset maps
def load_datasetmap(path_to_datasetmap: Path, is_old: bool = False) -> np.ndarray:
    """utils.py:190-227: comma-separated 4-column map, or the space-separated ``{model}.txt``
    (3 header lines); a single row is re-wrapped to 2-D."""
    path_to_datasetmap = Path(path_to_datasetmap)
    assert path_to_datasetmap.suffix == ".txt", \
        f"Expected Path {path_to_datasetmap} to be a .txt file but got {path_to_datasetmap.suffix}."
    if is_old:
        dataset_map = np.genfromtxt(path_to_datasetmap, delimiter=",", dtype=str)
    else:
        dataset_map = np.genfromtxt(path_to_datasetmap, delimiter=" ", dtype=str, skip_header=3)
    dataset_map = np.asarray(dataset_map)
    if dataset_map.ndim == 1:
        dataset_map = dataset_map[None, :]
    return dataset_map


def get_pdb_keys_to_filter(pdb_key_path: Path, file_extension: str = ".txt") -> t.List[str]:
    """utils.py:283-315: first four characters of every key in every list file under the dir."""
    files = list(Path(pdb_key_path).glob(f"**/*{file_extension}"))
    assert len(files) >= 1, "Expected at least 1 pdb key file."
    keys: t.List[str] = []
    for f in files:
        keys.extend(str(k)[:4] for k in np.atleast_1d(np.genfromtxt(f, dtype=str)))
    return keys


def convert_dataset_map_for_srb(flat_dataset_map, model_name: str, path_to_output: Path = Path.cwd()):
    """utils.py:533-566 -> ``{model}.txt``: frames per ``pdb`` (+chain when the code has four
    characters), a trailing ``_0`` state suffix stripped."""
    counts: dict = {}
    for pdb, chain, _, _ in np.asarray(flat_dataset_map):
        pdb = str(pdb)
        if "_0" in pdb:
            pdb = pdb.split("_0")[0]
        if len(pdb) == 4:
            pdb += str(chain)
        counts[pdb] = counts.get(pdb, 0) + 1
    with open(Path(path_to_output) / f"{model_name}.txt", "w") as d:
        d.write(SRB_HEADER)
        d.writelines(f"{pdb} {n}\n" for pdb, n in counts.items())


def save_consensus_probs(pdb_to_consensus_prob: dict, model_name: str, path_to_output: Path = Path.cwd()):
    """utils.py:569-592.  Quirk kept: the ``.csv`` goes to the CWD (bare filename, append)."""
    with open(Path(path_to_output) / f"{model_name}_consensus.txt", "w") as d, \
            open(f"{model_name}_consensus.csv", "a") as p:
        d.write(SRB_HEADER)
        for pdb, predictions in pdb_to_consensus_prob.items():
            d.write(f"{pdb} {len(predictions)}\n")
            np.savetxt(p, predictions, delimiter=",")


def save_dict_to_fasta(pdb_to_sequence: dict, model_name: str, path_to_output: Path = Path.cwd()):
    """utils.py:595-613."""
    with open(Path(path_to_output) / f"{model_name}.fasta", "w") as f:
        f.writelines(f">{pdb}\n{seq}\n" for pdb, seq in pdb_to_sequence.items())


_FP16_TEXT: dict = {}


def _fp16_text_tables():
    """'%.18e' text of every float16 bit pattern, once (65 536 strings, ~60 ms): the CSV the reference writes holds
    float16-cast probabilities (utils.py:768-771), so formatting is a table lookup instead of a Python loop per number."""
    if not _FP16_TEXT:
        vals = np.arange(65536, dtype=np.uint16).view(np.float16).astype(np.float64)
        with np.errstate(all="ignore"):
            text = ["%.18e" % v for v in vals]
        ok = np.array([len(t) == 24 for t in text])              # non-negative finite values: fixed 24 characters
        _FP16_TEXT["ok"] = ok
        _FP16_TEXT["comma"] = np.array([(t + ",").encode() if k else b"" for t, k in zip(text, ok)], dtype="S25")
        _FP16_TEXT["nl"] = np.array([(t + "\n").encode() if k else b"" for t, k in zip(text, ok)], dtype="S25")
    return _FP16_TEXT


def savetxt_fp16(f, arr16: np.ndarray) -> None:
    """Byte-identical to ``np.savetxt(f, arr16, delimiter=",")`` for a 2-D float16 array (numpy's default '%.18e'), written
    from a 65 536-entry text table; rows holding negative or non-finite values fall back to ``np.savetxt``."""
    a = np.ascontiguousarray(arr16, dtype=np.float16)
    if a.ndim != 2 or a.size == 0:
        np.savetxt(f, a, delimiter=",")
        return
    t = _fp16_text_tables()
    idx = a.view(np.uint16)
    if not t["ok"][idx].all():
        np.savetxt(f, a, delimiter=",")
        return
    out = t["comma"][idx]
    out[:, -1] = t["nl"][idx[:, -1]]
    data = out.tobytes()
    f.write(data.decode("ascii") if isinstance(f, io.TextIOBase) else data)


def _warn_fallback(native: str, fallback: str, err: Exception) -> None:
    """The host-side text helpers have a numpy equivalent (same bytes / values, an order of magnitude slower); falling
    back silently would hide a broken or missing libtimed_b200 symbol."""
    import warnings
    warnings.warn(f"{native} unavailable ({type(err).__name__}: {err}); using {fallback} (identical output, much slower)",
                  RuntimeWarning, stacklevel=3)


def savetxt_e18(f, arr) -> None:
    """Byte-identical to ``np.savetxt(f, arr, delimiter=",")`` for a 2-D float32/float64 array (the raw 338-wide rotamer
    dump, predict.py:145-146), formatted by libtimed_b200's host-side writer on all cores; any other input goes through
    ``np.savetxt``."""
    import ctypes as C
    import os
    a = np.asarray(arr)
    if a.ndim != 2 or a.size == 0 or a.dtype not in (np.float32, np.float64):
        np.savetxt(f, a, delimiter=",")
        return
    try:
        from . import _lib
        lib = _lib.load()
    except Exception as e:                                 # library missing: correct but ~13x slower -- say so
        _warn_fallback("timed_b200_format_csv_e18", "np.savetxt", e)
        np.savetxt(f, a, delimiter=",")
        return
    a = np.ascontiguousarray(a)
    buf = np.empty(a.size * 26, dtype=np.uint8)
    n = C.c_int64()
    _lib.check(lib.timed_b200_format_csv_e18(C.c_void_p(a.ctypes.data), _lib.np_dtype_code(a), a.shape[0], a.shape[1],
                                             C.c_void_p(buf.ctypes.data), buf.size, C.byref(n),
                                             min(32, os.cpu_count() or 1)))
    data = buf[:n.value].tobytes()
    f.write(data.decode("ascii") if isinstance(f, io.TextIOBase) else data)


def load_matrix_csv(path) -> np.ndarray:
    """``np.genfromtxt(path, delimiter=",", dtype=np.float64)`` for the prediction matrices this package writes, parsed
    by libtimed_b200's host-side reader (strtod on all cores); anything it does not accept goes through ``genfromtxt``."""
    import ctypes as C
    import mmap
    import os
    try:
        from . import _lib
        lib = _lib.load()
        with open(path, "rb") as fh:
            if os.fstat(fh.fileno()).st_size == 0:
                raise ValueError("empty file")
            mm = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
            view = np.frombuffer(mm, dtype=np.uint8)
            rows, cols = C.c_int64(), C.c_int64()
            _lib.check(lib.timed_b200_parse_csv(C.c_void_p(view.ctypes.data), view.size, None, 0, C.byref(rows),
                                                C.byref(cols), 1))
            out = np.empty((rows.value, cols.value), dtype=np.float64)
            _lib.check(lib.timed_b200_parse_csv(C.c_void_p(view.ctypes.data), view.size, C.c_void_p(out.ctypes.data),
                                                out.size, C.byref(rows), C.byref(cols), min(32, os.cpu_count() or 1)))
            del view
            mm.close()
        return out[0] if out.shape[0] == 1 else out              # genfromtxt returns 1-D for a single row
    except Exception as e:                                 # not a rectangular numeric matrix, or the library is missing
        _warn_fallback("timed_b200_parse_csv", "np.genfromtxt", e)
        return np.genfromtxt(path, delimiter=",", dtype=np.float64)


def savetxt_onehot(f, arr) -> None:
    """Byte-identical to ``np.savetxt(f, arr, delimiter=",", fmt="%i")`` when every entry is 0 or 1 (the label one-hots);
    anything else goes through ``np.savetxt``."""
    a = np.asarray(arr)
    if a.ndim != 2 or a.size == 0 or not np.isin(a, (0, 1)).all():
        np.savetxt(f, a, delimiter=",", fmt="%i")
        return
    out = np.where(a.astype(bool), b"1,", b"0,").astype("S2")
    out[:, -1] = np.where(a[:, -1].astype(bool), b"1\n", b"0\n")
    data = out.tobytes()
    f.write(data.decode("ascii") if isinstance(f, io.TextIOBase) else data)


def save_outputs_to_file(y_true, y_pred, flat_dataset_map, model: int, model_name: str,
                         path_to_output: Path = Path.cwd()):
    """utils.py:726-771.  Appends labels (``%i``) for model 0, writes ``datasetmap.txt`` once,
    appends the predictions cast to float16 with numpy's default ``%.18e``."""
    path_to_output = Path(path_to_output)
    if model == 0:
        with open(path_to_output / "encoded_labels.csv", "a") as f:
            savetxt_onehot(f, np.asarray(y_true))
    map_path = path_to_output / "datasetmap.txt"
    if not map_path.exists():
        with open(map_path, "a") as f:
            np.savetxt(f, np.asarray(flat_dataset_map), delimiter=",", fmt="%s")
    with open(path_to_output / f"{model_name}.csv", "a") as f:
        savetxt_fp16(f, np.array(y_pred[model], dtype=np.float16))
