"""Per-sequence metrics written next to every sampled sequence (charge, isoelectric point,
molecular weight, molar extinction at 280 nm).

The reference computes them inside its Monte-Carlo inner loop through four ampal functions
(``calculate_seq_metrics``, /root/reference/design_utils/analyse_utils.py:351-371, called at
sampling_utils.py:132).  ampal==1.5.1 is not vendored and not installable here, so the residue
tables below are a RECOLLECTION of ``ampal/amino_acids.py`` and the formulas a restatement of
``ampal/analyse_protein.py`` -- **unverified against ampal** (SURVEY.md 8(f)-3 / App. G); parity
is claimed for the file *formats* only.  All four metrics depend on the residue composition
alone, so they are evaluated for a whole batch of sampled sequences at once from a
(n_sequences, 20) histogram instead of per sequence in a Python loop.
"""
from __future__ import annotations

import numpy as np

LETTERS = "ACDEFGHIKLMNPQRSTVWY"
WATER_MASS = 18.0153
_MWT = dict(A=71.0779, C=103.1429, D=115.0874, E=129.114, F=147.1739, G=57.0513, H=137.1393,
            I=113.1576, K=128.1723, L=113.1576, M=131.1961, N=114.1026, P=97.1152, Q=128.1292,
            R=156.1857, S=87.0773, T=101.1039, V=99.1311, W=186.2099, Y=163.1733)
_EXT280 = dict(C=125.0, W=5690.0, Y=1280.0)
_CHARGE = dict(C=-1, D=-1, E=-1, H=1, K=1, R=1, Y=-1)
_PKA = dict(C=8.3, D=3.65, E=4.25, H=6.1, K=10.53, R=12.48, Y=10.1)
_NTERM = (1, 8.0)
_CTERM = (-1, 3.1)

MWT = np.array([_MWT[a] for a in LETTERS])
EXT280 = np.array([_EXT280.get(a, 0.0) for a in LETTERS])
CHARGE = np.array([_CHARGE.get(a, 0) for a in LETTERS], dtype=np.float64)
PKA = np.array([_PKA.get(a, 0.0) for a in LETTERS])
_LUT = np.full(256, -1, dtype=np.int64)
for _i, _a in enumerate(LETTERS):
    _LUT[ord(_a)] = _i


def composition(seqs_u8: np.ndarray) -> np.ndarray:
    """(n_seq, n_res) ASCII codes -> (n_seq, 20) residue counts."""
    idx = _LUT[seqs_u8]
    if (idx < 0).any():
        raise ValueError("sequence contains a non-standard residue letter")
    n = seqs_u8.shape[0]
    flat = idx + (np.arange(n)[:, None] * 20)
    return np.bincount(flat.ravel(), minlength=n * 20).reshape(n, 20)


def _partial(charge_sign: np.ndarray, pka: np.ndarray, ph: np.ndarray) -> np.ndarray:
    """ampal partial_charge: 10^d / (1 + 10^d), d = pH - pKa, negated for positive groups."""
    diff = ph[..., None] - pka
    diff = np.where(charge_sign > 0, -diff, diff)
    e = 10.0 ** diff
    return e / (1.0 + e)


def charge_at(counts: np.ndarray, ph) -> np.ndarray:
    """Net charge of every sequence at every pH in ``ph`` -> (n_seq, n_ph)."""
    ph = np.atleast_1d(np.asarray(ph, dtype=np.float64))
    ion = CHARGE != 0
    per_res = _partial(CHARGE[ion], PKA[ion], ph) * CHARGE[ion]           # (n_ph, n_ion)
    total = counts[:, ion].astype(np.float64) @ per_res.T                   # (n_seq, n_ph)
    for sign, pka in (_NTERM, _CTERM):
        total += (_partial(np.array([sign]), np.array([pka]), ph) * sign)[:, 0]
    return total


_warned = {"done": False}


def warn_unverified() -> None:
    """One RuntimeWarning per process the first time metric values are produced: the residue tables are unverified
    against ampal 1.5.1 (absent offline), so charge / pI / MW / ext280 must not be taken as drop-in values."""
    if not _warned["done"]:
        _warned["done"] = True
        import warnings
        warnings.warn("sequence metrics (charge, isoelectric point, molecular weight, extinction) use residue tables that "
                      "are UNVERIFIED against ampal 1.5.1 (not installable offline); file formats match the reference, "
                      "metric values may differ", RuntimeWarning, stacklevel=3)


def metrics_from_composition(counts: np.ndarray):
    """(charge at pH 7.4, isoelectric point on the 1.0..12.9 step-0.1 grid, molecular weight,
    molar extinction at 280 nm), one row per sequence."""
    counts = np.asarray(counts)
    charge = charge_at(counts, 7.4)[:, 0]
    grid = np.arange(1, 13, 0.1)
    pi = grid[np.argmin(np.abs(charge_at(counts, grid)), axis=1)]          # first minimum, like min()
    mw = counts @ MWT + WATER_MASS
    ext = counts @ EXT280
    return charge, pi, mw, ext


def calculate_seq_metrics(seq: str):
    """Drop-in signature of analyse_utils.calculate_seq_metrics for one sequence."""
    warn_unverified()
    c, p, m, e = metrics_from_composition(composition(np.frombuffer(seq.encode(), dtype=np.uint8)[None, :]))
    return float(c[0]), float(p[0]), float(m[0]), float(e[0])


def device_tables() -> np.ndarray:
    """Packed float64 table for ``timed_b200_seq_metrics`` (include/timed_b200.h):
    [mw(20) | ext280(20) | q(pH 7.4)(20) | term(pH 7.4) | water | n_grid | grid | term(grid) | q(grid) (n_grid x 20)]
    where q(pH)[i] = charge_i * partial_charge_i(pH) and term(pH) = N-terminus + C-terminus contribution."""
    grid = np.arange(1, 13, 0.1)

    def per_res(ph):
        ph = np.atleast_1d(np.asarray(ph, dtype=np.float64))
        q = np.zeros((len(ph), 20))
        ion = CHARGE != 0
        q[:, ion] = _partial(CHARGE[ion], PKA[ion], ph) * CHARGE[ion]
        term = np.zeros(len(ph))
        for sign, pka in (_NTERM, _CTERM):
            term += (_partial(np.array([sign]), np.array([pka]), ph) * sign)[:, 0]
        return q, term

    q74, t74 = per_res(7.4)
    qg, tg = per_res(grid)
    return np.concatenate([MWT, EXT280, q74[0], t74, [WATER_MASS], [float(len(grid))], grid, tg, qg.ravel()]).astype(np.float64)
