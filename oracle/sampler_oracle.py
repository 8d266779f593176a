"""ORACLE (test infrastructure, never shipped on the product path).

numpy restatement of the reference's Monte-Carlo sampler arithmetic:
  * apply_temp_to_probs       /root/reference/design_utils/sampling_utils.py:139-161
  * random_choice_prob_index  /root/reference/design_utils/sampling_utils.py:53-90
  * sample_from_sequences     /root/reference/design_utils/sampling_utils.py:93-136 (metrics call
    excluded: ampal is absent)
plus a pure-Python Philox4x32-10 so the GPU's counter-based uniforms can be checked bit for bit.

Pinned: checked against tests/golden/sampler.npz and temperature.npz, which were produced by
executing the reference's own functions (tests/golden/make_golden.py), and against the golden
probability row + properties of /root/reference/tests/test_sampling_utils.py:5-62.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np

LETTERS20 = list("ACDEFGHIKLMNPQRSTVWY")     # ampal.standard_amino_acids key order (utils.py:425)


def apply_temp_to_probs(probs: np.ndarray, t: float = 1.0) -> np.ndarray:
    """sampling_utils.py:159-161 -- p ** (1/t), np.sum(axis=1), divide (float64)."""
    probs = np.array(probs) ** (1 / t)
    p_sum = np.sum(probs, axis=1)
    return probs / p_sum[:, None]


def choice_index(probs: np.ndarray, r: np.ndarray) -> np.ndarray:
    """sampling_utils.py:81-82 with the uniforms injected: (cumsum(axis=1) > r[:,None]).argmax(1).
    argmax of an all-False row is 0: rows whose cumsum never exceeds r yield class 0."""
    return (probs.cumsum(axis=1) > np.asarray(r)[:, None]).argmax(axis=1)


def sample_sequences(probs: np.ndarray, uniforms: np.ndarray, categories=None) -> list:
    """sampling_utils.py:123-130 for one chain: one sequence string per row of ``uniforms``."""
    res = np.array(list(categories) if categories is not None else LETTERS20)
    return ["".join(res[choice_index(probs, r)]) for r in uniforms]


def sample_loop_numpy(probs: np.ndarray, sample_n: int, categories=None) -> list:
    """The reference's inner loop verbatim (global legacy RNG, cumsum recomputed per sample);
    used as the timed CPU baseline of the sampler."""
    res = np.array(list(categories) if categories is not None else LETTERS20)
    out = []
    for _ in range(sample_n):
        r = np.expand_dims(np.random.rand(probs.shape[0]), axis=1)
        idxs = (probs.cumsum(axis=1) > r).argmax(axis=1)
        out.append("".join(res[idxs]))
    return out


# ----------------------------------------------------------------------------- Philox4x32-10
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    c = [int(x) & _MASK for x in counter]
    k0, k1 = int(key[0]) & _MASK, int(key[1]) & _MASK
    for _ in range(10):
        p0 = _M0 * c[0]
        p1 = _M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c[3] ^ k1) & _MASK, p0 & _MASK]
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c


def philox_uniform(sample: int, res: int, seed: int, stream_id: int) -> float:
    """Same construction as kernels.cuh::philox_uniform (53-bit double in [0,1))."""
    mix = (stream_id * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    key = ((seed & _MASK) ^ (mix >> 32), ((seed >> 32) & _MASK) ^ (mix & _MASK))
    pair = res >> 1                       # one Philox block serves two residues: words 0-1 the even one, 2-3 the odd one
    c = philox4x32_10([sample & _MASK, (sample >> 32) & _MASK, pair & _MASK, (pair >> 32) & _MASK], key)
    a, b = (c[2] >> 5, c[3] >> 6) if res & 1 else (c[0] >> 5, c[1] >> 6)
    return (a * 67108864.0 + b) / 9007199254740992.0


def philox_uniforms(n_res: int, n_samples: int, first_sample: int, seed: int, stream_id: int) -> np.ndarray:
    out = np.empty((n_samples, n_res))
    for s in range(n_samples):
        for r in range(n_res):
            out[s, r] = philox_uniform(first_sample + s, r, seed, stream_id)
    return out
