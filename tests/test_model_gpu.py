"""GPU parity of the whole inference graph against the CPU oracle (numpy fp64 and torch fp32).

Tolerance from BASELINE.json north_star: per-residue softmax probabilities within 1e-4 of the
fp32 CPU path; fp16-rounded argmax identical outside the reported near-tie set."""
import numpy as np
import pytest

from oracle import keras_oracle as ko
from timed_design_b200 import standins

PROB_TOL = 1e-4


def _check(model_fn, frames_fn, n, tol=PROB_TOL, use_numpy64=True):
    from timed_design_b200.model import Model
    cfg, w = model_fn()
    X = frames_fn(n)
    m = Model(cfg, w)
    p = m.predict(X)
    assert p.shape == (n, m.n_classes) and p.dtype == np.float32
    ref32 = ko.forward_torch(cfg, w, X)
    d32 = float(np.abs(p - ref32).max())
    assert np.isfinite(p).all()
    assert d32 <= tol, f"max|dp| vs torch-fp32 oracle = {d32:.3e}"
    if use_numpy64:
        ref64 = ko.forward_numpy(cfg, w, X)
        assert float(np.abs(p - ref64).max()) <= tol
    safe = ~ko.near_tie_rows(ref32)
    assert (ko.fp16_argmax(p)[safe] == ko.fp16_argmax(ref32)[safe]).all()
    np.testing.assert_allclose(p.sum(1), 1.0, atol=1e-5)
    return p, ref32


@pytest.mark.gpu
def test_tiny_timed_parity():
    _check(lambda: standins.tiny_standin(), lambda n: standins.synthetic_frames(n, side=9), 37)


@pytest.mark.gpu
def test_tiny_timed_parity_bool_and_f64_inputs():
    from timed_design_b200.model import Model
    cfg, w = standins.tiny_standin()
    m = Model(cfg, w)
    X = standins.synthetic_frames(9, side=9)
    Xb = X > 0.2
    for arr in (X.astype(np.float64), Xb, Xb.astype(np.uint8)):
        p = m.predict(arr)
        ref = ko.forward_torch(cfg, w, arr)
        assert np.abs(p - ref).max() <= PROB_TOL


@pytest.mark.gpu
def test_timed_standin_parity_20():
    p, ref = _check(lambda: standins.timed_standin(20), lambda n: standins.synthetic_frames(n), 6,
                    use_numpy64=False)
    assert len(set(ko.fp16_argmax(ref))) > 1      # the stand-in is not a constant predictor


@pytest.mark.gpu
def test_timed_standin_parity_338():
    _check(lambda: standins.timed_standin(338, seed=8), lambda n: standins.synthetic_frames(n), 3,
           use_numpy64=False)


@pytest.mark.gpu
def test_prodconn_standin_parity():
    _check(lambda: standins.prodconn_standin(side=13, calib_frames=3),
           lambda n: standins.synthetic_frames(n, side=13), 5)


@pytest.mark.gpu
def test_densecpd_small_parity():
    _check(lambda: standins.densecpd_standin(side=12, n_layers=2, calib_frames=3),
           lambda n: standins.synthetic_frames(n, side=12), 5)


@pytest.mark.gpu
def test_precise_mode_is_tighter_and_consistent():
    """Model(precise=True): 2-CTA cluster path for the full-width layers.  Same answers within tolerance,
    measurably closer to the oracle on a batch large enough to engage the cluster tiles."""
    from timed_design_b200.model import Model
    cfg, w = standins.timed_standin(20)
    uniq = standins.synthetic_frames(48, seed=5)
    X = np.tile(uniq, (8, 1, 1, 1, 1))                      # 384 frames: enough rows for cluster tiles
    ref = ko.forward_torch(cfg, w, uniq)
    fast = Model(cfg, w).predict(X, batch_size=4096)
    prec = Model(cfg, w, precise=True).predict(X, batch_size=4096)
    e_fast = np.abs(fast[:48] - ref).max()
    e_prec = np.abs(prec[:48] - ref).max()
    assert e_fast <= PROB_TOL and e_prec <= PROB_TOL
    assert e_prec < e_fast
    np.testing.assert_array_equal(prec[:48], prec[48:96])    # batch-position independent
    safe = ~ko.near_tie_rows(ref)
    assert (ko.fp16_argmax(prec[:48])[safe] == ko.fp16_argmax(ref)[safe]).all()
