"""CPU: the table-driven CSV writers are byte-identical to the numpy calls the reference makes
(np.savetxt(..., delimiter=",") on the float16-cast probabilities, utils.py:768-771; fmt="%i" on the labels, :757-760)."""
import io

import numpy as np

from timed_design_b200 import postprocess as pp


def _np_text(arr, **kw):
    f = io.StringIO()
    np.savetxt(f, arr, delimiter=",", **kw)
    return f.getvalue()


def test_fp16_csv_writer_is_byte_identical_to_savetxt(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.dirichlet(np.ones(20), size=3000).astype(np.float16)
    a[0, :6] = [0.0, 5.96e-8, 6.1e-5, 1.0, 65504.0, 0.333251953125]          # zero, subnormals, extremes
    f = io.StringIO()
    pp.savetxt_fp16(f, a)
    assert f.getvalue() == _np_text(a)
    assert f.getvalue().splitlines()[1].split(",")[0].__len__() == len("1.234741210937500000e-01")
    # every non-negative finite float16, 338 wide (rotamer one-hots are written through the same call)
    allv = np.arange(0x7C00, dtype=np.uint16).view(np.float16)
    wide = np.resize(allv, (94, 338))
    f = io.StringIO()
    pp.savetxt_fp16(f, wide)
    assert f.getvalue() == _np_text(wide)
    # negative / non-finite entries take the numpy path: still identical
    b = a[:50].copy()
    b[3, 2], b[7, 1], b[9, 0] = -0.5, np.nan, np.inf
    f = io.StringIO()
    pp.savetxt_fp16(f, b)
    assert f.getvalue() == _np_text(b)
    # append mode on a real text file, and a binary handle
    p = tmp_path / "m.csv"
    for chunk in (a[:7], a[7:19]):
        with open(p, "a") as fh:
            pp.savetxt_fp16(fh, chunk)
    assert p.read_text() == _np_text(a[:19])
    with open(tmp_path / "b.csv", "wb") as fh:
        pp.savetxt_fp16(fh, a[:5])
    assert (tmp_path / "b.csv").read_text() == _np_text(a[:5])


def test_label_writer_is_byte_identical_to_savetxt():
    rng = np.random.default_rng(1)
    y = np.eye(20)[rng.integers(0, 20, 777)]
    f = io.StringIO()
    pp.savetxt_onehot(f, y)
    assert f.getvalue() == _np_text(y, fmt="%i")
    y2 = y.copy()
    y2[5, 3] = 2                                               # not a one-hot: numpy path
    f = io.StringIO()
    pp.savetxt_onehot(f, y2)
    assert f.getvalue() == _np_text(y2, fmt="%i")
    f = io.StringIO()
    pp.savetxt_onehot(f, list(map(list, y[:3])))              # the drivers hand over lists of rows
    assert f.getvalue() == _np_text(y[:3], fmt="%i")


def test_native_e18_writer_is_byte_identical_to_savetxt(tmp_path):
    """timed_b200_format_csv_e18 (host threads, snprintf "%.18e") against np.savetxt: float32 and float64, specials,
    one row, more threads than rows."""
    rng = np.random.default_rng(2)
    a = rng.dirichlet(np.ones(338), size=257).astype(np.float32)
    a[0, :6] = [0.0, -1.5, np.inf, -np.inf, np.nan, 1e-45]
    for arr in (a, a[:1], rng.standard_normal((33, 7)), rng.standard_normal((5, 1)).astype(np.float32)):
        f = io.StringIO()
        pp.savetxt_e18(f, arr)
        assert f.getvalue() == _np_text(arr)
    p = tmp_path / "rot.csv"
    for chunk in (a[:20], a[20:41]):
        with open(p, "a") as fh:
            pp.savetxt_e18(fh, chunk)
    assert p.read_text() == _np_text(a[:41])
    f = io.StringIO()
    pp.savetxt_e18(f, np.arange(6).reshape(2, 3))            # integers: numpy path
    assert f.getvalue() == _np_text(np.arange(6).reshape(2, 3))


def test_native_csv_reader_matches_genfromtxt(tmp_path):
    """timed_b200_parse_csv behind load_matrix_csv (what sample.py reads the prediction matrix with) against
    np.genfromtxt: exact float64 values, specials, CRLF / blank lines, the 1-D result for a single row, and the
    reference's ValueError for ragged text."""
    import pytest
    rng = np.random.default_rng(3)
    a = rng.dirichlet(np.ones(338), size=300).astype(np.float16)
    p = tmp_path / "m.csv"
    with open(p, "w") as f:
        pp.savetxt_fp16(f, a)
    got = pp.load_matrix_csv(p)
    ref = np.genfromtxt(p, delimiter=",", dtype=np.float64)
    assert got.dtype == np.float64 and got.shape == (300, 338)
    np.testing.assert_array_equal(got, ref)
    p.write_text("1.5,nan,inf\n-2e-3,4,5\r\n\n")
    np.testing.assert_array_equal(pp.load_matrix_csv(p), np.genfromtxt(p, delimiter=","))
    p.write_text("1,2,3\n")
    assert pp.load_matrix_csv(p).shape == (3,)
    p.write_text("1,2,3\n4,5\n")
    with pytest.raises(ValueError):
        pp.load_matrix_csv(p)
