"""csrc/csv_io.cpp against numpy: the files scUnif.py reads and writes (scUnif.py:85-94)
must be byte-identical to np.savetxt(delimiter=',') and parse to the same doubles as
np.loadtxt(delimiter=',').  CPU only."""
import io
import time

import numpy as np
import pytest

from ursm_b200 import _lib, csvio


def _np_bytes(a):
    buf = io.BytesIO()
    np.savetxt(buf, a, delimiter=",")
    return buf.getvalue()


CASES = {
    "counts": lambda rs: rs.poisson(3.0, (37, 53)).astype(float),
    "simplex": lambda rs: rs.dirichlet(np.ones(41), 5).T,
    "exp_S": lambda rs: rs.randint(0, 51, (64, 33)) / 50.0,
    "wide_range": lambda rs: np.exp(rs.normal(0, 60, (29, 7))) * rs.choice([-1, 1], (29, 7)),
    "vector": lambda rs: rs.normal(size=17),
    "single_row": lambda rs: rs.normal(size=(1, 9)),
    "single_col": lambda rs: rs.normal(size=(9, 1)),
    "scalar_like": lambda rs: np.array([[3.25]]),
    "specials": lambda rs: np.array([[0.0, -0.0, 1e-320, np.inf], [-np.inf, np.nan, 1.7976931348623157e308, 5e-324]]),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_write_is_byte_identical_to_savetxt(tmp_path, name):
    a = CASES[name](np.random.RandomState(3))
    path = tmp_path / (name + ".csv")
    csvio.savetxt_csv(str(path), a, nthreads=3)
    assert path.read_bytes() == _np_bytes(a)


@pytest.mark.parametrize("name", sorted(CASES))
def test_read_matches_loadtxt(tmp_path, name):
    a = CASES[name](np.random.RandomState(4))
    path = tmp_path / (name + ".csv")
    np.savetxt(str(path), a, delimiter=",")
    want = np.loadtxt(str(path), delimiter=",")
    got = csvio.loadtxt_csv(str(path), nthreads=3)
    assert got.shape == want.shape
    np.testing.assert_array_equal(got, want)     # NaN == NaN under assert_array_equal
    assert np.array_equal(np.signbit(got), np.signbit(want))


def test_plain_integers_crlf_and_blank_lines(tmp_path):
    path = tmp_path / "ints.csv"
    path.write_bytes(b"1,2,3\r\n\r\n4, 5,6\n7,8,9")
    np.testing.assert_array_equal(csvio.loadtxt_csv(str(path)), np.arange(1, 10.0).reshape(3, 3))


def test_errors(tmp_path):
    with pytest.raises(_lib.UrsmError):
        csvio.loadtxt_csv(str(tmp_path / "missing.csv"))
    bad = tmp_path / "ragged.csv"
    bad.write_text("1,2,3\n4,5\n")
    with pytest.raises(_lib.UrsmError):
        csvio.loadtxt_csv(str(bad))
    bad.write_text("1,2,x\n")
    with pytest.raises(_lib.UrsmError):
        csvio.loadtxt_csv(str(bad))
    with pytest.raises(_lib.UrsmError):
        csvio.savetxt_csv(str(tmp_path / "no_such_dir" / "a.csv"), np.ones((2, 2)))


def test_round_trip_and_speed(tmp_path):
    """Size-independent property (write -> read is the identity on doubles) on a matrix the
    size of the demo's exp_S x 400, plus the measurement quoted in DESIGN.md."""
    rs = np.random.RandomState(0)
    a = rs.randint(0, 101, (2000, 1000)) / 100.0
    path = str(tmp_path / "big.csv")
    t0 = time.perf_counter()
    csvio.savetxt_csv(path, a)
    t1 = time.perf_counter()
    b = csvio.loadtxt_csv(path)
    t2 = time.perf_counter()
    np.testing.assert_array_equal(a, b)
    print("native: write %.3f s, read %.3f s for %d values" % (t1 - t0, t2 - t1, a.size))
