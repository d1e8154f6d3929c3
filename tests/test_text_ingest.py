"""The library's text reader against the reference's text-format rules
(apop_conversions.m4.c:339-403) on the host -- no device needed -- and, on a B200, the same files
ingested straight into the resident layout."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ab():
    import apophenia_b200 as m
    return m


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return p


def test_text_rules_on_host(ab, tmp_path):
    # header line, comments, blank lines, white space, missing values, text fields, number syntax
    p = _write(tmp_path, "a.csv", "\n".join([
        "# a comment before the names",
        "y, x1 ,x2|x3",
        "1, 2.5 ,3e2|-4",
        "",
        "   \t ",
        "0,,NaN| .5   # two commas: missing value",
        "1,abc,+7.25|1e-3",
        "0,\t9,8,7",                       # comma then tab: the tab is a formatting glitch, not a missing value
        "1,inf,-INFINITY,0x1p-2",
        "0,123456789012345678,1.7976931348623157e308,4.9e-324",
        "1,2,",                             # short row ending on a delimiter: trailing NaN, then padded with 0
    ]) + "\n")
    got = ab.text_to_host(p)
    # "abc": strtod converts nothing -> NaN, as the reference's code writes (apop_conversions.m4.c:659-666; its
    # format page, :363, still says such fields are "taken as zeros")
    want = np.array([[1, 2.5, 300, -4], [0, np.nan, np.nan, .5], [1, np.nan, 7.25, 1e-3], [0, 9, 8, 7],
                     [1, np.inf, -np.inf, 0.25], [0, 123456789012345678.0, 1.7976931348623157e308, 4.9e-324],
                     [1, 2, np.nan, 0]])
    assert got.shape == want.shape
    assert np.array_equal(got, want, equal_nan=True)
    # no names, row names, blank-delimited
    p = _write(tmp_path, "b.txt", "r1 1   2 3\nr2 4 5\t6\n")
    got = ab.text_to_host(p, has_row_names=True, has_col_names=False, delimiters=" \t")
    assert np.array_equal(got, [[1, 2, 3], [4, 5, 6]])
    # CRLF line ends and no final newline
    p = _write(tmp_path, "c.csv", "a,b\r\n1,2\r\n3,4")
    assert np.array_equal(ab.text_to_host(p), [[1, 2], [3, 4]])
    # quoted fields and backslash escapes (apop_conversions.m4.c:352-358): the marks are stripped and the text is read
    # verbatim (a delimiter or # inside is text; a quoted number is a number, other text does not convert: NaN); the
    # character after a backslash is an ordinary character
    p = _write(tmp_path, "q.csv", "\n".join([
        '"Males, 30-40",plain,"with # inside"',
        '"2.5",3,4  # comment',
        '1,"text, with comma",7',
        '" 6 ",\\-2,"8"',                       # inner spaces stay inside the field (still the number 6); \- is a plain minus
        '5,\\,,9',                              # an escaped comma is text (NaN), then an ordinary field
        '"",1,2',                                 # an empty quoted field has length 0: NaN
        '7,12abc,"3 apples"',                     # strtod converts the leading number of a field
    ]) + "\n")
    got = ab.text_to_host(p)
    nan = np.nan
    assert np.array_equal(got, [[2.5, 3, 4], [1, nan, 7], [6, -2, 8], [5, nan, 9], [nan, 1, 2], [7, 12, 3]], equal_nan=True)
    # refused: over-long rows, missing file
    for bad in ("a,b\n1,2\n3,4,5\n",):
        with pytest.raises(ab.B200Error):
            ab.text_to_host(_write(tmp_path, "bad.csv", bad))
    with pytest.raises(ab.B200Error):
        ab.text_to_host(tmp_path / "does_not_exist.csv")


def test_fixed_width_layout_on_host(ab, tmp_path):
    """.field_ends of apop_text_to_data (parse_a_fixed_line, apop_conversions.m4.c:434-464): the example of the
    reference's format page, numeric fields, a missing last end, short lines, row names"""
    nan = np.nan
    # the page's example (:405-411): NUM | LE | OL; the text columns do not convert
    p = _write(tmp_path, "f1.txt", "NUMLEOL\n123AABB\n456CCDD\n")
    assert np.array_equal(ab.text_to_host(p, field_ends=[3, 5, 7]), [[123, nan, nan], [456, nan, nan]], equal_nan=True)
    # numbers with blanks inside their fields; '#' and ',' are ordinary characters here; an all-blank field is NaN;
    # a line without characters is skipped; no name line
    p = _write(tmp_path, "f2.txt", "  1 2.5-3e0\n\n 10   7  .5\n  2    #9,1\n")
    got = ab.text_to_host(p, has_col_names=False, field_ends=[3, 7, 11])
    assert np.array_equal(got, [[1, 2.5, -3], [10, 7, .5], [2, nan, nan]], equal_nan=True)
    # the last end left out: what follows the last given end is one more field; a line that stops early has fewer
    # fields and its row is padded with 0 like the freshly reallocated matrix row of the reference
    p = _write(tmp_path, "f3.txt", "aabbcc\n 1 2 3\n 4 5\n")
    assert np.array_equal(ab.text_to_host(p, field_ends=[2, 4]), [[1, 2, 3], [4, 5, 0]])
    # row names in the first field
    p = _write(tmp_path, "f4.txt", "r1  1.5 -2\nr2 3.25  4\n")
    assert np.array_equal(ab.text_to_host(p, has_row_names=True, has_col_names=False, field_ends=[2, 7, 10]), [[1.5, -2], [3.25, 4]])
    # a line longer than the first data line has more fields than the matrix has columns: refused (the reference stops
    # reading there with error 't'); ends are counted from 1
    with pytest.raises(ab.B200Error):
        ab.text_to_host(_write(tmp_path, "f5.txt", " 1 2\n 3 4 5\n"), has_col_names=False, field_ends=[2, 4])
    with pytest.raises(ab.B200Error):
        ab.text_to_host(_write(tmp_path, "f6.txt", " 1 2\n"), has_col_names=False, field_ends=[0, 2])
    # the same answer from many segments and threads as from a plain split
    rng = np.random.default_rng(5)
    big = rng.integers(-9999, 99999, (200000, 6)).astype(float) / 100
    lines = ["".join(f"{v:10.2f}" for v in row) for row in big]
    pb = _write(tmp_path, "fbig.txt", "\n".join(lines) + "\n")
    assert np.array_equal(ab.text_to_host(pb, has_col_names=False, field_ends=[10, 20, 30, 40, 50, 60]), big)


def test_decimal_fast_path_is_correctly_rounded(ab, tmp_path):
    """every value printed with repr() must come back bit for bit (fast path and strtod fallback)"""
    rng = np.random.default_rng(3)
    vals = np.concatenate([rng.uniform(-1, 1, 4000), rng.standard_normal(2000) * 1e6, 10.0 ** rng.uniform(-300, 300, 2000),
                           rng.integers(-10**15, 10**15, 1000).astype(float), [0.1, 0.2, 0.3, 1e22, 1e23, 5e-324, 123456.789e3]])
    vals = vals[: vals.size // 4 * 4].reshape(-1, 4)
    lines = ["c0,c1,c2,c3"] + [",".join(repr(float(v)) for v in row) for row in vals]
    short = [",".join(f"{v:.6f}" for v in row) for row in vals[:500]]      # the common short-decimal form
    p = _write(tmp_path, "v.csv", "\n".join(lines + short) + "\n")
    got = ab.text_to_host(p)
    assert np.array_equal(got[: vals.shape[0]], vals)
    assert np.array_equal(got[vals.shape[0]:], np.array([[float(f"{v:.6f}") for v in row] for row in vals[:500]]))
    # many segments, many threads: a file larger than the segment size parses to the same matrix as numpy
    big = rng.uniform(-1, 1, (120000, 8))
    pb = tmp_path / "big.csv"
    np.savetxt(pb, big, delimiter=",", header="a,b,c,d,e,f,g,h", comments="", fmt="%.17g")
    assert np.array_equal(ab.text_to_host(pb), big)


@pytest.mark.gpu
def test_text_to_resident_and_fit(ab, tmp_path):
    """file -> HBM -> device prep -> fit, no host matrix at any point"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import gen_probit_logit_sample
    from oracle import oracle as o
    ab.init(0)
    n, k = 70013, 6
    X, y, beta = gen_probit_logit_sample(n, k, seed=31)
    raw = X.copy(); raw[:, 0] = y
    p = tmp_path / "probit.csv"
    np.savetxt(p, raw, delimiter=",", header=",".join(f"c{j}" for j in range(k)), comments="", fmt="%.17g")
    d = ab.text_to_resident(p, outcome_col0=True)
    Xd, yd = d.download()
    Xs = raw.copy(); Xs[:, 0] = 1.0
    assert np.array_equal(Xd, Xs) and np.array_equal(yd, y)
    cats, start = ab.prep_outcome(d)
    assert np.array_equal(cats, [0.0, 1.0])
    ll, g = ab.ll_score(ab.PROBIT, d, beta)
    want = float(o.probit_ll(Xs, y, beta, o.EXACT))
    assert abs(ll - want) <= 1e-12 * abs(want)
    # apop_estimate through the mirror on the device-only data set
    from apophenia_b200 import abi, host
    est, params, rep = host.estimate(abi.PROBIT, d.ptr, method="Newton", tolerance=1e-10, relative_tolerance=True, max_iterations=100)
    assert rep["status"] == 0 and np.max(np.abs(params - beta)) < 0.07      # the reference's own recovery test, tests/test_apop.c:935-959
    # without an outcome column the matrix alone is resident (iid models)
    d2 = ab.text_to_resident(p, outcome_col0=False)
    X2, _ = d2.download()
    assert np.array_equal(X2, raw)
    d.free(); d2.free()
