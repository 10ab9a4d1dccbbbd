"""The polynomial tables behind the input converter's E_2(x) (bart_b200/csrc/expint2_table.inc, written
by tools/gen_expint2_table.py; device code retrieval.cu expint2): evaluated here exactly as the kernel
evaluates them (same branches, same Horner order) against scipy.special.expn(2, x) -- what
code/PT.py:736 calls -- and against mpmath."""
import os
import re
import numpy as np
from scipy.special import expn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_tables():
    src = open(os.path.join(ROOT, "bart_b200", "csrc", "expint2_table.inc")).read()
    ns, nc, nk = [int(v) for v in re.search(r"kE2SeriesN = (\d+), kE2ChebN = (\d+), kE2Intervals = (\d+)", src).groups()]
    vals = np.array([float(v) for v in re.findall(r"^\s+(-?[0-9][0-9.e+-]*)[,}]", src, re.M)])
    assert vals.size == ns + nc * nk
    return vals[:ns], vals[ns:].reshape(nk, nc)


def e2_like_device(x, ser, tab):
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lo = x < 1.0
    h = np.full(lo.sum(), ser[-1])
    for c in ser[-2::-1]:
        h = h * x[lo] + c
    out[lo] = np.exp(-x[lo]) - x[lo] * (h - np.log(x[lo]))
    xh = x[~lo]
    k = np.frexp(xh)[1] - 1                                   # binade: 2^k <= x < 2^(k+1)
    u = xh * np.ldexp(1.0, 1 - k) - 3.0
    p = tab[k, -1].copy()
    for j in range(tab.shape[1] - 2, -1, -1):
        p = p * u + tab[k, j]
    out[~lo] = p * np.exp(-xh) / xh
    return out


def test_tables_against_scipy_expn():
    ser, tab = load_tables()
    rng = np.random.default_rng(5)
    x = np.concatenate([10.0 ** rng.uniform(-14, 0, 20000), rng.uniform(0.9, 1.1, 5000),
                        2.0 ** rng.uniform(0, 9.3, 40000), 2.0 ** np.arange(0, 10), np.nextafter(2.0 ** np.arange(1, 10), 0)])
    got, ref = e2_like_device(x, ser, tab), expn(2, x)
    rel = np.abs(got - ref) / ref
    # scipy's expn (cephes: series / continued fraction run to convergence) is itself up to ~3.5e-15
    # from the true value just below x = 1; the tables are held to mpmath below
    assert rel.max() < 5e-15, (rel.max(), x[rel.argmax()])


def test_tables_against_mpmath_and_regenerate():
    import mpmath as mp
    mp.mp.dps = 40
    ser, tab = load_tables()
    rng = np.random.default_rng(6)
    x = np.concatenate([[1e-9, 0.03, 0.5, 0.999, 1.0, 1.5, 2.0, 3.999, 7.3, 31.0, 100.0, 511.9, 600.0],
                        10.0 ** rng.uniform(-10, 0, 600), rng.uniform(0.8, 1.0, 400), 2.0 ** rng.uniform(0, 9.2, 1500)])
    got = e2_like_device(x, ser, tab)
    for xv, g in zip(x, got):
        ref = mp.expint(2, mp.mpf(float(xv)))
        assert abs((mp.mpf(float(g)) - ref) / ref) < 1.5e-15, xv
    # the series coefficients are -gamma and -(-1)^k / (k k!)
    assert ser[0] == float(-mp.euler) and ser[1] == 1.0 and ser[2] == -0.25
