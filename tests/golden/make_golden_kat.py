#!/usr/bin/env python
"""tests/golden/kat_voigt.npz: a subsample of the Voigt table the reference's own test directory
records (modules/transit/transit/test/voigt.080904.dat: unit-area Voigt profile, Lorentz half
width 1.5, Doppler half width 1, on [-75, 75] in 10000 bins with 16 sub-bin samples, 4 significant
digits, written by the 2004 code).  Every 25th row is kept, all 16 sub-bin columns.

    python tests/golden/make_golden_kat.py
"""
import os
import re
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/modules/transit/transit/test/voigt.080904.dat"


def main():
    d = np.loadtxt(SRC)
    n = d.shape[0]
    rows = np.arange(0, n, 25)
    # the first column is printed with 5 digits: the grid is -75 + i * 150 / (n - 1)
    np.savez_compressed(os.path.join(HERE, "kat_voigt.npz"), nrows=n, rows=rows, table=d[rows, 1:],
                        alphaL=1.5, alphaD=1.0, halfrange=75.0, subbins=32, first_sub=-8)
    print("kat_voigt.npz: %d of %d rows, %d columns" % (len(rows), n, d.shape[1] - 1))


if __name__ == "__main__":
    main()
