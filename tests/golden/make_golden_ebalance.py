#!/usr/bin/env python
"""tests/golden/make_golden_ebalance.py -- fixture for the energy-balance rejection of the input
converter (the third rejection test).

Executes the statements of the reference's code/BARTfunc.py:366-383 verbatim, with the reference's
own code/constants.py and code/reader.py imported from /root/reference and the TEP file the
reference ships (examples/WASP-12b/WASP-12b.tep), on seeded spectra that straddle the threshold.
tests/test_retrieval_oracle.py checks oracle.retrieval_oracle.energy_balance against it (CPU);
tests/test_gpu_retrieval.py checks the CUDA kernel (bart_energy_balance) against it.
usage: python tests/golden/make_golden_ebalance.py
"""
import os
import sys
import numpy as np
import scipy.constants as sc

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REF, "code"))
if not hasattr(np, "trapz"):
    np.trapz = np.trapezoid
import constants as c   # noqa: E402  (the reference module)
import reader as rd     # noqa: E402  (the reference module)


def main():
    tep = rd.File(os.path.join(REF, "examples", "WASP-12b", "WASP-12b.tep"))
    # BARTfunc.py:169
    rplanet = float(tep.getvalue('Rp')[0]) * c.Rjup
    specwn = np.arange(910.0, 3333.0 + 0.5, 1.0)
    rng = np.random.default_rng(20261018)
    M = 16
    shape = 1.0 + 0.5 * rng.random((M, specwn.size))
    # scale the spectra so that E_out spans 0.25 .. 4 E_in
    tstar = float(tep.getvalue('Ts')[0])
    rstar = float(tep.getvalue('Rs')[0]) * c.Rsun
    sma = float(tep.getvalue('a')[0]) * sc.au
    e_in0 = c.sig * tstar ** 4 * rstar ** 2 * np.pi * rplanet ** 2 / sma ** 2 * 1e7
    target = e_in0 * np.exp(np.linspace(np.log(0.25), np.log(4.0), M))
    spectra = np.empty_like(shape)
    for m in range(M):
        raw = np.trapz(shape[m], specwn) * 4 * (rplanet * 100) ** 2
        spectra[m] = shape[m] * target[m] / raw
    e_in_all, e_out_all, rejected = np.zeros(M), np.zeros(M), np.zeros(M, dtype=bool)
    for m in range(M):
        spectrum = spectra[m]
        # --- BARTfunc.py:367-379, verbatim
        # Stellar temperature in K:
        tstar = float(tep.getvalue('Ts')[0])
        # Stellar radius (in meters):
        rstar = float(tep.getvalue('Rs')[0]) * c.Rsun
        # Semi-major axis (in meters):
        sma   = float(tep.getvalue( 'a')[0]) * sc.au

        # Calculate energy in and energy out, in cgs
        j2erg = 1e7
        e_in  = c.sig*tstar**4 * rstar**2 * np.pi*rplanet**2 / sma**2 * j2erg
        e_out = np.trapz(spectrum, specwn) * 4 * (rplanet*100)**2     
        if e_out > e_in:
            rejected[m] = True
        # ---
        e_in_all[m], e_out_all[m] = e_in, e_out
    np.savez_compressed(os.path.join(HERE, "ebalance.npz"), specwn=specwn, spectra=spectra,
                        tstar=tstar, rstar=rstar, sma=sma, rplanet=rplanet, e_in=e_in_all,
                        e_out=e_out_all, rejected=rejected)
    print("ebalance.npz: %d spectra, %d rejected; e_in %.6e" % (M, rejected.sum(), e_in_all[0]))


if __name__ == "__main__":
    main()
