#!/usr/bin/env python
"""tests/golden/make_golden_wine.py -- fixture for stage (c), the band integration.

Imports the REFERENCE's own code/wine.py (and code/kurucz_inten.py) from /root/reference in this
container and records what readfilter / readkurucz / resample / bandintegrate return for the
filters the reference ships (inputs/filters/demo/fdemo01-10.dat on the demo wavenumber grid, the
four Spitzer IRAC *_fa.dat on the WASP-12b grid), with the shipped Kurucz model as the star.
tests/test_wine_golden.py checks oracle.readfilter/resample/bandintegrate and
bart_b200.api.filters_from_files against it.  Run once here; the .npz travels, /root/reference
does not.  usage: python tests/golden/make_golden_wine.py
"""
import os
import sys
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REF, "code"))
if not hasattr(np, "trapz"):            # numpy >= 2.4 dropped the old name wine.py uses
    np.trapz = np.trapezoid
import wine  # noqa: E402  (the reference module, unmodified)


def main():
    out = {}
    kfile = os.path.join(REF, "inputs", "kurucz", "fp00k2odfnew.pck")
    starfl, starwn, tmodel, gmodel = wine.readkurucz(kfile, 6000.0, 4.5)
    out["starwn"] = starwn
    out["starfl"] = starfl
    out["star_model"] = np.array([tmodel, gmodel])
    rng = np.random.default_rng(12345)
    sets = {
        "demo": (np.arange(2500.0, 5000.0 + 0.5, 1.0),
                 [os.path.join("inputs", "filters", "demo", "fdemo%02d.dat" % i) for i in range(1, 11)]),
        "w12": (np.arange(910.0, 3333.0 + 0.5, 1.0),
                [os.path.join("inputs", "filters", "spitzer_irac%d_fa.dat" % i) for i in range(1, 5)]),
    }
    for name, (specwn, files) in sets.items():
        spectrum = 1e4 * (1.0 + rng.random(specwn.size))          # planet flux, arbitrary positive
        out[name + "_specwn"] = specwn
        out[name + "_spectrum"] = spectrum
        out[name + "_files"] = np.array(files)
        for i, rel in enumerate(files):
            fwn, ftr = wine.readfilter(os.path.join(REF, rel))
            nif, istar, idx = wine.resample(specwn, fwn, ftr, starwn, starfl)
            idx = idx[0]
            key = "%s_%d_" % (name, i)
            out[key + "fwn"] = fwn
            out[key + "ftr"] = ftr
            out[key + "nifilter"] = nif
            out[key + "istarfl"] = istar
            out[key + "idx"] = idx
            # BARTfunc.py:386-396: eclipse (flux ratio x rprs^2) and transit (modulation) forms
            rprs = 0.117
            out[key + "band_eclipse"] = wine.bandintegrate(spectrum[idx] / istar * rprs ** 2, specwn, nif, (idx,))
            out[key + "band_transit"] = wine.bandintegrate(spectrum[idx], specwn, nif, (idx,))
    np.savez_compressed(os.path.join(HERE, "wine.npz"), **out)
    print("wrote", os.path.join(HERE, "wine.npz"), "with", len(out), "arrays")


if __name__ == "__main__":
    main()
