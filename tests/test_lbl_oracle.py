"""CPU: the oracle's line-by-line forward mode (no opacity file; tau.c:163-175,253-264 ->
computemolext(permol=0), restated in oracle/transit_oracle.c) against golden vectors produced by
the UNMODIFIED reference run without `opacityfile` (tests/golden/make_golden.py lbl_golden)."""
import numpy as np
import pytest

import cases
from util import relerr, tau_relerr


@pytest.mark.parametrize("name", list(cases.LBL_CASES))
def test_lbl_oracle_matches_reference_golden(name, built, workdir):
    from oracle import oracle as orc
    case, models = cases.build_lbl_case(name, workdir)
    g = np.load(cases.golden_path(name))
    assert cases.sha(np.fromfile(case["tli"], dtype=np.uint8)) == str(g["tli_sha"])
    assert cases.sha(models) == str(g["models_sha"])
    O = orc.Oracle(case["cfg"])
    assert O.lbl is not None
    assert np.array_equal(O.wn, g["wn"])
    for m in range(models.shape[0]):
        o = O.run(models[m], inter=True)
        assert relerr(o["radius"], g["radius"][m]) < 1e-13
        assert np.array_equal(o["last"], g["last"][m])
        ext_ref = g["ext"][m]
        comp = np.abs(ext_ref).sum(axis=1) > 0        # the reference evaluates layers lazily
        assert comp.sum() >= 3
        assert np.array_equal(o["ext"][comp] > 0, ext_ref[comp] > 0)
        # float32 Voigt table evaluated in long double by the reference's pu library: 1e-6 is the
        # north-star tolerance for built opacities
        assert relerr(o["ext"][comp], ext_ref[comp]) < 1e-6
        assert tau_relerr(o["tau"], g["tau"][m], g["last"][m]) < 1e-6
        assert relerr(o["spectrum"], g["spectra"][m]) < 1e-6


def test_total_mode_differs_from_permol(built, workdir):
    """permol = 0 is not the density-weighted sum of the per-molecule rows: the weak-line cut
    uses ONE strongest line across all molecules (extinction.c:296,405-427)."""
    from oracle import oracle as orc
    case, models = cases.build_lbl_case("lbl_transit_2mol", workdir)
    B = orc.BuilderOracle(case["cfg"], grid_temps=False)
    T = 1500.0
    atm = B.atm
    r = 3
    dens = 1.66053886e-24 * atm["q"][r] * (atm["press"][r] * atm["pfct"]) / 1.380658e-16 / T * B.spec_mass
    Z = np.array([np.interp(T, i["T"], i["Z"]) for i in B.tli["isos"]])
    k = B.total(T, dens, Z)
    assert k.shape == (len(B.wn),) and np.isfinite(k).all() and (k > 0).any()


@pytest.mark.parametrize("name", list(cases.LBL_CASES))
def test_column_math_in_lbl_mode(name, built, workdir):
    """The device-side half of the line-by-line mode on the CPU emulator (tests/cpu_emu): atm_prep
    with DevConfig::lbl presents ext[model][layer][wave] to the unchanged column code as a
    one-molecule grid with bracket weights (1, 0).  Given the oracle's per-layer extinction the
    emulated columns must reproduce the oracle's (= the reference's) spectra, tau and last[]."""
    from oracle import oracle as orc
    from emu import Emu
    case, models = cases.build_lbl_case(name, workdir)
    g = np.load(cases.golden_path(name))
    O = orc.Oracle(case["cfg"])
    E = Emu(case["cfg"])
    for m in range(models.shape[0]):
        o = O.run(models[m], inter=True)
        E.set_lbl_ext(o["ext"])
        e = E.run(models[m])
        assert e["status"] == 0
        assert np.array_equal(e["ext"], o["ext"])            # weight 1, second bracket weight 0: exact
        assert np.array_equal(e["last"], g["last"][m])
        assert tau_relerr(e["tau"], g["tau"][m], g["last"][m]) < 1e-6
        assert relerr(e["spectrum"], o["spectrum"]) < 1e-9
        assert relerr(e["spectrum"], g["spectra"][m]) < 1e-6
    # a layer outside the TLI temperature range is rejected (makesample.c:488-503)
    bad = models[0].copy()
    bad[2] = 69.0
    assert E.run(bad)["status"] != 0
