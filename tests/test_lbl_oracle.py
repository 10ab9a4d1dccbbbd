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
