"""GPU (B200): the line-by-line forward mode (no opacity file).  run_transit /
bart_run_batch compute every layer's molecular extinction with the builder kernels at the layer's
own temperature (tau.c:163-175,253-264 -> computemolext(permol=0)) and feed the column kernels.
Checked through the C ABI against the oracle and against golden vectors of the UNMODIFIED
reference run without `opacityfile`."""
import os
import numpy as np
import pytest

import cases
from util import relerr, tau_relerr

pytestmark = pytest.mark.gpu
TOL = 1e-6


@pytest.fixture(scope="module")
def api(built):
    from bart_b200 import api as a
    a.device_info()
    return a


@pytest.mark.parametrize("name", list(cases.LBL_CASES))
def test_lbl_forward_vs_oracle_and_reference(name, api, workdir):
    from oracle import oracle as orc
    case, models = cases.build_lbl_case(name, workdir)
    g = np.load(cases.golden_path(name))
    assert cases.sha(models) == str(g["models_sha"])
    tr = api.Transit(case["cfg"])
    O = orc.Oracle(case["cfg"])
    assert np.array_equal(tr.get_waveno_arr(), g["wn"])
    ext = tr.extinction_batch(models, total=False)
    tr.debug_keep(True)
    spectra, status = tr.run_batch(models)
    assert (status == 0).all()
    for m in range(models.shape[0]):
        o = O.run(models[m], inter=True)
        assert np.array_equal(ext[m] > 0, o["ext"] > 0)
        assert relerr(ext[m], o["ext"]) < TOL
        ext_ref = g["ext"][m]
        comp = np.abs(ext_ref).sum(axis=1) > 0           # the reference evaluates layers lazily
        assert relerr(ext[m][comp], ext_ref[comp]) < TOL
        last = tr.debug_get("last", m).astype(np.int64)
        assert np.array_equal(last, o["last"]) and np.array_equal(last, g["last"][m])
        tau = tr.debug_get("tau", m).reshape(tr.nwave, tr.nlayer)
        assert tau_relerr(tau, g["tau"][m], g["last"][m]) < TOL
        assert relerr(spectra[m], o["spectrum"]) < TOL
        assert relerr(spectra[m], g["spectra"][m]) < TOL
    tr.debug_keep(False)
    spectra2, _ = tr.run_batch(models)
    # specialised (exp(-tau/cos 60) by squaring) vs run-time-count kernel: degree-4 exponentials
    assert relerr(spectra2, spectra) < 1e-10
    one = tr.run_transit(models[1])                      # the reference's own entry point
    assert np.array_equal(one, spectra2[1])
    tr.free_memory()


def test_lbl_rejects_out_of_range_temperature(api, workdir):
    """A layer outside the TLI temperature range makes the reference exit (makesample.c:488-503);
    the batched call rejects that model only."""
    case, models = cases.build_lbl_case("lbl_eclipse", workdir)
    tr = api.Transit(case["cfg"])
    bad = models.copy()
    bad[0, 3] = 69.0                                     # below the 70 K floor of the synthetic TLI
    spectra, status = tr.run_batch(bad)
    assert status[0] != 0 and (spectra[0] == -1).all()
    assert status[1] == 0
    ok, _ = tr.run_batch(models)
    assert np.array_equal(spectra[1], ok[1])
    # a rejected model in the middle of a batch: the accepted runs on both sides are unaffected
    tri = np.concatenate([models[:1], bad[:1], models[1:2]])
    s3, st3 = tr.run_batch(tri)
    assert st3[0] == 0 and st3[1] != 0 and st3[2] == 0
    assert np.array_equal(s3[0], ok[0]) and np.array_equal(s3[2], ok[1]) and (s3[1] == -1).all()
    # the legacy single-model call fails like the reference (makesample.c:488-503)
    with pytest.raises(api.BartError):
        tr.run_transit(bad[0])
    tr.free_memory()


def test_lbl_equals_grid_mode_at_a_grid_temperature(api, workdir):
    """Cross-check of the two opacity paths: for an isothermal atmosphere whose temperature is a
    node of the opacity grid the T-interpolation is exact, and with a single line-list molecule
    computemolext(permol=0) = density x computemolext(permol=1) -- so the forward model without an
    opacity file must reproduce the one that reads the grid built from the same line list."""
    from bart_b200 import synth
    kw = dict(shape="tiny", solution="eclipse", nlayer=16, with_grid=False, nlines=2500, seed=5152,
              ethresh=1e-6, tlow=400.0, thigh=2800.0, tempdelt=600.0)
    grid_case = synth.make_case(os.path.join(workdir, "lbl_vs_grid_g"), **kw)
    if os.path.exists(grid_case["opacity"]):
        os.remove(grid_case["opacity"])
    lbl_case = synth.make_case(os.path.join(workdir, "lbl_vs_grid_l"), no_opacity=True, **kw)
    models = synth.make_models(grid_case, 2, seed=63, molfit=("CH4",))
    nl = grid_case["nlayer"]
    models[0, :nl] = 1000.0                              # grid nodes: 400, 1000, 1600, 2200, 2800
    models[1, :nl] = 1600.0
    tr = api.Transit(grid_case["cfg"])                   # builds the grid file first (missing)
    ext_g = tr.extinction_batch(models, total=False)
    spec_g, st = tr.run_batch(models)
    assert (st == 0).all()
    tr.free_memory()
    tr = api.Transit(lbl_case["cfg"])
    ext_l = tr.extinction_batch(models, total=False)
    spec_l, st = tr.run_batch(models)
    assert (st == 0).all()
    tr.free_memory()
    assert (ext_l > 0).any()
    assert np.array_equal(ext_l > 0, ext_g > 0)
    assert relerr(ext_l, ext_g) < 1e-12
    assert relerr(spec_l, spec_g) < 1e-10


@pytest.mark.parametrize("k", list(cases.FUZZ_LBL))
def test_randomised_line_by_line_configurations(k, api, workdir):
    """cases.build_lbl_fuzz_case (the oracle agrees with the compiled reference on each,
    tests/test_oracle_vs_reference.py) through the CUDA path."""
    from oracle import oracle as orc
    case, models = cases.build_lbl_fuzz_case(k, workdir)
    tr = api.Transit(case["cfg"])
    O = orc.Oracle(case["cfg"])
    ext = tr.extinction_batch(models, total=False)
    tr.debug_keep(True)
    spectra, status = tr.run_batch(models)
    assert (status == 0).all()
    for m in range(models.shape[0]):
        o = O.run(models[m], inter=True)
        assert np.array_equal(ext[m] > 0, o["ext"] > 0)
        assert relerr(ext[m], o["ext"]) < TOL
        assert np.array_equal(tr.debug_get("last", m).astype(np.int64), o["last"])
        assert relerr(spectra[m], o["spectrum"]) < TOL
    tr.free_memory()
