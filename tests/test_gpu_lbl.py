"""GPU (B200): the line-by-line forward mode (no opacity file).  run_transit /
bart_run_batch compute every layer's molecular extinction with the builder kernels at the layer's
own temperature (tau.c:163-175,253-264 -> computemolext(permol=0)) and feed the column kernels.
Checked through the C ABI against the oracle and against golden vectors of the UNMODIFIED
reference run without `opacityfile`."""
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
    assert relerr(spectra2, spectra) < 1e-13
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
    tr.free_memory()
