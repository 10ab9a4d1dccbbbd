"""GPU (B200): the CUDA path, called through the C ABI (ctypes on libbart_b200.so), against
(1) the oracle on the same seeded inputs, (2) the golden vectors of the unmodified reference,
(3) size-independent properties at the full BASELINE shapes.  Tolerance for spectra and band
fluxes: 1e-6 relative in fp64 (BASELINE.json north star); `last[]` (the layer index where the
optical depth exceeds toomuch) must be identical."""
import os
import numpy as np
import pytest

import cases
from util import relerr, tau_relerr, apply_setters

pytestmark = pytest.mark.gpu
TOL = 1e-6            # north-star tolerance (against the reference's golden vectors)
TIGHT = 1e-8          # what the CUDA path is held to against the oracle; observed ~1e-11


@pytest.fixture(scope="module")
def api(built):
    from bart_b200 import api as a
    info = a.device_info()
    assert info["cc"][0] >= 10, info
    return a


@pytest.mark.parametrize("name", list(cases.CASES))
def test_spectra_vs_oracle_and_golden(name, api, get_case, monkeypatch):
    from oracle import oracle as orc
    case, models, setters = get_case(name)
    g = np.load(cases.golden_path(name))
    assert cases.sha(case["grid"]) == str(g["grid_sha"])
    assert cases.sha(models) == str(g["models_sha"])
    tr = api.Transit(case["cfg"])
    apply_setters(tr, setters)
    O = orc.Oracle(case["cfg"])
    apply_setters(O, setters)
    assert np.array_equal(tr.get_waveno_arr(), g["wn"])
    tr.debug_keep(True)
    spectra, status = tr.run_batch(models)
    assert (status == 0).all()
    wsel = g["tau_wsel"]
    for m in range(models.shape[0]):
        o = O.run(models[m], inter=True)
        assert relerr(tr.debug_get("radius", m), o["radius"]) < 1e-12
        last = tr.debug_get("last", m).astype(np.int64)
        assert np.array_equal(last, o["last"])
        assert np.array_equal(last, g["last"][m]), "last[] differs from the reference"
        tau = tr.debug_get("tau", m).reshape(tr.nwave, tr.nlayer)
        assert tau_relerr(tau, o["tau"], o["last"]) < 1e-8
        assert tau_relerr(tau[wsel], g["tau_sample"][m], g["last"][m][wsel]) < 1e-8
        assert relerr(spectra[m], o["spectrum"]) < TIGHT
        assert relerr(spectra[m], g["spectra"][m]) < TOL
    tr.debug_keep(False)
    spectra2, _ = tr.run_batch(models)
    # the production kernel is specialised (compile-time counts, exp(-tau/cos 60) by squaring); the
    # introspection kernel above is the run-time-count instantiation of the same code.  The angle
    # exponentials are degree-4 (2.6e-12 each), so squaring one instead of evaluating it differs
    # at that level
    assert relerr(spectra2, spectra) < 1e-10, "keep/no-keep kernels disagree"
    for m in range(models.shape[0]):
        assert relerr(spectra2[m], o["spectrum"]) < TIGHT if m == models.shape[0] - 1 else True
        assert relerr(spectra2[m], g["spectra"][m]) < TOL
        # the reference's own single-model entry point gives the same numbers as the batch
        one = tr.run_transit(models[m])
        assert np.array_equal(one, spectra2[m])
    # batches this small take the latency kernel (eclipse: lanes <-> layers scan); the throughput
    # kernel and the slot kernel on the same models, against the same oracle / golden spectra
    for mode in ("0", "1"):
        monkeypatch.setenv("BART_ECL_SMALL", mode)
        spectra3, _ = tr.run_batch(models)
        assert relerr(spectra3, spectra2) < 1e-9
        for m in range(models.shape[0]):
            assert relerr(spectra3[m], g["spectra"][m]) < TOL
        assert relerr(spectra3[-1], o["spectrum"]) < TIGHT
    monkeypatch.delenv("BART_ECL_SMALL")
    tr.free_memory()


@pytest.mark.parametrize("name", ["tiny_eclipse", "small4_eclipse_cloud"])
def test_opacity_lookup_kernel(name, api, get_case):
    """Stand-alone K1 (extinction.c:534-581): e[layer][wn] against the oracle, <= 1e-12."""
    from oracle import oracle as orc
    case, models, setters = get_case(name)
    tr = api.Transit(case["cfg"])
    O = orc.Oracle(case["cfg"])
    ext = tr.extinction_batch(models, total=False)
    for m in range(models.shape[0]):
        o = O.run(models[m], inter=True)
        assert relerr(ext[m], o["ext"]) < 1e-12
    tr.free_memory()


def test_tma_and_plain_staging_agree(api, get_case, monkeypatch):
    """The bulk-async (TMA) staging of the per-model table and the plain cooperative copy must give
    bit-identical spectra."""
    import os
    import subprocess
    import sys
    case, models, _ = get_case("tiny_eclipse")
    tr = api.Transit(case["cfg"])
    ref, _ = tr.run_batch(models)
    tr.free_memory()
    mp = os.path.join(case["workdir"], "m_tma.npy")
    op = os.path.join(case["workdir"], "o_tma.npy")
    np.save(mp, models)
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from bart_b200 import api; "
            "tr = api.Transit(%r); s, st = tr.run_batch(np.load(%r)); np.save(%r, s)"
            % (cases.ROOT, case["cfg"], mp, op))
    env = dict(os.environ, BART_NO_TMA="1")
    subprocess.check_call([sys.executable, "-c", code], env=env)
    assert np.array_equal(np.load(op), ref)


def test_per_model_knobs(api, get_case):
    """bart_set_batch_knobs == calling the reference setters before each single model."""
    case, models, _ = get_case("small4_transit_cloud")
    tr = api.Transit(case["cfg"])
    M = models.shape[0]
    r0 = np.array([94000.0, 95200.0])[:M]
    ct = np.array([-1.0, -2.5])[:M]
    sf = np.array([1, 1], dtype=np.int32)[:M]
    sl = np.array([0.5, 2.0])[:M]
    singles = []
    for m in range(M):
        tr.set_radius(r0[m]); tr.set_cloudtop(ct[m]); tr.set_scattering(1, sl[m])
        singles.append(tr.run_transit(models[m]))
    tr.set_batch_knobs(M, r0, ct, sf, sl)
    batch, status = tr.run_batch(models)
    assert (status == 0).all()
    assert np.array_equal(batch, np.stack(singles))
    tr.free_memory()


def test_rejected_models_are_flagged(api, get_case):
    case, models, _ = get_case("tiny_eclipse")
    tr = api.Transit(case["cfg"])
    batch = np.repeat(models[:1], 4, axis=0)
    batch[1, 5] = 3500.0                                   # T above the opacity grid
    nl = tr.nlayer
    batch[2, nl * 6: nl * 7] = 0.9                         # sum of abundances > 1.001
    spectra, status = tr.run_batch(batch)
    assert status[0] == 0 and status[3] == 0
    assert status[1] & api.REJ_TGRID and status[2] & api.REJ_SUMQ
    assert (spectra[1] == -1).all() and (spectra[2] == -1).all()
    assert np.array_equal(spectra[0], spectra[3])
    with pytest.raises(api.BartError):                     # the reference exit()s; we raise
        tr.run_transit(batch[1])
    tr.free_memory()


def test_modlevel_m1_needs_toomuch(api, workdir):
    """modlevel -1 (slantpath.c:446-473): a column whose optical depth never reaches toomuch makes
    the reference exit (slantpath.c:308-316); here the model is flagged and -1-filled."""
    from bart_b200 import synth
    case = synth.make_case(os.path.join(workdir, "m1_thin"), shape="tiny", solution="transit",
                           seed=12346, refradius_km=95000.0, extra_cfg=["modlevel -1"],
                           overrides={"toomuch": 1e30})
    models = synth.make_models(case, 2, seed=91, molfit=("CH4",))
    tr = api.Transit(case["cfg"])
    spectra, status = tr.run_batch(models)
    assert (status & api.REJ_NOTOOMUCH).all() and (spectra == -1).all()
    with pytest.raises(api.BartError):
        tr.run_transit(models[0])
    tr.free_memory()


def test_band_integration(api, get_case):
    """Stage (c): K4 against wine.resample/bandintegrate restated in numpy, <= 1e-12; and the fused
    profiles -> band flux call against spectra -> band flux."""
    from oracle import oracle as orc
    case, models, _ = get_case("demo_eclipse")
    tr = api.Transit(case["cfg"])
    wn = tr.get_waveno_arr()
    starwn = np.linspace(wn[0] - 10, wn[-1] + 10, 4000)
    hc_k = 6.6260755e-27 * 2.99792458e10 / 1.380658e-16
    starfl = 2 * 6.6260755e-27 * 2.99792458e10 ** 2 * starwn ** 3 / np.expm1(hc_k * starwn / 6075.0) * np.pi
    rprs = 0.12
    start, count, weight, star = api.filters_from_files(wn, case["filters"], starwn, starfl)
    tr.set_filters(start, count, weight, star, rprs)
    spectra, _ = tr.run_batch(models)
    bf = tr.band_integrate(spectra)
    filt = []
    for f in case["filters"]:
        fwn, ftr = orc.readfilter(f)
        filt.append(orc.resample(wn, fwn, ftr, starwn, starfl))
    for m in range(models.shape[0]):
        ref = orc.bandflux(spectra[m], wn, filt, star=True, rprs=rprs)
        assert relerr(bf[m], ref) < 1e-12
    bf2, status = tr.bandflux_batch(models)
    assert np.array_equal(bf2, bf) and (status == 0).all()
    # without a stellar spectrum (transit / direct modes)
    tr.set_filters(start, count, weight, None, 1.0)
    bf3 = tr.band_integrate(spectra)
    for m in range(models.shape[0]):
        ref = orc.bandflux(spectra[m], wn, filt, star=False)
        assert relerr(bf3[m], ref) < 1e-12
    tr.free_memory()


@pytest.mark.parametrize("name", ["w12", "tiny_eclipse", "small4_eclipse_cloud"])
def test_small_batch_kernel_bit_identical(name, api, get_case, workdir, monkeypatch):
    """The slot kernel that small batches select (one warp per 32 columns, kernels.cu
    eclipse_slot_kernel) against the throughput kernel on the same models: identical bits, so a
    spectrum does not depend on the size of the batch it was computed in."""
    import os
    from bart_b200 import synth
    if name == "w12":
        case = synth.make_case(os.path.join(workdir, "w12"), shape="w12", solution="eclipse", seed=2026)
        models = synth.make_models(case, 12, seed=11, molfit=("H2O", "CO2", "CO", "CH4"))
        setters = {}
    else:
        case, models, setters = get_case(name)
    tr = api.Transit(case["cfg"])
    apply_setters(tr, setters)
    monkeypatch.setenv("BART_ECL_SMALL", "0")
    big, st = tr.run_batch(models)
    monkeypatch.setenv("BART_ECL_SMALL", "1")
    small, st2 = tr.run_batch(models)
    assert (st == 0).all() and (st2 == 0).all()
    assert np.array_equal(big, small)
    tr.free_memory()


@pytest.mark.parametrize("name", ["w12", "tiny_eclipse", "small4_eclipse_cloud", "demo_eclipse",
                                  "small4_eclipse_polar", "small4_eclipse_2cia", "tiny_eclipse_3ang",
                                  "tiny_eclipse_9", "tiny_eclipse_t20", "real_inputs_eclipse"])
def test_scan_kernel(name, api, get_case, workdir, monkeypatch):
    """The latency kernel of the smallest batches (lanes <-> layers, kernels.cu eclipse_scan_kernel:
    Simpson panels summed by a warp scan, series / exponentials choice of D(tau) per lane) against
    the throughput kernel on the same models: equal to the accuracy of the two kernels' approximations
    (1e-9; the parity gate is 1e-6), and a model's spectrum does not depend on the batch it came in
    (bit-exact within the kernel)."""
    import os
    from bart_b200 import synth
    if name == "w12":
        case = synth.make_case(os.path.join(workdir, "w12"), shape="w12", solution="eclipse", seed=2026)
        models = synth.make_models(case, 12, seed=11, molfit=("H2O", "CO2", "CO", "CH4"))
        setters = {}
    else:
        case, models, setters = get_case(name)
    tr = api.Transit(case["cfg"])
    apply_setters(tr, setters)
    monkeypatch.setenv("BART_ECL_SMALL", "0")
    big, st = tr.run_batch(models)
    monkeypatch.setenv("BART_ECL_SMALL", "2")
    scan, st2 = tr.run_batch(models)
    one, st3 = tr.run_batch(models[1:2])
    assert (st == 0).all() and (st2 == 0).all() and (st3 == 0).all()
    assert relerr(scan, big) < 1e-9
    assert np.array_equal(one[0], scan[1])
    tr.free_memory()


def test_band_integration_vs_wine_golden(api, get_case):
    """K4 through the C ABI on the arrays the reference's own code/wine.py produced
    (tests/golden/wine.npz: shipped demo filters + Kurucz star on the demo wavenumber grid):
    band fluxes <= 1e-12 of wine.bandintegrate's."""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "wine.npz"))
    case, models, _ = get_case("demo_eclipse")
    tr = api.Transit(case["cfg"])
    wn = tr.get_waveno_arr()
    assert np.array_equal(wn, G["demo_specwn"])
    n = 10
    start = np.array([G["demo_%d_idx" % i][0] for i in range(n)], dtype=np.int32)
    count = np.array([len(G["demo_%d_idx" % i]) for i in range(n)], dtype=np.int32)
    weight = np.concatenate([G["demo_%d_nifilter" % i] for i in range(n)])
    star = np.concatenate([G["demo_%d_istarfl" % i] for i in range(n)])
    tr.set_filters(start, count, weight, star, 0.117)
    bf = tr.band_integrate(G["demo_spectrum"])[0]
    ref = np.array([float(G["demo_%d_band_eclipse" % i]) for i in range(n)])
    assert relerr(bf, ref) < 1e-12
    tr.set_filters(start, count, weight, None, 1.0)
    bf = tr.band_integrate(G["demo_spectrum"])[0]
    ref = np.array([float(G["demo_%d_band_transit" % i]) for i in range(n)])
    assert relerr(bf, ref) < 1e-12
    tr.free_memory()


def test_batch_properties_w12_shape(api, workdir):
    """Full WASP-12b shape (2424 wn x 100 layers x 27 T x 4 molecules): (i) a sample of models
    against the oracle; (ii) permutation equivariance and batch-size independence, bit-exact
    (between the throughput and the slot kernel; the one-model latency kernel to 1e-9);
    (iii) monotone optical depth and last[] consistent with toomuch."""
    import os
    from bart_b200 import synth
    from oracle import oracle as orc
    case = synth.make_case(os.path.join(workdir, "w12"), shape="w12", solution="eclipse", seed=2026)
    models = synth.make_models(case, 48, seed=11, molfit=("H2O", "CO2", "CO", "CH4"))
    tr = api.Transit(case["cfg"])
    assert tr.nwave == 2424 and tr.nlayer == 100
    spectra, status = tr.run_batch(models)
    assert (status == 0).all() and np.isfinite(spectra).all() and (spectra > 0).all()
    O = orc.Oracle(case["cfg"])
    for m in (0, 17, 47):
        assert relerr(spectra[m], O.run(models[m])) < TIGHT
    perm = np.random.default_rng(3).permutation(models.shape[0])
    sp2, _ = tr.run_batch(models[perm])
    assert np.array_equal(sp2, spectra[perm])
    sp3, _ = tr.run_batch(models[4:28])           # 24 models: another batch size, same bits
    assert np.array_equal(sp3, spectra[4:28])
    sp4, _ = tr.run_batch(models[5:6])            # one model: the scan kernel, equal to its approximations
    assert relerr(sp4[0], spectra[5]) < 1e-9
    assert relerr(sp4[0], O.run(models[5])) < TIGHT
    tr.debug_keep(True)
    tr.run_batch(models[:2])
    for m in range(2):
        tau = tr.debug_get("tau", m).reshape(tr.nwave, tr.nlayer)
        last = tr.debug_get("last", m).astype(int)
        for w in range(0, tr.nwave, 97):
            col = tau[w, :last[w] + 1]
            assert (np.diff(col) >= 0).all()
            assert (col[:-1] <= 10.0).all()
            assert col[-1] > 10.0 or last[w] == tr.nlayer - 1
    tr.debug_keep(False)
    # the pipelined host paths (>= 1024 models: chunked H2D / kernels / D2H on three streams) give
    # the bits of plain small calls, with a rejected model inside a chunk, for both entry points
    wn = tr.get_waveno_arr()
    tr.set_filters(*api.filters_from_files(wn, case["filters"], wn, np.ones_like(wn)), 0.1)
    big = np.tile(models, (25, 1))[:1100].copy()
    big[700, :tr.nlayer] = 5000.0                      # outside the opacity grid's temperatures
    sp_big, st_big = tr.run_batch(big)
    bf_big, st_bf = tr.bandflux_batch(big)
    assert st_big[700] != 0 and (np.delete(st_big, 700) == 0).all() and np.array_equal(st_big, st_bf)
    assert (sp_big[700] == -1.0).all() and (bf_big[700] == -1.0).all()
    ok = np.arange(1100) != 700
    assert np.array_equal(sp_big[ok], np.tile(spectra, (25, 1))[:1100][ok])
    bf_small, _ = tr.bandflux_batch(models)
    assert np.array_equal(bf_big[ok], np.tile(bf_small, (25, 1))[:1100][ok])
    tr.free_memory()


def test_reference_on_box_if_present(api, workdir):
    """When oracle/_ref travelled to the box, run the unmodified reference on a fresh seed there."""
    import os
    import conftest
    if not conftest.has_ref():
        pytest.skip("oracle/_ref not present")
    from bart_b200 import synth
    case = synth.make_case(os.path.join(workdir, "fresh_gpu"), shape="tiny", solution="eclipse", seed=60221)
    models = synth.make_models(case, 2, seed=1414)
    mp, op = os.path.join(case["workdir"], "m.npy"), os.path.join(case["workdir"], "ref.npz")
    np.save(mp, models)
    conftest.run_reference(case["cfg"], mp, op, {}, inter=False)
    d = np.load(op)
    tr = api.Transit(case["cfg"])
    spectra, _ = tr.run_batch(models)
    assert relerr(spectra, d["spectra"]) < TOL
    tr.free_memory()


def test_transit_module_drop_in(api, get_case):
    """The CPython module with the reference's SWIG surface, driven the way BARTfunc.py does."""
    import os
    import sys
    from oracle import oracle as orc
    sys.path.insert(0, os.path.join(cases.ROOT, "bart_b200", "python"))
    import transit_module as trm
    case, models, _ = get_case("tiny_eclipse")
    args = ["transit", "-c", case["cfg"]]
    trm.transit_init(len(args), args)
    nwave = trm.get_no_samples()
    specwn = trm.get_waveno_arr(nwave)
    O = orc.Oracle(case["cfg"])
    assert np.array_equal(specwn, O.wn)
    spectrum = trm.run_transit(models[0].flatten(), nwave)
    assert spectrum.shape == (nwave,) and relerr(spectrum, O.run(models[0])) < TOL
    spectra, status = trm.run_transit_batch(models)
    assert np.array_equal(spectra[0], spectrum)
    trm.free_memory()


def test_wavelength_specified_range(api, workdir):
    """Spectral range given as wavelengths (makecfg.py's transit configuration): get_waveno_arr and
    the spectrum against the reference's."""
    from bart_b200 import synth
    g = np.load(cases.golden_path("wl_ranges"))
    for k in range(len(cases.WL_CASES)):
        case = cases.build_wl_case(k, workdir)
        tr = api.Transit(case["cfg"])
        assert np.array_equal(tr.get_waveno_arr(), g["wn%d" % k])
        spec = tr.run_transit(synth.make_models(case, 1, seed=6)[0])
        assert relerr(spec, g["spec%d" % k]) < TOL
        tr.free_memory()


@pytest.mark.parametrize("k", list(cases.FUZZ_GPU))
def test_randomised_configurations(k, api, workdir):
    """The seeded random configurations of tests/cases.py build_fuzz_case (the oracle agrees with
    the compiled reference on every one of them, tests/test_oracle_vs_reference.py) through the CUDA
    path: spectra within the north-star tolerance, last[] identical."""
    from oracle import oracle as orc
    case, models, setters = cases.build_fuzz_case(k, workdir)
    tr = api.Transit(case["cfg"])
    O = orc.Oracle(case["cfg"])
    apply_setters(tr, setters)
    apply_setters(O, setters)
    tr.debug_keep(True)
    spectra, status = tr.run_batch(models)
    assert (status == 0).all()
    for m in range(models.shape[0]):
        o = O.run(models[m], inter=True)
        assert np.array_equal(tr.debug_get("last", m).astype(np.int64), o["last"])
        assert relerr(spectra[m], o["spectrum"]) < TIGHT
    tr.debug_keep(False)
    fast, _ = tr.run_batch(models)
    assert relerr(fast, spectra) < 1e-10          # degree-4 angle exponentials, squared vs evaluated
    tr.free_memory()
