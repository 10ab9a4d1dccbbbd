"""GPU: the retrieval loop on the device (include/bart_b200.h part 3) through the C ABI, against
the retrieval oracle and the golden vectors made by the reference's own Python code."""
import os
import numpy as np
import pytest
import cases
from util import relerr

pytestmark = pytest.mark.gpu
G = cases.GOLDEN_DIR


@pytest.fixture(scope="module")
def api():
    from bart_b200 import api as a
    a.lib()
    return a


def setup(api, name, workdir):
    case, spec, extra = cases.build_retrieval(name, workdir)
    tr = api.Transit(case["cfg"])
    wn = tr.get_waveno_arr()
    start, count, weight, star = api.filters_from_files(wn, case["filters"], extra["starwn"], extra["starfl"])
    tr.set_filters(start, count, weight, star, extra["rprs"])
    tr.converter_init(case["press_bar"], case["species"], case["abund"], spec["molfit"], spec["pt"],
                      pt_args=extra["pt_args"], nrad=spec["nrad"], ncloud=spec["ncloud"],
                      nray=spec["nray"])
    return case, spec, extra, tr


def band_oracle(case, spec, extra):
    from oracle import retrieval_oracle as ro
    conv = ro.Converter(case["press_bar"], case["species"], case["abund"], spec["molfit"], spec["pt"],
                        pt_args=extra["pt_args"], nrad=spec["nrad"], ncloud=spec["ncloud"],
                        nray=spec["nray"])
    return ro.BandOracle(case["cfg"], conv, case["filters"], extra["starwn"], extra["starfl"],
                         extra["rprs"])


@pytest.mark.parametrize("name", list(cases.RETRIEVAL))
def test_converter_vs_reference_golden(name, api, workdir):
    """K-1 against the profiles the reference's BARTfunc statements + PT.py produce."""
    case, spec, extra, tr = setup(api, name, workdir)
    d = np.load(os.path.join(G, "retrieval_conv_%s.npz" % name))
    prof, status, knobs = tr.profiles_from_params(d["params"])
    assert np.array_equal(status, d["status"])
    ok = status == 0
    assert ok.sum() >= 10 and (status == 16).any() and (status == 32).any()
    assert np.max(np.abs(prof[ok] - d["profiles"][ok]) / np.abs(d["profiles"][ok])) < 1e-12
    c = 5
    for k, on in enumerate((spec["nrad"], spec["ncloud"], spec["nray"])):
        if on:
            assert np.array_equal(knobs[k], d["params"][:, c]); c += 1
    tr.free_memory()


def test_pt_line_thorngren_and_adiabatic(api, workdir):
    case, spec, extra, tr = setup(api, "retr_tiny_eclipse", workdir)
    d = np.load(os.path.join(G, "retrieval_pt.npz"))
    assert np.array_equal(d["pressure"], case["press_bar"]) or np.allclose(d["pressure"], case["press_bar"], rtol=1e-14)
    nl = tr.nlayer
    for tint_type, key in (("const", "T_const"), ("thorngren", "T_thorngren")):
        tr.converter_init(case["press_bar"], case["species"], case["abund"], (), "line",
                          pt_args=extra["pt_args"], tint_type=tint_type, tmin=0.0, tmax=1e9, nrad=0)
        prof, status, _ = tr.profiles_from_params(d["pars"])
        assert (status == 0).all()
        assert np.max(np.abs(prof[:, :nl] / d[key] - 1)) < 1e-12
    tr.converter_init(case["press_bar"], case["species"], case["abund"], (), "adiabatic", tmin=-1e9,
                      tmax=1e9, nrad=0)
    prof, status, _ = tr.profiles_from_params(d["apars"])
    assert np.max(np.abs(prof[:, :nl] / d["T_adiabatic"] - 1)) < 1e-13
    tr.converter_init(case["press_bar"], case["species"], case["abund"], (), "iso", nrad=0)
    prof, status, _ = tr.profiles_from_params(np.array([[1500.0], [399.0]]))
    assert np.all(prof[0, :nl] == 1500.0) and list(status) == [0, 16]
    tr.free_memory()


def test_smoothed_pt_models_on_device(api, workdir):
    """Madhusudhan-Seager (inverted / non-inverted) and Piette profiles incl. the Gaussian smoothing
    over the layers, on the device, against code/PT.py (golden) and the oracle; parameter sets
    PT.py refuses come back as BART_REJ_PTMODEL."""
    from oracle import retrieval_oracle as ro
    case, spec, extra, tr = setup(api, "retr_tiny_eclipse", workdir)
    g = np.load(os.path.join(G, "retrieval_pt_smooth.npz"))
    assert np.allclose(g["pressure_a"], case["press_bar"], rtol=1e-14)
    nl = tr.nlayer
    for name, key in (("madhu_noinv", "noinv"), ("madhu_inv", "inv"), ("piette", "piette")):
        tr.converter_init(case["press_bar"], case["species"], case["abund"], (), name, tmin=0.0,
                          tmax=1e9, nrad=0)
        pars, T, phys = (g["%s_%s_a" % (key, k)] for k in ("pars", "T", "phys"))
        prof, status, _ = tr.profiles_from_params(pars)
        assert np.array_equal(status == 256, phys == 0) and np.all(status[phys == 1] == 0)
        ok = phys == 1
        assert np.max(np.abs(prof[ok, :nl] / T[ok] - 1)) < 1e-13
        conv = ro.Converter(case["press_bar"], case["species"], case["abund"], (), name, tmin=0.0, tmax=1e9)
        oprof, ostatus, _ = conv.profiles(pars)
        assert np.array_equal(ostatus, status)
        assert np.max(np.abs(prof[ok] / oprof[ok] - 1)) < 1e-13
    # a second pressure grid (60 layers, 0.144 dex per layer: Piette's sigma 2.08 layers)
    import os as _os
    from bart_b200 import synth
    c60 = synth.make_case(_os.path.join(workdir, "pt60"), shape="tiny", solution="eclipse", seed=5, nlayer=60)
    tr.free_memory()
    tr60 = api.Transit(c60["cfg"])
    pb = g["pressure_b"]
    for name, key in (("madhu_noinv", "noinv"), ("madhu_inv", "inv"), ("piette", "piette")):
        tr60.converter_init(pb, c60["species"], c60["abund"], (), name, tmin=0.0, tmax=1e9, nrad=0)
        pars, T, phys = (g["%s_%s_b" % (key, k)] for k in ("pars", "T", "phys"))
        prof, status, _ = tr60.profiles_from_params(pars)
        assert np.array_equal(status == 256, phys == 0)
        ok = phys == 1
        assert np.max(np.abs(prof[ok, :60] / T[ok] - 1)) < 1e-13
    tr60.free_memory()
    tr = api.Transit(case["cfg"])
    # temperature bounds act on the smoothed profile (BARTfunc.py:327-330)
    tr.converter_init(case["press_bar"], case["species"], case["abund"], (), "piette", tmin=400.0,
                      tmax=3000.0, nrad=0)
    pars, T, phys = (g["piette_%s_a" % k] for k in ("pars", "T", "phys"))
    prof, status, _ = tr.profiles_from_params(pars)
    want = np.where((T < 400.0).any(axis=1) | (T > 3000.0).any(axis=1), 16, 0)
    assert np.array_equal(status, want) and (want == 0).any() and (want == 16).any()
    tr.free_memory()


@pytest.mark.parametrize("name", list(cases.RETRIEVAL))
def test_bandflux_from_params_vs_oracle(name, api, workdir):
    """parameters -> band fluxes in one call (one BARTfunc worker iteration) against the oracle
    chain converter -> forward model -> band integration; rejected proposals are -1."""
    case, spec, extra, tr = setup(api, name, workdir)
    d = np.load(os.path.join(G, "retrieval_conv_%s.npz" % name))
    params = d["params"]
    bf, status = tr.bandflux_from_params(params)
    ref = band_oracle(case, spec, extra)(params)
    rej = d["status"] != 0
    assert np.all(bf[rej] == -1.0) and np.all(ref[rej] == -1.0)
    assert np.array_equal(status[rej], d["status"][rej])
    ok = ~rej & (status == 0)
    assert ok.sum() >= 8
    assert relerr(bf[ok], ref[ok]) < 1e-6
    # the same through the host-converter path (profiles -> band fluxes)
    prof, st2, knobs = tr.profiles_from_params(params[ok])
    kw = {}
    if spec["nrad"]: kw["refradius"] = knobs[0]
    if spec["ncloud"]: kw["cloudtop"] = knobs[1]
    if spec["nray"]: kw["scat_flag"] = np.ones(ok.sum(), dtype=np.int32); kw["scat_logext"] = knobs[2]
    if kw:
        tr.set_batch_knobs(int(ok.sum()), **kw)
    bf2, _ = tr.bandflux_batch(prof)
    tr.set_batch_knobs(0)
    assert np.array_equal(bf2, bf[ok])
    tr.free_memory()


@pytest.mark.parametrize("graph", ["1", "0"])
@pytest.mark.parametrize("name", list(cases.RETRIEVAL))
def test_demc_on_device_reproduces_reference_mc3(name, graph, api, workdir, monkeypatch):
    """The seeded run of the reference's MCcubed.mc.mcmc (golden) chain for chain: same random
    streams, the generation loop on the GPU (CUDA-graph replay and plain launches)."""
    from bart_b200 import driver
    monkeypatch.setenv("BART_MCMC_GRAPH", graph)
    case, spec, extra, tr = setup(api, name, workdir)
    d = np.load(os.path.join(G, "retrieval_mc3_%s.npz" % name))
    np.random.seed(spec["seed"])
    n0 = api.lib().bart_launch_count()
    out = driver.run_demc(tr, d["data"], d["uncert"], spec["params"], spec["pmin"], spec["pmax"],
                          spec["stepsize"], spec["numit"], spec["nchains"], burnin=spec["burnin"])
    assert np.array_equal(out["allparams"], d["allparams"])          # every chain, every iteration
    assert np.array_equal(out["allstack"], d["allstack"])
    assert np.array_equal(out["bestp"], d["bestp"])
    # MC3's savemodel array (mcmc.py:636-651, 849-850) and its files
    assert np.array_equal(out["allmodel"] == 0, d["allmodel"] == 0)
    assert relerr(out["allmodel"], d["allmodel"]) < 1e-6
    sf, sm = os.path.join(workdir, "output_%s.npy" % name), os.path.join(workdir, "band_%s.npy" % name)
    driver._save_mc3_files(out, sf, sm)
    assert np.array_equal(np.load(sf), d["allparams"]) and np.load(sm).shape == d["allmodel"].shape
    chainsize = d["allparams"].shape[2]
    per_gen = 5 if tr.eclipse else 7
    assert api.lib().bart_launch_count() - n0 >= per_gen * chainsize
    # the oracle loop on the same streams: acceptance counts, chi-squared, best model
    from oracle import retrieval_oracle as ro
    np.random.seed(spec["seed"])
    ref = ro.demc(band_oracle(case, spec, extra), d["data"], d["uncert"], spec["params"], spec["pmin"],
                  spec["pmax"], spec["stepsize"], spec["numit"], spec["nchains"], burnin=spec["burnin"])
    assert np.array_equal(out["numaccept"], ref["numaccept"])
    assert np.array_equal(out["outbounds"], ref["outbounds"])
    assert np.array_equal(out["params"], ref["params"])
    assert relerr(out["currchisq"], ref["currchisq"]) < 1e-6
    assert abs(out["bestchisq"] / ref["bestchisq"] - 1) < 1e-6
    assert relerr(out["bestmodel"], ref["bestmodel"]) < 1e-6
    assert relerr(out["models"], ref["allmodels"][-1]) < 1e-6
    # MC3's resume=True (mcmc.py:254-269) on the files just written, against the reference's own
    # resumed run: chains restart from their last states, traces are appended
    np.save(sm, out["allmodel"])
    np.random.seed(spec["seed"] + 1)
    res = driver.run_demc(tr, d["data"], d["uncert"], spec["params"], spec["pmin"], spec["pmax"],
                          spec["stepsize"], spec["numit"], spec["nchains"], burnin=spec["burnin"],
                          savefile=sf, savemodel=sm, resume=True)
    assert res["allparams"].shape[2] == 2 * chainsize
    assert np.array_equal(res["allparams"], d["resume_allparams"])
    assert np.array_equal(res["allstack"], d["resume_allstack"])
    assert np.array_equal(res["bestp"], d["resume_bestp"])
    assert np.array_equal(res["allmodel"] == 0, d["resume_allmodel"] == 0)
    assert relerr(res["allmodel"], d["resume_allmodel"]) < 1e-6
    assert np.array_equal(np.load(sf), d["resume_allparams"])
    tr.free_memory()


def test_demc_continues_across_calls(api, workdir):
    """Two bart_mcmc_run calls of n/2 generations == one call of n (state stays on the device)."""
    from bart_b200 import driver
    name = "retr_tiny_eclipse"
    case, spec, extra, tr = setup(api, name, workdir)
    d = np.load(os.path.join(G, "retrieval_mc3_%s.npz" % name))
    nch = spec["nchains"]
    stepsize = np.array(spec["stepsize"])
    rng = np.random.RandomState(5)
    p0 = np.repeat(np.atleast_2d(spec["params"]), nch, 0)
    p0[:, stepsize > 0] += rng.normal(0, 0.01, (nch, int((stepsize > 0).sum())))
    dr = driver.demc_draws(rng, nch, 16, stepsize[stepsize > 0])
    one = driver.run_demc(tr, d["data"], d["uncert"], p0, spec["pmin"], spec["pmax"], stepsize,
                          16 * nch, nch, draws=dr)
    tr.mcmc_init(p0, spec["pmin"], spec["pmax"], stepsize, d["data"], d["uncert"])
    halves = []
    for h in (slice(0, 8), slice(8, 16)):
        tr.mcmc_run(dr["support"][h], dr["r1"][:, h], dr["r2"][:, h], dr["unif"][h], dr["ugamma"][h])
        halves.append(tr.mcmc_get("allparams"))
    assert np.array_equal(np.concatenate(halves, axis=2), one["allparams"])
    assert np.array_equal(tr.mcmc_get("params"), one["params"])
    tr.free_memory()


@pytest.mark.parametrize("graph", ["1", "0"])
@pytest.mark.parametrize("thinning", [1, 3])
@pytest.mark.parametrize("name", list(cases.RETRIEVAL))
def test_snooker_on_device_reproduces_reference_mc3(name, thinning, graph, api, workdir, monkeypatch):
    """walk='snooker' (BART's configured walk): the seeded reference MCcubed run (MPI mode) chain
    for chain, with history Z, proposals, Metropolis rule and forward models on the GPU."""
    from bart_b200 import driver
    monkeypatch.setenv("BART_MCMC_GRAPH", graph)
    case, spec, extra, tr = setup(api, name, workdir)
    d = np.load(os.path.join(G, "retrieval_snooker_%s_thin%d.npz" % (name, thinning)))
    np.random.seed(spec["seed"] + thinning)
    out = driver.run_snooker(tr, d["data"], d["uncert"], spec["params"], spec["pmin"], spec["pmax"],
                             spec["stepsize"], spec["numit"], spec["nchains"], burnin=spec["burnin"],
                             thinning=thinning)
    assert np.array_equal(out["allparams"], d["allparams"])          # every chain, every iteration
    assert np.array_equal(out["allstack"], d["allstack"])
    assert np.array_equal(out["bestp"], d["bestp"])
    assert np.array_equal(out["allmodel"] == 0, d["allmodel"] == 0)
    assert relerr(out["allmodel"], d["allmodel"]) < 1e-6
    from oracle import retrieval_oracle as ro
    np.random.seed(spec["seed"] + thinning)
    ref = ro.snooker(band_oracle(case, spec, extra), d["data"], d["uncert"], spec["params"],
                     spec["pmin"], spec["pmax"], spec["stepsize"], spec["numit"], spec["nchains"],
                     burnin=spec["burnin"], thinning=thinning)
    assert np.array_equal(out["numaccept"], ref["numaccept"])
    assert np.array_equal(out["outbounds"], ref["outbounds"])
    assert np.array_equal(out["params"], ref["params"])
    assert out["Z"].shape[0] == ref["Zsize"]
    assert np.array_equal(out["Z"], ref["Z"][:ref["Zsize"]])
    assert relerr(out["Zchisq"], ref["Zchisq"][:ref["Zsize"]]) < 1e-6
    assert relerr(out["currchisq"], ref["currchisq"]) < 1e-6
    assert abs(out["bestchisq"] / ref["bestchisq"] - 1) < 1e-6
    assert relerr(out["bestmodel"], ref["bestmodel"]) < 1e-6
    tr.free_memory()


def test_snooker_large_population_and_continuation(api, workdir):
    """64 chains x 11 free parameters (numpy's 8-way pairwise sums in the projection), split over
    two bart_mcmc_run_snooker calls, against the oracle on the same streams."""
    from bart_b200 import driver
    from oracle import retrieval_oracle as ro
    name = "retr_small4_transit"
    case, spec, extra, tr = setup(api, name, workdir)
    d = np.load(os.path.join(G, "retrieval_mc3_%s.npz" % name))
    nch, niter, thin = 64, 12, 2
    stepsize = np.array(spec["stepsize"], dtype=float)
    stepsize[2], stepsize[3] = 0.05, 0.02                           # free all but the shared one
    pmin, pmax = np.array(spec["pmin"]), np.array(spec["pmax"])
    ifree = np.where(stepsize > 0)[0]
    assert len(ifree) == 11
    rng = np.random.RandomState(11)
    p0 = np.repeat(np.atleast_2d(spec["params"]), nch, 0)
    p0[:, ifree] += rng.normal(0, 0.2, (nch, len(ifree))) * stepsize[ifree]
    hsize = nch + 1
    dr = driver.snooker_draws(rng, nch, len(ifree), niter, hsize, thin, stepsize[ifree], pmin[ifree], pmax[ifree])
    dr["ugamma"][:, ::3] *= 0.2                                      # more snooker jumps than 10 %
    sj = dr["ugamma"] < 0.1
    dr["usn_offset"] = np.concatenate([[0], np.cumsum(sj.sum(axis=1))])
    dr["usnooker"] = rng.uniform(1.2, 2.2, (dr["usn_offset"][-1], len(ifree)))
    ref = ro.snooker(band_oracle(case, spec, extra), d["data"], d["uncert"], p0, pmin, pmax, stepsize,
                     niter * nch, nch, thinning=thin, draws=dr)
    tr.mcmc_init(p0, pmin, pmax, stepsize, d["data"], d["uncert"])
    tr.mcmc_snooker_init(dr["z0"], thin)
    halves = []
    for h in (slice(0, 6), slice(6, 12)):
        off = dr["usn_offset"][h.start:h.stop + 1]
        tr.mcmc_run_snooker(dr["support"][h], dr["i1"][h], dr["i2"][h], dr["iz"][h], dr["ic"][h],
                            dr["usnooker"][off[0]:off[-1]], off - off[0], dr["unif"][h], dr["ugamma"][h])
        halves.append(tr.mcmc_get("allparams"))
    got = np.concatenate(halves, axis=2)
    assert np.array_equal(got, ref["allparams"])
    assert np.array_equal(tr.mcmc_get("Z"), ref["Z"][:ref["Zsize"]])
    assert np.any(ref["mrfactor"] != 1.0)
    tr.free_memory()


@pytest.mark.parametrize("walk", ["demc", "snooker"])
def test_grtest_segments_leave_the_chains_unchanged(walk, api, workdir):
    """grtest=True runs the device loop in the segments between MC3's convergence checkpoints
    (mcmc.py:238,663-686): same chains as the single-call run / the reference, PSRF history equal to
    the reference's gelman_rubin at those checkpoints.  Snooker with thinning 3 and checkpoints every
    2 generations exercises the global iteration numbering of the history update."""
    from bart_b200 import driver
    name = "retr_small4_transit"                                   # chainsize 20 -> checkpoints 1,3,5..
    case, spec, extra, tr = setup(api, name, workdir)
    if walk == "demc":
        d = np.load(os.path.join(G, "retrieval_mc3_%s.npz" % name))
        np.random.seed(spec["seed"])
        out = driver.run_demc(tr, d["data"], d["uncert"], spec["params"], spec["pmin"], spec["pmax"],
                              spec["stepsize"], spec["numit"], spec["nchains"], burnin=spec["burnin"],
                              grtest=True)
        g = np.load(os.path.join(G, "retrieval_gr.npz"))
        assert [h[0] for h in out["psrf"]] == list(g["its_" + name])
        for (i, psrf), ref in zip(out["psrf"], g["psrf_" + name]):
            assert np.allclose(psrf, ref, rtol=1e-12, equal_nan=True)
    else:
        d = np.load(os.path.join(G, "retrieval_snooker_%s_thin3.npz" % name))
        np.random.seed(spec["seed"] + 3)
        out = driver.run_snooker(tr, d["data"], d["uncert"], spec["params"], spec["pmin"], spec["pmax"],
                                 spec["stepsize"], spec["numit"], spec["nchains"], burnin=spec["burnin"],
                                 thinning=3, grtest=True)
        assert len(out["psrf"]) >= 5
    assert np.array_equal(out["allparams"], d["allparams"])
    assert np.array_equal(out["allstack"], d["allstack"])
    assert np.array_equal(out["bestp"], d["bestp"])
    assert np.array_equal(out["allmodel"] == 0, d["allmodel"] == 0)
    assert relerr(out["allmodel"], d["allmodel"]) < 1e-6
    tr.free_memory()


@pytest.mark.gpu
def test_energy_balance_rejection_vs_bartfunc_golden(workdir):
    """The device energy-balance test (retrieval.cu energy_balance_kernel through
    bart_energy_balance / bart_set_energy_balance) against code/BARTfunc.py:366-383 run verbatim
    (tests/golden/ebalance.npz: 16 spectra on the WASP-12b wavenumber grid straddling the
    threshold), and end to end: a rejected model returns -1 band fluxes and BART_REJ_ENERGY."""
    import os
    from bart_b200 import api, synth
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ebalance.npz"))
    case = synth.make_case(os.path.join(workdir, "w12_eb"), shape="w12", solution="eclipse", seed=2026)
    tr = api.Transit(case["cfg"])
    wn = tr.get_waveno_arr()
    assert np.array_equal(wn, g["specwn"])
    e_in = tr.set_energy_balance(float(g["tstar"]), float(g["rstar"]), float(g["sma"]), float(g["rplanet"]))
    assert abs(e_in / g["e_in"][0] - 1) < 1e-14
    flags = tr.energy_balance(g["spectra"])
    assert np.array_equal(flags != 0, g["rejected"])
    assert set(np.unique(flags)) <= {0, api.REJ_ENERGY}
    # end to end through the band-flux call: the same models with the test off, with a budget no
    # spectrum can exceed, and with one every spectrum exceeds
    start, count, weight, star = api.filters_from_files(wn, case["filters"], wn, np.ones_like(wn))
    tr.set_filters(start, count, weight, star, 0.1)
    models = synth.make_models(case, 6, seed=3, molfit=("H2O", "CO2", "CO", "CH4"))
    api._check(api.lib().bart_set_energy_balance(0, 1.0, 1.0))
    bf0, st0 = tr.bandflux_batch(models)
    assert (st0 == 0).all() and (bf0 > 0).all()
    api._check(api.lib().bart_set_energy_balance(1, 1e300, 1.0))
    bf1, st1 = tr.bandflux_batch(models)
    assert np.array_equal(bf1, bf0) and (st1 == 0).all()
    api._check(api.lib().bart_set_energy_balance(1, 1e-300, 1.0))
    bf2, st2 = tr.bandflux_batch(models)
    assert (st2 == api.REJ_ENERGY).all() and (bf2 == -1).all()
    tr.free_memory()
