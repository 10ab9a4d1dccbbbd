"""the reference's worker drives the module unchanged.

tests/golden/bartfunc_trace.npz was recorded by running the reference's UNMODIFIED
code/BARTfunc.py::main(comm) (tests/golden/make_golden_bartfunc.py: stand-in mpi4py master,
recording transit_module with oracle spectra).  Here the recorded calls -- same names, same order,
same argument types and shapes -- are replayed on the real CUDA-backed
bart_b200/python/transit_module, and the band fluxes BARTfunc computed from the returned spectra
(BARTfunc.py:386-396, with the star / filter arrays wine.py derived) must come out the same."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "bartfunc_trace.npz"))
CALLS = json.loads(str(G["calls"]))


def test_recorded_trace_is_the_swig_surface():
    """CPU: what BARTfunc.main called, against the surface of transit/src/transit.i:12-31."""
    names = [c[0] for c in CALLS]
    assert names[:3] == ["transit_init", "get_no_samples", "get_waveno_arr"]
    assert names[-1] == "free_memory" and names.count("run_transit") == 4
    init = CALLS[0][1]
    assert init[0] == {"int": 3} and [list(d.keys())[0] for d in init[1]["list"]] == ["str", "str", "str"]
    run = [c for c in CALLS if c[0] == "run_transit"][0][1]
    assert run[0]["dtype"] == "float64" and len(run[0]["ndarray"]) == 1 and run[0]["contiguous"]
    assert list(run[1].keys()) == ["int"]
    bf = G["bandflux"]
    assert bf.shape == (6, 10) and (bf[3] == -1).all() and (bf[4] == -1).all()      # BARTfunc.py:327-344
    assert (bf[[0, 1, 2, 5]] > 0).all()


@pytest.mark.gpu
def test_replay_on_cuda_transit_module(workdir):
    from bart_b200 import synth
    case = synth.make_case(os.path.join(workdir, "bartfunc_case"), **json.loads(str(G["case"])))
    assert hashlib.sha256(np.fromfile(case["opacity"], dtype=np.uint8).tobytes()).hexdigest() == str(G["grid_sha"])
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bart_b200", "python"))
    import transit_module as trm                      # the module BARTfunc.py:28-30 imports
    specwn, rprs = G["specwn"], float(G["rprs"])
    start, count = G["start"], G["count"]
    off = np.concatenate([[0], np.cumsum(count)])
    ok_rows = [i for i in range(len(G["bandflux"])) if G["bandflux"][i, 0] != -1]
    k = 0
    nwave = None
    for name, args in CALLS:
        if name == "transit_init":
            argv = ["transit", "-c", case["cfg"]]
            trm.transit_init(len(argv), argv)
        elif name == "get_no_samples":
            nwave = trm.get_no_samples()
            assert isinstance(nwave, int) and nwave == len(specwn)
        elif name == "get_waveno_arr":
            wn = trm.get_waveno_arr(nwave)
            assert isinstance(wn, np.ndarray) and wn.dtype == np.float64 and np.array_equal(wn, specwn)
        elif name == "run_transit":
            prof = np.ascontiguousarray(G["profiles"][k])
            assert list(prof.shape) == args[0]["ndarray"]
            spectrum = trm.run_transit(prof, nwave)
            assert isinstance(spectrum, np.ndarray) and spectrum.shape == (nwave,)
            band = np.zeros(len(start))
            for i in range(len(start)):                # BARTfunc.py:386-391 + wine.bandintegrate
                idx = np.arange(start[i], start[i] + count[i])
                fluxrat = spectrum[idx] / G["star"][off[i]:off[i + 1]] * rprs * rprs
                y = fluxrat * G["weight"][off[i]:off[i + 1]]
                band[i] = np.sum(np.diff(specwn[idx]) * (y[1:] + y[:-1]) / 2.0)
            ref = G["bandflux"][ok_rows[k]]
            assert np.max(np.abs(band / ref - 1)) < 1e-8, (k, band, ref)
            k += 1
        elif name == "free_memory":
            # before letting go: the additive batched entry gives the same band fluxes in one call
            trm.set_filters(start.astype(np.int32), count.astype(np.int32), G["weight"], G["star"], rprs)
            bf, st = trm.band_flux_batch(np.ascontiguousarray(G["profiles"]), len(start))
            assert (st == 0).all()
            assert np.max(np.abs(bf / G["bandflux"][ok_rows] - 1)) < 1e-8
            trm.free_memory()
        else:
            raise AssertionError("the worker called %s, which the replay does not know" % name)
    assert k == 4


@pytest.mark.gpu
def test_device_worker_equals_unmodified_bartfunc(workdir):
    """The whole worker iteration on the device -- parameters -> PT profile, abundance scaling, the
    two rejections, forward model, band integration (bart_bandflux_from_params) -- set up the way
    BARTfunc.py:139-222 sets itself up (TEP file, atmosphere, molfit, filters), against the band
    fluxes the UNMODIFIED BARTfunc.main sent back to MC3 for the same parameter vectors."""
    from bart_b200 import api, driver, synth
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_bartfunc as gen                 # the generator's constants (TEP text, molfit)
    case = synth.make_case(os.path.join(workdir, "bartfunc_case_dev"), **json.loads(str(G["case"])))
    tepfile = os.path.join(workdir, "planet.tep")
    open(tepfile, "w").write(gen.TEP)
    sysp = driver.system_from_tep(tepfile, tint=100.0)
    assert abs(sysp["rprs"] / float(G["rprs"]) - 1) < 1e-15
    tr = api.Transit(case["cfg"])
    tr.set_filters(G["start"].astype(np.int32), G["count"].astype(np.int32), G["weight"], G["star"],
                   sysp["rprs"])
    species, press, _, abund = driver.read_atm(case["atm"])          # BARTfunc.py:185 (mat.readatm)
    assert species == case["species"] and np.allclose(press, case["press_bar"], rtol=1e-4)
    npars = tr.converter_init(press, species, abund, gen.MOLFIT, "line",
                              pt_args=sysp["pt_args"], tmin=400.0, tmax=3000.0)   # BARTfunc's defaults
    assert npars == G["params"].shape[1]
    bf, status = tr.bandflux_from_params(G["params"])
    ref = G["bandflux"]
    rej = ref[:, 0] == -1
    assert list(status[rej]) == [16, 32] and (bf[rej] == -1).all()    # BARTfunc.py:327-330, 339-344
    assert (status[~rej] == 0).all()
    assert np.max(np.abs(bf[~rej] / ref[~rej] - 1)) < 1e-8
    tr.free_memory()


@pytest.mark.gpu
def test_bartworker_from_the_mcmc_configuration(workdir):
    """driver.BartWorker reads the [MCMC] section BARTfunc reads (the one the golden run used, minus
    the Kurucz file, which does not travel: the stellar spectrum wine.readkurucz returned is passed
    in) and returns what the unmodified BARTfunc gathered."""
    from bart_b200 import driver, synth
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_bartfunc as gen
    case = synth.make_case(os.path.join(workdir, "bartfunc_case_cfg"), **json.loads(str(G["case"])))
    tepfile = os.path.join(workdir, "planet2.tep")
    open(tepfile, "w").write(gen.TEP)
    cfg = os.path.join(workdir, "MCMC.cfg")
    with open(cfg, "w") as f:                              # as make_golden_bartfunc.main writes it
        f.write("[MCMC]\n")
        f.write("params = %s\n" % " ".join("%r" % v for v in gen.PARAMS[0]))
        f.write("molfit = %s\n" % " ".join(gen.MOLFIT))
        f.write("atmfile = %s\nPTtype = line\ntint = 100.0\ntint_type = const\n" % case["atm"])
        f.write("tconfig = %s\n" % case["cfg"])
        f.write("filters = %s\n" % "\n    ".join(case["filters"]))
        f.write("tep_name = %s\nsolution = eclipse\n" % tepfile)
    # the star on the spectrum grid, as wine.resample left it in the fixture: rebuild (wn, flux) pairs
    # that interpolate to exactly those values on every filter's sample range
    wn = G["specwn"]
    starfl = np.interp(wn, wn, np.ones_like(wn))
    off = np.concatenate([[0], np.cumsum(G["count"])])
    for i in range(len(G["start"])):
        starfl[G["start"][i]:G["start"][i] + G["count"][i]] = G["star"][off[i]:off[i + 1]]
    w = driver.BartWorker(cfg, star=(wn, starfl))
    assert w.npars == G["params"].shape[1] and w.tr.nfilters == len(G["start"])
    bf = w(G["params"])
    ref = G["bandflux"]
    rej = ref[:, 0] == -1
    assert (bf[rej] == -1).all()
    assert np.max(np.abs(bf[~rej] / ref[~rej] - 1)) < 1e-8
    w.close()
