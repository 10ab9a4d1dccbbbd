"""CPU: the retrieval-loop oracle (oracle/retrieval_oracle.py) against golden vectors produced by
the reference's own Python code (tests/golden/make_golden_retrieval.py): code/PT.py, the
BARTfunc.py input converter, and a seeded run of the reference's MCcubed.mc.mcmc (DE-MC)."""
import os
import numpy as np
import pytest
import cases
from oracle import retrieval_oracle as ro

G = cases.GOLDEN_DIR


def test_expn2_and_pt_models_vs_reference():
    d = np.load(os.path.join(G, "retrieval_pt.npz"))
    e = ro.expn(2, d["x"])
    ok = d["expn2"] > 0
    assert np.max(np.abs(e[ok] - d["expn2"][ok]) / d["expn2"][ok]) < 1e-15
    assert np.all(e[~ok] == 0)
    p = d["pressure"]
    rstar, tstar, tint, sma, grav = d["ptargs"]
    for q, Tc, Tt in zip(d["pars"], d["T_const"], d["T_thorngren"]):
        assert np.max(np.abs(ro.PT_line(p, *q, rstar, tstar, tint, sma, grav) / Tc - 1)) < 1e-14
        assert np.max(np.abs(ro.PT_line(p, *q, rstar, tstar, tint, sma, grav, "thorngren") / Tt - 1)) < 1e-14
    for q, T in zip(d["apars"], d["T_adiabatic"]):
        assert np.array_equal(ro.PT_adiabatic(p, *q), T)


@pytest.mark.parametrize("name", list(cases.RETRIEVAL))
def test_converter_vs_reference(name, workdir):
    case, spec, extra = cases.build_retrieval(name, workdir)
    d = np.load(os.path.join(G, "retrieval_conv_%s.npz" % name))
    conv = ro.Converter(case["press_bar"], case["species"], case["abund"], spec["molfit"], spec["pt"],
                        pt_args=extra["pt_args"], nrad=spec["nrad"], ncloud=spec["ncloud"],
                        nray=spec["nray"])
    prof, status, knobs = conv.profiles(d["params"])
    assert np.array_equal(status, d["status"])
    assert set(np.unique(status)) == {0, 16, 32}
    ok = status == 0
    assert np.max(np.abs(prof[ok] - d["profiles"][ok]) / np.abs(d["profiles"][ok])) < 1e-14
    c = conv.npt
    if spec["nrad"]:
        assert np.array_equal(knobs["refradius"], d["params"][:, c]); c += 1
    if spec["ncloud"]:
        assert np.array_equal(knobs["cloudtop"], d["params"][:, c]); c += 1
    if spec["nray"]:
        assert np.array_equal(knobs["scat_logext"], d["params"][:, c])


@pytest.mark.parametrize("name", list(cases.RETRIEVAL))
def test_demc_oracle_reproduces_reference_mc3(name, built, workdir):
    """Same seed, same model function -> the reference's chain trace bit for bit."""
    case, spec, extra = cases.build_retrieval(name, workdir)
    d = np.load(os.path.join(G, "retrieval_mc3_%s.npz" % name))
    conv = ro.Converter(case["press_bar"], case["species"], case["abund"], spec["molfit"], spec["pt"],
                        pt_args=extra["pt_args"], nrad=spec["nrad"], ncloud=spec["ncloud"],
                        nray=spec["nray"])
    band = ro.BandOracle(case["cfg"], conv, case["filters"], extra["starwn"], extra["starfl"],
                         extra["rprs"])
    np.random.seed(spec["seed"])
    out = ro.demc(band, d["data"], d["uncert"], spec["params"], spec["pmin"], spec["pmax"],
                  spec["stepsize"], spec["numit"], spec["nchains"], burnin=spec["burnin"])
    assert np.array_equal(out["allparams"], d["allparams"])
    assert np.array_equal(out["bestp"], d["bestp"])
    # MC3's `savemodel` file (BART.cfg: band_eclipse.npy): models of the chains' current states,
    # zeros until a chain's first acceptance (mcmc.py:649-651)
    assert np.array_equal(out["allmodel"], d["allmodel"])
    assert (d["allmodel"][:, :, 0] == 0).all(axis=1).any()
    nacc = out["numaccept"].sum()
    assert 0 < nacc < spec["numit"]


@pytest.mark.parametrize("thinning", [1, 3])
@pytest.mark.parametrize("name", list(cases.RETRIEVAL))
def test_snooker_oracle_reproduces_reference_mc3(name, thinning, built, workdir):
    """walk='snooker' (the walk every BART example configures), MC3's MPI mode: same seed, same
    model function -> the reference's chain trace bit for bit."""
    case, spec, extra = cases.build_retrieval(name, workdir)
    d = np.load(os.path.join(G, "retrieval_snooker_%s_thin%d.npz" % (name, thinning)))
    conv = ro.Converter(case["press_bar"], case["species"], case["abund"], spec["molfit"], spec["pt"],
                        pt_args=extra["pt_args"], nrad=spec["nrad"], ncloud=spec["ncloud"],
                        nray=spec["nray"])
    band = ro.BandOracle(case["cfg"], conv, case["filters"], extra["starwn"], extra["starfl"],
                         extra["rprs"])
    np.random.seed(spec["seed"] + thinning)
    out = ro.snooker(band, d["data"], d["uncert"], spec["params"], spec["pmin"], spec["pmax"],
                     spec["stepsize"], spec["numit"], spec["nchains"], burnin=spec["burnin"],
                     thinning=thinning)
    assert np.array_equal(out["allparams"], d["allparams"])
    assert np.array_equal(out["bestp"], d["bestp"])
    assert np.array_equal(out["allmodel"], d["allmodel"])
    assert out["hsize"] == spec["nchains"] + 1
    assert out["Zsize"] == out["hsize"] + len(range(0, out["allparams"].shape[2], thinning))
    # the walk exercised projected snooker jumps and their Metropolis factor
    assert np.any(out["mrfactor"] != 1.0)
    nacc = out["numaccept"].sum()
    assert 0 < nacc < spec["numit"]
