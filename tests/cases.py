"""Named parity configurations shared by tests/ and tests/golden/make_golden.py.

Every case is regenerated bit-identically from its seed by bart_b200.synth (the golden files
carry sha256 digests of the generated opacity grid and model batch to prove it)."""
import hashlib
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bart_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

REF_INPUTS = {k: os.path.join(GOLDEN_DIR, "ref_inputs", v) for k, v in
              (("atm", "HD209458b_demo.atm"), ("mol", "molecules.dat"), ("cia", "CIA_H2H2_400-7000K.dat"))}

# name -> (make_case kwargs, n_models, model seed, setters)
CASES = {
    "tiny_eclipse": (dict(shape="tiny", solution="eclipse", seed=12345), 3, 99, {}),
    "tiny_transit": (dict(shape="tiny", solution="transit", seed=12345, refradius_km=95000.0),
                     3, 98, {"radius": 94000.0}),
    "small4_eclipse_cloud": (dict(shape="small4", solution="eclipse", seed=777), 3, 97,
                             {"cloudtop": -1.0, "scattering": 1.5}),
    "small4_transit_cloud": (dict(shape="small4", solution="transit", seed=778,
                                  refradius_km=95000.0), 2, 96,
                             {"radius": 94500.0, "cloudtop": -2.0, "scattering": 0.5}),
    "demo_eclipse": (dict(shape="demo", solution="eclipse", seed=12345), 2, 95, {}),
    "demo_transit": (dict(shape="demo", solution="transit", seed=12345, refradius_km=95000.0),
                     2, 94, {"radius": 94000.0}),
    "tiny_transit_m1": (dict(shape="tiny", solution="transit", seed=12346, refradius_km=95000.0,
                             extra_cfg=["modlevel -1"]), 3, 91, {"radius": 94200.0}),
    "tiny_eclipse_3ang": (dict(shape="tiny", solution="eclipse", seed=4242,
                               extra_cfg=["raygrid 0 30 70"]), 2, 93, {}),
    # set_scattering(2, .): polarizability (Rayleigh) scattering of every species, extinction.c:586-624
    "small4_eclipse_polar": (dict(shape="small4", solution="eclipse", seed=881, nlayer=40), 2, 90,
                             {"scatflag": 2}),
    # two CIA tables, the second a two-species (H2-He) one with its own temperature range
    # (examples/WASP-12b/BART.cfg: csfile CIA_H2H2..., CIA_H2He...)
    "small4_eclipse_2cia": (dict(shape="small4", solution="eclipse", seed=882, nlayer=40, cia_h2he=True),
                            2, 89, {}),
    # layer counts that are not a multiple of the transit kernel's depth chunk (20), and fewer
    # layers than one chunk
    "tiny_transit_37": (dict(shape="tiny", solution="transit", seed=883, nlayer=37, refradius_km=95000.0),
                        2, 88, {"radius": 94100.0}),
    "tiny_transit_9": (dict(shape="tiny", solution="transit", seed=884, nlayer=9, refradius_km=95000.0),
                       2, 87, {}),
    "tiny_eclipse_9": (dict(shape="tiny", solution="eclipse", seed=885, nlayer=9), 2, 86, {}),
    "tiny_eclipse_t20": (dict(shape="tiny", solution="eclipse", seed=4243,
                              overrides={"toomuch": 20.0}, nlayer=60), 2, 92, {}),
    # the input files the reference ships (tests/golden/ref_inputs/: TEA demo atmosphere, molecules.dat,
    # the H2-H2 CIA table of examples/demo) through the product's readers, demo spectral range
    "real_inputs_eclipse": (dict(shape="demo", solution="eclipse", seed=2024, atm_path=REF_INPUTS["atm"],
                                 mol_path=REF_INPUTS["mol"], cia_path=REF_INPUTS["cia"]), 3, 85, {}),
    "real_inputs_transit": (dict(shape="demo", solution="transit", seed=2025, atm_path=REF_INPUTS["atm"],
                                 mol_path=REF_INPUTS["mol"], cia_path=REF_INPUTS["cia"],
                                 refradius_km=95000.0), 2, 84, {"radius": 94300.0}),
}


# Opacity-grid builder (--justOpacity) cases: name -> make_case kwargs (no pre-made grid)
BUILD_CASES = {
    "build_ch4": dict(shape="tiny", nlayer=12, with_grid=False, nlines=3000, tempdelt=600.0,
                      seed=4711, ethresh=1e-6),
    "build_h2o_ch4": dict(shape=dict(wnlow=2000.0, wnhigh=2120.0, wndelt=1.0, mols=["H2O", "CH4"],
                                     toomuch=10.0),
                          nlayer=10, with_grid=False, nlines=5000, tempdelt=800.0, seed=4712,
                          ethresh=1e-4, wnosamp=1080, nwidth=30),
}


# Line-by-line forward mode (no opacity file: tau.c:163-175,253-264 -> computemolext(permol=0)):
# name -> (make_case kwargs, n_models, model seed, molfit)
LBL_CASES = {
    "lbl_eclipse": (dict(shape="tiny", solution="eclipse", nlayer=24, with_grid=False, no_opacity=True,
                         nlines=3000, seed=5150, ethresh=1e-6), 2, 61, ("CH4",)),
    "lbl_transit_2mol": (dict(shape=dict(wnlow=2000.0, wnhigh=2120.0, wndelt=1.0, mols=["H2O", "CH4"],
                                         toomuch=10.0),
                              solution="transit", refradius_km=95000.0, nlayer=20, with_grid=False,
                              no_opacity=True, nlines=5000, seed=5151, ethresh=1e-4, wnosamp=1080,
                              nwidth=30), 2, 62, ("H2O", "CH4")),
}


def build_lbl_case(name, workdir):
    kw, nm, mseed, molfit = LBL_CASES[name]
    case = synth.make_case(os.path.join(workdir, name), **kw)
    models = synth.make_models(case, nm, seed=mseed, molfit=molfit)
    return case, models


# `savefiles yes` dumps (tau.c:179-190,308-329; tau.dat feeds code/cf.py)
SAVEFILES_CASE = (dict(shape="tiny", solution="eclipse", seed=6001, nlayer=30,
                       extra_cfg=["savefiles yes", "cloudtop -1.0", "scattering 1.5"]), 1, 60)


# what code/cf.py:40-65 does for the contribution functions: the best-fit configuration with
# `toomuch 1e100` (every layer is integrated) and `savefiles yes`, run through the executable
SAVEFILES_CF_CASE = dict(shape="tiny", solution="eclipse", seed=6002, nlayer=30, outputs=True,
                         overrides={"toomuch": 1e100}, extra_cfg=["savefiles yes"])


def build_savefiles_cf_case(workdir):
    return synth.make_case(os.path.join(workdir, "savefiles_cf"), **SAVEFILES_CF_CASE)


def build_savefiles_case(workdir):
    kw, nm, mseed = SAVEFILES_CASE
    case = synth.make_case(os.path.join(workdir, "savefiles"), **kw)
    return case, synth.make_models(case, nm, seed=mseed)


# CLI runs (`transit -c cfg`, transit.c:230-242; BART.py:632-634 best-fit run): name -> make_case kwargs
CLI_CASES = {
    "cli_eclipse": dict(shape="tiny", solution="eclipse", seed=7001, nlayer=40, outputs=True),
    "cli_transit": dict(shape="tiny", solution="transit", seed=7002, nlayer=40, outputs=True,
                        refradius_km=95000.0),
}


def build_cli_case(name, workdir):
    return synth.make_case(os.path.join(workdir, name), **CLI_CASES[name])


# spectral range given as wavelengths (wllow / wlhigh / wlfct: what code/makecfg.py writes from
# BART.cfg), makewnsample makesample.c:282-404
WL_CASES = [(3.0, 5.0, 1.0), (2.9, 4.77, 0.7), (3.3, 3.9, 0.25)]


def build_wl_case(k, workdir):
    lo, hi, d = WL_CASES[k]
    case = synth.make_case(os.path.join(workdir, "wl%d" % k),
                           shape=dict(wnlow=1e4 / hi, wnhigh=1e4 / lo, wndelt=d, mols=["CH4"], toomuch=10.0),
                           solution="eclipse", seed=5, nlayer=20)
    with open(case["cfg"]) as f:
        txt = [l for l in f.read().splitlines() if not l.startswith("wnlow") and not l.startswith("wnhigh")]
    txt += ["wllow %.10g" % lo, "wlhigh %.10g" % hi]
    with open(case["cfg"], "w") as f:
        f.write("\n".join(txt) + "\n")
    return case


def build_fuzz_case(k, workdir):
    """Seeded random configuration k: geometry, layer count, toomuch, ray grid, knobs, spectral
    window -> (case, models, setters)."""
    rng = np.random.default_rng(9000 + k)
    solution = ("eclipse", "transit")[k % 2]
    nlayer = int(rng.integers(12, 70))
    lo = float(rng.uniform(1800.0, 3000.0))
    shape = dict(wnlow=lo, wnhigh=lo + float(rng.uniform(40.0, 160.0)), wndelt=float(rng.choice([0.5, 1.0, 2.0])),
                 mols=[["CH4"], ["H2O", "CO2", "CO", "CH4"]][int(rng.integers(0, 2))],
                 toomuch=float(rng.choice([5.0, 10.0, 20.0, 1e100])))
    extra = []
    if solution == "eclipse":
        extra.append("raygrid " + ["0 20 40 60 80", "0 30 70", "0 15 30 45 60 75", "10 50"][int(rng.integers(0, 4))])
    case = synth.make_case(os.path.join(workdir, "fuzz%d" % k), shape=shape, solution=solution,
                           seed=7000 + k, nlayer=nlayer, extra_cfg=extra,
                           refradius_km=95000.0 if solution == "transit" else 123820.0)
    molfit = ("CH4",) if len(shape["mols"]) == 1 else ("H2O", "CO2", "CO", "CH4")
    models = synth.make_models(case, 2, seed=100 + k, molfit=molfit)
    setters = {}
    if rng.uniform() < 0.5:
        setters["cloudtop"] = float(rng.uniform(-3.0, 0.5))
    if rng.uniform() < 0.5:
        setters["scattering"] = float(rng.uniform(0.0, 2.5))
    if solution == "transit":
        setters["radius"] = float(rng.uniform(93500.0, 96000.0))
    return case, models, setters


def build_builder_fuzz_case(k, workdir):
    """Seeded random builder configuration k (no opacity file on disk): molecules, layers, line
    count, oversampling, profile width, weak-line threshold, spectral step, temperature grid."""
    rng = np.random.default_rng(9500 + k)
    mols = [["CH4"], ["H2O", "CH4"], ["H2O", "CO2", "CO", "CH4"], ["CO", "CH4"]][int(rng.integers(0, 4))]
    lo = float(rng.uniform(1900.0, 2900.0))
    shape = dict(wnlow=lo, wnhigh=lo + float(rng.uniform(60.0, 200.0)), wndelt=float(rng.choice([0.25, 0.5, 1.0, 2.0])),
                 mols=mols, toomuch=10.0)
    kw = dict(shape=shape, nlayer=int(rng.integers(6, 16)), with_grid=False, nlines=int(rng.integers(800, 6000)),
              tempdelt=float(rng.choice([500.0, 650.0, 1300.0])), seed=9600 + k,
              ethresh=float(rng.choice([1e-8, 1e-6, 1e-4, 1e-2])), wnosamp=int(rng.choice([360, 720, 1080, 2160])),
              nwidth=int(rng.choice([10, 20, 40])))
    case = synth.make_case(os.path.join(workdir, "bfuzz%d" % k), **kw)
    if os.path.exists(case["opacity"]):
        os.remove(case["opacity"])
    return case


def build_lbl_fuzz_case(k, workdir):
    """Seeded random line-by-line forward configuration k (no opacity file) -> (case, models)."""
    rng = np.random.default_rng(9700 + k)
    mols = [["CH4"], ["H2O", "CH4"], ["H2O", "CO2", "CO", "CH4"]][int(rng.integers(0, 3))]
    lo = float(rng.uniform(1900.0, 2900.0))
    solution = ("eclipse", "transit")[k % 2]
    shape = dict(wnlow=lo, wnhigh=lo + float(rng.uniform(60.0, 180.0)), wndelt=float(rng.choice([0.5, 1.0, 2.0])),
                 mols=mols, toomuch=float(rng.choice([10.0, 1e100])))
    case = synth.make_case(os.path.join(workdir, "lfuzz%d" % k), shape=shape, solution=solution,
                           nlayer=int(rng.integers(10, 30)), with_grid=False, no_opacity=True,
                           nlines=int(rng.integers(800, 5000)), seed=9800 + k,
                           ethresh=float(rng.choice([1e-8, 1e-5, 1e-3])), wnosamp=int(rng.choice([720, 1080, 2160])),
                           nwidth=int(rng.choice([10, 20, 40])),
                           refradius_km=95000.0 if solution == "transit" else 123820.0)
    molfit = tuple(m for m in ("H2O", "CO2", "CO", "CH4") if m in mols)
    return case, synth.make_models(case, 2, seed=300 + k, molfit=molfit)


FUZZ_LBL = range(4)
FUZZ_BUILDER = range(6)
FUZZ_CPU = range(6)          # against the compiled reference (build container)
FUZZ_GPU = range(12)         # CUDA path against the oracle


def build_builder_case(name, workdir):
    import os as _os
    case = synth.make_case(_os.path.join(workdir, name), **BUILD_CASES[name])
    if _os.path.exists(case["opacity"]):
        _os.remove(case["opacity"])
    return case


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build_case(name, workdir):
    kw, nm, mseed, setters = CASES[name]
    case = synth.make_case(os.path.join(workdir, name), **kw)
    molfit = ("CH4",) if len(case["shape"]["mols"]) == 1 else ("H2O", "CO2", "CO", "CH4")
    models = synth.make_models(case, nm, seed=mseed, molfit=molfit)
    return case, models, setters


def golden_path(name):
    return os.path.join(GOLDEN_DIR, name + ".npz")


# Retrieval-loop cases: input converter + DE-MC around the forward model.
# WASP-12b system constants (examples/WASP-12b/WASP-12b.tep; code/constants.py; BARTfunc.py:159-211)
RSUN, RJUP, MJUP, AU, GNEWT = 6.96e8, 7.1492e7, 1.8983e27, 149597870700.0, 6.6743e-11
W12_PTARGS = (1.57 * RSUN, 6300.0, 100.0, 0.0229 * AU,
              100.0 * GNEWT * 1.41 * MJUP / (1.79 * RJUP) ** 2)    # rstar, tstar, tint, sma, gplanet
RETRIEVAL = {
    # name -> forward-model case, PT model, knob parameters, MC3 set-up
    "retr_tiny_eclipse": dict(
        case=dict(shape="tiny", solution="eclipse", seed=2031), molfit=("CH4",), pt="line",
        nrad=0, ncloud=0, nray=0,
        params=[-0.5, -0.2, 1.0, 0.0, 1.1, 0.5],
        pmin=[-5.0, -3.0, -2.0, 0.0, 0.55, -9.0], pmax=[2.0, 2.0, 3.0, 1.0, 1.4, 3.0],
        stepsize=[0.05, 0.05, 0.0, 0.0, 0.01, 0.3], truth=[-0.7, -0.1, 1.0, 0.0, 1.05, 1.0],
        nchains=6, numit=180, burnin=5, seed=314),
    "retr_small4_transit": dict(
        case=dict(shape="small4", solution="transit", seed=2032, refradius_km=95000.0),
        molfit=("H2O", "CO2", "CO", "CH4"), pt="line", nrad=1, ncloud=1, nray=1,
        params=[-0.5, -0.2, 1.0, 0.0, 1.1, 94000.0, -1.0, 0.5, 0.0, 0.5, -0.5, 0.2],
        pmin=[-5.0, -3.0, -2.0, 0.0, 0.55, 90000.0, -4.0, -3.0, -9.0, -9.0, -9.0, -9.0],
        pmax=[2.0, 2.0, 3.0, 1.0, 1.4, 99000.0, 1.5, 3.0, 3.0, 3.0, 3.0, 3.0],
        stepsize=[0.05, 0.05, 0.0, 0.0, 0.01, 100.0, 0.2, 0.2, 0.3, 0.3, -9, 0.3],
        truth=[-0.6, -0.15, 1.0, 0.0, 1.08, 94300.0, -0.8, 0.8, 0.3, 0.2, 0.3, 0.0],
        nchains=5, numit=100, burnin=0, seed=2718),
}


def build_retrieval(name, workdir):
    """-> (case, spec, dict(pt_args, star, rprs)) for a RETRIEVAL entry."""
    spec = RETRIEVAL[name]
    case = synth.make_case(os.path.join(workdir, name), **spec["case"])
    wn = case["wn"]
    starwn = np.linspace(wn[0] - 10, wn[-1] + 10, 3000)
    hc_k = 6.6260755e-27 * 2.99792458e10 / 1.380658e-16
    starfl = 2 * 6.6260755e-27 * 2.99792458e10 ** 2 * starwn ** 3 / np.expm1(hc_k * starwn / 6300.0) * np.pi
    eclipse = spec["case"]["solution"] == "eclipse"
    extra = dict(pt_args=W12_PTARGS, starwn=starwn if eclipse else None,
                 starfl=starfl if eclipse else None, rprs=0.117 if eclipse else 1.0)
    return case, spec, extra
