#!/usr/bin/env python
"""Generate tests/golden/retrieval_*.npz from the REFERENCE's own Python code, in the build
container (neither /root/reference nor the /tmp build exist on the GPU box):

  * retrieval_pt.npz      code/PT.py PT_line / PT_iso / PT_adiabatic (imported from
                          /root/reference/code) on seeded parameter draws, incl. scipy's expn(2, x)
  * retrieval_pt_smooth.npz  code/PT.py PT_NoInversion / PT_Inversion / PT_piette (the models that
                          smooth over the layers with scipy's gaussian_filter1d)
  * retrieval_conv_*.npz  the per-proposal input converter: the statements of
                          code/BARTfunc.py:320-347 executed verbatim in numpy around the
                          reference's PT_generator
  * retrieval_mc3_*.npz   a seeded run of the reference's MCcubed.mc.mcmc (walk='demc'; C
                          extensions compiled from a copy under /tmp/bart_mc3_ref with the
                          reference's own setup.py) whose model function is the forward-model
                          oracle: the chain trace, best fit, acceptance counts
  * retrieval_snooker_*.npz  the same with walk='snooker' (BART's configured walk) in MC3's MPI
                          mode, the worker side played by an in-process stand-in (FakeWorkers)

Shims applied from the outside, none to the reference sources: `numpy.int/float` aliases (removed
from numpy 2), stub `matplotlib` and `dwt` modules (absent / built on a removed numpy C API; neither
is used on this path).

    python tests/golden/make_golden_retrieval.py [pt] [ptsmooth] [tep] [conv] [demc] [snooker]
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile
import types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from oracle import retrieval_oracle as ro  # noqa: E402


def shim():
    np.int = int
    np.float = float
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "dwt"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    sys.modules["matplotlib"].use = lambda *a, **k: None


def build_mc3():
    dst = "/tmp/bart_mc3_ref"
    if not glob.glob(os.path.join(dst, "MCcubed", "lib", "chisq*.so")):
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree(os.path.join(REF, "modules", "MCcubed"), dst)
        subprocess.check_call([sys.executable, "setup.py", "build"], cwd=dst,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for so in glob.glob(os.path.join(dst, "build", "lib.*", "*.so")):
            if "dwt" not in so:
                shutil.copy(so, os.path.join(dst, "MCcubed", "lib"))
    sys.path.insert(0, dst)
    import MCcubed
    return MCcubed


def golden_pt(pt):
    rng = np.random.default_rng(12345)
    p = np.logspace(2, -5, 100)
    rstar, tstar, tint, sma, grav = cases.W12_PTARGS
    lo = np.array([-5.0, -3.0, -2.0, 0.0, 0.55])
    hi = np.array([2.0, 2.0, 3.0, 1.0, 1.4])
    pars = rng.uniform(lo, hi, (64, 5))
    pars[0] = [-0.5, -0.2, 1.0, 0.0, 1.1]                    # examples/WASP-12b/BART.cfg:104
    T_const = np.array([pt.PT_line(p, *q, rstar, tstar, tint, sma, grav, "const") for q in pars])
    T_thorn = np.array([pt.PT_line(p, *q, rstar, tstar, tint, sma, grav, "thorngren") for q in pars])
    apars = np.column_stack([rng.uniform(800, 2500, 16), rng.uniform(1.1, 1.67, 16),
                             rng.uniform(-1, 2, 16)])
    T_adia = np.array([pt.PT_adiabatic(p, *q) for q in apars])
    import scipy.special as sp
    x = np.concatenate([10 ** rng.uniform(-10, 2.85, 400), [1.0, 0.999999, 1.000001, 709.0, 720.0]])
    np.savez_compressed(os.path.join(HERE, "retrieval_pt.npz"), pressure=p, ptargs=cases.W12_PTARGS,
                        pars=pars, T_const=T_const, T_thorngren=T_thorn, apars=apars, T_adiabatic=T_adia,
                        x=x, expn2=sp.expn(2, x))
    print("retrieval_pt.npz: T range %.0f..%.0f K" % (T_const.min(), T_const.max()))


def golden_pt_smooth(pt):
    """retrieval_pt_smooth.npz: code/PT.py PT_NoInversion / PT_Inversion (Madhusudhan & Seager 2009)
    and PT_piette on seeded draws, called the way BARTfunc.py:176,321 does (pressure reversed to
    top -> bottom, the result reversed back to the atmosphere file's order).  Draws the reference
    refuses (ValueError: negative boundary temperatures) are recorded with phys = 0."""
    rng = np.random.default_rng(4242)
    out = {}
    for tag, pfile in (("a", np.logspace(2, -5, 100)), ("b", np.logspace(2.5, -6, 60))):
        p = pfile[::-1]                                       # BARTfunc.py:176

        def run(fn, pars):
            T, ok = np.zeros((len(pars), len(p))), np.zeros(len(pars), dtype=np.int32)
            for i, q in enumerate(pars):
                try:
                    T[i] = pt.PT_generator(p, q, fn)[::-1]
                    ok[i] = 1
                except ValueError:
                    pass
            return T, ok
        n = 48
        noinv = np.column_stack([rng.uniform(0.25, 1.0, n), rng.uniform(0.08, 0.6, n),
                                 10 ** rng.uniform(-4.5, -1.5, n), 10 ** rng.uniform(-0.5, 1.5, n),
                                 rng.uniform(900, 2600, n)])
        noinv[0] = [0.99, 0.19, 1e-3, 1.0, 1500.0]
        inv = np.column_stack([rng.uniform(0.25, 1.0, n), rng.uniform(0.15, 0.8, n),
                               10 ** rng.uniform(-4.5, -2.5, n), 10 ** rng.uniform(-2.0, -0.5, n),
                               10 ** rng.uniform(0.0, 1.5, n), rng.uniform(900, 2600, n)])
        piette = np.column_stack([rng.uniform(1000, 2200, n)] + [rng.uniform(0, 350, n) for _ in range(7)])
        out["pressure_" + tag] = pfile
        for name, fn, pars in (("noinv", pt.PT_NoInversion, noinv), ("inv", pt.PT_Inversion, inv),
                               ("piette", pt.PT_piette, piette)):
            T, ok = run(fn, pars)
            out["%s_pars_%s" % (name, tag)] = pars
            out["%s_T_%s" % (name, tag)] = T
            out["%s_phys_%s" % (name, tag)] = ok
            print("retrieval_pt_smooth %s/%s: %d physical of %d, T %.0f..%.0f K" % (
                name, tag, ok.sum(), len(ok), T[ok == 1].min(), T[ok == 1].max()))
    np.savez_compressed(os.path.join(HERE, "retrieval_pt_smooth.npz"), **out)


def golden_tep():
    """tep_systems.npz: the system parameters BARTfunc.py:157-172,204-211,246 extracts from a TEP file
    with the reference's code/reader.py and constants -- statements executed verbatim -- on the two TEP
    files the reference ships (copied to tests/golden/ref_inputs/ as data fixtures)."""
    import shutil
    import reader as rd
    import constants as c
    import scipy.constants as sc
    out = {}
    for name, path in (("wasp12b", os.path.join(REF, "examples", "WASP-12b", "WASP-12b.tep")),
                       ("hd209458b", os.path.join(REF, "inputs", "tep", "HD209458b.tep"))):
        tep = rd.File(path)
        tstar = float(tep.getvalue('Ts')[0])
        rstar = float(tep.getvalue('Rs')[0]) * c.Rsun
        sma = float(tep.getvalue('a')[0]) * sc.au
        rplanet = float(tep.getvalue('Rp')[0]) * c.Rjup
        mplanet = float(tep.getvalue('Mp')[0]) * c.Mjup
        gplanet = 100.0 * sc.G * mplanet / rplanet**2
        out[name] = np.array([tstar, rstar, sma, rplanet, mplanet, gplanet, rplanet / rstar])
        shutil.copy(path, os.path.join(HERE, "ref_inputs", os.path.basename(path)))
    np.savez(os.path.join(HERE, "tep_systems.npz"), **out)
    print("tep_systems.npz:", {k: v[:3] for k, v in out.items()})


def golden_converter(pt, name, tmp):
    case, spec, extra = cases.build_retrieval(name, tmp)
    rng = np.random.default_rng(777)
    lo, hi = np.array(spec["pmin"]), np.array(spec["pmax"])
    M = 48
    params = rng.uniform(lo, hi, (M, len(lo)))
    params[0] = spec["params"]
    params[1] = spec["truth"]
    params[2:12, :5] = np.array(spec["params"][:5]) + rng.normal(0, 0.05, (10, 5))
    params[8:12, -1] = 4.2                                   # sum of metals > 1 (BARTfunc.py:339-344)
    # --- BARTfunc.py:139-222 set-up, verbatim
    species = np.asarray(case["species"])
    pressure = case["press_bar"][::-1]
    abundances = case["abund"]
    nlayers, nspecies = len(pressure), len(species)
    iH2 = np.where(species == "H2")[0]
    iHe = np.where(species == "He")[0]
    ratio = (abundances[:, iH2] / abundances[:, iHe]).squeeze()
    imetals = np.where((species != "He") & (species != "H2") & (species != "H-") & (species != "e-"))[0]
    molfit = spec["molfit"]
    nmolfit = len(molfit)
    imol = np.zeros(nmolfit, dtype="i")
    for i in np.arange(nmolfit):
        imol[i] = np.where(np.asarray(species) == molfit[i])[0][0]
    nradfit, ncloud, nray = spec["nrad"], spec["ncloud"], spec["nray"]
    nPT = len(lo) - nmolfit - ncloud - nray - nradfit
    PTargs = list(extra["pt_args"]) + ["const"]
    profiles = np.zeros((nspecies + 1, nlayers), dtype="d")
    tprofile, aprofiles = profiles[0, :], profiles[1:, :]
    for i in np.arange(nspecies):
        aprofiles[i] = abundances[:, i]
    Tmin, Tmax = 400.0, 3000.0
    out = np.zeros((M, (nspecies + 1) * nlayers))
    status = np.zeros(M, dtype=np.int32)
    for m in range(M):
        par = params[m]
        # --- BARTfunc.py:320-347, verbatim
        tprofile[:] = pt.PT_generator(pressure, par[0:nPT], pt.PT_line, PTargs)[::-1]
        if np.any(tprofile < Tmin) or np.any(tprofile > Tmax):
            status[m] = 16
            continue
        for i in np.arange(nmolfit):
            mm = imol[i]
            aprofiles[mm] = abundances[:, mm] * 10.0 ** par[nPT + nradfit + ncloud + nray + i]
        q = 1.0 - np.sum(aprofiles[imetals], axis=0)
        if np.any(q < 0.0):
            status[m] = 32
            continue
        aprofiles[iH2] = ratio * q / (1.0 + ratio)
        aprofiles[iHe] = q / (1.0 + ratio)
        out[m] = profiles.flatten()
    np.savez_compressed(os.path.join(HERE, "retrieval_conv_%s.npz" % name), params=params,
                        profiles=out, status=status)
    print("retrieval_conv_%s.npz: %d ok, %d T-rejected, %d abundance-rejected" % (
        name, (status == 0).sum(), (status == 16).sum(), (status == 32).sum()))


def make_band_oracle(name, tmp):
    case, spec, extra = cases.build_retrieval(name, tmp)
    conv = ro.Converter(case["press_bar"], case["species"], case["abund"], spec["molfit"], spec["pt"],
                        pt_args=extra["pt_args"], nrad=spec["nrad"], ncloud=spec["ncloud"],
                        nray=spec["nray"])
    band = ro.BandOracle(case["cfg"], conv, case["filters"], extra["starwn"], extra["starfl"],
                         extra["rprs"])
    return case, spec, band


def retrieval_data(spec, band):
    """Synthetic observation: band fluxes at `truth` + seeded 1 % noise."""
    truth = band(np.array(spec["truth"]))[0]
    rng = np.random.default_rng(spec["seed"])
    uncert = 0.01 * np.abs(truth)
    return truth + rng.normal(0, 1, truth.shape) * uncert, uncert


def golden_mc3(mc3, name, tmp):
    case, spec, band = make_band_oracle(name, tmp)
    data, uncert = retrieval_data(spec, band)
    sav = os.path.join(tmp, name + "_trace.npy")
    savm = os.path.join(tmp, name + "_models.npy")
    log = open(os.path.join(tmp, name + ".log"), "w")
    np.random.seed(spec["seed"])
    out = mc3.mc.mcmc(data, uncert, band, [], params=np.array(spec["params"]),
                      pmin=np.array(spec["pmin"]), pmax=np.array(spec["pmax"]),
                      stepsize=np.array(spec["stepsize"], dtype=float), numit=spec["numit"],
                      nchains=spec["nchains"], walk="demc", leastsq=False, grtest=False,
                      burnin=spec["burnin"], plots=False, savefile=sav, savemodel=savm, log=log)
    allparams = np.load(sav)                       # [nchains][nfree][chainsize] (mcmc.py:842-843)
    allmodel = np.load(savm)                       # [nchains][ndata][chainsize] (mcmc.py:647-651,849-850)
    # the same run resumed (mcmc.py:254-269): MC3 loads savefile / savemodel, starts every chain from
    # its last state and appends; a fresh seed for the new random streams
    np.random.seed(spec["seed"] + 1)
    log2 = open(os.path.join(tmp, name + "_resume.log"), "w")
    out2 = mc3.mc.mcmc(data, uncert, band, [], params=np.array(spec["params"]),
                       pmin=np.array(spec["pmin"]), pmax=np.array(spec["pmax"]),
                       stepsize=np.array(spec["stepsize"], dtype=float), numit=spec["numit"],
                       nchains=spec["nchains"], walk="demc", leastsq=False, grtest=False,
                       burnin=spec["burnin"], plots=False, savefile=sav, savemodel=savm, log=log2,
                       resume=True)
    np.savez_compressed(os.path.join(HERE, "retrieval_mc3_%s.npz" % name), data=data, uncert=uncert,
                        allparams=allparams, allstack=out[0], bestp=out[1], allmodel=allmodel,
                        resume_allparams=np.load(sav), resume_allmodel=np.load(savm),
                        resume_allstack=out2[0], resume_bestp=out2[1])
    print("retrieval_mc3_%s.npz: trace %s, posterior %s, best %s" % (
        name, allparams.shape, out[0].shape, np.array2string(out[1], precision=4)))


class FakeWorkers:
    """Stands in for the MPI intercommunicator MC3 talks to in BART (mcmc.py:277-286,317-322,
    582-585; MCcubed/utils/mcutils.py:208-281): Scatter hands over the proposals of all chains,
    Gather returns the workers' models.  Test-side shim so that the reference takes its MPI-mode
    branches (every proposal evaluated, no early `continue`) without mpi4py/mpiexec."""

    def __init__(self, func, nchains):
        self.func, self.nchains, self.pending = func, nchains, None

    def Barrier(self):
        pass

    def Bcast(self, buf, root=0):
        pass

    def Scatter(self, send, recv, root=0):
        arr = np.asarray(send[0], dtype=float)
        self.pending = arr.reshape(self.nchains, -1).copy()

    def Gather(self, send, recv, root=0):
        recv[:] = np.asarray(self.func(self.pending), dtype=float).ravel()

    def Disconnect(self):
        pass


def golden_snooker(mc3, name, tmp):
    case, spec, band = make_band_oracle(name, tmp)
    data, uncert = retrieval_data(spec, band)
    mpi4py = types.ModuleType("mpi4py")
    mpi4py.MPI = types.SimpleNamespace(DOUBLE="d", INT="i", ROOT=-3)
    sys.modules["mpi4py"] = mpi4py
    sys.modules["mpi4py.MPI"] = mpi4py.MPI
    import MCcubed.utils.mcutils as mcu
    mcu.MPI = mpi4py.MPI
    for thinning in (1, 3):
        sav = os.path.join(tmp, "%s_snk%d_trace.npy" % (name, thinning))
        savm = os.path.join(tmp, "%s_snk%d_models.npy" % (name, thinning))
        log = open(os.path.join(tmp, "%s_snk%d.log" % (name, thinning)), "w")
        np.random.seed(spec["seed"] + thinning)
        comm = FakeWorkers(band, spec["nchains"])
        out = mc3.mc.mcmc(data, uncert, band, [], params=np.array(spec["params"]),
                          pmin=np.array(spec["pmin"]), pmax=np.array(spec["pmax"]),
                          stepsize=np.array(spec["stepsize"], dtype=float), numit=spec["numit"],
                          nchains=spec["nchains"], walk="snooker", leastsq=False, grtest=False,
                          burnin=spec["burnin"], thinning=thinning, plots=False, savefile=sav,
                          savemodel=savm, log=log, comm=comm)
        allparams = np.load(sav)
        allmodel = np.load(savm)
        np.savez_compressed(os.path.join(HERE, "retrieval_snooker_%s_thin%d.npz" % (name, thinning)),
                            data=data, uncert=uncert, allparams=allparams, allstack=out[0],
                            bestp=out[1], thinning=thinning, allmodel=allmodel)
        print("retrieval_snooker_%s_thin%d.npz: trace %s, best %s" % (
            name, thinning, allparams.shape, np.array2string(out[1], precision=4)))


def golden_gr():
    """Gelman-Rubin factors of the reference's MCcubed/mc/gelman_rubin.py at MC3's own checkpoints
    (mcmc.py:238,663-686) on the stored DE-MC traces and on a synthetic converged population."""
    import MCcubed.mc.gelman_rubin as gr
    out = {}
    for name in cases.RETRIEVAL:
        d = np.load(os.path.join(HERE, "retrieval_mc3_%s.npz" % name))
        allp, burnin = d["allparams"], cases.RETRIEVAL[name]["burnin"]
        chainsize = allp.shape[2]
        intsteps = chainsize / 10
        its, vals = [], []
        for i in range(chainsize):
            if ((i + 1) % intsteps == 0) and (i > 0) and i > burnin:
                its.append(i)
                vals.append(gr.convergetest(allp[:, :, burnin:i + 1:1]))
        out["its_" + name], out["psrf_" + name] = np.array(its), np.array(vals)
    rng = np.random.RandomState(11)
    conv = rng.normal(0.0, 1.0, (7, 3, 400)) + np.array([1.0, -2.0, 0.5])[None, :, None]
    out["conv_chains"] = conv
    out["conv_psrf_thin2"] = gr.convergetest(conv[:, :, 50:300:2])
    np.savez_compressed(os.path.join(HERE, "retrieval_gr.npz"), **out)
    print("retrieval_gr.npz:", {k: np.shape(v) for k, v in out.items()})


def main():
    only = set(sys.argv[1:])                     # e.g. `make_golden_retrieval.py snooker`
    want = lambda k: not only or k in only
    shim()
    sys.path.insert(0, os.path.join(REF, "code"))
    import PT as pt
    mc3 = build_mc3()
    if want("pt"):
        golden_pt(pt)
    if want("ptsmooth"):
        golden_pt_smooth(pt)
    if want("gr"):
        golden_gr()
    if want("tep"):
        golden_tep()
    with tempfile.TemporaryDirectory() as tmp:
        for name in cases.RETRIEVAL:
            if want("conv"):
                golden_converter(pt, name, tmp)
            if want("demc"):
                golden_mc3(mc3, name, tmp)
            if want("snooker"):
                golden_snooker(mc3, name, tmp)


if __name__ == "__main__":
    main()
