#!/usr/bin/env python
"""tests/golden/make_golden_bartfunc.py -- the reference's worker, UNMODIFIED, against the
transit_module surface ("code/BARTfunc.py drives it unchanged").

Runs /root/reference/code/BARTfunc.py::main(comm) as shipped in this (GPU-less) container with
  * a stand-in `mpi4py` whose communicator plays MC3's master: it broadcasts (npars, niter),
    scatters a seeded list of parameter vectors (incl. one T-bounds and one abundance rejection,
    BARTfunc.py:327-344), sends MC3's end flag, and collects what the worker gathers;
  * a RECORDING `transit_module` with exactly the SWIG surface of transit/src/transit.i:12-31,
    whose spectra come from the forward-model oracle (oracle/transit_oracle.c): every call the
    worker makes -- name, argument types / shapes, order -- is written down.
The recorded call trace, the parameter vectors, and the band fluxes the worker sent back are the
fixture tests/golden/bartfunc_trace.npz.  tests/test_gpu_bartfunc.py replays the SAME calls on the
real CUDA-backed bart_b200/python/transit_module and must reproduce the band fluxes; the case's
input files are regenerated from bart_b200.synth (seeded), the star / filter arrays wine.py derived
are carried in the fixture (the reference's Kurucz file does not travel).
Shims from the outside only: numpy.int/float aliases, stub matplotlib.
usage: python tests/golden/make_golden_bartfunc.py
"""
import json
import os
import sys
import tempfile
import types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from bart_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASE = dict(shape="tiny", solution="eclipse", seed=2040)
MOLFIT = ["CH4"]
# PT_line (kappa, g1, g2, alpha, beta) + log10 CH4 factor
PARAMS = np.array([[-0.5, -0.2, 1.0, 0.0, 1.10, 0.5],
                   [-0.7, -0.1, 1.0, 0.0, 1.05, 1.0],
                   [-0.4, -0.3, 0.8, 0.1, 1.00, -0.5],
                   [-0.5, -0.2, 1.0, 0.0, 3.00, 0.5],      # far too hot: T-bounds rejection
                   [-0.5, -0.2, 1.0, 0.0, 1.10, 4.5],      # CH4 x 10^4.5: sum of metals > 1
                   [-0.6, -0.25, 1.1, 0.05, 1.08, 0.0]])

TEP = """# synthetic TEP file (WASP-12b values, examples/WASP-12b/WASP-12b.tep)
planetname WASP-12b -1 - -
Ts 6300 150 K -
Rs 1.57 0.07 Rsun -
loggstar 4.17 0.03 cgs -
a 0.0229 0.0008 AU -
Rp 1.79 0.09 Rjup -
Mp 1.41 0.10 Mjup -
"""


def shim():
    np.int = int
    np.float = float
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.cm"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].use = lambda *a, **k: None


class Master:
    """MC3's side of the worker protocol (MCcubed/utils/mcutils.py comm_bcast / comm_scatter /
    comm_gather as the WORKER calls them: mpitype None = receive, else send)."""

    def __init__(self, params):
        self.params = list(params)
        self.gathered = []
        self.sent = 0

    def Get_rank(self):
        return 0

    def Barrier(self):
        pass

    def Bcast(self, buf, root=0):
        buf[:] = [len(self.params[0]), len(self.params)]

    def Scatter(self, send, recv, root=0):
        if self.sent < len(self.params):
            recv[:] = self.params[self.sent]
        else:
            recv[:] = np.inf                                   # MC3's end flag (BARTfunc.py:313)
        self.sent += 1

    def Gather(self, send, recv, root=0):
        arr = send[0] if isinstance(send, (list, tuple)) else send
        self.gathered.append(np.array(arr, dtype=float).copy())

    def Disconnect(self):
        pass


def describe(x):
    if isinstance(x, np.ndarray):
        return {"ndarray": list(x.shape), "dtype": str(x.dtype), "contiguous": bool(x.flags["C_CONTIGUOUS"])}
    if isinstance(x, (list, tuple)):
        return {"list": [describe(v) for v in x]}
    return {type(x).__name__: x if isinstance(x, (int, float, str)) else repr(x)}


def main():
    shim()
    tmp = tempfile.mkdtemp(prefix="bartfunc_")
    case = synth.make_case(os.path.join(tmp, "case"), **CASE)
    calls = []
    state = {}

    trm = types.ModuleType("transit_module")

    def transit_init(argc, argv):
        calls.append(("transit_init", [describe(argc), describe(argv)]))
        assert argc == len(argv) and argv[:2] == ["transit", "-c"]
        state["O"] = orc.Oracle(argv[2])

    def get_no_samples():
        calls.append(("get_no_samples", []))
        return len(state["O"].wn)

    def get_waveno_arr(n):
        calls.append(("get_waveno_arr", [describe(n)]))
        return np.array(state["O"].wn[:n])

    def set_radius(r):
        calls.append(("set_radius", [describe(r)]))
        state["O"].set_radius(r)

    def set_cloudtop(v):
        calls.append(("set_cloudtop", [describe(v)]))
        state["O"].set_cloudtop(v)

    def set_scattering(f, v):
        calls.append(("set_scattering", [describe(f), describe(v)]))
        state["O"].set_scattering(f, v)

    def run_transit(profiles, nwave):
        calls.append(("run_transit", [describe(profiles), describe(nwave)]))
        state.setdefault("profiles", []).append(np.array(profiles))
        return state["O"].run(np.ascontiguousarray(profiles))

    def free_memory():
        calls.append(("free_memory", []))

    for f in (transit_init, get_no_samples, get_waveno_arr, set_radius, set_cloudtop, set_scattering,
              run_transit, free_memory):
        setattr(trm, f.__name__, f)
    sys.modules["transit_module"] = trm
    mpi = types.ModuleType("mpi4py")
    mpi.MPI = types.SimpleNamespace(DOUBLE="DOUBLE", INT="INT", ROOT=-3)
    sys.modules["mpi4py"] = mpi

    tepfile = os.path.join(tmp, "planet.tep")
    open(tepfile, "w").write(TEP)
    kurucz = os.path.join(REF, "inputs", "kurucz", "fp00k2odfnew.pck")
    cfg = os.path.join(tmp, "MCMC.cfg")
    with open(cfg, "w") as f:
        f.write("[MCMC]\n")
        f.write("params = %s\n" % " ".join("%r" % v for v in PARAMS[0]))
        f.write("molfit = %s\n" % " ".join(MOLFIT))
        f.write("atmfile = %s\nPTtype = line\ntint = 100.0\ntint_type = const\n" % case["atm"])
        f.write("tconfig = %s\n" % case["cfg"])
        f.write("filters = %s\n" % "\n    ".join(case["filters"]))
        f.write("tep_name = %s\nkurucz = %s\nsolution = eclipse\n" % (tepfile, kurucz))
    # MC3's C extensions (chisq, ...) are compiled from a copy of the reference's MCcubed under /tmp by
    # the reference's own setup.py (make_golden_retrieval.build_mc3); BARTfunc appends the reference's
    # MCcubed directory to sys.path, the built copy -- the same sources -- is found first
    sys.path.insert(0, HERE)
    import make_golden_retrieval as mgr
    mgr.shim()
    mgr.build_mc3()
    sys.path.insert(0, os.path.join(REF, "code"))
    sys.argv = ["BARTfunc.py", "-c", cfg]
    import BARTfunc  # noqa: E402  (the reference worker, unmodified)
    master = Master(PARAMS)
    BARTfunc.main(master)
    band = np.array(master.gathered)
    # what wine.py derived for this case (carried for the replay: the Kurucz file does not travel)
    import wine as w
    import reader as rd
    import constants as c
    tep = rd.File(tepfile)
    tstar, gstar = float(tep.getvalue('Ts')[0]), float(tep.getvalue('loggstar')[0])
    rprs = float(tep.getvalue('Rp')[0]) * c.Rjup / (float(tep.getvalue('Rs')[0]) * c.Rsun)
    starfl, starwn, _, _ = w.readkurucz(kurucz, tstar, gstar)
    specwn = np.array(state["O"].wn)
    nif, ist, idx = [], [], []
    for ff in case["filters"]:
        fw, ft = w.readfilter(ff)
        a, b, i = w.resample(specwn, fw, ft, starwn, starfl)
        nif.append(a); ist.append(b); idx.append(i[0])
    out = dict(params=PARAMS, bandflux=band, rprs=rprs, specwn=specwn,
               calls=json.dumps(calls), profiles=np.array(state["profiles"]),
               start=np.array([i[0] for i in idx]), count=np.array([len(i) for i in idx]),
               weight=np.concatenate(nif), star=np.concatenate(ist),
               case=json.dumps(CASE), grid_sha=__import__("hashlib").sha256(
                   np.fromfile(case["opacity"], dtype=np.uint8).tobytes()).hexdigest())
    np.savez_compressed(os.path.join(HERE, "bartfunc_trace.npz"), **out)
    names = [c_[0] for c_ in calls]
    print("bartfunc_trace.npz: %d calls (%s ...), %d gathers, rejected rows: %s" % (
        len(calls), " ".join(names[:6]), len(band), [int(i) for i in np.where(band[:, 0] == -1)[0]]))


if __name__ == "__main__":
    main()
