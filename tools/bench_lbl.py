#!/usr/bin/env python
"""Line-by-line forward mode (no opacity file; tau.c:163-175,253-264 -> computemolext(permol=0))
throughput: W12 shape, synthetic TLI of --nlines lines, M models per batch through bart_run_batch
(host profiles in, spectra out).  --ref N also times the UNMODIFIED reference (oracle/_ref) on N
models of the same batch, single process, and checks the spectra against it.
usage: bench_lbl.py [--nlines 2400000] [--models 8] [--reps 3] [--ref 1]"""
import argparse, json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bart_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nlines", type=int, default=2400000)
ap.add_argument("--models", type=int, default=8)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--nlayer", type=int, default=100)
ap.add_argument("--ref", type=int, default=0)
ap.add_argument("--solution", default="eclipse")
a = ap.parse_args()

tmp = tempfile.mkdtemp(prefix="bart_lbl_")
case = synth.make_case(tmp, shape="w12", solution=a.solution, nlayer=a.nlayer, with_grid=False,
                       no_opacity=True, nlines=a.nlines, seed=2026, ethresh=1e-6)
models = synth.make_models(case, a.models, seed=7, molfit=("H2O", "CO2", "CO", "CH4"))
t0 = time.time()
tr = api.Transit(case["cfg"])
t_init = time.time() - t0
L = api.lib()
spectra, status = tr.run_batch(models)           # warm-up (allocations)
assert (status == 0).all()
names = ("kmax", "strength", "widths", "accumulate")
before = {n: L.bart_builder_phase_ms(n.encode()) for n in names}
t0 = time.time()
for _ in range(a.reps):
    spectra, status = tr.run_batch(models)
wall = (time.time() - t0) / a.reps
after = {n: L.bart_builder_phase_ms(n.encode()) for n in names}
res = {"shape": {"nwave": tr.nwave, "nlayer": tr.nlayer, "nlines": a.nlines, "models_per_batch": a.models,
                 "solution": a.solution},
       "init_s": t_init, "s_per_batch": wall, "models_per_s": a.models / wall,
       "builder_ms_per_batch": {n: (after[n] - before[n]) / a.reps for n in names},
       "line_cells_per_s": a.nlines * a.models * tr.nlayer / wall}
if a.ref:
    n = min(a.ref, a.models)
    mp, op = os.path.join(tmp, "m.npy"), os.path.join(tmp, "ref.npz")
    np.save(mp, models[:n])
    t0 = time.time()
    subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "ref_driver.py"), case["cfg"], mp, op],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t_ref = time.time() - t0
    d = np.load(op)
    ref = d["spectra"]
    err = float(np.max(np.abs(spectra[:n] - ref) / np.abs(ref)))
    res["reference"] = {"models": n, "wall_s_incl_init": t_ref, "init_s": float(d["t_init"]),
                        "s_per_model": (t_ref - float(d["t_init"])) / n, "cores": 1,
                        "max_rel_err_vs_reference": err}
    res["speedup_vs_reference_core"] = res["reference"]["s_per_model"] * res["models_per_s"]
print(json.dumps(res))
