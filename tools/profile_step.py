#!/usr/bin/env python
"""Run a few device-resident generations of the bench workload (for ncu / launch lists).
usage: profile_step.py [--models M] [--steps K] [--shape w12|demo] [--solution eclipse|transit]"""
import argparse
import os
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bart_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--models", type=int, default=1024)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--shape", default="w12")
ap.add_argument("--solution", default="eclipse")
ap.add_argument("--lookup", action="store_true", help="also run the stand-alone lookup kernel")
a = ap.parse_args()
tmp = tempfile.mkdtemp(prefix="bart_prof_")
case = synth.make_case(tmp, shape=a.shape, solution=a.solution, seed=2026,
                       refradius_km=95000.0 if a.solution == "transit" else 123820.0)
molfit = ("CH4",) if len(case["shape"]["mols"]) == 1 else ("H2O", "CO2", "CO", "CH4")
models = synth.make_models(case, a.models, seed=2026, molfit=molfit)
tr = api.Transit(case["cfg"])
L = api.lib()
wn = tr.get_waveno_arr()
start, count, weight, star = api.filters_from_files(wn, case["filters"], wn, np.ones_like(wn))
tr.set_filters(start, count, weight, star, 0.1)
M, n_in = a.models, tr.n_in
d_prof = L.bart_dev_alloc(M * n_in * 8)
d_band = L.bart_dev_alloc(M * tr.nfilters * 8)
L.bart_memcpy_h2d(d_prof, models.ctypes.data, M * n_in * 8)
L.bart_profile_enable(1)
for _ in range(a.steps):
    L.bart_flush_l2()
    api._check(L.bart_bandflux_batch_device(d_prof, M, n_in, d_band, None))
if a.lookup:
    tr.extinction_batch(models[:min(M, 64)], total=False, fetch=False)
print(api.kernel_stats())
