#!/usr/bin/env python
"""Transit-geometry forward models (BASELINE.json configs[1]: examples/demo BART_transit.cfg shape --
CH4 grid, 2501 wavenumbers, 100 layers, slantpath.c chord optical depth + modulation) per second,
device-resident proposals -> band fluxes, one batch of --models per step, L2 flushed between steps.
usage: bench_transit.py [--models 4096] [--steps 8] [--shape demo|w12]"""
import argparse, json, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bart_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--models", type=int, default=4096)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--shape", default="demo")
a = ap.parse_args()
tmp = tempfile.mkdtemp(prefix="bart_tr_")
case = synth.make_case(tmp, shape=a.shape, solution="transit", seed=2026, refradius_km=95000.0)
molfit = ("CH4",) if len(case["shape"]["mols"]) == 1 else ("H2O", "CO2", "CO", "CH4")
models = synth.make_models(case, a.models, seed=2026, molfit=molfit)
tr = api.Transit(case["cfg"])
L = api.lib()
wn = tr.get_waveno_arr()
start, count, weight, _ = api.filters_from_files(wn, case["filters"])
tr.set_filters(start, count, weight, None, 1.0)
M, n_in = a.models, tr.n_in
d_prof = L.bart_dev_alloc(M * n_in * 8)
d_band = L.bart_dev_alloc(M * tr.nfilters * 8)
L.bart_memcpy_h2d(d_prof, models.ctypes.data, M * n_in * 8)
for _ in range(3):
    api._check(L.bart_bandflux_batch_device(d_prof, M, n_in, d_band, None))
L.bart_profile_reset()
L.bart_profile_enable(1)
ms = []
for _ in range(a.steps):
    L.bart_flush_l2()
    L.bart_sync()
    L.bart_timer_begin()
    api._check(L.bart_bandflux_batch_device(d_prof, M, n_in, d_band, None))
    ms.append(L.bart_timer_end())
L.bart_profile_enable(0)
ks = {k: v["ms"] / v["launches"] for k, v in api.kernel_stats().items()}
band = np.empty((M, tr.nfilters))
L.bart_memcpy_d2h(band.ctypes.data, d_band, band.nbytes)
print(json.dumps({"workload": "transit geometry, %s shape: %d wavenumbers x %d layers, %d grid molecule(s), %d filters"
                              % (a.shape, tr.nwave, tr.nlayer, L.bart_ngridmol(), tr.nfilters),
                  "models_per_step": M, "ms_per_step": float(np.mean(ms)),
                  "spectra_per_s": M / (float(np.mean(ms)) * 1e-3), "kernel_ms": ks,
                  "chord_product": "fp64 tensor cores (DMMA m8n8k4)" if os.environ.get("BART_TRANSIT_MMA", "1") != "0" else "DFMA",
                  "finite": bool(np.isfinite(band).all())}))
