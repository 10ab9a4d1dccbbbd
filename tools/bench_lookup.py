#!/usr/bin/env python
"""Stand-alone opacity-lookup kernel (K1, extinction.c:534-581) at the high-resolution sweep shape:
achieved algorithmic GB/s against the measured HBM copy peak (per model
16*Nmol*Nlayer*Nwave bytes read + 8*Nlayer*Nwave written), plus the fused eclipse kernel at the
same shape.  usage: bench_lookup.py [--nwave 100001] [--models 16] [--steps 5] [--fused-models 256]"""
import argparse, json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bart_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nwave", type=int, default=100001)
ap.add_argument("--ntemp", type=int, default=20)
ap.add_argument("--models", type=str, default="1,4,16,64", help="comma list of batch sizes for K1")
ap.add_argument("--fused-models", type=int, default=256)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--profile-only", action="store_true", help="one launch of each kernel (for ncu)")
a = ap.parse_args()
tmp = tempfile.mkdtemp(prefix="bart_hr_")
t0 = time.time()
case = synth.make_hr_case(tmp, nwave=a.nwave, ntemp=a.ntemp)
t_gen = time.time() - t0
tmax = float(case["grid_temps"][-1])
mlist = [int(x) for x in a.models.split(",")]
models = synth.make_models(case, max(max(mlist), a.fused_models), seed=2026,
                           molfit=("H2O", "CO2", "CO", "CH4"), tmax=tmax)
t0 = time.time()
tr = api.Transit(case["cfg"])
t_init = time.time() - t0
L = api.lib()
nw, nl, nmol = tr.nwave, tr.nlayer, L.bart_ngridmol()
peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = {"shape": {"nwave": nw, "nlayer": nl, "ntemp": a.ntemp, "nmol": nmol,
                 "grid_GB": L.bart_grid_bytes() / 1e9}, "gen_s": t_gen, "init_s": t_init, "hbm_peak_gbs": peak}
steps = 1 if a.profile_only else a.steps
# --- K1
out["lookup"] = []
L.bart_profile_enable(1)
for M in mlist:
    tr.extinction_batch(models[:M], total=False, fetch=False)          # warm-up (allocations)
    L.bart_profile_reset()
    for _ in range(steps):
        L.bart_flush_l2()
        tr.extinction_batch(models[:M], total=False, fetch=False)
    st = api.kernel_stats()["opacity_lookup"]
    ms = st["ms"] / st["launches"]
    alg = M * (16.0 * nmol * nl * nw + 8.0 * nl * nw)
    out["lookup"].append({"models": M, "ms_per_launch": ms, "algorithmic_GB": alg / 1e9,
                          "achieved_gbs": alg / ms / 1e6, "frac_of_hbm_peak": alg / ms / 1e6 / peak})
# --- fused eclipse kernel at the same shape
M2 = a.fused_models
wn = tr.get_waveno_arr()
start, count, weight, star = api.filters_from_files(wn, case["filters"], wn, np.ones_like(wn))
tr.set_filters(start, count, weight, star, 0.1)
d_prof = L.bart_dev_alloc(M2 * tr.n_in * 8)
d_band = L.bart_dev_alloc(M2 * tr.nfilters * 8)
L.bart_memcpy_h2d(d_prof, models[:M2].ctypes.data, M2 * tr.n_in * 8)
api._check(L.bart_bandflux_batch_device(d_prof, M2, tr.n_in, d_band, None))
L.bart_profile_reset()
for _ in range(steps):
    L.bart_flush_l2()
    api._check(L.bart_bandflux_batch_device(d_prof, M2, tr.n_in, d_band, None))
ks = api.kernel_stats()
st = ks["eclipse_column"]
ms = st["ms"] / st["launches"]
alg = M2 * (16.0 * nmol * nl * nw + 8.0 * nw)
step_ms = sum(v["ms"] for v in ks.values()) / st["launches"]
out["fused_eclipse"] = {"models": M2, "ms_per_launch": ms, "spectra_per_s": M2 / (step_ms * 1e-3),
                        "algorithmic_GB": alg / 1e9, "achieved_gbs": alg / ms / 1e6,
                        "frac_of_hbm_peak": alg / ms / 1e6 / peak}
print(json.dumps(out))
