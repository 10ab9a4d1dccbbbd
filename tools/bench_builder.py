#!/usr/bin/env python
"""Opacity-grid builder (--justOpacity; opacity.c:218-427, extinction.c:281-529) throughput on a
bounded sample of BASELINE.json configs[3] (1e8 lines onto 1e5 wn x 100 layers x 20 T): the same
line density per wavenumber bin (default 1000 lines/bin), 100 layers, a slice of the temperature
axis.  Reports line x (T,layer) cells per second and the per-phase device times.
usage: bench_builder.py [--nlines 2400000] [--shape w12|demo] [--ntemp 2] [--nlayer 100]
                        [--wndelt 1.0] [--check N]"""
import argparse, ctypes as C, json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bart_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nlines", type=int, default=2400000)
ap.add_argument("--wnlow", type=float, default=910.0)
ap.add_argument("--wnhigh", type=float, default=3333.0)
ap.add_argument("--wndelt", type=float, default=1.0)
ap.add_argument("--wnosamp", type=int, default=2160)
ap.add_argument("--mols", default="H2O,CO2,CO,CH4")
ap.add_argument("--ntemp", type=int, default=2, help="temperatures of the slice that is built")
ap.add_argument("--nlayer", type=int, default=100)
ap.add_argument("--ethresh", type=float, default=1e-6)
ap.add_argument("--check", type=int, default=0, help="compare N (layer,T) cells with the builder oracle")
ap.add_argument("--ntemp-grid", type=int, default=27, help="temperatures of the grid (400 K upwards, 100 K apart)")
ap.add_argument("--planes", default="", help="comma list of temperature indices to build (default: --ntemp planes spread over the grid)")
a = ap.parse_args()

tmp = tempfile.mkdtemp(prefix="bart_build_")
t0 = time.time()
shape = dict(wnlow=a.wnlow, wnhigh=a.wnhigh, wndelt=a.wndelt, mols=a.mols.split(","), toomuch=10.0)
case = synth.make_case(tmp, shape=shape, nlayer=a.nlayer, with_grid=False, nlines=a.nlines,
                       tempdelt=100.0, thigh=400.0 + 100.0 * (a.ntemp_grid - 1), seed=2026, ethresh=a.ethresh,
                       wnosamp=a.wnosamp)
t_gen = time.time() - t0
L = api.lib()
t0 = time.time()
api.device_info()                                   # CUDA initialisation, timed on its own
t_cuda = time.time() - t0
# init without building the file: BART_TSLICE=0:0 makes --justOpacity build an empty slice
os.environ["BART_TSLICE"] = "0:0"
t0 = time.time()
tr = api.Transit(argv=["transit", "-c", case["cfg"], "--justOpacity"])
t_init = time.time() - t0
nl, nw = a.nlayer, len(case["wn"])
nmol = len(shape["mols"])
nt_all = len(case["grid_temps"])
# spread the slice over the grid's temperature range (cold and hot planes cost differently)
picks = sorted(set(int(round(x)) for x in np.linspace(0, nt_all - 1, a.ntemp + 2)[1:-1]))
if a.planes:
    picks = [int(x) for x in a.planes.split(",")]
out = np.zeros((nl, 1, nmol, nw))
names = ("read_tli_host", "grouping_host", "voigt_table", "kmax", "strength", "widths", "accumulate", "d2h")
once = ("read_tli_host", "grouping_device", "lines_h2d_index_d2h", "grouping_host", "groups_h2d", "voigt_table")
t0 = time.time()
api._check(L.bart_build_opacity_slice(picks[-1], picks[-1] + 1, out.ctypes.data_as(api.dp)))   # warm-up: allocations
t_warm = time.time() - t0
before = {n: L.bart_builder_phase_ms(n.encode()) for n in names}
t0 = time.time()
for it in picks:
    api._check(L.bart_build_opacity_slice(it, it + 1, out.ctypes.data_as(api.dp)))
wall = time.time() - t0
after = {n: L.bart_builder_phase_ms(n.encode()) for n in names}
nlines, ngroups, neval = C.c_longlong(), C.c_longlong(), C.c_longlong()
L.bart_builder_stats(C.byref(nlines), C.byref(ngroups), C.byref(neval))
cells = len(picks) * nl
dev_ms = sum(after[n] - before[n] for n in ("kmax", "strength", "widths", "accumulate"))
res = {"shape": {"nwave": nw, "nlayer": nl, "ntemp_built": len(picks), "ntemp_grid": nt_all, "nmol": nmol,
                 "wnosamp": a.wnosamp, "wndelt": a.wndelt, "lines_per_bin": a.nlines / nw},
       "nlines_in_range": nlines.value, "ngroups": ngroups.value, "evaluated_group_cells": neval.value,
       "gen_s": t_gen, "cuda_init_s": t_cuda, "init_s": t_init,
       "first_slice_s": t_warm,
       "one_time_ms": {n: L.bart_builder_phase_ms(n.encode()) for n in once},
       "per_slice_ms": {n: after[n] - before[n] for n in names if after[n] - before[n] > 0},
       "wall_s": wall, "device_ms": dev_ms,
       "line_cells_per_s_device": nlines.value * cells / (dev_ms * 1e-3) if dev_ms > 0 else None,
       "line_cells_per_s_wall": nlines.value * cells / wall,
       "full_config_estimate_s": None}
if dev_ms > 0:
    # BASELINE configs[3]: 1e8 lines x 100 layers x 20 T
    res["full_config_estimate_s"] = 1e8 * 2000 / res["line_cells_per_s_device"]
if a.check:
    from oracle import oracle as orc
    B = orc.BuilderOracle(case["cfg"])
    rng = np.random.default_rng(1)
    worst = 0.0
    for _ in range(a.check):
        r = int(rng.integers(0, nl)); it = picks[int(rng.integers(0, len(picks)))]
        ref = B.build(layers=[r], temps=[it])
        api._check(L.bart_build_opacity_slice(it, it + 1, out.ctypes.data_as(api.dp)))
        got = out[r, 0]
        refc = np.asarray(ref).reshape(got.shape)
        m = refc > 0
        worst = max(worst, float(np.max(np.abs(got[m] - refc[m]) / refc[m])))
        assert np.array_equal(got > 0, m)
    res["oracle_check"] = {"cells": a.check, "max_rel_err": worst}
print(json.dumps(res))
