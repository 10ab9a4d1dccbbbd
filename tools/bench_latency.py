#!/usr/bin/env python
"""Small-population latency at the WASP-12b eclipse shape: BART's real populations are ~10 chains
(examples/WASP-12b/BART.cfg; mccubed.py:277), so a generation is 10 forward models, far too few to
fill 148 SMs with the throughput mapping.  Reports
  us_per_generation   the device-resident DE-MC loop (propose -> converter -> atm_prep -> eclipse
                      columns -> band integration -> chi^2/accept), CUDA-graph replay, 10 chains
  forward_only        one batched call of 10 proposal models -> band fluxes, device-resident inputs
  run_transit         the reference's own one-model call (transit_module.run_transit), host arrays
and the kernel that was picked for the small batch.
usage: bench_latency.py [--chains 10] [--gens 1200]"""
import argparse, json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bart_b200 import api, driver, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=10)
ap.add_argument("--gens", type=int, default=1200)
ap.add_argument("--solution", default="eclipse", choices=("eclipse", "transit"))
a = ap.parse_args()

RSUN, RJUP, MJUP, AU, G = 6.955e10, 7.1492e9, 1.8986e30, 1.4959787e13, 6.67384e-8
PTARGS = (1.57 * RSUN, 6300.0, 100.0, 0.0229 * AU, 100.0 * G * 1.41 * MJUP / (1.79 * RJUP) ** 2)
tmp = tempfile.mkdtemp(prefix="bart_lat_")
if a.solution == "eclipse":
    case = synth.make_case(tmp, shape="w12", solution="eclipse", seed=2026)
else:
    case = synth.make_case(tmp, shape="w12", solution="transit", seed=2026, refradius_km=95000.0)
tr = api.Transit(case["cfg"])
L = api.lib()
wn = tr.get_waveno_arr()
hc_k = 6.6260755e-27 * 2.99792458e10 / 1.380658e-16
star = 2 * 6.6260755e-27 * 2.99792458e10 ** 2 * wn ** 3 / np.expm1(hc_k * wn / 6300.0) * np.pi
start, count, weight, st = api.filters_from_files(wn, case["filters"], wn, star)
tr.set_filters(start, count, weight, st if a.solution == "eclipse" else None, 0.117 if a.solution == "eclipse" else 1.0)
molfit = ("H2O", "CO2", "CO", "CH4")
tr.converter_init(case["press_bar"], case["species"], case["abund"], molfit, "line", pt_args=PTARGS)
# 5 PT_line parameters (kappa, gamma1, gamma2, alpha, beta) + 4 log abundance factors
params = np.array([-0.5, -0.2, 1.0, 0.0, 1.1, 0.3, -0.2, 0.1, 0.2])
pmin = np.array([-5.0, -3.0, -2.0, 0.0, 0.55, -9.0, -9.0, -9.0, -9.0])
pmax = np.array([2.0, 2.0, 3.0, 1.0, 1.4, 3.0, 3.0, 3.0, 3.0])
stepsize = np.array([0.05, 0.05, 0.0, 0.0, 0.01, 0.3, 0.3, 0.3, 0.3])
if a.solution == "transit":          # the planet radius at the reference pressure (km) rides along
    ins = lambda v, x: np.insert(v, 5, x)
    params, pmin, pmax, stepsize = ins(params, 94500.0), ins(pmin, 90000.0), ins(pmax, 99000.0), ins(stepsize, 20.0)
nch = a.chains
truth, status = tr.bandflux_from_params(params[None, :])
assert status[0] == 0, "the reference point of the latency run is rejected by the converter"
data = truth[0]
uncert = 0.02 * np.abs(data)
rng = np.random.RandomState(7)
p0 = np.repeat(params[None, :], nch, 0)
free = stepsize > 0
p0[:, free] += rng.normal(0, 0.01, (nch, int(free.sum())))
ngen = a.gens
dr = driver.demc_draws(rng, nch, ngen, stepsize[free])
tr.mcmc_init(p0, pmin, pmax, stepsize, data, uncert)
warm = slice(0, 100)
tr.mcmc_run(dr["support"][warm], dr["r1"][:, warm], dr["r2"][:, warm], dr["unif"][warm], dr["ugamma"][warm])
rest = slice(100, ngen)
t0 = time.perf_counter()
tr.mcmc_run(dr["support"][rest], dr["r1"][:, rest], dr["r2"][:, rest], dr["unif"][rest], dr["ugamma"][rest])
dt = time.perf_counter() - t0
us_gen = 1e6 * dt / (ngen - 100)
numaccept = tr.mcmc_get("numaccept")
# per-kernel device time of a generation (plain launches with CUDA events; not the latency figure)
L.bart_profile_reset()
L.bart_profile_enable(1)
sl = slice(0, 60)
tr.mcmc_run(dr["support"][sl], dr["r1"][:, sl], dr["r2"][:, sl], dr["unif"][sl], dr["ugamma"][sl])
L.bart_profile_enable(0)
gen_kernels = {k: 1e3 * v["ms"] / v["launches"] for k, v in api.kernel_stats().items()}

# forward only: profiles (device) -> band fluxes (device), 10 models per call
prof, pst, _ = tr.profiles_from_params(tr.mcmc_get("params"))
d_prof = L.bart_dev_alloc(prof.size * 8)
d_band = L.bart_dev_alloc(nch * tr.nfilters * 8)
L.bart_memcpy_h2d(d_prof, prof.ctypes.data, prof.size * 8)
for _ in range(20):
    api._check(L.bart_bandflux_batch_device(d_prof, nch, tr.n_in, d_band, None))
L.bart_sync()
L.bart_profile_reset()
L.bart_profile_enable(1)
reps = 200
t0 = time.perf_counter()
for _ in range(reps):
    api._check(L.bart_bandflux_batch_device(d_prof, nch, tr.n_in, d_band, None))
L.bart_sync()
us_fwd = 1e6 * (time.perf_counter() - t0) / reps
L.bart_profile_enable(0)
ks = {k: 1e3 * v["ms"] / v["launches"] for k, v in api.kernel_stats().items()}
L.bart_dev_free(d_prof); L.bart_dev_free(d_band)

# the reference's own entry point, one model per call, host arrays in and out: what an unmodified
# BARTfunc.py worker pays per proposal (H2D profile, atm_prep, eclipse columns, D2H spectrum, sync)
for _ in range(20):
    tr.run_transit(prof[0])
t0 = time.perf_counter()
for k in range(reps):
    tr.run_transit(prof[k % nch])
us_legacy = 1e6 * (time.perf_counter() - t0) / reps

# BART's configured walk, DE-MC with snooker updates (MC3 walk='snooker'): the same measurement through
# bart_mcmc_run_snooker, random streams drawn up front like driver.run_snooker does
nfree = int(free.sum())
hsize = nch + 1
ds = driver.snooker_draws(np.random.RandomState(11), nch, nfree, ngen, hsize, 1, stepsize[free], pmin[free], pmax[free])
tr.mcmc_init(p0, pmin, pmax, stepsize, data, uncert)
tr.mcmc_snooker_init(ds["z0"], 1)
def run_snk(lo, hi):
    h = slice(lo, hi)
    off = np.asarray(ds["usn_offset"])[lo:hi + 1]
    tr.mcmc_run_snooker(ds["support"][h], ds["i1"][h], ds["i2"][h], ds["iz"][h], ds["ic"][h],
                        ds["usnooker"][off[0]:off[-1]], off - off[0], ds["unif"][h], ds["ugamma"][h])
run_snk(0, 100)
t0 = time.perf_counter()
run_snk(100, ngen)
us_snooker = 1e6 * (time.perf_counter() - t0) / (ngen - 100)

out = {"workload": "WASP-12b %s shape, %d chains, 9 free parameters (PT_line + 4 abundances)" % (a.solution, nch),
       "us_per_generation": us_gen, "generations_timed": ngen - 100,
       "accept_rate": float(np.sum(numaccept)) / (nch * ngen),
       "generation_kernels_us": gen_kernels,
       "forward_only_us_per_call": us_fwd, "forward_kernels_us": ks,
       "run_transit_us_per_call": us_legacy,
       "us_per_generation_snooker": us_snooker,
       "small_batch_kernel": os.environ.get("BART_ECL_SMALL", "auto"),
       "reference_note": "the reference evaluates a generation as 10 concurrent run_transit calls, one MPI "
                         "process per chain: 1 / cpu_baseline.single_thread_value seconds"}
print(json.dumps(out))
