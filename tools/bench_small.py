#!/usr/bin/env python
"""Eclipse kernel choice for small batches at the WASP-12b shape: device time of the forward call
(profiles on the device -> spectra) for M models under each kernel ($BART_ECL_SMALL = 0 throughput,
1 slot, 2 scan), per-kernel times from the library's CUDA-event profile, and the largest relative
difference of the scan kernel's spectra from the throughput kernel's.
usage: bench_small.py [--models 1,2,4,10,16,24,32,48,64]"""
import argparse, json, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bart_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--models", default="1,2,4,10,16,24,32,48,64")
ap.add_argument("--reps", type=int, default=200)
a = ap.parse_args()
tmp = tempfile.mkdtemp(prefix="bart_small_")
case = synth.make_case(tmp, shape="w12", solution="eclipse", seed=2026)
tr = api.Transit(case["cfg"])
L = api.lib()
out = {"workload": "WASP-12b eclipse shape, eclipse kernel per batch size", "rows": []}
for M in [int(x) for x in a.models.split(",")]:
    models = synth.make_models(case, M, seed=11, molfit=("H2O", "CO2", "CO", "CH4"))
    d_prof = L.bart_dev_alloc(models.size * 8)
    d_spec = L.bart_dev_alloc(M * tr.nwave * 8)
    L.bart_memcpy_h2d(d_prof, models.ctypes.data, models.size * 8)
    row = {"models": M}
    spectra = {}
    for mode in (0, 1, 2):
        os.environ["BART_ECL_SMALL"] = str(mode)
        for _ in range(10):
            api._check(L.bart_run_batch_device(d_prof, M, tr.n_in, d_spec, tr.nwave, None))
        L.bart_sync()
        L.bart_profile_reset()
        L.bart_profile_enable(1)
        for _ in range(a.reps):
            api._check(L.bart_run_batch_device(d_prof, M, tr.n_in, d_spec, tr.nwave, None))
        L.bart_sync()
        L.bart_profile_enable(0)
        ks = api.kernel_stats()
        row["us_mode%d" % mode] = 1e3 * ks["eclipse_column"]["ms"] / ks["eclipse_column"]["launches"]
        sp = np.empty((M, tr.nwave))
        L.bart_memcpy_d2h(sp.ctypes.data, d_spec, sp.size * 8)
        spectra[mode] = sp
    row["slot_identical"] = bool(np.array_equal(spectra[0], spectra[1]))
    row["scan_max_rel_diff"] = float(np.max(np.abs(spectra[2] - spectra[0]) / np.abs(spectra[0])))
    out["rows"].append(row)
    L.bart_dev_free(d_prof); L.bart_dev_free(d_spec)
os.environ.pop("BART_ECL_SMALL", None)
print(json.dumps(out))
