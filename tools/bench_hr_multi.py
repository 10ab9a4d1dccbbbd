#!/usr/bin/env python
"""High-resolution sweep (BASELINE.json configs[4]) on N GPUs: every GPU holds its own replica of the
6.4 GB grid (1e5 wavenumbers x 100 layers x 20 T x 4 molecules) and evaluates --models proposal
models per generation with the fused eclipse kernel; no data-path collective (forward models are
independent), so the aggregate is the sum over ranks at the slowest rank's step time.
usage: bench_hr_multi.py [--gpus 8] [--nwave 100001] [--models 1000]"""
import argparse, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=8)
ap.add_argument("--nwave", type=int, default=100001)
ap.add_argument("--models", type=int, default=1000)
a = ap.parse_args()
t0 = time.time()
procs = []
for r in range(a.gpus):
    env = dict(os.environ, BART_DEVICE=str(r))
    env.pop("LOCAL_RANK", None)
    procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tools", "bench_lookup.py"), "--nwave",
                                   str(a.nwave), "--models", "1", "--fused-models", str(a.models), "--steps", "5"],
                                  env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
ms = [o["fused_eclipse"]["ms_per_launch"] for o in outs]
step = [a.models / o["fused_eclipse"]["spectra_per_s"] * 1e3 for o in outs]
print(json.dumps({"n_gpus": a.gpus, "shape": outs[0]["shape"], "models_per_gpu_per_generation": a.models,
                  "eclipse_kernel_ms_per_rank": ms, "step_ms_per_rank": step,
                  "spectra_per_s_aggregate": a.gpus * a.models / (max(step) * 1e-3),
                  "spectra_per_s_per_rank": [o["fused_eclipse"]["spectra_per_s"] for o in outs],
                  "lookup_m1_frac_of_hbm_peak_per_rank": [o["lookup"][0]["frac_of_hbm_peak"] for o in outs],
                  "elapsed_incl_grid_generation_s": time.time() - t0}))
