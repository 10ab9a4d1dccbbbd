#!/usr/bin/env python
"""Per-generation latency of the device-resident DE-MC loop on `--world` GPUs (one process each),
with the fused band-integration + peer-window all-gather (BART_P2P=1) and with ncclAllGather
(BART_P2P=0).  usage: bench_mcmc_multi.py [--world 2] [--gens 2000]"""
import argparse, json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import numpy as np
import test_gpu_multirank as t

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=2)
ap.add_argument("--gens", type=int, default=2000)
a = ap.parse_args()
os.environ["BART_TIME_GENS"] = str(a.gens)
res = {"world": a.world, "generations": a.gens - 50, "case": "retr_tiny_eclipse (6 chains, 201 wn x 100 layers)"}
ref = None
for p2p in (1, 0):
    with tempfile.TemporaryDirectory() as tmp:
        outs = t.run_world("time", tmp, a.world, p2p)
    us = max(float(o["us_per_generation"]) for o in outs)
    key = "fused_peer_window" if p2p else "nccl_allgather"
    res[key] = {"us_per_generation": us, "p2p_mapped": int(outs[0]["p2p"])}
    if ref is None:
        ref = outs[0]["params"]
    else:
        res["identical_chains"] = bool(np.array_equal(ref, outs[0]["params"]))
print(json.dumps(res))
