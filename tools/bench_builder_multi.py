#!/usr/bin/env python
"""Opacity-grid build sharded over the temperature axis on N GPUs (one process per GPU, no
collective: every rank builds its planes).  Each rank runs tools/bench_builder.py's
sample on its own device for its share of a --ntemp-total temperature grid; the aggregate is the
total line x (T,layer) cells over the slowest rank's wall time.
usage: bench_builder_multi.py [--gpus 8] [--ntemp-total 16] [--nlines N] [--wndelt D]"""
import argparse, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=8)
ap.add_argument("--ntemp-total", type=int, default=16)
ap.add_argument("--nlines", type=int, default=2400000)
ap.add_argument("--wndelt", type=float, default=1.0)
ap.add_argument("--full", action="store_true",
                help="BASELINE configs[3] as stated: a grid of --ntemp-total temperatures, rank r builds planes "
                     "r, r + gpus, ... for all layers (every rank generates and loads the same seeded line list)")
a = ap.parse_args()
share = [a.ntemp_total // a.gpus + (1 if r < a.ntemp_total % a.gpus else 0) for r in range(a.gpus)]
procs = []
t0 = time.time()
for r in range(a.gpus):
    if share[r] == 0:
        continue
    # one process per GPU, each seeing only its own device (CUDA initialisation then does not walk
    # all eight GPUs in every process)
    env = dict(os.environ, BART_DEVICE="0", CUDA_VISIBLE_DEVICES=str(r))
    env.pop("LOCAL_RANK", None)
    extra = []
    if a.full:
        extra = ["--ntemp-grid", str(a.ntemp_total), "--planes", ",".join(str(t) for t in range(r, a.ntemp_total, a.gpus))]
    procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tools", "bench_builder.py"),
                                   "--ntemp", str(share[r]), "--nlines", str(a.nlines), "--wndelt", str(a.wndelt)] + extra,
                                  env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
t_all = time.time() - t0
cells = sum(o["nlines_in_range"] * o["shape"]["nlayer"] * o["shape"]["ntemp_built"] for o in outs)
wall = max(o["wall_s"] for o in outs)
dev = max(o["device_ms"] for o in outs) * 1e-3
print(json.dumps({"n_gpus": a.gpus, "planes_per_rank": share, "nlines": outs[0]["nlines_in_range"],
                  "nwave": outs[0]["shape"]["nwave"], "nlayer": outs[0]["shape"]["nlayer"],
                  "line_cells_total": cells, "slowest_rank_wall_s": wall, "slowest_rank_device_s": dev,
                  "line_cells_per_s_wall": cells / wall, "line_cells_per_s_device": cells / dev,
                  "per_rank_line_cells_per_s_device": [o["line_cells_per_s_device"] for o in outs],
                  "per_rank_wall_s": [o["wall_s"] for o in outs], "per_rank_gen_s": [o["gen_s"] for o in outs],
                  "per_rank_init_s": [o["init_s"] for o in outs], "per_rank_cuda_init_s": [o.get("cuda_init_s") for o in outs], "one_time_ms_rank0": outs[0]["one_time_ms"],
                  "per_slice_ms_rank0": outs[0]["per_slice_ms"], "full_config": bool(a.full),
                  "elapsed_incl_init_s": t_all}))
