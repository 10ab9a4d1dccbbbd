#!/usr/bin/env python
"""bench.py -- forward-model spectra/s on the WASP-12b eclipse shape (BASELINE.json metric).

A step = one DE-MC generation on this rank: M proposal models -> spectra -> band fluxes
(atm_prep + fused eclipse column kernel + band integration), plus, for N>1, the per-generation
all-gather of the band fluxes over NCCL.  `value` is measured with the proposals already
resident in HBM; `e2e` goes through the reference-facing batched call with pinned HOST buffers
(profiles H2D, spectra D2H inside the timed region).  Timing: CUDA events on the library's own
stream, max over ranks.  `--impl reference` times the unmodified reference C (oracle/_ref; else
the oracle port) on the host cores for the same configuration.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "forward_model_spectra_per_s"
UNIT = "spectra/s"
WORKLOAD = ("WASP-12b eclipse (examples/WASP-12b/BART.cfg shape): 2424 wavenumbers (910-3333 cm-1) "
            "x 100 layers x 27 grid temperatures x 4 molecules (H2O CO2 CO CH4) + H2-H2 CIA, "
            "5 ray angles, toomuch 10, 4 filters")


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ---------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.proc = None
        self.lines = []

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.lines.append(line.strip())
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)),
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------
def make_workload(tmp, rank, M):
    from bart_b200 import synth
    case = synth.make_case(os.path.join(tmp, "w12_rank%d" % rank), shape="w12", solution="eclipse",
                           seed=2026)
    models = synth.make_models(case, M, seed=2026 + rank, molfit=("H2O", "CO2", "CO", "CH4"))
    return case, models


def star_planck(wn, tstar=6300.0):
    hc_k = 6.6260755e-27 * 2.99792458e10 / 1.380658e-16
    return 2 * 6.6260755e-27 * 2.99792458e10 ** 2 * wn ** 3 / np.expm1(hc_k * wn / tstar) * np.pi


def reference_rate(case, models, nproc, per_proc, timeout=900):
    """Time the CPU implementation of run_transit with `nproc` independent processes (MC3's own
    one-process-per-chain model, mccubed.py:277,393).  Returns (models/s, kind, cores)."""
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libtransit_ref.so")
    workdir = case["workdir"]
    if os.path.exists(ref_so):
        procs = []
        for p in range(nproc):
            mp = os.path.join(workdir, "ref_models_%d.npy" % p)
            np.save(mp, models[(p * per_proc) % len(models):][:per_proc] if len(models) >= per_proc
                    else models)
            op = os.path.join(workdir, "ref_out_%d.npz" % p)
            procs.append((subprocess.Popen(
                [sys.executable, os.path.join(ROOT, "oracle", "ref_driver.py"), case["cfg"], mp, op,
                 "--time", str(per_proc)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL), op))
        rate = 0.0
        for pr, op in procs:
            pr.wait(timeout=timeout)
            if pr.returncode == 0 and os.path.exists(op):
                rate += 1.0 / float(np.load(op)["sec_per_model"])
        return rate, "reference", nproc
    # the oracle port (single-threaded restatement), one process per core via fork
    from oracle import oracle as orc
    O = orc.Oracle(case["cfg"])
    O.run(models[0])
    t0 = time.perf_counter()
    n = min(per_proc, len(models))
    for m in range(n):
        O.run(models[m])
    return n / (time.perf_counter() - t0), "port", 1


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    nproc = max(1, min(ncores, 64))
    with tempfile.TemporaryDirectory(prefix="bart_bench_ref_") as tmp:
        case, models = make_workload(tmp, 0, 64)
        per_proc = 6
        for _ in range(args.warmup and 1):
            reference_rate(case, models, nproc, 2)
        t0 = time.perf_counter()
        rates = []
        for _ in range(args.steps):
            r, kind, cores = reference_rate(case, models, nproc, per_proc)
            rates.append(r)
        wall = time.perf_counter() - t0
    value = float(np.mean(rates))
    sample = "%d processes x %d models each per step (run_transit wall time after 2 warm-ups)" % (nproc, per_proc)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "models_per_step": nproc * per_proc},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--models", type=int, default=4096, help="proposal models per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from bart_b200 import api
    L = api.lib()
    M = args.models
    W = max(3, args.warmup)
    tmp = tempfile.mkdtemp(prefix="bart_bench_")
    case, models = make_workload(tmp, rank, M)
    tr = api.Transit(case["cfg"], device=local)
    info = api.device_info()
    wn = tr.get_waveno_arr()
    start, count, weight, star = api.filters_from_files(wn, case["filters"], wn, star_planck(wn))
    tr.set_filters(start, count, weight, star, 0.117)
    nf, nw, n_in = tr.nfilters, tr.nwave, tr.n_in

    # multi-GPU exchange: library-owned NCCL communicator, unique id through torch.distributed
    if world > 1:
        import ctypes as C
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            api._check(L.bart_comm_unique_id(idbuf))
        obj = [idbuf.raw]
        dist.broadcast_object_list(obj, src=0)
        api._check(L.bart_comm_init(rank, world, obj[0]))

    # device-resident proposals and results
    d_prof = L.bart_dev_alloc(M * n_in * 8)
    d_band = L.bart_dev_alloc(M * nf * 8)
    d_all = L.bart_dev_alloc(world * M * nf * 8)
    api._check(L.bart_memcpy_h2d(d_prof, models.ctypes.data, M * n_in * 8))

    def step_device():
        if world > 1:
            # band integration fused with the all-gather (stores into the peers' NVLink windows;
            # ncclAllGather when the windows could not be mapped)
            api._check(L.bart_bandflux_allgather_device(d_prof, M, n_in, d_band, d_all))
        else:
            api._check(L.bart_bandflux_batch_device(d_prof, M, n_in, d_band, None))

    def barrier():
        L.bart_sync()
        if dist is not None:
            dist.barrier()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(W):
        step_device()
    barrier()
    L.bart_profile_reset()
    L.bart_profile_enable(1)
    launches0 = L.bart_launch_count()
    ms = []
    for _ in range(args.steps):
        L.bart_flush_l2()                      # untimed: evict the L2 between steps
        barrier()
        L.bart_timer_begin()
        step_device()
        ms.append(L.bart_timer_end())
    barrier()
    launches = L.bart_launch_count() - launches0                   # L2-flush fills are not counted
    L.bart_profile_enable(0)
    stats = api.kernel_stats()
    total_ms = float(np.sum(ms))

    # end to end through the reference-facing batched call: pinned host in/out
    h_in = api.PinnedArray((M, n_in))
    h_out = api.PinnedArray((M, nw))
    h_in.array[:] = models
    st = np.zeros(M, dtype=np.int32)
    for _ in range(2):
        tr.run_batch(h_in.array, out=h_out.array, status=st)
    e2e_ms = []
    for _ in range(args.steps):
        L.bart_flush_l2()
        barrier()
        t0 = time.perf_counter()
        tr.run_batch(h_in.array, out=h_out.array, status=st)
        e2e_ms.append(1e3 * (time.perf_counter() - t0))
    e2e_total = float(np.sum(e2e_ms))
    # nvidia-smi can take longer to start than a short run lasts (8 ranks on one box): keep the
    # same step running, untimed, until the sampler has seen the GPU under this load
    t_wait = time.time()
    while len(sampler.lines) < 5 and time.time() - t_wait < 8.0:
        # rank-local work only: the number of iterations differs between ranks
        api._check(L.bart_bandflux_batch_device(d_prof, M, n_in, d_band, None))
    clocks = sampler.stop()

    if dist is not None:
        import torch
        t = torch.tensor([total_ms, e2e_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_total = float(t[0]), float(t[1])

    if rank == 0:
        value = world * M * args.steps / (total_ms * 1e-3)
        e2e_value = world * M * args.steps / (e2e_total * 1e-3)
        # roofline of the dominant kernel (SURVEY.md section 8d accounting)
        nmol, nlayer = L.bart_ngridmol(), tr.nlayer
        alg_bytes = M * (16.0 * nmol * nlayer * nw + 8.0 * nw)
        dom = "eclipse_column"
        k = stats.get(dom, {"launches": 0, "ms": 0.0})
        peak, peak_src = 6650.0, "fallback"
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            peak_src = "measured"
        except Exception:
            pass
        achieved = alg_bytes * k["launches"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else None
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "eclipse_column_traffic.json")))
            if prof.get("models_per_launch") == M:
                traffic = prof.get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak,
                    "peak_source": peak_src, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "kernel_ms_per_launch": (k["ms"] / k["launches"]) if k["launches"] else None,
                    "kernel_share_of_step": (k["ms"] / total_ms) if total_ms else None,
                    "note": "algorithmic bytes counted without credit for the toomuch early exit or "
                            "for L2/L1 reuse of grid planes across models; >1 means cache reuse"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ncores = os.cpu_count() or 1
            nproc = max(1, min(ncores, 64))
            rate, kind, cores = reference_rate(case, models, nproc, 6)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": "%d processes x 6 models of the same batch (run_transit wall time after 2 "
                             "warm-ups; reference writes its output spectrum to /dev/null)" % nproc}
        # the other named throughput of the north star: the --justOpacity grid builder, on a bounded
        # sample (its own process: the library holds one configuration per process)
        builder = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_builder.py")],
                                   capture_output=True, text=True, timeout=300)
                b = json.loads(r.stdout.strip().splitlines()[-1])
                builder = {"metric": "opacity_grid_builder_line_cells_per_s",
                           "value": b["line_cells_per_s_device"], "unit": "lines x (T,layer) cells / s",
                           "wall_value": b["line_cells_per_s_wall"],
                           "sample": "%d synthetic lines x %d layers x %d temperature planes on the W12 "
                                     "wavenumber grid (%d samples, wnosamp %d)" % (
                                         b["nlines_in_range"], b["shape"]["nlayer"], b["shape"]["ntemp_built"],
                                         b["shape"]["nwave"], b["shape"]["wnosamp"]),
                           "per_slice_ms": b["per_slice_ms"]}
            except Exception as e:                    # the headline line must not depend on it
                builder = {"error": repr(e)[:200]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": W, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "models_per_gpu_per_step": M,
                           "parallelism": "chains partitioned by rank (dp%d); per-generation all-gather "
                                          "of band fluxes %s" % (world, "fused into the band-integration "
                                          "kernel (stores into peer windows over NVLink)"
                                          if world > 1 and L.bart_comm_p2p() else "(ncclAllGather)"
                                          if world > 1 else "(single rank: none)"),
                           "l2": "flushed between steps (256 MB write); grid 209 MB > L2",
                           "device": info["name"], "sm_count": info["sm_count"]},
                "roofline": roofline, "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": M * n_in * 8,
                        "d2h_bytes_per_step": M * nw * 8 + M * 4,
                        "path": "bart_run_batch: pinned host profiles -> H2D -> kernels -> D2H spectra"},
                "gpu_launches": int(launches), "kernels": stats, "clocks": clocks}
        if builder is not None:
            line["builder"] = builder
        print(json.dumps(line))
    if dist is not None:
        L.bart_comm_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
