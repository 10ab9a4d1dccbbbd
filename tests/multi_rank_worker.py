"""Worker of tests/test_gpu_multirank.py: one process per GPU (rank = argv[1]), NCCL unique id
exchanged through a file.  mode 'demc': the device-resident DE-MC loop with the chains partitioned
over the ranks; mode 'gather': forward models + band integration fused with the all-gather."""
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    rank, world, mode, workdir, outpath = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
    import ctypes as C
    import cases
    from bart_b200 import api, driver
    L = api.lib()
    name = "retr_small4_transit" if mode in ("demc_transit", "snooker_transit") else "retr_tiny_eclipse"
    case, spec, extra = cases.build_retrieval(name, os.path.join(workdir, "rank%d" % rank))
    tr = api.Transit(case["cfg"], device=rank)
    wn = tr.get_waveno_arr()
    start, count, weight, star = api.filters_from_files(wn, case["filters"], extra["starwn"], extra["starfl"])
    tr.set_filters(start, count, weight, star, extra["rprs"])
    tr.converter_init(case["press_bar"], case["species"], case["abund"], spec["molfit"], spec["pt"],
                      pt_args=extra["pt_args"], nrad=spec["nrad"], ncloud=spec["ncloud"], nray=spec["nray"])
    idfile = os.path.join(workdir, "nccl_id_%s.bin" % mode)
    if rank == 0:
        buf = C.create_string_buffer(128)
        api._check(L.bart_comm_unique_id(buf))
        with open(idfile + ".tmp", "wb") as f:
            f.write(buf.raw)
        os.rename(idfile + ".tmp", idfile)
    t0 = time.time()
    while not os.path.exists(idfile):
        if time.time() - t0 > 60:
            raise SystemExit("no NCCL id")
        time.sleep(0.05)
    with open(idfile, "rb") as f:
        uid = f.read()
    api._check(L.bart_comm_init(rank, world, uid))
    out = {"p2p": L.bart_comm_p2p()}
    if mode == "time":
        # generations per second of the device-resident loop (graph replay) at MC3's population size
        d = np.load(os.path.join(cases.GOLDEN_DIR, "retrieval_mc3_%s.npz" % name))
        nch, ngen = spec["nchains"], int(os.environ.get("BART_TIME_GENS", "2000"))
        stepsize = np.array(spec["stepsize"])
        rng = np.random.RandomState(5)
        p0 = np.repeat(np.atleast_2d(spec["params"]), nch, 0)
        p0[:, stepsize > 0] += rng.normal(0, 0.01, (nch, int((stepsize > 0).sum())))
        dr = driver.demc_draws(rng, nch, ngen, stepsize[stepsize > 0])
        tr.mcmc_init(p0, spec["pmin"], spec["pmax"], stepsize, d["data"], d["uncert"])
        warm = slice(0, 50)
        tr.mcmc_run(dr["support"][warm], dr["r1"][:, warm], dr["r2"][:, warm], dr["unif"][warm], dr["ugamma"][warm])
        rest = slice(50, ngen)
        t0 = time.perf_counter()
        tr.mcmc_run(dr["support"][rest], dr["r1"][:, rest], dr["r2"][:, rest], dr["unif"][rest], dr["ugamma"][rest])
        dt = time.perf_counter() - t0
        out.update(gens=ngen - 50, seconds=dt, us_per_generation=1e6 * dt / (ngen - 50),
                   params=tr.mcmc_get("params"))
    elif mode.startswith("snooker"):
        thinning = 3
        d = np.load(os.path.join(cases.GOLDEN_DIR, "retrieval_snooker_%s_thin%d.npz" % (name, thinning)))
        np.random.seed(spec["seed"] + thinning)
        r = driver.run_snooker(tr, d["data"], d["uncert"], spec["params"], spec["pmin"], spec["pmax"],
                               spec["stepsize"], spec["numit"], spec["nchains"], burnin=spec["burnin"],
                               thinning=thinning)
        out.update(allparams=r["allparams"], bestp=r["bestp"], numaccept=r["numaccept"], Z=r["Z"])
    elif mode.startswith("demc"):
        d = np.load(os.path.join(cases.GOLDEN_DIR, "retrieval_mc3_%s.npz" % name))
        np.random.seed(spec["seed"])
        r = driver.run_demc(tr, d["data"], d["uncert"], spec["params"], spec["pmin"], spec["pmax"],
                            spec["stepsize"], spec["numit"], spec["nchains"], burnin=spec["burnin"])
        out.update(allparams=r["allparams"], bestp=r["bestp"], numaccept=r["numaccept"], models=r["models"])
    else:
        # every rank evaluates its own M models; all ranks receive all blocks
        g = np.load(os.path.join(cases.GOLDEN_DIR, "retrieval_conv_%s.npz" % name))
        prof_all, status, _ = tr.profiles_from_params(g["params"])
        ok = np.where(status == 0)[0]
        M = len(ok) // world
        mine = np.ascontiguousarray(prof_all[ok[rank * M:(rank + 1) * M]])
        nf, n_in = tr.nfilters, tr.n_in
        d_prof = L.bart_dev_alloc(M * n_in * 8)
        d_band = L.bart_dev_alloc(M * nf * 8)
        d_all = L.bart_dev_alloc(world * M * nf * 8)
        api._check(L.bart_memcpy_h2d(d_prof, mine.ctypes.data, M * n_in * 8))
        gathered = []
        for rep in range(5):                                  # slot parity, flag reuse
            api._check(L.bart_bandflux_allgather_device(d_prof, M, n_in, d_band, d_all))
            a = np.zeros((world, M, nf))
            api._check(L.bart_memcpy_d2h(a.ctypes.data, d_all, a.nbytes))
            gathered.append(a)
        single, _ = tr.bandflux_batch(np.ascontiguousarray(prof_all[ok[:world * M]]))
        out.update(gathered=np.stack(gathered), single=single.reshape(world, M, nf))
    np.savez(outpath, **out)
    L.bart_comm_finalize()
    tr.free_memory()


if __name__ == "__main__":
    main()
