#!/usr/bin/env python
"""oracle/ref_driver.py -- TEST INFRASTRUCTURE ONLY.

Drives the UNMODIFIED reference transit code (oracle/_ref/libtransit_ref.so, built by
oracle/Makefile from /root/reference sources) through the 8 exported C functions of
modules/transit/transit/src/transit.c:14-22 with ctypes, exactly as the SWIG module would.

The reference keeps all state in process globals and its option parser is one-shot
(pu/src/procopt.c:196-202), so every configuration needs its own process: this file is run as
a subprocess (`python oracle/ref_driver.py cfg models.npy out.npz [--inter] [--last] [--time K]`).

Outputs (npz): wn[nwave], spectra[M,nwave]; with --inter also, for every model, radius,
density, ext[layer,wn], cia[wn,layer], tau[wn,depth], last[wn] read from the reference's
globals through oracle/ref_shim.c; with --last only last[wn] (bench.py's in-run parity gate).
"""
import ctypes as C
import os
import sys
import time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libtransit_ref.so")

dp = C.POINTER(C.c_double)


def load():
    lib = C.CDLL(LIB)
    lib.transit_init.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    lib.get_no_samples.restype = C.c_int
    lib.get_waveno_arr.argtypes = [dp, C.c_int]
    lib.set_radius.argtypes = [C.c_double]
    lib.set_cloudtop.argtypes = [C.c_double]
    lib.set_scattering.argtypes = [C.c_int, C.c_double]
    lib.run_transit.argtypes = [dp, C.c_int, dp, C.c_int]
    lib.ref_run_keep.argtypes = [dp, dp]
    for name in ("ref_radius", "ref_temp", "ref_press", "ref_mm", "ref_ext", "ref_cia",
                 "ref_tau", "ref_angles", "ref_op_temp", "ref_adop", "ref_alor", "ref_owns"):
        getattr(lib, name).restype = dp
    lib.ref_density.restype = dp
    lib.ref_density.argtypes = [C.c_int]
    lib.ref_abund.restype = dp
    lib.ref_abund.argtypes = [C.c_int]
    lib.ref_last.restype = C.POINTER(C.c_long)
    for name in ("ref_nlayers", "ref_nwave", "ref_nmol", "ref_nowns"):
        getattr(lib, name).restype = C.c_long
    lib.ref_toomuch.restype = C.c_double
    lib.ref_radfct.restype = C.c_double
    lib.ref_prof_size.restype = C.c_long
    lib.ref_prof_size.argtypes = [C.c_int, C.c_int]
    lib.ref_prof.restype = C.POINTER(C.c_float)
    lib.ref_prof.argtypes = [C.c_int, C.c_int]
    return lib


def arr(ptr, n, dtype=np.float64):
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def init(lib, cfg, extra=()):
    args = [b"transit", b"-c", cfg.encode()] + [e.encode() for e in extra]
    argv = (C.c_char_p * (len(args) + 1))(*args, None)
    lib.transit_init(len(args), argv)


def main():
    cfg, models_path, out_path = sys.argv[1:4]
    flags = sys.argv[4:]
    inter = "--inter" in flags
    want_last = "--last" in flags
    ntime = int(flags[flags.index("--time") + 1]) if "--time" in flags else 0
    setters = {}
    for f in flags:
        if f.startswith("--radius="):
            setters["radius"] = float(f.split("=")[1])
        if f.startswith("--cloudtop="):
            setters["cloudtop"] = float(f.split("=")[1])
        if f.startswith("--scattering="):
            setters["scattering"] = float(f.split("=")[1])
        if f.startswith("--scatflag="):
            setters["scatflag"] = int(f.split("=")[1])
    lib = load()
    t0 = time.perf_counter()
    init(lib, cfg)
    t_init = time.perf_counter() - t0
    nwave = lib.get_no_samples()
    wn = np.zeros(nwave)
    lib.get_waveno_arr(wn.ctypes.data_as(dp), nwave)
    if "radius" in setters:
        lib.set_radius(setters["radius"])
    if "cloudtop" in setters:
        lib.set_cloudtop(setters["cloudtop"])
    if "scattering" in setters or "scatflag" in setters:
        lib.set_scattering(setters.get("scatflag", 1), setters.get("scattering", 0.0))
    models = np.ascontiguousarray(np.load(models_path), dtype=np.float64)
    if models.ndim == 1:
        models = models[None, :]
    M = models.shape[0]
    spectra = np.zeros((M, nwave))
    out = dict(wn=wn, t_init=t_init)
    nl = None
    inter_store = {k: [] for k in ("radius", "temp", "mm", "density", "ext", "cia", "tau",
                                   "last", "computed")}
    for m in range(M):
        spec = np.zeros(nwave)
        if inter:
            lib.ref_run_keep(models[m].ctypes.data_as(dp), spec.ctypes.data_as(dp))
            nl = lib.ref_nlayers()
            nmol = lib.ref_nmol()
            inter_store["radius"].append(arr(lib.ref_radius(), nl))
            inter_store["temp"].append(arr(lib.ref_temp(), nl))
            inter_store["mm"].append(arr(lib.ref_mm(), nl))
            inter_store["density"].append(np.stack([arr(lib.ref_density(j), nl)
                                                    for j in range(nmol)]))
            inter_store["ext"].append(arr(lib.ref_ext(), nl * nwave).reshape(nl, nwave))
            inter_store["cia"].append(arr(lib.ref_cia(), nl * nwave).reshape(nwave, nl))
            inter_store["tau"].append(arr(lib.ref_tau(), nl * nwave).reshape(nwave, nl))
            inter_store["last"].append(
                np.ctypeslib.as_array(lib.ref_last(), shape=(nwave,)).astype(np.int64))
        elif want_last:
            lib.ref_run_keep(models[m].ctypes.data_as(dp), spec.ctypes.data_as(dp))
            inter_store["last"].append(
                np.ctypeslib.as_array(lib.ref_last(), shape=(nwave,)).astype(np.int64))
        else:
            lib.run_transit(models[m].ctypes.data_as(dp), models.shape[1],
                            spec.ctypes.data_as(dp), nwave)
        spectra[m] = spec
    out["spectra"] = spectra
    if want_last and not inter:
        out["last"] = np.stack(inter_store["last"])
    if inter:
        for k, v in inter_store.items():
            if v:
                out[k] = np.stack(v)
        out["toomuch"] = lib.ref_toomuch()
        out["radfct"] = lib.ref_radfct()
    if ntime:
        spec = np.zeros(nwave)
        for m in range(min(2, M)):
            lib.run_transit(models[m].ctypes.data_as(dp), models.shape[1],
                            spec.ctypes.data_as(dp), nwave)
        t0 = time.perf_counter()
        for k in range(ntime):
            m = k % M
            lib.run_transit(models[m].ctypes.data_as(dp), models.shape[1],
                            spec.ctypes.data_as(dp), nwave)
        out["sec_per_model"] = (time.perf_counter() - t0) / ntime
    np.savez(out_path, **out)


if __name__ == "__main__":
    main()
