#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libtransit_ref.so,
compiled from /root/reference by oracle/Makefile) on the synthetic configurations of
tests/cases.py.  Run in the build container (the reference sources do not exist on the GPU box):

    python tests/golden/make_golden.py

Each file stores the reference's spectra, last[] (index where tau exceeds toomuch), hydrostatic
radii and a strided sample of tau, plus sha256 digests of the generated inputs."""
import os
import subprocess
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402


def main():
    only = sys.argv[1:]
    with tempfile.TemporaryDirectory() as tmp:
        for name in cases.CASES:
            if only and name not in only:
                continue
            case, models, setters = cases.build_case(name, tmp)
            mpath = os.path.join(case["workdir"], "models.npy")
            opath = os.path.join(case["workdir"], "ref.npz")
            np.save(mpath, models)
            cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_driver.py"), case["cfg"], mpath,
                   opath, "--inter"] + ["--%s=%r" % (k, v) for k, v in setters.items()]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise SystemExit("reference failed on %s:\n%s\n%s" % (name, r.stdout[-2000:], r.stderr[-2000:]))
            d = np.load(opath)
            nw = len(d["wn"])
            wsel = np.unique(np.linspace(0, nw - 1, 16).astype(int))
            np.savez_compressed(
                cases.golden_path(name), wn=d["wn"], spectra=d["spectra"], last=d["last"],
                radius=d["radius"], tau_sample=d["tau"][:, wsel, :], tau_wsel=wsel,
                ext_sample=d["ext"][:, :, wsel], cia_sample=d["cia"][:, wsel, :],
                grid_sha=cases.sha(case["grid"]), models_sha=cases.sha(models),
                toomuch=d["toomuch"])
            print("%-24s nwave %5d  models %d  last[min,max]=%d,%d  -> %s (%d bytes)" % (
                name, nw, models.shape[0], d["last"].min(), d["last"].max(),
                os.path.basename(cases.golden_path(name)), os.path.getsize(cases.golden_path(name))))


def builder_golden(only):
    """Grids built by the reference's `transit --justOpacity` from synthetic TLI files."""
    from bart_b200 import synth
    exe = os.path.join(ROOT, "oracle", "_ref", "transit_ref")
    with tempfile.TemporaryDirectory() as tmp:
        for name in cases.BUILD_CASES:
            if only and name not in only:
                continue
            case = cases.build_builder_case(name, tmp)
            r = subprocess.run([exe, "-c", case["cfg"], "--justOpacity"], capture_output=True, text=True)
            if r.returncode != 0 or not os.path.exists(case["opacity"]):
                raise SystemExit("reference builder failed on %s:\n%s\n%s" % (name, r.stdout[-2000:], r.stderr[-2000:]))
            g = synth.read_opacity(case["opacity"], mmap=False)
            np.savez_compressed(cases.golden_path(name), grid=g["o"], temps=g["temps"], molids=g["molids"],
                                press=g["press"], wn=g["wn"], tli_sha=cases.sha(np.fromfile(case["tli"], dtype=np.uint8)),
                                file_bytes=os.path.getsize(case["opacity"]))
            print("%-24s grid %s  max %.3g  nonzero %.3f -> %d bytes" % (
                name, g["o"].shape, g["o"].max(), (g["o"] > 0).mean(), os.path.getsize(cases.golden_path(name))))


def lbl_golden(only):
    """Forward models of the reference WITHOUT an opacity file (line-by-line extinction at every
    layer's own temperature, tau.c:163-175 -> computemolext(permol=0))."""
    with tempfile.TemporaryDirectory() as tmp:
        for name in cases.LBL_CASES:
            if only and name not in only:
                continue
            case, models = cases.build_lbl_case(name, tmp)
            mpath = os.path.join(case["workdir"], "models.npy")
            opath = os.path.join(case["workdir"], "ref.npz")
            np.save(mpath, models)
            cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_driver.py"), case["cfg"], mpath,
                   opath, "--inter"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise SystemExit("reference failed on %s:\n%s\n%s" % (name, r.stdout[-2000:], r.stderr[-2000:]))
            d = np.load(opath)
            np.savez_compressed(
                cases.golden_path(name), wn=d["wn"], spectra=d["spectra"], last=d["last"],
                radius=d["radius"], ext=d["ext"], tau=d["tau"],
                tli_sha=cases.sha(np.fromfile(case["tli"], dtype=np.uint8)),
                models_sha=cases.sha(models), toomuch=d["toomuch"])
            print("%-24s nwave %5d  models %d  last[min,max]=%d,%d  ext max %.3g -> %d bytes" % (
                name, len(d["wn"]), models.shape[0], d["last"].min(), d["last"].max(), d["ext"].max(),
                os.path.getsize(cases.golden_path(name))))


def wl_golden(only):
    """Wavenumber grids the reference derives from wavelength limits (get_waveno_arr)."""
    if only and "wl" not in only:
        return
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for k in range(len(cases.WL_CASES)):
            case = cases.build_wl_case(k, tmp)
            from bart_b200 import synth
            models = synth.make_models(case, 1, seed=6)
            mpath, opath = os.path.join(case["workdir"], "m.npy"), os.path.join(case["workdir"], "o.npz")
            np.save(mpath, models)
            r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_driver.py"), case["cfg"], mpath, opath],
                               capture_output=True, text=True)
            if r.returncode != 0:
                raise SystemExit("reference failed:\n%s\n%s" % (r.stdout[-2000:], r.stderr[-2000:]))
            d = np.load(opath)
            out["wn%d" % k], out["spec%d" % k] = d["wn"], d["spectra"][0]
    np.savez_compressed(cases.golden_path("wl_ranges"), **out)
    print("wl_ranges:", {k: v.shape for k, v in out.items()})


def cli_golden(only):
    """Spectrum files written by the reference's command-line program (printflux eclipse.c:355-380,
    printmod slantpath.c:510-555) for the atmosphere file's own profiles."""
    exe = os.path.join(ROOT, "oracle", "_ref", "transit_ref")
    with tempfile.TemporaryDirectory() as tmp:
        for name in cases.CLI_CASES:
            if only and name not in only:
                continue
            case = cases.build_cli_case(name, tmp)
            out = os.path.join(case["workdir"], "outspec.dat")
            r = subprocess.run([exe, "-c", case["cfg"]], capture_output=True, text=True, cwd=case["workdir"])
            if r.returncode != 0 or not os.path.exists(out):
                raise SystemExit("reference CLI failed on %s:\n%s\n%s" % (name, r.stdout[-2000:], r.stderr[-2000:]))
            with open(out) as f:
                text = f.read()
            np.savez_compressed(cases.golden_path(name), text=text, grid_sha=cases.sha(case["grid"]))
            print("%-24s %d lines: %r ..." % (name, text.count("\n"), text[:60]))


def savefiles_cf_golden(only):
    """tau.dat of the reference EXECUTABLE for cf.py's configuration (toomuch 1e100, savefiles yes)."""
    if only and "savefiles_cf" not in only:
        return
    from util import parse_dump
    exe = os.path.join(ROOT, "oracle", "_ref", "transit_ref")
    with tempfile.TemporaryDirectory() as tmp:
        case = cases.build_savefiles_cf_case(tmp)
        r = subprocess.run([exe, "-c", case["cfg"]], capture_output=True, text=True, cwd=case["workdir"])
        if r.returncode != 0:
            raise SystemExit("reference CLI failed:\n%s\n%s" % (r.stdout[-2000:], r.stderr[-2000:]))
        keys, tau = parse_dump(os.path.join(case["workdir"], "tau.dat"))
        with open(os.path.join(case["workdir"], "outspec.dat")) as f:
            spec = f.read()
        np.savez_compressed(cases.golden_path("savefiles_cf"), tau_keys=keys, tau=tau, outspec=spec,
                            grid_sha=cases.sha(case["grid"]))
        print("savefiles_cf: tau %s, all layers integrated: %s" % (tau.shape, bool((tau[:, 1:] > 0).all())))


def savefiles_golden(only):
    """The six `savefiles yes` text dumps written by the reference for one model, parsed."""
    if only and "savefiles" not in only:
        return
    from util import parse_dump, DUMPS
    with tempfile.TemporaryDirectory() as tmp:
        case, models = cases.build_savefiles_case(tmp)
        mpath = os.path.join(case["workdir"], "models.npy")
        opath = os.path.join(case["workdir"], "ref.npz")
        np.save(mpath, models)
        cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_driver.py"), case["cfg"], mpath, opath]
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=case["workdir"])
        if r.returncode != 0:
            raise SystemExit("reference failed:\n%s\n%s" % (r.stdout[-2000:], r.stderr[-2000:]))
        out = dict(models_sha=cases.sha(models), grid_sha=cases.sha(case["grid"]),
                   spectra=np.load(opath)["spectra"])
        for name in DUMPS:
            path = os.path.join(case["workdir"], name)
            keys, rows = parse_dump(path)
            key = name.split(".")[0]
            out[key + "_keys"], out[key] = keys, rows
            with open(path) as f:
                out[key + "_head"] = f.read(300)
        np.savez_compressed(cases.golden_path("savefiles"), **out)
        print("savefiles: tau %s, mol %s -> %d bytes" % (out["tau"].shape, out["mol_extion"].shape,
                                                          os.path.getsize(cases.golden_path("savefiles"))))


if __name__ == "__main__":
    wl_golden(sys.argv[1:])
    cli_golden(sys.argv[1:])
    savefiles_cf_golden(sys.argv[1:])
    savefiles_golden(sys.argv[1:])
    builder_golden(sys.argv[1:])
    lbl_golden(sys.argv[1:])
    main()
