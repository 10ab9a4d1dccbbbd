"""GPU (B200): `savefiles yes` (tau.c:179-190,308-329) -- run_transit writes the reference's six
text dumps into the working directory; tau.dat is what code/cf.py:68-135 reads for the contribution
functions.  Compared with the files the UNMODIFIED reference wrote for the same model
(tests/golden/savefiles.npz)."""
import os
import numpy as np
import pytest

import cases
from util import relerr, parse_dump, DUMPS

pytestmark = pytest.mark.gpu


def test_savefiles_dumps_match_reference(built, workdir, monkeypatch):
    from bart_b200 import api
    case, models = cases.build_savefiles_case(workdir)
    g = np.load(cases.golden_path("savefiles"))
    assert cases.sha(models) == str(g["models_sha"]) and cases.sha(case["grid"]) == str(g["grid_sha"])
    monkeypatch.chdir(case["workdir"])
    for n in DUMPS:
        if os.path.exists(n):
            os.remove(n)
    tr = api.Transit(case["cfg"])
    spec = tr.run_transit(models[0])
    assert relerr(spec, g["spectra"][0]) < 1e-6
    got = {}
    for n in DUMPS:
        key = n.split(".")[0]
        assert os.path.exists(n), n
        keys, rows = parse_dump(n)
        got[key] = rows
        assert np.array_equal(keys, g[key + "_keys"]) or relerr(keys, g[key + "_keys"]) < 2e-9
        assert rows.shape == g[key].shape
        with open(n) as f:
            head = f.read(300)
        # identical header block and first record label (layout of print2dArrayDouble / save1Darray)
        ref_head = str(g[key + "_head"])
        nhead = ref_head.index("\n", ref_head.index(":")) + 1
        assert head[:nhead] == ref_head[:nhead], (head[:nhead], ref_head[:nhead])
    # values: 10 significant digits are printed; 2e-9 = one unit in the last printed digit
    tau, tau_ref = got["tau"], g["tau"]
    assert np.array_equal(tau > 0, tau_ref > 0)          # zero beyond `last`, like the reference
    assert relerr(tau, tau_ref) < 1e-8
    assert relerr(got["CIA"], g["CIA"]) < 2e-9
    assert relerr(got["cloud_extion"], g["cloud_extion"]) < 2e-9
    assert relerr(got["scatt_extion"], g["scatt_extion"]) < 2e-9
    mol_ref = g["mol_extion"]
    comp = np.abs(mol_ref).sum(axis=1) > 0               # the reference computes layers lazily
    assert comp.sum() > 5
    assert relerr(got["mol_extion"][comp], mol_ref[comp]) < 2e-9
    assert relerr(got["total_extion"][:, comp], g["total_extion"][:, comp]) < 2e-9
    # what cf.py does with tau.dat (cf.py:68-96): rows are wavenumbers, transposed to [layer][wn]
    assert got["tau"].T.shape == (tr.nlayer, tr.nwave)
    tr.free_memory()


def test_contribution_function_run_like_cf_py(built, workdir):
    """code/cf.py:40-65 re-runs the best-fit configuration through the transit EXECUTABLE with
    `toomuch 1e100` (no early exit: every layer integrated) and `savefiles yes`, then reads tau.dat
    (cf.py:68-96).  Same run with bart_b200/bin/transit against the reference executable's files."""
    import subprocess
    case = cases.build_savefiles_cf_case(workdir)
    g = np.load(cases.golden_path("savefiles_cf"))
    assert cases.sha(case["grid"]) == str(g["grid_sha"])
    for n in DUMPS + ("outspec.dat",):
        p = os.path.join(case["workdir"], n)
        if os.path.exists(p):
            os.remove(p)
    exe = os.path.join(cases.ROOT, "bart_b200", "bin", "transit")
    r = subprocess.run([exe, "-c", case["cfg"]], capture_output=True, text=True, cwd=case["workdir"])
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    keys, tau = parse_dump(os.path.join(case["workdir"], "tau.dat"))
    assert np.array_equal(keys, g["tau_keys"]) and tau.shape == g["tau"].shape
    assert (tau[:, 1:] > 0).all()                            # toomuch 1e100: down to the bottom layer
    assert relerr(tau, g["tau"]) < 1e-8
    # cf.readTauDat's view: tau[layer][wn] after the transpose
    assert tau.T.shape == (case["nlayer"], len(case["wn"]))
    ref = str(g["outspec"]).splitlines()
    with open(os.path.join(case["workdir"], "outspec.dat")) as f:
        mine = f.read().splitlines()
    assert mine[0] == ref[0] and len(mine) == len(ref)
    worst = max(abs(float(a[15:]) - float(b[15:])) / abs(float(b[15:])) for a, b in zip(mine[1:], ref[1:]))
    assert worst < 2e-9
