"""CPU (build container only): fresh seeds through the compiled reference and the oracle.
Skipped where oracle/_ref is absent."""
import os
import numpy as np
import pytest

import conftest
from util import relerr, apply_setters

pytestmark = pytest.mark.skipif(not conftest.has_ref(), reason="oracle/_ref not built here")


@pytest.mark.parametrize("solution,setters", [("eclipse", {}),
                                              ("transit", {"radius": 94200.0}),
                                              ("eclipse", {"cloudtop": -0.5, "scattering": 2.0})])
def test_fresh_seed(solution, setters, built, workdir):
    from bart_b200 import synth
    from oracle import oracle as orc
    tag = "fresh_%s_%d" % (solution, len(setters))
    case = synth.make_case(os.path.join(workdir, tag), shape="tiny", solution=solution, seed=31337,
                           refradius_km=95000.0 if solution == "transit" else 123820.0)
    models = synth.make_models(case, 2, seed=2718)
    mp, op = os.path.join(case["workdir"], "m.npy"), os.path.join(case["workdir"], "ref.npz")
    np.save(mp, models)
    conftest.run_reference(case["cfg"], mp, op, setters)
    d = np.load(op)
    O = orc.Oracle(case["cfg"])
    apply_setters(O, setters)
    for m in range(2):
        o = O.run(models[m], inter=True)
        assert np.array_equal(o["last"], d["last"][m])
        assert relerr(o["spectrum"], d["spectra"][m]) < 1e-9


@pytest.mark.parametrize("solution", ["eclipse", "transit"])
def test_fresh_seed_line_by_line(solution, built, workdir):
    """No opacity file: the reference computes every layer line by line (tau.c:163-175); the
    oracle's restatement on a fresh line list and fresh models."""
    from bart_b200 import synth
    from oracle import oracle as orc
    case = synth.make_case(os.path.join(workdir, "fresh_lbl_" + solution), shape="tiny", solution=solution,
                           seed=4242, nlayer=18, with_grid=False, no_opacity=True, nlines=1500,
                           ethresh=1e-5, refradius_km=95000.0 if solution == "transit" else 123820.0)
    models = synth.make_models(case, 2, seed=1618)
    mp, op = os.path.join(case["workdir"], "m.npy"), os.path.join(case["workdir"], "ref.npz")
    np.save(mp, models)
    conftest.run_reference(case["cfg"], mp, op, {})
    d = np.load(op)
    O = orc.Oracle(case["cfg"])
    for m in range(2):
        o = O.run(models[m], inter=True)
        assert np.array_equal(o["last"], d["last"][m])
        comp = np.abs(d["ext"][m]).sum(axis=1) > 0
        assert relerr(o["ext"][comp], d["ext"][m][comp]) < 1e-6
        assert relerr(o["spectrum"], d["spectra"][m]) < 1e-6


@pytest.mark.parametrize("k", list(__import__("cases").FUZZ_GPU))
def test_randomised_configurations(k, built, workdir):
    """Seeded random configurations (geometry, layer count, toomuch incl. 1e100, ray grids of 2-6
    angles, knobs, spectral window; tests/cases.py build_fuzz_case) through the compiled reference
    and the oracle.  The GPU suite runs the same configurations through the CUDA path."""
    import cases
    from oracle import oracle as orc
    case, models, setters = cases.build_fuzz_case(k, workdir)
    mp, op = os.path.join(case["workdir"], "m.npy"), os.path.join(case["workdir"], "ref.npz")
    np.save(mp, models)
    conftest.run_reference(case["cfg"], mp, op, setters)
    d = np.load(op)
    O = orc.Oracle(case["cfg"])
    apply_setters(O, setters)
    for m in range(2):
        o = O.run(models[m], inter=True)
        assert np.array_equal(o["last"], d["last"][m])
        assert relerr(o["spectrum"], d["spectra"][m]) < 1e-8


@pytest.mark.parametrize("k", list(__import__("cases").FUZZ_LBL))
def test_randomised_line_by_line_configurations(k, built, workdir):
    """Seeded random configurations WITHOUT an opacity file (cases.build_lbl_fuzz_case): the oracle's
    line-by-line forward mode against the compiled reference; the GPU suite runs the same through
    the CUDA path."""
    import cases
    from oracle import oracle as orc
    case, models = cases.build_lbl_fuzz_case(k, workdir)
    mp, op = os.path.join(case["workdir"], "m.npy"), os.path.join(case["workdir"], "ref.npz")
    np.save(mp, models)
    conftest.run_reference(case["cfg"], mp, op, {})
    d = np.load(op)
    O = orc.Oracle(case["cfg"])
    for m in range(2):
        o = O.run(models[m], inter=True)
        assert np.array_equal(o["last"], d["last"][m])
        comp = np.abs(d["ext"][m]).sum(axis=1) > 0
        assert np.array_equal(o["ext"][comp] > 0, d["ext"][m][comp] > 0)
        assert relerr(o["ext"][comp], d["ext"][m][comp]) < 1e-6
        assert relerr(o["spectrum"], d["spectra"][m]) < 1e-6
