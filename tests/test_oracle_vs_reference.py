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
