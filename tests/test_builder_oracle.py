"""CPU: the line-by-line builder oracle (orc_voigtn + orc_computemolext driven like
calcprofiles/calcopacity) against grids built by the reference's `transit --justOpacity`."""
import numpy as np
import pytest

import cases
from util import relerr


@pytest.mark.parametrize("name", list(cases.BUILD_CASES))
def test_builder_oracle_vs_reference_grid(name, built, workdir):
    from oracle import oracle as orc
    case = cases.build_builder_case(name, workdir)
    g = np.load(cases.golden_path(name))
    assert cases.sha(np.fromfile(case["tli"], dtype=np.uint8)) == str(g["tli_sha"])
    B = orc.BuilderOracle(case["cfg"])
    assert np.array_equal(B.temps, g["temps"])
    assert list(B.gmol_id) == list(g["molids"])
    nl = g["grid"].shape[0]
    layers = sorted({0, nl // 2, nl - 1})
    o = B.build(layers=layers)
    ref = g["grid"][layers]
    assert np.array_equal(o > 0, ref > 0)
    assert relerr(o, ref) < 1e-12          # float32 profiles are reproduced bit for bit


@pytest.mark.parametrize("k", list(cases.FUZZ_BUILDER))
def test_randomised_builder_configurations(k, built, workdir):
    """Seeded random builder configurations (cases.build_builder_fuzz_case): the builder oracle
    against the grid the compiled reference builds (`transit_ref --justOpacity`; build container
    only).  The GPU suite runs the CUDA builder on the same configurations against the oracle."""
    import os
    import subprocess
    import conftest
    if not conftest.has_ref():
        pytest.skip("oracle/_ref not built here")
    from oracle import oracle as orc
    from bart_b200 import synth
    case = cases.build_builder_fuzz_case(k, workdir)
    exe = os.path.join(cases.ROOT, "oracle", "_ref", "transit_ref")
    r = subprocess.run([exe, "-c", case["cfg"], "--justOpacity"], capture_output=True, text=True)
    assert r.returncode == 0 and os.path.exists(case["opacity"]), r.stdout[-1500:] + r.stderr[-1500:]
    ref = synth.read_opacity(case["opacity"], mmap=False)
    os.remove(case["opacity"])
    B = orc.BuilderOracle(case["cfg"])
    nl, nt = ref["o"].shape[0], ref["o"].shape[1]
    layers, temps = sorted({0, nl // 2, nl - 1}), sorted({0, nt - 1})
    o = B.build(layers=layers, temps=temps)
    want = ref["o"][layers][:, temps]
    assert np.array_equal(o > 0, want > 0)
    assert relerr(o, want) < 1e-12


def test_voigt_kat():
    """Known values of the Voigt function the profile table samples: pure-Doppler and
    pure-Lorentz limits (analytic), for the region formulas of voigt.c:132-200."""
    from oracle import oracle as orc
    import ctypes as C
    L = orc.lib()
    # fine spacing << alphaD selects the adjacent-average branch; the bin-averaged profile must
    # integrate to ~1 over +-20 widths and peak near the analytic centre value
    for aL, aD in ((1e-4, 0.05), (0.05, 0.05), (0.5, 0.01)):
        dw = 1e-3
        n = 2 * int(20 * max(aL, aD) / dw) + 1
        out = np.zeros(n, dtype=np.float32)
        L.orc_voigtn(n, dw * (n // 2), aL, aD, out.ctypes.data_as(C.POINTER(C.c_float)), 0)
        area = float(out.sum()) * dw
        assert abs(area - 1.0) < 0.04, (aL, aD, area)      # Lorentz wings beyond 20 widths: ~3 %
        assert out.argmax() in (n // 2 - 1, n // 2)
    # Doppler limit: centre value sqrt(ln2/pi)/aD
    aD = 0.05
    out = np.zeros(3, dtype=np.float32)
    L.orc_voigtn(3, 1e-6, 1e-9, aD, out.ctypes.data_as(C.POINTER(C.c_float)), 1)
    assert abs(out[1] - np.sqrt(np.log(2) / np.pi) / aD) / out[1] < 1e-5
