"""CPU: the forward-model arithmetic shared with the CUDA kernels (bart_b200/csrc/column_math.cuh:
prefix-scan optical depth, pre-folded CIA splines, chord weights, modulation scan) mapped over
host loops by tests/cpu_emu and compared with the oracle and the reference's golden vectors.
This exercises the host-side readers/config parser of the product as well.  The GPU parity tests
(tests/test_gpu_parity.py) are the real gate; this catches math errors without a GPU."""
import numpy as np
import pytest

import cases
from emu import Emu
from util import relerr, tau_relerr, apply_setters


@pytest.mark.parametrize("name", list(cases.CASES))
def test_column_math_vs_oracle_and_golden(name, built, get_case):
    from oracle import oracle as orc
    case, models, setters = get_case(name)
    g = np.load(cases.golden_path(name))
    E = Emu(case["cfg"])
    O = orc.Oracle(case["cfg"])
    apply_setters(E, setters)
    apply_setters(O, setters)
    assert np.array_equal(E.wn(), g["wn"])
    for m in range(models.shape[0]):
        e = E.run(models[m])
        o = O.run(models[m], inter=True)
        assert e["status"] == 0
        assert relerr(e["radius"], o["radius"]) < 1e-13
        assert relerr(e["ext"], o["ext"]) < 1e-12        # the oracle performs the identity spline resample
        assert np.array_equal(e["last"], g["last"][m]), "last[] differs from the reference"
        assert tau_relerr(e["tau"], o["tau"], o["last"]) < 5e-9
        assert relerr(e["spectrum"], o["spectrum"]) < 1e-9
        assert relerr(e["spectrum"], g["spectra"][m]) < 1e-9


def test_rejection_status(built, get_case):
    case, models, _ = get_case("tiny_eclipse")
    E = Emu(case["cfg"])
    bad = models[0].copy()
    bad[5] = 3500.0                      # above the opacity grid (3000 K)
    assert E.run(bad)["status"] & 1
    bad = models[0].copy()
    nl = E.nlayer
    bad[nl * 6: nl * 7] = 0.9            # H2 abundance -> sum > 1.001
    assert E.run(bad)["status"] & 4


def test_fast_exp_accuracy(built):
    """The kernels' exp (column_math.cuh exp_core / fast_exp) against libm over the ranges the
    forward model uses (-tau/mu down to the clamp at -700, Planck exponents up to the clamp at
    +700): relative error <= 2e-15 + 2e-17 |x| (the one-constant argument reduction carries the
    rounding of N/ln2 into the exponent; 1e-14 at the clamps, 8 orders below the 1e-6 budget)."""
    import ctypes as C
    lib = C.CDLL(__import__("os").path.join(cases.ROOT, "tests", "cpu_emu", "libemu.so"))
    lib.emu_fast_exp.restype = C.c_double
    lib.emu_fast_exp.argtypes = [C.c_double]
    rng = np.random.default_rng(1)
    xs = np.concatenate([-10.0 ** rng.uniform(-12, 2.845, 20000), rng.uniform(0, 700, 5000),
                         [0.0, -0.0, -1e-300, -700.0, 700.0, 1.0, -1.0, 0.5 * np.log(2),
                          -0.5 * np.log(2), np.log(2) / 256, -np.log(2) / 256]])
    worst = 0.0
    for x in xs:
        got, ref = lib.emu_fast_exp(float(x)), float(np.exp(x))
        worst = max(worst, abs(got - ref) / ref / (2e-15 + 2e-17 * abs(x)))
    assert worst < 1.0, worst
    assert lib.emu_fast_exp(0.0) == 1.0
    # arguments are clamped to [-700, 700]: an opaque deck transmits e^-700 ~ 1e-304, not NaN
    assert lib.emu_fast_exp(-1e9) == lib.emu_fast_exp(-700.0) < 1e-300
    assert lib.emu_fast_exp(1e5) == lib.emu_fast_exp(700.0) and np.isfinite(lib.emu_fast_exp(1e5))


def test_wavelength_limits_give_the_reference_grid(built, workdir):
    """wllow / wlhigh / wlfct (what makecfg.py writes from BART.cfg) through the product's host code
    (cfg.cpp, readers.cpp make_sampling) and through the oracle: the reference's wavenumber grid."""
    import numpy as np
    import cases
    from emu import Emu
    from oracle import oracle as orc
    g = np.load(cases.golden_path("wl_ranges"))
    for k in range(len(cases.WL_CASES)):
        case = cases.build_wl_case(k, workdir)
        assert np.array_equal(Emu(case["cfg"]).wn(), g["wn%d" % k])
        assert np.array_equal(orc.Oracle(case["cfg"]).wn, g["wn%d" % k])
