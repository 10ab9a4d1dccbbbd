"""Known-answer tests of the reference's own test design (modules/transit/transit/test/
test_slantpath.c): analytic chord optical depths for constant, outward-increasing and
inward-increasing extinction (const_ex_anal / incout_ex_anal / incin_ex_anal, lines 177-198) and
analytic modulations for constant and linear optical-depth profiles (mod_ctau / mod_itau /
mod_dtau, lines 231-307), on the reference's layer counts (100 and 1000; its pass threshold is 1e-4,
line 62 -- the current integrator is at 3e-7 with 1000 layers).
Applied to (i) the oracle's restatement of totaltau1 / modulation1 and (ii) the product's chord
weights as the device code builds them (column_math.cuh, through tests/cpu_emu).  The observed
values recorded in test/slantpath.080604.dat come from the 2004 integrator and are not reproduced
by the current reference code; the analytic values are version independent."""
import ctypes as C
import os
import numpy as np
import pytest

from oracle import oracle as orc

dp = C.POINTER(C.c_double)
HERE = os.path.dirname(os.path.abspath(__file__))


def calcex(alpha, rm, r):                                   # test_slantpath.c:67-78
    if alpha == 0:
        return np.ones_like(r)
    return -alpha * (rm - r) if alpha < 0 else alpha * r


def analytic_tau(alpha, rm, ip):                            # test_slantpath.c:177-198
    rat = rm / ip
    if alpha == 0:
        return 2 * np.sqrt(rm * rm - ip * ip)
    a = rm * ip * np.sqrt(rat * rat - 1)
    b = ip * ip * np.log(np.sqrt(rat * rat - 1) + rat)
    return alpha * (a + b) if alpha > 0 else -alpha * (a - b)


def emu():
    lib = C.CDLL(os.path.join(HERE, "cpu_emu", "libemu.so"))
    lib.emu_chord_tau.argtypes = [C.c_int, dp, dp, dp]
    return lib


@pytest.mark.parametrize("alpha", [0.0, 1.0, -1.0])
def test_chord_optical_depth_analytic(alpha, built):
    """tau_dens / tau_ip of the reference's test: planets of radius 10, 100, 1000, rays crossing at
    0.1, 0.5, 0.75 and 0.9 of the radius, 100 and 1000 layers."""
    L = orc.lib()
    E = emu()
    for rm in (10.0, 100.0, 1000.0):
        for frac in (0.1, 0.5, 0.75, 0.9):
            ip = frac * rm
            want = analytic_tau(alpha, rm, ip)
            for n, tol in ((100, 3e-3), (1000, 1e-6)):   # the scheme converges like n^-3: 1.9e-3, 3.2e-7
                rad = (np.arange(n) + 1.0) * (rm / n)        # bottom -> top, like tau_dens
                ex = calcex(alpha, rm, rad)
                got = L.orc_totaltau1(ip, rad.ctypes.data_as(dp), ex.ctypes.data_as(dp), n)
                assert abs(got / want - 1) < tol, (alpha, rm, frac, n, got, want)
                # the product's weights: impact parameters are layer radii there (depth d <-> layer
                # n-1-d); every ray of the reference's list sits on a layer of these grids
                k = int(round(ip / (rm / n))) - 1
                assert abs(rad[k] - ip) < 1e-9 * rm
                tau = np.zeros(n)
                rtd, etd = rad[::-1].copy(), ex[::-1].copy()
                E.emu_chord_tau(n, rtd.ctypes.data_as(dp), etd.ctypes.data_as(dp), tau.ctypes.data_as(dp))
                dev = tau[n - 1 - k]
                assert abs(dev / want - 1) < tol, (alpha, rm, frac, n, dev, want)
                assert abs(dev / got - 1) < 1e-12            # device weights == oracle integration


def test_modulation_analytic(built):
    """mod_ctau / mod_itau / mod_dtau: constant tau, tau increasing outwards and inwards; star of
    radius 10 x the planet's, atmosphere from `first` x ipmax upwards, toomuch far away."""
    L = orc.lib()
    toomuch = 1e300
    star = 50.0
    for ipmax in (1.0, 5.0):
        for first in (0.5, 0.9):
            for nip, tol in ((101, 1e-4), (1001, 1e-6)):
                ipv = first * ipmax + (nip - 1 - np.arange(nip)) * (ipmax * (1 - first) / (nip - 1))  # top -> bottom
                rath, ratl = ipmax / star, first * ipmax / star
                for kind, prm in (("const", 1.0), ("out", 0.8), ("in", 0.8)):
                    delt = (1 - first) / (nip - 1)
                    idx = np.arange(nip)
                    if kind == "const":
                        tau = np.full(nip, prm)
                        want = -np.exp(-prm) * (rath * rath - ratl * ratl) + rath * rath
                    elif kind == "out":
                        tau = prm * ipmax * (1 - idx * delt)
                        want = -2 * (np.exp(-prm * ipmax * first) * (first * ipmax + 1 / prm) -
                                     np.exp(-prm * ipmax) * (ipmax + 1 / prm)) / star / star / prm + rath * rath
                    else:
                        tau = prm * ipmax * idx * delt
                        want = -2 * ((ipmax - 1 / prm) - np.exp(-prm * ipmax * (1 - first)) *
                                     (ipmax * first - 1 / prm)) / star / star / prm + rath * rath
                    # the reference's formulas carry - exp(-toomuch) ratl^2 for the opaque core
                    # (mod_*tau); with toomuch out of reach that term vanishes and the level-1
                    # modulation is called without the transparent-core correction
                    got = L.orc_modulation1(tau.ctypes.data_as(dp), nip - 1, toomuch, ipv.ctypes.data_as(dp),
                                            nip, 1.0, star, 0)
                    assert abs(got / want - 1) < tol, (ipmax, first, nip, kind, got, want)


def test_voigt_against_recorded_table_and_faddeeva(built):
    """The reference's test directory records a Voigt table (test/voigt.080904.dat, alpha_L 1.5,
    alpha_D 1, 4 significant digits; fixture tests/golden/kat_voigt.npz).  It was written by the 2004
    code and sits up to 1.1 % (0.45 % in its bin-centre column) off the true Voigt function; the current voigtxy (Pierluissi's
    three-region approximation, pu/src/voigt.c:132-200), which the oracle restates, is within 4e-5 of
    it.  Both statements are checked: the oracle against an independent Faddeeva evaluation
    (scipy.special.voigt_profile) and against the recorded table."""
    from scipy.special import voigt_profile
    L = orc.lib()
    g = np.load(os.path.join(HERE, "golden", "kat_voigt.npz"))
    n, sub, first = int(g["nrows"]), int(g["subbins"]), int(g["first_sub"])
    h = 2 * float(g["halfrange"]) / (n - 1) / sub
    nf = (n - 1) * sub + 17                                  # 8 fine samples of margin on both sides
    out = np.zeros(nf, dtype=np.float32)
    L.orc_voigtn(nf, float(g["halfrange"]) + 8 * h, float(g["alphaL"]), float(g["alphaD"]),
                 out.ctypes.data_as(C.POINTER(C.c_float)), 1)
    x = -float(g["halfrange"]) - 8 * h + np.arange(nf) * h
    exact = voigt_profile(x, float(g["alphaD"]) / np.sqrt(2 * np.log(2)), float(g["alphaL"]))
    assert np.max(np.abs(out / exact - 1)) < 5e-5
    worst = 0.0
    for j in range(g["table"].shape[1]):
        idx = g["rows"] * sub + (first + j) + 8
        worst = max(worst, float(np.max(np.abs(out[idx] / g["table"][:, j] - 1))))
    assert worst < 1.5e-2, worst


@pytest.mark.parametrize("nl", [1, 2, 3, 4, 9, 37, 100, 101])
def test_chord_weight_rows_split_over_workers(nl, built):
    """transit_weights_kernel deals a depth's row to several threads, each taking a share of the
    row's Simpson panels (column_math.cuh transit_weight_row_parts): for 1..7 workers the assembled
    row equals the whole row bit for bit and every element is written, at layer counts with odd and
    even panel tails and fewer panels than workers."""
    E = emu()
    E.emu_chord_parts_mismatch.argtypes = [C.c_int, dp, C.c_int]
    rng = np.random.default_rng(nl)
    rad = np.sort(7.0e9 + np.cumsum(rng.uniform(2e6, 9e6, nl)))[::-1].copy()     # by depth: top first
    for nparts in (1, 2, 3, 4, 7):
        assert E.emu_chord_parts_mismatch(nl, rad.ctypes.data_as(dp), nparts) == 0
