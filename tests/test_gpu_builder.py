"""GPU (B200): the CUDA opacity-grid builder (K5 Voigt table, K6 line binning) through the C ABI
against (1) grids built by the unmodified reference (`transit --justOpacity`, tests/golden),
(2) the builder oracle, with line-to-bin indices bit-exact and the opacity file byte layout of
opacity.c:406-421."""
import os
import numpy as np
import pytest

import cases
from util import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-6      # grid values: float32 Voigt table => north-star tolerance 1e-6 relative


@pytest.fixture(scope="module")
def api(built):
    from bart_b200 import api as a
    a.device_info()
    return a


@pytest.mark.parametrize("name", list(cases.BUILD_CASES))
def test_built_grid_vs_reference(name, api, workdir):
    from bart_b200 import synth
    case = cases.build_builder_case(name, workdir)
    g = np.load(cases.golden_path(name))
    assert cases.sha(np.fromfile(case["tli"], dtype=np.uint8)) == str(g["tli_sha"])
    assert not os.path.exists(case["opacity"])
    tr = api.Transit(argv=["transit", "-c", case["cfg"], "--justOpacity"])      # BART.py:563-565
    assert os.path.getsize(case["opacity"]) == int(g["file_bytes"])
    mine = synth.read_opacity(case["opacity"], mmap=False)
    assert np.array_equal(mine["molids"], g["molids"])
    assert np.array_equal(mine["temps"], g["temps"])
    assert relerr(mine["press"], g["press"]) < 1e-14     # reference: after its identity spline resample
    assert np.array_equal(mine["wn"], g["wn"])
    ref = g["grid"]
    assert np.array_equal(mine["o"] > 0, ref > 0)
    assert relerr(mine["o"], ref) < TOL
    tr.free_memory()
    # the file just written drives a forward model (init path with an existing grid)
    tr2 = api.Transit(case["cfg"])
    models = synth.make_models(case, 2, seed=5)
    spectra, status = tr2.run_batch(models)
    assert (status == 0).all() and np.isfinite(spectra).all()
    tr2.free_memory()


def test_bins_and_profiles_vs_oracle(api, workdir):
    """Bit-exact line -> oversampled-bin indices (incl. the co-add grouping) and Voigt profiles
    within one float32 ulp of the oracle's long-double evaluation."""
    import ctypes as C
    from oracle import oracle as orc
    case = cases.build_builder_case("build_h2o_ch4", workdir)
    tr = api.Transit(argv=["transit", "-c", case["cfg"], "--justOpacity"])
    L = api.lib()
    B = orc.BuilderOracle(case["cfg"])
    B.build(layers=[0], temps=[0], trace=True)
    n = len(B.wl)
    bins = np.zeros(n, dtype=np.int64)
    got = L.bart_line_bins(bins.ctypes.data_as(C.POINTER(C.c_longlong)), n)
    assert got == n
    assert np.array_equal(bins, B.trace), "line-to-bin indices differ"
    worst = 0.0
    for (i, j) in ((0, 0), (5, 40), (30, 10), (59, 59), (20, 59), (59, 0)):
        ref, ps = B.profile(i, j) if B.profiles[i * B.nLor + j] is None else \
            (B.profiles[i * B.nLor + j], B.profsize[i * B.nLor + j])
        hs = C.c_longlong()
        out = np.zeros(2 * int(ps) + 1, dtype=np.float32)
        api._check(L.bart_voigt_profile(i, j, out.ctypes.data_as(C.POINTER(C.c_float)), out.size, C.byref(hs)))
        assert hs.value == ps
        worst = max(worst, relerr(out, ref))
    assert worst < 5e-7, worst
    tr.free_memory()


def test_line_bins_million_lines(api, workdir, monkeypatch):
    """a19 index hazards at scale (extinction.c:431-462: `(wavn - wns.i)/odwn` truncation, the
    nearest-oversampled-node test, the co-add run): 1.2e6 random lines onto 601 x 1080 oversampled
    bins through the CUDA path (line_index_kernel + host grouping) against the oracle's trace --
    leader bin of every line, or the leader it was co-added into.  Disagreements are counted, and
    the count must be zero."""
    import ctypes as C
    import os
    from bart_b200 import synth
    from oracle import oracle as orc
    case = synth.make_case(os.path.join(workdir, "bins_1e6"),
                           shape=dict(wnlow=2000.0, wnhigh=2600.0, wndelt=1.0, mols=["H2O", "CH4"], toomuch=10.0),
                           nlayer=10, with_grid=False, nlines=1200000, tempdelt=800.0, seed=4713,
                           ethresh=1e-4, wnosamp=1080, nwidth=30)
    monkeypatch.setenv("BART_TSLICE", "0:0")             # load and index the lines, build no plane
    tr = api.Transit(argv=["transit", "-c", case["cfg"], "--justOpacity"])
    monkeypatch.delenv("BART_TSLICE")
    B = orc.BuilderOracle(case["cfg"])
    n = len(B.wl)
    assert n >= 1000000
    B.build(layers=[len(B.atm["press"]) - 1], temps=[0], trace=True)
    bins = np.zeros(n, dtype=np.int64)
    assert api.lib().bart_line_bins(bins.ctypes.data_as(C.POINTER(C.c_longlong)), n) == n
    ndiff = int((bins != B.trace).sum())
    nlead, nco = int((B.trace >= 0).sum()), int((B.trace <= -2).sum())
    print("line bins: %d lines, %d leaders, %d co-added, %d disagreements" % (n, nlead, nco, ndiff))
    assert nlead > 300000 and nco > 100000
    assert ndiff == 0
    tr.free_memory()


def test_device_grouping_equals_host_walk(api, workdir, monkeypatch):
    """The co-add grouping on the device (speculative per-block chain walks + one fix-up walk,
    streamed TLI upload) against the sequential host walk of extinction.c:450-462
    ($BART_GROUP_HOST=1): the per-line trace and the grid planes built from the groups are identical,
    on a dense line list (0.5 oversampled bins between lines, a third of the lines co-added), a
    sparse one and an extremely dense one."""
    import ctypes as C
    import os
    from bart_b200 import synth
    L = api.lib()
    # "extreme": 40 lines per oversampled bin everywhere (groups of dozens of lines)
    for tag, nlines, osamp, wnhigh in (("dense", 300000, 1080, 2300.0), ("sparse", 20000, 2160, 2300.0),
                                       ("extreme", 250000, 120, 2050.0)):
        case = synth.make_case(os.path.join(workdir, "grp_" + tag),
                               shape=dict(wnlow=2000.0, wnhigh=wnhigh, wndelt=1.0, mols=["H2O", "CH4"], toomuch=10.0),
                               nlayer=6, with_grid=False, nlines=nlines, tempdelt=1300.0, seed=99,
                               ethresh=1e-5, wnosamp=osamp)
        got = {}
        for host in ("1", "0"):
            monkeypatch.setenv("BART_GROUP_HOST", host)
            monkeypatch.setenv("BART_TSLICE", "0:0")
            tr = api.Transit(argv=["transit", "-c", case["cfg"], "--justOpacity"])
            monkeypatch.delenv("BART_TSLICE")
            nl, nw = tr.nlayer, tr.nwave
            nmol = 2
            nlines_c, ngroups_c, neval_c = C.c_longlong(), C.c_longlong(), C.c_longlong()
            out = np.zeros((nl, 1, nmol, nw))
            api._check(L.bart_build_opacity_slice(1, 2, out.ctypes.data_as(api.dp)))
            L.bart_builder_stats(C.byref(nlines_c), C.byref(ngroups_c), C.byref(neval_c))
            n = nlines_c.value
            bins = np.zeros(n, dtype=np.int64)
            assert L.bart_line_bins(bins.ctypes.data_as(C.POINTER(C.c_longlong)), n) == n
            got[host] = (n, ngroups_c.value, neval_c.value, bins, out)
            L.bart_builder_phase_ms.restype = C.c_double
            on_host = L.bart_builder_phase_ms(b"grouping_host") > 0
            assert on_host == (host == "1"), (tag, host)
            tr.free_memory()
        h, d = got["1"], got["0"]
        assert h[0] == d[0] > 0.9 * nlines and h[1] == d[1] and h[2] == d[2]
        assert np.array_equal(h[3], d[3])
        assert np.array_equal(h[4], d[4]) and (h[4] > 0).any()
        nco = int((h[3] <= -2).sum())
        assert nco > (0.2 * nlines if tag == "dense" else 0)
    monkeypatch.delenv("BART_GROUP_HOST")


def test_high_resolution_wide_profiles_vs_oracle(api, workdir, monkeypatch):
    """The sweep's resolution (0.02423 cm-1 per sample, wnosamp 2160): pressure-broadened profiles
    span hundreds to thousands of bins, i.e. the accumulate kernel's lane-per-bin paths (tile-covering
    loop and range-tested ends) carry the result.  Random (layer, T) cells against the builder oracle."""
    import os
    from bart_b200 import synth
    from oracle import oracle as orc
    case = synth.make_case(os.path.join(workdir, "hr_wide"),
                           shape=dict(wnlow=2000.0, wnhigh=2000.0 + 0.02423 * 12000, wndelt=0.02423,
                                      mols=["H2O", "CO2", "CO", "CH4"], toomuch=10.0),
                           nlayer=100, with_grid=False, nlines=60000, tempdelt=100.0, seed=2026,
                           ethresh=1e-6, wnosamp=2160)
    monkeypatch.setenv("BART_TSLICE", "0:0")
    tr = api.Transit(argv=["transit", "-c", case["cfg"], "--justOpacity"])
    monkeypatch.delenv("BART_TSLICE")
    L = api.lib()
    nl, nw, nmol = tr.nlayer, tr.nwave, 4
    assert nw == 12001
    B = orc.BuilderOracle(case["cfg"])
    out = np.zeros((nl, 1, nmol, nw))
    it = 9
    api._check(L.bart_build_opacity_slice(it, it + 1, out.ctypes.data_as(api.dp)))
    worst = 0.0
    for r in (0, 3, 41, 99):                          # bottom (widest profiles) ... top (Doppler cores)
        ref = np.asarray(B.build(layers=[r], temps=[it])).reshape(nmol, nw)
        got = out[r, 0]
        m = ref > 0
        assert np.array_equal(got > 0, m) and m.any()
        worst = max(worst, float(np.max(np.abs(got[m] - ref[m]) / ref[m])))
    assert worst < 1e-12, worst
    tr.free_memory()


def test_temperature_sharded_build(api, workdir):
    """T-sharded build (bart_build_opacity_slice): slices reassemble to the full grid bit for bit."""
    case = cases.build_builder_case("build_ch4", workdir)
    g = np.load(cases.golden_path("build_ch4"))
    tr = api.Transit(argv=["transit", "-c", case["cfg"], "--justOpacity"])
    L = api.lib()
    from bart_b200 import synth
    full = synth.read_opacity(case["opacity"], mmap=False)["o"]
    nl, nt, nm, nw = full.shape
    parts = []
    for (a, b) in ((0, 2), (2, 3), (3, nt)):
        out = np.zeros((nl, b - a, nm, nw))
        api._check(L.bart_build_opacity_slice(a, b, out.ctypes.data_as(api.dp)))
        parts.append(out)
    assert np.array_equal(np.concatenate(parts, axis=1), full)
    assert relerr(full, g["grid"]) < TOL
    tr.free_memory()


def test_cli_temperature_slices_assemble_the_file(api, workdir):
    """`transit -c cfg --justOpacity` run as separate processes with $BART_TSLICE (one per GPU on a
    multi-GPU box; sequential here): every process streams its planes to their offsets in the SAME
    file (opacity.c:418-421 order); the result equals the single-process file byte for byte."""
    import subprocess
    from bart_b200 import synth
    case = cases.build_builder_case("build_h2o_ch4", workdir)
    g = np.load(cases.golden_path("build_h2o_ch4"))
    exe = os.path.join(cases.ROOT, "bart_b200", "bin", "transit")
    nt = len(g["temps"])
    env = dict(os.environ)
    for sl in ("%d:%d" % (nt // 2, nt), "0:%d" % (nt // 2)):           # header written by the 0: slice
        env["BART_TSLICE"] = sl
        r = subprocess.run([exe, "-c", case["cfg"], "--justOpacity"], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    sliced = np.fromfile(case["opacity"], dtype=np.uint8)
    assert sliced.size == int(g["file_bytes"])
    os.remove(case["opacity"])
    env.pop("BART_TSLICE")
    r = subprocess.run([exe, "-c", case["cfg"], "--justOpacity"], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    whole = np.fromfile(case["opacity"], dtype=np.uint8)
    assert np.array_equal(sliced, whole)
    mine = synth.read_opacity(case["opacity"], mmap=False)
    assert relerr(mine["o"], g["grid"]) < TOL


@pytest.mark.parametrize("k", list(cases.FUZZ_BUILDER))
def test_randomised_builder_configurations(k, api, workdir):
    """The seeded random builder configurations of cases.build_builder_fuzz_case (the oracle agrees
    with the compiled reference's grids on all of them, tests/test_builder_oracle.py) through the
    CUDA builder: whole file written, sampled cells against the oracle, line bins bit-exact."""
    import ctypes as C
    from oracle import oracle as orc
    from bart_b200 import synth
    case = cases.build_builder_fuzz_case(k, workdir)
    tr = api.Transit(argv=["transit", "-c", case["cfg"], "--justOpacity"])
    mine = synth.read_opacity(case["opacity"], mmap=False)
    B = orc.BuilderOracle(case["cfg"])
    nl, nt = mine["o"].shape[0], mine["o"].shape[1]
    assert np.array_equal(mine["temps"], B.temps) and list(mine["molids"]) == list(B.gmol_id)
    layers, temps = sorted({0, nl // 2, nl - 1}), sorted({0, nt - 1})
    ref = B.build(layers=layers, temps=temps, trace=True)
    got = mine["o"][layers][:, temps]
    assert np.array_equal(got > 0, ref > 0)
    assert relerr(got, ref) < TOL
    n = len(B.wl)
    bins = np.zeros(n, dtype=np.int64)
    assert api.lib().bart_line_bins(bins.ctypes.data_as(C.POINTER(C.c_longlong)), n) == n
    B.build(layers=[0], temps=[0], trace=True)           # the trace of the cell the library reports
    assert np.array_equal(bins, B.trace)
    tr.free_memory()
