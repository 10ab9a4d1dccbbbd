"""CPU: host-side readers of the product (bart_b200/csrc/readers.cpp, compiled into the TEST-ONLY
emulation library) against the oracle's independent Python readers.  The TLI line block is
memory-mapped and sliced per isotope with the reference's binary search + linear refinement
(readlineinfo.c:16-77, 416-537)."""
import ctypes as C
import os
import numpy as np
import pytest

import cases

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("window", [(2500.0, 2700.0), (2555.5, 2601.25), (2699.0, 2700.0), (100.0, 200.0)])
def test_mmap_tli_selection_matches_oracle(window, built, workdir):
    from oracle import oracle as orc
    from bart_b200 import synth
    case = synth.make_case(os.path.join(workdir, "tli_reader"), shape="tiny", nlayer=8, with_grid=False,
                           nlines=4000, seed=77)
    lib = C.CDLL(os.path.join(HERE, "cpu_emu", "libemu.so"))
    lib.emu_read_tli.restype = C.c_longlong
    dp, sp = C.POINTER(C.c_double), C.POINTER(C.c_short)
    lib.emu_read_tli.argtypes = [C.c_char_p, C.c_double, C.c_double, C.c_longlong, dp, dp, dp, sp]
    tli = orc.read_tli(case["tli"])
    idx = orc.select_lines(tli, *window)
    cap = len(tli["wl"]) + 1
    wl, el, gf = np.zeros(cap), np.zeros(cap), np.zeros(cap)
    iso = np.zeros(cap, dtype=np.int16)
    n = lib.emu_read_tli(case["tli"].encode(), window[0], window[1], cap, wl.ctypes.data_as(dp),
                         el.ctypes.data_as(dp), gf.ctypes.data_as(dp), iso.ctypes.data_as(sp))
    assert n == len(idx)
    assert np.array_equal(wl[:n], tli["wl"][idx])
    assert np.array_equal(el[:n], tli["elow"][idx])
    assert np.array_equal(gf[:n], tli["gf"][idx])
    assert np.array_equal(iso[:n], tli["isoid"][idx])


def test_tli_reader_rejects_truncated_file(built, workdir):
    from bart_b200 import synth
    case = synth.make_case(os.path.join(workdir, "tli_trunc"), shape="tiny", nlayer=8, with_grid=False,
                           nlines=500, seed=78)
    raw = open(case["tli"], "rb").read()
    cut = os.path.join(workdir, "tli_trunc", "cut.tli")
    with open(cut, "wb") as f:
        f.write(raw[:len(raw) - 1000])
    lib = C.CDLL(os.path.join(HERE, "cpu_emu", "libemu.so"))
    lib.emu_read_tli.restype = C.c_longlong
    dp, sp = C.POINTER(C.c_double), C.POINTER(C.c_short)
    lib.emu_read_tli.argtypes = [C.c_char_p, C.c_double, C.c_double, C.c_longlong, dp, dp, dp, sp]
    z = np.zeros(1)
    zi = np.zeros(1, dtype=np.int16)
    assert lib.emu_read_tli(cut.encode(), 2500.0, 2700.0, 0, z.ctypes.data_as(dp), z.ctypes.data_as(dp),
                            z.ctypes.data_as(dp), zi.ctypes.data_as(sp)) == -1
    lib.emu_error.restype = C.c_char_p
    assert b"truncated" in lib.emu_error()
