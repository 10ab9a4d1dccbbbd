"""CPU: the C-ABI library loads and exports every symbol include/bart_b200.h declares; the
CPython transit_module has the SWIG surface and its argument checking.  No compute calls."""
import ctypes as C
import os
import re
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bart_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)\s*\(", src)
    return sorted({n for n in names if n.startswith(("bart_", "transit_", "get_", "set_", "run_",
                                                      "free_"))})


def test_header_symbols_exported(built):
    lib = C.CDLL(os.path.join(ROOT, "bart_b200", "libbart_b200.so"))
    syms = declared_symbols()
    assert len(syms) >= 40
    for ref in ("transit_init", "get_no_samples", "get_waveno_arr", "set_radius", "set_cloudtop",
                "set_scattering", "run_transit", "free_memory"):
        assert ref in syms
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, "declared but not exported: %s" % missing


def test_library_is_sm100a_only(built):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "bart_b200", "libbart_b200.so")],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_oracle_in_product(built):
    """The product must not link, import or reference anything under oracle/ or tests/."""
    import subprocess
    so = os.path.join(ROOT, "bart_b200", "libbart_b200.so")
    ldd = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "emu" not in ldd
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bart_b200")):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".c", ".hpp", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in txt and "from oracle" not in txt and \
                    "import oracle" not in txt and "cpu_emu" not in txt.replace("tests/cpu_emu", ""), f


def test_transit_module_surface(built):
    sys.path.insert(0, os.path.join(ROOT, "bart_b200", "python"))
    import transit_module as trm
    for name in ("transit_init", "get_no_samples", "get_waveno_arr", "set_radius", "set_cloudtop",
                 "set_scattering", "run_transit", "free_memory"):
        assert callable(getattr(trm, name))
    with pytest.raises(TypeError):
        trm.transit_init(3, "not a list")
    with pytest.raises(TypeError):
        trm.transit_init(3, ["transit", "-c", 42])
    assert trm.get_no_samples() == 0                      # not initialised
    arr = trm.get_waveno_arr(4)                           # -1 fill like the reference
    assert arr.shape == (4,) and (arr == -1).all()
    with pytest.raises(TypeError):
        trm.run_transit([[1.0, 2.0], [3.0, 4.0]], 4)      # IN_ARRAY1 wants 1-D


def test_missing_device_fails_loudly(built):
    """On a box without a GPU every compute entry point must raise, not fall back."""
    from bart_b200 import api
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(api.BartError) as e:
        api.device_info()
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)
