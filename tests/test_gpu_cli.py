"""GPU (B200): the command-line program `bart_b200/bin/transit -c cfg` (transit.c:230-242; what
BART.py:632-634 runs for the best-fit spectrum) against the spectrum files the UNMODIFIED reference
program wrote for the same configuration (printflux eclipse.c:355-380, printmod
slantpath.c:510-555): same header, same wavelength column text, values to the printed precision."""
import os
import subprocess
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(cases.CLI_CASES))
def test_cli_spectrum_file_matches_reference(name, built, workdir):
    case = cases.build_cli_case(name, workdir)
    g = np.load(cases.golden_path(name))
    assert cases.sha(case["grid"]) == str(g["grid_sha"])
    ref = str(g["text"]).splitlines()
    out = os.path.join(case["workdir"], "outspec.dat")
    if os.path.exists(out):
        os.remove(out)
    exe = os.path.join(cases.ROOT, "bart_b200", "bin", "transit")
    r = subprocess.run([exe, "-c", case["cfg"]], capture_output=True, text=True, cwd=case["workdir"])
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    with open(out) as f:
        mine = f.read().splitlines()
    assert len(mine) == len(ref)
    assert mine[0] == ref[0]                                   # header line
    width = 15 if case["solution"] == "eclipse" else 17        # "%-15.10g" / "%-17.9g" wavelength column
    worst = 0.0
    for a, b in zip(mine[1:], ref[1:]):
        assert a[:width] == b[:width]                          # wavelength text identical
        va, vb = float(a[width:]), float(b[width:])
        worst = max(worst, abs(va - vb) / abs(vb))
    assert worst < 2e-9                                        # 9 significant digits are printed
