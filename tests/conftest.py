import os
import sys
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the in-tree artefacts exist (no-op when __graft_entry__.build() already ran)."""
    import __graft_entry__ as g
    if not (os.path.exists(os.path.join(ROOT, "bart_b200", "libbart_b200.so")) and
            os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")) and
            os.path.exists(os.path.join(ROOT, "tests", "cpu_emu", "libemu.so"))):
        g.build()
    return True


@pytest.fixture(scope="session")
def workdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("bart_cases"))


_case_cache = {}


@pytest.fixture(scope="session")
def get_case(workdir):
    import cases

    def _get(name):
        if name not in _case_cache:
            _case_cache[name] = cases.build_case(name, workdir)
        return _case_cache[name]
    return _get


def has_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libtransit_ref.so"))


def run_reference(cfg, models_path, out_path, setters=None, inter=True, extra=()):
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_driver.py"), cfg, models_path, out_path]
    if inter:
        cmd.append("--inter")
    for k, v in (setters or {}).items():
        cmd.append("--%s=%r" % (k, v))
    cmd += list(extra)
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, "reference failed:\n%s\n%s" % (r.stdout[-2000:], r.stderr[-2000:])
