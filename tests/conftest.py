import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _built():
    need = [os.path.join(ROOT, "gpu-pathtracer_b200", "csrc", "libb200pt.so"), os.path.join(ROOT, "oracle", "libpt_oracle.so"),
            os.path.join(ROOT, "tests", "emu", "libb200pt_emu.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g
        g.build()


@pytest.fixture(scope="session", autouse=True)
def built_libs():
    _built()


@pytest.fixture(scope="session")
def oracle():
    from tests.oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
