import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    import oracle_lib as ol
    if not os.path.exists(ol.PORT_SO):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
    lib = os.path.join(ROOT, "jackal-navigation_b200", "libjn_elas.so")
    if not os.path.exists(lib):
        importlib.import_module("jackal-navigation_b200.build").build()


@pytest.fixture(scope="session")
def jn():
    _ensure_built()
    return importlib.import_module("jackal-navigation_b200")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("jackal-navigation_b200.synth")


@pytest.fixture(scope="session")
def port():
    _ensure_built()
    import oracle_lib as ol
    o = ol.load("port")
    assert o is not None
    return o


@pytest.fixture(scope="session")
def ref():
    import oracle_lib as ol
    o = ol.load("ref")
    if o is None:
        pytest.skip("oracle/_ref/libelas_ref.so not built (needs /root/reference)")
    return o


@pytest.fixture(scope="session")
def oracle(port):
    """The strongest checker available: the compiled reference if present, else the port."""
    import oracle_lib as ol
    return ol.load("ref") or port
