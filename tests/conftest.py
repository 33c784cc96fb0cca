import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_bindings import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref); tests that need it skip when it is absent."""
    from oracle_bindings import load_reference
    r = load_reference()
    if r is None:
        pytest.skip("oracle/_ref not built (reference sources not mounted)")
    return r


@pytest.fixture(scope="session")
def synth():
    import importlib
    return importlib.import_module("binary-image-compression_b200.synth")
