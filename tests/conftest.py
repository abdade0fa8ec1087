"""Shared fixtures.  `-m "not gpu"` runs here on CPU; `-m gpu` needs a B200."""
import importlib
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.swo import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


@pytest.fixture(scope="session")
def swb():
    """The product package (directory name has a hyphen, hence importlib)."""
    return importlib.import_module("smith-waterman_b200")
