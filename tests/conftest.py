import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def have_ref():
    from oracle.pyoracle import have_reference
    return have_reference()


@pytest.fixture(scope="session")
def engine():
    """One GPU engine context for the session (fails loudly when CUDA is unusable)."""
    from minotaur_b200.engine import GpuBoundEngine
    eng = GpuBoundEngine(0)
    yield eng
    eng.close()
