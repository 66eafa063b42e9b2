import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "order_last: run after every other test (tests whose cuda variant has not had a GPU "
                                       "run yet: a surprise in one of them must not hide the established ones under -x)")


def pytest_collection_modifyitems(config, items):
    last = [it for it in items if it.get_closest_marker("order_last") is not None]
    if last:
        first = [it for it in items if it.get_closest_marker("order_last") is None]
        items[:] = first + last


@pytest.fixture(params=[pytest.param("cuda", marks=pytest.mark.gpu), "simt"])
def engine(request):
    """Every test that uses it runs twice: on the GPU through the product library (marked gpu) and, in the CPU suite,
    through the SIMT-on-CPU emulator build of the same kernel sources (tests/engines.py)."""
    import engines

    with engines.running(request.param) as e:
        yield e


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False))
    return load
