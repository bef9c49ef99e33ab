import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `pytest -m gpu`)")
    config.addinivalue_line("markers", "slow: larger corpora")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def g1_dump():
    from diagon_b200 import read_dump

    return read_dump(os.path.join(GOLDEN, "g1.dmp.gz"))


@pytest.fixture(scope="session")
def g2_dump():
    from diagon_b200 import read_dump

    return read_dump(os.path.join(GOLDEN, "g2.dmp.gz"))
