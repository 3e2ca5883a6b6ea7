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
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def ehmm():
    """the product package; building it needs nvcc only, loading it needs no GPU"""
    so = os.path.join(ROOT, "epigrahmm_b200", "libehmm_b200.so")
    if not os.path.exists(so):
        from epigrahmm_b200 import build as B  # noqa: WPS433  (module import without loading the .so)
        B.build()
    import epigrahmm_b200
    return epigrahmm_b200


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
