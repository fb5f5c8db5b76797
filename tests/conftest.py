import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    return O.oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled into oracle/_ref/libref.so (built here when /root/reference exists)."""
    from oracle import oracle as O
    if not os.path.exists(O.REF_SO) and os.path.isdir("/root/reference/src"):
        O.build(ref=True)
    r = O.reference()
    if r is None:
        pytest.skip("oracle/_ref/libref.so not built (no /root/reference here)")
    return r


@pytest.fixture(scope="session")
def gpu_api():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cp_pfdr_graph_d1_b200 import api
    return api
