import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference at build time)")
    return Ref()


@pytest.fixture(scope="session")
def pkg():
    import obs_color_monitor_b200 as p
    return p


@pytest.fixture(scope="session")
def engine(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    eng = pkg.ScopeEngine()
    yield eng
    eng.close()
