import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from m17_oracles import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    from m17_oracles import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref not built (reference sources absent)")
    return Ref()


@pytest.fixture(scope="session")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import m17_sdr_b200 as m
    m.build()
    c = m.Context(0)
    yield c
    c.close()
