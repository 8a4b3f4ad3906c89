import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port_lib():
    """The plain-C restatement (oracle/lcp_oracle.c), built on demand -- the checker, never the product."""
    from oracle import pyoracle
    pyoracle.build_port()
    return pyoracle


@pytest.fixture(scope="session")
def small_problem():
    from physimglobalpose_b200 import synth
    prob = synth.make_problem(500, 20000, 0.01, seed=7)
    T = synth.make_hypotheses(prob, 400, seed=11)
    return prob, T


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from physimglobalpose_b200.engine import PoseEngine
    e = PoseEngine(0)
    yield e
    e.close()
