import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


# BASELINE.json model constants (configs 1-3)
CFG1 = dict(S0=100.0, K=100.0, r=0.05, sigma=0.2, T=1.0, n=50)
CFG2 = dict(S0=100.0, r=0.05, xi=0.04, H=0.1, eta=1.9, rho=-0.9, n=252, dt=1.0 / 252.0)


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as O
    O.build(ref=os.path.isdir("/root/reference"))
    return O


@pytest.fixture(scope="session")
def port(orc):
    return orc.port()


@pytest.fixture(scope="session")
def ref(orc):
    if not orc.have_ref():
        pytest.skip("oracle/_ref/libmcp_ref.so not built (needs /root/reference)")
    return orc.ref()


@pytest.fixture(scope="session")
def engine():
    import montecarlooptionspricer_b200 as m
    eng = m.Engine(0)  # raises McpError without a GPU: gpu tests must never silently fall back
    yield eng
    eng.close()


def f32_draws(rng, shape):
    """iid N(0,1) rounded to fp32 -- the injected-draw convention of BASELINE config 2."""
    return rng.standard_normal(shape).astype(np.float32)
