import os
import sys
import ctypes as C

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # Build the checker (oracle) and the test-only host helpers.  The product
    # CUDA library is built by __graft_entry__.build(); tests never build a
    # CPU version of the product.
    import __graft_entry__ as g
    g.build_oracle()
    g.build_test_helpers()


DEFAULT_PRIOR = (2.522017292522833465, 15.37301672440559130,
                 0.2184765374862465137, 0.2184765374862465137)   # prior.py:731-734


def ls1980(x, d=14e3, L=160e3):
    """LS1980 anomaly of a straight fault (src/anomaly/ls1980.cpp:103-127), c in W/m^2 per W."""
    dist = np.abs(x)
    return 2 / (np.pi * d * L) * (1 - dist / d * np.arctan(d / dist))


def gamma_dataset(N, seed, alpha=40.0, beta=0.4, P=100e6):
    """Recipe of testing/test_anomaly_posterior.py:43-75 (one anomaly, dmin = 0)."""
    rng = np.random.default_rng(seed)
    Q = np.sort(rng.gamma(alpha, size=N) / beta)
    x = 160e3 * (rng.random(N) - 0.5)
    c = ls1980(x)
    q = Q + 1e3 * P * c
    return q, c * 1e3      # posterior.py:286-288 passes c * 1e3


@pytest.fixture(scope="session")
def emu_api():
    from reheatfunq_b200._lib import Api
    so = os.path.join(ROOT, "tests", "hostmath", "libengine_emu.so")
    return Api(C.CDLL(so), "emu_rhfq_")


@pytest.fixture(scope="session")
def hostmath():
    so = os.path.join(ROOT, "tests", "hostmath", "libhostmath.so")
    H = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    H.hm_sizeof_locals.restype = C.c_int
    H.hm_locals.argtypes = [dp, dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                            C.c_double, C.c_void_p, dp]
    H.hm_log_scale.argtypes = [C.c_void_p, dp, dp]
    H.hm_ymax.argtypes = [C.c_void_p, dp]
    H.hm_log_outer.argtypes = [C.c_void_p, dp, dp, C.c_int, dp, dp]
    H.hm_log_large_z.argtypes = [C.c_void_p, dp, C.c_int, C.c_int, dp, dp]
    for f in (H.hm_lgdiff, H.hm_digdiff, H.hm_trigdiff):
        f.restype = C.c_double
    H.hm_lgdiff.argtypes = [C.c_double] * 5
    H.hm_digdiff.argtypes = [C.c_double] * 3
    H.hm_trigdiff.argtypes = [C.c_double] * 3
    return H


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
