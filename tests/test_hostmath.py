"""
The product's fixed-node device math (reheatfunq_b200/csrc/rhfq_math.cuh,
compiled for the host by tests/hostmath) against the long-double oracle's
adaptive quadrature: the hot inner kernel of the path, checked without a GPU.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import gamma_dataset, DEFAULT_PRIOR
from oracle.oracle import OraclePosterior

dp = C.POINTER(C.c_double)


def d(a):
    return a.ctypes.data_as(dp)


CASES = [(2, 8, 40, .4), (3, 9, 40, .4), (5, 7, 40, .4), (10, 1, 40, .4), (10, 5, 200, 200 / 68.3),
         (25, 29189, 40, .4), (100, 6, 10, 10 / 68.3), (100, 3, 40, .4), (1000, 4, 40, .4)]


@pytest.mark.parametrize("N,seed,alpha,beta", CASES)
def test_log_outer_and_large_z(hostmath, N, seed, alpha, beta):
    H = hostmath
    q, c = gamma_dataset(N, seed, alpha, beta)
    post = OraclePosterior([q], [c], [1.0], *DEFAULT_PRIOR, 1.0, 1e-8, pdf_algorithm="explicit",
                           precision="long double")
    loc = post.locals()
    zz = np.concatenate([np.linspace(0, 0.99, 34), 1 - np.logspace(-2.5, -6, 8)])
    zz = zz[zz <= loc["ztrans"]]
    ref = post.log_outer(0, zz) + loc["log_scale"]
    Lb = C.create_string_buffer(H.hm_sizeof_locals())
    k = np.zeros(N)
    assert H.hm_locals(d(q), d(c), N, *DEFAULT_PRIOR, 1.0, Lb, d(k)) == 0
    out = np.zeros_like(zz)
    info = np.zeros(4 * len(zz))
    assert H.hm_log_outer(Lb, d(k), d(zz), len(zz), d(out), d(info)) == 0
    # relative error of the integral = absolute error of its log
    assert np.abs(out - ref).max() < 2e-10
    assert info.reshape(-1, 4)[:, 3].max() <= 2 * 48 + 30     # bounded work per node
    # large-z branch, integrated and not
    ys = np.logspace(-14, 0, 15) * loc["ymax"]
    for integ in (0, 1):
        refz = post.log_large_z(0, ys, bool(integ)) + loc["log_scale"]
        oz = np.zeros_like(ys)
        assert H.hm_log_large_z(Lb, d(ys), len(ys), integ, d(oz), None) == 0
        assert np.abs(oz - refz).max() < 2e-9   # |C1| kink, see DESIGN.md


@pytest.mark.parametrize("N,seed", [(10, 1), (100, 3)])
def test_locals_logscale_ymax(hostmath, N, seed):
    H = hostmath
    q, c = gamma_dataset(N, seed)
    post = OraclePosterior([q], [c], [1.0], *DEFAULT_PRIOR, 1.0, 1e-8, pdf_algorithm="explicit",
                           precision="double")
    loc = post.locals()
    Lb = C.create_string_buffer(H.hm_sizeof_locals())
    k = np.zeros(N)
    assert H.hm_locals(d(q), d(c), N, *DEFAULT_PRIOR, 1.0, Lb, d(k)) == 0
    ls = C.c_double()
    assert H.hm_log_scale(Lb, d(k), C.byref(ls)) == 0
    assert abs(ls.value - loc["log_scale"]) < 1e-9
    ym = C.c_double()
    assert H.hm_ymax(Lb, C.byref(ym)) == 0
    # y_max comes out of a 2-bit TOMS748 bracket (large_z.hpp:340): factor < 1.5
    assert 1 / 1.5 < ym.value / loc["ymax"] < 1.5


def test_special_function_series(hostmath):
    import mpmath as mp
    H = hostmath
    mp.mp.dps = 40
    for a in (0.3, 1.0, 5.0, 11.9, 12.0, 47.0, 300.0, 4e4):
        for n, v in ((10.2, 10.2), (100.5, 100.2), (1000.2, 1000.2), (3.0, 2.5)):
            C_ = -v * mp.log(v) + 0.37
            ref = mp.loggamma(v * a) - n * mp.loggamma(a) + C_ * a
            got = H.hm_lgdiff(a, n, v, float(C_), float(mp.log(v)))
            # rounding scales with the size of the (cancelling) Stirling terms
            scale = max(abs(ref), 1.5 * n * (a + 12) * max(1.0, np.log(a + 12)))
            assert abs(got - ref) <= 5e-16 * scale + 1e-13
            rd = v * mp.digamma(v * a) - n * mp.digamma(a)
            assert abs(H.hm_digdiff(a, n, v) - rd) <= 1e-12 * max(1, abs(rd), float(v * abs(mp.log(v * a))))
            rt = v * v * mp.polygamma(1, v * a) - n * mp.polygamma(1, a)
            assert abs(H.hm_trigdiff(a, n, v) - rt) <= 1e-11 * max(abs(rt), float(n * mp.polygamma(1, a)))
