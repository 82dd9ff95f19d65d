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
        # (2e-9 before the panels were cut at the |C1| kink)
        assert np.abs(oz - refz).max() < 5e-10


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


def test_lgamma_table_reads(hostmath):
    """The binade-uniform lgamma-difference table (rhfq_math.cuh: GTAB) against mpmath over
    its whole range [1, 1024): H by quintic Hermite to the rounding of H, H' and H'' by cubic
    Hermite well inside what the Newton searches (tolerance 1e-9) need; cell boundaries,
    binade boundaries and the last cell included."""
    import ctypes as C
    import mpmath as mp
    H = hostmath
    mp.mp.dps = 40
    H.hm_gtab_doubles.restype = C.c_int
    H.hm_tetragamma.restype = C.c_double
    H.hm_tetragamma.argtypes = [C.c_double]
    for x in (0.7, 3.0, 11.99, 12.0, 150.0):
        ref = mp.polygamma(2, x)
        assert abs(H.hm_tetragamma(x) - ref) <= 2e-12 * abs(ref)
    nd = H.hm_gtab_doubles()
    rng = np.random.default_rng(5)
    a = np.concatenate([np.exp(rng.uniform(0, np.log(1024.0), 300)),
                        [1.0, 1.0 + 2.0 ** -7, 2.0, 2.0 - 1e-12, 12.0, 512.0, 1024.0 * (1 - 1e-16),
                         1023.999]])
    a = a[(a >= 1.0) & (a < 1024.0)]
    d = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
    for n, v in ((10.2, 10.2), (100.5, 100.2), (1000.2, 1000.2), (3.0, 2.5), (30.0, 41.0)):
        tab = np.zeros(nd)
        H.hm_gtab_build(C.c_double(n), C.c_double(v), d(tab))
        out = np.zeros((a.size, 3))
        H.hm_gtab_eval(d(tab), d(a), C.c_int(a.size), d(out))
        for ai, (h0, h1, h2) in zip(a, out):
            am = mp.mpf(float(ai))
            r0 = mp.loggamma(v * am) - n * mp.loggamma(am) - am * v * mp.log(v)
            r1 = v * mp.digamma(v * am) - n * mp.digamma(am) - v * mp.log(v)
            r2 = v * v * mp.polygamma(1, v * am) - n * mp.polygamma(1, am)
            # rounding of the cancelling Stirling terms the records are computed from
            scale = max(abs(r0), 1.5 * n * (ai + 12) * max(1.0, np.log(ai + 12)))
            assert abs(h0 - r0) <= 5e-16 * scale + 1e-13, (n, v, ai, h0, float(r0))
            # cubic Hermite: h^4 H^(5) / 384 with h <= a / 128, H^(5) ~ 6 (n - v) / a^4 + 12 (n - 1) / a^5
            assert abs(h1 - r1) <= 6e-11 * abs(n - v) + 2e-10 * n / ai + 1e-12 * max(1, abs(r1)), (n, v, ai)
            assert abs(h2 - r2) <= 1e-7 * max(abs(r2), n / ai ** 2), (n, v, ai)
