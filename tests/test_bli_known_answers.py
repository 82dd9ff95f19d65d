"""
The reference's known-answer vectors for its barycentric Lagrange interpolator
(/root/reference/src/testing/barylagrint.cpp:32-145, run by testing/test_numerics.py:94-95),
restated for the oracle's BLI and for the product's refinement / coefficient / evaluation
kernels (C ABI rhfq_bli_known_answer).  Same functions, constructor arguments, evaluation grid
(897389 points on [0, 5]) and tolerances as the reference's test.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle

XMAX = 5.0
NGRID = 897389
I = np.arange(NGRID, dtype=np.float64)
X = I * XMAX / (NGRID - 1)                      # barylagrint.cpp:46-48
XB = (NGRID - 1 - I) * XMAX / (NGRID - 1)
EPS4 = np.finfo(np.float64).eps ** 0.25         # default tol_rel, barylagrint.hpp:108

# (kind, ctor arguments as in the reference's test, reference function)
CASES = {
    "exp": (0, dict(tol_rel=1e-8, tol_abs=0.0, fmin=0.0, fmax=1.0, max_refinements=6),
            lambda x: np.exp(-x)),
    "sin2": (1, dict(tol_rel=1e-10, tol_abs=0.0, fmin=0.0, fmax=1.0, max_refinements=8),
             lambda x: np.sin(x) ** 2),
    "heaviside": (2, dict(tol_rel=1e-11, tol_abs=1e-11, fmin=0.0, fmax=1.0, max_refinements=9),
                  lambda x: (x > 2.0).astype(float)),
    "gauss": (3, dict(tol_rel=1e-8, tol_abs=0.0, fmin=-np.inf, fmax=np.inf, max_refinements=6),
              lambda x: np.exp(-0.5 * (x - 2.0) ** 2 / 0.25)),
}


def check_answers(name, y, abs_floor=0.0, sub=slice(None)):
    """The assertions of barylagrint.cpp:49-54, 72-88, 105-123, 139-147.

    abs_floor: the reference's sin^2 tolerance max(1e-7, 1e-16 x / y_ref) * y_ref is an ABSOLUTE
    1e-16 x around the zeros of the function; the reference meets it by accumulating the
    barycentric sums in 80-bit long double (barylagrint.hpp:58-71, restated in the oracle).  The
    GPU has no such type: its sums (Clenshaw or barycentric) are good to a few ulp of max |f|,
    i.e. the product is held to max(reference tolerance, abs_floor = 2e-15) there."""
    X = globals()["X"][sub]
    y_ref = CASES[name][2](X)
    if name == "exp":
        assert np.all(np.abs(y - y_ref) <= 1e-8 * y_ref)
    elif name == "sin2":
        with np.errstate(divide="ignore", invalid="ignore"):
            tol = np.maximum(1e-7, 1e-16 / (y_ref / X))
        bad = (np.abs(y - y_ref) > np.maximum(tol * y_ref, abs_floor)) & (y_ref > 0)
        assert not bad.any(), (X[bad][:5], y[bad][:5], y_ref[bad][:5])
    elif name == "heaviside":
        # the reference only reports a failure here ("expected due to the harsh discontinuity")
        assert np.all(np.isfinite(y)) and y.min() >= 0.0 and y.max() <= 1.0
        far = np.abs(X - 2.0) > 1.0
        assert np.abs(y - y_ref)[far].max() < 0.2
    else:
        assert np.all(np.abs(y - y_ref) <= 1e-7 * y_ref)


def product_bli(api, name, barycentric=False):
    kind, kw, _ = CASES[name]
    out = np.empty_like(X)
    info = np.zeros(2, dtype=np.int64)
    err = C.create_string_buffer(512)
    rc = api.bli_known_answer(kind, 0.0, XMAX, kw["tol_rel"], kw["tol_abs"], kw["fmin"], kw["fmax"],
                              kw["max_refinements"], X.ctypes.data_as(C.c_void_p), X.size,
                              out.ctypes.data_as(C.c_void_p), int(barycentric),
                              info.ctypes.data_as(C.c_void_p), 0, err, 512)
    assert rc == 0, err.value.decode()
    return out, int(info[0])


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_bli_known_answers(name):
    kind, kw, _ = CASES[name]
    prec = "long double" if name == "heaviside" else "double"      # barylagrint.cpp:101-102
    # (the Heaviside interpolant keeps 4097 samples and the reference only REPORTS its deviations:
    #  every 16th grid point keeps the long-double evaluation at seconds on one core)
    sub = slice(None, None, 16) if name == "heaviside" else slice(None)
    y, samples, evals = oracle.bli_known_answer(kind, X[sub], XB[sub], 0.0, XMAX, precision=prec, **kw)
    assert samples >= 9 and evals >= samples
    check_answers(name, y, sub=sub)


@pytest.mark.parametrize("name", list(CASES))
def test_engine_bli_known_answers_emu(emu_api, name):
    """The product's stage bodies on the host emulation: same answers, and the refinement stops
    at the same number of samples as the oracle's restatement of barylagrint.hpp:576-668."""
    kind, kw, _ = CASES[name]
    y, samples = product_bli(emu_api, name)
    check_answers(name, y, abs_floor=2e-15)
    _, s_orc, _ = oracle.bli_known_answer(kind, X[:2], XB[:2], 0.0, XMAX, **kw)
    assert samples == s_orc
    if name != "heaviside":
        yb, _ = product_bli(emu_api, name, barycentric=True)
        assert np.max(np.abs(y - yb)) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_device_bli_known_answers(name):
    from reheatfunq_b200 import _lib
    kind, kw, _ = CASES[name]
    y, samples = product_bli(_lib.api(), name)
    check_answers(name, y, abs_floor=2e-15)
    _, s_orc, _ = oracle.bli_known_answer(kind, X[:2], XB[:2], 0.0, XMAX, **kw)
    assert samples == s_orc
    yb, _ = product_bli(_lib.api(), name, barycentric=True)
    check_answers(name, yb, abs_floor=4e-15)
