"""
Legacy (v1) functional API of ``reheatfunq.anomaly.bayes`` (reheatfunq/anomaly/bayes.pyx:
187-640) as thin adapters over the batched B200 posterior.

The legacy C++ (external/pdtoolbox/src/ziebarth2022a.cpp) evaluates the same model with its
own quadratures at ``dest_tol``; here the pdf is the explicit evaluation (agreement with the
reference's mpmath values ~1e-12) and cdf / tail / tail quantiles go through the
barycentric-Lagrange + Simpson-tree path with ``rtol = min(dest_tol, 1e-8)``.  Return types
follow the legacy functions (``numpy.longdouble`` arrays for pdf / cdf / tail).
``working_precision`` is accepted for compatibility; arithmetic is IEEE double on the GPU.
"""
import numpy as np

from ..batch import PosteriorBatch

_PRECISIONS = ("auto", "double", "long double", "float128", "dec50", "dec100")


def _support_float128():
    return False


def _support_dec50():
    return False


def _support_dec100():
    return False


def _check_precision(wp):
    if wp not in _PRECISIONS:
        raise RuntimeError("`working_precision` must be one of 'double', 'long double', "
                           "'float128', 'dec50', 'dec100', or 'auto'.")
    if wp in ("float128", "dec50", "dec100"):
        raise RuntimeError("The requested working precision is not available in this build.")


def _batch(Qi, Ci, p, s, n, v, amin, dest_tol, algorithm, device, _api):
    if len(Qi) != len(Ci):
        raise RuntimeError("Length of `Qi` and `Ci` needs to match!")
    sets = []
    for q, c in zip(Qi, Ci):
        q = np.ascontiguousarray(q, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        if q.shape[0] != c.shape[0]:
            raise RuntimeError("`qi` and `ci` have to be of same shape in marginal_posterior.")
        if q.shape[0] == 0:
            raise NotImplementedError("No data given - use prior (not implemented).")
        sets.append([(q, c)])
    B = PosteriorBatch(sets, [[1.0]] * len(sets), (p, s, n, v), amin=amin,
                       rtol=min(float(dest_tol), 1e-8), pdf_algorithm=algorithm, device=device,
                       api=_api)
    st = B.status()
    if np.any(st != 0):
        raise RuntimeError(B.status_message(int(st[st != 0][0])))
    return B


def _eval(what, x, p, s, n, v, Qi, Ci, amin, dest_tol, working_precision, device, _api):
    _check_precision(working_precision)
    x = np.ascontiguousarray(x, dtype=np.float64)
    algorithm = "explicit" if what == "pdf" else "barycentric_lagrange"
    B = _batch(Qi, Ci, p, s, n, v, amin, dest_tol, algorithm, device, _api)
    out, qst = getattr(B, what)(x, return_status=True)
    if np.any(qst != 0):
        raise RuntimeError(B.status_message(int(qst[qst != 0][0])))
    return out


def marginal_posterior_pdf(P_H, p, s, n, v, qi, ci, amin=1.0, dest_tol=1e-8,
                           working_precision="auto", *, device=0, _api=None):
    """bayes.pyx:187-227."""
    return _eval("pdf", P_H, p, s, n, v, [qi], [ci], amin, dest_tol, working_precision, device,
                 _api)[0].astype(np.longdouble)


def marginal_posterior_pdf_batch(P_H, p, s, n, v, Qi, Ci, amin=1.0, dest_tol=1e-8,
                                 working_precision="auto", *, device=0, _api=None):
    """bayes.pyx:231-284: one row per data set (independent posteriors)."""
    return _eval("pdf", P_H, p, s, n, v, Qi, Ci, amin, dest_tol, working_precision, device,
                 _api).astype(np.longdouble)


def marginal_posterior_cdf(P_H, p, s, n, v, qi, ci, amin=1.0, dest_tol=1e-8,
                           working_precision="auto", *, device=0, _api=None):
    """bayes.pyx:288-329."""
    return _eval("cdf", P_H, p, s, n, v, [qi], [ci], amin, dest_tol, working_precision, device,
                 _api)[0].astype(np.longdouble)


def marginal_posterior_cdf_batch(P_H, p, s, n, v, Qi, Ci, amin=1.0, dest_tol=1e-8,
                                 working_precision="auto", *, device=0, _api=None):
    """bayes.pyx:333-385."""
    return _eval("cdf", P_H, p, s, n, v, Qi, Ci, amin, dest_tol, working_precision, device,
                 _api).astype(np.longdouble)


def marginal_posterior_tail(P_H, p, s, n, v, qi, ci, amin=1.0, dest_tol=1e-8,
                            working_precision="auto", *, device=0, _api=None):
    """bayes.pyx:389-432."""
    return _eval("tail", P_H, p, s, n, v, [qi], [ci], amin, dest_tol, working_precision, device,
                 _api)[0].astype(np.longdouble)


def marginal_posterior_tail_batch(P_H, p, s, n, v, Qi, Ci, amin=1.0, dest_tol=1e-8,
                                  working_precision="auto", *, device=0, _api=None):
    """bayes.pyx:436-489."""
    return _eval("tail", P_H, p, s, n, v, Qi, Ci, amin, dest_tol, working_precision, device,
                 _api).astype(np.longdouble)


def marginal_posterior_tail_quantiles(quantiles, p, s, n, v, qi, ci, amin=1.0, dest_tol=1e-8,
                                      inplace=False, *, device=0, _api=None):
    """bayes.pyx:544-581."""
    res = _eval("tail_quantiles", quantiles, p, s, n, v, [qi], [ci], amin, dest_tol, "auto",
                device, _api)[0]
    if inplace:
        quantiles[...] = res
        return quantiles
    return res


def marginal_posterior_tail_quantiles_batch(quantiles, p, s, n, v, Qi, Ci, amin=1.0,
                                            dest_tol=1e-8, *, device=0, _api=None):
    """bayes.pyx:585-640: (len(Qi), Nquant)."""
    return _eval("tail_quantiles", quantiles, p, s, n, v, Qi, Ci, amin, dest_tol, "auto", device,
                 _api)
