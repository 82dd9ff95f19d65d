"""
Drop-in for the Kullback-Leibler / normalisation entry points of
``reheatfunq.regional.backend`` (reheatfunq/regional/backend.pyx:271-272,
497-528, 563-734) on the B200 backend.  Same names, argument order and
exceptions; ``epsabs`` / ``epsrel`` are accepted and ignored like the
reference's C++ ignores them for these integrals (its quadrature tolerance is
the constant sqrt(eps), gamma_conjugate_prior.cpp:439).
"""
import ctypes as C

import numpy as np

from .. import _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def gcp_ln_Phi(lp, s, n, v, amin=1.0, *, device=0, _api=None):
    api = _api if _api is not None else _lib.api()
    params = np.array([[lp, s, n, v]], dtype=np.float64)
    out = np.empty(1)
    st = np.zeros(1, dtype=np.int32)
    err = C.create_string_buffer(512)
    rc = api.gcp_ln_phi(_ptr(params), 1, float(amin), _ptr(out), _ptr(st), device, err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    if st[0] != 0:
        raise RuntimeError("Could not compute the normalization constant of the gamma "
                           "conjugate prior.")
    return float(out[0])


def _ln_phi_many(params, amin, device=0, _api=None, chunk=1 << 22):
    """ln Phi of K parameter rows (lp, s, n, v): one launch per chunk of 4M rows."""
    api = _api if _api is not None else _lib.api()
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 4)
    K = params.shape[0]
    out = np.empty(K)
    st = np.zeros(K, dtype=np.int32)
    err = C.create_string_buffer(512)
    for a in range(0, K, chunk):
        b = min(K, a + chunk)
        pc = np.ascontiguousarray(params[a:b])
        oc = np.empty(b - a)
        sc = np.zeros(b - a, dtype=np.int32)
        rc = api.gcp_ln_phi(_ptr(pc), b - a, float(amin), _ptr(oc), _ptr(sc), device, err, 512)
        if rc != 0:
            raise RuntimeError(err.value.decode())
        out[a:b] = oc
        st[a:b] = sc
    return out, st


# Gauss-Legendre (16) on [-1, 1] for the predictive cdf
_GL16_X, _GL16_W = np.polynomial.legendre.leggauss(16)


def _predictive_cdf(q, lp, s, n, v, amin, what, device, _api):
    """
    Piecewise integration of the (mixture) predictive pdf between the sorted q nodes
    (gamma_conjugate_prior.cpp:996-1157).  The reference uses adaptive GK(7,15) with 5
    bisection levels at rel. tolerance 1e-9 per piece; here every piece is integrated
    non-adaptively with Gauss-Legendre-16 panels (8 uniform ones; geometric grading towards
    q = 0), all pdf nodes of all pieces in one batched ln Phi launch — at least as accurate.
    """
    q = np.asarray(q, dtype=np.float64)
    lp, s, n, v = (np.atleast_1d(np.asarray(x, dtype=np.float64)) for x in (lp, s, n, v))
    M, N = lp.shape[0], q.shape[0]
    if N == 0:
        return np.zeros(0)
    order = np.argsort(q, kind="stable")
    qs = q[order]
    ql = np.concatenate([[0.0], qs[:-1]])
    # panels per piece: 8 uniform ones; the piece that starts at q = 0 is graded
    # geometrically towards 0 (the pdf behaves like 1/|ln q| there), 2^-30 ... 1
    P = 8
    lo_l, hi_l, piece = [], [], []
    for i in range(N):
        if qs[i] <= 0.0 or qs[i] == ql[i]:
            continue
        if ql[i] <= 0.0:
            e = qs[i] * np.concatenate([[0.0], 2.0 ** np.arange(-30, 1)])
        else:
            e = np.linspace(ql[i], qs[i], P + 1)
        lo_l.append(e[:-1]); hi_l.append(e[1:]); piece.append(np.full(e.shape[0] - 1, i))
    if not lo_l:
        return np.zeros(N)
    lo = np.concatenate(lo_l); hi = np.concatenate(hi_l); piece = np.concatenate(piece)
    mid = 0.5 * (hi + lo)
    hw = 0.5 * (hi - lo)
    nodes = mid[:, None] + hw[:, None] * _GL16_X[None, :]                  # (panels, 16)
    wts = hw[:, None] * _GL16_W[None, :]
    qn = nodes.reshape(-1)
    ok = qn > 0
    base = np.column_stack([lp, s, n, v])
    rows = np.empty((M, int(ok.sum()), 4))
    rows[:, :, 0] = lp[:, None] + np.log(qn[ok])[None, :]
    rows[:, :, 1] = s[:, None] + qn[ok][None, :]
    rows[:, :, 2] = (n + 1)[:, None]
    rows[:, :, 3] = (v + 1)[:, None]
    lnphi, st = _ln_phi_many(np.concatenate([base, rows.reshape(-1, 4)]), amin, device, _api)
    if np.any(st != 0):
        raise RuntimeError("Error in %s: 'could not compute the normalization constant'." % what)
    mx = lnphi[:M].max()
    tot = np.exp(lnphi[:M] - mx).sum()
    pdf = np.zeros(qn.shape[0])
    pdf[ok] = np.exp(lnphi[M:].reshape(M, -1) - mx).sum(axis=0) / tot
    pieces = np.bincount(piece, weights=(pdf.reshape(nodes.shape) * wts).sum(axis=1), minlength=N)
    cs = np.cumsum(pieces)
    res = np.empty(N)
    res[order] = cs
    if cs[-1] > 1.0:
        res /= cs[-1]
    return res


def gamma_conjugate_prior_predictive_cdf(q, lp, s, n, v, amin, epsabs=0.0, epsrel=1e-10,
                                         inplace=False, *, device=0, _api=None):
    """backend.pyx:398-418."""
    res = _predictive_cdf(q, lp, s, n, v, amin, "posterior_predictive_cdf", device, _api)
    if inplace:
        q[...] = res
        return q
    return res


def gamma_conjugate_prior_predictive_cdf_batch(q, lp, s, n, v, amin, epsabs=0.0, epsrel=1e-10,
                                               out=None, *, device=0, _api=None):
    """backend.pyx:422-452: (M, N) cdfs, one per parameter set."""
    lp, s, n, v = (np.asarray(x, dtype=np.float64) for x in (lp, s, n, v))
    M = lp.shape[0]
    for name, arr in (("s", s), ("n", n), ("v", v)):
        if arr.shape[0] != M:
            raise RuntimeError("`lp` and `%s` need to be of same shape." % name)
    res = np.empty((M, len(q))) if out is None else out
    for i in range(M):
        res[i] = _predictive_cdf(q, lp[i], s[i], n[i], v[i], amin, "posterior_predictive_cdf_batch",
                                 device, _api)
    return res


def gamma_conjugate_prior_predictive_cdf_common_norm(q, lp, s, n, v, amin, epsabs=0.0,
                                                     epsrel=1e-10, out=None, *, device=0,
                                                     _api=None):
    """backend.pyx:455-494 / gamma_conjugate_prior.cpp:1054-1157."""
    lp, s, n, v = (np.asarray(x, dtype=np.float64) for x in (lp, s, n, v))
    M = lp.shape[0]
    for name, arr in (("s", s), ("n", n), ("v", v)):
        if arr.shape[0] != M:
            raise RuntimeError("`lp` and `%s` need to be of same shape." % name)
    res = _predictive_cdf(q, lp, s, n, v, amin, "posterior_predictive_cdf_common_norm", device, _api)
    if out is not None:
        out[...] = res
        return out
    return res


def gamma_conjugate_prior_predictive(q, lp, s, n, v, amin, inplace=False, *, device=0,
                                     _api=None):
    """
    Posterior predictive pdf of the heat flow q (backend.pyx:305-325;
    gamma_conjugate_prior.cpp:794-849):
    p(q) = Phi(lp + ln q, s + q, n + 1, v + 1) / Phi(lp, s, n, v); p(q <= 0) = 0.
    All q nodes are one batched ln Phi launch.
    """
    qa = np.asarray(q, dtype=np.float64)
    out = qa if inplace else np.empty_like(qa)
    pos = qa > 0
    rows = np.empty((int(pos.sum()) + 1, 4))
    rows[0] = (lp, s, n, v)
    rows[1:, 0] = lp + np.log(qa[pos])
    rows[1:, 1] = s + qa[pos]
    rows[1:, 2] = n + 1
    rows[1:, 3] = v + 1
    lnphi, st = _ln_phi_many(rows, amin, device, _api)
    if np.any(st != 0):
        raise RuntimeError("Error in posterior_predictive_pdf: 'could not compute the "
                           "normalization constant'.")
    res = np.zeros_like(qa)
    res[pos] = np.exp(lnphi[1:] - lnphi[0])
    out[...] = res
    return out


def gamma_conjugate_prior_predictive_batch(q, lp, s, n, v, amin, epsabs=0.0, epsrel=1e-10,
                                           out=None, *, device=0, _api=None):
    """backend.pyx:329-356: (M, N) predictive pdfs for M parameter sets."""
    q = np.asarray(q, dtype=np.float64)
    lp, s, n, v = (np.asarray(x, dtype=np.float64) for x in (lp, s, n, v))
    M, N = lp.shape[0], q.shape[0]
    for name, arr in (("s", s), ("n", n), ("v", v)):
        if arr.shape[0] != M:
            raise RuntimeError("`lp` and `%s` need to be of same shape." % name)
    res = np.empty((M, N)) if out is None else out
    pos = q > 0
    base = np.column_stack([lp, s, n, v])
    rows = np.empty((M, int(pos.sum()), 4))
    rows[:, :, 0] = lp[:, None] + np.log(q[pos])[None, :]
    rows[:, :, 1] = s[:, None] + q[pos][None, :]
    rows[:, :, 2] = (n + 1)[:, None]
    rows[:, :, 3] = (v + 1)[:, None]
    lnphi, st = _ln_phi_many(np.concatenate([base, rows.reshape(-1, 4)]), amin, device, _api)
    if np.any(st != 0):
        raise RuntimeError("Error in posterior_predictive_pdf_batch: 'could not compute the "
                           "normalization constant'.")
    res[...] = 0.0
    res[:, pos] = np.exp(lnphi[M:].reshape(M, -1) - lnphi[:M, None])
    return res


def gamma_conjugate_prior_predictive_common_norm(q, lp, s, n, v, amin, epsabs=0.0,
                                                 epsrel=1e-10, out=None, *, device=0,
                                                 _api=None):
    """backend.pyx:360-393 / gamma_conjugate_prior.cpp:926-990: mixture of the M predictive
    pdfs weighted by their normalisation constants."""
    q = np.asarray(q, dtype=np.float64)
    lp, s, n, v = (np.asarray(x, dtype=np.float64) for x in (lp, s, n, v))
    M, N = lp.shape[0], q.shape[0]
    for name, arr in (("s", s), ("n", n), ("v", v)):
        if arr.shape[0] != M:
            raise RuntimeError("`lp` and `%s` need to be of same shape." % name)
    res = np.empty(N) if out is None else out
    base = np.column_stack([lp, s, n, v])
    rows = np.empty((M, N, 4))
    with np.errstate(divide="ignore", invalid="ignore"):
        rows[:, :, 0] = lp[:, None] + np.log(q)[None, :]
    rows[:, :, 1] = s[:, None] + q[None, :]
    rows[:, :, 2] = (n + 1)[:, None]
    rows[:, :, 3] = (v + 1)[:, None]
    lnphi, st = _ln_phi_many(np.concatenate([base, rows.reshape(-1, 4)]), amin, device, _api)
    if np.any(st != 0):
        raise RuntimeError("Error in posterior_predictive_pdf_common_norm: 'could not compute "
                           "the normalization constant'.")
    mx = lnphi[:M].max()
    tot = np.exp(lnphi[:M] - mx).sum()
    res[...] = np.exp(lnphi[M:].reshape(M, N) - mx).sum(axis=0) / tot
    return res


def gamma_conjugate_prior_kullback_leibler_batch_max_many(params, refs, amin=1.0, *,
                                                          return_all=False, device=0,
                                                          _api=None):
    """Cost of K candidate references at once: out[k] = max_i KL(params_i || refs_k)."""
    api = _api if _api is not None else _lib.api()
    params = np.ascontiguousarray(params, dtype=np.float64)
    refs = np.ascontiguousarray(refs, dtype=np.float64)
    if params.ndim != 2 or params.shape[1] != 4:
        raise RuntimeError("`params` has to be of shape (N,4).")
    if refs.ndim != 2 or refs.shape[1] != 4:
        raise RuntimeError("`refs` has to be of shape (K,4).")
    K, N = refs.shape[0], params.shape[0]
    out = np.empty(K)
    kl_all = np.empty((K, N)) if return_all else None
    err = C.create_string_buffer(512)
    rc = api.gcp_kl_batch_max(_ptr(params), N, _ptr(refs), K, float(amin), _ptr(out),
                              None if kl_all is None else _ptr(kl_all), device, err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    if return_all:
        return out, kl_all
    return out


def gamma_conjugate_prior_kullback_leibler(lp, s, n, v, lp_ref, s_ref, n_ref, v_ref, amin=1.0,
                                           epsabs=0, epsrel=1e-10, *, device=0, _api=None):
    """backend.pyx:497-511."""
    out = gamma_conjugate_prior_kullback_leibler_batch_max_many(
        [[lp, s, n, v]], [[lp_ref, s_ref, n_ref, v_ref]], amin, device=device, _api=_api)
    if not np.isfinite(out[0]):
        raise RuntimeError("Could not compute the Kullback-Leibler distance (normalization "
                           "or quadrature failed).")
    return float(out[0])


def gamma_conjugate_prior_kullback_leibler_batch_max(params, lp_ref, s_ref, n_ref, v_ref,
                                                     amin=1.0, epsabs=0, epsrel=1e-10, *,
                                                     device=0, _api=None):
    """backend.pyx:514-528: +inf when any distance cannot be computed."""
    params = np.ascontiguousarray(params, dtype=np.float64)
    if params.ndim != 2 or params.shape[1] != 4:
        raise RuntimeError("`params` has to be of shape (N,4).")
    out = gamma_conjugate_prior_kullback_leibler_batch_max_many(
        params, [[lp_ref, s_ref, n_ref, v_ref]], amin, device=device, _api=_api)
    return float(out[0])


def gamma_conjugate_prior_logL(a, b, lp, s, n, v, amin=1.0, *, device=0, _api=None):
    """Log-likelihood of the gamma conjugate prior (backend.pyx:240-268)."""
    from scipy.special import gammaln
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape[0] != b.shape[0]:
        raise RuntimeError("Shapes of `a` and `b` do not match.")
    ll = -gcp_ln_Phi(lp, s, n, v, amin, device=device, _api=_api)
    return float(ll + ((v * a - 1.0) * np.log(b) + (a - 1.0) * lp - s * b - n * gammaln(a)).sum())


def gamma_conjugate_prior_bulk_log_p(a, b, lp, s, n, v, amin=1.0, *, device=0, _api=None):
    """Log-probability of the conjugate prior at each (a[i], b[i]) (backend.pyx:275-301)."""
    from scipy.special import gammaln
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape[0] != b.shape[0]:
        raise RuntimeError("Shapes of `a` and `b` do not match.")
    norm = gcp_ln_Phi(lp, s, n, v, amin, device=device, _api=_api)
    return (v * a - 1.0) * np.log(b) + (a - 1.0) * lp - s * b - n * gammaln(a) - norm


def gamma_mle(x, amin=1.0):
    """
    Gamma distribution maximum-likelihood estimate (backend.pyx:532-553,
    external/pdtoolbox/src/gamma.cpp:32-66: Minka's start + Newton).  Host side, O(N).
    """
    from scipy.special import digamma, polygamma
    x = np.asarray(x, dtype=np.float64)
    xm = x.mean()
    lxm = np.log(x).mean()
    s = np.log(xm) - lxm
    a = (3.0 - s + np.sqrt((s - 3.0) ** 2 + 24.0 * s)) / (12.0 * s)
    a1 = a - (np.log(a) - digamma(a) - s) / (1.0 / a - polygamma(1, a))
    i = 0
    while abs(a1 - a) > 1e-13 and i < 20:
        a = a1
        a1 = a - (np.log(a) - digamma(a) - s) / (1.0 / a - polygamma(1, a))
        a1 = max(a1, amin)
        i += 1
    return float(a1), float(a1 / xm)


def gamma_conjugate_prior_minimum_surprise_backend(gcp_updated_params, amin, bounds, kwargs,
                                                   verbose, cache=None, *, device=0, _api=None):
    """
    Backend of ``GammaConjugatePrior.minimum_surprise_estimate`` (backend.pyx:741-799): SHGO
    over (lp, s, ln(n/v - 1), v) of cost = max_i KL(sample_i || (lp, s, n, v)).  The cost is
    served by the B200 backend through ``GCPMSECache`` (one ln Phi integral + O(N) arithmetic
    per evaluation after the per-sample moments have been computed once).
    """
    from math import exp
    from scipy.optimize import shgo
    params = np.ascontiguousarray(gcp_updated_params, dtype=np.float64)
    if params.ndim != 2 or params.shape[1] != 4:
        raise RuntimeError("`gcp_updated_params` has to be of shape (n,4).")
    if params.shape[0] == 0:
        raise RuntimeError("`gcp_updated_params` is empty.")
    (lpmin, lpmax), (smin, smax), (lnvspmin, lnvspmax), (vmin, vmax) = bounds
    if cache is None or cache.size() == 0 and cache._params.shape[0] == 0:
        cache = GCPMSECache(params, amin, device=device, _api=_api)

    def cost(x):
        # SHGO's local minimiser may step out of bounds (backend.pyx:775-781)
        if x[0] < lpmin or x[0] > lpmax or x[1] < smin or x[1] > smax or x[3] < vmin \
                or x[3] > vmax or x[2] < lnvspmin or x[2] > lnvspmax:
            return np.inf
        return cache(x[0], x[1], x[3] * (1.0 + exp(x[2])), x[3])

    return shgo(cost, bounds=bounds, **kwargs)


class GCPMSECache:
    """
    Memo of the minimum-surprise cost function (backend.pyx:563-734,
    external/pdtoolbox/include/funccache.hpp:43-129): exact-key cache of
    cost(lp, s, n, v) = max_i KL(params_i || (lp, s, n, v)), picklable.
    ``evaluate_many`` is the batched entry the B200 backend adds.
    """

    def __init__(self, params, amin, *, device=0, _api=None):
        params = np.ascontiguousarray(params, dtype=np.float64)
        if params.ndim != 2 or params.shape[1] != 4:
            raise RuntimeError("`params` has to be of shape (N,4).")
        self._params = params.copy()
        self._amin = float(amin)
        self._cache = {}
        self._hits = 0
        self._misses = 0
        self._device = device
        self._api = _api

    def __call__(self, lp, s, n, v):
        key = (float(lp), float(s), float(n), float(v))
        val = self._cache.get(key)
        if val is not None:
            self._hits += 1
            return val
        self._misses += 1
        val = gamma_conjugate_prior_kullback_leibler_batch_max(
            self._params, *key, self._amin, device=self._device, _api=self._api)
        self._cache[key] = val
        return val

    def evaluate_many(self, refs):
        refs = np.ascontiguousarray(refs, dtype=np.float64).reshape(-1, 4)
        keys = [tuple(float(x) for x in r) for r in refs]
        todo = [i for i, k in enumerate(keys) if k not in self._cache]
        self._hits += len(keys) - len(todo)
        self._misses += len(todo)
        if todo:
            vals = gamma_conjugate_prior_kullback_leibler_batch_max_many(
                self._params, refs[todo], self._amin, device=self._device, _api=self._api)
            for i, val in zip(todo, vals):
                self._cache[keys[i]] = float(val)
        return np.array([self._cache[k] for k in keys])

    def hits(self):
        return self._hits

    def misses(self):
        return self._misses

    def hit_rate(self):
        tot = self._hits + self._misses
        return self._hits / tot if tot else 0.0

    def size(self):
        return len(self._cache)

    def __eq__(self, other):
        return (isinstance(other, GCPMSECache) and self._amin == other._amin
                and np.array_equal(self._params, other._params))

    @staticmethod
    def empty():
        return GCPMSECache(np.zeros((0, 4)), 1.0)

    def __reduce__(self):
        dump = np.array([k + (v,) for k, v in sorted(self._cache.items())]).reshape(-1, 5)
        return (_restore_cache, (self._params, self._amin, dump))


def _restore_cache(params, amin, dump):
    c = GCPMSECache(params, amin)
    for row in np.asarray(dump).reshape(-1, 5):
        c._cache[tuple(float(x) for x in row[:4])] = float(row[4])
    return c
