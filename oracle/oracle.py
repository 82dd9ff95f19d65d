"""
ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front end of ``liborc.so`` (the CPU restatement of the reference's
algorithm, see ``orc_posterior.hpp``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package
``reheatfunq_b200`` never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    if force or not os.path.exists(so):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(so)
        L.orc_posterior_create.restype = C.c_void_p
        L.orc_posterior_create.argtypes = [C.c_int, _dp, _dp, _ip, _dp, C.c_int64,
                                           C.c_double, C.c_double, C.c_double, C.c_double,
                                           C.c_double, C.c_double, C.c_int, C.c_int,
                                           C.c_char_p, C.c_size_t]
        L.orc_posterior_create_ymax.restype = C.c_void_p
        L.orc_posterior_create_ymax.argtypes = [C.c_int, _dp, _dp, _ip, _dp, C.c_int64,
                                                C.c_double, C.c_double, C.c_double, C.c_double,
                                                C.c_double, C.c_double, C.c_int, C.c_int, _dp,
                                                C.c_char_p, C.c_size_t]
        L.orc_posterior_destroy.argtypes = [C.c_void_p]
        L.orc_posterior_qmax.restype = C.c_double
        L.orc_posterior_qmax.argtypes = [C.c_void_p]
        for name in ("pdf", "cdf", "tail", "tail_quantiles"):
            f = getattr(L, "orc_posterior_" + name)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_char_p, C.c_size_t]
        L.orc_posterior_locals.argtypes = [C.c_void_p, C.c_int64, _dp]
        L.orc_posterior_log_outer.argtypes = [C.c_void_p, C.c_int64, _dp, C.c_int64, _dp,
                                              C.c_char_p, C.c_size_t]
        L.orc_posterior_log_large_z.argtypes = [C.c_void_p, C.c_int64, _dp, C.c_int64, _dp,
                                                C.c_int, C.c_char_p, C.c_size_t]
        L.orc_posterior_bli.restype = C.c_int64
        L.orc_posterior_bli.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int64, _dp]
        L.orc_posterior_saq_leaves.restype = C.c_int64
        L.orc_posterior_saq_leaves.argtypes = [C.c_void_p, _dp, C.c_int64]
        L.orc_counters.argtypes = [_ip, _ip, C.c_int]
        L.orc_resilience.restype = C.c_int
        L.orc_resilience.argtypes = [C.c_int, _dp, C.c_int64, C.c_int64, C.c_double, _dp, C.c_int,
                                     _dp, C.c_double, C.c_double, C.c_uint64, C.c_int, C.c_int,
                                     _dp, _ip, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int]
        L.orc_resilience_levels.restype = C.c_int
        L.orc_resilience_levels.argtypes = [C.c_int, _dp, C.c_int64, C.c_int64, C.c_double, _dp,
                                            C.c_int, _dp, C.c_double, C.c_double, C.c_uint64,
                                            C.c_int, C.c_int, _dp, _ip, C.c_void_p]
        L.orc_bli_known_answer.restype = C.c_int
        L.orc_bli_known_answer.argtypes = [C.c_int, C.c_int] + [C.c_double] * 6 + [
            C.c_int, _dp, _dp, C.c_int64, _dp, _ip, C.c_char_p, C.c_size_t]
        L.orc_gcp_ln_phi.argtypes = [C.c_double] * 5 + [_dp]
        L.orc_gcp_kl.argtypes = [C.c_double] * 9 + [_dp]
        L.orc_gcp_kl_batch_max.argtypes = [_dp, C.c_int64, _dp, C.c_int64, C.c_double, _dp]
        L.orc_gcp_predictive_pdf.argtypes = [_dp, C.c_int64] + [C.c_double] * 5 + [_dp]
        L.orc_gcp_predictive_pdf_common_norm.argtypes = [_dp, C.c_int64, _dp, C.c_int64,
                                                         C.c_double, _dp]
        L.orc_gcp_predictive_cdf.argtypes = [_dp, C.c_int64, _dp, C.c_int64, C.c_double, _dp]
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_get_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


_ALG = {"explicit": 0, "barycentric_lagrange": 1, "adaptive_simpson": 2}
LOCALS_FIELDS = ("lp ls n v amin Qmax h0 h1 h2 h3 w lh0 l1p_w lv log_scale ymax ztrans S "
                 "taylor norm weight global_norm global_Qmax imax z_levels z_err z_l1").split()


class OraclePosterior:
    """Mirror of ``CppAnomalyPosterior`` (postbackend.pyx:167-366) on the CPU oracle.

    ``precision`` here names the *arithmetic type actually used* ('double' or
    'long double'), not the reference's (swapped) keyword.
    """

    def __init__(self, qij, cij, wi, p, s, n, v, amin, rtol,
                 pdf_algorithm="barycentric_lagrange", bli_max_refinements=7,
                 precision="double", ymax_override=None):
        """``ymax_override`` (tests only): one y_max per sample set (<= 0: compute); see
        ``LocalsAndNorm`` in orc_posterior.hpp."""
        L = lib()
        q = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.float64) for x in qij]))
        c = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.float64) for x in cij]))
        off = np.zeros(len(qij) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(x) for x in qij])
        w = np.ascontiguousarray(np.asarray(wi, dtype=np.float64))
        err = C.create_string_buffer(512)
        yo = None
        if ymax_override is not None:
            yo = np.ascontiguousarray(np.asarray(ymax_override, dtype=np.float64))
            if yo.shape != (len(qij),):
                raise ValueError("ymax_override needs one value per sample set")
        self._h = L.orc_posterior_create_ymax(
            {"double": 0, "long double": 1}[precision], _d(q), _d(c),
            off.ctypes.data_as(_ip), _d(w), len(qij), p, s, n, v, amin, rtol,
            _ALG[pdf_algorithm], bli_max_refinements, _d(yo) if yo is not None else None, err, 512)
        if not self._h:
            raise RuntimeError(err.value.decode())
        self._M = len(qij)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_posterior_destroy(self._h)
            self._h = None

    def Qmax(self):
        return lib().orc_posterior_qmax(self._h)

    def _call(self, name, x):
        x = np.array(x, dtype=np.float64, copy=True).ravel()
        err = C.create_string_buffer(512)
        rc = getattr(lib(), "orc_posterior_" + name)(self._h, _d(x), x.size, err, 512)
        if rc != 0:
            raise RuntimeError(err.value.decode())
        return x

    def pdf(self, PH):
        return self._call("pdf", PH)

    def cdf(self, PH):
        return self._call("cdf", PH)

    def tail(self, PH):
        return self._call("tail", PH)

    def tail_quantiles(self, t):
        return self._call("tail_quantiles", t)

    def locals(self, l=0):
        out = np.zeros(27)
        lib().orc_posterior_locals(self._h, l, _d(out))
        return dict(zip(LOCALS_FIELDS, out))

    def log_outer(self, l, z):
        z = np.ascontiguousarray(z, dtype=np.float64)
        out = np.empty_like(z)
        err = C.create_string_buffer(512)
        rc = lib().orc_posterior_log_outer(self._h, l, _d(z), z.size, _d(out), err, 512)
        if rc != 0:
            raise RuntimeError(err.value.decode())
        return out

    def log_large_z(self, l, y, y_integrated=False):
        y = np.ascontiguousarray(y, dtype=np.float64)
        out = np.empty_like(y)
        err = C.create_string_buffer(512)
        rc = lib().orc_posterior_log_large_z(self._h, l, _d(y), y.size, _d(out),
                                             int(y_integrated), err, 512)
        if rc != 0:
            raise RuntimeError(err.value.decode())
        return out

    def bli(self, which):
        cap = 5000
        x = np.empty(cap)
        f = np.empty(cap)
        info = np.zeros(4)
        ns = lib().orc_posterior_bli(self._h, which, _d(x), _d(f), cap, _d(info))
        return x[:ns].copy(), f[:ns].copy(), dict(refinements=int(info[0]),
                                                  evaluations=int(info[1]),
                                                  err_abs=info[2], err_rel=info[3])

    def saq_leaves(self):
        cap = 1 << 22
        x = np.empty(cap)
        nl = lib().orc_posterior_saq_leaves(self._h, _d(x), cap)
        return x[:min(nl, cap)].copy()


def counters(reset=False):
    a = C.c_int64()
    b = C.c_int64()
    lib().orc_counters(C.byref(a), C.byref(b), int(reset))
    return a.value, b.value


def resilience(dist_kind, dist_params, N, M, P_MW, quantile, prior, amin, tol, seed,
               nthread=1, precision="long double", return_data=False, data_only=False):
    """Oracle of src/resilience/resilience.cpp: returns res[2, Nq, M], failures[2][, data]."""
    L = lib()
    quantile = np.ascontiguousarray(quantile, dtype=np.float64)
    params = np.ascontiguousarray(dist_params, dtype=np.float64)
    prior = np.ascontiguousarray(prior, dtype=np.float64)
    Nq = quantile.shape[0]
    out = np.zeros((2, Nq, M))
    failures = np.zeros(2, dtype=np.int64)
    dumps = [None] * 4
    if return_data or data_only:
        dumps = [np.zeros((M, N)) for _ in range(4)]
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = L.orc_resilience(int(dist_kind), _d(params), int(N), int(M), float(P_MW), _d(quantile),
                          Nq, _d(prior), float(amin), float(tol), int(seed), int(nthread),
                          {"double": 0, "long double": 1}[precision], _d(out),
                          failures.ctypes.data_as(_ip), ptr(dumps[0]), ptr(dumps[1]),
                          ptr(dumps[2]), ptr(dumps[3]), int(data_only))
    if rc == 2:
        raise RuntimeError("oracle resilience failed")
    if return_data or data_only:
        return out, failures, dict(q=dumps[0], c=dumps[1], u=dumps[2], q0=dumps[3])
    return out, failures


def resilience_levels(dist_kind, dist_params, N, M, P_MW, quantile, prior, amin, tol, seed,
                      nthread=1, precision="long double"):
    """`resilience` plus the number of samples each barycentric interpolant kept:
    returns res[2, Nq, M], failures[2], samples[M, 4] (informed low/high, flat low/high)."""
    L = lib()
    quantile = np.ascontiguousarray(quantile, dtype=np.float64)
    params = np.ascontiguousarray(dist_params, dtype=np.float64)
    prior = np.ascontiguousarray(prior, dtype=np.float64)
    Nq = quantile.shape[0]
    out = np.zeros((2, Nq, M))
    failures = np.zeros(2, dtype=np.int64)
    samples = np.zeros((M, 4), dtype=np.int32)
    rc = L.orc_resilience_levels(int(dist_kind), _d(params), int(N), int(M), float(P_MW),
                                 _d(quantile), Nq, _d(prior), float(amin), float(tol), int(seed),
                                 int(nthread), {"double": 0, "long double": 1}[precision], _d(out),
                                 failures.ctypes.data_as(_ip), samples.ctypes.data_as(C.c_void_p))
    if rc == 2:
        raise RuntimeError("oracle resilience failed")
    return out, failures, samples


def bli_known_answer(kind, x, xmax_minus_x, xmin=0.0, xmax=5.0, tol_rel=None, tol_abs=0.0,
                     fmin=-np.inf, fmax=np.inf, max_refinements=6, precision="double"):
    """The oracle's interpolator on the reference's known-answer functions
    (src/testing/barylagrint.cpp:32-145).  Returns (values, samples kept, evaluations)."""
    if tol_rel is None:
        tol_rel = np.finfo(np.float64).eps ** 0.25      # barylagrint.hpp:108
    x = np.ascontiguousarray(x, dtype=np.float64)
    xb = np.ascontiguousarray(xmax_minus_x, dtype=np.float64)
    out = np.empty_like(x)
    info = np.zeros(2, dtype=np.int64)
    err = C.create_string_buffer(512)
    rc = lib().orc_bli_known_answer(int(kind), {"double": 0, "long double": 1}[precision], xmin,
                                    xmax, tol_rel, tol_abs, fmin, fmax, int(max_refinements),
                                    _d(x), _d(xb), x.size, _d(out), info.ctypes.data_as(_ip),
                                    err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return out, int(info[0]), int(info[1])


def gcp_ln_phi(lp, s, n, v, amin=1.0):
    out = C.c_double()
    if lib().orc_gcp_ln_phi(lp, s, n, v, amin, C.byref(out)) != 0:
        raise RuntimeError("oracle ln_Phi failed")
    return out.value


def gcp_kl(lp, s, n, v, lp_ref, s_ref, n_ref, v_ref, amin=1.0):
    """GammaConjugatePriorBase::kullback_leibler with CONDITION_ERROR semantics."""
    out = C.c_double()
    rc = lib().orc_gcp_kl(lp, s, n, v, lp_ref, s_ref, n_ref, v_ref, amin, C.byref(out))
    if rc == 1:
        raise RuntimeError("oracle KL failed")
    if rc == 2:
        raise RuntimeError("Large condition number in GammaConjugatePriorBase::kullback_leibler")
    return out.value


def gcp_kl_batch_max(params, refs, amin=1.0):
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 4)
    refs = np.ascontiguousarray(refs, dtype=np.float64).reshape(-1, 4)
    out = np.empty(refs.shape[0])
    lib().orc_gcp_kl_batch_max(_d(params), params.shape[0], _d(refs), refs.shape[0], amin, _d(out))
    return out


def gcp_predictive_pdf(q, lp, s, n, v, amin=1.0):
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = np.empty_like(q)
    if lib().orc_gcp_predictive_pdf(_d(q), q.size, lp, s, n, v, amin, _d(out)) != 0:
        raise RuntimeError("oracle predictive pdf failed")
    return out


def gcp_predictive_pdf_common_norm(q, params, amin=1.0):
    q = np.ascontiguousarray(q, dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 4)
    out = np.empty_like(q)
    if lib().orc_gcp_predictive_pdf_common_norm(_d(q), q.size, _d(params), params.shape[0], amin,
                                                _d(out)) != 0:
        raise RuntimeError("oracle predictive pdf failed")
    return out


def gcp_predictive_cdf(q, params, amin=1.0):
    """posterior_predictive_cdf (one parameter row) / _cdf_common_norm (several rows)."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 4)
    out = np.empty_like(q)
    if lib().orc_gcp_predictive_cdf(_d(q), q.size, _d(params), params.shape[0], amin, _d(out)) != 0:
        raise RuntimeError("oracle predictive cdf failed")
    return out
