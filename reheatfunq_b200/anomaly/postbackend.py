"""
``CppAnomalyPosterior`` — drop-in for the reference's Cython class
(reheatfunq/anomaly/postbackend.pyx:167-455) on top of the B200 C ABI.

Same constructor signature, method names, argument meaning, return types and
exception types.  Differences, all documented in DESIGN.md:
  * arithmetic is IEEE double on the GPU for every ``precision`` the reference
    accepts without an optional build ("double", "long double").  In the
    reference the default keyword "long double" selects *double* arithmetic and
    "double" selects long double (postbackend.pyx:219-222 vs
    variableprecisionposterior.cpp:35-41); both agree with each other to
    ~1.5e-6 in the reference's own tests, and with this backend to 1e-8.
  * ``validate=True`` (debug cross-check against the legacy v1 code,
    src/anomaly/posterior.cpp) is not part of the hot path and raises.
"""
import numpy as np

from ..batch import PosteriorBatch

_PRECISIONS = ("double", "long double", "float128", "dec50", "dec100")


def _has_float128():
    return False


def _has_dec50():
    return False


def _has_dec100():
    return False


class CppAnomalyPosterior:
    def __init__(self, qij, cij, wi, p, s, n, v, amin, rtol, validate=False,
                 pdf_algorithm="barycentric_lagrange", bli_max_splits=100,
                 bli_max_refinements=7, precision="long double", device=0, _api=None):
        pdf_algorithm = pdf_algorithm.lower()
        if pdf_algorithm not in ("explicit", "barycentric_lagrange", "adaptive_simpson"):
            raise ValueError("`pdf_algorithm` must be one of 'explicit', "
                             "'barycentric_lagrange', or 'adaptive_simpson'.")
        M = len(qij)
        if len(cij) != M:
            raise RuntimeError("Length of `cij` and `qij` do not match.")
        wi = np.asarray(wi, dtype=np.float64)
        if wi.shape[0] != M:
            raise RuntimeError("Length of `wi` and `qij` do not match.")
        if precision in ("float128", "dec50", "dec100"):
            raise RuntimeError("To use float128 backend, REHEATFUNQ needs to be "
                               "compiled with the anomaly_posterior_float128 option.")
        if precision not in _PRECISIONS:
            raise RuntimeError("Unknown precision '" + precision + "'.")
        if validate:
            raise RuntimeError("validate=True (cross-check against the legacy v1 backend) "
                               "is not available in the B200 backend.")
        if M == 0:
            raise RuntimeError("No samples given.")
        sets = []
        for i in range(M):
            q = np.ascontiguousarray(qij[i], dtype=np.float64)
            c = np.ascontiguousarray(cij[i], dtype=np.float64)
            if q.shape[0] != c.shape[0]:
                raise RuntimeError("Length of `cij` and `qij` do not match in "
                                   "element " + str(int(i)) + ".")
            sets.append((q, c))
        self._batch = PosteriorBatch([sets], [wi], (p, s, n, v), amin=amin, rtol=rtol,
                                     pdf_algorithm=pdf_algorithm,
                                     bli_max_refinements=int(bli_max_refinements),
                                     device=device, api=_api,
                                     # the reference's keywords are swapped (postbackend.pyx:
                                     # 219-222): "double" runs its long double build
                                     reference_arithmetic=("long double" if precision == "double"
                                                           else "double"))
        st = int(self._batch.status()[0])
        if st != 0:
            # the reference throws std::runtime_error / domain_error -> RuntimeError
            b = self._batch
            self._batch = None
            raise RuntimeError(b.status_message(st))
        self._M = M

    def _check(self):
        if self._batch is None:
            raise RuntimeError("Not properly initialized.")

    def Qmax(self) -> float:
        self._check()
        return float(self._batch.Qmax()[0])

    def _eval(self, name, x, what):
        self._check()
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim != 1:
            raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % x.ndim)
        out, qst = getattr(self._batch, name)(x, return_status=True)
        if qst[0] != 0:
            raise RuntimeError("An error occurred while computing the posterior " + what + ": "
                               + ("Encountered quantile out of bounds [0,1]." if qst[0] == 2
                                  else self._batch.status_message(int(qst[0]))))
        return out[0].copy()

    def pdf(self, PH, parallel=True):
        return self._eval("pdf", PH, "PDF")

    def cdf(self, PH, parallel=True):
        return self._eval("cdf", PH, "CDF")

    def tail(self, PH, parallel=True):
        return self._eval("tail", PH, "tail CDF")

    def tail_quantiles(self, t, parallel=True):
        return self._eval("tail_quantiles", t, "tail quantiles")

    # -- debug getters (postbackend.pyx:368-455) --------------------------
    def _get_locals(self, l):
        self._check()
        if l >= self._M:
            raise RuntimeError("Index 'l' out of bounds.")
        L = self._batch.locals(int(l))
        k = self._batch.k(int(l))
        return (L["lp"], L["ls"], L["n"], L["v"], L["amin"], L["Qmax"], k,
                (L["h0"], L["h1"], L["h2"], L["h3"]), L["w"], L["lh0"], L["l1p_w"],
                L["log_scale"], L["ymax"], L["norm"])

    def _get_C(self, a, l):
        """postbackend.pyx:422-431: (C0, C1, C2, C3) of the large-z expansion at `a`."""
        self._check()
        if l >= self._M:
            raise RuntimeError("Index 'l' out of bounds.")
        return self._batch.large_z_C(float(a), int(l))

    def _get_log_pdf_bli_data(self):
        self._check()
        return [self._batch.bli_samples(0, 0), self._batch.bli_samples(0, 1)]
