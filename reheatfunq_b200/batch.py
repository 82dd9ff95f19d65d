"""
Batched heat-flow anomaly posteriors on one B200.

``PosteriorBatch`` is the numpy-facing wrapper of the C ABI
(``include/reheatfunq_b200.h``): R independent posteriors, each a weighted list
of (q_i, c_i) sample sets, created in one call and evaluated with ragged
queries.  It replaces R constructions of the reference's
``Posterior<real,alg>`` (include/anomaly/posterior.hpp:88-108) plus the
OpenMP loops of its pdf/cdf/tail/tail_quantiles (:117-337).
"""
import ctypes as C
import numpy as np

from . import _lib

ALGORITHMS = {"explicit": 0, "barycentric_lagrange": 1, "adaptive_simpson": 2}

LOCALS_FIELDS = ("lp ls n v amin Qmax h0 h1 h2 h3 w lh0 l1p_w lv log_scale ymax ztrans S "
                 "taylor norm weight log_fac N imax status k_off flags "
                 "c1_0 c1_1 c2_0 c2_1 c2_2 c3_0 c3_1 c3_2 c3_3").split()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class PosteriorBatch:
    """
    Parameters
    ----------
    sets : list (length R) of lists of (q, c) array pairs, or, with
        ``post_off`` given, a flat list of (q, c) pairs.
    weights : list (length R) of weight arrays (one weight per sample set), or
        a flat array with ``post_off``.
    prior : (p, s, n, v) or array (R, 4)
    """

    def __init__(self, sets, weights, prior, amin=1.0, rtol=1e-8,
                 pdf_algorithm="barycentric_lagrange", bli_max_refinements=7,
                 device=0, post_off=None, api=None, reference_arithmetic="double"):
        self._api = api if api is not None else _lib.api()
        self._h = None
        if pdf_algorithm not in ALGORITHMS:
            raise ValueError("`pdf_algorithm` must be one of 'explicit', "
                             "'barycentric_lagrange', or 'adaptive_simpson'.")
        if post_off is None:
            flat = [qc for post in sets for qc in post]
            post_off = np.zeros(len(sets) + 1, dtype=np.int64)
            post_off[1:] = np.cumsum([len(p) for p in sets])
            w = np.concatenate([np.asarray(x, dtype=np.float64).ravel() for x in weights])
        else:
            flat = list(sets)
            post_off = np.ascontiguousarray(post_off, dtype=np.int64)
            w = np.asarray(weights, dtype=np.float64).ravel()
        R = len(post_off) - 1
        S = len(flat)
        if w.shape[0] != S:
            raise RuntimeError("Length of `wi` and `qij` do not match.")
        set_off = np.zeros(S + 1, dtype=np.int64)
        for i, (q, c) in enumerate(flat):
            if len(q) != len(c):
                raise RuntimeError("Length of `cij` and `qij` do not match in element "
                                   + str(i) + ".")
        set_off[1:] = np.cumsum([len(q) for q, _ in flat])
        if S:
            q = np.ascontiguousarray(np.concatenate([np.asarray(q, dtype=np.float64).ravel()
                                                     for q, _ in flat]))
            c = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.float64).ravel()
                                                     for _, c in flat]))
        else:
            q = np.zeros(0)
            c = np.zeros(0)
        self._init_packed(q, c, set_off, w, post_off, prior, amin, rtol, pdf_algorithm,
                          bli_max_refinements, device, reference_arithmetic)

    @classmethod
    def from_packed(cls, q, c, set_off, w, post_off, prior, amin=1.0, rtol=1e-8,
                    pdf_algorithm="barycentric_lagrange", bli_max_refinements=7, device=0,
                    api=None, reference_arithmetic="double"):
        """CSR-packed constructor (no Python-level list handling).

        ``reference_arithmetic``: which build of the reference is the model -- "double", or
        "long double" (its thresholds that are powers of the machine epsilon take their long
        double values; the arithmetic here is IEEE double either way)."""
        self = cls.__new__(cls)
        self._api = api if api is not None else _lib.api()
        self._h = None
        if pdf_algorithm not in ALGORITHMS:
            raise ValueError("`pdf_algorithm` must be one of 'explicit', "
                             "'barycentric_lagrange', or 'adaptive_simpson'.")
        self._init_packed(np.ascontiguousarray(q, dtype=np.float64),
                          np.ascontiguousarray(c, dtype=np.float64),
                          np.ascontiguousarray(set_off, dtype=np.int64),
                          np.ascontiguousarray(w, dtype=np.float64),
                          np.ascontiguousarray(post_off, dtype=np.int64),
                          prior, amin, rtol, pdf_algorithm, bli_max_refinements, device,
                          reference_arithmetic)
        return self

    @classmethod
    def from_device(cls, q_ptr, c_ptr, set_off_ptr, w_ptr, post_off_ptr, R, prior_ptr, n_sets,
                    n_data, prior_stride=0, amin=1.0, rtol=1e-8,
                    pdf_algorithm="barycentric_lagrange", bli_max_refinements=7, device=0,
                    api=None):
        """Inputs already resident in HBM: the pointer arguments are device addresses (ints),
        e.g. ``tensor.data_ptr()`` of float64 / int64 torch tensors on ``device``; ``n_sets`` and
        ``n_data`` are the lengths of ``w`` and of ``q`` / ``c`` (the offsets are checked against
        them)."""
        self = cls.__new__(cls)
        self._api = api if api is not None else _lib.api()
        self._h = None
        err = C.create_string_buffer(1024)
        h = C.c_void_p()
        rc = self._api.batch_create(C.byref(h), C.c_void_p(q_ptr), C.c_void_p(c_ptr),
                                    C.c_void_p(set_off_ptr), C.c_void_p(w_ptr),
                                    C.c_void_p(post_off_ptr), int(R), int(n_sets), int(n_data),
                                    C.c_void_p(prior_ptr),
                                    int(prior_stride), amin, rtol, ALGORITHMS[pdf_algorithm],
                                    int(bli_max_refinements), int(device), 1, err, len(err))
        if rc != 0:
            raise RuntimeError(err.value.decode())
        self._h = h
        self.R = int(R)
        self.S = None
        return self

    def query_device(self, what, x_ptr, x_off_ptr, out_ptr, n_points):
        """Evaluation with device-resident points / offsets / output (device addresses);
        ``n_points`` is the length of the point and output arrays."""
        fun = {"pdf": self._api.batch_pdf, "cdf": self._api.batch_cdf,
               "tail": self._api.batch_tail,
               "tail_quantiles": self._api.batch_tail_quantiles}[what]
        err = C.create_string_buffer(1024)
        rc = fun(self._h, C.c_void_p(x_ptr), C.c_void_p(x_off_ptr), int(n_points),
                 C.c_void_p(out_ptr), 1, None, err, len(err))
        if rc != 0:
            raise RuntimeError(err.value.decode())

    def _init_packed(self, q, c, set_off, w, post_off, prior, amin, rtol, pdf_algorithm,
                     bli_max_refinements, device, reference_arithmetic="double"):
        if reference_arithmetic not in ("double", "long double"):
            raise ValueError("reference_arithmetic must be 'double' or 'long double'")
        flags = 2 if reference_arithmetic == "long double" else 0
        R = len(post_off) - 1
        # the same consistency checks the C ABI makes, with Python-side messages
        if R < 1:
            raise RuntimeError("No posteriors given.")
        if len(w) != len(set_off) - 1 or post_off[0] != 0 or post_off[-1] != len(w) \
                or np.any(np.diff(post_off) < 0):
            raise ValueError("post_off must start at 0, be non-decreasing and end at len(w).")
        if len(q) != len(c) or set_off[0] != 0 or set_off[-1] != len(q) \
                or np.any(np.diff(set_off) < 0):
            raise ValueError("set_off must start at 0, be non-decreasing and end at len(q) == len(c).")
        prior = np.ascontiguousarray(prior, dtype=np.float64)
        if prior.ndim == 1:
            if prior.shape[0] != 4:
                raise ValueError("prior must be (p, s, n, v)")
            stride = 0
        else:
            if prior.shape != (R, 4):
                raise ValueError("prior must have shape (R, 4)")
            stride = 4
        err = C.create_string_buffer(1024)
        h = C.c_void_p()
        rc = self._api.batch_create(C.byref(h), _ptr(q), _ptr(c), _ptr(set_off), _ptr(w),
                                    _ptr(post_off), R, len(w), len(q), _ptr(prior), stride, amin,
                                    rtol,
                                    ALGORITHMS[pdf_algorithm], int(bli_max_refinements),
                                    int(device), flags, err, len(err))
        if rc != 0:
            raise RuntimeError(err.value.decode())
        self._h = h
        self.R = R
        self.S = len(set_off) - 1

    def __del__(self):
        if getattr(self, "_h", None):
            self._api.batch_destroy(self._h)
            self._h = None

    # -- properties -------------------------------------------------------
    def Qmax(self):
        out = np.empty(self.R)
        self._api.batch_qmax(self._h, out.ctypes.data_as(_lib._dp))
        return out

    def status(self):
        out = np.empty(self.R, dtype=np.int32)
        self._api.batch_status(self._h, out.ctypes.data_as(_lib._i32p))
        return out

    def status_message(self, code):
        return self._api.status_message(int(code)).decode()

    # -- evaluation -------------------------------------------------------
    def _query(self, fun, x, x_off=None, return_status=False):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x_off is None:
            # same points for every posterior: x is (R, n) or (n,)
            if x.ndim == 1:
                x = np.ascontiguousarray(np.broadcast_to(x, (self.R, x.shape[0])))
            if x.ndim != 2 or x.shape[0] != self.R:
                raise ValueError("x must have shape (n,) or (R, n) when x_off is not given.")
            n = x.shape[1]
            x_off = np.arange(self.R + 1, dtype=np.int64) * n
            shape = x.shape
        else:
            x_off = np.ascontiguousarray(x_off, dtype=np.int64)
            if x_off.shape != (self.R + 1,) or x_off[0] != 0 or x_off[-1] != x.size \
                    or np.any(np.diff(x_off) < 0):
                raise ValueError("x_off must have R + 1 entries, start at 0, be non-decreasing "
                                 "and end at x.size.")
            shape = x.shape
        flat = np.ascontiguousarray(x.ravel())
        out = np.empty_like(flat)
        qst = np.zeros(self.R, dtype=np.int32)
        err = C.create_string_buffer(1024)
        rc = fun(self._h, _ptr(flat), _ptr(x_off), flat.size, _ptr(out), 0, _ptr(qst), err,
                 len(err))
        if rc != 0:
            raise RuntimeError(err.value.decode())
        out = out.reshape(shape)
        if return_status:
            return out, qst
        return out

    def pdf(self, x, x_off=None, return_status=False):
        return self._query(self._api.batch_pdf, x, x_off, return_status)

    def cdf(self, x, x_off=None, return_status=False):
        return self._query(self._api.batch_cdf, x, x_off, return_status)

    def tail(self, x, x_off=None, return_status=False):
        return self._query(self._api.batch_tail, x, x_off, return_status)

    def tail_quantiles(self, t, t_off=None, return_status=False):
        return self._query(self._api.batch_tail_quantiles, t, t_off, return_status)

    # -- diagnostics ------------------------------------------------------
    def locals(self, s=0):
        out = np.zeros(36)
        rc = self._api.batch_get_locals(self._h, int(s), out.ctypes.data_as(_lib._dp))
        if rc != 0:
            raise RuntimeError("Index 'l' out of bounds.")
        return dict(zip(LOCALS_FIELDS, out))

    def k(self, s=0):
        """k_i = c_i Qmax / q_i of sample set s (Locals::ki, locals.hpp:75-87)."""
        n = int(self.locals(s)["N"])
        out = np.empty(max(n, 1))
        if self._api.batch_get_k(self._h, int(s), out.ctypes.data_as(_lib._dp), n) < 0:
            raise RuntimeError("Index 'l' out of bounds.")
        return out[:n].copy()

    def large_z_C(self, a, s=0):
        """(C0, C1, C2, C3)(a) of the large-z expansion (Posterior::get_C, posterior.hpp:427-440)."""
        L = self.locals(s)
        return np.array([1.0, L["c1_0"] + a * L["c1_1"],
                         L["c2_0"] + a * (L["c2_1"] + a * L["c2_2"]),
                         L["c3_0"] + a * (L["c3_1"] + a * (L["c3_2"] + a * L["c3_3"]))])

    def bli_samples(self, r=0, which=0):
        cap = 8 * 4096 + 1
        f = np.empty(cap)
        n = self._api.batch_get_bli(self._h, int(r), int(which), f.ctypes.data_as(_lib._dp), cap)
        if n < 0:
            raise RuntimeError("No interpolant available.")
        return f[:n].copy()

    def saq_leaves(self):
        return int(self._api.batch_saq_leaves(self._h))

    def counters(self):
        out = np.zeros(21, dtype=np.uint64)
        self._api.batch_counters(self._h, out.ctypes.data_as(_lib._u64p))
        names = ("nodes", "lz_nodes", "g_evals", "newton", "nsum", "launches",
                 "node_kernel_ns", "node_kernel_launches", "norm_kernel_ns",
                 "norm_kernel_launches", "cheb_terms", "saq_steps", "saq_kernel_ns",
                 "saq_kernel_launches", "fine_evals", "fine_bad", "plan_kernel_ns",
                 "plan_kernel_launches", "g_evals_quad", "z_unconverged", "quad_checks")
        return dict(zip(names, (int(v) for v in out)))
