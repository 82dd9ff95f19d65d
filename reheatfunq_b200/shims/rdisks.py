"""
Stand-in for ``reheatfunq.coverings.rdisks`` (reheatfunq/coverings/rdisks.pyx:165-414):
the d_min-restricted sample enumeration that feeds the bootstrap sample sets of
``HeatFlowAnomalyPosterior`` / ``HeatFlowPredictive`` (posterior.py:299-343).

The combinatorics live in the library (``csrc/rhfq_coverings.inl``, a restatement
of src/coverings/dminpermutations.cpp) behind ``rhfq_restricted_samples``,
``rhfq_bootstrap_data_selection`` and ``rhfq_global_phmax``; this module mirrors
the Cython glue: argument checks, the numpy BitGenerator hand-over, result lists.
Random draws go through the generator's ``bitgen_t`` like in the reference, so the
same ``rng`` gives the same samples.
"""
import ctypes as C

import numpy as np

from .._lib import api as _product_api


def _generator(rng):
    """rdisks.pyx:78-87"""
    if isinstance(rng, np.random.Generator):
        return rng.bit_generator
    if isinstance(rng, np.random.BitGenerator):
        return rng
    return np.random.default_rng(rng).bit_generator


def _bitgen_ptr(bg):
    return C.c_void_p(bg.ctypes.bit_generator.value)


def _xy(xy, empty_message="Empty coordinates."):
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    if xy.ndim != 2 or xy.shape[1] != 2:
        raise RuntimeError("Shape of xy must be (N,2).")
    if xy.shape[0] == 0:
        raise RuntimeError(empty_message)
    return xy


def _collect(api, handle):
    try:
        S = api.samples_count(handle)
        w = np.empty(S)
        off = np.empty(S + 1, dtype=np.int64)
        idx = np.empty(api.samples_index_count(handle), dtype=np.int64)
        api.samples_get(handle, w.ctypes.data, off.ctypes.data, idx.ctypes.data)
    finally:
        api.samples_destroy(handle)
    return w, [idx[off[i]:off[i + 1]].copy() for i in range(S)]


def all_restricted_samples(xy, dmin, max_samples, max_iter, rng=127, extra_debug_checks=False,
                           _api=None):
    """rdisks.pyx:324-353 -> [(probability, ascending index array), ...] ordered
    lexicographically by the packed conflict-node indices."""
    api = _api or _product_api()
    xy = _xy(xy)
    bg = _generator(rng) if rng is not None else None
    handle = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = api.restricted_samples(C.byref(handle), xy.ctypes.data, xy.shape[0], float(dmin),
                                int(max_samples), int(max_iter),
                                _bitgen_ptr(bg) if bg is not None else None, err, 512)
    if rc:
        raise RuntimeError(err.value.decode())
    w, samples = _collect(api, handle)
    return [(float(wi), s) for wi, s in zip(w, samples)]


def bootstrap_data_selection(xy, dmin_m, B, rng=127, _api=None):
    """rdisks.pyx:229-318 -> [[weight, index array], ...] in order of first occurrence."""
    api = _api or _product_api()
    xy = _xy(xy)
    bg = _generator(rng)
    handle = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = api.bootstrap_data_selection(C.byref(handle), xy.ctypes.data, xy.shape[0],
                                      float(dmin_m), int(B), _bitgen_ptr(bg), err, 512)
    if rc:
        raise RuntimeError(err.value.decode())
    w, samples = _collect(api, handle)
    return [[float(wi), s] for wi, s in zip(w, samples)]


def conforming_data_selection(xy, dmin_m, rng=128, _api=None):
    """rdisks.pyx:165-207: boolean mask of one random conforming subset (one draw of the
    bootstrap above, which consumes the generator identically)."""
    xy = np.asarray(xy, dtype=np.float64)
    if xy.ndim != 2 or xy.shape[1] != 2:
        raise RuntimeError("xy has to be shape (N,2).")
    mask = np.zeros(xy.shape[0], dtype=bool)
    if xy.shape[0]:
        mask[bootstrap_data_selection(xy, dmin_m, 1, rng, _api=_api)[0][1]] = True
    return mask


def determine_global_PHmax(xy, PHmax_i, dmin, _api=None):
    """rdisks.pyx:358-369"""
    api = _api or _product_api()
    xy = _xy(xy)
    PHmax_i = np.ascontiguousarray(PHmax_i, dtype=np.float64)
    if PHmax_i.ndim != 1 or PHmax_i.shape[0] != xy.shape[0]:
        raise RuntimeError("Shapes of xy and PHmax_i not compatible.")
    out = C.c_double()
    err = C.create_string_buffer(512)
    rc = api.global_phmax(xy.ctypes.data, PHmax_i.ctypes.data, xy.shape[0], float(dmin),
                          C.byref(out), err, 512)
    if rc:
        raise RuntimeError(err.value.decode())
    return out.value


def samples_from_discrete_distribution(w, N, rng=127):
    """rdisks.pyx:374-414: inverse-cdf sampling with one standard uniform per draw
    (``Generator.random`` is ``random_standard_uniform`` on the same bit stream)."""
    gen = np.random.Generator(_generator(rng))
    w = np.ascontiguousarray(w, dtype=np.float64)
    M = w.shape[0]
    cdf = np.empty(M)
    W = 0.0
    for i in range(M):
        W += w[i]
        cdf[i] = W
    if cdf[-1] < 1.0:
        cdf[-1] = 1.0
    elif cdf[-1] > 1.0:
        cdf /= W
    z = gen.random(int(N))
    j = np.searchsorted(cdf, z, side="left")
    if np.any(j >= M):
        raise RuntimeError("Computed index out of bounds.")
    return j.astype(np.int64)
