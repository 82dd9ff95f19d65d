"""
Drop-in for the hot functions of ``reheatfunq.resilience.zeal2022hfresil``
(reheatfunq/resilience/zeal2022hfresil.pyx:103-348) on the B200 backend.

``test_performance_cython`` / ``test_performance_mixture_cython`` keep the
reference's signatures and return the same ``(2, len(Nset), Nq, M)`` array.
Extras (keyword-only, ignored by reference callers): ``device``, ``nstreams``,
``chunk``, ``realisations=(m0, m1)`` for sharding one experiment over ranks.

Differences to the reference, all deliberate and listed in DESIGN.md:
  * any Nq >= 1 works (the reference silently returns zeros for Nq not in
    {1, 41} resp. raises for Nq not in {4, 41});
  * posteriors are computed in double (the reference uses long double);
  * an informed-prior failure raises RuntimeError("Internal numerics failure.")
    after the run instead of terminating the process (resilience.cpp:333-348).
"""
import ctypes as C

import numpy as np

from .. import _lib


def resilience_run(dist_kind, dist_params, N, M, P_MW, quantile, prior, amin, tol, seed,
                   nstreams=1, device=0, chunk=0, realisations=None, return_data=False,
                   api=None):
    """One sample size.  Returns (res[2, Nq, M], failures[2], stats dict[, data dict])."""
    api = api if api is not None else _lib.api()
    quantile = np.ascontiguousarray(quantile, dtype=np.float64)
    params = np.ascontiguousarray(dist_params, dtype=np.float64)
    prior = np.ascontiguousarray(prior, dtype=np.float64)
    Nq = quantile.shape[0]
    m0, m1 = (0, M) if realisations is None else realisations
    out = np.zeros((2, Nq, M))
    failures = np.zeros(2, dtype=np.int64)
    stats = np.zeros(8, dtype=np.uint64)
    dumps = [None] * 4
    if return_data:
        dumps = [np.zeros((M, N)) for _ in range(4)]
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    err = C.create_string_buffer(1024)
    rc = api.resilience_run(int(dist_kind), ptr(params), int(N), int(M), int(m0), int(m1),
                            float(P_MW), ptr(quantile), int(Nq), ptr(prior), float(amin),
                            float(tol), int(seed) & 0xFFFFFFFFFFFFFFFF, int(nstreams),
                            int(device), int(chunk), ptr(out), ptr(failures), ptr(stats),
                            ptr(dumps[0]), ptr(dumps[1]), ptr(dumps[2]), ptr(dumps[3]),
                            err, len(err))
    if rc != 0:
        raise RuntimeError(err.value.decode())
    names = ("nodes", "lz_nodes", "g_evals", "newton", "nsum", "launches", "node_kernel_ns",
             "host_draw_ns")
    st = dict(zip(names, (int(v) for v in stats)))
    if return_data:
        return out, failures, st, dict(q=dumps[0], c=dumps[1], u=dumps[2], q0=dumps[3])
    return out, failures, st


def _sweep(kind, params, Nset, M, P_MW, quantile, prior, amin, verbose, seed, tol, nthread,
           device, chunk, api):
    Nset = np.asarray(Nset, dtype=np.int64)
    quantile = np.asarray(quantile, dtype=np.float64)
    res = np.zeros((2, Nset.shape[0], quantile.shape[0], M))
    for i, N in enumerate(Nset):
        if verbose:
            print("i = " + str(i + 1) + " / " + str(Nset.shape[0]))
        out, failures, _ = resilience_run(kind, params, int(N), M, P_MW, quantile, prior, amin,
                                          tol, seed, nstreams=nthread, device=device,
                                          chunk=chunk, api=api)
        if failures[0] > 0:
            raise RuntimeError("Internal numerics failure.")
        res[:, i, :, :] = out
    return res


def test_performance_cython(Nset, M, P_MW, K, T, quantile, PRIOR_P, PRIOR_S, PRIOR_N, PRIOR_V,
                            amin=1.0, verbose=True, show_failures=False, seed=848782,
                            use_cpp_quantiles=True, tol=1e-3, nthread=0, *, device=0, chunk=0,
                            _api=None):
    """zeal2022hfresil.pyx:103-215 (gamma distributed heat flow)."""
    return _sweep(0, (K, T), Nset, M, P_MW, quantile, (PRIOR_P, PRIOR_S, PRIOR_N, PRIOR_V),
                  amin, verbose, seed, tol, nthread, device, chunk, _api)


def test_performance_mixture_cython(Nset, M, P_MW, x0, s0, a0, x1, s1, a1, quantile, PRIOR_P,
                                    PRIOR_S, PRIOR_N, PRIOR_V, amin, verbose=True,
                                    show_failures=False, seed=848782, use_cpp_quantiles=True,
                                    tol=1e-4, *, nthread=0, device=0, chunk=0, _api=None):
    """zeal2022hfresil.pyx:218-348 (two-normal mixture heat flow)."""
    return _sweep(1, (x0, s0, a0, x1, s1, a1), Nset, M, P_MW, quantile,
                  (PRIOR_P, PRIOR_S, PRIOR_N, PRIOR_V), amin, verbose, seed, tol, nthread,
                  device, chunk, _api)


# pytest must not collect the reference-named functions as tests
test_performance_cython.__test__ = False
test_performance_mixture_cython.__test__ = False
