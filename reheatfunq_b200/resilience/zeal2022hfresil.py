"""
Drop-in for the hot functions of ``reheatfunq.resilience.zeal2022hfresil``
(reheatfunq/resilience/zeal2022hfresil.pyx:103-348) on the B200 backend.

``test_performance_cython`` / ``test_performance_mixture_cython`` keep the
reference's signatures and return the same ``(2, len(Nset), Nq, M)`` array.
Extras (keyword-only, ignored by reference callers): ``device``, ``nstreams``,
``chunk``; ``resilience_run(..., shard=(rank, world))`` computes one rank's RNG streams of an
experiment (``reheatfunq_b200.sharding.resilience_sharded`` gathers them).

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


def shard_index(M, nstreams, shard=(0, 1), api=None):
    """Global realisation indices (increasing) of shard (rank, world): rank owns the RNG
    streams t = rank, rank + world, ..., stream t the batches j = t, t + nstreams, ... of 10
    realisations (resilience.cpp:397,420-437)."""
    api = api if api is not None else _lib.api()
    n = api.resilience_shard_count(int(M), int(nstreams), int(shard[0]), int(shard[1]))
    if n < 0:
        raise ValueError("invalid shard")
    idx = np.empty(n, dtype=np.int64)
    if api.resilience_shard_index(int(M), int(nstreams), int(shard[0]), int(shard[1]),
                                  idx.ctypes.data_as(C.c_void_p)) != 0:
        raise ValueError("invalid shard")
    return idx


def resilience_run(dist_kind, dist_params, N, M, P_MW, quantile, prior, amin, tol, seed,
                   nstreams=1, device=0, chunk=0, shard=(0, 1), return_data=False,
                   out_device_ptr=None, return_samples=False, api=None):
    """One sample size.  Returns (res[2, Nq, M_local], failures[2], stats dict[, data dict]);
    M_local = M for the default shard (0, 1).  With ``out_device_ptr`` (device address of a
    float64 [2, Nq, M_local] array) the result stays in HBM and ``res`` is None."""
    api = api if api is not None else _lib.api()
    quantile = np.ascontiguousarray(quantile, dtype=np.float64)
    params = np.ascontiguousarray(dist_params, dtype=np.float64)
    prior = np.ascontiguousarray(prior, dtype=np.float64)
    Nq = quantile.shape[0]
    M_local = api.resilience_shard_count(int(M), int(nstreams), int(shard[0]), int(shard[1]))
    if M_local < 0:
        raise ValueError("invalid shard")
    out = None if out_device_ptr is not None else np.zeros((2, Nq, M_local))
    failures = np.zeros(2, dtype=np.int64)
    stats = np.zeros(9, dtype=np.uint64)
    dumps = [None] * 4
    if return_data:
        dumps = [np.zeros((M_local, N)) for _ in range(4)]
    samples = np.zeros((M_local, 4), dtype=np.int32) if return_samples else None
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    err = C.create_string_buffer(1024)
    rc = api.resilience_run(int(dist_kind), ptr(params), int(N), int(M), int(shard[0]),
                            int(shard[1]), float(P_MW), ptr(quantile), int(Nq), ptr(prior),
                            float(amin), float(tol), int(seed) & 0xFFFFFFFFFFFFFFFF,
                            int(nstreams), int(device), int(chunk),
                            C.c_void_p(out_device_ptr) if out_device_ptr is not None else ptr(out),
                            1 if out_device_ptr is not None else 0, ptr(failures), ptr(stats),
                            ptr(dumps[0]), ptr(dumps[1]), ptr(dumps[2]), ptr(dumps[3]),
                            ptr(samples), err, len(err))
    if rc != 0:
        raise RuntimeError(err.value.decode())
    names = ("nodes", "lz_nodes", "g_evals", "newton", "nsum", "launches", "node_kernel_ns",
             "host_draw_ns", "z_unconverged")
    st = dict(zip(names, (int(v) for v in stats)))
    if return_samples:
        st["bli_samples"] = samples
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
