"""
Resilience Monte Carlo (src/resilience/resilience.cpp): synthetic-data streams
bit-exact against the oracle's use of libstdc++ <random>, and quantiles of the
batched pipeline against the oracle's per-realisation long-double posteriors.
CPU part: host emulation of the engine (logic); GPU part: the CUDA path.
"""
import numpy as np
import pytest

from conftest import DEFAULT_PRIOR
from oracle import oracle
from reheatfunq_b200.resilience import resilience_run

MIX = (31, 3.5, 0.33, 62, 5.5, 0.67)      # A4-Resilience notebook, cell 20
GAM = (7.0, 10.0)                          # A3-Posterior-Impact notebook, cell 17
QUANT41 = np.linspace(0.99, 0.01, 41)


def check(api, kind, params, tol, N, M, T, exact_positions):
    ro, fo, do = oracle.resilience(kind, params, N, M, 100.0, QUANT41, DEFAULT_PRIOR, 1.0, tol,
                                   848782, nthread=T, precision="long double", return_data=True)
    rb, fb, st, db = resilience_run(kind, params, N, M, 100.0, QUANT41, DEFAULT_PRIOR, 1.0, tol,
                                    848782, nstreams=T, return_data=True, api=api)
    # bit-exact sample indexing: the raw draws of every realisation
    assert np.array_equal(do["u"], db["u"])
    assert np.array_equal(do["q0"], db["q0"])
    if exact_positions:
        assert np.array_equal(do["q"], db["q"]) and np.array_equal(do["c"], db["c"])
    else:
        # device asin/atan may differ from glibc in the last bit
        np.testing.assert_allclose(db["q"], do["q"], rtol=1e-13)
        np.testing.assert_allclose(db["c"], do["c"], rtol=1e-11)
    assert np.array_equal(fo, fb)
    ok = np.isfinite(ro) & np.isfinite(rb)
    assert np.array_equal(np.isfinite(ro), np.isfinite(rb))
    rel = np.abs(rb[ok] - ro[ok]) / np.abs(ro[ok])
    # north_star: 1e-6 on tail quantiles.  A flipped BLI refinement decision
    # (double vs long double at rtol = tol) would show up as a much larger error.
    assert rel.max() < 1e-6, "max rel %.2e, fraction > 1e-6: %.4f" % (rel.max(), (rel > 1e-6).mean())
    assert st["nodes"] > 0
    return rb


@pytest.mark.parametrize("kind,params,tol", [(1, MIX, 1e-4), (0, GAM, 1e-3)])
@pytest.mark.parametrize("N,M,T", [(10, 12, 1), (24, 23, 3)])
def test_resilience_emu(emu_api, kind, params, tol, N, M, T):
    check(emu_api, kind, params, tol, N, M, T, exact_positions=True)


def test_resilience_shards_emu(emu_api):
    """Sharding one experiment over ranks by RNG stream: the shards' columns, placed at their
    global indices, reproduce the single-rank run bit for bit -- for any number of ranks, with
    uneven stream counts, an incomplete last batch, and more ranks than streams."""
    from reheatfunq_b200.resilience.zeal2022hfresil import shard_index
    M, T = 47, 3
    args = (1, MIX, 12, M, 100.0, QUANT41[::10], DEFAULT_PRIOR, 1.0, 1e-4, 7)
    full, ff, _, dfull = resilience_run(*args, nstreams=T, return_data=True, api=emu_api)
    assert np.array_equal(shard_index(M, T, (0, 1), api=emu_api), np.arange(M))
    for world in (2, 3, 4):
        seen = np.zeros(M, dtype=int)
        fails = np.zeros(2, dtype=np.int64)
        for rank in range(world):
            idx = shard_index(M, T, (rank, world), api=emu_api)
            assert np.all(np.diff(idx) > 0)
            part, f, _, d = resilience_run(*args, nstreams=T, shard=(rank, world), chunk=10,
                                           return_data=True, api=emu_api)
            assert part.shape == (2, 5, len(idx))
            np.testing.assert_array_equal(part, full[:, :, idx])
            np.testing.assert_array_equal(d["u"], dfull["u"][idx])
            np.testing.assert_array_equal(d["q0"], dfull["q0"][idx])
            # batches of 10 realisations stay with their stream: stream = (m // 10) % T
            assert np.all(((idx // 10) % T) % world == rank)
            seen[idx] += 1
            fails += f
        assert np.all(seen == 1)
        assert np.array_equal(fails, ff)


def test_reference_signature_emu(emu_api):
    from reheatfunq_b200.resilience import test_performance_mixture_cython, test_performance_cython
    res = test_performance_mixture_cython(np.array([10, 12]), 4, 100.0, *MIX, QUANT41[::10][:4],
                                          *DEFAULT_PRIOR, 1.0, verbose=False, nthread=1,
                                          _api=emu_api)
    assert res.shape == (2, 2, 4, 4) and np.all(np.isfinite(res))
    res = test_performance_cython(np.array([10]), 3, 100.0, 7.0, 10.0, np.array([0.5]),
                                  *DEFAULT_PRIOR, verbose=False, nthread=1, _api=emu_api)
    assert res.shape == (2, 1, 1, 3) and np.all(res > 0)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,params,tol", [(1, MIX, 1e-4), (0, GAM, 1e-3)])
def test_resilience_gpu(kind, params, tol):
    from reheatfunq_b200 import _lib
    rb = check(_lib.api(), kind, params, tol, 24, 40, 2, exact_positions=False)
    assert np.all(np.isfinite(rb))


@pytest.mark.gpu
def test_resilience_gpu_chunks_and_properties():
    """Full-size style run: results do not depend on the chunking; quantiles are ordered."""
    from reheatfunq_b200 import _lib
    a, fa, _ = resilience_run(1, MIX, 50, 3000, 100.0, QUANT41, DEFAULT_PRIOR, 1.0, 1e-4, 848782,
                              nstreams=4, chunk=1024, api=_lib.api())
    b, fb, _ = resilience_run(1, MIX, 50, 3000, 100.0, QUANT41, DEFAULT_PRIOR, 1.0, 1e-4, 848782,
                              nstreams=4, chunk=0, api=_lib.api())
    # the segmented prefix sums of the Simpson masses round differently for
    # different batch layouts: equal to ~1e-13, not bitwise
    np.testing.assert_allclose(a, b, rtol=1e-11)
    ok = np.isfinite(a).all(axis=1)
    assert ok.mean() > 0.99
    # tail quantiles t = 0.99 ... 0.01 are increasing in P_H
    d = np.diff(a, axis=1)
    assert np.all(d[:, :, :][np.broadcast_to(ok[:, None, :], d.shape)] >= 0)


def _assert_resilience_sweep(rep, min_fraction):
    """Product against the LONG-DOUBLE oracle (the arithmetic of resilience.cpp:321) on the
    realisations of tests/golden/resilience_ld_golden.npz (tools/resilience_sweep.py): both
    sides fail on the same realisations, every quantile is within north_star's 1e-6 -- a
    flipped refinement decision of an interpolant (double vs x87 long double at rtol = tol,
    SURVEY section 7) may move a quantile, but not beyond the tolerance."""
    for cfg in rep["configs"]:
        assert cfg["finite_mismatch"] == 0, cfg
        assert cfg["failures_product"] == cfg["failures_oracle"], cfg
    assert rep["fraction_within_1e6"] >= min_fraction, rep["fraction_within_1e6"]
    assert rep["max_rel"] <= 1e-6, (rep["max_rel"], rep["flipped_refinements"][:5])


def test_resilience_sweep_emu():
    import sys
    from conftest import ROOT
    sys.path.insert(0, __import__("os").path.join(ROOT, "tools"))
    import resilience_sweep
    rep = resilience_sweep.sweep(emu=True, limit=30)
    assert rep["realisations"] == 240
    _assert_resilience_sweep(rep, 1.0)


@pytest.mark.gpu
def test_resilience_sweep_gpu():
    """10^4 realisations (N in {10, 25, 50, 100} x {mixture tol 1e-4, gamma tol 1e-3}) on the
    GPU against the committed long-double oracle results; the report goes to gpurun_out/."""
    import json
    import os
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import resilience_sweep
    rep = resilience_sweep.sweep(emu=False)
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "resilience_sweep_10000.json"), "w") as f:
            json.dump(rep, f, indent=1)
    except OSError:
        pass
    assert rep["realisations"] == 10000
    _assert_resilience_sweep(rep, 1.0)
