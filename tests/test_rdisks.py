"""
d_min-restricted sample enumeration: the hand-built cases of the
reference's testing/test_rdisks.py:24-92 with their expected (probability, sample)
lists, an independent brute-force enumeration of all selection orders, and the
invariants of the Monte-Carlo fallback.  Host-only code, exercised through the C ABI
of the test build of the library (same sources as the product).
"""
import itertools
from fractions import Fraction

import numpy as np
import pytest

from reheatfunq_b200.shims import rdisks


def _as_set(res):
    return sorted((round(p, 14), tuple(int(i) for i in s)) for p, s in res)


def _brute_force(xy, dmin):
    """Probability of each maximal conforming subset when points are selected uniformly at
    random, one after another, from those not yet excluded."""
    N = len(xy)
    d2 = ((xy[:, None, :] - xy[None, :, :]) ** 2).sum(axis=2)
    nb = [set(j for j in range(N) if j != i and d2[i, j] <= dmin * dmin) for i in range(N)]
    out = {}

    def rec(chosen, free, p):
        if not free:
            key = tuple(sorted(chosen))
            out[key] = out.get(key, Fraction(0)) + p
            return
        for i in sorted(free):
            rec(chosen + [i], free - {i} - nb[i], p / len(free))
    # isolated nodes are in every sample: enumerate over the conflicting ones only
    conf = set(i for i in range(N) if nb[i])
    iso = [i for i in range(N) if not nb[i]]
    if conf:
        rec([], conf, Fraction(1))
        out = {tuple(sorted(k + tuple(iso))): v for k, v in out.items()}
    else:
        out = {tuple(range(N)): Fraction(1)}
    return sorted((round(float(p), 14), s) for s, p in out.items())


CASES = [
    # X--X--X
    (np.array([(-2.0, 0.0), (0.0, 0.0), (2.0, 0.0)]), 1.1, [(1.0, (0, 1, 2))]),
    # X-X-X
    (np.array([(-1.0, 0.0), (0.0, 0.0), (1.0, 0.0)]), 1.1, [(1 / 3, (1,)), (2 / 3, (0, 2))]),
    #   X
    # X-X-X
    (np.array([(-1.0, 0.0), (0.0, 0.0), (1.0, 0.0), (0.0, 1.0)]), 1.1,
     [(1 / 4, (1,)), (3 / 4, (0, 2, 3))]),
    # X-XX-X
    (np.array([(-1.0, 0.0), (-0.05, 0.0), (0.05, 0.0), (1.0, 0.0)]), 1.2,
     [(1 / 4, (1,)), (1 / 4, (2,)), (1 / 2, (0, 3))]),
    # triangle, pair, single
    (np.array([(-0.5, 0.0), (0.5, 0.0), (0.0, np.sqrt(3) / 2), (2.0, 0.0), (3.0, 0.0),
               (5.0, 0.0)]), 1.2,
     [(1 / 6, (0, 3, 5)), (1 / 6, (0, 4, 5)), (1 / 6, (1, 3, 5)), (1 / 6, (1, 4, 5)),
      (1 / 6, (2, 3, 5)), (1 / 6, (2, 4, 5))]),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_reference_hand_built_cases(emu_api, case):
    xy, dmin, expected = CASES[case]
    res = rdisks.all_restricted_samples(xy, dmin, 100, 10000, np.random.default_rng(893289),
                                        True, _api=emu_api)
    assert _as_set(res) == sorted((round(p, 14), s) for p, s in expected)
    assert _as_set(res) == _brute_force(xy, dmin)
    # samples come back ordered by node tuple, indices ascending
    keys = [tuple(s) for _, s in res]
    assert all(list(k) == sorted(k) for k in keys)


def test_deterministic_matches_brute_force_random_sets(emu_api):
    rng = np.random.default_rng(77)
    for N in (5, 7, 9):
        xy = rng.random((N, 2))
        res = rdisks.all_restricted_samples(xy, 0.35, 10000, 1000000, None, _api=emu_api)
        assert _as_set(res) == _brute_force(xy, 0.35)
        assert abs(sum(p for p, _ in res) - 1.0) < 1e-14


def test_exceedance_without_generator_raises(emu_api):
    xy = np.random.default_rng(5).random((40, 2))
    with pytest.raises(RuntimeError, match="upper limit"):
        rdisks.all_restricted_samples(xy, 0.1, 100, 100, None, True, _api=emu_api)
    with pytest.raises(RuntimeError, match="Empty"):
        rdisks.all_restricted_samples(np.empty((0, 2)), 0.1, 100, 100, None, _api=emu_api)


@pytest.mark.parametrize("N", [40, 100])
def test_monte_carlo_fallback_invariants(emu_api, N):
    xy = np.random.default_rng(N).random((N, 2))
    dmin = 0.1
    res = rdisks.all_restricted_samples(xy, dmin, 10000, 20000, np.random.default_rng(1), True,
                                        _api=emu_api)
    again = rdisks.all_restricted_samples(xy, dmin, 10000, 20000, np.random.default_rng(1), True,
                                          _api=emu_api)
    assert len(res) == len(again) and all(
        a[0] == b[0] and np.array_equal(a[1], b[1]) for a, b in zip(res, again))
    assert abs(sum(p for p, _ in res) - 1.0) < 1e-12
    d2 = ((xy[:, None, :] - xy[None, :, :]) ** 2).sum(axis=2)
    np.fill_diagonal(d2, np.inf)
    seen = np.zeros(N, dtype=bool)
    for p, s in res:
        assert p > 0
        assert np.all(np.diff(s) > 0)
        sub = d2[np.ix_(s, s)]
        assert sub.min() > dmin * dmin                       # conforming
        rest = np.setdiff1d(np.arange(N), s)
        if rest.size:                                        # maximal
            assert np.all(d2[np.ix_(rest, s)].min(axis=1) <= dmin * dmin)
        seen[s] = True
    assert seen.all()           # the first pass starts one sample at every node


def test_bootstrap_data_selection(emu_api):
    xy = np.random.default_rng(3).random((30, 2))
    dmin = 0.15
    res = rdisks.bootstrap_data_selection(xy, dmin, 500, np.random.default_rng(9), _api=emu_api)
    assert abs(sum(w for w, _ in res) - 1.0) < 1e-12
    assert all(abs(w * 500 - round(w * 500)) < 1e-9 for w, _ in res)
    d2 = ((xy[:, None, :] - xy[None, :, :]) ** 2).sum(axis=2)
    np.fill_diagonal(d2, np.inf)
    for w, s in res:
        assert d2[np.ix_(s, s)].min() >= dmin * dmin
    assert len(set(tuple(s) for _, s in res)) == len(res)     # duplicates merged
    # frequencies follow the enumerated probabilities on a small case
    xy3, dm, expected = CASES[1]
    res3 = rdisks.bootstrap_data_selection(xy3, dm, 30000, 12, _api=emu_api)
    freq = {tuple(int(i) for i in s): w for w, s in res3}
    for p, s in expected:
        assert abs(freq[s] - p) < 0.01
    # one draw == conforming_data_selection with the same generator state
    m = rdisks.conforming_data_selection(xy, dmin, np.random.default_rng(9), _api=emu_api)
    first = rdisks.bootstrap_data_selection(xy, dmin, 1, np.random.default_rng(9), _api=emu_api)
    assert np.array_equal(np.flatnonzero(m), first[0][1])


def _global_phmax_expected(xy, ph, dmin):
    """max over maximal conforming samples of the sample's minimum, per connected component, then the
    minimum over components -- with the reference's pruning rule (dminpermutations.cpp:
    754-770): only components whose smallest value is STRICTLY below the smallest component
    maximum are evaluated (at least one)."""
    N = len(xy)
    d2 = ((xy[:, None, :] - xy[None, :, :]) ** 2).sum(axis=2)
    adj = (d2 <= dmin * dmin) & ~np.eye(N, dtype=bool)
    comp = -np.ones(N, dtype=int)
    for i in range(N):
        if comp[i] < 0:
            stack = [i]
            comp[i] = i
            while stack:
                j = stack.pop()
                for k in np.flatnonzero(adj[j]):
                    if comp[k] < 0:
                        comp[k] = i
                        stack.append(k)
    comps = [np.flatnonzero(comp == c) for c in np.unique(comp)]
    comps.sort(key=lambda c: ph[c].min())
    min_max = min(ph[c].max() for c in comps)
    keep = [c for c in comps if ph[c].min() < min_max] or comps[:1]
    out = np.inf
    for c in keep:
        best = -np.inf
        for r in range(1, len(c) + 1):
            for sub in itertools.combinations(c, r):
                sub = np.array(sub)
                rest = np.setdiff1d(c, sub)
                maximal = rest.size == 0 or adj[np.ix_(rest, sub)].any(axis=1).all()
                if maximal and not adj[np.ix_(sub, sub)].any():
                    best = max(best, ph[sub].min())
        out = min(out, best)
    return out


def test_global_phmax(emu_api):
    # no conflicts: the minimum over all points
    xy = np.array([(0.0, 0.0), (10.0, 0.0), (20.0, 0.0)])
    ph = np.array([3.0, 2.0, 5.0])
    assert rdisks.determine_global_PHmax(xy, ph, 1.0, _api=emu_api) == 2.0
    # X-X pair: either can be left out; the sample avoiding the small one has the larger
    # minimum -> max over samples of min over members
    xy = np.array([(0.0, 0.0), (0.5, 0.0), (20.0, 0.0)])
    assert rdisks.determine_global_PHmax(xy, np.array([1.0, 4.0, 5.0]), 1.0, _api=emu_api) == 4.0
    # reference pruning quirk: the isolated point with value 3.0 equals the
    # smallest component maximum and is pruned, so the pair's 4.0 is returned
    assert rdisks.determine_global_PHmax(xy, np.array([1.0, 4.0, 3.0]), 1.0, _api=emu_api) == 4.0
    rng = np.random.default_rng(21)
    for trial in range(20):
        xy = rng.random((12, 2))
        ph = rng.random(12) + 0.5
        assert rdisks.determine_global_PHmax(xy, ph, 0.3, _api=emu_api) \
            == _global_phmax_expected(xy, ph, 0.3)
    with pytest.raises(RuntimeError):
        rdisks.determine_global_PHmax(xy, ph[:5], 0.3, _api=emu_api)


def test_samples_from_discrete_distribution():
    w = np.array([0.1, 0.2, 0.3, 0.4])
    a = rdisks.samples_from_discrete_distribution(w, 20000, 5)
    b = rdisks.samples_from_discrete_distribution(w, 20000, np.random.default_rng(5))
    assert np.array_equal(a, b) and a.dtype == np.int64
    assert np.allclose(np.bincount(a, minlength=4) / 20000, w, atol=0.015)
    # the draw is the inverse cdf of Generator.random on the same stream
    z = np.random.default_rng(5).random(20000)
    assert np.array_equal(a, np.searchsorted(np.cumsum(w) / 1.0, z, side="left").clip(max=3))
