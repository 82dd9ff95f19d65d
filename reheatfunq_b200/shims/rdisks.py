"""
Stand-in for ``reheatfunq.coverings.rdisks`` (reheatfunq/coverings/rdisks.pyx:165-414)
for the case without a minimum-distance criterion (``dmin == 0``), where every
data point is admissible and the restricted-sample enumeration
(src/coverings/dminpermutations.cpp) degenerates to the single full sample.
``dmin > 0`` is SURVEY 8f item 2 ("next").
"""
import numpy as np


def _check(xy, dmin):
    xy = np.asarray(xy)
    if xy.ndim != 2 or xy.shape[1] != 2:
        raise RuntimeError("Shape of xy must be (N,2).")
    if xy.shape[0] == 0:
        raise RuntimeError("Empty coordinates.")
    if dmin > 0.0:
        d = np.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(axis=2))
        np.fill_diagonal(d, np.inf)
        if d.min() < dmin:
            raise NotImplementedError(
                "reheatfunq_b200: d_min-restricted sample enumeration (dmin > 0 with "
                "conflicting points) is not part of the accelerated hot path yet.")
    return xy


def all_restricted_samples(xy, dmin, max_samples, max_iter, rng=127, extra_debug_checks=False):
    xy = _check(xy, dmin)
    return [(1.0, np.arange(xy.shape[0]))]


def bootstrap_data_selection(xy, dmin_m, B, rng=127):
    """rdisks.pyx:229-318 returns [[weight, index array], ...]; without conflicts every
    bootstrap draw is the full data set (weight 1)."""
    xy = _check(xy, dmin_m)
    return [[1.0, np.arange(xy.shape[0])]]


def conforming_data_selection(xy, dmin_m, rng=128):
    xy = _check(xy, dmin_m)
    return np.arange(xy.shape[0])


def determine_global_PHmax(xy, PHmax_i, dmin):
    _check(xy, dmin)
    PHmax_i = np.asarray(PHmax_i, dtype=np.float64)
    pos = PHmax_i[PHmax_i > 0]
    return float(pos.min()) if pos.size else float("inf")


def samples_from_discrete_distribution(w, N, rng=127):
    if isinstance(rng, int):
        rng = np.random.default_rng(rng)
    w = np.asarray(w, dtype=np.float64)
    return rng.choice(w.shape[0], size=int(N), p=w / w.sum())
