"""
Stand-in for ``reheatfunq.anomaly.anomaly`` (reheatfunq/anomaly/anomaly.pyx:71-178).
``AnomalyLS1980`` restates src/anomaly/ls1980.cpp:103-127 with a brute-force
nearest-segment search instead of the Boost.Geometry rtree.
"""
import numpy as np


class Anomaly:
    """Base class (anomaly.pyx:71-103): ``__call__(xy, P_H=1.0) -> c_i``."""

    def __call__(self, xy, P_H=1.0):
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        if xy.ndim != 2 or xy.shape[1] != 2:
            raise RuntimeError("`xy` must be of shape (N,2).")
        return self._c_i(xy, float(P_H))

    def _c_i(self, xy, P_H):
        raise RuntimeError("Anomaly not properly initialized. Did you use the base class?")


class AnomalyLS1980(Anomaly):
    """Lachenbruch & Sass (1980) eq. (A23b) along a segmented fault trace."""

    def __init__(self, xy, d):
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        if xy.ndim != 2 or xy.shape[1] != 2 or xy.shape[0] < 2:
            raise RuntimeError("`xy` must be of shape (N,2) with N >= 2.")
        self._a = xy[:-1]
        self._b = xy[1:]
        self._d = float(d)
        self._L = float(np.sqrt(((self._b - self._a) ** 2).sum(axis=1)).sum())

    def length(self):
        return self._L

    def _c_i(self, xy, P_H):
        ab = self._b - self._a                                  # (S,2)
        ap = xy[:, None, :] - self._a[None, :, :]               # (N,S,2)
        t = np.clip((ap * ab[None]).sum(axis=2) / (ab ** 2).sum(axis=1)[None], 0.0, 1.0)
        closest = self._a[None] + t[..., None] * ab[None]
        dist = np.sqrt(((xy[:, None, :] - closest) ** 2).sum(axis=2)).min(axis=1)
        d, L = self._d, self._L
        Qstar = 2.0 * P_H / (d * L)
        with np.errstate(divide="ignore", invalid="ignore"):
            c = Qstar / np.pi * (1.0 - dist / d * np.arctan(d / dist))
        c[dist == 0.0] = Qstar / np.pi
        return c


class AnomalyNearestNeighbor(Anomaly):
    """anomaly.pyx:152-178: c_i of the nearest of a set of (x, y, c) points."""

    def __init__(self, xyc):
        xyc = np.ascontiguousarray(xyc, dtype=np.float64)
        if xyc.ndim != 2 or xyc.shape[1] != 3:
            raise RuntimeError("`xyc` must be of shape (N,3).")
        self._xyc = xyc

    def _c_i(self, xy, P_H):
        d2 = ((xy[:, None, :] - self._xyc[None, :, :2]) ** 2).sum(axis=2)
        return P_H * self._xyc[d2.argmin(axis=1), 2]
