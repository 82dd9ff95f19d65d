"""
reheatfunq_b200 — B200-native (sm_100a) backend for the hot path of
REHEATFUNQ: batched FP64 evaluation of the heat-flow anomaly posterior.

Public surface:
  * :class:`reheatfunq_b200.batch.PosteriorBatch` — R posteriors in one call;
  * :mod:`reheatfunq_b200.anomaly.postbackend` — ``CppAnomalyPosterior`` with the
    signature of the reference's Cython class (reheatfunq/anomaly/postbackend.pyx:167);
  * the C ABI in ``include/reheatfunq_b200.h`` (``librhfq_b200.so``).
"""
from .batch import PosteriorBatch  # noqa: F401

__version__ = "0.1.0"
