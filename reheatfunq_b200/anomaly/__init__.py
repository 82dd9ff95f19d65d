"""Mirror of ``reheatfunq.anomaly`` for the hot path (backend module only)."""
from .postbackend import CppAnomalyPosterior  # noqa: F401
