"""Mirror of ``reheatfunq.resilience`` (reheatfunq/resilience/__init__.py:20-24)."""
from .zeal2022hfresil import (test_performance_cython,            # noqa: F401
                              test_performance_mixture_cython,
                              resilience_run)
