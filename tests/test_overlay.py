"""
Drop-in check: the reference's unmodified ``HeatFlowAnomalyPosterior`` /
``GammaConjugatePrior`` front end (imported from /root/reference, which only
exists in the build container) on top of this repo's backend.  CPU: engine
host emulation behind the same mirror classes; GPU box: skipped (no reference
checkout there).
"""
import os
import warnings

import numpy as np
import pytest

from conftest import gamma_dataset, DEFAULT_PRIOR
from oracle.oracle import OraclePosterior

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "reheatfunq")),
                                reason="reference checkout not available")


@pytest.fixture()
def rq(emu_api):
    import reheatfunq_b200.overlay as ov
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = ov.install(REF, api=emu_api)
    yield mod
    import sys
    for name in list(sys.modules):
        if name == "reheatfunq" or name.startswith("reheatfunq."):
            del sys.modules[name]
    if REF in sys.path:
        sys.path.remove(REF)


def test_heat_flow_anomaly_posterior_unmodified_front_end(rq):
    """Recipe of testing/test_anomaly_posterior.py:43-116 (N = 25, one anomaly, dmin = 0)."""
    from reheatfunq import HeatFlowAnomalyPosterior, AnomalyLS1980
    from reheatfunq.regional import default_prior
    rng = np.random.default_rng(29189)
    N = 25
    Q = np.sort(rng.gamma(40.0, size=N) / 0.4)
    x = 160e3 * (rng.random(N) - 0.5)
    y = 160e3 * (rng.random(N) - 0.5)
    ano = AnomalyLS1980(np.array([(0, -80e3), (0, 80e3)], dtype=float), 14e3)
    q = Q + 1e3 * 100e6 * ano(np.stack((x, y), axis=1))
    gcp = default_prior()
    post = HeatFlowAnomalyPosterior(q, x, y, ano, gcp, 0.0)
    PH = np.linspace(0, post.PHmax, 50)
    # the same posterior straight from the oracle
    c = ano(np.stack((x, y), axis=1)) * 1e3
    O = OraclePosterior([q], [c], [1.0], np.exp(gcp.lp), gcp.s, gcp.n, gcp.v, gcp.amin, 1e-8,
                        precision="double")
    assert post.PHmax == O.Qmax()
    po = O.pdf(PH)
    m = po > 0
    np.testing.assert_allclose(post.pdf(PH)[m], po[m], rtol=1e-8)
    np.testing.assert_allclose(post.cdf(PH)[1:-1], O.cdf(PH)[1:-1], rtol=1e-8)
    np.testing.assert_allclose(post.tail(PH)[1:-1], O.tail(PH)[1:-1], rtol=1e-8)
    t = np.array([0.01, 0.5, 0.99])
    np.testing.assert_allclose(post.tail_quantiles(t), O.tail_quantiles(t), rtol=1e-6)


def test_gamma_conjugate_prior_kullback_leibler(rq):
    """testing/test_gcp.py:26-47 through the reference's own prior.py."""
    from reheatfunq import GammaConjugatePrior
    gcp0 = GammaConjugatePrior(3.0, 0.1, 0.1, 0.05)
    gcp1 = GammaConjugatePrior(3.0, 0.1, 1.0, 0.05)
    KL = gcp0.kullback_leibler(gcp1)
    assert abs(KL - 87169608.06187229) < 1e-6 * 87169608.06187229
    gcp2 = GammaConjugatePrior(3.0, 0.1, 1.0, 0.05, amin=2.0)
    with pytest.raises(RuntimeError):
        gcp0.kullback_leibler(gcp2)


def test_resilience_entry_points(rq):
    from reheatfunq.resilience import test_performance_mixture_cython
    res = test_performance_mixture_cython(np.array([10]), 3, 100.0, 31, 3.5, 0.33, 62, 5.5, 0.67,
                                          np.array([0.9, 0.5, 0.1, 0.01]), *DEFAULT_PRIOR, 1.0,
                                          verbose=False)
    assert res.shape == (2, 1, 4, 3) and np.all(np.isfinite(res))
