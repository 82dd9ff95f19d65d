"""
Drop-in check: the reference's unmodified ``HeatFlowAnomalyPosterior`` /
``GammaConjugatePrior`` front end on top of this repo's backend, twice: behind the engine's
host emulation (CPU) and behind librhfq_b200.so on the CUDA device (``-m gpu``).

The reference's Python files come from /root/reference in the build container.  The GPU box
has no such checkout: ``tools/stage_reference.sh`` copies the reference's pure-Python package
files into the git-ignored ``scratch/refstage/`` before a gpurun call (never committed); where
neither exists the tests are skipped.
"""
import os
import warnings

import numpy as np
import pytest

from conftest import gamma_dataset, DEFAULT_PRIOR
from oracle.oracle import OraclePosterior

from conftest import ROOT, has_gpu

REF = next((d for d in ("/root/reference", os.path.join(ROOT, "scratch", "refstage"))
            if os.path.isfile(os.path.join(d, "reheatfunq", "anomaly", "posterior.py"))), None)
pytestmark = pytest.mark.skipif(REF is None, reason="reference checkout not available")


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def rq(request):
    import reheatfunq_b200.overlay as ov
    if request.param == "emu":
        if has_gpu():
            pytest.skip("host emulation is the CPU-suite variant")
        api = request.getfixturevalue("emu_api")
    else:
        api = None          # the product: librhfq_b200.so on cuda:0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = ov.install(REF, api=api)
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


def test_minimum_surprise_estimate_front_end(rq):
    """GammaConjugatePrior.minimum_surprise_estimate (prior.py:405-539) through the B200-backed
    cost function; checks the optimum against a direct evaluation of the cost by the oracle."""
    from oracle import oracle
    from reheatfunq import GammaConjugatePrior
    rng = np.random.default_rng(7)
    hf = [rng.gamma(a, size=N) * 68.3 / a for a, N in ((20.0, 12), (60.0, 30), (35.0, 18))]
    gcp, res = GammaConjugatePrior.minimum_surprise_estimate(
        hf, shgo_kwargs=dict(iters=1, n=32), return_shgo_result=True)
    assert res.success and np.isfinite(res.fun)
    # the same cost from the oracle at the optimum
    params = np.array([[np.log(q).sum(), q.sum(), q.size, q.size] for q in hf])
    ref = oracle.gcp_kl_batch_max(params, [[gcp.lp, gcp.s, gcp.n, gcp.v]], 1.0)[0]
    assert abs(res.fun - ref) < 1e-7 * max(1.0, abs(ref))
    # log-probability helpers of the same front end
    lp = gcp.log_probability(np.array([30.0, 50.0]), np.array([0.4, 0.7]))
    assert np.all(np.isfinite(lp))


def test_heat_flow_predictive_front_end(rq):
    """reheatfunq.regional.HeatFlowPredictive (predictive.py:29-174) on the batched ln Phi kernel."""
    from oracle import oracle
    from reheatfunq import HeatFlowPredictive
    from reheatfunq.regional import default_prior
    rng = np.random.default_rng(3)
    q = rng.gamma(50.0, size=20) * 68.3 / 50.0
    x = 100e3 * rng.random(20)
    y = 100e3 * rng.random(20)
    gcp = default_prior()
    pred = HeatFlowPredictive(q, x, y, gcp, dmin=0.0, n_bootstrap=10)
    qq = np.array([40.0, 60.0, 68.3, 80.0, 100.0])
    u = gcp.updated(q)
    ref_pdf = oracle.gcp_predictive_pdf(qq, u.lp, u.s, u.n, u.v, u.amin)
    np.testing.assert_allclose(pred.pdf(qq), ref_pdf, rtol=1e-8)
    ref_cdf = oracle.gcp_predictive_cdf(qq, [[u.lp, u.s, u.n, u.v]], u.amin)
    np.testing.assert_allclose(pred.cdf(qq), ref_cdf, rtol=1e-6, atol=5e-8)


def test_heat_flow_anomaly_posterior_with_dmin(rq):
    """dmin > 0: the reference front end enumerates the d_min-restricted samples through
    rdisks (posterior.py:299-343), builds one (q, c) set per sample with the enumerated
    probabilities as weights, and the posterior is the weighted mixture."""
    from reheatfunq import HeatFlowAnomalyPosterior, AnomalyLS1980
    from reheatfunq.regional import default_prior
    rng = np.random.default_rng(4711)
    N = 14
    Q = np.sort(rng.gamma(40.0, size=N) / 0.4)
    x = 60e3 * (rng.random(N) - 0.5)
    y = 60e3 * (rng.random(N) - 0.5)
    ano = AnomalyLS1980(np.array([(0, -80e3), (0, 80e3)], dtype=float), 14e3)
    xy = np.stack((x, y), axis=1)
    q = Q + 1e3 * 100e6 * ano(xy)
    gcp = default_prior()
    dmin = 12e3
    post = HeatFlowAnomalyPosterior(q, x, y, ano, gcp, dmin, n_bootstrap=2000, rng=5)
    assert len(post.bootstrap) > 1
    w = np.array([b[0] for b in post.bootstrap])
    assert abs(w.sum() - 1.0) < 1e-12
    d = np.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(axis=2))
    np.fill_diagonal(d, np.inf)
    assert d.min() < dmin
    for _, _, ids in post.bootstrap:
        assert d[np.ix_(ids, ids)].min() > dmin
    c = ano(xy) * 1e3
    O = OraclePosterior([q[ids] for _, _, ids in post.bootstrap],
                        [c[ids] for _, _, ids in post.bootstrap], w,
                        np.exp(gcp.lp), gcp.s, gcp.n, gcp.v, gcp.amin, 1e-8, precision="double")
    assert post.PHmax == O.Qmax()
    PH = np.linspace(0, post.PHmax, 40)
    po = O.pdf(PH)
    m = po > 0
    np.testing.assert_allclose(post.pdf(PH)[m], po[m], rtol=1e-8)
    np.testing.assert_allclose(post.cdf(PH)[1:-1], O.cdf(PH)[1:-1], rtol=1e-8)
    t = np.array([0.01, 0.5, 0.99])
    np.testing.assert_allclose(post.tail_quantiles(t), O.tail_quantiles(t), rtol=1e-6)
    assert post.PHmax_global >= post.PHmax * (1 - 1e-15)


def test_heat_flow_predictive_with_dmin(rq):
    """HeatFlowPredictive with the default d_min: the conforming bootstrap
    (rdisks.bootstrap_data_selection) feeds one updated prior per distinct subset."""
    from oracle import oracle
    from reheatfunq import HeatFlowPredictive
    from reheatfunq.regional import default_prior
    rng = np.random.default_rng(8)
    q = rng.gamma(50.0, size=20) * 68.3 / 50.0
    x = 100e3 * rng.random(20)
    y = 100e3 * rng.random(20)
    gcp = default_prior()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pred = HeatFlowPredictive(q, x, y, gcp, dmin=20e3, n_bootstrap=50)
    assert len(pred.bootstrap) > 1
    assert abs(sum(b[0] for b in pred.bootstrap) - 1.0) < 1e-12
    xy = np.stack((x, y), axis=1)
    d = np.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(axis=2))
    np.fill_diagonal(d, np.inf)
    for _, ids in pred.bootstrap:
        assert d[np.ix_(ids, ids)].min() >= 20e3
    qq = np.array([40.0, 60.0, 68.3, 80.0, 100.0])
    ref = oracle.gcp_predictive_pdf_common_norm(
        qq, np.stack((pred.lp, pred.s, pred.n, pred.v), axis=1), gcp.amin)
    np.testing.assert_allclose(pred.pdf(qq), ref, rtol=1e-8)
