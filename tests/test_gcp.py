"""
Gamma conjugate prior integrals: oracle pinned to the reference's known answer
(testing/test_gcp.py:36-47) and to mpmath; the product's device math (host
emulation on CPU, CUDA on the GPU box) against the oracle.
"""
import numpy as np
import pytest

from oracle import oracle
from reheatfunq_b200.regional import backend as B

LP3 = np.log(3.0)
KL_REF = 87169608.06187229       # testing/test_gcp.py:44


def sample_params(seed=3, R=100):
    """per-sample updated flat priors (p = prod q, s = sum q, n = v = N)."""
    rng = np.random.default_rng(seed)
    params = []
    for r in range(R):
        N = rng.integers(10, 101)
        a = np.exp(rng.uniform(np.log(10), np.log(200)))
        q = rng.gamma(a, size=N) / (a / 68.3)
        params.append([np.log(q).sum(), q.sum(), N, N])
    return np.array(params)


REFS = np.array([[np.log(2.5), 15.4, 0.22, 0.218], [1.0, 10.0, 1.5, 1.0], [0.3, 30.0, 0.5, 0.3],
                 [2.0, 5.0, 3.0, 2.9], [LP3, 0.1, 0.1, 0.05], [0.5, 2.0, 1.0, 0.4]])


def test_oracle_known_answer():
    """GCP(3,0.1,0.1,0.05).kullback_leibler(GCP(3,0.1,1.0,0.05)); prior.py:281-302 makes the
    argument the first tuple and self the reference."""
    kl = oracle.gcp_kl(LP3, 0.1, 1.0, 0.05, LP3, 0.1, 0.1, 0.05, 1.0)
    assert abs(kl - KL_REF) < 1e-6 * KL_REF
    assert abs(kl - KL_REF) < 1e-8 * KL_REF        # in fact all printed digits agree
    assert oracle.gcp_kl_batch_max([[LP3, 0.1, 1.0, 0.05]], [[LP3, 0.1, 0.1, 0.05]])[0] == kl


def test_oracle_ln_phi_mpmath():
    import mpmath as mp
    mp.mp.dps = 30
    for (lp, s, n, v) in REFS[:4]:
        for amin in (0.5, 1.0, 4.0):
            F = lambda a: mp.exp((a - 1) * lp + mp.loggamma(v * a) - n * mp.loggamma(a)
                                 - v * a * mp.log(s))
            ref = mp.log(mp.quad(F, [amin, 5, 10, 50, 100, 1000, 1e4, 1e6, mp.inf]))
            assert abs(oracle.gcp_ln_phi(lp, s, n, v, amin) - float(ref)) < 1e-10 * max(1, abs(ref))


def check_against_oracle(api):
    kl = B.gamma_conjugate_prior_kullback_leibler(LP3, 0.1, 1.0, 0.05, LP3, 0.1, 0.1, 0.05, 1.0,
                                                  _api=api)
    assert abs(kl - KL_REF) < 1e-6 * KL_REF         # the reference's own tolerance
    assert kl == B.gamma_conjugate_prior_kullback_leibler_batch_max(
        np.array([[LP3, 0.1, 1.0, 0.05]]), LP3, 0.1, 0.1, 0.05, 1.0, _api=api)
    params = sample_params()
    for amin in (0.5, 1.0, 2.0, 4.0):
        for (lp, s, n, v) in REFS:
            a = B.gcp_ln_Phi(lp, s, n, v, amin, _api=api)
            b = oracle.gcp_ln_phi(lp, s, n, v, amin)
            assert abs(a - b) < 1e-9 * max(1.0, abs(b))
        o = oracle.gcp_kl_batch_max(params, REFS, amin)
        g, kl_all = B.gamma_conjugate_prior_kullback_leibler_batch_max_many(
            params, REFS, amin, return_all=True, _api=api)
        np.testing.assert_allclose(g, o, rtol=1e-8)
        assert np.all(kl_all.max(axis=1) == g)
        assert np.all(kl_all >= -1e-9)              # KL is non-negative


def check_predictive(api):
    """Posterior predictive pdf (gamma_conjugate_prior.cpp:794-990): SURVEY 8f item 1, pdf part."""
    q = np.array([-1.0, 0.0, 5.0, 20.0, 45.0, 68.3, 90.0, 150.0, 400.0])
    lp, s, n, v = np.log(2.522017292522833465), 15.37301672440559130, 0.2184765374862465137, \
        0.2184765374862465137
    got = B.gamma_conjugate_prior_predictive(q, lp, s, n, v, 1.0, _api=api)
    ref = oracle.gcp_predictive_pdf(q, lp, s, n, v, 1.0)
    assert np.all(got[:2] == 0)
    np.testing.assert_allclose(got[2:], ref[2:], rtol=1e-9)
    params = sample_params(R=6)
    qq = q[2:]
    got = B.gamma_conjugate_prior_predictive_common_norm(qq, params[:, 0], params[:, 1],
                                                         params[:, 2], params[:, 3], 1.0, _api=api)
    ref = oracle.gcp_predictive_pdf_common_norm(qq, params, 1.0)
    np.testing.assert_allclose(got, ref, rtol=1e-9)
    gb = B.gamma_conjugate_prior_predictive_batch(qq, params[:, 0], params[:, 1], params[:, 2],
                                                  params[:, 3], 1.0, _api=api)
    for i in range(params.shape[0]):
        np.testing.assert_allclose(gb[i], oracle.gcp_predictive_pdf(qq, *params[i], 1.0), rtol=1e-9)
    # predictive cdf (gamma_conjugate_prior.cpp:996-1157): single prior and common norm
    qc = np.array([70.0, 20.0, 45.0, 45.0, 110.0, 300.0])
    got = B.gamma_conjugate_prior_predictive_cdf(qc, lp, s, n, v, 1.0, _api=api)
    ref = oracle.gcp_predictive_cdf(qc, [[lp, s, n, v]], 1.0)
    # the reference integrates each piece with adaptive GK(7,15) at rel. tolerance 1e-9 and
    # at most 5 bisections: its own accuracy near q -> 0 (pdf ~ 1/|ln q|) is ~1e-8 absolute
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=5e-8)
    got = B.gamma_conjugate_prior_predictive_cdf_common_norm(qc, params[:, 0], params[:, 1],
                                                             params[:, 2], params[:, 3], 1.0,
                                                             _api=api)
    ref = oracle.gcp_predictive_cdf(qc, params, 1.0)
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=5e-8)
    assert np.all(np.diff(got[np.argsort(qc)]) >= 0) and got.max() <= 1.0
    # a pdf: integrates to one
    grid = np.linspace(1e-3, 600.0, 6001)
    p = B.gamma_conjugate_prior_predictive(grid, *params[0], 1.0, _api=api)
    assert abs(np.trapezoid(p, grid) - 1.0) < 1e-4


def test_device_math_emu(emu_api):
    check_against_oracle(emu_api)


def test_predictive_emu(emu_api):
    check_predictive(emu_api)


def test_cache_emu(emu_api):
    import pickle
    params = sample_params(R=10)
    c = B.GCPMSECache(params, 1.0, _api=emu_api)
    v1 = c(*REFS[1])
    assert c.misses() == 1 and c.hits() == 0
    assert c(*REFS[1]) == v1 and c.hits() == 1
    vals = c.evaluate_many(REFS[:3])
    assert vals[1] == v1 and c.size() == 3
    c2 = pickle.loads(pickle.dumps(c))
    assert c2 == c and c2.size() == 3
    c2._api = emu_api
    assert c2(*REFS[1]) == v1 and c2.hits() == 1
    with pytest.raises(RuntimeError):
        B.GCPMSECache(np.zeros((3, 5)), 1.0)


@pytest.mark.gpu
def test_device_math_gpu():
    from reheatfunq_b200 import _lib
    check_against_oracle(_lib.api())


@pytest.mark.gpu
def test_predictive_gpu():
    from reheatfunq_b200 import _lib
    check_predictive(_lib.api())


@pytest.mark.gpu
def test_cost_sweep_gpu():
    """Config 4 in miniature: many candidate references in one call."""
    from reheatfunq_b200 import _lib
    params = sample_params()
    rng = np.random.default_rng(0)
    K = 512
    refs = np.column_stack([rng.uniform(0.0, 4.0, K), rng.uniform(1.0, 50.0, K),
                            np.zeros(K), rng.uniform(0.05, 3.0, K)])
    refs[:, 2] = refs[:, 3] * (1.0 + np.exp(rng.uniform(-5, 1, K)))
    g = B.gamma_conjugate_prior_kullback_leibler_batch_max_many(params, refs, 1.0)
    idx = rng.choice(K, 24, replace=False)
    o = oracle.gcp_kl_batch_max(params, refs[idx], 1.0)
    fin = np.isfinite(o)
    np.testing.assert_allclose(g[idx][fin], o[fin], rtol=1e-8)


def check_cache_and_minimum_surprise(api):
    """GCPMSECache (backend.pyx:563-734: call, batch evaluation, hit counting, pickle round trip)
    and gamma_conjugate_prior_minimum_surprise_backend (backend.pyx:741-799: SHGO over the cached
    cost) on the given backend; the optimum's cost against the oracle's long-double cost."""
    import pickle
    params = sample_params(R=10)
    c = B.GCPMSECache(params, 1.0, _api=api)
    v1 = c(*REFS[1])
    assert np.isfinite(v1) and c.misses() == 1 and c.hits() == 0
    assert c(*REFS[1]) == v1 and c.hits() == 1
    vals = c.evaluate_many(REFS[:3])
    assert vals[1] == v1 and c.size() == 3
    o = oracle.gcp_kl_batch_max(params, REFS[:3], 1.0)
    fin = np.isfinite(o)
    np.testing.assert_allclose(np.asarray(vals)[fin], o[fin], rtol=1e-8)
    c2 = pickle.loads(pickle.dumps(c))
    assert c2 == c and c2.size() == 3
    c2._api = api
    assert c2(*REFS[1]) == v1 and c2.hits() == 1
    # SHGO front end (prior.py:484-536 bounds and variables)
    bounds = np.array([[0.0, 4.0], [1.0, 60.0], [-5.0, 1.0], [0.05, 3.0]])
    res = B.gamma_conjugate_prior_minimum_surprise_backend(params, 1.0, bounds,
                                                           dict(iters=1, n=32), False, _api=api)
    assert res.success and np.isfinite(res.fun)
    lp, s, v = res.x[0], res.x[1], res.x[3]
    n = v * (1.0 + np.exp(res.x[2]))                    # backend.pyx:783-788
    assert np.all(np.isfinite([lp, s, n, v])) and n > v > 0
    ref = oracle.gcp_kl_batch_max(params, [[lp, s, n, v]], 1.0)[0]
    got = B.gamma_conjugate_prior_kullback_leibler_batch_max(params, lp, s, n, v, 1.0, _api=api)
    assert abs(got - ref) < 1e-7 * max(1.0, abs(ref)) and abs(res.fun - ref) < 1e-7 * max(1.0, abs(ref))


def test_cache_and_minimum_surprise_emu(emu_api):
    check_cache_and_minimum_surprise(emu_api)


@pytest.mark.gpu
def test_cache_and_minimum_surprise_gpu():
    from reheatfunq_b200 import _lib
    check_cache_and_minimum_surprise(_lib.api())
