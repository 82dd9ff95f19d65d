"""
Import overlay: use the reference's UNMODIFIED Python front end
(``reheatfunq/anomaly/posterior.py``, ``reheatfunq/regional/prior.py``, ...) from a
reference checkout with the B200 backend behind it.

    import reheatfunq_b200.overlay as ov
    ov.install("/path/to/REHEATFUNQ")          # checkout root (contains reheatfunq/)
    from reheatfunq import HeatFlowAnomalyPosterior, GammaConjugatePrior, AnomalyLS1980

The reference's compiled modules are replaced in ``sys.modules`` before its
``__init__`` files run:

  reheatfunq.anomaly.postbackend        -> reheatfunq_b200.anomaly.postbackend   (hot path)
  reheatfunq.regional.backend           -> KL / ln Phi / cache from the B200 backend (hot
                                           path); the other names raise NotImplementedError
  reheatfunq.resilience.zeal2022hfresil -> reheatfunq_b200.resilience.zeal2022hfresil
  reheatfunq.anomaly.anomaly            -> numpy stand-in (off path)
  reheatfunq.coverings[.rdisks]         -> reheatfunq_b200.shims.rdisks (d_min sample enumeration
                                           in the library; also avoids the reference's
                                           pyproj / loaducerf3 imports)
"""
import os
import sys
import types


def _not_implemented(name):
    def f(*a, **k):
        raise NotImplementedError("reheatfunq_b200: `%s` is outside the accelerated hot path "
                                  "(SURVEY 8f)." % name)
    f.__name__ = name
    return f


def install(reference_root, api=None):
    """Registers the overlay.  ``api``: alternative C-ABI binding (tests pass the host
    emulation; the product default is librhfq_b200.so)."""
    pkg_dir = os.path.join(reference_root, "reheatfunq")
    if not os.path.isfile(os.path.join(pkg_dir, "anomaly", "posterior.py")):
        raise ImportError("No REHEATFUNQ checkout at " + reference_root)
    for name in list(sys.modules):
        if name == "reheatfunq" or name.startswith("reheatfunq."):
            del sys.modules[name]

    from .anomaly import postbackend
    from .regional import backend as kl
    from .resilience import zeal2022hfresil as resil
    from .shims import anomaly as shim_anomaly, rdisks as shim_rdisks

    if api is not None:
        import functools

        class _Bound(postbackend.CppAnomalyPosterior):
            def __init__(self, *a, **k):
                k.setdefault("_api", api)
                super().__init__(*a, **k)
        pb = types.ModuleType("reheatfunq.anomaly.postbackend")
        pb.CppAnomalyPosterior = _Bound
        pb._has_float128 = postbackend._has_float128
        pb._has_dec50 = postbackend._has_dec50
        pb._has_dec100 = postbackend._has_dec100
        bind = lambda f: functools.partial(f, _api=api)
    else:
        pb = postbackend
        bind = lambda f: f

    backend = types.ModuleType("reheatfunq.regional.backend")
    backend.gcp_ln_Phi = bind(kl.gcp_ln_Phi)
    backend.gamma_conjugate_prior_kullback_leibler = bind(kl.gamma_conjugate_prior_kullback_leibler)
    backend.gamma_conjugate_prior_kullback_leibler_batch_max = \
        bind(kl.gamma_conjugate_prior_kullback_leibler_batch_max)
    backend.GCPMSECache = kl.GCPMSECache if api is None else bind(kl.GCPMSECache)
    backend.gamma_conjugate_prior_predictive = bind(kl.gamma_conjugate_prior_predictive)
    backend.gamma_conjugate_prior_predictive_batch = bind(kl.gamma_conjugate_prior_predictive_batch)
    backend.gamma_conjugate_prior_predictive_common_norm = \
        bind(kl.gamma_conjugate_prior_predictive_common_norm)
    backend.gamma_conjugate_prior_logL = bind(kl.gamma_conjugate_prior_logL)
    backend.gamma_conjugate_prior_bulk_log_p = bind(kl.gamma_conjugate_prior_bulk_log_p)
    backend.gamma_conjugate_prior_minimum_surprise_backend = \
        bind(kl.gamma_conjugate_prior_minimum_surprise_backend)
    backend.gamma_mle = kl.gamma_mle
    backend.gamma_conjugate_prior_predictive_cdf = bind(kl.gamma_conjugate_prior_predictive_cdf)
    backend.gamma_conjugate_prior_predictive_cdf_batch = \
        bind(kl.gamma_conjugate_prior_predictive_cdf_batch)
    backend.gamma_conjugate_prior_predictive_cdf_common_norm = \
        bind(kl.gamma_conjugate_prior_predictive_cdf_common_norm)
    for name in ("gamma_conjugate_prior_mle",):
        setattr(backend, name, _not_implemented(name))

    resmod = types.ModuleType("reheatfunq.resilience.zeal2022hfresil")
    resmod.test_performance_cython = bind(resil.test_performance_cython)
    resmod.test_performance_mixture_cython = bind(resil.test_performance_mixture_cython)
    for name in ("generate_synthetic_heat_flow_coverings_mix2",
                 "generate_synthetic_heat_flow_coverings_mix3",
                 "generate_normal_mixture_errors_3"):
        setattr(resmod, name, _not_implemented(name))

    rdisks = types.ModuleType("reheatfunq.coverings.rdisks")
    for name in ("all_restricted_samples", "bootstrap_data_selection",
                 "conforming_data_selection", "determine_global_PHmax"):
        setattr(rdisks, name, bind(getattr(shim_rdisks, name)))
    rdisks.samples_from_discrete_distribution = shim_rdisks.samples_from_discrete_distribution
    coverings = types.ModuleType("reheatfunq.coverings")
    coverings.__path__ = []
    coverings.rdisks = rdisks
    coverings.conforming_data_selection = rdisks.conforming_data_selection
    coverings.bootstrap_data_selection = rdisks.bootstrap_data_selection
    coverings.random_global_R_disk_coverings = _not_implemented("random_global_R_disk_coverings")

    sys.modules["reheatfunq.anomaly.postbackend"] = pb
    sys.modules["reheatfunq.anomaly.anomaly"] = shim_anomaly
    sys.modules["reheatfunq.regional.backend"] = backend
    sys.modules["reheatfunq.resilience.zeal2022hfresil"] = resmod
    sys.modules["reheatfunq.coverings"] = coverings
    sys.modules["reheatfunq.coverings.rdisks"] = rdisks
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    import reheatfunq  # noqa: F401  (the reference's own, unmodified package files)
    return reheatfunq
