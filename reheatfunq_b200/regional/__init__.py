"""Mirror of the hot entry points of ``reheatfunq.regional.backend``."""
from .backend import (gcp_ln_Phi, gamma_conjugate_prior_kullback_leibler,   # noqa: F401
                      gamma_conjugate_prior_kullback_leibler_batch_max,
                      gamma_conjugate_prior_kullback_leibler_batch_max_many, GCPMSECache,
                      gamma_conjugate_prior_predictive, gamma_conjugate_prior_predictive_batch,
                      gamma_conjugate_prior_predictive_common_norm,
                      gamma_conjugate_prior_predictive_cdf,
                      gamma_conjugate_prior_predictive_cdf_batch,
                      gamma_conjugate_prior_predictive_cdf_common_norm,
                      gamma_conjugate_prior_logL, gamma_conjugate_prior_bulk_log_p, gamma_mle,
                      gamma_conjugate_prior_minimum_surprise_backend)
