"""
Minimal pure-numpy stand-ins for the reference's compiled modules that are NOT
on the hot path but are imported by the unmodified reference sources
(SURVEY 8b "import-time hazards").  They exist so that
``reheatfunq.anomaly.HeatFlowAnomalyPosterior`` and
``reheatfunq.regional.GammaConjugatePrior`` can be used unchanged on top of the
B200 backend (see ``reheatfunq_b200.overlay``); everything outside the simple
cases they cover raises ``NotImplementedError``.
"""
