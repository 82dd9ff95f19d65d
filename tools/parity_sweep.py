"""Randomised parity sweep: posteriors with random N, gamma shapes, priors, amin, anomaly
strengths and (sometimes) several weighted sample sets, GPU (C ABI) against the CPU oracle.
Prints the worst relative deviations and every case beyond the north-star tolerances."""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from conftest import gamma_dataset, DEFAULT_PRIOR
from reheatfunq_b200 import PosteriorBatch
from oracle.oracle import OraclePosterior

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 20261020)
PRIORS = [DEFAULT_PRIOR, (1.0, 0.0, 0.0, 0.0), (3.0, 40.0, 2.5, 2.0), (10.0, 500.0, 8.0, 7.5)]
worst = dict(pdf=0.0, cdf=0.0, tail=0.0, q=0.0)
bad = []
t0 = time.time()
for case in range(n_cases):
    M = 1 if rng.random() < 0.7 else int(rng.integers(2, 6))
    N = int(np.exp(rng.uniform(np.log(3), np.log(300))))
    alpha = float(np.exp(rng.uniform(np.log(5), np.log(300))))
    P = float(10 ** rng.uniform(6.5, 8.5))
    prior = PRIORS[int(rng.integers(0, len(PRIORS)))]
    amin = float(rng.choice([1.0, 1.0, 0.5, 2.0]))
    rtol = float(rng.choice([1e-8, 1e-8, 1e-5]))
    sets, w = [], []
    for m in range(M):
        q, c = gamma_dataset(N, int(rng.integers(1, 10**6)), alpha, alpha / 68.3, P)
        sets.append((q, c)); w.append(float(rng.uniform(0.2, 1.0)))
    try:
        O = OraclePosterior([s[0] for s in sets], [s[1] for s in sets], w, *prior, amin, rtol,
                            precision="double")
    except RuntimeError as e:
        B = PosteriorBatch([sets], [w], prior, amin=amin, rtol=rtol)
        if B.status()[0] == 0:
            bad.append(dict(case=case, what="oracle failed, GPU did not", err=str(e)[:60]))
        continue
    B = PosteriorBatch([sets], [w], prior, amin=amin, rtol=rtol)
    if B.status()[0] != 0:
        bad.append(dict(case=case, what="GPU status %d" % B.status()[0], N=N, M=M))
        continue
    Q = B.Qmax()[0]
    x = np.concatenate([np.linspace(0, 1, 61)[1:-1], [0.999, 0.9999]]) * Q
    qs = np.array([0.01, 0.1, 0.5, 0.9, 0.99])
    res = {}
    for name, fb, fo in (("pdf", B.pdf, O.pdf), ("cdf", B.cdf, O.cdf), ("tail", B.tail, O.tail)):
        b, o = fb(x)[0], fo(x)
        m = o > 1e-280
        res[name] = float(np.max(np.abs(b[m] - o[m]) / o[m])) if m.any() else 0.0
    qb, qo = B.tail_quantiles(qs)[0], O.tail_quantiles(qs)
    res["q"] = float(np.max(np.abs(qb - qo) / qo))
    for k in worst:
        worst[k] = max(worst[k], res[k])
    if res["pdf"] > 1e-8 or res["cdf"] > 1e-8 or res["tail"] > 1e-8 or res["q"] > 1e-6:
        # where, and did the two sides keep the same interpolant levels?
        b, o = B.pdf(x)[0], O.pdf(x)
        m = o > 1e-280
        dev = np.where(m, np.abs(b - o) / np.where(m, o, 1.0), 0.0)
        lev = []
        for which in (0, 1):
            try:
                lev.append((len(B.bli_samples(0, which)), len(O.bli(which)[0])))
            except Exception as e:      # noqa
                lev.append(str(e)[:40])
        # the explicit pdf of both sides at the worst point
        xe = x[np.argsort(dev)[-3:]]
        Be = PosteriorBatch([sets], [w], prior, amin=amin, rtol=rtol, pdf_algorithm="explicit")
        Oe = OraclePosterior([s_[0] for s_ in sets], [s_[1] for s_ in sets], w, *prior, amin, rtol,
                             precision="double", pdf_algorithm="explicit")
        Ol = OraclePosterior([s_[0] for s_ in sets], [s_[1] for s_ in sets], w, *prior, amin, rtol,
                             precision="long double", pdf_algorithm="explicit")
        pe_b, pe_o, pe_l = Be.pdf(xe)[0], Oe.pdf(xe), Ol.pdf(xe)
        ymax = [(B.locals(l)["ymax"], O.locals(l)["ymax"]) for l in range(M)]
        bad.append(dict(case=case, N=N, M=M, alpha=alpha, P=P, prior=prior, amin=amin, rtol=rtol,
                        ymax_gpu_oracle=ymax, Qmax_l_over_Q=[B.locals(l)["Qmax"] / Q for l in range(M)],
                        x_over_Q=(xe / Q).tolist(), dev_bli=np.sort(dev)[-3:].tolist(),
                        levels_gpu_oracle=lev,
                        explicit_gpu_vs_oracle=(np.abs(pe_b - pe_o) / pe_o).tolist(),
                        explicit_gpu_vs_longdouble=(np.abs(pe_b - pe_l) / pe_l).tolist(),
                        explicit_oracle_vs_longdouble=(np.abs(pe_o - pe_l) / pe_l).tolist(), **res))
print(json.dumps(dict(cases=n_cases, seconds=time.time() - t0, worst=worst, beyond_tolerance=bad)))
