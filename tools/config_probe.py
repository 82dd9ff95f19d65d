"""
Times BASELINE.json's five configurations on the GPU next to the CPU oracle
(bounded samples) and checks parity on the way.  Output: one JSON line per config.
"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
from reheatfunq_b200 import PosteriorBatch
from reheatfunq_b200.resilience import resilience_run
from reheatfunq_b200.regional import backend as gb
from oracle import oracle
from oracle.oracle import OraclePosterior

DP = bench.DEFAULT_PRIOR
threads = os.cpu_count() or 1


def ls1980(x, d=14e3, L=160e3):
    dist = np.abs(x)
    return 2 / (np.pi * d * L) * (1 - dist / d * np.arctan(d / dist))


def tmin(f, n=3):
    best = 1e99
    for _ in range(n):
        t0 = time.perf_counter(); r = f(); best = min(best, time.perf_counter() - t0)
    return best, r


out = []
# ---- C1: one N = 25 posterior, pdf/cdf/tail on 1000 nodes ----
rng = np.random.default_rng(29189)
N = 25
Q = np.sort(rng.gamma(40.0, size=N) / 0.4); x = 160e3 * (rng.random(N) - 0.5)
c = ls1980(x); q = Q + 1e3 * 100e6 * c; c = c * 1e3
def c1_gpu():
    B = PosteriorBatch([[(q, c)]], [[1.0]], DP)
    PH = np.linspace(0, B.Qmax()[0], 1000)
    return B.pdf(PH)[0], B.cdf(PH)[0], B.tail(PH)[0], PH
def c1_cpu():
    O = OraclePosterior([q], [c], [1.0], *DP, 1.0, 1e-8, precision="double")
    PH = np.linspace(0, O.Qmax(), 1000)
    return O.pdf(PH), O.cdf(PH), O.tail(PH), PH
c1_gpu()
tg, g = tmin(c1_gpu); tc, o = tmin(c1_cpu, 1)
m = o[0] > 0
out.append(dict(config="C1 single posterior N=25, pdf+cdf+tail on 1000 nodes", gpu_s=tg, cpu_s=tc,
                cpu_threads=threads, evals_per_s_gpu=3000 / tg, evals_per_s_cpu=3000 / tc,
                max_rel_pdf=float((abs(g[0][m] - o[0][m]) / o[0][m]).max()),
                max_rel_cdf=float((abs(g[1][1:-1] - o[1][1:-1]) / o[1][1:-1]).max())))

# ---- C2: N = 100, custom Gaussian anomaly, 99 quantiles; M = 1 and M = 1000 subsets ----
rng = np.random.default_rng(123329773)
N = 100
xy = 3 * rng.random((N, 2)) - 1.5
q0 = 0.39 * rng.gamma(175.2, size=N)
ci = 68.3e-3 / 10e6 * np.exp(-(xy ** 2).sum(axis=1))
q = q0 + 1e3 * 10e6 * ci
c = ci * 1e3
t = np.linspace(0.01, 0.99, 99)
def c2_gpu(sets, w):
    B = PosteriorBatch([sets], [w], DP)
    return B.tail_quantiles(t)[0]
def c2_cpu(sets, w):
    O = OraclePosterior([s[0] for s in sets], [s[1] for s in sets], w, *DP, 1.0, 1e-8, precision="double")
    return O.tail_quantiles(t)
sets1, w1 = [(q, c)], [1.0]
c2_gpu(sets1, w1)
tg, g = tmin(lambda: c2_gpu(sets1, w1)); tc, o = tmin(lambda: c2_cpu(sets1, w1), 1)
out.append(dict(config="C2 tail quantiles (99) N=100, M=1", gpu_s=tg, cpu_s=tc, cpu_threads=threads,
                max_rel_quantile=float((abs(g - o) / o).max())))
M = 1000
setsM = []
for m_ in range(M):
    idx = np.sort(rng.choice(N, size=N - 5, replace=False))
    setsM.append((q[idx], c[idx]))
wM = np.ones(M) / M
c2_gpu(setsM[:8], wM[:8])
tg, g = tmin(lambda: c2_gpu(setsM, wM), 2)
Mc = 16
tc, o = tmin(lambda: c2_cpu(setsM[:Mc], wM[:Mc]), 1)
g16 = c2_gpu(setsM[:Mc], wM[:Mc])
out.append(dict(config="C2 tail quantiles (99) N=95, M=1000 bootstrap sets", gpu_s=tg,
                cpu_s_M16=tc, cpu_s_extrapolated_M1000=tc * M / Mc, cpu_threads=threads,
                max_rel_quantile_M16=float((abs(g16 - o) / o).max())))

# ---- C4: KL cost over 100 sample tuples, 4096 candidates, amin sweep ----
rng = np.random.default_rng(3)
params = []
for r in range(100):
    Nn = rng.integers(10, 101); a = np.exp(rng.uniform(np.log(10), np.log(200)))
    qq = rng.gamma(a, size=Nn) / (a / 68.3)
    params.append([np.log(qq).sum(), qq.sum(), Nn, Nn])
params = np.array(params)
K = 4096
refs = np.column_stack([rng.uniform(0.0, 4.0, K), rng.uniform(1.0, 50.0, K), np.zeros(K),
                        rng.uniform(0.05, 3.0, K)])
refs[:, 2] = refs[:, 3] * (1.0 + np.exp(rng.uniform(-5, 1, K)))
def c4_gpu():
    return [gb.gamma_conjugate_prior_kullback_leibler_batch_max_many(params, refs, am) for am in (0.5, 1, 2, 4)]
c4_gpu()
tg, g = tmin(c4_gpu)
Kc = 64
tc, o = tmin(lambda: [oracle.gcp_kl_batch_max(params, refs[:Kc], am) for am in (0.5, 1, 2, 4)], 1)
rel = max(float(np.nanmax(np.abs(np.array(g[i][:Kc]) / o[i] - 1)[np.isfinite(o[i])])) for i in range(4))
out.append(dict(config="C4 KL cost: 100 tuples x 4096 candidates x 4 amin", gpu_s=tg,
                cost_evals_per_s_gpu=4 * K / tg, cpu_s_K64=tc, cost_evals_per_s_cpu=4 * Kc / tc,
                cpu_threads=threads, max_rel=rel))

# ---- C5: resilience (bounded) ----
MIX = (31, 3.5, 0.33, 62, 5.5, 0.67)
Q41 = np.linspace(0.99, 0.01, 41)
for Nn in (10, 50, 100):
    # warm-up with the full size (the stream-ordered memory pool grows to the footprint of this
    # shape: first-touch allocation and pool re-mapping are not what is measured), best of 2
    resilience_run(1, MIX, Nn, 100000, 100.0, Q41, DP, 1.0, 1e-4, 848782, nstreams=8)
    tg, (rg, fg, st) = tmin(lambda: resilience_run(1, MIX, Nn, 100000, 100.0, Q41, DP, 1.0, 1e-4, 848782, nstreams=8), 2)
    Mc = 4 * threads
    tc, (ro, fo) = tmin(lambda: oracle.resilience(1, MIX, Nn, Mc, 100.0, Q41, DP, 1.0, 1e-4, 848782, nthread=threads), 1)
    out.append(dict(config="C5 resilience mixture N=%d, 41 quantiles, 2 priors" % Nn, gpu_s_M100000=tg,
                    sets_per_s_gpu=100000 / tg, cpu_s=tc, cpu_sample=Mc, sets_per_s_cpu=Mc / tc,
                    cpu_threads=threads, failures=[int(fg[0]), int(fg[1])]))
for o_ in out:
    print(json.dumps(o_))
