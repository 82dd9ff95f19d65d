"""Resilience parity at scale: the product (GPU through the C ABI; `--emu`: the TEST-ONLY host
emulation) against the LONG-DOUBLE oracle -- the arithmetic the reference's resilience loop
uses (src/resilience/resilience.cpp:321) -- over 10^4 realisations: N in {10, 25, 50, 100},
both heat-flow distributions, tol 1e-3 (gamma) / 1e-4 (mixture) as zeal2022hfresil.pyx:111,227.

    python tools/resilience_sweep.py --make-golden [--jobs J]   # oracle -> tests/golden/ (CPU, minutes)
    python tools/resilience_sweep.py [--emu] [--limit M]        # product vs the golden file

Reports, per configuration and overall: the fraction of quantiles within 1e-6 (north_star), the
worst deviation, failures on either side, and every realisation whose barycentric interpolants
kept a different number of samples than the oracle's (a flipped refinement decision,
barylagrint.hpp:576-668: double vs x87 long double at rtol = tol).
The golden file holds only the oracle's outputs; the synthetic data are regenerated from the
seeds on both sides (bit-exact RNG restatement, tests/test_resilience.py).
"""
import os, sys, json, time
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..")
sys.path.insert(0, ROOT)
import numpy as np

MIX = (31, 3.5, 0.33, 62, 5.5, 0.67)      # A4-Resilience notebook, cell 20
GAM = (7.0, 10.0)                          # A3-Posterior-Impact notebook, cell 17
PRIOR = (2.522017292522833465, 15.37301672440559130, 0.2184765374862465137, 0.2184765374862465137)
QUANT = np.array([0.99, 0.9, 0.5, 0.1, 0.01])
SEED = 848782
NSTREAMS = 5          # odd on purpose: batches of 10 dealt over 5 streams
M_PER_CONFIG = 1250   # 8 configurations -> 10^4 realisations
CONFIGS = [(kind, params, tol, N) for kind, params, tol in ((1, MIX, 1e-4), (0, GAM, 1e-3))
           for N in (10, 25, 50, 100)]
GOLDEN = os.path.join(ROOT, "tests", "golden", "resilience_ld_golden.npz")


def make_golden(jobs):
    from oracle import oracle
    oracle.lib().orc_set_num_threads(jobs)
    out = {}
    for ci, (kind, params, tol, N) in enumerate(CONFIGS):
        t0 = time.time()
        # the oracle's OpenMP threads are its RNG streams: nthread = NSTREAMS
        res, fail, samples = oracle.resilience_levels(kind, params, N, M_PER_CONFIG, 100.0, QUANT,
                                                      PRIOR, 1.0, tol, SEED, nthread=NSTREAMS,
                                                      precision="long double")
        out["res_%d" % ci] = res
        out["fail_%d" % ci] = fail
        out["samples_%d" % ci] = samples
        print("config %d kind=%d N=%d tol=%g: %.1f s, failures %s" % (ci, kind, N, tol, time.time() - t0, fail),
              file=sys.stderr)
    np.savez_compressed(GOLDEN, quantiles=QUANT, nstreams=NSTREAMS, seed=SEED, m=M_PER_CONFIG, **out)


def sweep(emu=False, limit=None):
    from reheatfunq_b200.resilience import resilience_run
    api = None
    if emu:
        import ctypes as C
        from reheatfunq_b200._lib import Api
        api = Api(C.CDLL(os.path.join(ROOT, "tests", "hostmath", "libengine_emu.so")), "emu_rhfq_")
    G = np.load(GOLDEN)
    M = int(G["m"]) if limit is None else min(int(G["m"]), int(limit))
    report = dict(emu=bool(emu), realisations_per_config=M, configs=[], tolerance=1e-6)
    tot_q = tot_in = 0
    worst = 0.0
    flips_all = []
    t0 = time.time()
    for ci, (kind, params, tol, N) in enumerate(CONFIGS):
        # the first M realisations of an experiment do not depend on its length (stream t owns
        # batches t, t + T, ...), so a shorter product run is compared with a prefix of the golden
        rb, fb, st = resilience_run(kind, params, N, M, 100.0, QUANT, PRIOR, 1.0, tol, int(G["seed"]),
                                    nstreams=NSTREAMS, return_samples=True, api=api)
        ro = G["res_%d" % ci][:, :, :M]
        so = G["samples_%d" % ci][:M]
        sb = st["bli_samples"]
        fin_b, fin_o = np.isfinite(rb), np.isfinite(ro)
        ok = fin_b & fin_o
        rel = np.zeros_like(rb)
        rel[ok] = np.abs(rb[ok] - ro[ok]) / np.abs(ro[ok])
        flipped = np.nonzero(np.any(sb != so, axis=1))[0]
        flips = [dict(config=ci, realisation=int(m), product=sb[m].tolist(), oracle=so[m].tolist(),
                      max_rel=float(rel[:, :, m].max())) for m in flipped]
        flips_all += flips
        cfg = dict(config=ci, kind="mixture" if kind == 1 else "gamma", N=N, tol=tol,
                   quantiles_compared=int(ok.sum()), within_1e6=int((rel[ok] <= 1e-6).sum()),
                   max_rel=float(rel[ok].max()) if ok.any() else 0.0,
                   finite_mismatch=int((fin_b != fin_o).sum()),
                   failures_product=[int(v) for v in fb],
                   failures_oracle=[int((~fin_o[p, 0]).sum()) for p in (0, 1)],
                   flipped_refinements=len(flips), z_unconverged=int(st["z_unconverged"]))
        report["configs"].append(cfg)
        tot_q += cfg["quantiles_compared"]; tot_in += cfg["within_1e6"]
        worst = max(worst, cfg["max_rel"])
    report.update(realisations=M * len(CONFIGS), quantiles_compared=tot_q,
                  fraction_within_1e6=tot_in / max(tot_q, 1), max_rel=worst,
                  flipped_refinements=flips_all, seconds=time.time() - t0)
    return report


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--make-golden", action="store_true")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--emu", action="store_true")
    ap.add_argument("--limit", type=int, default=None)
    ap.add_argument("--seed", type=int, default=None, help="another experiment (with --golden: not the committed file)")
    ap.add_argument("--golden", default=None, help="path of the oracle results (default: tests/golden/)")
    ap.add_argument("--m", type=int, default=None, help="realisations per configuration when making a golden file")
    a = ap.parse_args()
    if a.seed is not None:
        SEED = a.seed
    if a.golden is not None:
        GOLDEN = a.golden
    if a.m is not None:
        M_PER_CONFIG = a.m
    if a.make_golden:
        make_golden(a.jobs)
    else:
        print(json.dumps(sweep(a.emu, a.limit)))
