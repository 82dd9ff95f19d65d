"""Randomised parity sweep: posteriors with random N, gamma shapes, priors, amin, anomaly
strengths and (sometimes) several weighted sample sets, the product (GPU through the C ABI;
`--emu`: the TEST-ONLY host emulation of the same stage bodies, for debugging without a GPU)
against the CPU oracle.  Prints the worst relative deviations and every case beyond the
north-star tolerances (1e-8 on pdf/cdf/tail, 1e-6 on tail quantiles).

    python tools/parity_sweep.py N_CASES [SEED] [--emu] [--jobs J] [--only i,j,...] [--ld]

`--ld` adds the comparison with the long-double oracle (the reference's "double" keyword,
postbackend.pyx:219-222).
"""
import os, sys, time, json
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import numpy as np

DEFAULT_PRIOR = (2.522017292522833465, 15.37301672440559130,
                 0.2184765374862465137, 0.2184765374862465137)
PRIORS = [DEFAULT_PRIOR, (1.0, 0.0, 0.0, 0.0), (3.0, 40.0, 2.5, 2.0), (10.0, 500.0, 8.0, 7.5)]
TOL = dict(pdf=1e-8, cdf=1e-8, tail=1e-8, q=1e-6)


def ls1980(x, d=14e3, L=160e3):
    dist = np.abs(x)
    return 2 / (np.pi * d * L) * (1 - dist / d * np.arctan(d / dist))


def gamma_dataset(N, seed, alpha=40.0, beta=0.4, P=100e6):
    rng = np.random.default_rng(seed)
    Q = np.sort(rng.gamma(alpha, size=N) / beta)
    x = 160e3 * (rng.random(N) - 0.5)
    c = ls1980(x)
    return Q + 1e3 * P * c, c * 1e3


def gen_cases(n_cases, seed):
    """The case list is a pure function of (n_cases, seed): the same draws as the round-1 tool."""
    rng = np.random.default_rng(seed)
    cases = []
    for case in range(n_cases):
        M = 1 if rng.random() < 0.7 else int(rng.integers(2, 6))
        N = int(np.exp(rng.uniform(np.log(3), np.log(300))))
        alpha = float(np.exp(rng.uniform(np.log(5), np.log(300))))
        P = float(10 ** rng.uniform(6.5, 8.5))
        prior = PRIORS[int(rng.integers(0, len(PRIORS)))]
        amin = float(rng.choice([1.0, 1.0, 0.5, 2.0]))
        rtol = float(rng.choice([1e-8, 1e-8, 1e-5]))
        seeds, w = [], []
        for m in range(M):
            seeds.append(int(rng.integers(1, 10**6)))
            w.append(float(rng.uniform(0.2, 1.0)))
        cases.append(dict(case=case, M=M, N=N, alpha=alpha, P=P, prior=prior, amin=amin, rtol=rtol,
                          seeds=seeds, w=w))
    return cases


def case_sets(cs):
    return [gamma_dataset(cs["N"], s, cs["alpha"], cs["alpha"] / 68.3, cs["P"]) for s in cs["seeds"]]


def query_points(Q):
    x = np.concatenate([np.linspace(0, 1, 61)[1:-1], [0.999, 0.9999]]) * Q
    qs = np.array([0.01, 0.1, 0.5, 0.9, 0.99])
    return x, qs


def _api(emu):
    if not emu:
        return None
    import ctypes as C
    from reheatfunq_b200._lib import Api
    return Api(C.CDLL(os.path.join(HERE, "..", "tests", "hostmath", "libengine_emu.so")), "emu_rhfq_")


def rel_dev(b, o):
    m = o > 1e-280
    return float(np.max(np.abs(b[m] - o[m]) / o[m])) if m.any() else 0.0


def _oracle_points(f, x, Q, raised):
    """f(x) of the oracle; if it raises, point by point.  The one place where the reference
    itself raises for a valid argument is the seam of its two interpolants: pdf at x = 0.9 Qmax
    can fail with "x out of bounds in BarycentricLagrangeInterpolator" in the long double build,
    when the double x lies a fraction of an ulp above the long double 0.9 Qmax while log(Qmax - x)
    rounds to the seam's value (posterior.hpp:981-987, barylagrint.hpp:134-137).  Such points are
    dropped from the comparison and reported; a raise anywhere else propagates."""
    try:
        return f(x), np.ones(x.shape, dtype=bool)
    except RuntimeError:
        v, ok = np.full(x.shape, np.nan), np.ones(x.shape, dtype=bool)
        for i, xi in enumerate(x):
            try:
                v[i] = f(np.array([xi]))[0]
            except RuntimeError:
                if abs(xi / Q - 0.9) > 1e-15:
                    raise
                ok[i] = False
                raised.append(float(xi / Q))
        return v, ok


def _compare(B, O, x, qs, vals=None, raised=None):
    if vals is None:
        vals = dict(pdf=B.pdf(x)[0], cdf=B.cdf(x)[0], tail=B.tail(x)[0], q=B.tail_quantiles(qs)[0])
    raised = [] if raised is None else raised
    Q = O.Qmax()
    ref, res = dict(q=O.tail_quantiles(qs)), {}
    for k, f in (("pdf", O.pdf), ("cdf", O.cdf), ("tail", O.tail)):
        ref[k], ok = _oracle_points(f, x, Q, raised)
        res[k] = rel_dev(vals[k][ok], ref[k][ok])
    res["q"] = float(np.max(np.abs(vals["q"] - ref["q"]) / ref["q"]))
    return res, vals, ref


def _within(res, tol=TOL):
    return all(res[k] <= tol[k] for k in tol)


# y_max is the centre of a TOMS748 bracket [a, b] around the root y* with b - a <= a / 2
# (eps_tolerance(2), large_z.hpp:338-348): every admissible value lies in [5/6, 5/4] y*, two
# admissible values within a factor 3/2 of each other.
YMAX_RATIO = 1.5
# the z-quadrature of the norm stops at level 12 in the oracle (Boost: 15) whether or not it has
# converged; where it has not, the reference's number is the last iterate of a diverging
# sequence and only a loose comparison is meaningful
TOL_Z_UNCONVERGED = dict(pdf=1e-5, cdf=1e-5, tail=1e-5, q=1e-5)


def run_case_ld(cs, emu=False):
    """The reference's "double" keyword (= its long double build, postbackend.pyx:219-222): the
    product with the long-double thresholds against the LONG-DOUBLE oracle as it stands; if the
    two hold different admissible y_max, against the long-double oracle at the product's."""
    from reheatfunq_b200 import PosteriorBatch
    from oracle.oracle import OraclePosterior
    api = _api(emu)
    sets, w, prior, amin, rtol = case_sets(cs), cs["w"], cs["prior"], cs["amin"], cs["rtol"]
    qij, cij = [s[0] for s in sets], [s[1] for s in sets]
    out = dict(case=cs["case"], N=cs["N"], M=cs["M"], rtol=rtol, amin=amin, prior=list(prior))
    try:
        O = OraclePosterior(qij, cij, w, *prior, amin, rtol, precision="long double")
    except RuntimeError as e:
        B = PosteriorBatch([sets], [w], prior, amin=amin, rtol=rtol, api=api,
                           reference_arithmetic="long double")
        out["verdict"] = "skipped" if B.status()[0] != 0 else "FAIL"
        out["what"] = "oracle: " + str(e)[:80] + "; product status %d" % B.status()[0]
        return out
    B = PosteriorBatch([sets], [w], prior, amin=amin, rtol=rtol, api=api,
                       reference_arithmetic="long double")
    if B.status()[0] != 0:
        out["verdict"] = "FAIL"
        out["what"] = "product status %d, oracle fine" % B.status()[0]
        return out
    x, qs = query_points(B.Qmax()[0])
    raised = []
    res, vals, ref = _compare(B, O, x, qs, raised=raised)
    out["res"] = res
    if raised:
        out["oracle_raises_at_x_over_Q"] = raised       # the reference's seam quirk, see _oracle_points
        assert np.all(np.isfinite(vals["pdf"]))         # the product returns the value there
    Lb = [B.locals(l) for l in range(cs["M"])]
    Lo = [O.locals(l) for l in range(cs["M"])]
    out["ymax"] = [(Lb[l]["ymax"], Lo[l]["ymax"]) for l in range(cs["M"])]
    out["ymax_admissible"] = bool(max(max(a / b, b / a) for a, b in out["ymax"]) <= YMAX_RATIO)
    zb = any(int(L["flags"]) & 1 for L in Lb)
    zo = any(L["z_levels"] >= 12 and L["z_err"] > 3.2927225399135963e-10 * L["z_l1"] for L in Lo)
    if zb or zo:
        out["z_unconverged"] = dict(product=zb, oracle=zo)
        out["verdict"] = "z-unconverged" if (zb == zo and _within(res, TOL_Z_UNCONVERGED)) else "FAIL"
    elif _within(res):
        out["verdict"] = "natural"
    else:
        yb = [L["ymax"] for L in Lb]
        Oy = OraclePosterior(qij, cij, w, *prior, amin, rtol, precision="long double", ymax_override=yb)
        res_y, _, _ = _compare(B, Oy, x, qs, vals)
        out["res_ymax"] = res_y
        out["verdict"] = "ymax" if (_within(res_y) and out["ymax_admissible"]) else "FAIL"
    return out


def run_case(cs, emu=False, long_double=False, detail=True):
    """One posterior, product against the oracle.

    verdict (zero exceptions are tolerated by the test):
      "natural"   within tolerance of the double oracle as it stands
      "ymax"      the two sides took different (both admissible) y_max values from the 2-bit
                  bracket; within tolerance of the double oracle run at the product's y_max
      "referee"   the double oracle's adaptive rule is the outlier: the product is within tolerance
                  of the LONG-DOUBLE oracle (same y_max) and the double oracle is farther from it
      "z-unconverged"  both sides ran out of z-quadrature levels (see TOL_Z_UNCONVERGED)
      "skipped"   the oracle raised (and so did the product)
      "FAIL"      anything else
    """
    from reheatfunq_b200 import PosteriorBatch
    from oracle.oracle import OraclePosterior
    api = _api(emu)
    sets, w, prior, amin, rtol = case_sets(cs), cs["w"], cs["prior"], cs["amin"], cs["rtol"]
    qij, cij = [s[0] for s in sets], [s[1] for s in sets]
    out = dict(case=cs["case"], N=cs["N"], M=cs["M"], rtol=rtol, amin=amin, prior=list(prior))
    try:
        O = OraclePosterior(qij, cij, w, *prior, amin, rtol, precision="double")
    except RuntimeError as e:
        B = PosteriorBatch([sets], [w], prior, amin=amin, rtol=rtol, api=api)
        out["verdict"] = "skipped" if B.status()[0] != 0 else "FAIL"
        out["what"] = "oracle: " + str(e)[:80] + "; product status %d" % B.status()[0]
        return out
    B = PosteriorBatch([sets], [w], prior, amin=amin, rtol=rtol, api=api)
    if B.status()[0] != 0:
        out["verdict"] = "FAIL"
        out["what"] = "product status %d, oracle fine" % B.status()[0]
        return out
    Q = B.Qmax()[0]
    x, qs = query_points(Q)
    res, vals, ref = _compare(B, O, x, qs)
    out["res"] = res
    Lb = [B.locals(l) for l in range(cs["M"])]
    Lo = [O.locals(l) for l in range(cs["M"])]
    out["ymax"] = [(Lb[l]["ymax"], Lo[l]["ymax"]) for l in range(cs["M"])]
    ratio = [max(a / b, b / a) for a, b in out["ymax"]]
    out["ymax_admissible"] = bool(max(ratio) <= YMAX_RATIO)
    zb = any(int(L["flags"]) & 1 for L in Lb)
    zo = any(L["z_levels"] >= 12 and L["z_err"] > 1.4901161193847656e-08 * L["z_l1"] for L in Lo)
    if zb or zo:
        out["z_unconverged"] = dict(product=zb, oracle=zo)
        out["verdict"] = "z-unconverged" if (zb == zo and _within(res, TOL_Z_UNCONVERGED)) else "FAIL"
        return out
    if _within(res):
        out["verdict"] = "natural"
    else:
        yb = [Lb[l]["ymax"] for l in range(cs["M"])]
        Oy = OraclePosterior(qij, cij, w, *prior, amin, rtol, precision="double", ymax_override=yb)
        res_y, _, ref_y = _compare(B, Oy, x, qs, vals)
        out["res_ymax"] = res_y
        if _within(res_y) and out["ymax_admissible"]:
            out["verdict"] = "ymax"
        else:
            Ol = OraclePosterior(qij, cij, w, *prior, amin, rtol, precision="long double",
                                 ymax_override=yb)
            res_l, _, ref_l = _compare(B, Ol, x, qs, vals)
            res_ol = {k: rel_dev(ref_y[k], ref_l[k]) for k in ("pdf", "cdf", "tail")}
            res_ol["q"] = float(np.max(np.abs(ref_y["q"] - ref_l["q"]) / ref_l["q"]))
            # per quantity: within tolerance of the double oracle (same y_max), or -- where the
            # double oracle's adaptive rule is the outlier -- within tolerance of the long-double
            # oracle with the double oracle the farther one.  (cdf / tail / quantiles of the
            # long-double build differ from the double build's by design, through the Simpson
            # tolerance: they are only consulted where the double comparison fails.)
            out["res_vs_long_double"] = res_l
            out["oracle_vs_referee"] = res_ol
            final, ok = {}, out["ymax_admissible"]
            for k in TOL:
                if res_y[k] <= TOL[k]:
                    final[k] = res_y[k]
                else:
                    final[k] = res_l[k]
                    ok = ok and res_l[k] <= TOL[k] and res_ol[k] >= res_l[k]
            out["res_referee"] = final
            out["verdict"] = "referee" if ok else "FAIL"
    if long_double:
        # the reference's "double" keyword = long double arithmetic (postbackend.pyx:219-222)
        Ol = OraclePosterior(qij, cij, w, *prior, amin, rtol, precision="long double")
        res_ld, _, ref_ld = _compare(B, Ol, x, qs, vals)
        out["res_ld"] = res_ld
        out["ymax_ld"] = [Ol.locals(l)["ymax"] for l in range(cs["M"])]
        ro = {k: rel_dev(ref[k], ref_ld[k]) for k in ("pdf", "cdf", "tail")}
        ro["q"] = float(np.max(np.abs(ref["q"] - ref_ld["q"]) / ref_ld["q"]))
        out["oracle_double_vs_ld"] = ro
    if detail and out["verdict"] != "natural":
        b, o = vals["pdf"], ref["pdf"]
        m = o > 1e-280
        dev = np.where(m, np.abs(b - o) / np.where(m, o, 1.0), 0.0)
        lev = []
        for which in (0, 1):
            try:
                lev.append((len(B.bli_samples(0, which)), len(O.bli(which)[0])))
            except Exception as e:      # noqa
                lev.append(str(e)[:40])
        out["detail"] = dict(alpha=cs["alpha"], P=cs["P"],
                             Qmax_l_over_Q=[Lb[l]["Qmax"] / Q for l in range(cs["M"])],
                             x_over_Q=(x[np.argsort(dev)[-3:]] / Q).tolist(),
                             dev_bli=np.sort(dev)[-3:].tolist(), levels_product_oracle=lev)
    return out


def _worker(args):
    cs, emu, ld = args
    try:
        if ld == "keyword":
            return run_case_ld(cs, emu)
        return run_case(cs, emu, ld)
    except Exception as e:      # noqa
        return dict(case=cs["case"], verdict="FAIL", what="exception: " + str(e)[:200])


def sweep(n_cases, seed, emu=False, jobs=1, only=None, long_double=False):
    cases = gen_cases(n_cases, seed)
    if only is not None:
        cases = [cases[i] for i in only]
    t0 = time.time()
    if jobs > 1:
        import multiprocessing as mp
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        with mp.get_context("spawn").Pool(jobs) as pool:
            results = pool.map(_worker, [(cs, emu, long_double) for cs in cases], chunksize=4)
    else:
        results = [_worker((cs, emu, long_double)) for cs in cases]
    keys = ("pdf", "cdf", "tail", "q")
    zero = lambda: {k: 0.0 for k in keys}
    worst_natural, worst_final, worst_ld, worst_o = zero(), zero(), zero(), zero()
    verdicts, listed, ymax_ratio_max = {}, [], 1.0
    for r in results:
        v = r.get("verdict", "FAIL")
        verdicts[v] = verdicts.get(v, 0) + 1
        wide = "ymax" in r and max(max(a / b, b / a) for a, b in r["ymax"]) > YMAX_RATIO
        if v != "natural" or "oracle_raises_at_x_over_Q" in r or wide:
            listed.append(r)
        if "res" in r:
            for k in keys:
                worst_natural[k] = max(worst_natural[k], r["res"][k])
            final = {"natural": "res", "ymax": "res_ymax", "referee": "res_referee"}.get(v)
            if final:
                for k in keys:
                    worst_final[k] = max(worst_final[k], r[final][k])
            for (a, b) in r["ymax"]:
                ymax_ratio_max = max(ymax_ratio_max, a / b, b / a)
        if "res_ld" in r:
            for k in keys:
                worst_ld[k] = max(worst_ld[k], r["res_ld"][k])
                worst_o[k] = max(worst_o[k], r["oracle_double_vs_ld"][k])
    out = dict(cases=len(cases), seed=seed, emu=bool(emu), seconds=time.time() - t0,
               tolerance=TOL, verdicts=verdicts, worst_natural=worst_natural,
               worst_after_protocol=worst_final, ymax_ratio_max=ymax_ratio_max,
               not_natural=listed)
    out["reference_raises_at_seam"] = sum(1 for r in results if "oracle_raises_at_x_over_Q" in r)
    if long_double == "keyword":
        out["model"] = "the reference's long double build (precision='double' keyword)"
    if long_double is True:
        out["worst_vs_long_double_oracle"] = worst_ld
        out["oracle_double_vs_long_double"] = worst_o
    return out


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("n_cases", type=int, nargs="?", default=120)
    ap.add_argument("seed", type=int, nargs="?", default=20261020)
    ap.add_argument("--emu", action="store_true")
    ap.add_argument("--ld", action="store_true")
    ap.add_argument("--keyword-double", action="store_true",
                    help="the reference's precision='double' keyword: product with the long-double "
                         "thresholds against the long-double oracle")
    ap.add_argument("--jobs", type=int, default=1)
    ap.add_argument("--only", type=str, default=None)
    a = ap.parse_args()
    only = [int(v) for v in a.only.split(",")] if a.only else None
    print(json.dumps(sweep(a.n_cases, a.seed, a.emu, a.jobs, only, "keyword" if a.keyword_double else a.ld)))
