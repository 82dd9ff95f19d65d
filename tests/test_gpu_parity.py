"""
Parity of the CUDA path (through the C ABI, librhfq_b200.so on a B200) against
the CPU oracle on the same seeded inputs.  Tolerances are north_star's:
relative 1e-8 on pdf / cdf / tail, 1e-6 on tail quantiles.
"""
import json
import os

import numpy as np
import pytest

from conftest import gamma_dataset, DEFAULT_PRIOR, ROOT
from oracle.oracle import OraclePosterior

pytestmark = pytest.mark.gpu

RTOL_PDF = 1e-8
RTOL_Q = 1e-6


def Batch(*a, **k):
    from reheatfunq_b200 import PosteriorBatch
    return PosteriorBatch(*a, **k)


def compare(B, r, O, nx=200, explicit=False):
    Q = B.Qmax()[r]
    assert Q == O.Qmax()
    PH = np.linspace(0, Q, nx)
    pb, po = B.pdf(PH)[r], O.pdf(PH)
    m = po > 0
    assert (np.abs(pb[m] - po[m]) / po[m]).max() < RTOL_PDF
    assert np.all(pb[~m] == 0)
    if explicit:
        return
    cb, co = B.cdf(PH)[r], O.cdf(PH)
    m = co > 0
    assert (np.abs(cb[m] - co[m]) / co[m]).max() < RTOL_PDF
    tb, to = B.tail(PH)[r], O.tail(PH)
    m = to > 1e-290
    assert (np.abs(tb[m] - to[m]) / to[m]).max() < RTOL_PDF
    qs = np.linspace(0.01, 0.99, 99)
    qb, qo = B.tail_quantiles(qs)[r], O.tail_quantiles(qs)
    assert (np.abs(qb - qo) / qo).max() < RTOL_Q


def test_golden_absolute_values():
    """testing/test_anomaly_posterior.py:261-337 through the CUDA explicit algorithm."""
    from test_oracle_golden import P1, Q1, C1, PH1, REF1
    B = Batch([[(Q1, C1)]], [[1.0]], (P1["p"], P1["s"], P1["n"], P1["v"]),
              pdf_algorithm="explicit")
    assert B.status()[0] == 0
    # the reference's gate is 1e-11 in long double; FP64 on the device reaches ~1e-12
    np.testing.assert_allclose(B.pdf(PH1)[0], REF1, rtol=1e-10)


def test_golden_mpmath():
    path = os.path.join(ROOT, "tests", "golden", "mpm_golden.json")
    g = json.load(open(path))["A"]
    pr = g["prior"]
    for alg, rt in (("explicit", 1e-10), ("barycentric_lagrange", 1e-6)):
        B = Batch([[(np.array(g["q"]), np.array(g["c"]))]], [[1.0]],
                  (pr["p"], pr["s"], pr["n"], pr["v"]), pdf_algorithm=alg)
        x = np.array(g["PH"])
        np.testing.assert_allclose(B.pdf(x)[0], g["pdf"], rtol=rt)
        np.testing.assert_allclose(B.cdf(x)[0], g["cdf"], rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("N,seed", [(10, 1), (25, 29189), (50, 29189), (100, 3)])
def test_single_posterior(N, seed, monkeypatch):
    """Config 1 of BASELINE.json (N = 25, seed 29189) and neighbours."""
    q, c = gamma_dataset(N, seed)
    monkeypatch.setenv("RHFQ_KERNEL_TIMERS", "1")       # event pairs around the node launches
    B = Batch([[(q, c)]], [[1.0]], DEFAULT_PRIOR)
    assert B.status()[0] == 0
    O = OraclePosterior([q], [c], [1.0], *DEFAULT_PRIOR, 1.0, 1e-8, precision="double")
    for which in (0, 1):
        xo, fo, info = O.bli(which)
        fb = B.bli_samples(0, which)
        assert len(fb) == len(fo)
        assert np.abs(fb - fo).max() < 1e-10
    compare(B, 0, O, nx=1000 if N == 25 else 200)
    cnt = B.counters()
    assert cnt["nodes"] > 0 and cnt["node_kernel_launches"] > 0 and cnt["node_kernel_ns"] > 0


def test_N1000():
    q, c = gamma_dataset(1000, 294982299, 120.0, 4.0)
    B = Batch([[(q, c)]], [[1.0]], DEFAULT_PRIOR)
    O = OraclePosterior([q], [c], [1.0], *DEFAULT_PRIOR, 1.0, 1e-8, precision="double")
    compare(B, 0, O, nx=20)


def test_explicit_and_simpson_algorithms():
    q, c = gamma_dataset(25, 29189)
    for alg in ("explicit", "adaptive_simpson"):
        B = Batch([[(q, c)]], [[1.0]], DEFAULT_PRIOR, pdf_algorithm=alg)
        O = OraclePosterior([q], [c], [1.0], *DEFAULT_PRIOR, 1.0, 1e-8, pdf_algorithm=alg,
                            precision="double")
        PH = np.linspace(0, B.Qmax()[0], 60)
        pb, po = B.pdf(PH)[0], O.pdf(PH)
        m = po > 0
        tol = RTOL_PDF if alg == "explicit" else 1e-6   # Simpson density: leaf-level flips
        assert (np.abs(pb[m] - po[m]) / po[m]).max() < tol
        cb, co = B.cdf(PH)[0], O.cdf(PH)
        m = co > 0
        assert (np.abs(cb[m] - co[m]) / co[m]).max() < RTOL_PDF


def test_multi_set_weighting():
    rng = np.random.default_rng(5)
    q, c = gamma_dataset(30, 11)
    sets, ws = [], []
    for m in range(16):
        idx = np.sort(rng.choice(30, size=24, replace=False))
        sets.append((q[idx], c[idx]))
        ws.append(0.5 + rng.random())
    B = Batch([sets], [ws], DEFAULT_PRIOR, bli_max_refinements=5)
    assert B.status()[0] == 0
    O = OraclePosterior([s[0] for s in sets], [s[1] for s in sets], ws, *DEFAULT_PRIOR, 1.0,
                        1e-8, bli_max_refinements=5, precision="double")
    compare(B, 0, O, nx=80)


def test_batch_of_regions():
    """Config 3 in miniature: ragged N, per-posterior priors, one launch sequence."""
    rng = np.random.default_rng(1)
    Ns = rng.integers(10, 101, size=24)
    data = [gamma_dataset(int(N), 100 + i, float(np.exp(rng.uniform(np.log(10), np.log(200)))))
            for i, N in enumerate(Ns)]
    B = Batch([[d] for d in data], [[1.0]] * len(data), DEFAULT_PRIOR)
    assert np.all(B.status() == 0)
    qs = np.array([0.1, 0.5, 0.9])
    tq = B.tail_quantiles(qs)
    for r in range(0, len(data), 3):
        q, c = data[r]
        O = OraclePosterior([q], [c], [1.0], *DEFAULT_PRIOR, 1.0, 1e-8, precision="double")
        ref = O.tail_quantiles(qs)
        assert (np.abs(tq[r] - ref) / ref).max() < RTOL_Q
    # size-independent property: tail(quantile(t)) == t, cdf + tail == 1
    t2 = B.tail(tq)
    np.testing.assert_allclose(t2, np.broadcast_to(qs, t2.shape), rtol=1e-10)
    x = np.linspace(0.05, 0.95, 7)[None, :] * B.Qmax()[:, None]
    np.testing.assert_allclose(B.cdf(x) + B.tail(x), 1.0, rtol=0, atol=1e-14)


def test_statuses_and_edge_cases():
    q, c = gamma_dataset(12, 31)
    bad_q = q.copy(); bad_q[3] = -1.0
    dup = (np.array([1.0, 2.0, 3.0]), np.array([1e-9, 2e-9, 1e-9]))
    B = Batch([[(q, c)], [(bad_q, c)], [(q, -np.abs(c))], [dup], [(q, c)],
               [(np.array([60.0]), np.array([1e-9]))]],
              [[1.0], [1.0], [1.0], [1.0], [-1.0], [1.0]], (1.0, 0.0, 0.0, 0.0),
              bli_max_refinements=3, rtol=1e-3)
    st = B.status()
    assert st[0] == 0 and np.all(st[1:5] == 2) and st[5] == 3
    out, qst = B.tail_quantiles(np.array([0.0, 1.0, 0.5]), return_status=True)
    assert out[0, 0] == B.Qmax()[0] and out[0, 1] == 0.0 and 0 < out[0, 2] < B.Qmax()[0]
    assert np.all(np.isnan(out[1:]))
    out, qst = B.tail_quantiles(np.array([1.5]), return_status=True)
    assert qst[0] == 2 and np.isnan(out[0, 0])
    Q = B.Qmax()[0]
    assert np.all(B.pdf(np.array([-1.0, Q, 2 * Q]))[0] == 0)
    np.testing.assert_array_equal(B.cdf(np.array([-1.0, 0.0, Q, 2 * Q]))[0], [0, 0, 1, 1])


def test_legacy_bayes_api_gpu():
    """The reference's known-answer test (test_anomaly_posterior.py:261-337) as written there,
    through the legacy-API mirror on the GPU."""
    from reheatfunq_b200.anomaly.bayes import marginal_posterior_pdf
    from test_oracle_golden import P1, Q1, C1, PH1, REF1
    pdf_compare = marginal_posterior_pdf(np.array(PH1), P1["p"], P1["s"], P1["n"], P1["v"],
                                         np.array(Q1), np.array(C1), amin=1.0, dest_tol=1e-15,
                                         working_precision="long double")
    np.testing.assert_allclose(pdf_compare, REF1, rtol=1e-11)


def test_dmin_restricted_mixture_posterior():
    """SURVEY 8f item 2 end to end through the product library: enumerate the d_min-restricted
    samples (host code of librhfq_b200.so), build the weighted mixture posterior on the GPU the
    way posterior.py:299-368 does, compare with the oracle on the same samples."""
    from reheatfunq_b200.shims import rdisks
    from test_rdisks import CASES, _as_set
    for xy, dmin, expected in CASES:
        res = rdisks.all_restricted_samples(xy, dmin, 100, 10000, 893289, True)
        assert _as_set(res) == sorted((round(p, 14), s) for p, s in expected)
    rng = np.random.default_rng(4711)
    N = 16
    q, c = gamma_dataset(N, 991)
    xy = 60e3 * (rng.random((N, 2)) - 0.5)
    samples = rdisks.all_restricted_samples(xy, 12e3, 2000, 2000, 5)
    assert len(samples) > 1
    w = np.array([s[0] for s in samples])
    sets = [(q[s[1]], c[s[1]]) for s in samples]
    B = Batch([sets], [w], DEFAULT_PRIOR)
    assert B.status()[0] == 0
    O = OraclePosterior([s[0] for s in sets], [s[1] for s in sets], w, *DEFAULT_PRIOR, 1.0, 1e-8,
                        precision="double")
    compare(B, 0, O)
    ph_i = q / c
    assert rdisks.determine_global_PHmax(xy, ph_i, 12e3) >= B.Qmax()[0] * (1 - 1e-15)


def test_fine_grid_and_lgamma_tables_against_their_direct_forms(monkeypatch):
    """The two tabulate-and-interpolate accelerations (fine theta grid of the interpolants,
    lgamma-difference tables) against the direct evaluations they replace, on the device."""
    data = [gamma_dataset(N, s) for N, s in ((10, 51), (30, 52), (100, 53), (300, 54))]
    mk = lambda **k: Batch([[d] for d in data], [[1.0]] * 4, DEFAULT_PRIOR, **k)
    qs = np.array([0.01, 0.5, 0.99])
    B = mk()
    x = np.linspace(0.02, 0.98, 25)[None, :] * B.Qmax()[:, None]
    ref_pdf, ref_cdf, ref_q = B.pdf(x), B.cdf(x), B.tail_quantiles(qs)
    cnt = B.counters()
    assert cnt["fine_bad"] == 0 and cnt["fine_evals"] > 0.99 * 2 * cnt["saq_steps"]
    monkeypatch.setenv("RHFQ_NO_FINE", "1")
    B1 = mk()
    np.testing.assert_allclose(B1.cdf(x), ref_cdf, rtol=1e-11)
    np.testing.assert_allclose(B1.tail_quantiles(qs), ref_q, rtol=1e-11)
    assert abs(B1.saq_leaves() - B.saq_leaves()) <= 2      # borderline accept decisions
    monkeypatch.delenv("RHFQ_NO_FINE")
    monkeypatch.setenv("RHFQ_NO_GTAB", "1")
    B2 = mk()
    np.testing.assert_allclose(B2.pdf(x), ref_pdf, rtol=2e-11)
    np.testing.assert_allclose(B2.cdf(x), ref_cdf, rtol=2e-11)
    monkeypatch.setenv("RHFQ_SAQ_FRONTIER", "1")
    B3 = mk()
    np.testing.assert_allclose(B3.cdf(x), B2.cdf(x), rtol=1e-13)


def test_simpson_levels_device_vs_host_controlled(monkeypatch):
    """The Simpson trees built by one persistent CTA per posterior (fine grids staged in shared
    memory, CTA-private frontier) against the level-synchronous launches queued from a
    device-side control block and against the host reading every level's frontier length;
    small blocks, frontier and pool estimates that overflow (redo path), fine grids that do
    not fit the shared memory (read from global memory instead)."""
    data = [gamma_dataset(N, s) for N, s in ((10, 81), (30, 82), (100, 83), (20, 84), (55, 85))]
    mk = lambda: Batch([[d] for d in data], [[1.0]] * 5, DEFAULT_PRIOR)
    qs = np.array([0.01, 0.5, 0.99])
    B = mk()
    x = np.linspace(0.02, 0.98, 25)[None, :] * B.Qmax()[:, None]
    ref_cdf, ref_tail, ref_q = B.cdf(x), B.tail(x), B.tail_quantiles(qs)
    for env in ({"RHFQ_SAQ_HOSTCTL": "1"}, {"RHFQ_SAQ_BLOCK": "2"},
                {"RHFQ_SAQ_FRONTIER": "1", "RHFQ_SAQ_NODES": "16"},
                {"RHFQ_SAQ_LEVELSYNC": "1"}, {"RHFQ_SAQ_LEVELSYNC": "1", "RHFQ_SAQ_BLOCK": "2"},
                {"RHFQ_SAQ_LEVELSYNC": "1", "RHFQ_SAQ_FRONTIER": "1", "RHFQ_SAQ_NODES": "16"},
                {"RHFQ_SAQP_SMEM_KB": "8"}, {"RHFQ_SAQP_SMEM_KB": "1"}, {"RHFQ_SAQP_CTAS": "1"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        B1 = mk()
        # values do not depend on the order in which the atomics hand out pool slots
        np.testing.assert_array_equal(B1.cdf(x), ref_cdf)
        np.testing.assert_array_equal(B1.tail(x), ref_tail)
        np.testing.assert_array_equal(B1.tail_quantiles(qs), ref_q)
        assert B1.saq_leaves() == B.saq_leaves()
        assert B1.counters()["saq_steps"] == B.counters()["saq_steps"]
        for k in env:
            monkeypatch.delenv(k)


def test_split_node_kernels_match_fused_on_device(monkeypatch):
    """prep / mode / window / Gauss-Legendre launches against the single node kernel."""
    data = [gamma_dataset(N, s) for N, s in ((4, 71), (25, 72), (120, 73))]
    mk = lambda: Batch([[d] for d in data] + [[data[0], data[2]]], [[1.0]] * 3 + [[0.3, 0.7]],
                       DEFAULT_PRIOR, pdf_algorithm="explicit")
    monkeypatch.setenv("RHFQ_NO_QUAD_CHECK", "1")      # the sampled precision check adds evaluations
    B = mk()
    x = np.linspace(0.0, 1.0, 41)[None, :] * B.Qmax()[:, None]
    p_split = B.pdf(x)
    monkeypatch.setenv("RHFQ_FUSED_NODES", "1")
    Bf = mk()
    p_fused = Bf.pdf(x)
    m = p_fused > 0
    np.testing.assert_allclose(p_split[m], p_fused[m], rtol=1e-13)
    assert np.array_equal(p_split[~m], p_fused[~m])
    for key in ("nodes", "lz_nodes", "g_evals", "newton", "nsum"):
        assert B.counters()[key] == Bf.counters()[key], key


def test_full_size_config3_properties():
    """BASELINE.json config 3 at its full size (10^4 regions, the bench workload) through the
    packed C ABI: size-independent properties (tail(quantile(t)) = t, cdf + tail = 1,
    monotone quantiles and cdf) plus oracle parity on a few regions picked across the batch."""
    import bench
    R = 10000
    from reheatfunq_b200 import PosteriorBatch
    q, c, off = bench.make_regions(R, seed=1)
    B = PosteriorBatch.from_packed(q, c, off, np.ones(R), np.arange(R + 1, dtype=np.int64),
                                   DEFAULT_PRIOR)
    assert np.all(B.status() == 0)
    qs = np.array([0.001, 0.1, 0.5, 0.9, 0.999])
    tq = B.tail_quantiles(qs)
    assert np.all(np.isfinite(tq)) and np.all(np.diff(tq, axis=1) < 0)      # decreasing in t
    np.testing.assert_allclose(B.tail(tq), np.broadcast_to(qs, tq.shape), rtol=1e-10)
    x = np.linspace(0.02, 0.98, 9)[None, :] * B.Qmax()[:, None]
    cdf, tail = B.cdf(x), B.tail(x)
    np.testing.assert_allclose(cdf + tail, 1.0, rtol=0, atol=1e-14)
    assert np.all(np.diff(cdf, axis=1) >= -1e-15)      # up to rounding (a few ulp) where cdf -> 1
    for r in (0, 1234, 5000, 9999):
        qr, cr = q[off[r]:off[r + 1]], c[off[r]:off[r + 1]]
        O = OraclePosterior([qr], [cr], [1.0], *DEFAULT_PRIOR, 1.0, 1e-8, precision="double")
        ref = O.tail_quantiles(qs[1:4])
        assert (np.abs(tq[r, 1:4] - ref) / ref).max() < RTOL_Q
        np.testing.assert_allclose(cdf[r], O.cdf(x[r]), rtol=1e-8)
        np.testing.assert_allclose(tail[r], O.tail(x[r]), rtol=1e-8)


def _assert_sweep(res):
    """The bar: pdf / cdf / tail within 1e-8 and tail quantiles within 1e-6 of the oracle in
    EVERY case -- no tolerated exceptions.  tools/parity_sweep.py classifies each case:
    "natural" (against the double oracle as it stands), "ymax" (the reference takes y_max from a
    2-bit TOMS748 bracket, large_z.hpp:338-348, whose last iterate flips with rounding-level
    changes of the root function; product and oracle then hold two different ADMISSIBLE values
    and are compared at the same one), "referee" (the double oracle's adaptive rule stopped
    short at a node: a quantity that misses the double oracle must then agree with the
    LONG-DOUBLE oracle, with the double oracle the farther one), "z-unconverged" (the z-quadrature ran out of levels on
    BOTH sides; the reference does not check, localsandnorm.hpp:221-226).  Any "FAIL" fails."""
    v = res["verdicts"]
    assert v.get("FAIL", 0) == 0, [r for r in res["not_natural"] if r["verdict"] == "FAIL"]
    w = res["worst_after_protocol"]
    assert w["pdf"] <= 1e-8 and w["cdf"] <= 1e-8 and w["tail"] <= 1e-8 and w["q"] <= 1e-6, w
    assert res["ymax_ratio_max"] <= 1.5
    # the protocol must stay the exception: > 98 % of the cases pass against the plain oracle
    assert v.get("natural", 0) >= 0.98 * res["cases"], v
    assert v.get("z-unconverged", 0) <= 0.005 * res["cases"] + 1, v


def test_randomised_parity_sweep():
    """1000 random posteriors (random N in [3, 300], gamma shapes, four priors, amin, anomaly
    powers, 1-5 weighted sample sets, rtol 1e-8 / 1e-5) through the C ABI against the oracle,
    zero exceptions; the result is written to gpurun_out/ for profiles/."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "parity_sweep.py"), "1000", "7",
                          "--jobs", str(min(16, os.cpu_count() or 1))],
                         capture_output=True, text=True, check=True).stdout
    res = json.loads(out.strip().splitlines()[-1])
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_sweep_1000_seed7.json"), "w") as f:
            json.dump(res, f, indent=1)
    except OSError:
        pass
    assert res["cases"] == 1000 and not res["emu"]
    _assert_sweep(res)


def test_keyword_double_sweep():
    """The reference's precision="double" keyword runs its LONG DOUBLE build
    (postbackend.pyx:219-222).  The product then takes the thresholds that are powers of the
    machine epsilon of the reference's type from long double (y_max criterion, tanh-sinh
    tolerance of the norm, Simpson tolerance, PointInInterval threshold) while its arithmetic
    stays IEEE double: 200 random posteriors against the long-double oracle at the same
    north-star tolerances, zero exceptions; the report goes to gpurun_out/."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "parity_sweep.py"), "200", "11",
                          "--keyword-double", "--jobs", str(min(16, os.cpu_count() or 1))],
                         capture_output=True, text=True, check=True).stdout
    res = json.loads(out.strip().splitlines()[-1])
    try:
        with open(os.path.join(ROOT, "gpurun_out", "parity_sweep_keyword_double_200_seed11.json"), "w") as f:
            json.dump(res, f, indent=1)
    except OSError:
        pass
    assert res["cases"] == 200 and not res["emu"]
    _assert_sweep(res)


def test_locals_by_warp_match_sequential(monkeypatch):
    """The per-set statistics computed by a warp per set (element-wise work dealt out to the
    lanes, order-dependent sums run redundantly over the stored values) are bit-identical with
    one thread running the sequential compute_locals (locals.hpp:94-113,193-329) on the device,
    k_i included (the explicit pdf is bit-identical too)."""
    data = [gamma_dataset(N, s) for N, s in ((1, 7), (2, 8), (31, 9), (32, 10), (33, 11), (100, 12),
                                             (257, 13))]
    mk = lambda: Batch([[x] for x in data], [[1.0]] * len(data), DEFAULT_PRIOR,
                       pdf_algorithm="explicit")
    B = mk()
    monkeypatch.setenv("RHFQ_LOCALS_SEQ", "1")
    Bs = mk()
    assert np.array_equal(B.status(), Bs.status())
    for i in range(len(data)):
        a, b = B.locals(i), Bs.locals(i)
        for name in a:
            assert a[name] == b[name] or (a[name] != a[name] and b[name] != b[name]), (i, name)
    x = np.linspace(0.05, 0.95, 7)[None, :] * B.Qmax()[:, None]
    assert np.array_equal(B.pdf(x), Bs.pdf(x))


def test_failure_semantics_on_device():
    """The enumerated inputs on which the reference algorithm raises
    (tests/golden/failure_cases.json): same statuses on the CUDA path as on the host emulation."""
    from reheatfunq_b200 import _lib
    from test_engine_emu import check_failure_semantics
    check_failure_semantics(_lib.api())


def test_sampled_precision_check_counts():
    """The sampled precision check of the alpha-quadrature runs on the device (one warp in 61)
    and raises no false alarm on the bench workload's kind of input."""
    import bench
    from reheatfunq_b200 import PosteriorBatch
    q, c, off = bench.make_regions(300, seed=5)
    B = PosteriorBatch.from_packed(q, c, off, np.ones(300), np.arange(301, dtype=np.int64), DEFAULT_PRIOR)
    assert np.all(B.status() == 0)
    cnt = B.counters()
    assert cnt["quad_checks"] > 0.004 * (cnt["nodes"] + cnt["lz_nodes"])


def test_legacy_bayes_adapters_gpu():
    """The rest of the legacy functional API (bayes.pyx:187-640) on the CUDA path: cdf, tail,
    tail quantiles and the *_batch forms against the oracle."""
    from reheatfunq_b200.anomaly import bayes
    from test_oracle_golden import P1, Q1, C1, PH1
    PH = np.array(PH1)
    q2, c2 = np.array(Q1[:10]), np.array(C1[:10])
    args = (P1["p"], P1["s"], P1["n"], P1["v"])
    O1 = OraclePosterior([np.array(Q1)], [np.array(C1)], [1.0], *args, 1.0, 1e-8, precision="double")
    O2 = OraclePosterior([q2], [c2], [1.0], *args, 1.0, 1e-8, precision="double")
    ts = np.array([0.9, 0.5, 0.1])
    for name, ref1, ref2, x, tol in (("cdf", O1.cdf, O2.cdf, PH, RTOL_PDF), ("tail", O1.tail, O2.tail, PH, RTOL_PDF),
                                     ("tail_quantiles", O1.tail_quantiles, O2.tail_quantiles, ts, RTOL_Q)):
        single = getattr(bayes, "marginal_posterior_" + name)(x, *args, Q1, C1)
        batch = getattr(bayes, "marginal_posterior_" + name + "_batch")(x, *args, [Q1, q2], [C1, c2])
        # pdf / cdf / tail come back as long double, quantiles as double (bayes.pyx:570,624)
        assert single.dtype == (np.float64 if name == "tail_quantiles" else np.longdouble)
        assert batch.shape == (2, len(x))
        r1 = ref1(x)
        m = r1 > 1e-300
        np.testing.assert_allclose(np.asarray(single, dtype=float)[m], r1[m], rtol=tol)
        np.testing.assert_allclose(np.asarray(batch[0], dtype=float)[m], r1[m], rtol=tol)
        xs = x[x < O2.Qmax()] if name != "tail_quantiles" else x
        b2 = getattr(bayes, "marginal_posterior_" + name + "_batch")(xs, *args, [Q1, q2], [C1, c2])
        r2 = ref2(xs)
        m2 = r2 > 1e-300
        np.testing.assert_allclose(np.asarray(b2[1], dtype=float)[m2], r2[m2], rtol=tol)
