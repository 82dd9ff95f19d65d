#!/usr/bin/env python
"""
bench.py — batched anomaly-posterior tail quantiles on B200 (BASELINE.json config 3).

One "step" = one pass of the hot path over one batch of synthetic regions:
create R posteriors (locals -> norm -> BLI -> Simpson tree) and solve Nq tail
quantiles each.  Metric: posterior P_H evaluations per second (R * Nq / time).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--regions R] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); regions are independent, so
ranks shard them with no data-path collective (weak scaling: R regions per rank).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_PRIOR = (2.522017292522833465, 15.37301672440559130,
                 0.2184765374862465137, 0.2184765374862465137)
QUANTILES = np.array([0.1, 0.5, 0.9])


def make_regions(R, seed=1):
    """N_r ~ U[10,100], alpha_r ~ logU[10,200], beta_r = alpha_r/68.3,
    LS1980 anomaly of a straight fault, P = 100 MW.  CSR-packed."""
    rng = np.random.default_rng(seed)
    Ns = rng.integers(10, 101, size=R)
    alpha = np.exp(rng.uniform(np.log(10.0), np.log(200.0), size=R))
    off = np.zeros(R + 1, dtype=np.int64)
    off[1:] = np.cumsum(Ns)
    tot = int(off[-1])
    a_rep = np.repeat(alpha, Ns)
    Q = rng.gamma(a_rep) / (a_rep / 68.3)
    x = 160e3 * (rng.random(tot) - 0.5)
    dist = np.abs(x)
    d, L = 14e3, 160e3
    c = 2 / (np.pi * d * L) * (1 - dist / d * np.arctan(d / dist))
    q = Q + 1e3 * 100e6 * c
    return q, c * 1e3, off


def f_node_flops(cnt):
    """F_node = 27 (N + 1) + 110 I + 130 K, large-z nodes have no N-sum."""
    return (27.0 * (cnt["nsum"] + cnt["nodes"]) + 110.0 * cnt["newton"] + 130.0 * cnt["g_evals"])


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML in-process (spawning
    nvidia-smi during the timed region would itself perturb the measurement)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            return
        bits = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for nm, b in bits.items():
                    if r & b:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(1.0)   # NVML queries stall concurrent CUDA calls (~50 ms each): sample sparsely

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def cpu_reference_rate(q, c, off, n_regions, threads):
    """Times the CPU oracle (restatement of the reference's adaptive algorithm, OpenMP-free
    per posterior; regions spread over `threads` processes' worth of OpenMP threads)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle.oracle import OraclePosterior, lib
    lib().orc_set_num_threads(1)

    def one(r):
        a, b = off[r], off[r + 1]
        O = OraclePosterior([q[a:b]], [c[a:b]], [1.0], *DEFAULT_PRIOR, 1.0, 1e-8,
                            precision="double")
        return O.tail_quantiles(QUANTILES)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:   # ctypes releases the GIL
        list(ex.map(one, range(n_regions)))
    dt = time.perf_counter() - t0
    return n_regions * len(QUANTILES) / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g
    g.build_oracle()
    q, c, off = make_regions(args.regions, seed=1)
    threads = os.cpu_count() or 1
    sample = max(threads, args.ref_sample)
    vals = []
    for i in range(args.warmup + args.steps):
        lo = (i * sample) % max(1, args.regions - sample)
        sub_off = off[lo:lo + sample + 1] - off[lo]
        v, dt = cpu_reference_rate(q[off[lo]:off[lo + sample]], c[off[lo]:off[lo + sample]],
                                   sub_off, sample, threads)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    line = {
        "impl": "reference", "metric": "posterior_PH_evals_per_s", "value": value,
        "unit": "P_H evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3 batched regions: tail quantiles [0.1,0.5,0.9] of R "
                               "independent posteriors, N in [10,100], default prior, rtol 1e-8, "
                               "BLI r<=7", "regions_per_step": sample, "quantiles": 3},
        "cpu_baseline": {"value": value, "unit": "P_H evals/s", "cores": threads, "kind": "port",
                         "sample": "%d regions per step on %d host threads (oracle port of the "
                                   "reference's adaptive algorithm, double)" % (sample, threads)},
        "e2e": {"value": value, "unit": "P_H evals/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


MIX = (31, 3.5, 0.33, 62, 5.5, 0.67)         # A4-Resilience notebook, cell 20
Q41 = np.linspace(0.99, 0.01, 41)
RESIL_STREAMS = 64                           # RNG streams of the experiment (any rank count <= 64)


def _max_over_ranks(x, world, dist):
    if world == 1:
        return x
    import torch
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bench_resilience(args, world, rank, local_rank, dist, barrier):
    """BASELINE.json config 5 as stated: one experiment of M = 10^6 realisations (mixture heat
    flow, 41 quantiles, informed + flat prior, tol 1e-4) per sample size N, computed through
    reheatfunq_b200.sharding.resilience_sharded: rank r owns the RNG streams r, r + G, ..., the
    [2][41][M_local] blocks stay in HBM and ONE NCCL all-gather inside the timed region builds
    the [2][41][M] result on every rank.  Strong scaling: M is fixed as G grows.  Median of 3."""
    M = args.resilience_m
    if M <= 0:
        return None
    import torch
    from reheatfunq_b200.sharding import resilience_sharded
    sizes = [int(v) for v in args.resilience_n.split(",") if v]
    per_n, headline = [], None
    for N in sizes:
        # warm-up: two device chunks per rank bring the memory pool to its steady state
        # (a full-size run the first time: it page-locks the result buffer and grows the pools)
        resilience_sharded(1, MIX, N, M if N == sizes[0] else min(M, 65520 * world), 100.0, Q41,
                           DEFAULT_PRIOR, 1.0, 1e-4, 848782, nstreams=RESIL_STREAMS, dist=dist,
                           device=local_rank, host_result="rank0")
        runs = []
        for rep in range(3):
            tim = {}
            barrier()
            t0 = time.perf_counter()
            res, fail, st = resilience_sharded(1, MIX, N, M, 100.0, Q41, DEFAULT_PRIOR, 1.0, 1e-4,
                                               848782, nstreams=RESIL_STREAMS, dist=dist,
                                               device=local_rank, timings=tim, host_result="rank0")
            torch.cuda.synchronize()
            dt = _max_over_ranks(time.perf_counter() - t0, world, dist)
            tim.pop("device_result", None)
            runs.append((dt, tim, st, fail, float(np.nansum(res[:, 20, :])) if res is not None else None))
        runs.sort(key=lambda r: r[0])
        dt, tim, st, fail, chk = runs[1]
        flops = f_node_flops(st)
        entry = {"N": N, "value": M / dt, "unit": "MC sets/s", "seconds": dt,
                 "seconds_all": [r[0] for r in runs],
                 "rank0_compute_s": tim.get("compute_s"), "rank0_gather_s": tim.get("gather_s"),
                 "rank0_exchange_ms": tim.get("exchange_ms"),
                 "rank0_host_draw_s": st["host_draw_ns"] * 1e-9,
                 "failures": [int(fail[0]), int(fail[1])],
                 "rank0_node_evals": int(st["nodes"] + st["lz_nodes"]),
                 "rank0_f_node_tflops": flops / dt / 1e12,
                 "median_quantile_checksum": chk}
        per_n.append(entry)
        if N == 50 or headline is None:
            headline = entry
    out = {"metric": "resilience_mc_sets_per_s", "value": headline["value"], "unit": "MC sets/s",
           "n_gpus": world, "scaling": "strong", "higher_is_better": True,
           "config": {"workload": "C5: ONE experiment of M realisations per N, mixture "
                                  "(31,3.5,0.33,62,5.5,0.67), P=100 MW, 41 quantiles, default + "
                                  "flat prior, tol 1e-4, seed 848782",
                      "realisations": M, "headline_N": headline["N"], "rng_streams": RESIL_STREAMS,
                      "sharding": "rank r owns RNG streams r, r+G, ...; one all-gather of the "
                                  "[2][41][M_local] blocks (NCCL) inside the timed region; the "
                                  "gathered result is in HBM on every rank, rank 0 copies it to "
                                  "host memory (also inside the timed region)",
                      "timing": "median of 3 runs, host wall clock around the whole call "
                                "(barrier + device synchronise on both sides), max over ranks"},
           "per_N": per_n,
           # host draws in (U, Q0) and the quantile block out, per experiment
           "e2e": {"value": headline["value"], "unit": "MC sets/s",
                   "h2d_bytes_per_step": int(2 * 8 * M * headline["N"]),
                   "d2h_bytes_per_step": int(2 * 41 * 8 * M),
                   "note": "the RNG streams are host work by design (bit-exact libstdc++ "
                           "restatement), so every run is end to end: draws from host memory "
                           "in, [2][41][M] result array in host memory out"}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.resilience_cpu_sample > 0:
        from oracle import oracle
        threads = os.cpu_count() or 1
        Mc = args.resilience_cpu_sample
        t0 = time.perf_counter()
        oracle.lib().orc_set_num_threads(threads)
        oracle.resilience(1, MIX, headline["N"], Mc, 100.0, Q41, DEFAULT_PRIOR, 1.0, 1e-4, 848782,
                          nthread=threads, precision="long double")
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": Mc / dt, "unit": "MC sets/s", "cores": threads,
                               "kind": "port",
                               "sample": "%d realisations at N=%d, long double (the reference's "
                                         "arithmetic, resilience.cpp:321), %.1f s"
                                         % (Mc, headline["N"], dt)}
    return out


def make_gcp_case(K, seed=3):
    """100 regional sample sets (N_r ~ U[10,100], gamma) -> updated
    flat-prior tuples; K candidate priors in the SHGO bounds of prior.py:484-487."""
    rng = np.random.default_rng(seed)
    params = []
    for r in range(100):
        Nn = rng.integers(10, 101)
        a = np.exp(rng.uniform(np.log(10), np.log(200)))
        qq = rng.gamma(a, size=Nn) / (a / 68.3)
        params.append([np.log(qq).sum(), qq.sum(), Nn, Nn])
    refs = np.column_stack([rng.uniform(0.0, 4.0, K), rng.uniform(1.0, 50.0, K), np.zeros(K),
                            rng.uniform(0.05, 3.0, K)])
    refs[:, 2] = refs[:, 3] * (1.0 + np.exp(rng.uniform(-5, 1, K)))
    return np.array(params), refs


def bench_gcp(args, world, rank, local_rank, dist, barrier):
    """BASELINE.json config 4: minimum-surprise cost (max over 100 Kullback-Leibler distances,
    gamma_conjugate_prior.cpp:667-712) of K candidates x a_min in {0.5, 1, 2, 4}; candidates
    sharded over the ranks, cost vectors all-gathered.  Strong scaling."""
    K = args.gcp_candidates
    if K <= 0:
        return None
    import torch
    from reheatfunq_b200.sharding import kl_batch_max_sharded
    params, refs = make_gcp_case(K)
    amins = (0.5, 1.0, 2.0, 4.0)

    def sweep():
        return [kl_batch_max_sharded(params, refs, am, dist=dist, device=local_rank) for am in amins]
    sweep()
    runs = []
    for rep in range(5):
        barrier()
        t0 = time.perf_counter()
        cost = sweep()
        torch.cuda.synchronize()
        runs.append(_max_over_ranks(time.perf_counter() - t0, world, dist))
    dt = float(np.median(runs))
    out = {"metric": "gcp_cost_evals_per_s", "value": len(amins) * K / dt, "unit": "cost evals/s",
           "n_gpus": world, "scaling": "strong", "seconds": dt, "seconds_all": runs,
           "kl_integrals_per_s": len(amins) * K * params.shape[0] / dt,
           "finite_fraction": float(np.mean([np.isfinite(c).mean() for c in cost])),
           "config": {"workload": "C4: cost = max over 100 regional tuples of KL(tuple || "
                                  "candidate), %d candidates x a_min in {0.5,1,2,4}" % K,
                      "sharding": "contiguous candidate blocks per rank, all-gather of the costs",
                      "timing": "median of 5 sweeps, host buffers in and out"}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        threads = os.cpu_count() or 1
        oracle.lib().orc_set_num_threads(threads)
        Kc = 64
        t0 = time.perf_counter()
        ref = [oracle.gcp_kl_batch_max(params, refs[:Kc], am) for am in amins]
        dtc = time.perf_counter() - t0
        rel = 0.0
        for c, o in zip(cost, ref):
            m = np.isfinite(o)
            if m.any():
                rel = max(rel, float(np.max(np.abs(c[:Kc][m] / o[m] - 1.0))))
        out["cpu_baseline"] = {"value": len(amins) * Kc / dtc, "unit": "cost evals/s",
                               "cores": threads, "kind": "port",
                               "sample": "first %d candidates x 4 a_min, long double adaptive "
                                         "quadrature as gamma_conjugate_prior.cpp, %.1f s" % (Kc, dtc)}
        out["max_rel_vs_cpu_sample"] = rel
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--regions", type=int, default=10000)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--ref-sample", type=int, default=1536)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--resilience-m", type=int, default=1000000,
                    help="realisations of the ONE resilience experiment (0: skip)")
    ap.add_argument("--resilience-n", default="10,50,100")
    ap.add_argument("--resilience-cpu-sample", type=int, default=2000)
    ap.add_argument("--gcp-candidates", type=int, default=4096, help="0: skip the C4 block")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from reheatfunq_b200 import PosteriorBatch
    from reheatfunq_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback.")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    R = args.regions
    # weak scaling: every rank owns R regions of its own (seed offset by rank)
    q, c, off = make_regions(R, seed=1 + rank)
    w = np.ones(R)
    post_off = np.arange(R + 1, dtype=np.int64)
    nq = len(QUANTILES)
    # pinned host staging (the library copies from these buffers inside the timed region)
    tq = torch.from_numpy(q).pin_memory()
    tc = torch.from_numpy(c).pin_memory()
    qn, cn = tq.numpy(), tc.numpy()
    h2d_bytes = q.nbytes + c.nbytes + off.nbytes + w.nbytes + post_off.nbytes + nq * 8 * R
    d2h_bytes = nq * 8 * R

    def step_e2e():
        """The call a user makes: host buffers in, host results out."""
        B = PosteriorBatch.from_packed(qn, cn, off, w, post_off, DEFAULT_PRIOR, device=local_rank)
        out = B.tail_quantiles(QUANTILES)
        return B, out

    # device-resident copies for `value` (inputs already in HBM when the timed region starts)
    dev = torch.device("cuda", local_rank)
    dq, dc = tq.to(dev), tc.to(dev)
    doff = torch.from_numpy(off).to(dev)
    dw = torch.from_numpy(w).to(dev)
    dpost = torch.from_numpy(post_off).to(dev)
    dprior = torch.tensor(DEFAULT_PRIOR, dtype=torch.float64, device=dev)
    dt = torch.from_numpy(np.tile(QUANTILES, R)).to(dev)
    dtoff = torch.arange(R + 1, dtype=torch.int64, device=dev) * nq
    dout = torch.empty(R * nq, dtype=torch.float64, device=dev)

    def step_dev():
        B = PosteriorBatch.from_device(dq.data_ptr(), dc.data_ptr(), doff.data_ptr(),
                                       dw.data_ptr(), dpost.data_ptr(), R, dprior.data_ptr(),
                                       dw.numel(), dq.numel(), device=local_rank)
        B.query_device("tail_quantiles", dt.data_ptr(), dtoff.data_ptr(), dout.data_ptr(),
                       dout.numel())
        return B, dout

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        B, out = step_e2e()
        del B
        B, out = step_dev()
        del B
    # L2 flush between timed iterations: write a buffer larger than the 126 MB L2
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    sampler = ClockSampler(local_rank)
    if not os.environ.get('RHFQ_NO_SAMPLER'):
        sampler.start()

    def timed(step):
        """K steps; returns (mean ms per step by CUDA events on the default stream with a
        device synchronise on both sides, counters of the last step)."""
        ms, cnt = [], None
        barrier()
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            B, out = step()
            torch.cuda.synchronize()   # the library works on its own stream: full device sync
            e1.record()
            torch.cuda.synchronize()
            # host wall clock is the honest figure here: the step contains host-side
            # control flow (level counts) between launches
            ms.append(max((time.perf_counter() - t0) * 1e3, e0.elapsed_time(e1)))
            cnt = B.counters()
            del B
        barrier()
        print("[bench] rank %d %s steps [ms]: %s" % (rank, step.__name__,
              " ".join("%.1f" % v for v in ms)), file=sys.stderr)
        m = float(np.mean(ms))
        if world > 1:
            t = torch.tensor([m], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            m = float(t.item())
        return m, cnt

    ms_e2e, _ = timed(step_e2e)
    ms, cnt = timed(step_dev)
    # The timed steps carry CUDA events around the Simpson kernel only (the dominant kernel: one
    # launch per step).  Event pairs around the ~60 node / norm launches cost 4-6 us of device
    # idle time each, so those kernels are timed in three extra, instrumented steps after the
    # timed region (RHFQ_KERNEL_TIMERS=1, read when a batch is created).
    os.environ["RHFQ_KERNEL_TIMERS"] = "1"
    cnt_inst = None
    for _ in range(3):
        B, out = step_dev()
        torch.cuda.synchronize()
        cnt_inst = B.counters()
        del B
    del os.environ["RHFQ_KERNEL_TIMERS"]

    # second headline: resilience Monte Carlo (BASELINE.json config 5) -- ONE experiment of M
    # realisations sharded over the ranks, all-gather of the result inside the timed region
    resil = bench_resilience(args, world, rank, local_rank, dist if world > 1 else None, barrier)
    # third: GCP minimum-surprise cost sweep (BASELINE.json config 4), candidates sharded
    gcp = bench_gcp(args, world, rank, local_rank, dist if world > 1 else None, barrier)
    sampler.stop_flag = True
    if sampler.is_alive():
        sampler.join(timeout=2)

    evals = R * nq * world
    value = evals / (ms * 1e-3)
    e2e_value = evals / (ms_e2e * 1e-3)
    node_ms = cnt_inst["node_kernel_ns"] * 1e-6
    flops = f_node_flops(cnt)

    if rank == 0:
        peak = None
        err = (__import__("ctypes")).create_string_buffer(256)
        tf = (__import__("ctypes")).c_double()
        if _lib.api().measure_fp64_peak(local_rank, __import__("ctypes").byref(tf), err, 256) == 0:
            peak = tf.value
        src = ("measured here: register-resident DFMA chain (MEASURED_PEAKS.json has no "
               "FP64 entry)")

        # DRAM bytes per launch from the committed ncu pass over the same command
        # (dram__bytes_read.sum + dram__bytes_write.sum, tools/traffic_summary.py)
        traffic = {}
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles",
                             "r02_traffic.json")
        if os.path.exists(tpath) and R == 10000:
            traffic = json.load(open(tpath))

        def roof(kernel, kernel_ms, kflops, extra):
            ach = kflops / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else None
            tr = traffic.get(kernel)
            d = {"bound": "fp64", "kernel": kernel, "achieved": ach, "peak": peak,
                 "unit": "TFLOP/s", "frac": (ach / peak) if (ach and peak) else None,
                 "traffic": tr["bytes_per_launch"] if tr else None,
                 "traffic_source": tr["source"] if tr else None,
                 "peak_source": src, "algorithmic_flops_per_step": kflops,
                 "kernel_ms_per_step": kernel_ms,
                 "kernel_share_of_step": kernel_ms / ms if ms > 0 else None}
            d.update(extra)
            return d

        # Node evaluation, F_node of SURVEY 8d = 27 (N + 1) + 110 I + 130 K per node, now spread
        # over four kernels: the Gauss-Legendre kernels (rhfq_node_quad for pdf nodes,
        # rhfq_norm_quad for the norm's z-nodes) do 130 flop per evaluation of the alpha
        # log-integrand they make; the prep / mode / window kernels carry the rest.
        # Simpson-tree kernel: 82 flop per fine-grid evaluation (12-point Lagrange: 12 sub,
        # 46 mul, 12 fma), 4 per Chebyshev term of a direct (fallback) evaluation, 66 per
        # evaluation for log / asin / exp / argument mapping, 24 per accept test (the
        # division-free form: 18 add/mul/compare + the leaf mass).
        saq_ms = cnt["saq_kernel_ns"] * 1e-6
        saq_flops = (82.0 * cnt["fine_evals"] + 4.0 * cnt["cheb_terms"]
                     + (2 * 66.0 + 24.0) * cnt["saq_steps"])
        quad_ms = (cnt_inst["node_kernel_ns"] + cnt_inst["norm_kernel_ns"]) * 1e-6
        quad_flops = 130.0 * cnt["g_evals_quad"]
        plan_ms = cnt_inst["plan_kernel_ns"] * 1e-6
        plan_flops = flops - quad_flops
        node_ms = quad_ms
        r_node = roof("rhfq_node_quad + rhfq_norm_quad", quad_ms, quad_flops,
                      {"timed_in": "3 instrumented steps after the timed region (RHFQ_KERNEL_TIMERS=1)",
                       "node_evals": int(cnt["nodes"] + cnt["lz_nodes"]),
                       "log_integrand_evals": int(cnt["g_evals_quad"]),
                       "planning_kernels": {"kernels": "rhfq_{node,norm}_{prep,mode,window}",
                                            "kernel_ms_per_step": plan_ms,
                                            "algorithmic_flops_per_step": plan_flops,
                                            "achieved": (plan_flops / (plan_ms * 1e-3) / 1e12)
                                            if plan_ms > 0 else None}})
        saq_kernel = "rhfq_saq_step_bli" if os.environ.get("RHFQ_SAQ_LEVELSYNC") else "rhfq_saq_post"
        r_saq = roof(saq_kernel, saq_ms, saq_flops,
                     {"intervals": int(cnt["saq_steps"]), "fine_grid_evals": int(cnt["fine_evals"]),
                      "chebyshev_terms": int(cnt["cheb_terms"]),
                      "fine_grids_rejected": int(cnt["fine_bad"])})
        dominant, other = (r_saq, r_node) if saq_ms > node_ms else (r_node, r_saq)
        line = {
            "metric": "posterior_PH_evals_per_s", "value": value, "unit": "P_H evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3 batched regions: tail quantiles [0.1,0.5,0.9] of R "
                                   "independent posteriors, N in [10,100], default prior, "
                                   "rtol 1e-8, BLI r<=7", "regions_per_gpu": R, "quantiles": 3,
                       "l2": "256 MiB buffer written between timed iterations",
                       "value": "inputs and outputs resident in HBM",
                       "e2e": "pinned host buffers through PosteriorBatch.from_packed + "
                              "tail_quantiles (H2D and D2H inside the timed region)"},
            "posteriors_per_s": R * world / (ms * 1e-3),
            "resilience": resil,
            "gcp": gcp,
            "gpu_launches": int(cnt["launches"]),
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "P_H evals/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes)},
            "roofline": dominant,
            "roofline_other": other,
        }
        if not args.no_cpu_baseline:
            import __graft_entry__ as g
            g.build_oracle()
            threads = os.cpu_count() or 1
            sample = max(threads, args.ref_sample)
            v, dt = cpu_reference_rate(q[:off[sample]], c[:off[sample]], off[:sample + 1],
                                       sample, threads)
            line["cpu_baseline"] = {"value": v, "unit": "P_H evals/s", "cores": threads,
                                    "kind": "port",
                                    "sample": "first %d regions of the same workload, %.1f s"
                                              % (sample, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
