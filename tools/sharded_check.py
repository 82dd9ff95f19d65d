"""Multi-GPU check (run under torchrun, one rank per GPU): ONE resilience experiment sharded by
RNG stream over the ranks with the NCCL all-gather must be bit-identical to the same experiment
on one GPU, on every rank; likewise the GCP cost sweep sharded over candidates / tuples.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29533 tools/sharded_check.py [M] [N]
"""
import os, sys, json, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist
import bench
from reheatfunq_b200.sharding import resilience_sharded, kl_batch_max_sharded
from reheatfunq_b200.resilience import resilience_run
from reheatfunq_b200.regional.backend import gamma_conjugate_prior_kullback_leibler_batch_max_many as many

M = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tim = {}
t0 = time.perf_counter()
full, fail, st = resilience_sharded(1, bench.MIX, N, M, 100.0, bench.Q41, bench.DEFAULT_PRIOR, 1.0, 1e-4,
                                    848782, nstreams=bench.RESIL_STREAMS, dist=dist, device=local, timings=tim)
t_sh = time.perf_counter() - t0
tim.pop("device_result", None)
report = dict(rank=rank, world=world, M=M, N=N, sharded_s=t_sh, **tim, host_draw_s=st["host_draw_ns"] * 1e-9)
# every rank recomputes the whole experiment alone on its GPU and compares
t0 = time.perf_counter()
single, fail1, _ = resilience_run(1, bench.MIX, N, M, 100.0, bench.Q41, bench.DEFAULT_PRIOR, 1.0, 1e-4,
                                  848782, nstreams=bench.RESIL_STREAMS, device=local)
report["single_s"] = time.perf_counter() - t0
report["resilience_array_equal"] = bool(np.array_equal(full, single, equal_nan=True))
report["failures_equal"] = bool(np.array_equal(fail, fail1))
report["finite_fraction"] = float(np.isfinite(full).mean())
# the same experiment with the result for rank 0 only (every rank writes its slice of the
# gathered result into the host buffer all ranks map): twice, the second call is warm
from reheatfunq_b200 import sharding
for rep in range(2):
    tim0 = {}
    t0 = time.perf_counter()
    r0, f0, _ = resilience_sharded(1, bench.MIX, N, M, 100.0, bench.Q41, bench.DEFAULT_PRIOR, 1.0, 1e-4,
                                   848782, nstreams=bench.RESIL_STREAMS, dist=dist, device=local,
                                   timings=tim0, host_result="rank0")
    report["rank0_s_%d" % rep] = time.perf_counter() - t0
    report["rank0_gather_s_%d" % rep] = tim0["gather_s"]
sh = sharding._SHARED.get("buf")
report["shared_host"] = dict(ok=bool(sh and sh.ok), registered=bool(sh and sh.registered))
report["rank0_equal"] = bool(np.array_equal(r0, single, equal_nan=True)) if rank == 0 else (r0 is None)
report["rank0_failures_equal"] = bool(np.array_equal(f0, fail1))
params, refs = bench.make_gcp_case(1024)
ref_cost = many(params, refs, 1.0, device=local)
for how in ("candidates", "tuples"):
    c = kl_batch_max_sharded(params, refs, 1.0, shard=how, dist=dist, device=local)
    report["gcp_%s_equal" % how] = bool(np.array_equal(c, ref_cost))
print("SHARDED_CHECK " + json.dumps(report), flush=True)
ok = report["resilience_array_equal"] and report["failures_equal"] and report["gcp_candidates_equal"] \
    and report["gcp_tuples_equal"] and report["rank0_equal"] and report["rank0_failures_equal"]
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
