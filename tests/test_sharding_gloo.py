"""
N > 1 host logic on CPU: two gloo ranks shard one resilience experiment (each
on the test-only host emulation of the engine) and all-gather the result; the
outcome must equal the single-rank run.
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, DEFAULT_PRIOR
from reheatfunq_b200.sharding import block_range


def test_block_range():
    for n in (0, 1, 7, 10, 1000003):
        for world in (1, 2, 3, 8):
            blocks = [block_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for (a, b), (c, d) in zip(blocks[:-1], blocks[1:]):
                assert b == c
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _gcp_case():
    rng = np.random.default_rng(3)
    params = []
    for r in range(7):
        Nn = rng.integers(10, 101)
        a = np.exp(rng.uniform(np.log(10), np.log(200)))
        qq = rng.gamma(a, size=Nn) / (a / 68.3)
        params.append([np.log(qq).sum(), qq.sum(), Nn, Nn])
    K = 11
    refs = np.column_stack([rng.uniform(0.0, 4.0, K), rng.uniform(1.0, 50.0, K), np.zeros(K),
                            rng.uniform(0.05, 3.0, K)])
    refs[:, 2] = refs[:, 3] * (1.0 + np.exp(rng.uniform(-5, 1, K)))
    return np.array(params), refs


def _worker(rank, world, port, out_dir):
    import ctypes as C
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from reheatfunq_b200._lib import Api
    from reheatfunq_b200.sharding import resilience_sharded
    emu = Api(C.CDLL(os.path.join(ROOT, "tests", "hostmath", "libengine_emu.so")), "emu_rhfq_")
    mix = (31, 3.5, 0.33, 62, 5.5, 0.67)
    q = np.array([0.9, 0.5, 0.1, 0.01])
    full, failures, _ = resilience_sharded(1, mix, 12, 21, 100.0, q, DEFAULT_PRIOR, 1.0, 1e-4, 5,
                                           nstreams=2, dist=dist, api=emu)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), full)
    # result for rank 0 only, through the host buffer all ranks map (each writes its slice);
    # a second, larger experiment makes the buffer grow
    for tag, M in (("a", 21), ("b", 37)):
        res0, fail0, _ = resilience_sharded(1, mix, 12, M, 100.0, q, DEFAULT_PRIOR, 1.0, 1e-4, 5,
                                            nstreams=2, dist=dist, api=emu, host_result="rank0")
        assert (res0 is None) == (rank != 0)
        if rank == 0:
            np.save(os.path.join(out_dir, "rank0_shared_%s.npy" % tag), np.array(res0))
            np.save(os.path.join(out_dir, "rank0_shared_fail_%s.npy" % tag), fail0)
    # GCP cost sweep sharded both ways
    from reheatfunq_b200.sharding import kl_batch_max_sharded
    params, refs = _gcp_case()
    for how in ("candidates", "tuples"):
        cost = kl_batch_max_sharded(params, refs, 1.0, shard=how, dist=dist, api=emu)
        np.save(os.path.join(out_dir, "kl_%s_rank%d.npy" % (how, rank)), cost)
    dist.destroy_process_group()


def test_two_rank_resilience(tmp_path):
    import ctypes as C
    import torch.multiprocessing as mp
    from reheatfunq_b200._lib import Api
    from reheatfunq_b200.resilience import resilience_run
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    np.testing.assert_array_equal(r0, r1)
    np.testing.assert_array_equal(np.load(tmp_path / "rank0_shared_a.npy"), r0)
    emu = Api(C.CDLL(os.path.join(ROOT, "tests", "hostmath", "libengine_emu.so")), "emu_rhfq_")
    mix = (31, 3.5, 0.33, 62, 5.5, 0.67)
    q = np.array([0.9, 0.5, 0.1, 0.01])
    single, _, _ = resilience_run(1, mix, 12, 21, 100.0, q, DEFAULT_PRIOR, 1.0, 1e-4, 5,
                                  nstreams=2, api=emu)
    np.testing.assert_array_equal(r0, single)
    single_b, fail_b, _ = resilience_run(1, mix, 12, 37, 100.0, q, DEFAULT_PRIOR, 1.0, 1e-4, 5,
                                         nstreams=2, api=emu)
    np.testing.assert_array_equal(np.load(tmp_path / "rank0_shared_b.npy"), single_b)
    np.testing.assert_array_equal(np.load(tmp_path / "rank0_shared_fail_b.npy"),
                                  np.asarray(fail_b, dtype=np.int64))
    # the sharded GCP cost equals the single-rank cost, whichever way it is split
    from reheatfunq_b200.regional.backend import \
        gamma_conjugate_prior_kullback_leibler_batch_max_many as many
    params, refs = _gcp_case()
    ref_cost = many(params, refs, 1.0, _api=emu)
    for how in ("candidates", "tuples"):
        for rank in (0, 1):
            np.testing.assert_array_equal(np.load(tmp_path / ("kl_%s_rank%d.npy" % (how, rank))),
                                          ref_cost)
