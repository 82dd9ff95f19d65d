"""
Multi-GPU sharding of the hot path: one process per GPU.  Work items (regions,
realisations, parameter tuples) are independent, so there is no data-path
collective inside the computation; the exchange step is the gather of the
result blocks at the end (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  The reference has no counterpart (OpenMP only,
src/resilience/resilience.cpp:405-421, gamma_conjugate_prior.cpp:667-712).
"""
import numpy as np


_PINNED = {}      # page-locked result buffers of resilience_sharded, by shape
_SHARED = {}      # host buffer mapped by all ranks of the node (see _shared_host)


class _SharedHost:
    """
    A host buffer that every rank of the node maps (POSIX shared memory) and page-locks in its
    own CUDA context, so that each GPU writes its slice of a gathered result over ITS OWN PCIe
    link: 8 GPUs move the 656 MB of a 10^6-realisation result in 1/8 of the time one link
    needs (26 GB/s measured).  Rank 0 creates and unlinks the segment.
    """

    def __init__(self, dist, nbytes, cuda):
        import atexit
        import torch
        from multiprocessing import shared_memory, resource_tracker
        self.rank = dist.get_rank()
        self.nbytes = int(nbytes)
        self.shm = None
        self.registered = False
        name = [None]
        if self.rank == 0:
            try:
                import os
                st = os.statvfs("/dev/shm")
                if st.f_bavail * st.f_frsize > 1.25 * nbytes:
                    self.shm = shared_memory.SharedMemory(create=True, size=self.nbytes)
                    name[0] = self.shm.name
            except Exception:
                self.shm = None
        dist.broadcast_object_list(name, src=0)
        ok = name[0] is not None
        if ok and self.rank != 0:
            try:
                self.shm = shared_memory.SharedMemory(name=name[0])
                # the creator unlinks; the attaching processes must not (Python < 3.13 registers
                # every attachment with its resource tracker)
                try:
                    resource_tracker.unregister(self.shm._name, "shared_memory")
                except Exception:
                    pass
            except Exception:
                ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32,
                            device="cuda" if cuda else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self.ok = bool(int(flag.item()))
        self.flat = None
        if self.ok:
            self.flat = torch.frombuffer(self.shm.buf, dtype=torch.float64, count=self.nbytes // 8)
            if cuda:
                rc = torch.cuda.cudart().cudaHostRegister(self.flat.data_ptr(), self.nbytes, 0)
                self.registered = int(rc) == 0
        atexit.register(self.close)

    def close(self):
        if self.shm is None:
            return
        try:
            if self.registered:
                import torch
                torch.cuda.cudart().cudaHostUnregister(self.flat.data_ptr())
        except Exception:
            pass
        self.registered = False
        self.flat = None
        shm, self.shm = self.shm, None
        try:
            shm.close()
        except Exception:
            pass
        if self.rank == 0:
            try:
                shm.unlink()
            except Exception:
                pass


def _shared_host(dist, n_doubles, cuda):
    """The node-wide result buffer, grown (collectively: every rank calls with the same size)."""
    cur = _SHARED.get("buf")
    if cur is not None and (not cur.ok or cur.nbytes >= 8 * n_doubles):
        return cur
    if cur is not None:
        cur.close()
    cur = _SHARED["buf"] = _SharedHost(dist, 8 * int(n_doubles), cuda)
    return cur


def block_range(n, rank, world):
    """Contiguous block [lo, hi) of n items owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def gather_columns(local, lo, hi, total, dist=None, device=None):
    """
    All-gathers column blocks: `local[..., lo:hi]` of every rank -> full `[..., total]`
    array on every rank.  `dist` is torch.distributed (initialised) or None (single rank).
    """
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch
    world = dist.get_world_size()
    lead = local.shape[:-1]
    rows = int(np.prod(lead)) if lead else 1
    width = max(block_range(total, r, world)[1] - block_range(total, r, world)[0]
                for r in range(world))
    send = torch.zeros((rows, width), dtype=torch.float64, device=device)
    blk = np.ascontiguousarray(local.reshape(rows, total)[:, lo:hi])
    send[:, :hi - lo] = torch.from_numpy(blk).to(send.device)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    out = np.empty((rows, total))
    for r in range(world):
        a, b = block_range(total, r, world)
        out[:, a:b] = recv[r][:, :b - a].cpu().numpy()
    return out.reshape(lead + (total,))


def resilience_sharded(dist_kind, dist_params, N, M, P_MW, quantile, prior, amin, tol, seed,
                       nstreams=64, chunk=0, dist=None, device=0, api=None, timings=None,
                       host_result="all"):
    """
    ONE resilience experiment of M realisations (src/resilience/resilience.cpp:375-508) split
    over the ranks of `dist`.  Rank r owns the RNG streams t = r, r + G, ... of the `nstreams`
    streams (whole streams: no rank replays another rank's draws; the reference deals the same
    batches of 10 out to its OpenMP threads, :397,420-437), computes its realisations on its
    GPU and leaves the [2][Nq][M_local] block in HBM; the blocks are exchanged by ONE
    all-gather (NCCL over NVLink on the GPU box, gloo in the CPU tests) and interleaved by
    column into the [2][Nq][M] result.  The result is a function of (seed, nstreams) alone --
    bit-identical for any number of ranks.

    ``host_result``: "all" -- every rank returns the result as a host array (8 ranks copying
    656 MB each to host memory at once share the host's memory bandwidth); "rank0" -- the
    gathered result stays in HBM on every rank (``timings["device_result"]``, a torch tensor)
    and only rank 0, the process a drop-in caller lives in, gets it as a host array; the other
    ranks return None.  Every rank writes 1/G of the gathered result into a host buffer all
    ranks map (`_SharedHost`: one PCIe link per GPU instead of rank 0's link for everything;
    ranks on one node -- otherwise, or without room in /dev/shm, rank 0 copies it alone).
    """
    from .resilience.zeal2022hfresil import resilience_run, shard_index
    import time
    on = dist is not None and dist.is_initialized()
    rank = dist.get_rank() if on else 0
    world = dist.get_world_size() if on else 1
    quantile = np.ascontiguousarray(quantile, dtype=np.float64)
    Nq = quantile.shape[0]
    t0 = time.perf_counter()
    if world == 1:
        out, failures, stats = resilience_run(dist_kind, dist_params, N, M, P_MW, quantile, prior,
                                              amin, tol, seed, nstreams=nstreams, device=device,
                                              chunk=chunk, api=api)
        if timings is not None:
            timings.update(compute_s=time.perf_counter() - t0, gather_s=0.0)
        return out, failures, stats
    import torch
    nccl = dist.get_backend() == "nccl"
    gdev = torch.device("cuda", device) if nccl else torch.device("cpu")
    # realisation -> column maps of all ranks: a function of (M, nstreams, world) alone, kept on
    # the device across calls
    key = ("index", int(M), int(nstreams), int(world), str(gdev))
    cached = _PINNED.get(key)
    if cached is None:
        index = [shard_index(M, nstreams, (r, world), api=api) for r in range(world)]
        cached = _PINNED[key] = (index, [torch.from_numpy(ix).to(gdev) for ix in index])
        for k in [k for k in _PINNED if isinstance(k, tuple) and k[0] == "index" and k != key]:
            del _PINNED[k]
    index, index_dev = cached
    width = max(len(ix) for ix in index)
    n_loc = len(index[rank])
    send = torch.zeros((2 * Nq, width), dtype=torch.float64, device=gdev)
    if nccl:
        local = torch.empty((2 * Nq, max(n_loc, 1)), dtype=torch.float64, device=gdev)
        _, failures, stats = resilience_run(dist_kind, dist_params, N, M, P_MW, quantile, prior,
                                            amin, tol, seed, nstreams=nstreams, device=device,
                                            chunk=chunk, shard=(rank, world),
                                            out_device_ptr=local.data_ptr(), api=api)
        torch.cuda.synchronize(gdev)
        send[:, :n_loc] = local[:, :n_loc]
    else:
        out, failures, stats = resilience_run(dist_kind, dist_params, N, M, P_MW, quantile, prior,
                                              amin, tol, seed, nstreams=nstreams, device=device,
                                              chunk=chunk, shard=(rank, world), api=api)
        send[:, :n_loc] = torch.from_numpy(out.reshape(2 * Nq, n_loc))
    t1 = time.perf_counter()
    marks = []                # (label, CUDA event) of the exchange's phases, resolved at the end

    def mark(label):
        if nccl and timings is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((label, e))

    mark("start")
    recv = torch.empty((world, 2 * Nq, width), dtype=torch.float64, device=gdev)
    dist.all_gather_into_tensor(recv, send) if nccl else dist.all_gather(list(recv.unbind(0)), send)
    mark("all_gather")
    full = torch.empty((2 * Nq, M), dtype=torch.float64, device=gdev)
    for r in range(world):
        ix = index_dev[r]
        full.index_copy_(1, ix, recv[r][:, :len(index[r])])
    mark("interleave")
    f = torch.tensor(np.asarray(failures, dtype=np.int64), device=gdev)
    if timings is not None:
        timings["device_result"] = full
    shared = _shared_host(dist, 2 * Nq * M, nccl) if host_result == "rank0" else None
    if shared is not None and shared.ok:
        a, b = block_range(2 * Nq * M, rank, world)
        shared.flat[a:b].copy_(full.view(-1)[a:b], non_blocking=True)
        mark("to_host")
        # the reduction of the failure counts is stream-ordered behind every rank's copy:
        # when it has arrived on rank 0 all slices are in host memory
        dist.all_reduce(f)
        failures_all = f.cpu().numpy()
        mark("all_reduce")
        result = shared.flat[:2 * Nq * M].numpy().reshape(2, Nq, M) if rank == 0 else None
        if timings is not None:
            timings.update(compute_s=t1 - t0, gather_s=time.perf_counter() - t1)
            if marks:
                torch.cuda.synchronize(gdev)
                timings["exchange_ms"] = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks, marks[1:])}
        return result, failures_all, stats
    dist.all_reduce(f)
    if host_result == "rank0" and rank != 0:
        result = None
        if nccl:
            torch.cuda.synchronize(gdev)
    elif nccl:
        # device -> host through ONE page-locked buffer that only ever grows (a pageable
        # destination copies at 2-3 GB/s: 0.25 s for the 656 MB of a 10^6-realisation result;
        # page-locking a fresh buffer of that size costs 0.7 s)
        need = 2 * Nq * M
        flat = _PINNED.get("flat")
        if flat is None or flat.numel() < need:
            flat = _PINNED["flat"] = torch.empty(need, dtype=torch.float64, pin_memory=True)
        host = flat[:need].view(2 * Nq, M)
        host.copy_(full, non_blocking=True)
        torch.cuda.synchronize(gdev)
        result = host.numpy().reshape(2, Nq, M)      # view of the pinned buffer: valid until the next call
    else:
        result = full.numpy().reshape(2, Nq, M)
    if timings is not None:
        timings.update(compute_s=t1 - t0, gather_s=time.perf_counter() - t1)
    return result, f.cpu().numpy(), stats


def kl_batch_max_sharded(params, refs, amin=1.0, shard="candidates", dist=None, device=0,
                         api=None):
    """
    Minimum-surprise cost of K candidate priors over N regional sample tuples,
    out[k] = max_i KL(params_i || refs_k) (gamma_conjugate_prior.cpp:667-712), split over the
    ranks of `dist`:

      shard="candidates"  each rank evaluates a contiguous block of the K candidates against
                          all N tuples; the cost blocks are all-gathered (an SHGO generation
                          or an a_min sweep spread over the GPUs);
      shard="tuples"      each rank evaluates all candidates against its block of the N
                          regional tuples; the partial maxima are combined by all-reduce(MAX)
                          (many regions, few candidates).

    Returns the full [K] cost vector on every rank, identical to the single-rank call.
    """
    from .regional.backend import gamma_conjugate_prior_kullback_leibler_batch_max_many as many
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 4)
    refs = np.ascontiguousarray(refs, dtype=np.float64).reshape(-1, 4)
    on = dist is not None and dist.is_initialized()
    rank = dist.get_rank() if on else 0
    world = dist.get_world_size() if on else 1
    if world == 1:
        return many(params, refs, amin, device=device, _api=api)
    import torch
    gdev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
    K, N = refs.shape[0], params.shape[0]
    if shard == "candidates":
        lo, hi = block_range(K, rank, world)
        width = max(block_range(K, r, world)[1] - block_range(K, r, world)[0] for r in range(world))
        send = torch.zeros(width, dtype=torch.float64, device=gdev)
        if hi > lo:
            send[:hi - lo] = torch.from_numpy(many(params, refs[lo:hi], amin, device=device,
                                                   _api=api)).to(gdev)
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(recv, send)
        out = np.empty(K)
        for r in range(world):
            a, b = block_range(K, r, world)
            out[a:b] = recv[r][:b - a].cpu().numpy()
        return out
    if shard == "tuples":
        lo, hi = block_range(N, rank, world)
        part = np.full(K, -np.inf)
        if hi > lo:
            part = many(params[lo:hi], refs, amin, device=device, _api=api)
        t = torch.from_numpy(np.ascontiguousarray(part)).to(gdev)
        # +inf (a failed normalisation, :698-704) must win over every finite value: MAX does that;
        # NaN cannot occur (the kernel maps failures to +inf)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu().numpy()
    raise ValueError("shard must be 'candidates' or 'tuples'")
