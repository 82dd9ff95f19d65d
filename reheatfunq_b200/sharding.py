"""
Multi-GPU sharding of the hot path: one process per GPU, work items (regions,
realisations, parameter tuples) are independent, so ranks take contiguous
blocks with no data-path collective; the only communication is the gather of
the fixed-size result block at the end (NCCL over NVLink on the GPU box, gloo
in the CPU tests).  The reference has no counterpart (OpenMP only,
src/resilience/resilience.cpp:405-421).
"""
import numpy as np


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
                       nstreams=1, chunk=0, dist=None, device=0, api=None):
    """
    One resilience experiment of M realisations split over the ranks of `dist`.
    Every rank replays the (cheap, host-side) RNG streams and computes its own
    block of realisations; the result blocks are all-gathered.  The result does
    not depend on the number of ranks.
    """
    from .resilience import resilience_run
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    lo, hi = block_range(M, rank, world)
    out, failures, stats = resilience_run(dist_kind, dist_params, N, M, P_MW, quantile, prior,
                                          amin, tol, seed, nstreams=nstreams, device=device,
                                          chunk=chunk, realisations=(lo, hi), api=api)
    gdev = None
    if dist is not None and dist.is_initialized() and dist.get_backend() == "nccl":
        import torch
        gdev = torch.device("cuda", device)
    full = gather_columns(out, lo, hi, M, dist, gdev)
    if world > 1:
        import torch
        f = torch.tensor(failures, dtype=torch.int64, device=gdev)
        dist.all_reduce(f)
        failures = f.cpu().numpy()
    return full, failures, stats
