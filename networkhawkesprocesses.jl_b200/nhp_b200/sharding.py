"""Time sharding of the continuous event stream across ranks: contiguous shards of
equal event count, each preceded by the read-only halo of predecessors inside the look-back horizon.
Host-side logic only (numpy + torch.distributed for the exchange); the kernels see a shard through
nhp_events_upload(n_halo, index_base, flags)."""
import numpy as np


def shard_bounds(n, world):
    """[a_r, b_r) of rank r: equal event counts, remainder spread over the first ranks."""
    base, rem = divmod(int(n), int(world))
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


def halo_start(times, a, horizon):
    """First index j with times[j] > times[a] - horizon (the window start of the shard's first event)."""
    if a == 0:
        return 0
    if not np.isfinite(horizon):
        return 0
    return int(np.searchsorted(times, times[a] - horizon, side="right"))


def make_shard(times, nodes, rank, world, horizon):
    """Slice of the global stream rank `rank` uploads: dict(times, nodes, n_halo, index_base, flags, a, b)."""
    bounds = shard_bounds(len(times), world)
    a, b = bounds[rank], bounds[rank + 1]
    lo = min(halo_start(times, a, horizon), a) if b > a else a
    return dict(times=np.ascontiguousarray(times[lo:b]), nodes=np.ascontiguousarray(nodes[lo:b]), n_halo=a - lo, index_base=lo,
                flags=1 if rank == 0 else 0, a=a, b=b)


def allreduce_(array, dist=None):
    """In-place sum over ranks of a numpy array (gloo) or torch tensor (nccl); no-op without a process group."""
    import torch
    import torch.distributed as td
    dist = dist or td
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return array
    if isinstance(array, np.ndarray):
        t = torch.from_numpy(array)
        dist.all_reduce(t)
        return array
    dist.all_reduce(array)
    return array


def sharded_statistics(local, dist=None):
    """Two-phase reduction protocol of the Gibbs statistics.
    `local` provides phase0() -> flat f64 array [ll_log, ll_row, M0, Mn, Mnm, S1] (reduced in place),
    second_pass(reduced_phase0) -> flat S2 array.  Returns (phase0, S2) summed over ranks."""
    p0 = allreduce_(local.phase0(), dist)
    s2 = allreduce_(local.second_pass(p0), dist)
    return p0, s2
