"""Multi-GPU plumbing of the pairwise analysis: one process per GPU (torch.distributed).

The path shards twice and exchanges once:
  * null replicates: contiguous ranges of the outer replicates (nb_rep_CPU); simulated site
    ids are global, so the union of the shards is bit-identical to a single-GPU run;
  * one exchange: all-gather of the unbinned null samples (Stat, Nmin) -- 16 MB at
    1000 x 1000 -- after which every rank bins and sorts the union itself;
  * pair rows: row i goes to rank r when i mod 2N is r or 2N-1-r, which pairs a long row
    with a short one (row i holds S-1-i pairs).
The product path does all of this inside the library (cmb_comm_* / cmb_null_intra_sharded: NCCL
on the context's stream).  This module states the same sharding rules for the launcher and the
CPU (gloo) tests of the host logic, plus an exchange over torch.distributed for callers that
bring their own plumbing; nothing here touches the device.
"""
import numpy as np


def replicate_bounds(rep_cpu, world):
    """[r0, r1) of every rank: contiguous, the first rep_cpu % world ranks hold one more
    (the rule cmb_null_intra_sharded applies inside the library)."""
    q, r = divmod(int(rep_cpu), int(world))
    out, b = [], 0
    for k in range(world):
        e = b + q + (1 if k < r else 0)
        out.append((b, e))
        b = e
    return out


def owned_rows(n_sites, rank, world):
    """Rows of the upper triangle scored by `rank` (same rule as cmb_pairs' shard arguments)."""
    i = np.arange(n_sites)
    if world == 1:
        return i
    m = i % (2 * world)
    return i[(m == rank) | (m == 2 * world - 1 - rank)]


def owned_pairs(n_sites, rank, world):
    rows = owned_rows(n_sites, rank, world)
    return int((n_sites - 1 - rows).sum())


def all_gather_null(stat, nmin, n_valid, max_per_rank, group=None):
    """All-gathers each rank's (stat, nmin) samples.

    stat / nmin: 1-D float64 tensors holding n_valid samples (device of the backend).
    Returns (stat_all, nmin_all) of length world * max_per_rank; slots a rank did not fill
    hold NaN in both arrays, which the binning drops exactly like out-of-domain samples
    (Domain.getIndex throws for them, AnalysisTools.cpp:645-648).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    send = torch.full((2, max_per_rank), float("nan"), dtype=torch.float64, device=stat.device)
    send[0, :n_valid] = stat[:n_valid]
    send[1, :n_valid] = nmin[:n_valid]
    out = torch.empty((world, 2, max_per_rank), dtype=torch.float64, device=stat.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out.view(-1), send.view(-1), group=group)
    else:
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send, group=group)
        for r in range(world):
            out[r] = parts[r]
    return out[:, 0, :].reshape(-1).contiguous(), out[:, 1, :].reshape(-1).contiguous()


def max_over_ranks(value, device, group=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
