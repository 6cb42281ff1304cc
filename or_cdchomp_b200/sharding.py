"""Multi-GPU plumbing: static sharding of independent runs and the final best-cost gather.

Runs are independent (they share only read-only data: robot, SDFs, metric), so a
batch of R runs is split into contiguous shards, one per rank / GPU, the SDFs are
replicated, and nothing is exchanged during the iterations.  The only collective
is the arg-min over the final cost_total, done once per `iterate` call:
an all_gather of one (cost, global run id) pair per rank followed by a broadcast
of the winning trajectory from its owner (NCCL over NVLink on GPUs, gloo in the
CPU tests).  The reference has no counterpart: its module advances runs serially
(README.md:86-87 of the reference).
"""
import torch
import torch.distributed as dist


def shard_bounds(n_runs, rank, world):
    """Contiguous shard [lo, hi) of `n_runs` for `rank`; sizes differ by at most 1."""
    base, extra = divmod(int(n_runs), int(world))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def local_best(costs, status, lo):
    """(cost, global id) of the best successful local run; (+inf, -1) if none.
    First run wins ties, NaN costs are skipped (matches ocb_batch_best)."""
    best_c, best_i = float("inf"), -1
    for r in range(len(costs)):
        c = float(costs[r])
        if int(status[r]) != 0 or c != c:
            continue
        if c < best_c:
            best_c, best_i = c, lo + r
    return best_c, best_i


def gather_best(best_cost, best_global_id, traj_of_best, n_points, n_dof, device):
    """All ranks learn (cost, global run id, trajectory) of the globally best run.

    best_cost / best_global_id: this rank's candidate (id -1 = none).
    traj_of_best: tensor [n_points, n_dof] (float64) on `device` holding this
    rank's candidate trajectory (contents ignored when id == -1).
    Ties are broken towards the lowest global run id, so the answer does not
    depend on the number of ranks.
    """
    world = dist.get_world_size() if dist.is_initialized() else 1
    pair = torch.tensor([best_cost, float(best_global_id)], dtype=torch.float64, device=device)
    if world == 1:
        return best_cost, best_global_id, traj_of_best
    pairs = [torch.empty_like(pair) for _ in range(world)]
    dist.all_gather(pairs, pair)
    table = torch.stack(pairs).cpu()
    win_rank, win_cost, win_id = -1, float("inf"), -1
    for rk in range(world):
        c, i = float(table[rk, 0]), int(table[rk, 1])
        if i < 0:
            continue
        if c < win_cost or (c == win_cost and i < win_id):
            win_rank, win_cost, win_id = rk, c, i
    out = torch.empty((n_points, n_dof), dtype=torch.float64, device=device)
    if win_rank < 0:
        out.zero_()
        return win_cost, -1, out
    if dist.get_rank() == win_rank:
        out.copy_(traj_of_best)
    dist.broadcast(out, src=win_rank)
    return win_cost, win_id, out
