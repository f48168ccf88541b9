"""Multi-GPU sharding of the candidate axis.

The path shards embarrassingly: every rank holds a replica of the model state and owns a contiguous
slice of the candidates (or q-batches, or L-BFGS starts).  The only exchange is ONE all-gather of a
(value, global index, x[D]) tuple per rank; every rank then selects the same winner with the
reference's tie rule (first maximal index, optimizer.py:149).  One process per GPU,
``torch.distributed`` (NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of ``total`` items for ``rank`` (first ``total % world``
    ranks get one extra)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if total < 0:
        raise ValueError("total must be non-negative")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_best(pairs) -> Tuple[float, int]:
    """First-max merge of (value, global_index) pairs: larger value wins, ties -> lower index.
    NaN values never win; an empty shard is encoded as (-inf, -1)."""
    bv, bi = -np.inf, -1
    for v, i in pairs:
        if i < 0 or v != v:
            continue
        if bi < 0 or v > bv or (v == bv and i < bi):
            bv, bi = float(v), int(i)
    return bv, bi


def allgather_best(value: float, global_index: int, point: Optional[np.ndarray] = None, group=None, device=None):
    """The path's single collective.  Returns (best_value, best_global_index, best_point or None),
    identical on every rank."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value), int(global_index), None if point is None else np.asarray(point, dtype=np.float64)
    world = dist.get_world_size(group)
    D = 0 if point is None else int(np.asarray(point).shape[-1])
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    payload = torch.zeros(2 + D, dtype=torch.float64, device=device)
    payload[0] = float(value) if value == value else float("-inf")
    payload[1] = float(global_index)  # exact for indices < 2^53
    if D:
        payload[2:] = torch.as_tensor(np.asarray(point, dtype=np.float64).reshape(-1), device=device)
    gathered = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload, group=group)
    rows = [g.cpu().numpy() for g in gathered]
    bv, bi = merge_best([(r[0], int(r[1])) for r in rows])
    bp = None
    if D:
        for r in rows:
            if int(r[1]) == bi:
                bp = r[2:].copy()
                break
    return bv, bi, bp


def sharded_argmax(fn, points: np.ndarray, group=None):
    """``_get_max_discrete_points`` over candidates sharded across the ranks of ``group``:
    each rank evaluates its slice with the fused argmax kernels and one all-gather picks the winner.
    ``points`` [M, D] must be identical on every rank.  Returns (point [1, D], value, global index)."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_bounds(points.shape[0], rank, world)
    if hi > lo:
        idx, val = fn.fused_argmax(points[lo:hi])
        gidx = lo + idx
        pt = points[gidx]
    else:
        gidx, val, pt = -1, float("-inf"), np.zeros(points.shape[1])
    bv, bi, bp = allgather_best(val, gidx, pt, group=group)
    return bp[None, :], bv, bi


def sharded_thompson_argmin(trajectory, points: np.ndarray, group=None):
    """BASELINE config 4: batch Thompson sampling over candidates sharded across the ranks.  Every rank holds the same
    trajectory (same W, b, theta — broadcast or seeded identically) and evaluates its slice with the fused eval+argmin
    kernel; one all-gather per trajectory of (-value, global index) picks the minimiser.  Returns (points [B, D],
    values [B], global indices [B])."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_bounds(points.shape[0], rank, world)
    if hi > lo:
        mv, mi = trajectory.argmin_over(points[lo:hi])
    else:
        mv, mi = None, None
    B = len(mv) if mv is not None else int(getattr(trajectory, "_batch_size", 1) or 1)
    out_p, out_v, out_i = [], [], []
    for b in range(B):
        if mv is not None:
            val, gidx = -float(mv[b]), lo + int(mi[b])
            pt = points[gidx]
        else:
            val, gidx, pt = float("-inf"), -1, np.zeros(points.shape[1])
        bv, bi, bp = allgather_best(val, gidx, pt, group=group)
        out_p.append(bp if bp is not None else pt)
        out_v.append(-bv)
        out_i.append(bi)
    return np.stack(out_p), np.asarray(out_v), np.asarray(out_i)


def sharded_multistart(optimize_starts, starts: np.ndarray, group=None):
    """BASELINE config 5: the R multi-starts of ``generate_continuous_optimizer`` sharded across the ranks — each rank
    runs its slice of starts to convergence with no communication (``optimize_starts(starts_slice) -> (x [r, D],
    values [r])``), then one all-gather of (best value, global start index, x) picks the winner
    (optimizer.py:556-559)."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_bounds(starts.shape[0], rank, world)
    if hi > lo:
        xs, vals = optimize_starts(starts[lo:hi])
        vals = np.asarray(vals, dtype=np.float64).reshape(-1)
        j = int(np.argmax(np.where(np.isfinite(vals), vals, -np.inf)))
        val, gidx, pt = float(vals[j]), lo + j, np.asarray(xs)[j]
    else:
        val, gidx, pt = float("-inf"), -1, np.zeros(starts.shape[1])
    bv, bi, bp = allgather_best(val, gidx, pt, group=group)
    return (bp if bp is not None else pt)[None, :], bv, bi
