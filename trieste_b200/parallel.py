"""Multi-GPU sharding of the candidate axis.

The path shards embarrassingly: every rank holds a replica of the model state and owns a contiguous
slice of the candidates (or q-batches, or L-BFGS starts).  The only exchange is ONE all-gather of a
(value, global index, x[D]) tuple per rank (k tuples for the running top-k of ``generate_initial_points``); every rank then
selects the same winner(s) with the reference's tie rule (first maximal index, optimizer.py:149).  One process per GPU,
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


def _to_host(x) -> np.ndarray:
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def allgather_best(value, global_index, point: Optional[np.ndarray] = None, group=None, device=None):
    """The path's single collective.  ``value`` / ``global_index`` may be scalars (one winner) or arrays of length B
    (B independent winners — the q trajectories of a Thompson batch travel in ONE all-gather); ``point`` is [D] or [B, D].
    Returns (best_value, best_global_index, best_point or None) with the same leading shape, identical on every rank:
    larger value wins, ties -> lower global index (first-max, optimizer.py:149); NaN and empty shards (index -1) never win."""
    import torch
    import torch.distributed as dist

    scalar = np.ndim(value) == 0
    vals = np.atleast_1d(np.asarray(value, dtype=np.float64))
    idxs = np.atleast_1d(np.asarray(global_index, dtype=np.int64))
    B = vals.shape[0]
    pts = None if point is None else np.asarray(_to_host(point), dtype=np.float64).reshape(B, -1)
    D = 0 if pts is None else pts.shape[1]
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        bv, bi, bp = vals.copy(), idxs.copy(), None if pts is None else pts.copy()
    else:
        world = dist.get_world_size(group)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        payload = np.zeros((B, 2 + D))
        payload[:, 0] = np.where(vals == vals, vals, -np.inf)
        payload[:, 1] = idxs  # exact for indices < 2^53
        if D:
            payload[:, 2:] = pts
        mine = torch.as_tensor(payload, device=device)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine, group=group)
        rows = torch.stack(gathered).cpu().numpy()  # [world, B, 2 + D]: one device-to-host copy
        bv, bi = np.empty(B), np.empty(B, dtype=np.int64)
        bp = None if not D else np.empty((B, D))
        for b in range(B):
            v, i = merge_best([(rows[r, b, 0], int(rows[r, b, 1])) for r in range(world)])
            bv[b], bi[b] = v, i
            if D:
                src = [r for r in range(world) if int(rows[r, b, 1]) == i]
                bp[b] = rows[src[0], b, 2:] if src else np.nan
    if scalar:
        return float(bv[0]), int(bi[0]), None if bp is None else bp[0]
    return bv, bi, bp


def _gather_rows(points, indices) -> np.ndarray:
    """points[indices] as a host array — ONE indexed gather + one device-to-host copy for a torch.cuda tensor."""
    idx = np.asarray(indices, dtype=np.int64).reshape(-1)
    if hasattr(points, "detach"):
        import torch

        return points[torch.as_tensor(idx, device=points.device)].detach().cpu().numpy()
    return np.asarray(points)[idx]


def _world_rank(group):
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    return world, (dist.get_rank(group) if world > 1 else 0)


def sharded_argmax_local(fn, local_points, global_offset: int, group=None):
    """The rank's own shard of the candidates is already in place (``local_points`` [m, D], NumPy or torch.cuda; global
    indices ``global_offset + i``): fused evaluation + argmax over the shard, then the single all-gather.
    Returns (point [1, D], value, global index), identical on every rank."""
    if local_points.shape[0] > 0:
        idx, val = fn.fused_argmax(local_points)
        gidx, pt = global_offset + idx, _to_host(local_points[idx])
    else:
        gidx, val, pt = -1, float("-inf"), np.zeros(local_points.shape[1])
    bv, bi, bp = allgather_best(val, gidx, pt, group=group)
    return bp[None, :], bv, bi


def sharded_argmax(fn, points, group=None):
    """``_get_max_discrete_points`` over candidates sharded across the ranks of ``group``:
    each rank evaluates its slice with the fused argmax kernels and one all-gather picks the winner.
    ``points`` [M, D] must be identical on every rank.  Returns (point [1, D], value, global index)."""
    world, rank = _world_rank(group)
    lo, hi = shard_bounds(points.shape[0], rank, world)
    return sharded_argmax_local(fn, points[lo:hi], lo, group=group)


def sharded_thompson_argmin_local(trajectory, local_points, global_offset: int, group=None):
    """BASELINE config 4 on the rank's own shard: every rank holds the same trajectories (same W, b, theta) and evaluates
    its candidates with the fused eval+argmin kernel; the B (value, index, point) triples of the batch travel in ONE
    all-gather.  Returns (points [B, D], values [B], global indices [B])."""
    D = local_points.shape[1]
    if local_points.shape[0] > 0:
        mv, mi = trajectory.argmin_over(local_points)
        B = len(mv)
        vals, gidx = -np.asarray(mv, dtype=np.float64), global_offset + np.asarray(mi, dtype=np.int64)
        pts = _gather_rows(local_points, mi)
    else:
        B = int(getattr(trajectory, "_batch_size", 1) or 1)
        vals, gidx, pts = np.full(B, -np.inf), np.full(B, -1, dtype=np.int64), np.zeros((B, D))
    bv, bi, bp = allgather_best(vals, gidx, pts, group=group)
    return bp, -bv, bi


def sharded_thompson_argmin(trajectory, points, group=None):
    """As above with ``points`` [M, D] identical on every rank (sliced here)."""
    world, rank = _world_rank(group)
    lo, hi = shard_bounds(points.shape[0], rank, world)
    return sharded_thompson_argmin_local(trajectory, points[lo:hi], lo, group=group)


def sharded_multistart_local(optimize_starts, local_starts, global_offset: int, group=None):
    """BASELINE config 5 on the rank's own shard of the multi-starts: ``optimize_starts(starts) -> (x [r, D], values [r])``
    runs them to convergence with no communication, then one all-gather of (best value, global start index, x) picks the
    winner (optimizer.py:556-559)."""
    if local_starts.shape[0] > 0:
        xs, vals = optimize_starts(local_starts)
        vals = np.asarray(_to_host(vals), dtype=np.float64).reshape(-1)
        j = int(np.argmax(np.where(np.isfinite(vals), vals, -np.inf)))
        val, gidx, pt = float(vals[j]), global_offset + j, _to_host(xs)[j]
    else:
        val, gidx, pt = float("-inf"), -1, np.zeros(local_starts.shape[1])
    bv, bi, bp = allgather_best(val, gidx, pt, group=group)
    return bp[None, :], bv, bi


def sharded_multistart(optimize_starts, starts, group=None):
    """As above with ``starts`` [R, D] identical on every rank (sliced here)."""
    world, rank = _world_rank(group)
    lo, hi = shard_bounds(starts.shape[0], rank, world)
    return sharded_multistart_local(optimize_starts, starts[lo:hi], lo, group=group)


def merge_topk(values, indices, k: int):
    """``tf.math.top_k`` over the union of per-rank candidates: values descending, ties -> lower global index
    (optimizer.py:321-335 keeps a running top-k with exactly this primitive).  NaN values and indices < 0 (padding of short
    shards) are dropped.  Returns the positions (into the flattened inputs) of the winners, at most k of them."""
    v = np.asarray(values, dtype=np.float64).reshape(-1)
    i = np.asarray(indices, dtype=np.int64).reshape(-1)
    keep = np.flatnonzero((i >= 0) & (v == v))
    order = keep[np.lexsort((i[keep], -v[keep]))]  # primary: value descending; secondary: global index ascending
    return order[: max(int(k), 0)]


def allgather_topk(values, global_indices, points, k: int, group=None, device=None):
    """Every rank contributes its local top-k (values [m], global indices [m], points [m, D], m <= k, any order); ONE
    all-gather of k (2 + D)-word tuples per rank; every rank returns the same global top-k
    (values [k'], global indices [k'], points [k', D]), k' = min(k, number of valid candidates)."""
    import torch
    import torch.distributed as dist

    vals = np.asarray(values, dtype=np.float64).reshape(-1)
    idxs = np.asarray(global_indices, dtype=np.int64).reshape(-1)
    pts = np.asarray(_to_host(points), dtype=np.float64).reshape(len(vals), -1)
    D = pts.shape[1]
    payload = np.zeros((k, 2 + D))
    payload[:, 1] = -1.0  # padding rows: never selected
    m = min(k, len(vals))
    payload[:m, 0], payload[:m, 1], payload[:m, 2:] = np.where(vals[:m] == vals[:m], vals[:m], -np.inf), idxs[:m], pts[:m]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        mine = torch.as_tensor(payload, device=device)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine, group=group)
        payload = torch.stack(gathered).cpu().numpy().reshape(world * k, 2 + D)
    sel = merge_topk(payload[:, 0], payload[:, 1].astype(np.int64), k)
    return payload[sel, 0], payload[sel, 1].astype(np.int64), payload[sel, 2:]


def sharded_topk_local(fn, local_points, global_offset: int, k: int, group=None):
    """The initial-point selection of ``generate_initial_points`` (optimizer.py:247-341) over candidates sharded across
    the ranks: the rank evaluates ``fn`` on its own shard ([m, D] -> values [m]), takes its local top-k with the native
    ``tb_topk`` and one all-gather of k tuples per rank yields the global top-k, identical on every rank.
    Returns (points [k, D], values [k], global indices [k])."""
    import torch

    from .sampler import top_k

    # the rank's own GPU, read before any native call: it sorts the shard's values and carries the all-gather
    dev = torch.cuda.current_device() if torch.cuda.is_available() else None
    m = int(local_points.shape[0])
    if m > 0:
        vals = fn(local_points[:, None, :])
        vals = vals.reshape(-1)
        tv, ti = top_k(vals, min(k, m), device=dev)
        ti_h = np.asarray(_to_host(ti), dtype=np.int64)
        tv_h, pts = np.asarray(_to_host(tv), dtype=np.float64), _gather_rows(local_points, ti_h)
        gi = global_offset + ti_h
    else:
        tv_h, gi, pts = np.zeros(0), np.zeros(0, dtype=np.int64), np.zeros((0, local_points.shape[1]))
    import torch.distributed as dist

    nccl = dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl"
    bv, bi, bp = allgather_topk(tv_h, gi, pts, k, group=group, device=torch.device("cuda", dev) if nccl and dev is not None else None)
    return bp, bv, bi


def sharded_topk(fn, points, k: int, group=None):
    """As above with ``points`` [M, D] identical on every rank (sliced here)."""
    world, rank = _world_rank(group)
    lo, hi = shard_bounds(points.shape[0], rank, world)
    return sharded_topk_local(fn, points[lo:hi], lo, k, group=group)
