"""Multi-GPU sharding of the sequential-window pair set (one process per GPU).

Every pair of the reference op is independent (``sequential_matching.cc:139-181`` has no cross-pair state), so
the table is cut into contiguous anchor windows: rank g owns anchor images ``[s_g, e_g)`` and additionally needs
the descriptors of the ``overlap - 1`` images after ``e_g`` (its halo), which live in the next rank's (ranks')
HBM and are fetched peer-to-peer (``torch.distributed`` send/recv: NCCL over NVLink on GPUs, gloo in the CPU
tests).  There is no reduction and no other collective; match lists go back to the host per rank.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def row_costs(sizes: Sequence[int], overlap: int) -> np.ndarray:
    """Work of anchor row i: n_i * sum of the sizes of its overlap-1 successors (dot products to evaluate)."""
    n = np.asarray(sizes, dtype=np.float64)
    c = np.concatenate([[0.0], np.cumsum(n)])
    N = len(n)
    hi = np.minimum(np.arange(N) + overlap, N)
    return n * (c[hi] - c[np.arange(N) + 1])


def partition(sizes: Sequence[int], overlap: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous anchor ranges, one per rank, balanced on ``row_costs`` (equal image counts when sizes are
    uniform up to the lighter tail rows).  Ranges may be empty when world > number of images."""
    N = len(sizes)
    cost = row_costs(sizes, overlap)
    total = cost.sum()
    if total <= 0:
        cuts = [round(N * g / world) for g in range(world + 1)]
    else:
        acc = np.concatenate([[0.0], np.cumsum(cost)])
        cuts = [0]
        for g in range(1, world):
            # first index whose prefix cost reaches g/world of the total
            cuts.append(int(np.searchsorted(acc, total * g / world, side="left")))
        cuts.append(N)
        for g in range(1, world + 1):
            cuts[g] = max(cuts[g], cuts[g - 1])
    return [(cuts[g], cuts[g + 1]) for g in range(world)]


@dataclass
class ShardPlan:
    rank: int
    world: int
    own: Tuple[int, int]            # anchor rows [s, e)
    halo: Tuple[int, int]           # extra rows [e, h) whose descriptors are needed
    recv: List[Tuple[int, int]]     # (row, owner rank) in ascending row order
    send: List[Tuple[int, int]]     # (row, destination rank) in ascending (row, destination) order
    pairs: np.ndarray               # uint32 [m, 2] (row1, row2) table-row indices this rank matches


def plan(sizes: Sequence[int], overlap: int, world: int, rank: int) -> ShardPlan:
    N = len(sizes)
    parts = partition(sizes, overlap, world)
    owner = np.empty(N, dtype=np.int64)
    for g, (s, e) in enumerate(parts):
        owner[s:e] = g
    halos = [(e, min(e + overlap - 1, N)) if e > s else (e, e) for (s, e) in parts]
    s, e = parts[rank]
    recv = [(i, int(owner[i])) for i in range(*halos[rank])]
    send = sorted((i, g) for g in range(world) if g != rank for i in range(*halos[g]) if owner[i] == rank)
    pr = [(r, r2) for r in range(s, e) for r2 in range(r + 1, min(r + overlap, N))]
    return ShardPlan(rank, world, (s, e), halos[rank], recv, send,
                     np.asarray(pr, dtype=np.uint32).reshape(-1, 2))


def exchange_halo(p: ShardPlan, get_tensor: Callable[[int], "object"], make_recv: Callable[[int], "object"],
                  group=None) -> dict:
    """Post all sends (this rank's rows other ranks need) and receives (this rank's halo rows), wait, and
    return {row: received tensor}.  ``get_tensor(row)`` returns the tensor to send (a zero-copy view of the
    descriptor pool on GPUs); ``make_recv(row)`` allocates the receive buffer."""
    import torch.distributed as dist
    ops, out = [], {}
    for row, dst in p.send:
        ops.append(dist.P2POp(dist.isend, get_tensor(row), dst, group=group))
    for row, src in p.recv:
        buf = make_recv(row)
        out[row] = buf
        ops.append(dist.P2POp(dist.irecv, buf, src, group=group))
    if ops:
        # one coalesced group: both sides list their ops in ascending row order per peer, so they pair up
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    return out


def exchange_halo_packed(p: ShardPlan, row_bytes: Callable[[int], int], get_send: Callable[[List[int]], "object"],
                         make_recv: Callable[[int, int], "object"], group=None) -> dict:
    """Same exchange with ONE message per peer instead of one per image (queuing 18 point-to-point operations
    through torch.distributed costs 0.2-0.45 ms of host time per step on 8 GPUs).  ``get_send(rows)`` returns a
    1-D uint8 tensor holding those rows back to back (a zero-copy view when they are adjacent in the descriptor
    pool, a packed copy otherwise); ``make_recv(src, nbytes)`` returns the 1-D uint8 receive buffer for everything
    coming from rank ``src``.  Returns {row: 1-D view of that row's bytes inside its receive buffer}.  Both sides
    derive the message sizes from the plan alone, so they always agree."""
    import torch.distributed as dist
    by_dst, by_src = {}, {}
    for row, dst in p.send:
        by_dst.setdefault(dst, []).append(row)
    for row, src in p.recv:
        by_src.setdefault(src, []).append(row)
    ops, out = [], {}
    for dst in sorted(by_dst):
        ops.append(dist.P2POp(dist.isend, get_send(by_dst[dst]), dst, group=group))
    for src in sorted(by_src):
        rows = by_src[src]
        buf = make_recv(src, sum(row_bytes(r) for r in rows))
        off = 0
        for r in rows:
            out[r] = buf[off:off + row_bytes(r)]
            off += row_bytes(r)
        ops.append(dist.P2POp(dist.irecv, buf, src, group=group))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    return out
