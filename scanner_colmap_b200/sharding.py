"""Multi-GPU sharding of the sequential-window pair set (one process per GPU).

Every pair of the reference op is independent (``sequential_matching.cc:139-181`` has no cross-pair state), so
the table is cut into contiguous anchor windows: rank g owns anchor images ``[s_g, e_g)`` and additionally needs
the descriptors of the ``overlap - 1`` images after ``e_g`` (its halo), which live in the next rank's (ranks')
HBM and are fetched peer-to-peer (``torch.distributed`` send/recv: NCCL over NVLink on GPUs, gloo in the CPU
tests).  There is no reduction and no other collective; match lists go back to the host per rank.

Exhaustive matching (every image against every later image, BASELINE configs[4]) has no window to cut: there the
upper triangle of the pair matrix is tiled in 2-D over image blocks (``plan_exhaustive``), tiles are dealt to
the ranks by cost, and a rank needs only the blocks its tiles touch (half of the descriptors on 8 GPUs) instead
of an all-gather of everything.
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


def halo_ops_packed(p, row_bytes: Callable[[int], int], get_send: Callable[[List[int]], "object"],
                    make_recv: Callable[[int, int], "object"], group=None):
    """The point-to-point operations of ``exchange_halo_packed`` without issuing them: (ops, {row: view}).  A caller
    whose send views and receive buffers do not move can build them once and re-issue them every step."""
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
    return ops, out


def exchange_halo_packed(p: ShardPlan, row_bytes: Callable[[int], int], get_send: Callable[[List[int]], "object"],
                         make_recv: Callable[[int, int], "object"], group=None) -> dict:
    """Same exchange with ONE message per peer instead of one per image (queuing 18 point-to-point operations
    through torch.distributed costs 0.2-0.45 ms of host time per step on 8 GPUs).  ``get_send(rows)`` returns a
    1-D uint8 tensor holding those rows back to back (a zero-copy view when they are adjacent in the descriptor
    pool, a packed copy otherwise); ``make_recv(src, nbytes)`` returns the 1-D uint8 receive buffer for everything
    coming from rank ``src``.  Returns {row: 1-D view of that row's bytes inside its receive buffer}.  Both sides
    derive the message sizes from the plan alone, so they always agree."""
    import torch.distributed as dist
    ops, out = halo_ops_packed(p, row_bytes, get_send, make_recv, group)
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    return out


# --------------------------------------------------------------------------------------------------------------
# Exhaustive matching: 2-D tiling of the upper-triangular pair matrix
# --------------------------------------------------------------------------------------------------------------
@dataclass
class ExhaustivePlan:
    rank: int
    world: int
    own: Tuple[int, int]                     # rows this rank uploads from the host (contiguous, balanced on bytes)
    blocks: List[Tuple[int, int]]            # the image blocks [s, e) of the tiling (same on every rank)
    tiles: List[Tuple[int, int]]             # (block i, block j), i <= j, assigned to this rank
    need: List[int]                          # every row this rank's tiles touch (ascending)
    recv: List[Tuple[int, int]]              # (row, owner rank): needed but not owned, ascending row
    send: List[Tuple[int, int]]              # (row, destination rank), ascending (row, destination)
    pairs: np.ndarray                        # uint32 [m, 2] (row1 < row2) this rank matches
    cost: float                              # sum n1*n2 over this rank's pairs
    imbalance: float                         # max over ranks / mean over ranks of cost (same on every rank)
    resident_fraction: float                 # largest share of all descriptor bytes any rank must hold


def _balanced_cuts(weights: np.ndarray, parts: int) -> List[Tuple[int, int]]:
    """Contiguous ranges with near-equal weight sums (ranges may be empty when parts > len(weights))."""
    n = len(weights)
    acc = np.concatenate([[0.0], np.cumsum(weights, dtype=np.float64)])
    cuts = [0]
    for g in range(1, parts):
        c = int(np.searchsorted(acc, acc[-1] * g / parts, side="left")) if acc[-1] > 0 else round(n * g / parts)
        cuts.append(min(max(c, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[g], cuts[g + 1]) for g in range(parts)]


def _tile_cost(sz: np.ndarray, pre: np.ndarray, pre2: np.ndarray, bi: Tuple[int, int], bj: Tuple[int, int]) -> float:
    """sum of n_i * n_j over the pairs i < j of tile (bi, bj); bi == bj is a diagonal tile."""
    si = pre[bi[1]] - pre[bi[0]]
    if bi == bj:
        return 0.5 * (si * si - (pre2[bi[1]] - pre2[bi[0]]))
    return si * (pre[bj[1]] - pre[bj[0]])


def _assign_tiles(sizes: np.ndarray, world: int, nblocks: int):
    sz = sizes.astype(np.float64)
    pre = np.concatenate([[0.0], np.cumsum(sz)])
    pre2 = np.concatenate([[0.0], np.cumsum(sz * sz)])
    blocks = [b for b in _balanced_cuts(sz, nblocks) if b[1] > b[0]]
    nb = len(blocks)
    tiles = [(i, j) for i in range(nb) for j in range(i, nb)]
    cost = {t: _tile_cost(sz, pre, pre2, blocks[t[0]], blocks[t[1]]) for t in tiles}
    tiles = [t for t in tiles if cost[t] > 0]
    bw = [pre[b[1]] - pre[b[0]] for b in blocks]
    load = [0.0] * world
    held = [set() for _ in range(world)]
    mine = [[] for _ in range(world)]
    # largest tile first, to the least loaded rank; among (near-)equally loaded ranks the one that already holds
    # the tile's blocks, so that a rank touches as few blocks as possible
    total = sum(cost.values())
    for t in sorted(tiles, key=lambda t: (-cost[t], t)):
        lo = min(load)
        cands = [r for r in range(world) if load[r] - lo <= 1e-9 * max(total, 1.0)]
        r = min(cands, key=lambda r: (sum(bw[b] for b in set(t) - held[r]), r))
        load[r] += cost[t]
        held[r] |= set(t)
        mine[r].append(t)
    mean = total / world if total > 0 else 1.0
    imbalance = max(load) / mean if total > 0 else 1.0
    resident = max((sum(bw[b] for b in h) for h in held), default=0.0) / max(pre[-1], 1.0)
    return blocks, mine, load, imbalance, resident


def plan_exhaustive(sizes: Sequence[int], world: int, rank: int, max_imbalance: float = 1.03) -> ExhaustivePlan:
    """All pairs i < j over ``len(sizes)`` images on ``world`` ranks.  The block count of the tiling is the one
    (up to 4 * world blocks) whose cost-balanced tile assignment stays within ``max_imbalance`` while a rank holds
    the smallest share of the descriptors; if none does, the best balanced one.  Every rank computes the same plan."""
    sizes = np.asarray(sizes, dtype=np.int64)
    N = len(sizes)
    best = None
    for nb in range(1, max(2, 4 * world) + 1):
        blocks, mine, load, imb, res = _assign_tiles(sizes, world, nb)
        key = (0, res, imb, nb) if imb <= max_imbalance else (1, imb, res, nb)
        if best is None or key < best[0]:
            best = (key, blocks, mine, load, imb, res)
        if len(blocks) < nb:
            break
    _, blocks, mine, load, imb, res = best
    owners = _balanced_cuts(sizes.astype(np.float64), world)
    owner = np.zeros(N, dtype=np.int64)
    for g, (s, e) in enumerate(owners):
        owner[s:e] = g

    def needed(g):
        rows = set()
        for (i, j) in mine[g]:
            rows.update(range(*blocks[i]))
            rows.update(range(*blocks[j]))
        return rows

    need = sorted(needed(rank))
    recv = [(i, int(owner[i])) for i in need if owner[i] != rank]
    send = sorted((i, g) for g in range(world) if g != rank for i in needed(g) if owner[i] == rank)
    pr = []
    for (i, j) in sorted(mine[rank]):
        (s1, e1), (s2, e2) = blocks[i], blocks[j]
        if i == j:
            pr += [(a, b) for a in range(s1, e1) for b in range(a + 1, e1)]
        else:
            pr += [(a, b) for a in range(s1, e1) for b in range(s2, e2)]
    return ExhaustivePlan(rank, world, owners[rank], blocks, sorted(mine[rank]), need, recv, send,
                          np.asarray(pr, dtype=np.uint32).reshape(-1, 2), float(load[rank]), float(imb), float(res))
