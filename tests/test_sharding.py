"""Host-side multi-GPU logic on CPU: the window + halo partition, and a world_size-2 gloo run of the halo
exchange (the N > 1 path of bench.py with the GPU matcher replaced by bookkeeping)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scanner_colmap_b200 import sequential_pairs, sharding, synth


@pytest.mark.parametrize("n,overlap,world", [(100, 10, 1), (100, 10, 2), (1000, 20, 8), (37, 10, 4), (5, 10, 4), (3, 4, 8),
                                             (200, 201, 8), (24, 24, 3)])  # the last two: exhaustive (configs[4])
def test_plans_cover_every_pair_exactly_once(n, overlap, world):
    sizes = [8192] * n
    want = {tuple(p) for p in sequential_pairs(list(range(n)), overlap).tolist()}
    seen = []
    for r in range(world):
        p = sharding.plan(sizes, overlap, world, r)
        seen += [tuple(x) for x in p.pairs.tolist()]
        s, e = p.own
        assert p.halo == ((e, min(e + overlap - 1, n)) if e > s else (e, e))
        needed = {int(x) for x in p.pairs.reshape(-1)}
        assert needed <= set(range(s, e)) | {row for row, _ in p.recv}
    assert len(seen) == len(set(seen)) and set(seen) == want
    # what one rank sends is exactly what the others expect to receive
    sends = {(row, r, dst) for r in range(world) for row, dst in sharding.plan(sizes, overlap, world, r).send}
    recvs = {(row, src, r) for r in range(world) for row, src in sharding.plan(sizes, overlap, world, r).recv}
    assert sends == recvs


def test_uniform_partition_is_near_equal_and_ragged_is_cost_balanced():
    parts = sharding.partition([8192] * 1000, 20, 8)
    assert parts[0][0] == 0 and parts[-1][1] == 1000
    counts = [e - s for s, e in parts]
    assert max(counts[:-1]) - min(counts[:-1]) <= 1 and 0 <= counts[-1] - counts[0] <= 20   # tail rows are lighter
    sizes = synth.ragged_sizes(2000).tolist()
    cost = sharding.row_costs(sizes, 10)
    per = [cost[s:e].sum() for s, e in sharding.partition(sizes, 10, 8)]
    assert max(per) / np.mean(per) < 1.05


@pytest.mark.parametrize("n,world", [(200, 8), (200, 4), (200, 2), (200, 1), (24, 3), (7, 8), (2, 4)])
def test_exhaustive_plan_tiles_the_triangle(n, world):
    """BASELINE configs[4]: all pairs i < j, 2-D tiled over image blocks; every pair exactly once, send/recv lists
    agree, every needed row is owned or received."""
    sizes = [16384] * n
    plans = [sharding.plan_exhaustive(sizes, world, r) for r in range(world)]
    seen = [tuple(x) for p in plans for x in p.pairs.tolist()]
    assert len(seen) == len(set(seen)) == n * (n - 1) // 2 and all(a < b for a, b in seen)
    sends = {(row, r, dst) for r, p in enumerate(plans) for row, dst in p.send}
    recvs = {(row, src, r) for r, p in enumerate(plans) for row, src in p.recv}
    assert sends == recvs
    owned = [set(range(*p.own)) for p in plans]
    assert sorted(x for o in owned for x in o) == list(range(n))           # every image has exactly one owner
    for r, p in enumerate(plans):
        needed = {int(x) for x in p.pairs.reshape(-1)}
        assert needed <= owned[r] | {row for row, _ in p.recv}
        assert set(p.need) >= needed


def test_exhaustive_plan_is_balanced_and_holds_half_the_descriptors_on_8_gpus():
    p = sharding.plan_exhaustive([16384] * 200, 8, 0)
    assert p.imbalance <= 1.03 and p.resident_fraction <= 0.51 and len(p.blocks) == 4
    sizes = synth.ragged_sizes(300).tolist()
    q = sharding.plan_exhaustive(sizes, 8, 3)
    assert q.imbalance <= 1.03 and q.resident_fraction < 0.8


def _worker(rank, world, port, n, overlap, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sizes = [40 + 3 * (i % 5) for i in range(n)]
        p = sharding.plan(sizes, overlap, world, rank)
        own = {i: torch.from_numpy(synth.make_image(i, sizes[i], track_step=4)) for i in range(*p.own)}
        got = sharding.exchange_halo(p, lambda row: own[row].reshape(-1),
                                     lambda row: torch.empty(sizes[row] * 128, dtype=torch.uint8))
        ok = all(np.array_equal(got[row].numpy().reshape(-1, 128), synth.make_image(row, sizes[row], track_step=4))
                 for row, _ in p.recv)
        # one message per peer: same bytes, views into one receive buffer per source
        got2 = sharding.exchange_halo_packed(p, lambda row: sizes[row] * 128,
                                             lambda rows: torch.cat([own[r].reshape(-1) for r in rows]),
                                             lambda src, nbytes: torch.empty(nbytes, dtype=torch.uint8))
        ok = ok and sorted(got2) == sorted(got) and all(torch.equal(got2[row], got[row]) for row in got)
        # exhaustive plan: rows come from the owners of the blocks this rank's tiles touch
        pe = sharding.plan_exhaustive(sizes, world, rank)
        mine = {i: torch.from_numpy(synth.make_image(i, sizes[i], track_step=4)) for i in range(*pe.own)}
        got3 = sharding.exchange_halo_packed(pe, lambda row: sizes[row] * 128,
                                             lambda rows: torch.cat([mine[r].reshape(-1) for r in rows]),
                                             lambda src, nbytes: torch.empty(nbytes, dtype=torch.uint8))
        ok = ok and sorted(got3) == [row for row, _ in pe.recv] and all(
            np.array_equal(got3[row].numpy().reshape(-1, 128), synth.make_image(row, sizes[row], track_step=4)) for row in got3)
        q.put((rank, ok, len(p.recv), len(p.send), len(p.pairs)))
    finally:
        dist.destroy_process_group()


def test_halo_exchange_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n, overlap, world = 24, 6, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, overlap, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, *_ in res)
    assert res[0][2] == overlap - 1 and res[0][3] == 0          # rank 0 receives its halo, sends nothing
    assert res[1][2] == 0 and res[1][3] == overlap - 1
    assert res[0][4] + res[1][4] == (overlap - 1) * n - overlap * (overlap - 1) // 2


def test_plans_property_random_sizes():
    """Random ragged image sets, overlaps and world sizes: both plans cover every pair exactly once, the send / recv
    lists of all ranks agree, and every row a rank's pairs name is owned or received."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(1, 16384), min_size=1, max_size=40), st.integers(2, 45), st.integers(1, 8))
    def check(sizes, overlap, world):
        n = len(sizes)
        want = {tuple(p) for p in sequential_pairs(list(range(n)), overlap).tolist()}
        plans = [sharding.plan(sizes, overlap, world, r) for r in range(world)]
        seen = [tuple(x) for p in plans for x in p.pairs.tolist()]
        assert len(seen) == len(set(seen)) and set(seen) == want
        assert {(row, r, dst) for r, p in enumerate(plans) for row, dst in p.send} == \
               {(row, src, r) for r, p in enumerate(plans) for row, src in p.recv}
        for p in plans:
            s, e = p.own
            assert {int(x) for x in p.pairs.reshape(-1)} <= set(range(s, e)) | {row for row, _ in p.recv}
        ex = [sharding.plan_exhaustive(sizes, world, r) for r in range(world)]
        seen = [tuple(x) for p in ex for x in p.pairs.tolist()]
        assert len(seen) == len(set(seen)) == n * (n - 1) // 2
        assert {(row, r, dst) for r, p in enumerate(ex) for row, dst in p.send} == \
               {(row, src, r) for r, p in enumerate(ex) for row, src in p.recv}
        owned = [set(range(*p.own)) for p in ex]
        assert sorted(x for o in owned for x in o) == list(range(n))
        for r, p in enumerate(ex):
            assert {int(x) for x in p.pairs.reshape(-1)} <= owned[r] | {row for row, _ in p.recv}

    check()
