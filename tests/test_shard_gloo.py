"""Multi-rank host logic on CPU: keyframe shard plans are consistent across ranks (gloo, world_size 2
and 3): every halo pull names a slot its owner really holds, and the union of owned ranges is exact."""
import os
import socket

import numpy as np
import pytest

from sdmb200 import shard, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, per_rank, n_nbr, q, bounds=None):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nb = synth.neighbours(per_rank * world, n_nbr)
    plan = shard.make_plan(nb, per_rank, rank, world, bounds)
    mine = {"lo": plan.lo, "hi": plan.hi, "own": (plan.own_lo, plan.own_hi),
            "pulls": [(int(a) + plan.lo, int(r), int(s)) for a, r, s in zip(plan.halo_local, plan.halo_rank, plan.halo_peer_slot)]}
    allp = [None] * world
    dist.all_gather_object(allp, mine)
    ok = True
    for g, r, s in mine["pulls"]:
        owner = allp[r]
        ok &= owner["own"][0] <= g < owner["own"][1]          # the peer owns that keyframe
        ok &= owner["lo"] + s == g                            # and the slot index addresses it in the peer's arena
    owned = sorted(sum([list(range(*p["own"])) for p in allp], []))
    ok &= owned == list(range(per_rank * world))
    # every neighbour of an owned keyframe is held locally (pass 1 needs its inputs, pass 2 its planes)
    for kf in plan.owned_local:
        ok &= bool((plan.nbr_local[kf] >= 0).all())
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok), len(mine["pulls"])))


@pytest.mark.parametrize("world,per_rank,n_nbr,bounds", [(2, 16, 6, None), (3, 10, 6, None), (2, 12, 10, None),
                                                         (3, 10, 6, [0, 4, 19, 30]), (2, 16, 6, [0, 23, 32])])
def test_shard_plans_consistent(world, per_rank, n_nbr, bounds):
    """equal ranges and weighted ones (shard.balanced_bounds: a 4-keyframe rank whose neighbours reach into two peers)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, per_rank, n_nbr, q, bounds)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(n > 0 for _, _, n in res), "every rank of a chain has at least one halo keyframe"


def test_single_rank_plan_has_no_halo():
    nb = synth.neighbours(20, 6)
    plan = shard.make_plan(nb, 20, 0, 1)
    assert plan.lo == 0 and plan.hi == 20 and len(plan.halo_local) == 0
    assert np.array_equal(plan.nbr_local, nb)


def test_balanced_bounds():
    """contiguous ranges of nearly equal weight; every rank keeps at least one keyframe; equal weights give equal ranges"""
    assert shard.balanced_bounds(np.ones(40), 4) == [0, 10, 20, 30, 40]
    rng = np.random.default_rng(3)
    w = 50e3 + 45e3 * (np.sin(np.arange(1000) / 13.0) > 0) + rng.uniform(0, 5e3, 1000)   # the bench trajectory's pattern
    b = shard.balanced_bounds(w, 8)
    sums = np.array([w[b[r]:b[r + 1]].sum() for r in range(8)])
    equal = np.array([w[r * 125:(r + 1) * 125].sum() for r in range(8)])
    assert b[0] == 0 and b[-1] == 1000 and all(b1 > b0 for b0, b1 in zip(b, b[1:]))
    assert sums.max() / sums.mean() < 1.01 < equal.max() / equal.mean()
    assert shard.balanced_bounds([5, 0, 0, 0, 100, 0], 4) == [0, 1, 2, 4, 6] or len(shard.balanced_bounds([5, 0, 0, 0, 100, 0], 4)) == 5
    b = shard.balanced_bounds([1e9, 1, 1], 3)
    assert b == [0, 1, 2, 3]
    nb = synth.neighbours(30, 6)
    plans = [shard.make_plan(nb, 0, r, 3, [0, 4, 19, 30]) for r in range(3)]
    assert [(p.own_lo, p.own_hi) for p in plans] == [(0, 4), (4, 19), (19, 30)]
    assert sorted(set(int(r) for r in plans[0].halo_rank)) == [1]          # keyframes 4 .. 6 live on rank 1
