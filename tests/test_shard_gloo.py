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


def _worker(rank, world, port, per_rank, n_nbr, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nb = synth.neighbours(per_rank * world, n_nbr)
    plan = shard.make_plan(nb, per_rank, rank, world)
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


@pytest.mark.parametrize("world,per_rank,n_nbr", [(2, 16, 6), (3, 10, 6), (2, 12, 10)])
def test_shard_plans_consistent(world, per_rank, n_nbr):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, per_rank, n_nbr, q)) for r in range(world)]
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
