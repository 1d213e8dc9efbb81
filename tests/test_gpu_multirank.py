"""Two ranks (two processes, gloo rendezvous): keyframe shards + CUDA-IPC halo pulls must reproduce the single-context
result, i.e. the oracle over the whole trajectory, on every owned keyframe.  The ranks use two GPUs when the box has
them (`gpurun --gpus 2`: the planes then cross NVLink) and share cuda:0 otherwise.  Two protocols:
  * host-ordered: sdm_export_arena / sdm_import_peer_arena / sdm_pull_halo between synchronize + barrier pairs;
  * device-ordered: sdm_export_peer_handle / sdm_import_peer / sdm_set_halo once, then sdm_pass1 / sdm_exchange /
    sdm_pass2 per step with no host synchronisation or barrier in between (flags in peer memory), run for several
    steps so that the acknowledgement path (the next pass 1 waits for the pullers) is exercised, and once more through
    sdm_run_loop(exchange = 1)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, per_rank, q, mode="host"):
    try:
        for p in (os.path.join(ROOT, "eao-slam_b200", "python"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
            sys.path.insert(0, p)
        import torch.distributed as dist
        import oracle_py as O
        from helpers import compare_planes
        from sdmb200 import api, shard, synth
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        W, H, N = 320, 240, 6
        nb_global = synth.neighbours(per_rank * world, N)
        plan = shard.make_plan(nb_global, per_rank, rank, world)
        nb_local = np.where(plan.nbr_local >= 0, plan.nbr_local, 0).astype(np.int32)
        sc = synth.make_scene(plan.n_local, W, H, N, seed=41, first=plan.lo, nbr_idx=nb_local)
        import torch
        ndev = torch.cuda.device_count()
        ctx = api.Context(width=W, height=H, max_keyframes=plan.n_local, device=rank % max(1, ndev))
        owned = list(plan.owned_local)
        items = api.make_items(owned, sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)
        handles = [None] * world
        if mode == "host":
            ctx.upload_scene(sc)
            dist.all_gather_object(handles, ctx.export_arena())
            for r in set(int(x) for x in plan.halo_rank):
                ctx.import_peer_arena(r, handles[r])
            ctx.pass1(items)
            ctx.synchronize(); dist.barrier()            # every rank's pass 1 is complete
            ctx.pull_halo(plan.halo_local, plan.halo_rank, plan.halo_peer_slot)
            ctx.pass2(items)
            ctx.synchronize(); dist.barrier()            # nobody frees / overwrites planes a peer still reads
        else:
            dist.all_gather_object(handles, ctx.export_peer_handle(rank))
            for r in set(int(x) for x in plan.halo_rank):
                ctx.import_peer(r, handles[r])
            ctx.set_halo(plan.halo_local, plan.halo_rank, plan.halo_peer_slot)
            dist.barrier()                               # every import (puller registration) is done: the only barrier
            if mode == "device":
                for step in range(3):                    # re-upload + both passes: no host sync, no barrier per step
                    ctx.upload_scene(sc)
                    ctx.pass1(items)
                    ctx.exchange()
                    ctx.pass2(items)
            else:                                        # the same through the library's pipelined loop
                for step in range(2):
                    ctx.run_loop(upload=ctx.upload_descs(sc, range(plan.n_local)), pass1=items, pass2=items, chunk=3,
                                 exchange=True)
            ctx.synchronize()
        dev = {k: np.zeros((len(owned), H, W) + ((3,) if k == "points" else ()), np.float32)
               for k in ("depth", "sigma", "checked", "points")}
        for j, s in enumerate(owned):
            r = ctx.download(s)
            for k in dev:
                dev[k][j] = r[k]
        # reference: the oracle over the WHOLE trajectory (global neighbour lists), owned slice
        g = synth.make_scene(per_rank * world, W, H, N, seed=41, nbr_idx=nb_global)
        osc = O.OracleScene(g)
        osc.run()
        sl = slice(plan.own_lo, plan.own_hi)
        for k in ("im", "grad", "theta", "Tcw", "min_depth", "max_depth"):  # a rank's slice == the global trajectory
            assert np.array_equal(getattr(sc, k), getattr(g, k)[plan.lo:plan.hi]), k

        class Ref:
            depth, sigma, checked, points = osc.depth[sl], osc.sigma[sl], osc.checked[sl], osc.points[sl]
        rep = compare_planes(dev, Ref)
        bad = np.argwhere(dev["checked"].view(np.uint32) != Ref.checked.view(np.uint32))
        rep["bad"] = [(int(i) + plan.own_lo, int(y), int(x), float(dev["checked"][i, y, x]), float(Ref.checked[i, y, x])) for i, y, x in bad[:8]]
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()
        q.put((rank, "ok", rep, len(plan.halo_local)))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "error", traceback.format_exc(), 0))


@pytest.mark.parametrize("mode", ["host", "device", "run_loop"])
def test_two_ranks_halo_exchange(mode):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 10, q, mode)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    if os.environ.get("SDM_DUMP"):
        import json
        json.dump([(r, st, rp) for r, st, rp, _ in res], open(os.environ["SDM_DUMP"], "w"), default=str)
    for rank, status, rep, n_halo in res:
        assert status == "ok", rep
        assert n_halo == 3, "N=6 neighbours -> 3 halo keyframes across the shard boundary"
        assert rep["pass2_set_mismatch"] == 0 and rep["checked_bit_mismatch"] == 0, rep
