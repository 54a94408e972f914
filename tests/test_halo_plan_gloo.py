"""N>1 path on CPU: world_size-2 (and 4) gloo processes build their partition / halo plan through the C ABI
(host-only entry points, no GPU), exchange the planned cells with torch.distributed and check that every rank's
ghost range ends up holding exactly the neighbour cells its owned cells reference."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_sub, refine, bc, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import exadg_b200
        from oracle.oracle import OracleOperator
        plan = exadg_b200.PartitionPlan(n_sub, refine, rank, world, bc)
        n = n_sub << refine
        n_cells = n ** 3
        # p4est-style contiguous equal-count chunks of the (coarse cell, Morton) curve
        assert plan.global_offset == n_cells * rank // world
        assert plan.n_owned == n_cells * (rank + 1) // world - n_cells * rank // world
        # global reference connectivity from the oracle's (independent) mesh generator
        _, nb_ref, _, _ = OracleOperator(1, n_sub, refine, 1, 0.0, 2, bc).mesh()
        g0 = plan.global_offset
        for c in range(plan.n_owned):
            for f in range(6):
                loc = plan.neighbors[c, f]
                glob = -1 if loc < 0 else (g0 + loc if loc < plan.n_owned else plan.ghost_global_ids[loc - plan.n_owned])
                assert glob == nb_ref[g0 + c, f]
        # "cell data" = global cell id repeated; exchange exactly what the plan says
        payload = 5
        owned = torch.arange(g0, g0 + plan.n_owned, dtype=torch.float64).repeat_interleave(payload).reshape(plan.n_owned, payload)
        ghost = torch.full((plan.n_ghost, payload), -1.0, dtype=torch.float64)
        reqs, keep = [], []
        for p in plan.peers:
            send = owned[torch.from_numpy(p["send_cells"].astype(np.int64))].contiguous()
            keep.append(send)
            reqs.append(dist.isend(send, p["rank"]))
            reqs.append(dist.irecv(ghost[p["recv_begin"]:p["recv_begin"] + p["recv_count"]], p["rank"]))
        for r in reqs:
            r.wait()
        expect = torch.from_numpy(plan.ghost_global_ids.astype(np.float64)).repeat_interleave(payload).reshape(plan.n_ghost, payload)
        assert torch.equal(ghost, expect)
        # a global reduction the way CG's dot products use it
        s = torch.tensor([float(plan.n_owned)])
        dist.all_reduce(s)
        assert int(s.item()) == n_cells
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_sub,refine,bc", [(2, 1, 2, (0,) * 6), (2, 3, 1, (1, 2, 1, 1, 1, 1)), (4, 1, 2, (0,) * 6), (3, 5, 0, (0,) * 6)])
def test_halo_plan_exchange(world, n_sub, refine, bc):
    import __graft_entry__ as ge
    ge.build()
    ret = mp.get_context("spawn").Manager().dict()
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, n_sub, refine, bc, ret), nprocs=world, join=True)
    assert sorted(ret.keys()) == list(range(world))
