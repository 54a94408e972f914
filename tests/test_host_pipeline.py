"""Host logic of the pipelined host-buffer vmult (exadg_b200_vmult_host_pipelined, csrc/host_pipeline.hpp) on the CPU: the chunk plan
must never apply a chunk before every chunk holding a face neighbour of its cells has been uploaded, and on the benchmark mesh it
must overlap most of the two PCIe directions (model: duration of one call in units of a one-direction transfer; 2 = no overlap)."""
import numpy as np
import pytest

import exadg_b200
from exadg_b200.laplace_operator import PartitionPlan


def check_plan(n_sub, refine, cells_per_chunk, boundary=(0,) * 6):
    P = exadg_b200.host_pipeline_plan(n_sub, refine, cells_per_chunk, boundary)
    K = P["n_chunks"]
    nb = PartitionPlan(n_sub, refine, 0, 1, boundary).neighbors  # [cells][6], -1 on the boundary
    n_cells = nb.shape[0]
    assert K == -(-n_cells // cells_per_chunk)
    assert sorted(P["upload_order"]) == list(range(K)) and sorted(P["compute_order"]) == list(range(K))
    pos = np.empty(K, dtype=np.int64)
    pos[P["upload_order"]] = np.arange(K)
    chunk = np.arange(n_cells) // cells_per_chunk
    for c in range(K):
        cells = np.nonzero(chunk == c)[0]
        neigh = nb[cells].ravel()
        deps = set(chunk[neigh[neigh >= 0]].tolist()) | {c}
        ready = P["ready_chunk"][c]
        assert ready in deps
        assert max(pos[d] for d in deps) == pos[ready], "chunk %d would be applied before its neighbours have arrived" % c
    ready_pos = pos[P["ready_chunk"][P["compute_order"]]]
    assert np.all(np.diff(ready_pos) >= 0), "compute order follows the upload order"
    return P


@pytest.mark.parametrize("n_sub,refine,cpc", [(3, 2, 48), (1, 3, 24), (5, 1, 72), (3, 2, 1536)])
def test_chunks_are_applied_only_after_their_neighbours_arrived(n_sub, refine, cpc):
    check_plan(n_sub, refine, cpc)


def test_plan_with_boundaries():
    check_plan(3, 1, 24, (1, 2, 1, 1, 0, 0))


def test_partitioned_operators_keep_the_sequential_path():
    assert exadg_b200.host_pipeline_plan(3, 2, 48, rank=0, world=2)["n_chunks"] == 0


def test_benchmark_mesh_overlaps_most_of_the_two_transfers():
    P = exadg_b200.host_pipeline_plan(3, 5)  # 96^3 cells, library default: 512 batches of 24 cells per chunk (measured optimum)
    assert P["n_chunks"] == 72
    assert P["model"] < 1.5
    P = exadg_b200.host_pipeline_plan(3, 5, 1536)  # finer chunks overlap more on paper (and cost more per chunk in practice)
    assert P["n_chunks"] == 576
    assert P["model"] < 1.3


# ---- direct variant: pieces in address order, readiness per kernel unit (HostStreamPlan) ----
def check_stream_plan(n_sub, refine, unit, cells_per_piece, boundary=(0,) * 6):
    P = exadg_b200.host_stream_plan(n_sub, refine, unit, cells_per_piece, boundary)
    nb = PartitionPlan(n_sub, refine, 0, 1, boundary).neighbors
    n_cells = nb.shape[0]
    K = P["n_steps"]
    assert K == -(-n_cells // cells_per_piece)
    assert P["piece_begin"][0] == 0 and P["piece_begin"][-1] == n_cells and np.all(np.diff(P["piece_begin"]) > 0)
    n_units = -(-n_cells // unit)
    assert sorted(P["units"].tolist()) == list(range(n_units)), "every unit is applied exactly once"
    assert P["step_begin"][0] == 0 and P["step_begin"][-1] == n_units and np.all(np.diff(P["step_begin"]) >= 0)
    piece = np.searchsorted(P["piece_begin"], np.arange(n_cells), side="right") - 1  # piece (= upload step) of every cell
    for i in range(K):
        us = P["units"][P["step_begin"][i]:P["step_begin"][i + 1]]
        assert np.all(np.diff(us) > 0)
        for u in us:
            cells = np.arange(u * unit, min((u + 1) * unit, n_cells))
            neigh = nb[cells].ravel()
            need = max(piece[cells].max(), piece[neigh[neigh >= 0]].max(initial=0))
            assert need == i, "unit %d is applied behind upload %d but is complete with upload %d" % (u, i, need)
    return P


@pytest.mark.parametrize("n_sub,refine,unit,cpp", [(3, 2, 24, 48), (1, 3, 24, 96), (5, 1, 16, 80), (3, 2, 1, 100), (3, 2, 32, 1536), (3, 1, 64, 64)])
def test_units_are_applied_behind_the_upload_that_completes_them(n_sub, refine, unit, cpp):
    check_stream_plan(n_sub, refine, unit, cpp)


def test_stream_plan_with_boundaries_and_ragged_last_unit():
    check_stream_plan(3, 1, 24, 48, (1, 2, 1, 1, 0, 0))  # 216 cells: 9 units
    check_stream_plan(5, 0, 24, 48)                      # 125 cells: the last unit has 5 cells, the last piece 29


@pytest.mark.parametrize("rank", [0, 1])
def test_stream_plan_of_a_partition_leaves_the_units_with_ghost_neighbours_to_the_ghost_import(rank):
    unit, cpp = 24, 96
    P = exadg_b200.host_stream_plan(3, 2, unit, cpp, rank=rank, world=2)
    part = PartitionPlan(3, 2, rank, 2)
    nb, n_owned = part.neighbors, part.n_owned  # local indices, ghosts >= n_owned
    assert P["n_steps"] == -(-n_owned // cpp)
    n_units = -(-n_owned // unit)
    touches = np.array([(nb[u * unit:min((u + 1) * unit, n_owned)] >= n_owned).any() for u in range(n_units)])
    assert touches.any() and not touches.all()
    assert sorted(P["units"].tolist()) == np.nonzero(~touches)[0].tolist(), "exactly the units without ghost neighbours are in the steps"
    piece = np.arange(n_owned) // cpp
    for i in range(P["n_steps"]):
        for u in P["units"][P["step_begin"][i]:P["step_begin"][i + 1]]:
            cells = np.arange(u * unit, min((u + 1) * unit, n_owned))
            neigh = nb[cells].ravel()
            assert max(piece[cells].max(), piece[neigh[neigh >= 0]].max(initial=0)) == i


def test_benchmark_mesh_streams_with_little_more_than_one_transfer():
    P = exadg_b200.host_stream_plan(3, 5)  # 96^3 cells, 24-cell batches, library default: 36 pieces of 24576 cells (measured optimum)
    assert P["n_steps"] == 36
    assert P["model"] < 1.13               # chunk plan: 1.46
    assert exadg_b200.host_stream_plan(3, 5, 24, 12288)["model"] < 1.12
    assert exadg_b200.host_stream_plan(3, 5, 24, 98304)["model"] < 1.21
