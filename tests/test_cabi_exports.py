"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/exadg_b200.h declares; argument errors are reported through status codes (no GPU calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    import exadg_b200
    return exadg_b200.load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "exadg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(exadg_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    import exadg_b200
    names = declared_symbols()
    assert len(names) >= 35
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(exadg_b200.EXPORTED_SYMBOLS) == names


def test_version_and_error_string(lib):
    assert lib.exadg_b200_version() >= 100
    assert isinstance(lib.exadg_b200_last_error(), bytes)


def test_null_arguments_return_status(lib):
    assert lib.exadg_b200_create_hypercube(None, None) == 1
    assert b"null" in lib.exadg_b200_last_error()
    assert lib.exadg_b200_destroy(None) == 0
    assert lib.exadg_b200_n(None) == -1


def test_bad_degree_is_rejected_before_any_cuda_call(lib):
    import exadg_b200
    d = exadg_b200.HypercubeDesc()
    d.degree, d.n_subdivisions, d.n_refinements, d.mapping_degree, d.world = 9, 1, 1, 1, 1
    h = C.c_void_p()
    assert lib.exadg_b200_create_hypercube(C.byref(d), C.byref(h)) == 1
    assert b"degree" in lib.exadg_b200_last_error()


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under exadg_b200/ or include/ may import, link or call it."""
    for base in ("exadg_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "liboracle" not in text and "from oracle" not in text and "import oracle" not in text and "orc_" not in text, f


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import exadg_b200
    with pytest.raises(exadg_b200.ExaDGError):
        exadg_b200.LaplaceOperator.hypercube(degree=2, n_subdivisions=2)


def test_create_rejects_inconsistent_mesh_before_any_cuda_call(lib):
    """exadg_b200_create validates the arrays a binding passes (status code, no exception across the C ABI)."""
    import numpy as np
    import exadg_b200
    d = exadg_b200.MeshDesc()
    pts = np.zeros((1, 8, 3))
    nb = np.full((1, 6), -1, dtype=np.int32)
    nf = np.zeros((1, 6), dtype=np.uint8)
    bt = np.zeros((1, 6), dtype=np.uint8)  # says "interior" although there is no neighbour
    d.degree, d.mapping_degree, d.n_cells_owned, d.n_cells_ghost = 2, 1, 1, 0
    d.mapping_points = pts.ctypes.data_as(C.POINTER(C.c_double))
    d.neighbors = nb.ctypes.data_as(C.POINTER(C.c_int32))
    d.neighbor_face = nf.ctypes.data_as(C.POINTER(C.c_uint8))
    d.boundary_type = bt.ctypes.data_as(C.POINTER(C.c_uint8))
    d.ip_factor = 1.0
    h = C.c_void_p()
    assert lib.exadg_b200_create(C.byref(d), C.byref(h)) == 1
    assert b"boundary_type" in lib.exadg_b200_last_error()
    nb[0, 0] = 7  # out of range
    bt[:] = 1
    bt[0, 0] = 0
    assert lib.exadg_b200_create(C.byref(d), C.byref(h)) == 1
    assert b"out of range" in lib.exadg_b200_last_error()


def test_partition_plan_is_contiguous_and_complete(lib):
    import exadg_b200
    world, total, ghosts = 4, 0, 0
    for r in range(world):
        p = exadg_b200.PartitionPlan(3, 1, r, world)
        assert p.global_offset == total
        total += p.n_owned
        ghosts += p.n_ghost
        assert all(pp["rank"] != r for pp in p.peers)
        assert sum(pp["recv_count"] for pp in p.peers) == p.n_ghost
    assert total == 6 ** 3 and ghosts > 0
