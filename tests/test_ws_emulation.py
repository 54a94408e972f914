"""CPU emulation of the warp-specialised Cartesian kernel (exadg_b200/csrc/cart_ws.hpp) against the oracle.

The CTA body of the CUDA kernel is written against a small run-time interface; tests/cpp/ws_emulate.cpp compiles the same body
with g++ on OS threads (192 or 256 per CTA, pthread barriers, synchronous bulk copies with a late-read check).  This pins the indexing,
the producer/consumer protocol and the folded 1-D tables of that kernel to the oracle without a GPU; a second build under
ThreadSanitizer reports any shared-memory access that the kernel's barriers do not order."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle.oracle import OracleOperator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "ws_emulate.cpp")
N3 = 125


def _build(path, extra):
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", *extra, SRC, "-o", path])
    lib = ctypes.CDLL(path)
    lib.wse_create.restype = ctypes.c_void_p
    lib.wse_create.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double]
    lib.wse_destroy.argtypes = [ctypes.c_void_p]
    for f in ("wse_n_owned", "wse_n_ghost", "wse_global_offset", "wse_smem_bytes"):
        getattr(lib, f).restype = ctypes.c_int64
        getattr(lib, f).argtypes = [ctypes.c_void_p]
    lib.wse_halo_max.argtypes = [ctypes.c_void_p]
    lib.wse_n_batches.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.wse_ghost_global.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.wse_vmult.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 3
    return lib


@pytest.fixture(scope="module", params=[(8, 2), (12, 2), (4, 4), (3, 4)], ids=["depth8", "depth12", "depth4_4producers", "staged3_4producers"])
def emu(request, tmp_path_factory):
    """the kernel instantiations of the library: neighbour cells fetched per producer round, producer warps"""
    depth, producers = request.param
    return _build(str(tmp_path_factory.mktemp("wse") / ("libwse%d_%d.so" % request.param)), ["-DWSE_R=%d" % depth, "-DWSE_NP=%d" % producers])


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _run(lib, n_sub, refine, rank, world, x_global, n_ctas, add=False, split=False):
    """vmult of rank's partition through the emulated kernel; returns (dst_local, lo, hi)."""
    h = lib.wse_create(n_sub, refine, rank, world, 1.0)
    try:
        n_owned, n_ghost, off = lib.wse_n_owned(h), lib.wse_n_ghost(h), lib.wse_global_offset(h)
        gg = np.zeros(max(n_ghost, 1), dtype=np.int64)
        lib.wse_ghost_global(h, _ptr(gg))
        # two doubles of padding behind the vectors: the staged variant copies whole neighbour cells in 16-byte granularity (the
        # library only selects it for even cell counts, where no copy leaves the vector; the emulation runs it on every mesh)
        src_store = np.zeros(n_owned * N3 + 2)
        src_store[:n_owned * N3] = x_global[off * N3:(off + n_owned) * N3]
        src = src_store[:n_owned * N3]
        ghost_store = np.zeros(max(n_ghost, 1) * N3 + 2)
        if n_ghost:
            ghost_store[:n_ghost * N3] = np.concatenate([x_global[g * N3:(g + 1) * N3] for g in gg[:n_ghost]])
        ghost = ghost_store[:max(n_ghost, 1) * N3]
        assert src.ctypes.data % 16 == 0 and ghost.ctypes.data % 16 == 0
        dst = np.full(n_owned * N3, 3.0) if add else np.full(n_owned * N3, np.nan)
        if split:  # interior batches, then the batches with ghost neighbours (the two launches of the multi-GPU path)
            assert lib.wse_n_batches(h, 1) + lib.wse_n_batches(h, 2) == lib.wse_n_batches(h, 0)
            err = lib.wse_vmult(h, _ptr(src), _ptr(ghost), _ptr(dst), int(add), n_ctas, 1)
            err += lib.wse_vmult(h, _ptr(src), _ptr(ghost), _ptr(dst), int(add), n_ctas, 2)
        else:
            err = lib.wse_vmult(h, _ptr(src), _ptr(ghost), _ptr(dst), int(add), n_ctas, 0)
        if lib.wse_halo_max(h) > 64:  # too irregular for the producers' staging area: the library keeps the pipelined kernel
            assert err < 0
            return None, off * N3, (off + n_owned) * N3
        assert err == 0, "bulk-copy protocol violated"
        assert lib.wse_smem_bytes(h) <= 113 * 1024
        return dst, off * N3, (off + n_owned) * N3
    finally:
        lib.wse_destroy(h)


def _oracle(n_sub, refine, x):
    return OracleOperator(4, n_sub, refine).vmult(x)


@pytest.mark.parametrize("n_sub,refine,n_ctas", [(1, 2, 1), (1, 2, 3), (3, 1, 2), (3, 1, 9), (1, 3, 4)])
def test_emulated_kernel_matches_oracle(emu, n_sub, refine, n_ctas):
    n = (n_sub << refine) ** 3 * N3
    x = np.random.default_rng(7).uniform(-1, 1, n)
    y, lo, hi = _run(emu, n_sub, refine, 0, 1, x, n_ctas)
    ref = _oracle(n_sub, refine, x)
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-13


@pytest.mark.parametrize("n_sub,refine,world", [(3, 2, 1), (1, 4, 1), (3, 3, 1), (3, 3, 2), (1, 4, 8)])
def test_plan_of_octet_aligned_meshes_fits_two_ctas_per_sm(emu, n_sub, refine, world):
    """24-cell batches = 3 octets of the Morton curve: at most 64 out-of-batch faces, 111 KB of shared memory per CTA"""
    for rank in range(world):
        h = emu.wse_create(n_sub, refine, rank, world, 1.0)
        try:
            assert emu.wse_halo_max(h) == 64
            assert emu.wse_smem_bytes(h) <= 228 * 1024 // 2 - 1024
            assert emu.wse_n_batches(h, 1) + emu.wse_n_batches(h, 2) == emu.wse_n_batches(h, 0)
        finally:
            emu.wse_destroy(h)


def test_emulated_kernel_dynamic_item_claiming(emu):
    """items claimed from a work counter through the ring in shared memory (single-launch partitioned vmult): same result"""
    n = 6 ** 3 * N3
    x = np.random.default_rng(10).uniform(-1, 1, n)
    emu.wse_set_dynamic(1)
    try:
        y, _, _ = _run(emu, 3, 1, 0, 1, x, 3)
    finally:
        emu.wse_set_dynamic(0)
    ref = _oracle(3, 1, x)
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-13


def test_emulated_kernel_add(emu):
    n = 4 ** 3 * N3
    x = np.random.default_rng(8).uniform(-1, 1, n)
    y, _, _ = _run(emu, 1, 2, 0, 1, x, 2, add=True)
    ref = _oracle(1, 2, x) + 3.0
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-13


@pytest.mark.parametrize("n_sub,refine,world,expect_supported", [(1, 3, 2, 2), (1, 3, 4, 4), (3, 2, 2, 2), (5, 0, 2, 1), (3, 1, 2, 1), (1, 3, 3, 1)])
def test_emulated_kernel_partitions(emu, n_sub, refine, world, expect_supported):
    """partitions with ghost cells, interior/boundary launches, ragged last batches; partitions that cut through the octets of the
    Morton curve have more than 64 out-of-batch faces per batch and are left to the pipelined kernel"""
    n = (n_sub << refine) ** 3 * N3
    x = np.random.default_rng(9).uniform(-1, 1, n)
    ref = _oracle(n_sub, refine, x)
    supported = 0
    for rank in range(world):
        for split in (False, True):
            y, lo, hi = _run(emu, n_sub, refine, rank, world, x, 2, split=split)
            if y is not None:
                supported += 1
                assert np.linalg.norm(y - ref[lo:hi]) / np.linalg.norm(ref) < 1e-13
    assert supported == 2 * expect_supported


@pytest.mark.parametrize("flags", [[], ["-DWSE_R=3", "-DWSE_NP=4"]], ids=["depth8", "staged3_4producers"])
def test_emulated_kernel_thread_sanitizer(tmp_path, flags):
    """the same run under ThreadSanitizer: the kernel's barriers must order every shared-memory access"""
    exe = str(tmp_path / "tsan_driver.py")
    lib = str(tmp_path / "libwse_tsan.so")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-g", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-fsanitize=thread", *flags, SRC, "-o", lib])
    tsan_rt = subprocess.run(["/usr/bin/g++", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(tsan_rt):
        pytest.skip("libtsan not available")
    with open(exe, "w") as f:
        f.write(
            "import ctypes, numpy as np, sys\n"
            "lib = ctypes.CDLL(sys.argv[1])\n"
            "lib.wse_create.restype = ctypes.c_void_p; lib.wse_create.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double]\n"
            "lib.wse_vmult.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 3\n"
            "h = lib.wse_create(1, 2, 0, 1, 1.0)\n"
            "x = np.random.default_rng(1).uniform(-1, 1, 64 * 125); g = np.zeros(1); y = np.zeros(64 * 125)\n"
            "p = lambda a: a.ctypes.data_as(ctypes.c_void_p)\n"
            "print('errors', lib.wse_vmult(h, p(x), p(g), p(y), 0, 1, 0), float(np.abs(y).sum()))\n")
    env = dict(os.environ, LD_PRELOAD=tsan_rt, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0")
    r = subprocess.run([os.sys.executable, exe, lib], capture_output=True, text=True, env=env, timeout=900)
    assert "errors 0" in r.stdout, r.stdout + r.stderr[-2000:]
    assert "WARNING: ThreadSanitizer: data race" not in r.stderr, r.stderr[-4000:]
