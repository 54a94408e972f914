"""CPU checks of the multigrid restatement (oracle/multigrid.py) and of the level logic the library exports.

The multigrid-preconditioned solve is pinned to the reference's golden fixtures: the sine application runs CG + multigrid
(applications/poisson/sine/application.h:166-177), and its printed L2 errors only depend on the converged solution."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle.multigrid import OracleMultigrid, Transfer, initialize_levels
from oracle.oracle import OracleOperator, basis_tables

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sine_l2_errors.json")))


def test_level_list_matches_the_documented_example_of_the_reference():
    # multigrid_preconditioner_base.cpp:103-134: h_levels = [0 1 2], p_levels = [1 3 7]
    assert initialize_levels("pMG", "Bisect", 7, 3) == [(2, 1), (2, 3), (2, 7)]
    assert initialize_levels("phMG", "Bisect", 7, 3) == [(0, 1), (1, 1), (2, 1), (2, 3), (2, 7)]
    assert initialize_levels("hMG", "Bisect", 7, 3) == [(0, 7), (1, 7), (2, 7)]
    assert initialize_levels("hpMG", "Bisect", 7, 3) == [(0, 1), (0, 3), (0, 7), (1, 7), (2, 7)]
    assert initialize_levels("pMG", "DecreaseByOne", 4, 1) == [(0, 1), (0, 2), (0, 3), (0, 4)]
    assert initialize_levels("pMG", "GoToOne", 5, 2) == [(1, 1), (1, 5)]


def test_library_level_list_equals_the_restatement():
    import exadg_b200
    for t in ("hMG", "pMG", "hpMG", "phMG"):
        for seq in ("GoToOne", "DecreaseByOne", "Bisect"):
            for k in range(1, 8):
                for nh in (1, 2, 4):
                    assert exadg_b200.multigrid_levels(t, seq, k, nh) == initialize_levels(t, seq, k, nh), (t, seq, k, nh)
    with pytest.raises(exadg_b200.ExaDGError):
        exadg_b200.multigrid_levels("cphMG", "Bisect", 4, 2)


@pytest.mark.parametrize("kf,kc,h", [(4, 2, False), (2, 1, False), (7, 3, False), (3, 3, True), (1, 1, True)])
def test_transfer_embeds_polynomials_and_restriction_is_the_transpose(kf, kc, h):
    T = Transfer(kf, kc, h)
    nf, nc = kf + 1, kc + 1
    n_coarse = 3
    rng = np.random.default_rng(1)
    # a polynomial of the coarse degree is reproduced exactly at the fine nodes
    xc, xf = basis_tables(kc)["xn"], basis_tables(kf)["xn"]
    poly = lambda x, y, z: (1 + x) ** kc * (2 - y) ** kc + z ** kc - x * y * z  # noqa: E731
    uc = np.zeros((n_coarse, nc, nc, nc))
    for c in range(n_coarse):
        uc[c] = poly(xc[None, None, :] + c, xc[None, :, None], xc[:, None, None])
    fine = np.zeros(n_coarse * (8 if h else 1) * nf ** 3)
    T.prolongate_add(fine, uc.ravel())
    uf = fine.reshape(n_coarse, 8 if h else 1, nf, nf, nf)
    for c in range(n_coarse):
        for child in range(8 if h else 1):
            s = 0.5 if h else 1.0
            ox, oy, oz = (0.5 * (child & 1), 0.5 * ((child >> 1) & 1), 0.5 * ((child >> 2) & 1)) if h else (0, 0, 0)
            ref = poly(ox + s * xf[None, None, :] + c, oy + s * xf[None, :, None], oz + s * xf[:, None, None])
            assert np.allclose(uf[c, child], ref, rtol=1e-12, atol=1e-12)
    # <P c, f> = <c, R f>
    c = rng.standard_normal(n_coarse * nc ** 3)
    f = rng.standard_normal(fine.size)
    Pc = T.prolongate_add(np.zeros_like(f), c)
    Rf = T.restrict_add(np.zeros_like(c), f)
    assert abs(Pc @ f - c @ Rf) < 1e-12 * np.linalg.norm(Pc) * np.linalg.norm(f)


def test_galerkin_consistency_of_the_p_transfer_on_the_periodic_box():
    """R A_fine P = A_coarse for the p-transfer on affine cells (both sides integrate the same polynomials exactly) -- up to the
    penalty factor, which depends on the degree: compare with the coarse operator built with the fine penalty (ip_factor scaled)."""
    kf, kc = 4, 2
    fine, coarse = OracleOperator(kf, 1, 1), OracleOperator(kc, 1, 1, ip_factor=(kf + 1.0) ** 2 / (kc + 1.0) ** 2)
    T = Transfer(kf, kc, False)
    rng = np.random.default_rng(3)
    c = rng.standard_normal(coarse.n_dofs)
    lhs = T.restrict_add(np.zeros(coarse.n_dofs), fine.vmult(T.prolongate_add(np.zeros(fine.n_dofs), c)))
    rhs = coarse.vmult(c)
    assert np.linalg.norm(lhs - rhs) < 1e-11 * np.linalg.norm(rhs)


@pytest.mark.parametrize("mg_type", ["pMG", "hMG", "phMG", "hpMG"])
def test_v_cycle_is_a_contraction_and_symmetric(mg_type):
    k, refine = 3, 2
    bc = (1, 2, 1, 1, 1, 1)
    levels = initialize_levels(mg_type, "Bisect", k, refine + 1)
    mg = OracleMultigrid(levels, n_sub=1, deformation=0.1, bc=bc)
    A = mg.ops[-1]
    rng = np.random.default_rng(0)
    u, v = rng.standard_normal(A.n_dofs), rng.standard_normal(A.n_dofs)
    # the coarse solve is iterative (rel 1e-3) and warm-started: symmetry only approximately
    Mu, Mv = mg.vmult(u), mg.vmult(v)
    assert abs(Mu @ v - u @ Mv) < 2e-2 * abs(Mu @ v)
    # error propagation: ||(I - M A) e||_A < ||e||_A
    e = rng.standard_normal(A.n_dofs)
    e1 = e - mg.vmult(A.vmult(e))
    assert np.sqrt(e1 @ A.vmult(e1)) < 0.6 * np.sqrt(e @ A.vmult(e))


@pytest.mark.parametrize("mesh,degree", [("cartesian", 2), ("cartesian", 4), ("curvilinear", 3), ("curvilinear", 4)])
def test_multigrid_preconditioned_solve_reproduces_the_reference_golden_l2_errors(mesh, degree):
    cfg = GOLD["config"]
    deform = cfg["deformation_curvilinear"] if mesh == "curvilinear" else 0.0
    levels = initialize_levels("phMG", "Bisect", degree, cfg["refine"] + 1)
    mg = OracleMultigrid(levels, n_sub=cfg["n_cells_1d_coarse"], mapping_degree=cfg["mapping_degree"], deformation=deform,
                         frequency=cfg["frequency"], bc=tuple(cfg["bc"]))
    A = mg.ops[-1]
    b = A.rhs_sine()
    x, its, hist = mg.pcg(A, b, rel_tol=cfg["cg_rel_tol"])
    assert hist[-1] < cfg["cg_rel_tol"] * hist[0]
    assert its < 25, its   # multigrid: mesh- and degree-independent counts (Jacobi-preconditioned CG needs > 100)
    err = A.l2_error_sine(x)
    gold = GOLD[mesh][degree - 1]
    assert abs(err / gold - 1.0) < 6e-6, (err, gold)
