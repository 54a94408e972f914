"""ctypes wrapper around the CPU oracle (oracle/sipg_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under exadg_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

PERIODIC, DIRICHLET, NEUMANN = 0, 1, 2


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("sipg_oracle.c", "sipg_fast.inc", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.orc_create_hypercube.restype = C.c_void_p
        L.orc_create_hypercube.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), C.c_double]
        L.orc_create_periodic_box_lean.restype = C.c_void_p
        L.orc_create_periodic_box_lean.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double]
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_long, dp, C.POINTER(C.c_long), C.POINTER(C.c_ubyte), C.POINTER(C.c_ubyte), C.c_double]
        L.orc_destroy.argtypes = [C.c_void_p]
        for name in ("orc_n_dofs", "orc_n_cells", "orc_n_interior_faces", "orc_n_boundary_faces"):
            getattr(L, name).restype = C.c_long
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_get_tau.argtypes = [C.c_void_p, dp]
        L.orc_mass_vmult.argtypes = [C.c_void_p, dp, dp]
        L.orc_get_cell_jxw.argtypes = [C.c_void_p, dp]
        L.orc_get_mesh.argtypes = [C.c_void_p, dp, C.POINTER(C.c_long), C.POINTER(C.c_ubyte), C.POINTER(C.c_ubyte)]
        L.orc_vmult.argtypes = [C.c_void_p, dp, dp]
        L.orc_vmult_add.argtypes = [C.c_void_p, dp, dp]
        L.orc_vmult_cellwise.argtypes = [C.c_void_p, dp, dp, C.c_int]
        L.orc_vmult_fast.restype = C.c_int
        L.orc_vmult_fast.argtypes = [C.c_void_p, dp, dp, C.c_int]
        L.orc_max_threads.restype = C.c_int
        L.orc_calculate_diagonal.argtypes = [C.c_void_p, dp]
        L.orc_calculate_inverse_diagonal.argtypes = [C.c_void_p, dp]
        L.orc_cg.restype = C.c_int
        L.orc_cg.argtypes = [C.c_void_p, dp, dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, dp, C.c_int, C.POINTER(C.c_int)]
        L.orc_cheb_create.restype = C.c_void_p
        L.orc_cheb_create.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int]
        L.orc_cheb_destroy.argtypes = [C.c_void_p]
        L.orc_cheb_get.argtypes = [C.c_void_p, dp]
        L.orc_cheb_set.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_cheb_vmult.argtypes = [C.c_void_p, dp, dp]
        L.orc_cheb_step.argtypes = [C.c_void_p, dp, dp]
        L.orc_cg_chebyshev.restype = C.c_int
        L.orc_cg_chebyshev.argtypes = [C.c_void_p, C.c_void_p, dp, dp, C.c_double, C.c_double, C.c_int, C.c_int, dp, C.c_int, C.POINTER(C.c_int)]
        L.orc_rhs_sine.argtypes = [C.c_void_p, dp]
        L.orc_l2_error_sine.restype = C.c_double
        L.orc_l2_error_sine.argtypes = [C.c_void_p, dp]
        L.orc_dof_coordinates.argtypes = [C.c_void_p, dp]
        L.orc_get_basis.argtypes = [C.c_int] + [dp] * 7
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OracleOperator:
    """SIPG Laplace operator on the reference's hypercube grids (CPU, FP64)."""

    def __init__(self, degree, n_sub=1, refine=0, mapping_degree=1, deformation=0.0, frequency=2,
                 bc=(PERIODIC,) * 6, ip_factor=1.0, lean=False):
        L = lib()
        bc_arr = (C.c_int * 6)(*bc)
        if lean:  # periodic Cartesian box without stored geometry: vmult_fast only (the CPU baseline at benchmark size)
            assert deformation == 0.0 and tuple(bc) == (PERIODIC,) * 6 and mapping_degree == 1
            self.h = L.orc_create_periodic_box_lean(degree, n_sub, refine, float(ip_factor))
        else:
            self.h = L.orc_create_hypercube(degree, n_sub, refine, mapping_degree, float(deformation), frequency, bc_arr, float(ip_factor))
        if not self.h:
            raise ValueError("oracle: unsupported parameters")
        self.degree = degree
        self.mapping_degree = mapping_degree
        self.n_dofs = L.orc_n_dofs(self.h)
        self.n_cells = L.orc_n_cells(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def _vec(self):
        return np.zeros(self.n_dofs, dtype=np.float64)

    def vmult(self, src):
        dst = self._vec()
        lib().orc_vmult(self.h, _p(dst), _p(np.ascontiguousarray(src, dtype=np.float64)))
        return dst

    def vmult_add(self, dst, src):
        lib().orc_vmult_add(self.h, _p(dst), _p(np.ascontiguousarray(src, dtype=np.float64)))
        return dst

    def vmult_cellwise(self, src, n_threads=0, dst=None):
        dst = self._vec() if dst is None else dst
        lib().orc_vmult_cellwise(self.h, _p(dst), _p(np.ascontiguousarray(src, dtype=np.float64)), n_threads)
        return dst

    def vmult_fast(self, src, n_threads=0, dst=None):
        """Vectorised CPU baseline (8 cells per SIMD batch, compile-time degree; uniform boxes with interior faces only, otherwise it
        falls back to vmult_cellwise).  self.fast_path_used tells which of the two ran."""
        dst = self._vec() if dst is None else dst
        self.fast_path_used = bool(lib().orc_vmult_fast(self.h, _p(dst), _p(np.ascontiguousarray(src, dtype=np.float64)), n_threads))
        return dst

    def diagonal(self):
        d = self._vec()
        lib().orc_calculate_diagonal(self.h, _p(d))
        return d

    def inverse_diagonal(self):
        d = self._vec()
        lib().orc_calculate_inverse_diagonal(self.h, _p(d))
        return d

    def cg(self, b, x0=None, jacobi=False, abs_tol=1e-20, rel_tol=1e-12, max_it=10000, cellwise=True):
        x = self._vec() if x0 is None else np.array(x0, dtype=np.float64)
        hist = np.zeros(max_it + 1)
        conv = C.c_int(0)
        it = lib().orc_cg(self.h, _p(x), _p(np.ascontiguousarray(b, dtype=np.float64)), int(jacobi), abs_tol, rel_tol, max_it, int(cellwise), _p(hist), len(hist), C.byref(conv))
        return x, it, hist[: it + 1], bool(conv.value)

    def rhs_sine(self):
        r = self._vec()
        lib().orc_rhs_sine(self.h, _p(r))
        return r

    def l2_error_sine(self, u):
        return lib().orc_l2_error_sine(self.h, _p(np.ascontiguousarray(u, dtype=np.float64)))

    def dof_coordinates(self):
        xyz = np.zeros((self.n_dofs, 3))
        lib().orc_dof_coordinates(self.h, _p(xyz))
        return xyz

    def mass_vmult(self, src):
        """MassKernel with QGauss(k+1) (mass_kernel.h:32-93), one component."""
        dst = self._vec()
        lib().orc_mass_vmult(self.h, _p(dst), _p(np.ascontiguousarray(src, dtype=np.float64)))
        return dst

    def cell_jxw(self):
        w = np.zeros((self.n_cells, (self.degree + 1) ** 3))
        lib().orc_get_cell_jxw(self.h, _p(w))
        return w

    def tau(self):
        t = np.zeros(self.n_cells)
        lib().orc_get_tau(self.h, _p(t))
        return t

    def mesh(self):
        np3 = (self.mapping_degree + 1) ** 3
        xmap = np.zeros((self.n_cells, np3, 3))
        nb = np.zeros((self.n_cells, 6), dtype=np.int64)
        nbface = np.zeros((self.n_cells, 6), dtype=np.uint8)
        bt = np.zeros((self.n_cells, 6), dtype=np.uint8)
        lib().orc_get_mesh(self.h, _p(xmap), nb.ctypes.data_as(C.POINTER(C.c_long)), nbface.ctypes.data_as(C.POINTER(C.c_ubyte)), bt.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return xmap, nb, nbface, bt


class OracleHelmholtz:
    """scaling_factor_mass * M + viscosity * A_SIPG applied to every component of a vector-valued DG field (FESystem(FE_DGQ(k)^n):
    cell-major, then component, then lexicographic node): the momentum / viscous operator of the incompressible Navier-Stokes module in
    Laplace formulation with constant viscosity (momentum_operator.cpp:376-426, viscous_operator.h:365-386, 489-560)."""

    def __init__(self, op, n_components=3, scaling_factor_mass=1.0, viscosity=1.0):
        self.op, self.nc, self.alpha, self.nu = op, n_components, scaling_factor_mass, viscosity
        self.n_dofs = op.n_dofs * n_components

    def _split(self, v):
        n3 = (self.op.degree + 1) ** 3
        return np.ascontiguousarray(v, dtype=np.float64).reshape(self.op.n_cells, self.nc, n3)

    def _each(self, f, v):
        u = self._split(v)
        out = np.empty_like(u)
        for c in range(self.nc):
            out[:, c, :] = f(np.ascontiguousarray(u[:, c, :]).ravel()).reshape(self.op.n_cells, -1)
        return out.ravel()

    def vmult(self, src):
        return self._each(lambda u: self.alpha * self.op.mass_vmult(u) + self.nu * self.op.vmult(u), src)

    def mass_vmult(self, src):
        return self._each(self.op.mass_vmult, src)

    def diagonal(self):
        e = np.ones(self.op.n_dofs)
        n = self.op.degree + 1
        S = basis_tables(self.op.degree)["S"]  # S[q, j] = l_j(x_q)
        S2 = S * S
        w = self.op.cell_jxw().reshape(self.op.n_cells, n, n, n)
        dm = np.einsum("qi,rj,sk,nsrq->nkji", S2, S2, S2, w, optimize=True).ravel()
        d = self.alpha * dm + self.nu * self.op.diagonal()
        return np.repeat(d.reshape(self.op.n_cells, 1, -1), self.nc, axis=1).ravel() + 0 * e[0]


class OracleChebyshev:
    """dealii::PreconditionChebyshev with point-Jacobi, as configured by ExaDG's ChebyshevSmoother."""

    def __init__(self, op, degree=5, smoothing_range=20.0, eig_cg_n_iterations=20, cellwise=True):
        self.op = op
        self.c = lib().orc_cheb_create(op.h, degree, smoothing_range, eig_cg_n_iterations, int(cellwise))
        out = np.zeros(4)
        lib().orc_cheb_get(self.c, _p(out))
        self.lambda_min_est, self.lambda_max_est, self.theta, self.delta = out

    def __del__(self):
        if getattr(self, "c", None):
            lib().orc_cheb_destroy(self.c)
            self.c = None

    def set_interval(self, theta, delta):
        lib().orc_cheb_set(self.c, theta, delta)
        self.theta, self.delta = theta, delta

    def vmult(self, src):
        dst = self.op._vec()
        lib().orc_cheb_vmult(self.c, _p(dst), _p(np.ascontiguousarray(src, dtype=np.float64)))
        return dst

    def step(self, dst, src):
        dst = np.array(dst, dtype=np.float64)
        lib().orc_cheb_step(self.c, _p(dst), _p(np.ascontiguousarray(src, dtype=np.float64)))
        return dst

    def cg(self, b, abs_tol=1e-20, rel_tol=1e-10, max_it=10000, cellwise=True):
        x = self.op._vec()
        hist = np.zeros(max_it + 1)
        conv = C.c_int(0)
        it = lib().orc_cg_chebyshev(self.op.h, self.c, _p(x), _p(np.ascontiguousarray(b, dtype=np.float64)), abs_tol, rel_tol, max_it, int(cellwise), _p(hist), len(hist), C.byref(conv))
        return x, it, hist[: it + 1], bool(conv.value)


def basis_tables(degree):
    n = degree + 1
    xn, xq, w = np.zeros(n), np.zeros(n), np.zeros(n)
    S, D = np.zeros((n, n)), np.zeros((n, n))
    fv, fd = np.zeros((2, n)), np.zeros((2, n))
    lib().orc_get_basis(degree, _p(xn), _p(xq), _p(w), _p(S), _p(D), _p(fv), _p(fd))
    return dict(xn=xn, xq=xq, w=w, S=S, D=D, fv=fv, fd=fd)


def synthetic_vector(n_dofs, seed=42):
    """Counter-based hash -> uniform(-1,1); layout-independent, reproducible on CPU and GPU
    (SURVEY 8(d)): splitmix64 of (global dof index + seed * 2^32)."""
    i = np.arange(n_dofs, dtype=np.uint64) + (np.uint64(seed) << np.uint64(32))
    with np.errstate(over="ignore"):
        z = i + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0
