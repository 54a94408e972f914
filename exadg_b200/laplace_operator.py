"""Python mirror of the reference's operator / solver interface for the SIPG Laplace path.

Names, argument meaning and error behaviour follow
  ExaDG::OperatorBase                I/operators/operator_base.h:123-398
  ExaDG::Poisson::LaplaceOperator    I/poisson/spatial_discretization/laplace_operator.h:238-300
  ExaDG::Krylov::KrylovSolver        I/solvers_and_preconditioners/solvers/iterative_solvers_dealii_wrapper.h:101-246
  ExaDG::JacobiPreconditioner        I/solvers_and_preconditioners/preconditioners/jacobi_preconditioner.h:33-86
  ExaDG::ChebyshevSmoother           I/solvers_and_preconditioners/multigrid/smoothers/chebyshev_smoother.h:35-175
Everything forwards to the C ABI; vectors are torch CUDA float64 tensors (device memory only).
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np


class ExaDGError(RuntimeError):
    """The reference throws dealii::ExcMessage through AssertThrow; the C ABI returns a status."""


def _lib():
    from . import load_library
    return load_library()


def _check(status):
    if status != 0:
        raise ExaDGError("exadg_b200 status %d: %s" % (status, _lib().exadg_b200_last_error().decode()))


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise ExaDGError("exadg_b200 needs a CUDA device (there is no CPU fallback)")
    return torch


def _ptr(t, n=None):
    torch = _torch()
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise ExaDGError("vectors must be contiguous CUDA float64 tensors")
    if n is not None and t.numel() != n:
        raise ExaDGError("vector has %d entries, operator expects %d" % (t.numel(), n))
    return C.c_void_p(t.data_ptr())


class LaplaceOperator:
    """Poisson::LaplaceOperator<3, double, 1> (DG) on the GPU."""

    value_type = np.float64  # typedef Number value_type (laplace_operator.h:248)

    def __init__(self, handle):
        self._h = handle
        L = _lib()
        self._n_global = L.exadg_b200_n(handle)
        self._n_local = L.exadg_b200_local_size(handle)
        self.n_cells_owned = L.exadg_b200_n_cells_owned(handle)
        self.n_cells_ghost = L.exadg_b200_n_cells_ghost(handle)
        self.is_cartesian_path = L.exadg_b200_is_cartesian_path(handle)
        # kernels are launched on an operator-owned stream; callers synchronise through synchronize()

    # -- construction ------------------------------------------------------------------------------
    @classmethod
    def hypercube(cls, degree, n_subdivisions=1, n_refinements=0, mapping_degree=1, deformation=0.0, frequency=2,
                  boundary=(0,) * 6, ip_factor=1.0, rank=0, world=1, force_general=False):
        """Grid of applications/poisson/throughput (periodic box) or applications/poisson/sine."""
        _torch()
        d = _desc(degree, n_subdivisions, n_refinements, mapping_degree, deformation, frequency, boundary, ip_factor, rank, world, force_general)
        h = C.c_void_p()
        _check(_lib().exadg_b200_create_hypercube(C.byref(d), C.byref(h)))
        op = cls(h)
        op.degree = degree
        return op

    @classmethod
    def hypercube_helmholtz(cls, degree, n_components=3, scaling_factor_mass=1.0, viscosity=1.0, n_subdivisions=1, n_refinements=0, mapping_degree=1,
                            deformation=0.0, frequency=2, boundary=(0,) * 6, ip_factor=1.0, rank=0, world=1):
        """IncNS::MomentumOperator with the viscous term in Laplace formulation and constant viscosity (momentum_operator.cpp:376-480,
        viscous_operator.h:365-560): scaling_factor_mass * M + viscosity * A_SIPG on every component.  rank / world: partition
        of the grid as for LaplaceOperator.hypercube (whole cell blocks - all components - travel in the ghost import)."""
        from . import HelmholtzData
        _torch()
        d = _desc(degree, n_subdivisions, n_refinements, mapping_degree, deformation, frequency, boundary, ip_factor, rank, world, False)
        hd = HelmholtzData(n_components, scaling_factor_mass, viscosity)
        h = C.c_void_p()
        _check(_lib().exadg_b200_create_hypercube_helmholtz(C.byref(d), C.byref(hd), C.byref(h)))
        op = cls(h)
        op.degree = degree
        op.n_components = n_components
        return op

    def set_scaling_factor_mass(self, scaling_factor_mass):
        """MomentumOperator::set_scaling_factor_mass_operator: gamma_0 / dt of the current time step."""
        self.synchronize()
        _check(_lib().exadg_b200_set_scaling_factor_mass(self._h, float(scaling_factor_mass)))

    def inverse_mass_vmult(self, dst, src):
        """InverseMassOperator::apply (inverse_mass_operator.h): the InverseMassPreconditioner of the momentum equation."""
        self._order_after_torch()
        _check(_lib().exadg_b200_inverse_mass_vmult(self._h, _ptr(dst, self._n_local), _ptr(src, self._n_local)))
        self.synchronize()

    @classmethod
    def from_mesh(cls, degree, mapping_degree, mapping_points, neighbors, neighbor_face, boundary_type, n_cells_ghost=0,
                  ip_factor=1.0, force_general=False):
        """What a reference-side binding passes after extracting the mesh from dealii::MatrixFree."""
        from . import MeshDesc
        _torch()
        xm = np.ascontiguousarray(mapping_points, dtype=np.float64)
        nb = np.ascontiguousarray(neighbors, dtype=np.int32)
        nf = np.ascontiguousarray(neighbor_face, dtype=np.uint8)
        bt = np.ascontiguousarray(boundary_type, dtype=np.uint8)
        d = MeshDesc()
        d.degree, d.mapping_degree = degree, mapping_degree
        d.n_cells_owned, d.n_cells_ghost = nb.shape[0], n_cells_ghost
        d.mapping_points = xm.ctypes.data_as(C.POINTER(C.c_double))
        d.neighbors = nb.ctypes.data_as(C.POINTER(C.c_int32))
        d.neighbor_face = nf.ctypes.data_as(C.POINTER(C.c_uint8))
        d.boundary_type = bt.ctypes.data_as(C.POINTER(C.c_uint8))
        d.ip_factor, d.n_global_cells, d.global_cell_offset, d.force_general = ip_factor, nb.shape[0], 0, int(force_general)
        d.operator_is_singular = -1  # derived from the boundary types (OperatorBaseData::operator_is_singular)
        h = C.c_void_p()
        _check(_lib().exadg_b200_create(C.byref(d), C.byref(h)))
        op = cls(h)
        op.degree = degree
        return op

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib().exadg_b200_destroy(h)
            except Exception:
                pass

    # -- OperatorBase surface ----------------------------------------------------------------------
    def m(self):
        return self._n_global

    def n(self):
        return self._n_global

    def local_size(self):
        return self._n_local

    def el(self, i, j):
        # operator_base.cpp:216-222
        raise ExaDGError("Matrix-free does not allow for entry access")

    def initialize_dof_vector(self):
        torch = _torch()
        return torch.zeros(self._n_local, dtype=torch.float64, device="cuda")

    def _order_after_torch(self):
        """The operator launches on its own non-blocking stream: order it behind the work torch has queued on its current stream
        (the producers of src, the zero fill of dst) without a host synchronisation."""
        torch = _torch()
        _check(_lib().exadg_b200_wait_stream(self._h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def vmult(self, dst, src):
        self._order_after_torch()
        _check(_lib().exadg_b200_vmult(self._h, _ptr(dst, self._n_local), _ptr(src, self._n_local)))
        self.synchronize_with_torch()

    def vmult_add(self, dst, src):
        self._order_after_torch()
        _check(_lib().exadg_b200_vmult_add(self._h, _ptr(dst, self._n_local), _ptr(src, self._n_local)))
        self.synchronize_with_torch()

    apply = vmult          # operator_base.cpp:156-168: vmult forwards to apply on the matrix-free path
    apply_add = vmult_add
    vmult_interface_down = vmult        # operator_base.cpp:184-190
    vmult_add_interface_up = vmult_add  # operator_base.cpp:192-198

    def vmult_async(self, dst, src):
        """vmult without the trailing stream synchronisation (benchmark loop)."""
        self._order_after_torch()
        _check(_lib().exadg_b200_vmult(self._h, _ptr(dst, self._n_local), _ptr(src, self._n_local)))

    def vmult_host(self, dst_host, src_host):
        """dst = A src through pinned/pageable HOST tensors (H2D, vmult, D2H inside)."""
        for t in (dst_host, src_host):
            if t.is_cuda or t.numel() != self._n_local or not t.is_contiguous():
                raise ExaDGError("vmult_host expects contiguous host tensors of the local size")
        _check(_lib().exadg_b200_vmult_host(self._h, C.c_void_p(dst_host.data_ptr()), C.c_void_p(src_host.data_ptr())))

    def vmult_host_pipelined(self, dst_host, src_host):
        """Same result as vmult_host with upload, operator and download overlapped chunk by chunk inside the call (unpartitioned
        operators only; pinned host tensors)."""
        for t in (dst_host, src_host):
            if t.is_cuda or t.numel() != self._n_local or not t.is_contiguous():
                raise ExaDGError("vmult_host_pipelined expects contiguous host tensors of the local size")
        _check(_lib().exadg_b200_vmult_host_pipelined(self._h, C.c_void_p(dst_host.data_ptr()), C.c_void_p(src_host.data_ptr())))

    def set_host_pipeline_mode(self, mode):
        """Variant of vmult_host_pipelined: 0 automatic, 1 / "staged" chunk plan with a copy-engine download per chunk, 2 / "direct"
        piece-wise upload with the kernels storing dst straight into the pinned host tensor; returns the previous mode (an int)."""
        mode = {"auto": 0, "staged": 1, "direct": 2}.get(mode, mode)
        return _lib().exadg_b200_set_host_pipeline_mode(self._h, int(mode))

    def calculate_diagonal(self, diagonal):
        self._order_after_torch()
        _check(_lib().exadg_b200_calculate_diagonal(self._h, _ptr(diagonal, self._n_local)))
        self.synchronize_with_torch()

    def add_diagonal(self, diagonal):
        self._order_after_torch()
        _check(_lib().exadg_b200_add_diagonal(self._h, _ptr(diagonal, self._n_local)))
        self.synchronize_with_torch()

    def calculate_inverse_diagonal(self, diagonal):
        self._order_after_torch()
        _check(_lib().exadg_b200_calculate_inverse_diagonal(self._h, _ptr(diagonal, self._n_local)))
        self.synchronize_with_torch()

    def operator_is_singular(self):
        """operator_base.h:196: constants lie in the kernel (no Dirichlet face, e.g. the all-periodic throughput box)."""
        return bool(_lib().exadg_b200_operator_is_singular(self._h))

    # getters of operator_base.h:144-197 that have a meaning without deal.II objects
    def get_level(self):
        return 0xFFFFFFFF  # dealii::numbers::invalid_unsigned_int = the active level (operator_base.cpp:224-230)

    def get_dof_index(self):
        return 0

    def get_quad_index(self):
        return 0

    def set_time(self, time):
        self._time = float(time)  # the Laplace operator has no time-dependent coefficient; kept for the interface

    def get_time(self):
        return getattr(self, "_time", 0.0)

    def set_kernel_variant(self, variant):
        """Kernel of the affine fast path of THIS operator (see cartesian_kernel); -1 follows the process-wide default."""
        _check(_lib().exadg_b200_set_kernel_variant(self._h, int(variant)))

    def get_kernel_variant(self):
        return _lib().exadg_b200_get_kernel_variant(self._h)

    def is_empty_locally(self):
        return self.n_cells_owned == 0

    def subtract_mean_value(self, vec):
        """dealii::VectorTools::subtract_mean_value (global mean): consistency of the singular system (operator_is_singular)."""
        self._order_after_torch()
        _check(_lib().exadg_b200_subtract_mean_value(self._h, _ptr(vec, self._n_local)))
        self.synchronize()

    # -- inhomogeneous boundary data, right-hand side, error (operator_base.h:314-344; error_calculation.cpp) ------
    def boundary_quadrature_points(self):
        """(xyz [n_faces, (k+1)^2, 3], type [n_faces]) of the boundary faces of the owned cells."""
        n = C.c_int64()
        _check(_lib().exadg_b200_n_boundary_faces(self._h, C.byref(n)))
        nq2 = (self.degree + 1) ** 2
        xyz = np.zeros((n.value, nq2, 3))
        bt = np.zeros(n.value, dtype=np.uint8)
        _check(_lib().exadg_b200_boundary_quadrature_points(self._h, xyz.ctypes.data_as(C.c_void_p), bt.ctypes.data_as(C.c_void_p)))
        return xyz, bt

    def set_boundary_values(self, values):
        """g at the quadrature points of Dirichlet faces, h at those of Neumann faces ([n_faces, (k+1)^2])."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        _check(_lib().exadg_b200_set_boundary_values(self._h, v.ctypes.data_as(C.c_void_p)))

    def rhs(self, dst):
        self._order_after_torch()
        _check(_lib().exadg_b200_rhs(self._h, _ptr(dst, self._n_local)))
        self.synchronize()

    def rhs_add(self, dst):
        self._order_after_torch()
        _check(_lib().exadg_b200_rhs_add(self._h, _ptr(dst, self._n_local)))
        self.synchronize()

    def evaluate(self, dst, src):
        self._order_after_torch()
        _check(_lib().exadg_b200_evaluate(self._h, _ptr(dst, self._n_local), _ptr(src, self._n_local)))
        self.synchronize()

    def evaluate_add(self, dst, src):
        self._order_after_torch()
        _check(_lib().exadg_b200_evaluate_add(self._h, _ptr(dst, self._n_local), _ptr(src, self._n_local)))
        self.synchronize()

    def cell_quadrature_points(self, n_q_points_1d):
        xyz = np.zeros((self.n_cells_owned, n_q_points_1d ** 3, 3))
        _check(_lib().exadg_b200_cell_quadrature_points(self._h, n_q_points_1d, xyz.ctypes.data_as(C.c_void_p)))
        return xyz

    def integrate_source_add(self, dst, f_at_quadrature_points):
        """RHSOperator: dst += (f, phi_i), f at cell_quadrature_points(k + 1)."""
        f = np.ascontiguousarray(f_at_quadrature_points, dtype=np.float64)
        self._order_after_torch()
        _check(_lib().exadg_b200_integrate_source_add(self._h, _ptr(dst, self._n_local), f.ctypes.data_as(C.c_void_p)))
        self.synchronize()

    def l2_error(self, u, exact_at_quadrature_points, relative=True):
        """calculate_error with the L2 norm: exact solution at cell_quadrature_points(k + 3)."""
        ex = np.ascontiguousarray(exact_at_quadrature_points, dtype=np.float64)
        err = C.c_double()
        self._order_after_torch()
        _check(_lib().exadg_b200_l2_error(self._h, _ptr(u, self._n_local), ex.ctypes.data_as(C.c_void_p), int(relative), C.byref(err)))
        return err.value

    # -- stream handling ---------------------------------------------------------------------------
    def synchronize(self):
        _check(_lib().exadg_b200_synchronize(self._h))

    synchronize_with_torch = synchronize

    def use_torch_stream(self):
        """Launch on torch's current stream so torch.cuda.Event timing sees the kernels."""
        torch = _torch()
        _check(_lib().exadg_b200_set_stream(self._h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def kernel_launches(self):
        c = C.c_int64(0)
        _check(_lib().exadg_b200_kernel_launches(self._h, C.byref(c)))
        return c.value

    # -- halo plan (multi-GPU) ---------------------------------------------------------------------
    def halo_plan(self):
        L = _lib()
        peers = []
        for i in range(L.exadg_b200_halo_n_peers(self._h)):
            r, ns, rb, rc = C.c_int(), C.c_int64(), C.c_int64(), C.c_int64()
            _check(L.exadg_b200_halo_peer(self._h, i, C.byref(r), C.byref(ns), C.byref(rb), C.byref(rc)))
            cells = np.zeros(ns.value, dtype=np.int32)
            _check(L.exadg_b200_halo_send_list(self._h, i, cells.ctypes.data_as(C.POINTER(C.c_int32))))
            peers.append(dict(rank=r.value, send_cells=cells, recv_begin=rb.value, recv_count=rc.value))
        return peers

    def ghost_global_ids(self):
        ids = np.zeros(self.n_cells_ghost, dtype=np.int64)
        if self.n_cells_ghost:
            _check(_lib().exadg_b200_ghost_global_ids(self._h, ids.ctypes.data_as(C.POINTER(C.c_int64))))
        return ids

    def init_nccl(self, id_bytes):
        _check(_lib().exadg_b200_nccl_init(self._h, id_bytes))

    def enable_p2p(self, dist):
        """Switch the ghost import to NVLink peer-memory stores.  `dist` = initialised torch.distributed (any backend
        that can all-gather small tensors); only used to hand the IPC handles and receive offsets around."""
        torch = _torch()
        world = dist.get_world_size()
        handle = C.create_string_buffer(64)
        recv = (C.c_int64 * (world + 1))()
        _check(_lib().exadg_b200_p2p_export(self._h, handle, recv))
        mine = torch.cat([torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(torch.int64),
                          torch.tensor(list(recv), dtype=torch.int64)]).cuda()
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        handles = b"".join(bytes(g[:64].cpu().to(torch.uint8).numpy().tobytes()) for g in gathered)
        table = (C.c_int64 * (world * (world + 1)))(*[int(v) for g in gathered for v in g[64:].cpu().tolist()])
        _check(_lib().exadg_b200_p2p_connect(self._h, handles, table))
        dist.barrier()


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    _check(_lib().exadg_b200_nccl_unique_id(buf))
    return buf.raw


def _desc(degree, n_subdivisions, n_refinements, mapping_degree, deformation, frequency, boundary, ip_factor, rank, world, force_general):
    from . import HypercubeDesc
    d = HypercubeDesc()
    d.degree, d.n_subdivisions, d.n_refinements, d.mapping_degree = degree, n_subdivisions, n_refinements, mapping_degree
    d.deformation, d.frequency, d.ip_factor = float(deformation), frequency, float(ip_factor)
    for i in range(6):
        d.boundary[i] = int(boundary[i])
    d.rank, d.world, d.force_general = rank, world, int(force_general)
    return d


class JacobiPreconditioner:
    """jacobi_preconditioner.h:33-86: stores 1/diag(A); vmult is a pointwise scaling."""

    def __init__(self, op):
        self.op = op
        self.inverse_diagonal = op.initialize_dof_vector()
        self.update()

    def update(self):
        self.op.calculate_inverse_diagonal(self.inverse_diagonal)

    def vmult(self, dst, src):
        self.op._order_after_torch()
        _check(_lib().exadg_b200_jacobi_vmult(self.op._h, _ptr(dst), _ptr(src), _ptr(self.inverse_diagonal)))
        self.op.synchronize()


class ChebyshevSmoother:
    """chebyshev_smoother.h:35-175 with the defaults of multigrid_parameters.h:169-180."""

    def __init__(self, op, iterations=5, smoothing_range=20.0, iterations_eigenvalue_estimation=20):
        self.op = op
        h = C.c_void_p()
        op._order_after_torch()
        _check(_lib().exadg_b200_chebyshev_create(op._h, iterations, smoothing_range, iterations_eigenvalue_estimation, C.byref(h)))
        self._h = h
        v = [C.c_double() for _ in range(4)]
        _check(_lib().exadg_b200_chebyshev_get(h, *[C.byref(x) for x in v]))
        self.lambda_min_est, self.lambda_max_est, self.theta, self.delta = [x.value for x in v]

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib().exadg_b200_chebyshev_destroy(h)
            except Exception:
                pass

    def set_interval(self, theta, delta):
        _check(_lib().exadg_b200_chebyshev_set_interval(self._h, theta, delta))
        self.theta, self.delta = theta, delta

    def vmult(self, dst, src):
        self.op._order_after_torch()
        _check(_lib().exadg_b200_chebyshev_vmult(self._h, _ptr(dst), _ptr(src)))
        self.op.synchronize()

    def step(self, dst, src):
        self.op._order_after_torch()
        _check(_lib().exadg_b200_chebyshev_step(self._h, _ptr(dst), _ptr(src)))
        self.op.synchronize()


MG_TYPES = {"hMG": 0, "pMG": 1, "hpMG": 2, "phMG": 3}
P_SEQUENCES = {"GoToOne": 0, "DecreaseByOne": 1, "Bisect": 2}


def multigrid_levels(mg_type, p_sequence, degree, n_h_levels):
    """MultigridPreconditionerBase::initialize_levels (multigrid_preconditioner_base.cpp:97-323) for the DG-only multigrid types:
    list of (h_level, degree), coarse -> fine.  No CUDA call."""
    if mg_type not in MG_TYPES:
        raise ExaDGError("This multigrid type is not implemented! (%s; the c-transfer types need a continuous FE_Q operator)" % mg_type)
    n = C.c_int(0)
    _check(_lib().exadg_b200_multigrid_levels(MG_TYPES[mg_type], P_SEQUENCES[p_sequence], degree, n_h_levels, 0, C.byref(n), None, None))
    h = (C.c_int * n.value)()
    k = (C.c_int * n.value)()
    _check(_lib().exadg_b200_multigrid_levels(MG_TYPES[mg_type], P_SEQUENCES[p_sequence], degree, n_h_levels, n.value, C.byref(n), h, k))
    return [(h[i], k[i]) for i in range(n.value)]


class MultigridPreconditioner:
    """Poisson::MultigridPreconditioner / MultigridPreconditionerBase (I/poisson/preconditioners/multigrid_preconditioner.h:41-54,
    I/solvers_and_preconditioners/multigrid/multigrid_preconditioner_base.cpp) on DG levels: Chebyshev(point Jacobi) smoothers,
    CG + point Jacobi on the coarsest level, V-cycle of multigrid_algorithm.h:173-243."""

    def __init__(self, level_operators, smoother_iterations=5, smoothing_range=20.0, iterations_eigenvalue_estimation=20,
                 coarse_abs_tol=1e-12, coarse_rel_tol=1e-3, coarse_max_iter=10000):
        self.operators = list(level_operators)  # coarse -> fine; kept alive here
        arr = (C.c_void_p * len(self.operators))(*[op._h for op in self.operators])
        h = C.c_void_p()
        for op in self.operators:
            op._order_after_torch()
        _check(_lib().exadg_b200_multigrid_create(len(self.operators), arr, smoother_iterations, smoothing_range, iterations_eigenvalue_estimation,
                                                 coarse_abs_tol, coarse_rel_tol, coarse_max_iter, C.byref(h)))
        self._h = h
        self.op = self.operators[-1]

    @classmethod
    def hypercube(cls, fine_operator_args, mg_type="phMG", p_sequence="Bisect", fine_operator=None, **kw):
        """Level operators for a hypercube grid: fine_operator_args are the keyword arguments of LaplaceOperator.hypercube of the
        finest level; level (h, k) is the same grid with n_refinements = h and degree = k (what initialize_operator does per level,
        multigrid_preconditioner_base.cpp:593-640)."""
        a = dict(fine_operator_args)
        levels = multigrid_levels(mg_type, p_sequence, a["degree"], a.get("n_refinements", 0) + 1)
        ops = []
        for i, (h, k) in enumerate(levels):
            if i == len(levels) - 1 and fine_operator is not None:
                ops.append(fine_operator)
            else:
                b = dict(a)
                b["degree"], b["n_refinements"] = k, h
                ops.append(LaplaceOperator.hypercube(**b))
        mg = cls(ops, **kw)
        mg.levels = levels
        return mg

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib().exadg_b200_multigrid_destroy(h)
            except Exception:
                pass

    def vmult(self, dst, src):
        self.op._order_after_torch()
        _check(_lib().exadg_b200_multigrid_vmult(self._h, _ptr(dst, self.op.local_size()), _ptr(src, self.op.local_size())))
        self.op.synchronize()

    def info(self):
        n, ci, cy = C.c_int(0), C.c_int64(0), C.c_int64(0)
        _check(_lib().exadg_b200_multigrid_info(self._h, C.byref(n), C.byref(ci), C.byref(cy)))
        return {"n_levels": n.value, "coarse_iterations": ci.value, "cycles": cy.value}

    def smoother_interval(self, level):
        """(lambda_min_est, lambda_max_est, theta, delta) of the Chebyshev smoother of a level >= 1."""
        ch = C.c_void_p()
        _check(_lib().exadg_b200_multigrid_smoother(self._h, level, C.byref(ch)))
        v = [C.c_double() for _ in range(4)]
        _check(_lib().exadg_b200_chebyshev_get(ch, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def set_smoother_interval(self, level, theta, delta):
        ch = C.c_void_p()
        _check(_lib().exadg_b200_multigrid_smoother(self._h, level, C.byref(ch)))
        _check(_lib().exadg_b200_chebyshev_set_interval(ch, theta, delta))



@dataclass
class SolverData:
    """solver_data.h: SolverData(max_iter, abs_tol, rel_tol); Poisson defaults parameters.cpp:47."""
    max_iter: int = 10000
    abs_tol: float = 1e-20
    rel_tol: float = 1e-12


class KrylovSolverCG:
    """Krylov::KrylovSolver with linear_solver = CG (dealii::SolverCG + ReductionControl)."""

    def __init__(self, op, preconditioner=None, solver_data=None):
        self.op = op
        self.preconditioner = preconditioner
        self.solver_data = solver_data or SolverData()
        self.l2_0 = self.l2_n = self.rho = self.n_10 = 0.0
        self.n = 0
        self.residuals = None

    def solve(self, dst, rhs):
        """Returns the number of iterations (solver_control.last_step()); raises if not converged."""
        sd = self.solver_data
        kind, cheb = 0, None
        if isinstance(self.preconditioner, JacobiPreconditioner):
            kind = 1
        elif isinstance(self.preconditioner, ChebyshevSmoother):
            kind, cheb = 2, self.preconditioner._h
        elif isinstance(self.preconditioner, MultigridPreconditioner):
            kind = 3
        elif self.preconditioner is not None:
            raise ExaDGError("unsupported preconditioner")
        hist = np.zeros(sd.max_iter + 1)
        it = C.c_int(0)
        self.op._order_after_torch()
        if kind == 3:
            status = _lib().exadg_b200_cg_solve_multigrid(self.op._h, _ptr(dst), _ptr(rhs), self.preconditioner._h, sd.abs_tol, sd.rel_tol, sd.max_iter,
                                                         C.byref(it), hist.ctypes.data_as(C.POINTER(C.c_double)))
        else:
            status = _lib().exadg_b200_cg_solve(self.op._h, _ptr(dst), _ptr(rhs), kind, cheb, sd.abs_tol, sd.rel_tol, sd.max_iter,
                                               C.byref(it), hist.ctypes.data_as(C.POINTER(C.c_double)))
        self.n = it.value
        self.residuals = hist[: it.value + 1].copy()
        if status == 4:
            raise ExaDGError("SolverControl::NoConvergence after %d iterations" % it.value)
        _check(status)
        if not np.isfinite(self.residuals[-1]):
            raise ExaDGError("Last iteration step contained NaN or Inf values.")
        # do_compute_performance_metrics (iterative_solvers_dealii_wrapper.h:63-78)
        self.l2_0, self.l2_n = self.residuals[0], self.residuals[-1]
        if self.n > 0 and self.l2_0 > 0 and self.l2_n > 0:
            self.rho = (self.l2_n / self.l2_0) ** (1.0 / self.n)
            self.n_10 = -10.0 * np.log(10.0) / np.log(self.rho)
        return self.n


def fp64_peak():
    """Measured FP64 rates (TFLOP/s): register-resident DFMA chains and DMMA m8n8k4."""
    _torch()
    a, b = C.c_double(), C.c_double()
    _check(_lib().exadg_b200_fp64_peak(C.byref(a), C.byref(b)))
    return a.value, b.value


def cartesian_kernel(variant=-1):
    """Kernel of the affine fast path for degree 4: 0 pipelined 4-warp kernel, 1 (default) / 2 warp-specialised kernel with
    producer depth 8 / 12; -1 only queries.
    Process-wide tuning switch (no reference counterpart); returns the previous value."""
    return _lib().exadg_b200_cartesian_kernel(int(variant))


def host_pipeline_plan(n_subdivisions, n_refinements, cells_per_chunk=0, boundary=(0,) * 6, rank=0, world=1):
    """Host-only view (no GPU needed) of the chunk plan of vmult_host_pipelined on a hypercube grid: dict with n_chunks, upload_order,
    compute_order, ready_chunk (per chunk: the chunk whose upload makes it computable) and the modelled duration of one call in units
    of a one-direction transfer (2 = no overlap)."""
    L = _lib()
    d = _desc(1, n_subdivisions, n_refinements, 1, 0.0, 2, boundary, 1.0, rank, world, False)
    n, model = C.c_int32(), C.c_double()
    _check(L.exadg_b200_host_pipeline_plan(C.byref(d), int(cells_per_chunk), C.byref(n), None, None, None, C.byref(model)))
    out = {k: np.zeros(n.value, dtype=np.int32) for k in ("upload_order", "compute_order", "ready_chunk")}
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
    _check(L.exadg_b200_host_pipeline_plan(C.byref(d), int(cells_per_chunk), C.byref(n), p(out["upload_order"]), p(out["compute_order"]), p(out["ready_chunk"]),
                                           C.byref(model)))
    out["n_chunks"], out["model"] = n.value, model.value
    return out


def host_stream_plan(n_subdivisions, n_refinements, unit=24, cells_per_piece=0, boundary=(0,) * 6, rank=0, world=1):
    """Host-only view (no GPU needed) of the plan of the direct variant of vmult_host_pipelined on a hypercube grid: dict with n_steps,
    piece_begin (cell ranges in upload order), step_begin / units (the kernel units of `unit` cells applied behind every upload) and
    the modelled duration of one call in units of a one-direction transfer (1 = perfect overlap)."""
    L = _lib()
    d = _desc(1, n_subdivisions, n_refinements, 1, 0.0, 2, boundary, 1.0, rank, world, False)
    n, nu, model = C.c_int32(), C.c_int64(), C.c_double()
    _check(L.exadg_b200_host_stream_plan(C.byref(d), int(unit), C.c_int64(int(cells_per_piece)), C.byref(n), C.byref(nu), None, None, None, C.byref(model)))
    out = {"piece_begin": np.zeros(n.value + 1, dtype=np.int64), "step_begin": np.zeros(n.value + 1, dtype=np.int64), "units": np.zeros(nu.value, dtype=np.int32)}
    if n.value > 0:
        _check(L.exadg_b200_host_stream_plan(C.byref(d), int(unit), C.c_int64(int(cells_per_piece)), C.byref(n), C.byref(nu),
                                             out["piece_begin"].ctypes.data_as(C.POINTER(C.c_int64)), out["step_begin"].ctypes.data_as(C.POINTER(C.c_int64)),
                                             out["units"].ctypes.data_as(C.POINTER(C.c_int32)), C.byref(model)))
    out["n_steps"], out["unit"], out["model"] = n.value, int(unit), model.value
    return out


class PartitionPlan:
    """Host-only partition / halo plan of a hypercube grid (no GPU needed): what MatrixFree's
    Utilities::MPI::Partitioner holds for the ghost import of src."""

    def __init__(self, n_subdivisions, n_refinements, rank, world, boundary=(0,) * 6):
        L = _lib()
        d = _desc(1, n_subdivisions, n_refinements, 1, 0.0, 2, boundary, 1.0, rank, world, False)
        h = C.c_void_p()
        _check(L.exadg_b200_plan_create(C.byref(d), C.byref(h)))
        self._h = h
        no, ng, off, npeer = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        _check(L.exadg_b200_plan_sizes(h, C.byref(no), C.byref(ng), C.byref(off), C.byref(npeer)))
        self.n_owned, self.n_ghost, self.global_offset = no.value, ng.value, off.value
        self.neighbors = np.zeros((self.n_owned, 6), dtype=np.int32)
        self.ghost_global_ids = np.zeros(self.n_ghost, dtype=np.int64)
        _check(L.exadg_b200_plan_tables(h, self.neighbors.ctypes.data_as(C.POINTER(C.c_int32)),
                                        self.ghost_global_ids.ctypes.data_as(C.POINTER(C.c_int64))))
        self.peers = []
        for i in range(npeer.value):
            r, ns, rb, rc = C.c_int(), C.c_int64(), C.c_int64(), C.c_int64()
            _check(L.exadg_b200_plan_peer(h, i, C.byref(r), C.byref(ns), C.byref(rb), C.byref(rc), None))
            cells = np.zeros(ns.value, dtype=np.int32)
            _check(L.exadg_b200_plan_peer(h, i, None, None, None, None, cells.ctypes.data_as(C.POINTER(C.c_int32))))
            self.peers.append(dict(rank=r.value, send_cells=cells, recv_begin=rb.value, recv_count=rc.value))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib().exadg_b200_plan_destroy(h)
