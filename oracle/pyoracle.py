"""ctypes binding of the CPU oracle (oracle/oracle.h). TEST INFRASTRUCTURE ONLY.

Importers allowed by the project rules: tests/, __graft_entry__.smoke(), and bench.py's
cpu_baseline / --impl reference legs. The product package never imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NEOHOOKEAN, LINEAR_ELASTICITY, LAPLACIAN, MASS = 0, 1, 2, 3
SAINT_VENANT = 4
MOONEY_RIVLIN = 5
VISCOUS_DAMPING = 6
FIXED_COROTATIONAL = 7
MATERIAL_IDS = {"NeoHookean": NEOHOOKEAN, "LinearElasticity": LINEAR_ELASTICITY, "Laplacian": LAPLACIAN, "Mass": MASS, "SaintVenant": SAINT_VENANT,
                "MooneyRivlin": MOONEY_RIVLIN, "ViscousDamping": VISCOUS_DAMPING,
                "FixedCorotational": FIXED_COROTATIONAL}

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)


class _Desc(ctypes.Structure):
    _fields_ = [
        ("material", ctypes.c_int32), ("n_elements", ctypes.c_int32), ("n_loc", ctypes.c_int32),
        ("n_bases", ctypes.c_int32), ("n_qp", ctypes.c_int32), ("basis_order", ctypes.c_int32),
        ("node_lattice", _ip), ("conn", _ip), ("vertices", _dp), ("quad_points", _dp),
        ("quad_weights", _dp), ("ref_grads", _dp), ("lambda_", _dp), ("mu", _dp),
        ("use_cache", ctypes.c_int32), ("n_threads", ctypes.c_int32),
        ("ref_vals", _dp), ("density", _dp),
        ("geom_order", ctypes.c_int32), ("n_geom_loc", ctypes.c_int32), ("geom_lattice", _ip), ("geom_nodes", _dp),
        ("param3", _dp),
    ]


def build():
    """Compile liboracle.so (and oracle/_ref when the reference tree is mounted)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    if not os.path.exists(path) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(path)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    L = ctypes.CDLL(path)
    vp = ctypes.c_void_p
    L.oracle_create.restype = vp
    L.oracle_create.argtypes = [ctypes.POINTER(_Desc)]
    L.oracle_destroy.argtypes = [vp]
    L.oracle_size.argtypes = [vp]
    L.oracle_set_previous.restype = None
    L.oracle_set_previous.argtypes = [vp, _dp, ctypes.c_double]
    L.oracle_assemble_energy.restype = ctypes.c_double
    L.oracle_assemble_energy.argtypes = [vp, _dp]
    L.oracle_assemble_energy_per_element.argtypes = [vp, _dp, _dp]
    L.oracle_assemble_gradient.argtypes = [vp, _dp, _dp]
    L.oracle_assemble_hessian.restype = ctypes.c_int64
    L.oracle_assemble_hessian.argtypes = [vp, _dp, ctypes.c_int]
    L.oracle_assemble_linear.restype = ctypes.c_int64
    L.oracle_assemble_linear.argtypes = [vp]
    L.oracle_csc_nnz.restype = ctypes.c_int64
    L.oracle_csc_nnz.argtypes = [vp]
    for name, rt in (("oracle_csc_outer", _ip), ("oracle_csc_inner", _ip), ("oracle_csc_values", _dp)):
        getattr(L, name).restype = rt
        getattr(L, name).argtypes = [vp]
    for name in ("oracle_last_loop_seconds", "oracle_last_merge_seconds"):
        getattr(L, name).restype = ctypes.c_double
        getattr(L, name).argtypes = [vp]
    L.oracle_local_energy.restype = ctypes.c_double
    L.oracle_local_energy.argtypes = [vp, ctypes.c_int, _dp, ctypes.c_int]
    L.oracle_local_gradient.argtypes = [vp, ctypes.c_int, _dp, ctypes.c_int, _dp]
    L.oracle_local_hessian.argtypes = [vp, ctypes.c_int, _dp, ctypes.c_int, _dp]
    L.oracle_local_stiffness.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp]
    L.oracle_assembly_values.argtypes = [vp, ctypes.c_int, _dp, _dp, _dp]
    L.oracle_assembly_values.restype = None
    L.oracle_project_to_psd.argtypes = [ctypes.c_int, _dp]
    L.oracle_cache_new.restype = vp
    L.oracle_cache_new.argtypes = [ctypes.c_int]
    L.oracle_cache_copy.restype = vp
    L.oracle_cache_copy.argtypes = [vp]
    L.oracle_cache_free.argtypes = [vp]
    L.oracle_cache_add_value.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double]
    L.oracle_cache_prune.argtypes = [vp]
    L.oracle_cache_set_zero.argtypes = [vp]
    L.oracle_cache_add.argtypes = [vp, vp]
    L.oracle_cache_get_matrix.restype = ctypes.c_int64
    L.oracle_cache_get_matrix.argtypes = [vp]
    for name, rt in (("oracle_cache_outer", _ip), ("oracle_cache_inner", _ip), ("oracle_cache_values", _dp)):
        getattr(L, name).restype = rt
        getattr(L, name).argtypes = [vp]
    _LIB = L
    return L


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class CSC:
    """Column-compressed matrix as Eigen::SparseMatrix<double, ColMajor, int> stores it."""

    def __init__(self, n, outer, inner, values):
        self.n, self.outer, self.inner, self.values = n, outer, inner, values

    @property
    def nnz(self):
        return int(self.values.size)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.values, self.inner, self.outer), shape=(self.n, self.n))


class OracleProblem:
    """One assembler + FE space, as the reference's NLAssembler/LinearAssembler sees it."""

    def __init__(self, material, conn, vertices, n_bases, quad_points, quad_weights, ref_grads,
                 lam=None, mu=None, basis_order=1, node_lattice=None, use_cache=True, n_threads=1,
                 ref_vals=None, density=None, geom_order=0, geom_lattice=None, geom_nodes=None, param3=None):
        """MooneyRivlin: (c1, c2, k) = (lam, mu, param3). geom_order > 1 with geom_nodes [n_elements, n_geom_loc, 3] and geom_lattice [n_geom_loc, 3]: isoparametric geometry
        (curved elements); otherwise P1 geometry from `vertices`."""
        L = lib()
        self.material = MATERIAL_IDS[material] if isinstance(material, str) else int(material)
        self.conn = np.ascontiguousarray(conn, dtype=np.int32)
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        self.qp = np.ascontiguousarray(quad_points, dtype=np.float64)
        self.qw = np.ascontiguousarray(quad_weights, dtype=np.float64)
        self.rg = np.ascontiguousarray(ref_grads, dtype=np.float64)
        ne, nl = self.conn.shape
        self.n_elements, self.n_loc, self.n_bases = ne, nl, int(n_bases)
        self.lam = np.ascontiguousarray(np.broadcast_to(0.0 if lam is None else lam, (ne,)), dtype=np.float64)
        self.mu = np.ascontiguousarray(np.broadcast_to(0.0 if mu is None else mu, (ne,)), dtype=np.float64)
        self.lattice = None if node_lattice is None else np.ascontiguousarray(node_lattice, dtype=np.int32)
        d = _Desc()
        d.material, d.n_elements, d.n_loc, d.n_bases = self.material, ne, nl, self.n_bases
        d.n_qp, d.basis_order = int(self.qw.size), int(basis_order)
        d.node_lattice = _i(self.lattice) if self.lattice is not None else None
        d.conn, d.vertices = _i(self.conn), _d(self.vertices)
        d.quad_points, d.quad_weights, d.ref_grads = _d(self.qp), _d(self.qw), _d(self.rg)
        d.lambda_, d.mu = _d(self.lam), _d(self.mu)
        self.p3 = np.ascontiguousarray(np.broadcast_to(0.0 if param3 is None else param3, (ne,)), dtype=np.float64)
        d.param3 = _d(self.p3)
        d.use_cache, d.n_threads = int(bool(use_cache)), int(n_threads)
        if self.material == MATERIAL_IDS["Mass"]:
            assert ref_vals is not None, "Mass needs the basis values at the (mass) quadrature points"
            self.rv = np.ascontiguousarray(ref_vals, dtype=np.float64)
            assert self.rv.shape == (self.qw.size, nl)
            self.rho = np.ascontiguousarray(np.broadcast_to(1.0 if density is None else density, (ne,)), dtype=np.float64)
            d.ref_vals, d.density = _d(self.rv), _d(self.rho)
        if geom_order and geom_order > 1:
            self.gnodes = np.ascontiguousarray(geom_nodes, dtype=np.float64)
            self.glat = np.ascontiguousarray(geom_lattice, dtype=np.int32)
            assert self.gnodes.shape == (ne, self.glat.shape[0], 3) and quad_points is not None
            d.geom_order, d.n_geom_loc = int(geom_order), int(self.glat.shape[0])
            d.geom_lattice, d.geom_nodes = _i(self.glat), _d(self.gnodes)
        self._h = L.oracle_create(ctypes.byref(d))
        self.size = L.oracle_size(self._h)
        self.ndof = self.n_bases * self.size

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_destroy(self._h)
            self._h = None

    def _x(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        assert x.size == self.ndof
        return x

    def set_previous(self, x_prev, dt):
        """displacement_prev and dt of the NL entry points (ViscousDamping); x_prev None: no previous displacement"""
        lib().oracle_set_previous(self._h, None if x_prev is None else _d(self._x(x_prev)), float(dt))

    def assemble_energy(self, x):
        x = self._x(x)
        return float(lib().oracle_assemble_energy(self._h, _d(x)))

    def assemble_energy_per_element(self, x):
        x = self._x(x)
        out = np.zeros(self.n_elements)
        lib().oracle_assemble_energy_per_element(self._h, _d(x), _d(out))
        return out

    def assemble_gradient(self, x):
        x = self._x(x)
        rhs = np.zeros(self.ndof)
        lib().oracle_assemble_gradient(self._h, _d(x), _d(rhs))
        return rhs

    def _csc(self):
        L = lib()
        nnz = int(L.oracle_csc_nnz(self._h))
        outer = np.ctypeslib.as_array(L.oracle_csc_outer(self._h), shape=(self.ndof + 1,)).copy()
        inner = np.ctypeslib.as_array(L.oracle_csc_inner(self._h), shape=(nnz,)).copy() if nnz else np.zeros(0, np.int32)
        vals = np.ctypeslib.as_array(L.oracle_csc_values(self._h), shape=(nnz,)).copy() if nnz else np.zeros(0)
        return CSC(self.ndof, outer, inner, vals)

    def assemble_hessian(self, x, project_to_psd=False):
        x = self._x(x)
        lib().oracle_assemble_hessian(self._h, _d(x), int(bool(project_to_psd)))
        return self._csc()

    def assemble(self):
        lib().oracle_assemble_linear(self._h)
        return self._csc()

    def last_times(self):
        L = lib()
        return float(L.oracle_last_loop_seconds(self._h)), float(L.oracle_last_merge_seconds(self._h))

    def local_energy(self, e, x, autodiff=False):
        x = self._x(x)
        return float(lib().oracle_local_energy(self._h, int(e), _d(x), int(autodiff)))

    def local_gradient(self, e, x, autodiff=False):
        x = self._x(x)
        g = np.zeros(self.n_loc * self.size)
        lib().oracle_local_gradient(self._h, int(e), _d(x), int(autodiff), _d(g))
        return g

    def local_hessian(self, e, x, autodiff=False):
        x = self._x(x)
        n = self.n_loc * self.size
        h = np.zeros((n, n))
        lib().oracle_local_hessian(self._h, int(e), _d(x), int(autodiff), _d(h))
        return h

    def assembly_values(self, e):
        """(det[n_qp], jac_it[n_qp,3,3], grad_t_m[n_qp,n_loc,3]) of element e (ElementAssemblyValues.cpp:65-104)."""
        nq = int(self.qw.size)
        det, jit, gt = np.zeros(nq), np.zeros((nq, 3, 3)), np.zeros((nq, self.n_loc, 3))
        lib().oracle_assembly_values(self._h, int(e), _d(det), _d(jit), _d(gt))
        return det, jit, gt

    def local_stiffness(self, e, i, j):
        blk = np.zeros(self.size * self.size)
        lib().oracle_local_stiffness(self._h, int(e), int(i), int(j), _d(blk))
        return blk


def project_to_psd(a):
    a = np.array(a, dtype=np.float64, order="C")
    lib().oracle_project_to_psd(a.shape[0], _d(a))
    return a


def inertia(mass_csc, x, x_tilde):
    """InertiaForm::value_unweighted / first_derivative_unweighted (solver/forms/InertiaForm.cpp:17-29):
    (0.5 tmp^T M tmp, M tmp) with tmp = x - x_tilde; the Hessian is M itself (:31-34)."""
    tmp = np.asarray(x, dtype=np.float64).reshape(-1) - np.asarray(x_tilde, dtype=np.float64).reshape(-1)
    M = mass_csc.to_scipy()
    g = M @ tmp
    return 0.5 * float(tmp @ g), g


def project_gradient(grad, constrained, n_dofs=None):
    """BCLagrangianForm::project_gradient (solver/forms/lagrangian/BCLagrangianForm.cpp:149-155):
    keeps the entries of the sorted not_constraints_ list (:96-118), in order."""
    grad = np.asarray(grad, dtype=np.float64).reshape(-1)
    n = grad.size if n_dofs is None else n_dofs
    is_c = np.zeros(n, dtype=bool)
    is_c[np.asarray(constrained, dtype=np.int64)] = True
    not_constraints = np.flatnonzero(~is_c)
    out = np.empty(not_constraints.size)
    for i in range(not_constraints.size):  # the reference's loop, verbatim in meaning
        out[i] = grad[not_constraints[i]]
    return out


def project_hessian(csc, constrained):
    """BCLagrangianForm::project_hessian (BCLagrangianForm.cpp:167-213): two passes over the
    ColMajor CSC storage, dropping rows and columns whose indices are constrained dofs; the
    filtered rows stay ascending inside each kept column. Plain loops (small cases only)."""
    n = csc.n
    old_to_new = -np.ones(n, dtype=np.int64)
    is_c = np.zeros(n, dtype=bool)
    is_c[np.asarray(constrained, dtype=np.int64)] = True
    not_constraints = np.flatnonzero(~is_c)
    old_to_new[not_constraints] = np.arange(not_constraints.size)
    n_red = not_constraints.size
    total = 0
    for k in range(n):  # pass 1: count
        if old_to_new[k] < 0:
            continue
        for p in range(csc.outer[k], csc.outer[k + 1]):
            if old_to_new[csc.inner[p]] >= 0:
                total += 1
    outer = np.zeros(n_red + 1, dtype=np.int32)
    inner = np.zeros(total, dtype=np.int32)
    values = np.zeros(total)
    pos = 0
    for k in range(n):  # pass 2: fill column by column
        new_col = old_to_new[k]
        if new_col < 0:
            continue
        outer[new_col] = pos
        for p in range(csc.outer[k], csc.outer[k + 1]):
            new_row = old_to_new[csc.inner[p]]
            if new_row < 0:
                continue
            inner[pos] = new_row
            values[pos] = csc.values[p]
            pos += 1
    outer[n_red] = pos
    return CSC(n_red, outer, inner, values)


class Cache:
    """SparseMatrixCache restatement (the reference's tests/test_matrix.cpp "cache" test)."""

    def __init__(self, size=None, _h=None):
        self._h = _h if _h is not None else lib().oracle_cache_new(int(size))
        self.size = size

    def copy(self):
        c = Cache(_h=lib().oracle_cache_copy(self._h))
        c.size = self.size
        return c

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_cache_free(self._h)
            self._h = None

    def add_value(self, e, i, j, v):
        lib().oracle_cache_add_value(self._h, e, i, j, float(v))

    def prune(self):
        lib().oracle_cache_prune(self._h)

    def set_zero(self):
        lib().oracle_cache_set_zero(self._h)

    def add(self, other):
        """SparseMatrixCache::operator+= (merge of a thread copy, MatrixCache.cpp:289-322)."""
        lib().oracle_cache_add(self._h, other._h)

    def get_matrix(self):
        L = lib()
        nnz = int(L.oracle_cache_get_matrix(self._h))
        outer = np.ctypeslib.as_array(L.oracle_cache_outer(self._h), shape=(self.size + 1,)).copy()
        inner = np.ctypeslib.as_array(L.oracle_cache_inner(self._h), shape=(nnz,)).copy()
        vals = np.ctypeslib.as_array(L.oracle_cache_values(self._h), shape=(nnz,)).copy()
        return CSC(self.size, outer, inner, vals)


def problem_from_mesh(mesh, material, E=1e5, nu=0.3, order=None, rho=1.0, **kw):
    """Convenience: oracle problem on a polyfem_b200.mesh.TetMesh with a single material.
    Mass uses the mass quadrature order 2p of AssemblerUtils.cpp:204-211 unless `order` is given."""
    from polyfem_b200 import tables
    from polyfem_b200.mesh import lame_from_E_nu
    if material == "Mass":
        t = tables.reference_tables(mesh.p, order if order is not None else tables.quadrature_order(mesh.p, is_mass=True))
        return OracleProblem(material, mesh.conn, mesh.vertices, mesh.n_bases, t["points"], t["weights"], t["grad"],
                             basis_order=mesh.p, ref_vals=t["val"], density=rho, **kw)
    t = tables.reference_tables(mesh.p, order)
    lam, mu = lame_from_E_nu(E, nu)
    if material == "MooneyRivlin":  # c1, c2, k given explicitly (keywords), else a split of mu = 2 (c1 + c2), k = bulk modulus
        c1 = kw.pop("c1", 0.3 * mu)
        c2 = kw.pop("c2", 0.2 * mu)
        k = kw.pop("k", lam + 2.0 * mu / 3.0)
        lam, mu = c1, c2
        kw["param3"] = k
    if material == "ViscousDamping":  # (psi, phi) in the (lam, mu) slots
        lam, mu = kw.pop("psi", 30.0), kw.pop("phi", 20.0)
    return OracleProblem(material, mesh.conn, mesh.vertices, mesh.n_bases, t["points"], t["weights"], t["grad"],
                         lam=lam, mu=mu, basis_order=mesh.p,
                         node_lattice=np.array(tables.P_NODES_LATTICE[mesh.p], dtype=np.int32), **kw)
