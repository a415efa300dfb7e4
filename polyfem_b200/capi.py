"""ctypes binding of libpfa.so — the same C ABI (include/pfa.h) a PolyFEM-side shim binds.

Arrays may be numpy arrays (host memory) or torch CUDA tensors (device memory, passed by
`data_ptr()`); the library detects which with cudaPointerGetAttributes. There is no CPU
fallback: if libpfa.so is missing or no sm_100 device is present the calls raise.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PFA_LIB", os.path.join(_HERE, "libpfa.so"))  # PFA_LIB: kernel-variant experiments

PFA_OK = 0
PFA_ERR_INVALID, PFA_ERR_UNSUPPORTED, PFA_ERR_CUDA, PFA_ERR_NOMEM, PFA_ERR_NO_DEVICE = -1, -2, -3, -4, -5
NEOHOOKEAN, LINEAR_ELASTICITY, LAPLACIAN, MASS, SAINT_VENANT, MOONEY_RIVLIN, VISCOUS_DAMPING, FIXED_COROTATIONAL = 0, 1, 2, 3, 4, 5, 6, 7
MATERIAL_IDS = {"NeoHookean": NEOHOOKEAN, "LinearElasticity": LINEAR_ELASTICITY, "Laplacian": LAPLACIAN, "Mass": MASS, "SaintVenant": SAINT_VENANT,
                "MooneyRivlin": MOONEY_RIVLIN, "ViscousDamping": VISCOUS_DAMPING,
                "FixedCorotational": FIXED_COROTATIONAL}

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)

# every symbol include/pfa.h declares (tests check that the library exports all of them)
EXPORTS = [
    "pfa_create", "pfa_destroy", "pfa_last_error", "pfa_sizes", "pfa_pattern", "pfa_block_pattern", "pfa_pattern_device",
    "pfa_pattern_wide", "pfa_pattern_wide_device",
    "pfa_set_materials", "pfa_set_material_params", "pfa_set_previous", "pfa_energy", "pfa_energy_per_element", "pfa_gradient", "pfa_hessian",
    "pfa_linear_stiffness", "pfa_grad_hess", "pfa_grad_hess_weighted", "pfa_synchronize", "pfa_stream", "pfa_set_stream", "pfa_profile_enable",
    "pfa_profile_read", "pfa_launch_count", "pfa_setup_seconds",
    "pfa_is_step_valid", "pfa_set_constrained_dofs", "pfa_reduced_sizes", "pfa_reduced_pattern", "pfa_reduced_pattern_device",
    "pfa_project_gradient", "pfa_project_hessian", "pfa_grad_hess_reduced", "pfa_grad_hess_part",
    "pfa_symv", "pfa_inertia", "pfa_axpy",
    "pfa_partition_create", "pfa_partition_sizes", "pfa_partition_elements", "pfa_partition_conn", "pfa_partition_local_to_global",
    "pfa_partition_owned", "pfa_partition_destroy",
    "pfa_host_pattern_create", "pfa_host_pattern_arrays", "pfa_host_pattern_destroy", "pfa_host_element_order",
    "pfa_host_alloc", "pfa_host_free",
]


class MeshDesc(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_int32), ("material", ctypes.c_int32), ("n_elements", ctypes.c_int32),
        ("n_loc", ctypes.c_int32), ("n_bases", ctypes.c_int32), ("n_qp", ctypes.c_int32),
        ("conn", _ip), ("quad_weights", _dp), ("ref_grads", _dp),
        ("vertices", _dp), ("jac_it", _dp), ("da", _dp),
        ("lambda_", _dp), ("mu", _dp),
        ("material_stride", ctypes.c_int32), ("device", ctypes.c_int32), ("flags", ctypes.c_int32),
        ("n_ghost_elements", ctypes.c_int32), ("n_first_elements", ctypes.c_int32),
        ("ref_vals", _dp), ("density", _dp),
        ("owned_nodes", ctypes.POINTER(ctypes.c_uint8)),
        ("param3", _dp),
    ]


# pfa_mesh_desc.flags (include/pfa.h)
FLAG_KEEP_ELEMENT_ORDER = 1
FLAG_INKERNEL_ZERO = 2
FLAG_COLUMN_LANE = 4  # accepted, no effect: the owner-computes kernels are the default for NeoHookean P1/P2
FLAG_ROW_LANE = 8  # NeoHookean P1/P2: the round-1 row-lane reduction kernels (RED into a zero-filled values[])
FLAG_LARGE_INDEX = 32  # int64 pattern (POLYSOLVE_LARGE_INDEX), nnz may exceed 2^31
FLAG_GHOST_GEOMETRY = 16  # vertices / lambda / mu cover the ghost elements too (multi-GPU owner-computes form)


class PfaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pfa error {code}: {msg}")
        self.code = code


_LIB = None


def lib():
    """Loads libpfa.so; fails loudly when it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m polyfem_b200.build` (nvcc, sm_100a). "
                          "polyfem_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, c_int, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    L.pfa_create.argtypes = [ctypes.POINTER(MeshDesc), ctypes.POINTER(vp)]
    L.pfa_destroy.argtypes = [vp]
    L.pfa_destroy.restype = None
    L.pfa_last_error.argtypes = [vp]
    L.pfa_last_error.restype = ctypes.c_char_p
    L.pfa_sizes.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(i64), ctypes.POINTER(i64)]
    L.pfa_pattern.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(_ip), ctypes.POINTER(_ip)]
    L.pfa_block_pattern.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(_ip), ctypes.POINTER(_ip)]
    L.pfa_pattern_device.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp)]
    _lp = ctypes.POINTER(ctypes.c_int64)
    L.pfa_pattern_wide.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(_lp), ctypes.POINTER(_lp)]
    L.pfa_pattern_wide_device.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp)]
    L.pfa_set_materials.argtypes = [vp, vp, vp, ctypes.c_int32]
    L.pfa_set_material_params.argtypes = [vp, vp, vp, vp, ctypes.c_int32]
    L.pfa_set_previous.argtypes = [vp, vp, ctypes.c_double]
    L.pfa_energy.argtypes = [vp, vp, vp]
    L.pfa_energy_per_element.argtypes = [vp, vp, vp]
    L.pfa_gradient.argtypes = [vp, vp, vp]
    L.pfa_hessian.argtypes = [vp, vp, c_int, vp]
    L.pfa_linear_stiffness.argtypes = [vp, vp]
    L.pfa_grad_hess.argtypes = [vp, vp, c_int, vp, vp, vp]
    L.pfa_grad_hess_weighted.argtypes = [vp, vp, c_int, ctypes.c_double, vp, vp, vp]
    L.pfa_is_step_valid.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_int32), vp]
    L.pfa_set_constrained_dofs.argtypes = [vp, vp, i64]
    L.pfa_reduced_sizes.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    L.pfa_reduced_pattern.argtypes = [vp, ctypes.POINTER(_ip), ctypes.POINTER(_ip)]
    L.pfa_reduced_pattern_device.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp)]
    L.pfa_project_gradient.argtypes = [vp, vp, ctypes.c_double, vp]
    L.pfa_project_hessian.argtypes = [vp, vp, ctypes.c_double, vp]
    L.pfa_symv.argtypes = [vp, vp, vp, vp]
    L.pfa_inertia.argtypes = [vp, vp, vp, vp, vp, vp]
    L.pfa_axpy.argtypes = [vp, i64, ctypes.c_double, vp, vp]
    L.pfa_grad_hess_part.argtypes = [vp, vp, c_int, vp, vp, vp, c_int]
    L.pfa_grad_hess_reduced.argtypes = [vp, vp, c_int, ctypes.c_double, vp, vp, vp]
    L.pfa_synchronize.argtypes = [vp]
    L.pfa_stream.argtypes = [vp]
    L.pfa_stream.restype = vp
    L.pfa_set_stream.argtypes = [vp, vp]
    L.pfa_profile_enable.argtypes = [vp, c_int]
    L.pfa_profile_read.argtypes = [vp, c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_float)]
    L.pfa_launch_count.argtypes = [vp]
    L.pfa_launch_count.restype = i64
    L.pfa_setup_seconds.argtypes = [vp]
    L.pfa_setup_seconds.restype = ctypes.c_double
    i32p, u8p = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint8)
    L.pfa_partition_create.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _ip, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(vp)]
    L.pfa_partition_sizes.argtypes = [vp, i32p, i32p, i32p, i32p]
    for name, rt in (("pfa_partition_elements", i32p), ("pfa_partition_conn", i32p), ("pfa_partition_local_to_global", i32p), ("pfa_partition_owned", u8p)):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = rt
    L.pfa_partition_destroy.argtypes = [vp]
    L.pfa_partition_destroy.restype = None
    _LIB = L
    return L


def _ptr(a):
    """Raw address of a numpy array (host) / torch tensor (host or device) / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return ctypes.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return ctypes.c_void_p(a.data_ptr())
    raise TypeError(f"unsupported buffer type {type(a)}")


class Handle:
    """One pfa_handle: one mesh + material on one GPU."""

    def __init__(self, material, conn, n_bases, quad_weights, ref_grads, vertices=None, jac_it=None, da=None,
                 lam=None, mu=None, device=0, n_ghost_elements=0, flags=0, n_first_elements=0, ref_vals=None, density=None,
                 owned_nodes=None, param3=None):
        L = lib()
        self.material = MATERIAL_IDS[material] if isinstance(material, str) else int(material)
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        qw = np.ascontiguousarray(quad_weights, dtype=np.float64)
        ne, nl = conn.shape
        ne -= int(n_ghost_elements)  # trailing rows of conn are ghost elements
        # elements with geometry / material: with FLAG_GHOST_GEOMETRY the ghost elements as well
        ngeo = ne + (int(n_ghost_elements) if (int(flags) & FLAG_GHOST_GEOMETRY) else 0)
        nq = qw.size
        rg = None if ref_grads is None else np.ascontiguousarray(ref_grads, dtype=np.float64)
        assert rg is None or rg.shape == (nq, nl, 3), f"ref_grads must be [n_qp, n_loc, 3], got {rg.shape}"
        d = MeshDesc()
        d.struct_size = ctypes.sizeof(MeshDesc)
        d.material, d.n_elements, d.n_loc, d.n_bases, d.n_qp = self.material, ne, nl, int(n_bases), nq
        keep = [conn, qw, rg]
        d.conn = conn.ctypes.data_as(_ip)
        d.quad_weights = qw.ctypes.data_as(_dp)
        if rg is not None:
            d.ref_grads = rg.ctypes.data_as(_dp)
        if self.material == MASS:
            rv = np.ascontiguousarray(ref_vals, dtype=np.float64)
            assert rv.shape == (nq, nl), f"ref_vals must be [n_qp, n_loc], got {rv.shape}"
            rho = np.asarray(1.0 if density is None else density, dtype=np.float64)
            rho = np.ascontiguousarray(np.full(ne, float(rho)) if rho.ndim == 0 else rho)
            assert rho.size in (ne, ne * nq)
            d.ref_vals, d.density = rv.ctypes.data_as(_dp), rho.ctypes.data_as(_dp)
            keep += [rv, rho]
        if vertices is not None:
            v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(ngeo, 4, 3)
            d.vertices = v.ctypes.data_as(_dp)
            keep.append(v)
        if jac_it is not None:
            j = np.ascontiguousarray(jac_it, dtype=np.float64).reshape(ne, nq, 9)
            a = np.ascontiguousarray(da, dtype=np.float64).reshape(ne, nq)
            d.jac_it, d.da = j.ctypes.data_as(_dp), a.ctypes.data_as(_dp)
            keep += [j, a]
        stride = 1
        if self.material == MASS:
            stride = 1 if rho.size == ne else nq
        elif self.material != LAPLACIAN:
            lam = np.asarray(lam, dtype=np.float64)
            mu = np.asarray(mu, dtype=np.float64)
            if lam.ndim == 0:
                lam = np.full(ngeo, float(lam))
                mu = np.full(ngeo, float(mu))
            lam = np.ascontiguousarray(lam)
            mu = np.ascontiguousarray(mu)
            stride = 1 if lam.size == ngeo else nq
            assert lam.size == ngeo * stride and mu.size == ngeo * stride
            d.lambda_, d.mu = lam.ctypes.data_as(_dp), mu.ctypes.data_as(_dp)
            keep += [lam, mu]
            if param3 is not None:  # MooneyRivlin: (c1, c2, k) = (lam, mu, param3)
                p3 = np.asarray(param3, dtype=np.float64)
                p3 = np.ascontiguousarray(np.full(ngeo * stride, float(p3)) if p3.ndim == 0 else p3)
                assert p3.size == ngeo * stride
                d.param3 = p3.ctypes.data_as(_dp)
                keep.append(p3)
        d.material_stride, d.device, d.flags = stride, int(device), int(flags)
        d.n_ghost_elements = int(n_ghost_elements)
        d.n_first_elements = int(n_first_elements)
        if owned_nodes is not None:
            own = np.ascontiguousarray(owned_nodes, dtype=np.uint8)
            assert own.size == int(n_bases)
            d.owned_nodes = own.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
            keep.append(own)
        h = ctypes.c_void_p()
        rc = L.pfa_create(ctypes.byref(d), ctypes.byref(h))
        if rc != PFA_OK:
            raise PfaError(rc, L.pfa_last_error(None).decode())
        self._h = h
        self.device = int(device)
        self.n_elements, self.n_loc, self.n_bases, self.n_qp = ne, nl, int(n_bases), nq
        size, ndof, nnz = ctypes.c_int32(), ctypes.c_int64(), ctypes.c_int64()
        L.pfa_sizes(self._h, ctypes.byref(size), ctypes.byref(ndof), ctypes.byref(nnz))
        self.size, self.ndof, self.nnz = size.value, ndof.value, nnz.value

    def close(self):
        if getattr(self, "_h", None):
            lib().pfa_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise PfaError(rc, lib().pfa_last_error(self._h).decode())
        return rc

    # ---- pattern ----
    def pattern(self):
        """(outer[ndof+1], inner[nnz]) int32 host arrays, Eigen CSC layout."""
        L = lib()
        nnz = ctypes.c_int64()
        po, pi = _ip(), _ip()
        self._check(L.pfa_pattern(self._h, ctypes.byref(nnz), ctypes.byref(po), ctypes.byref(pi)))
        outer = np.ctypeslib.as_array(po, shape=(self.ndof + 1,)).copy()
        inner = np.ctypeslib.as_array(pi, shape=(nnz.value,)).copy()
        return outer, inner

    def pattern_wide(self):
        """(outer[ndof+1], inner[nnz]) int64 host arrays of a FLAG_LARGE_INDEX handle."""
        n = ctypes.c_int64()
        lp = ctypes.POINTER(ctypes.c_int64)
        po, pi = lp(), lp()
        self._check(lib().pfa_pattern_wide(self._h, ctypes.byref(n), ctypes.byref(po), ctypes.byref(pi)))
        return (np.ctypeslib.as_array(po, shape=(self.ndof + 1,)).copy(), np.ctypeslib.as_array(pi, shape=(n.value,)).copy())

    def block_pattern(self):
        """(adj_off[n_bases+1], adj[n_pairs]) int32 host arrays: node-block form of the pattern."""
        n = ctypes.c_int64()
        po, pa = _ip(), _ip()
        self._check(lib().pfa_block_pattern(self._h, ctypes.byref(n), ctypes.byref(po), ctypes.byref(pa)))
        return (np.ctypeslib.as_array(po, shape=(self.n_bases + 1,)).copy(),
                np.ctypeslib.as_array(pa, shape=(n.value,)).copy())

    def pattern_device_ptrs(self):
        po, pi = ctypes.c_void_p(), ctypes.c_void_p()
        self._check(lib().pfa_pattern_device(self._h, ctypes.byref(po), ctypes.byref(pi)))
        return po.value, pi.value

    def set_materials(self, lam, mu, stride=1, param3=None):
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        if param3 is not None:
            p3 = np.ascontiguousarray(param3, dtype=np.float64)
            self._check(lib().pfa_set_material_params(self._h, _ptr(lam), _ptr(mu), _ptr(p3), int(stride)))
            return
        self._check(lib().pfa_set_materials(self._h, _ptr(lam), _ptr(mu), int(stride)))

    def set_previous(self, x_prev, dt):
        """displacement_prev / dt of the NL virtuals (PFA_VISCOUS_DAMPING); x_prev None: no previous displacement"""
        if x_prev is None:
            self._check(lib().pfa_set_previous(self._h, None, float(dt)))
            return
        xp = np.ascontiguousarray(x_prev, dtype=np.float64).reshape(-1)
        assert xp.size == self.ndof
        self._check(lib().pfa_set_previous(self._h, _ptr(xp), float(dt)))

    # ---- host-array convenience wrappers (numpy in, numpy out) ----
    def energy(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        e = np.zeros(1)
        self._check(lib().pfa_energy(self._h, _ptr(x), _ptr(e)))
        return float(e[0])

    def energy_per_element(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        out = np.zeros(self.n_elements)
        self._check(lib().pfa_energy_per_element(self._h, _ptr(x), _ptr(out)))
        return out

    def gradient(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        g = np.zeros(self.ndof)
        self._check(lib().pfa_gradient(self._h, _ptr(x), _ptr(g)))
        return g

    def hessian(self, x, project_to_psd=False):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        v = np.zeros(self.nnz)
        self._check(lib().pfa_hessian(self._h, _ptr(x), int(bool(project_to_psd)), _ptr(v)))
        return v

    def linear_stiffness(self):
        v = np.zeros(self.nnz)
        self._check(lib().pfa_linear_stiffness(self._h, _ptr(v)))
        return v

    def grad_hess(self, x, project_to_psd=False):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        e, g, v = np.zeros(1), np.zeros(self.ndof), np.zeros(self.nnz)
        self._check(lib().pfa_grad_hess(self._h, _ptr(x), int(bool(project_to_psd)), _ptr(e), _ptr(g), _ptr(v)))
        return float(e[0]), g, v

    # ---- raw pointer entry (host numpy arrays or torch CUDA tensors, any may be None) ----
    def grad_hess_raw(self, x, energy=None, grad=None, values=None, project_to_psd=False):
        self._check(lib().pfa_grad_hess(self._h, _ptr(x), int(bool(project_to_psd)), _ptr(energy), _ptr(grad), _ptr(values)))

    def grad_hess_weighted_raw(self, x, weight, energy=None, grad=None, values=None, project_to_psd=False):
        """Every output times `weight` (Form::weight, e.g. dt^2 under implicit Euler)."""
        self._check(lib().pfa_grad_hess_weighted(self._h, _ptr(x), int(bool(project_to_psd)), float(weight), _ptr(energy), _ptr(grad), _ptr(values)))

    # ---- InertiaForm on a Mass handle ----
    def symv(self, values, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        values = np.ascontiguousarray(values, dtype=np.float64)
        y = np.zeros(self.ndof)
        self._check(lib().pfa_symv(self._h, _ptr(values), _ptr(x), _ptr(y)))
        return y

    def inertia(self, mass_values, x, x_tilde):
        """(0.5 d^T M d, M d), d = x - x_tilde (InertiaForm::value / first_derivative)."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        xt = np.ascontiguousarray(x_tilde, dtype=np.float64).reshape(-1)
        mass_values = np.ascontiguousarray(mass_values, dtype=np.float64)
        e, g = np.zeros(1), np.zeros(self.ndof)
        self._check(lib().pfa_inertia(self._h, _ptr(mass_values), _ptr(x), _ptr(xt), _ptr(e), _ptr(g)))
        return float(e[0]), g

    def inertia_raw(self, mass_values, x, x_tilde, energy, grad):
        self._check(lib().pfa_inertia(self._h, _ptr(mass_values), _ptr(x), _ptr(x_tilde), _ptr(energy), _ptr(grad)))

    def axpy(self, a, x, y):
        """y += a * x on device tensors of equal length."""
        self._check(lib().pfa_axpy(self._h, int(x.numel()), float(a), _ptr(x), _ptr(y)))

    def grad_hess_part_raw(self, x, energy, grad, values, part, project_to_psd=False):
        """part 1: clear outputs + elements [0, n_first_elements); part 2: add the rest (device tensors only)."""
        self._check(lib().pfa_grad_hess_part(self._h, _ptr(x), int(bool(project_to_psd)), _ptr(energy), _ptr(grad), _ptr(values), int(part)))

    def linear_stiffness_raw(self, values):
        self._check(lib().pfa_linear_stiffness(self._h, _ptr(values)))

    # ---- the step after the assembly: line-search validity, Dirichlet projection ----
    def is_step_valid(self, x, want_energy=True):
        """(valid, energy): ElasticForm::is_step_valid plus assemble_energy from the same pass.
        x: host numpy array or torch CUDA tensor."""
        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        valid = ctypes.c_int32(-1)
        e = np.zeros(1)
        self._check(lib().pfa_is_step_valid(self._h, _ptr(x), ctypes.byref(valid), _ptr(e) if want_energy else None))
        return bool(valid.value), (float(e[0]) if want_energy else None)

    def set_constrained_dofs(self, dofs):
        """BCLagrangianForm's boundary_nodes_: builds the reduced pattern and gather map on the device."""
        dofs = np.ascontiguousarray(dofs, dtype=np.int32).reshape(-1)
        self._check(lib().pfa_set_constrained_dofs(self._h, _ptr(dofs) if dofs.size else None, int(dofs.size)))
        nd, nz = ctypes.c_int64(), ctypes.c_int64()
        self._check(lib().pfa_reduced_sizes(self._h, ctypes.byref(nd), ctypes.byref(nz)))
        self.ndof_reduced, self.nnz_reduced = nd.value, nz.value

    def reduced_pattern(self):
        po, pi = _ip(), _ip()
        self._check(lib().pfa_reduced_pattern(self._h, ctypes.byref(po), ctypes.byref(pi)))
        outer = np.ctypeslib.as_array(po, shape=(self.ndof_reduced + 1,)).copy()
        inner = np.ctypeslib.as_array(pi, shape=(self.nnz_reduced,)).copy() if self.nnz_reduced else np.zeros(0, np.int32)
        return outer, inner

    def reduced_pattern_device_ptrs(self):
        po, pi = ctypes.c_void_p(), ctypes.c_void_p()
        self._check(lib().pfa_reduced_pattern_device(self._h, ctypes.byref(po), ctypes.byref(pi)))
        return po.value, pi.value

    def project_gradient(self, grad_full, scale=1.0, out=None):
        """numpy in -> numpy out; with `out` given both may be torch CUDA tensors (nothing leaves the device)."""
        if out is None:
            grad_full = np.ascontiguousarray(grad_full, dtype=np.float64).reshape(-1)
            out = np.zeros(self.ndof_reduced)
        self._check(lib().pfa_project_gradient(self._h, _ptr(grad_full), float(scale), _ptr(out)))
        return out

    def project_hessian(self, values_full, scale=1.0, out=None):
        if out is None:
            values_full = np.ascontiguousarray(values_full, dtype=np.float64).reshape(-1)
            out = np.zeros(self.nnz_reduced)
        self._check(lib().pfa_project_hessian(self._h, _ptr(values_full), float(scale), _ptr(out)))
        return out

    def grad_hess_reduced(self, x, scale=1.0, project_to_psd=False):
        """(energy, grad_reduced, values_reduced) numpy: fused assembly into the Dirichlet-reduced system."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        e, g, v = np.zeros(1), np.zeros(self.ndof_reduced), np.zeros(self.nnz_reduced)
        self._check(lib().pfa_grad_hess_reduced(self._h, _ptr(x), int(bool(project_to_psd)), float(scale), _ptr(e), _ptr(g), _ptr(v)))
        return float(e[0]), g, v

    def grad_hess_reduced_raw(self, x, scale=1.0, energy=None, grad=None, values=None, project_to_psd=False):
        self._check(lib().pfa_grad_hess_reduced(self._h, _ptr(x), int(bool(project_to_psd)), float(scale), _ptr(energy), _ptr(grad), _ptr(values)))

    def synchronize(self):
        self._check(lib().pfa_synchronize(self._h))

    def stream(self):
        return lib().pfa_stream(self._h)

    def set_stream(self, stream_ptr):
        """Launch on a caller-owned CUDA stream (e.g. torch.cuda.current_stream().cuda_stream)."""
        self._check(lib().pfa_set_stream(self._h, ctypes.c_void_p(stream_ptr)))

    def profile_enable(self, on=True):
        self._check(lib().pfa_profile_enable(self._h, int(bool(on))))

    def profile_read(self):
        cap = 8192
        names = (ctypes.c_char_p * cap)()
        ms = (ctypes.c_float * cap)()
        n = self._check(lib().pfa_profile_read(self._h, cap, names, ms))
        return [(names[i].decode(), float(ms[i])) for i in range(min(n, cap))]

    def launch_count(self):
        return int(lib().pfa_launch_count(self._h))

    def setup_seconds(self):
        return float(lib().pfa_setup_seconds(self._h))


def partition(conn, n_bases, world, rank):
    """pfa_partition_create (include/pfa.h): element partition for the multi-GPU owner-computes path. Returns a dict with
    `elements` (caller's element ids, own first, then ghost), `n_own`, `n_ghost`, `conn` (local node ids, [n_own + n_ghost, n_loc]),
    `l2g` (caller's node id per local node) and `owned` (uint8 per local node)."""
    L = lib()
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    ne, nl = conn.shape
    p = ctypes.c_void_p()
    rc = L.pfa_partition_create(ne, nl, int(n_bases), conn.ctypes.data_as(_ip), int(world), int(rank), ctypes.byref(p))
    if rc != PFA_OK:
        raise PfaError(rc, "pfa_partition_create failed")
    try:
        n_own, n_ghost, n_loc_b, n_owned = (ctypes.c_int32() for _ in range(4))
        L.pfa_partition_sizes(p, ctypes.byref(n_own), ctypes.byref(n_ghost), ctypes.byref(n_loc_b), ctypes.byref(n_owned))
        nt = n_own.value + n_ghost.value
        out = {
            "n_own": n_own.value, "n_ghost": n_ghost.value, "n_owned_bases": n_owned.value,
            "elements": np.ctypeslib.as_array(L.pfa_partition_elements(p), shape=(nt,)).copy(),
            "conn": np.ctypeslib.as_array(L.pfa_partition_conn(p), shape=(nt, nl)).copy(),
            "l2g": np.ctypeslib.as_array(L.pfa_partition_local_to_global(p), shape=(n_loc_b.value,)).copy(),
            "owned": np.ctypeslib.as_array(L.pfa_partition_owned(p), shape=(n_loc_b.value,)).copy(),
        }
    finally:
        L.pfa_partition_destroy(p)
    return out


def host_pattern(conn, n_bases):
    """pfa_host_pattern_create (include/pfa.h): the pattern builder of pfa_create without a device. Returns (adj_off[n_bases + 1],
    adj[n_pairs], slot[n_el, n_loc, n_loc]) as int32 arrays."""
    L = lib()
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    ne, nl = conn.shape
    p = ctypes.c_void_p()
    L.pfa_host_pattern_create.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _ip, ctypes.POINTER(ctypes.c_void_p)]
    rc = L.pfa_host_pattern_create(ne, nl, int(n_bases), conn.ctypes.data_as(_ip), ctypes.byref(p))
    if rc != PFA_OK:
        raise PfaError(rc, "pfa_host_pattern_create failed")
    try:
        n = ctypes.c_int64()
        po, pa, ps = _ip(), _ip(), _ip()
        L.pfa_host_pattern_arrays.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(_ip), ctypes.POINTER(_ip), ctypes.POINTER(_ip)]
        L.pfa_host_pattern_arrays(p, ctypes.byref(n), ctypes.byref(po), ctypes.byref(pa), ctypes.byref(ps))
        out = (np.ctypeslib.as_array(po, shape=(int(n_bases) + 1,)).copy(), np.ctypeslib.as_array(pa, shape=(n.value,)).copy(),
               np.ctypeslib.as_array(ps, shape=(ne, nl, nl)).copy())
    finally:
        L.pfa_host_pattern_destroy.argtypes = [ctypes.c_void_p]
        L.pfa_host_pattern_destroy(p)
    return out


def host_element_order(vertices):
    """pfa_host_element_order: perm[k] = caller's index of the k-th element of the internal (space-filling curve) order."""
    L = lib()
    v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 4, 3)
    perm = np.empty(v.shape[0], dtype=np.int32)
    L.pfa_host_element_order.argtypes = [ctypes.c_int32, _dp, _ip]
    rc = L.pfa_host_element_order(v.shape[0], v.ctypes.data_as(_dp), perm.ctypes.data_as(_ip))
    if rc != PFA_OK:
        raise PfaError(rc, "pfa_host_element_order failed")
    return perm

