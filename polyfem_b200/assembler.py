"""Host-side mirror of the reference's assembler interface for the hot path.

PolyFEM's host code is C++ (the C++ shim is `host/assembler_shim.hpp`); this Python module
mirrors the same operator interface — same class names, method names, argument order and
error behaviour as `polyfem::assembler::Assembler` / `LinearAssembler` / `NLAssembler`
(reference: src/polyfem/assembler/Assembler.hpp:53-376) — on top of the C ABI, so that the
parity tests read like the reference's own tests (tests/test_assembler.cpp).

What stands in for the C++ argument types:
  std::vector<basis::ElementBases> bases / gbases  ->  `FESpace` (connectivity + P1 nodes)
  AssemblyValsCache                                ->  `AssemblyValsCache` (reference tables)
  Eigen::MatrixXd displacement                     ->  numpy array (host) or torch CUDA tensor
  StiffnessMatrix                                  ->  scipy.sparse.csc_matrix with the CSC
                                                       arrays returned by the library
  utils::MatrixCache                               ->  `MatrixCache` (holds the pattern; the
                                                       GPU path keeps its own slot map)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import capi, tables
from .mesh import TetMesh, lame_from_E_nu


def log_and_throw_error(msg: str):
    """utils/Logger.hpp:42-49 analogue."""
    raise RuntimeError(msg)


@dataclass
class FESpace:
    """What the assembler reads from `bases` / `gbases` (basis/ElementBases.hpp:16-114)."""
    conn: np.ndarray      # [n_el, n_loc] bases[e].bases[j].global()[0].index
    vertices: np.ndarray  # [n_el, 4, 3]  gbases[e].bases[k].global()[0].node
    n_bases: int
    p: int
    body_ids: np.ndarray | None = None  # mesh.get_body_ids()

    @staticmethod
    def from_mesh(mesh: TetMesh, body_ids=None) -> "FESpace":
        return FESpace(mesh.conn, mesh.vertices, mesh.n_bases, mesh.p, body_ids)

    def __len__(self):
        return self.conn.shape[0]


@dataclass
class AssemblyValsCache:
    """assembler/AssemblyValsCache.{hpp,cpp}: here just the reference-element tables; the
    per-element values (J^-T, det) live on the device in SoA form."""
    p: int
    order: int | None = None
    is_mass_: bool = False
    t: dict = field(default_factory=dict)

    def __post_init__(self):
        self.t = tables.reference_tables(self.p, self.order if self.order is not None
                                         else tables.quadrature_order(self.p, self.is_mass_))

    def is_mass(self):
        return self.is_mass_


class MatrixCache:
    """utils::MatrixCache stand-in: caller-owned scratch that persists across Hessian calls
    (ElasticForm.hpp:116). The GPU path stores the CSC pattern here after the first call."""

    def __init__(self):
        self.outer = None
        self.inner = None


class Assembler:
    """assembler/Assembler.hpp:53-376 (the five virtuals ElasticForm calls)."""
    _material = None

    def __init__(self, device: int = 0):
        self.device = device
        self.size_ = 3
        self._materials = {}       # body id -> (lambda, mu)
        self._handle = None
        self._handle_key = None
        self._lam_mu_arrays = None

    # -- Assembler API --
    def name(self) -> str:
        return self._material

    def size(self) -> int:
        return self.size_

    def set_size(self, size: int):
        if size != 3 and self._material != "Laplacian":
            log_and_throw_error(f"{self.name()}: only 3D (size 3) is supported by the B200 path")
        self.size_ = size

    def is_linear(self) -> bool:
        return self._material in ("LinearElasticity", "Laplacian", "Mass")

    def add_multimaterial(self, index: int, params: dict):
        """JSON material spec (MatParams.cpp:403-446): E|young + nu, or lambda + mu."""
        if "E" in params or "young" in params:
            E = float(params.get("E", params.get("young")))
            lam, mu = lame_from_E_nu(E, float(params["nu"]))
        elif "lambda" in params and "mu" in params:
            lam, mu = float(params["lambda"]), float(params["mu"])
        else:
            log_and_throw_error("material needs E/young + nu or lambda + mu")
        self._materials[int(params.get("id", index))] = (lam, mu)
        self._materials.setdefault("default", (lam, mu))
        self._lam_mu_arrays = None

    def set_materials(self, body_ids, body_params):
        """Assembler::set_materials (Assembler.cpp:97-151): one JSON or a list with ids."""
        if isinstance(body_params, dict):
            body_params = [body_params]
        for i, p in enumerate(body_params):
            self.add_multimaterial(i, p)

    def set_lame_arrays(self, lam, mu):
        """Per-element (or per element x qp) arrays, for spatially varying / time-dependent
        expressions evaluated on the host (LameParameters::lambda_mu, MatParams.cpp:368-401)."""
        self._lam_mu_arrays = (np.asarray(lam, dtype=np.float64), np.asarray(mu, dtype=np.float64))
        if self._handle is not None:
            lam, mu = self._lam_mu_arrays
            self._handle.set_materials(lam, mu, 1 if lam.size == self._handle.n_elements else self._handle.n_qp)

    # -- device state (methods of the reference are const: state is handle-held) --
    def _lame(self, bases: FESpace):
        if self._material == "Laplacian":
            return None, None
        if self._lam_mu_arrays is not None:
            return self._lam_mu_arrays
        if not self._materials:
            log_and_throw_error(f"{self.name()}: no material set")
        ne = len(bases)
        if bases.body_ids is None:
            lam, mu = self._materials["default"]
            return np.full(ne, lam), np.full(ne, mu)
        lam = np.empty(ne)
        mu = np.empty(ne)
        for e, b in enumerate(bases.body_ids):
            lam[e], mu[e] = self._materials.get(int(b), self._materials["default"])
        return lam, mu

    def _param3(self, bases: FESpace):
        """third per-element material parameter (MooneyRivlin: k); None for the Lame materials"""
        return None

    def _get_handle(self, n_basis, bases: FESpace, gbases: FESpace, cache: AssemblyValsCache) -> capi.Handle:
        key = (id(bases), id(cache), n_basis)
        if self._handle is not None and self._handle_key == key:
            return self._handle
        if cache.is_mass():
            log_and_throw_error("mass-matrix quadrature is not part of this path")
        if n_basis != bases.n_bases:
            log_and_throw_error("n_basis does not match the FE space")
        lam, mu = self._lame(bases)
        try:
            h = capi.Handle(self._material, bases.conn, n_basis, cache.t["weights"], cache.t["grad"],
                            vertices=gbases.vertices, lam=lam, mu=mu, device=self.device, param3=self._param3(bases))
        except capi.PfaError as ex:
            log_and_throw_error(str(ex))
        self.invalidate()
        self._handle, self._handle_key = h, key
        return h

    def invalidate(self):
        """bases changed (adaptive quadrature, ElasticForm.cpp:220-233): drop the precompute."""
        if self._handle is not None:
            self._handle.close()
        self._handle, self._handle_key = None, None

    @staticmethod
    def _as_matrix(h: capi.Handle, values, mat_cache: MatrixCache | None):
        import scipy.sparse as sp
        if mat_cache is not None and mat_cache.outer is not None and mat_cache.inner.size == h.nnz:
            outer, inner = mat_cache.outer, mat_cache.inner
        else:
            outer, inner = h.pattern()
            if mat_cache is not None:
                mat_cache.outer, mat_cache.inner = outer, inner
        return sp.csc_matrix((values, inner, outer), shape=(h.ndof, h.ndof))


class LinearAssembler(Assembler):
    """Assembler.hpp:206-234, LinearAssembler::assemble (Assembler.cpp:157-384)."""

    def assemble(self, is_volume, n_basis, bases, gbases, cache, t, is_mass=False):
        if not is_volume:
            log_and_throw_error("only volumetric (tet) meshes are supported")
        h = self._get_handle(n_basis, bases, gbases, cache)
        try:
            values = h.linear_stiffness()
        except capi.PfaError as ex:
            log_and_throw_error(str(ex))
        return self._as_matrix(h, values, None)


class Mass(LinearAssembler):
    """assembler/Mass.{hpp,cpp}: rho * phi_i * phi_j on the block diagonal, assembled with the mass
    quadrature (`AssemblyValsCache(p, is_mass_=True)`, AssemblerUtils.cpp:204-211); the matrix of
    InertiaForm. JSON parameter: "rho" (Mass::add_multimaterial, Mass.cpp:33-38)."""
    _material = "Mass"

    def add_multimaterial(self, index: int, params: dict):
        if "rho" not in params:
            log_and_throw_error("Mass needs rho")
        self._materials[int(params.get("id", index))] = float(params["rho"])
        self._materials.setdefault("default", float(params["rho"]))

    def _density(self, bases: FESpace):
        if not self._materials:
            log_and_throw_error("Mass: no density set")
        if bases.body_ids is None:
            return np.full(len(bases), self._materials["default"])
        return np.array([self._materials.get(int(b), self._materials["default"]) for b in bases.body_ids])

    def _get_handle(self, n_basis, bases, gbases, cache):
        key = (id(bases), id(cache), n_basis)
        if self._handle is not None and self._handle_key == key:
            return self._handle
        if not cache.is_mass():
            log_and_throw_error("Mass::assemble needs the mass quadrature (AssemblyValsCache with is_mass)")
        try:
            h = capi.Handle("Mass", bases.conn, n_basis, cache.t["weights"], None, vertices=gbases.vertices,
                            ref_vals=cache.t["val"], density=self._density(bases), device=self.device)
        except capi.PfaError as ex:
            log_and_throw_error(str(ex))
        self.invalidate()
        self._handle, self._handle_key = h, key
        return h


class NLAssembler(Assembler):
    """Assembler.hpp:236-300, NLAssembler::assemble_* (Assembler.cpp:495-771)."""

    def _before_call(self, h, dt, displacement, displacement_prev):
        """materials that read dt / displacement_prev hand them to the handle here (ViscousDamping)"""

    def assemble_energy(self, is_volume, bases, gbases, cache, t, dt, displacement, displacement_prev):
        h = self._get_handle(bases.n_bases, bases, gbases, cache)
        self._before_call(h, dt, displacement, displacement_prev)
        try:
            return h.energy(displacement)
        except capi.PfaError as ex:
            log_and_throw_error(str(ex))

    def assemble_energy_per_element(self, is_volume, bases, gbases, cache, t, dt, displacement, displacement_prev):
        h = self._get_handle(bases.n_bases, bases, gbases, cache)
        self._before_call(h, dt, displacement, displacement_prev)
        try:
            return h.energy_per_element(displacement)
        except capi.PfaError as ex:
            log_and_throw_error(str(ex))

    def assemble_gradient(self, is_volume, n_basis, bases, gbases, cache, t, dt, displacement, displacement_prev):
        h = self._get_handle(n_basis, bases, gbases, cache)
        self._before_call(h, dt, displacement, displacement_prev)
        try:
            return h.gradient(displacement).reshape(-1, 1)
        except capi.PfaError as ex:
            log_and_throw_error(str(ex))

    def assemble_hessian(self, is_volume, n_basis, project_to_psd, bases, gbases, cache, t, dt,
                         displacement, displacement_prev, mat_cache: MatrixCache | None = None):
        h = self._get_handle(n_basis, bases, gbases, cache)
        self._before_call(h, dt, displacement, displacement_prev)
        try:
            values = h.hessian(displacement, project_to_psd)
        except capi.PfaError as ex:
            log_and_throw_error(str(ex))
        return self._as_matrix(h, values, mat_cache)

    def assemble_grad_hess(self, n_basis, bases, gbases, cache, displacement, project_to_psd=False,
                           mat_cache: MatrixCache | None = None):
        """Fused energy + gradient + Hessian of one Newton iteration (pfa_grad_hess) — not a
        reference virtual; what ElasticForm's value/first/second derivative calls amount to."""
        h = self._get_handle(n_basis, bases, gbases, cache)
        try:
            e, g, v = h.grad_hess(displacement, project_to_psd)
        except capi.PfaError as ex:
            log_and_throw_error(str(ex))
        return e, g.reshape(-1, 1), self._as_matrix(h, v, mat_cache)


class NeoHookeanElasticity(NLAssembler):
    """assembler/NeoHookeanElasticity.{hpp,cpp}; name() == "NeoHookean"."""
    _material = "NeoHookean"


class SaintVenantElasticity(NLAssembler):
    """assembler/SaintVenantElasticity.{hpp,cpp} with the isotropic elasticity tensor of (E, nu) / (lambda, mu)
    (MatParams.cpp:211-276); name() == "SaintVenant"."""
    _material = "SaintVenant"


class FixedCorotational(NLAssembler):
    """assembler/FixedCorotational.{hpp,cpp}; name() == "FixedCorotational"; Lame parameters like NeoHookean."""
    _material = "FixedCorotational"


class MooneyRivlinElasticity(NLAssembler):
    """assembler/MooneyRivlinElasticity.{hpp,cpp} (GenericElastic<MooneyRivlinElasticity>); name() == "MooneyRivlin".
    Material JSON: c1, c2, k (GenericMatParam, MooneyRivlinElasticity.cpp:5-15)."""
    _material = "MooneyRivlin"

    def add_multimaterial(self, index: int, params: dict):
        for key in ("c1", "c2", "k"):
            if key not in params:
                log_and_throw_error("MooneyRivlin material needs c1, c2 and k")
        val = (float(params["c1"]), float(params["c2"]), float(params["k"]))
        self._materials[int(params.get("id", index))] = val
        self._materials.setdefault("default", val)

    def _per_element(self, bases: FESpace, slot: int):
        if not self._materials:
            log_and_throw_error(f"{self.name()}: no material set")
        ne = len(bases)
        if bases.body_ids is None:
            return np.full(ne, self._materials["default"][slot])
        return np.array([self._materials.get(int(b), self._materials["default"])[slot] for b in bases.body_ids], dtype=np.float64)

    def _lame(self, bases: FESpace):
        return self._per_element(bases, 0), self._per_element(bases, 1)

    def _param3(self, bases: FESpace):
        return self._per_element(bases, 2)


class ViscousDamping(NLAssembler):
    """assembler/ViscousDamping.{hpp,cpp}; name() == "ViscousDamping". Material JSON: psi, phi (ViscousDamping.cpp:64-73).
    Reads displacement_prev and dt of the NL virtuals; a displacement_prev of another size means "no previous step" and
    gives zeros (ViscousDamping.cpp:125-126)."""
    _material = "ViscousDamping"

    def add_multimaterial(self, index: int, params: dict):
        old = self._materials.get("default", (0.0, 0.0))
        val = (float(params.get("psi", old[0])), float(params.get("phi", old[1])))
        self._materials[int(params.get("id", index))] = val
        self._materials["default"] = val
        self._lam_mu_arrays = None

    def set_params(self, psi: float, phi: float):
        self.add_multimaterial(0, {"psi": psi, "phi": phi})
        self.invalidate()

    def _before_call(self, h, dt, displacement, displacement_prev):
        prev = None if displacement_prev is None else np.asarray(displacement_prev)
        if prev is not None and prev.size != np.asarray(displacement).size:
            prev = None
        try:
            h.set_previous(prev, dt)
        except capi.PfaError as ex:
            log_and_throw_error(str(ex))


class LinearElasticity(NLAssembler, LinearAssembler):
    """assembler/LinearElasticity.{hpp,cpp}: linear `assemble` plus the NL energy / gradient /
    Hessian used when a linear material sits inside a nonlinear solve."""
    _material = "LinearElasticity"


class Laplacian(LinearAssembler):
    """assembler/Laplacian.{hpp,cpp}; scalar, size() == 1."""
    _material = "Laplacian"

    def __init__(self, device: int = 0):
        super().__init__(device)
        self.size_ = 1


def make_assembler(formulation: str, device: int = 0) -> Assembler:
    """AssemblerUtils::make_assembler (AssemblerUtils.cpp:55-122) for the hot-path names."""
    table = {"NeoHookean": NeoHookeanElasticity, "LinearElasticity": LinearElasticity, "Laplacian": Laplacian, "Mass": Mass,
             "SaintVenant": SaintVenantElasticity, "MooneyRivlin": MooneyRivlinElasticity, "ViscousDamping": ViscousDamping,
             "FixedCorotational": FixedCorotational}
    if formulation not in table:
        log_and_throw_error(f"Unsupported assembler on the B200 path: {formulation}")
    return table[formulation](device)
