/* pfa.h — C ABI of the B200 assembly path ("PolyFEM assembler", libpfa.so).
 *
 * This is the drop-in boundary for PolyFEM's per-element energy / gradient / Hessian
 * assembly. Each entry point replaces one virtual of
 * polyfem::assembler::Assembler (reference: src/polyfem/assembler/Assembler.hpp:68-125)
 * as ElasticForm calls it (src/polyfem/solver/forms/ElasticForm.cpp:287-328,411-418);
 * the C++ shim that overrides those virtuals and forwards here is host/assembler_shim.hpp,
 * the binding a PolyFEM maintainer adds is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C, no exceptions cross the boundary: every call returns PFA_OK (0) or a negative
 *    pfa_status; pfa_last_error() gives the message (the shim turns it into
 *    log_and_throw_error, utils/Logger.hpp:42-49). NaN/Inf results are NOT errors
 *    (ElasticForm.cpp:388-396 relies on NaN reaching the caller).
 *  - dofs are node-major: x[node*size + d] (NeoHookeanElasticity.cpp:352); size = 3 for
 *    NeoHookean / LinearElasticity, 1 for Laplacian.
 *  - matrices are returned as the values[] array of the CSC pattern reported by
 *    pfa_pattern(): column-major, int32 indices, inner indices ascending, all structural
 *    entries kept (explicit zeros included) — byte-identical to what
 *    SparseMatrixCache::get_matrix produces (utils/MatrixCache.cpp:134-228), so the shim can
 *    wrap it in Eigen::Map<const StiffnessMatrix>.
 *  - every `const double *x` / output pointer may be HOST or DEVICE memory (detected with
 *    cudaPointerGetAttributes). Calls are synchronous w.r.t. host pointers they fill; with
 *    device pointers the work is enqueued on the handle's stream and pfa_synchronize() (or
 *    any later host-pointer call) waits for it.
 *  - one handle = one GPU = one host thread at a time (the reference's callers are serial,
 *    SURVEY.md §8b "Threading").
 */
#ifndef PFA_H
#define PFA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

#define PFA_VERSION 1

	typedef enum
	{
		PFA_OK = 0,
		PFA_ERR_INVALID = -1,     /* bad argument / inconsistent description */
		PFA_ERR_UNSUPPORTED = -2, /* valid PolyFEM configuration this path does not cover */
		PFA_ERR_CUDA = -3,        /* CUDA runtime / launch failure */
		PFA_ERR_NOMEM = -4,       /* host or device allocation failed (std::bad_alloc analogue) */
		PFA_ERR_NO_DEVICE = -5    /* no usable sm_100 device: there is NO CPU fallback */
	} pfa_status;

	typedef enum
	{
		PFA_NEOHOOKEAN = 0,        /* assembler/NeoHookeanElasticity.cpp, name() == "NeoHookean" */
		PFA_LINEAR_ELASTICITY = 1, /* assembler/LinearElasticity.cpp,    name() == "LinearElasticity" */
		PFA_LAPLACIAN = 2,         /* assembler/Laplacian.cpp,           name() == "Laplacian" */
		PFA_MASS = 3,              /* assembler/Mass.cpp (LinearAssembler, size 3): rho phi_i phi_j on the block diagonal; the
		                            * mass matrix of InertiaForm. Needs ref_vals + density; quadrature is the
		                            * mass rule of order 2p (AssemblerUtils.cpp:204-211). */
		PFA_SAINT_VENANT = 4,      /* assembler/SaintVenantElasticity.cpp, name() == "SaintVenant", with the isotropic elasticity
		                            * tensor of (lambda, mu) (MatParams.cpp:211-253): an ElasticityNLAssembler whose energy is
		                            * differentiated by autodiff in the reference; here the closed forms
		                            * P = F S, S = 2 mu E + lambda tr(E) I and its tangent. Any order P1..P4 (generic kernel). */
		PFA_MOONEY_RIVLIN = 5,     /* assembler/MooneyRivlinElasticity.hpp, name() == "MooneyRivlin" (GenericElastic, autodiff in the
		                            * reference): psi = c1 (I1~ - 3) + c2 (I2~ - 3) + k/2 ln^2 J on the isochoric invariants. Parameters
		                            * (c1, c2, k) = (lambda[], mu[], param3[]). Here: the chain rule over (I1, I2, J) of F in closed
		                            * form (generic kernel, any order P1..P4). */
		PFA_VISCOUS_DAMPING = 6,   /* assembler/ViscousDamping.cpp, name() == "ViscousDamping": R = psi |dE/dt|^2 + phi/2 tr(dE/dt)^2,
		                            * dE/dt = sym(dF/dt^T F), dF/dt = (F - F_prev) / dt; (psi, phi) = (lambda[], mu[]). The previous
		                            * displacement and dt of the NLAssembler virtuals come through pfa_set_previous; until it is
		                            * called every result is zero (the reference's x_prev.size() != x.size() branch). Closed form:
		                            * with A = 2 F - F_prev the tangent is SaintVenant's with F -> A, mu -> psi / dt^2,
		                            * lambda -> phi / dt^2, S -> 2 (2 psi dE/dt + phi tr(dE/dt) I) / dt. */
		PFA_FIXED_COROTATIONAL = 7 /* assembler/FixedCorotational.cpp, name() == "FixedCorotational": psi = mu sum (sigma_i - 1)^2 +
		                            * lambda/2 (prod sigma - 1)^2 on the signed singular values of F (utils/svd.hpp); stress
		                            * lambda (J - 1) cof F + 2 mu (F - U V^T) and the 9 x 9 stiffness assembled from (U, sigma, V) per
		                            * quadrature point, contracted with the basis gradients by the whole warp. Inverted elements are
		                            * allowed (allow_inversion() is true in the reference). */
	} pfa_material;

	/* What the shim reads out of std::vector<basis::ElementBases> bases / gbases and the
	 * AssemblyValsCache once per mesh. All pointers are
	 * host memory and are copied during pfa_create. */
	typedef struct
	{
		int32_t struct_size; /* sizeof(pfa_mesh_desc), for ABI evolution */
		int32_t material;    /* pfa_material */
		int32_t n_elements;  /* bases.size() */
		int32_t n_loc;       /* bases[e].bases.size(): 4, 10, 20, 35 (P1..P4 tets) */
		int32_t n_bases;     /* n_basis argument of the assembler virtuals */
		int32_t n_qp;        /* quadrature points per element */
		/* conn[e*n_loc + j] = bases[e].bases[j].global()[0].index. Local2Global lists must have
		 * length 1 and weight 1 (conforming mesh, basis/Basis.hpp:21-38); otherwise the shim
		 * must not use this path. */
		const int32_t *conn;
		const double *quad_weights; /* [n_qp] vals.quadrature.weights (tet weights already /6) */
		const double *ref_grads;    /* [n_qp][n_loc][3] basis_values[j].grad.row(q) (reference element) */
		/* geometry, one of:
		 *  (a) affine (P1 gbases): vertices[e][4][3] = gbases[e].bases[k].global()[0].node, and
		 *      jac_it / da are NULL. J^-T and det are computed once here
		 *      (ElementAssemblyValues.cpp:65-104).
		 *  (b) general: jac_it[e][q][9] (row-major vals.jac_it[q]) and da[e][q] = det*weight. */
		const double *vertices;
		const double *jac_it;
		const double *da;
		/* Lame parameters, LameParameters::lambda_mu evaluated by the host (MatParams.cpp:368-401):
		 * material_stride == 1: [n_elements]; == n_qp: [n_elements][n_qp]. Ignored for Laplacian. */
		const double *lambda;
		const double *mu;
		int32_t material_stride;
		int32_t device; /* CUDA device ordinal */
		int32_t flags;  /* 0, or PFA_FLAG_* */
		/* Multi-GPU element partition: conn may carry n_ghost_elements extra rows
		 * after the n_elements computed ones. Ghost elements (elements of other ranks touching a
		 * node this rank owns) only widen the sparsity pattern so that owned columns have their
		 * full global row set; they are never evaluated and need no geometry/material entries. */
		int32_t n_ghost_elements;
		/* The first n_first_elements rows of conn form a group that pfa_grad_hess_part can assemble on
		 * its own (multi-GPU: the elements touching nodes owned by another rank, so that the interface
		 * exchange runs under the assembly of the rest). The internal re-ordering keeps the two groups
		 * apart. 0 = no split. */
		int32_t n_first_elements;
		/* PFA_MASS only: basis values [n_qp][n_loc] = basis_values[j].val(q) and the density
		 * (Mass::density_, evaluated by the host like lambda / mu) with the layout material_stride says.
		 * ref_grads, lambda and mu may be NULL for PFA_MASS. */
		const double *ref_vals;
		const double *density;
		/* Multi-GPU, owner-computes form (NeoHookean P1/P2 on affine elements; SURVEY.md §8e "device list": one process and one
		 * handle per GPU, the partition comes from pfa_partition_create): owned_nodes[n_bases] != 0 marks the nodes whose CSC
		 * columns and gradient entries this handle produces. Together with PFA_FLAG_GHOST_GEOMETRY - vertices, lambda and mu
		 * then cover the n_ghost_elements rows of conn as well - pfa_grad_hess / pfa_hessian write the FINISHED columns and
		 * gradient entries of the owned nodes (all incident elements are present on this rank) and leave the rest of
		 * values[] / grad[] untouched: no interface exchange, only the scalar energy (own elements) is summed over the ranks
		 * by the caller. NULL = every node is owned. */
		const uint8_t *owned_nodes;
		/* third material parameter, laid out like lambda / mu (PFA_MOONEY_RIVLIN: k; NULL for the other materials) */
		const double *param3;
	} pfa_mesh_desc;

/* pfa_mesh_desc.flags: keep the caller's element order internally (default: elements are
 * re-sorted along a space-filling curve once per mesh; results are identical either way) */
#define PFA_FLAG_KEEP_ELEMENT_ORDER 1
/* experimental: the row-lane kernels clear values[] themselves, block by block, a few warp
 * batches ahead of the scatter, instead of a cudaMemsetAsync before the launch (measured slower
 * on B200 in round 1: the cleared lines do not stay in L2 until their first RED) */
#define PFA_FLAG_INKERNEL_ZERO 2
/* NeoHookean P1/P2 on affine tets: pfa_grad_hess / pfa_hessian of the full (unreduced, unprojected) system run the
 * owner-computes (column-lane) kernels: the energy and every entry of values[] and of the gradient are summed in a fixed
 * order and written exactly once - bitwise reproducible results, no atomics, no zero fill. Costs a schedule of 16 bytes
 * per (element, local node) and a record buffer of 432 bytes per P2 element (144 per P1 element). This is the DEFAULT
 * since round 2 (the flag is accepted and has no effect); PFA_FLAG_ROW_LANE (or PFA_ROW_LANE=1 in the environment)
 * selects the round-1 row-lane reduction kernel (red.global.add.f64 into a zero-filled values[]) instead. */
#define PFA_FLAG_COLUMN_LANE 4
#define PFA_FLAG_ROW_LANE 8
/* vertices (and lambda, mu, density) have n_elements + n_ghost_elements entries: the ghost elements carry geometry and
 * material so that their records can be evaluated for the columns of owned nodes (see pfa_mesh_desc.owned_nodes) */
#define PFA_FLAG_GHOST_GEOMETRY 16
/* Large matrices (utils/Types.hpp:21-25: with POLYSOLVE_LARGE_INDEX StiffnessMatrix uses std::ptrdiff_t indices): the pattern is
 * handed out as int64 arrays by pfa_pattern_wide / pfa_pattern_wide_device and nnz may exceed 2^31 (BASELINE cfg 4, LinearElasticity P4
 * n = 32: 2.5 G nnz). Without the flag pfa_create refuses such a mesh, as Eigen's int indices would. With the flag the int32
 * accessors (pfa_pattern*, the Dirichlet projection, pfa_symv / pfa_inertia) report PFA_ERR_UNSUPPORTED; the NeoHookean P1/P2
 * kernels keep int32 entry tables and are limited to nnz < 2^31 either way. */
#define PFA_FLAG_LARGE_INDEX 32

	typedef struct pfa_handle pfa_handle;

	/* Builds the device-resident SoA precompute, the CSC pattern and the slot map. One-off per
	 * mesh (replaces AssemblyValsCache::init + the first-call pattern build of
	 * SparseMatrixCache, MatrixCache.cpp:88-100,134-213). */
	int pfa_create(const pfa_mesh_desc *desc, pfa_handle **out);

	/* Element partition of a mesh over `world` GPUs for the owner-computes path (host only, no device needed; replaces the
	 * per-thread element ranges of maybe_parallel_for, utils/MaybeParallelFor.tpp:18-68, and SURVEY.md §8e): contiguous element
	 * blocks in the caller's order; a node is owned by the rank of the FIRST element that touches it; the cuts balance the
	 * (element, node) incidences of the owned nodes, which is the work of the column kernels. Rank `rank` gets its own
	 * elements followed by its ghost elements (elements of other ranks touching a node it owns), renumbered locally in
	 * first-touch order. The arrays are owned by the partition object. */
	typedef struct pfa_partition pfa_partition;
	int pfa_partition_create(int32_t n_elements, int32_t n_loc, int32_t n_bases, const int32_t *conn, int32_t world, int32_t rank, pfa_partition **out);
	/* n_own / n_ghost elements, local nodes, owned local nodes */
	int pfa_partition_sizes(const pfa_partition *p, int32_t *n_own_elements, int32_t *n_ghost_elements, int32_t *n_local_bases, int32_t *n_owned_bases);
	const int32_t *pfa_partition_elements(const pfa_partition *p); /* [n_own + n_ghost] caller's element ids, own first */
	const int32_t *pfa_partition_conn(const pfa_partition *p);     /* [n_own + n_ghost][n_loc] local node ids */
	const int32_t *pfa_partition_local_to_global(const pfa_partition *p);      /* [n_local_bases] caller's node id of each local node */
	const uint8_t *pfa_partition_owned(const pfa_partition *p);    /* [n_local_bases] 1 = owned by this rank */
	void pfa_partition_destroy(pfa_partition *p);
	/* Host only, no device needed: the sparsity pattern pfa_create derives from a connectivity (it replaces the first-call path of
	 * SparseMatrixCache, utils/MatrixCache.cpp:88-213), in node-block form - adj_off[n_bases + 1], adj[n_pairs] as pfa_block_pattern
	 * describes them - and the slot map slot[e][i][j] = index into adj of (row node conn[e][i]) in the column list of node
	 * conn[e][j]. For a symbolic factorisation that starts before a GPU is attached, and for checking the pattern builder on a
	 * machine without one. The arrays are owned by the object. */
	typedef struct pfa_host_pattern pfa_host_pattern;
	int pfa_host_pattern_create(int32_t n_elements, int32_t n_loc, int32_t n_bases, const int32_t *conn, pfa_host_pattern **out);
	int pfa_host_pattern_arrays(const pfa_host_pattern *p, int64_t *n_pairs, const int32_t **adj_off, const int32_t **adj, const int32_t **slot);
	void pfa_host_pattern_destroy(pfa_host_pattern *p);
	/* Host only: the internal element order pfa_create uses for affine meshes without PFA_FLAG_KEEP_ELEMENT_ORDER (space-filling
	 * curve over the element centroids; vertices[e][4][3]): perm[k] = caller's index of the k-th internal element. */
	int pfa_host_element_order(int32_t n_elements, const double *vertices, int32_t *perm);
	void pfa_destroy(pfa_handle *h);
	/* message of the last failing call on this handle (or of pfa_create when h == NULL) */
	const char *pfa_last_error(const pfa_handle *h);

	/* size() of the assembler, ndof = n_bases*size, nnz of the CSC pattern */
	int pfa_sizes(const pfa_handle *h, int32_t *size, int64_t *ndof, int64_t *nnz);
	/* CSC pattern, host copies owned by the handle: outer[ndof+1], inner[nnz]
	 * (== mat.outerIndexPtr() / innerIndexPtr() of the reference result). */
	int pfa_pattern(pfa_handle *h, int64_t *nnz, const int32_t **outer, const int32_t **inner);
	/* node-block form of the same (symmetric) pattern: column node b lists its row nodes
	 * adj[adj_off[b] .. adj_off[b+1]) ascending; values index of (row (a,m), col (b,n)) with a at
	 * position k is size*size*adj_off[b] + n*size*deg(b) + size*k + m. Host arrays owned by the handle. */
	int pfa_block_pattern(pfa_handle *h, int64_t *n_pairs, const int32_t **adj_off, const int32_t **adj);
	/* PFA_FLAG_LARGE_INDEX handles: the same pattern with 64-bit indices (Eigen::SparseMatrix<double, ColMajor, std::ptrdiff_t>);
	 * built on first use (16 bytes per nnz on the device, and on the host for pfa_pattern_wide) */
	int pfa_pattern_wide(pfa_handle *h, int64_t *nnz, const int64_t **outer, const int64_t **inner);
	int pfa_pattern_wide_device(pfa_handle *h, const int64_t **outer_dev, const int64_t **inner_dev);
	/* same arrays in device memory (for a GPU linear solver / the multi-GPU exchange) */
	int pfa_pattern_device(pfa_handle *h, const int32_t **outer_dev, const int32_t **inner_dev);

	/* Re-upload Lame parameters when t changes (Assembler::set_materials, Assembler.cpp:97-151). */
	int pfa_set_materials(pfa_handle *h, const double *lambda, const double *mu, int32_t material_stride);
	/* displacement_prev and dt of the NLAssembler virtuals (Assembler.hpp:79-125), read by PFA_VISCOUS_DAMPING only: x_prev[ndof]
	 * (host or device pointer) is copied; NULL = no previous displacement. */
	int pfa_set_previous(pfa_handle *h, const double *x_prev, double dt);
	/* the same for materials with three parameters (PFA_MOONEY_RIVLIN: c1, c2, k) */
	int pfa_set_material_params(pfa_handle *h, const double *p1, const double *p2, const double *p3, int32_t material_stride);

	/* NLAssembler::assemble_energy (Assembler.cpp:495-531) */
	int pfa_energy(pfa_handle *h, const double *x, double *energy);
	/* NLAssembler::assemble_energy_per_element (Assembler.cpp:533-572); out[n_elements] */
	int pfa_energy_per_element(pfa_handle *h, const double *x, double *out);
	/* NLAssembler::assemble_gradient (Assembler.cpp:574-643); grad[ndof] fully overwritten */
	int pfa_gradient(pfa_handle *h, const double *x, double *grad);
	/* NLAssembler::assemble_hessian (Assembler.cpp:645-771); values[nnz] fully overwritten.
	 * project_to_psd != 0 applies ipc::project_to_psd per element (Assembler.cpp:693-694).
	 * For LinearElasticity this is the constant stiffness (ElasticForm.cpp:316-320). */
	int pfa_hessian(pfa_handle *h, const double *x, int project_to_psd, double *values);
	/* LinearAssembler::assemble (Assembler.cpp:157-384) for LinearElasticity / Laplacian */
	int pfa_linear_stiffness(pfa_handle *h, double *values);
	/* Fused energy + gradient + Hessian of one Newton iteration (the benchmarked entry).
	 * Any of energy / grad / values may be NULL to skip that output. */
	int pfa_grad_hess(pfa_handle *h, const double *x, int project_to_psd, double *energy, double *grad, double *values);
	/* The same with every output multiplied by `weight`: Form::value / first_derivative / second_derivative return
	 * weight() * the unweighted quantity (solver/forms/Form.hpp:30-56); for implicit Euler the elastic form's weight is
	 * dt^2 (time_integrator/ImplicitEuler.cpp:28-31, solver/SolveData.cpp:491). NeoHookean P1/P2 tets (fused in the kernels). */
	int pfa_grad_hess_weighted(pfa_handle *h, const double *x, int project_to_psd, double weight, double *energy, double *grad, double *values);

	/* ---- the step right after the assembly ---- */

	/* ElasticForm::is_step_valid (solver/forms/ElasticForm.cpp:388-396): assembles the gradient at x into
	 * device scratch and reports *valid = 0 when any entry is NaN (an inverted element under the
	 * Discrete inversion check), 1 otherwise; 4 bytes come back instead of the gradient. When
	 * energy != NULL the same kernel pass also returns assemble_energy(x), which the line search asks
	 * for next (solver/NLProblem.cpp:565-590). NLAssembler materials only. */
	int pfa_is_step_valid(pfa_handle *h, const double *x, int32_t *valid, double *energy);

	/* Dirichlet projection, BCLagrangianForm (solver/forms/lagrangian/BCLagrangianForm.cpp). The list of
	 * constrained dofs (boundary_nodes_, any order, host or device pointer) fixes not_constraints_ /
	 * old_to_new_ (:96-118); the reduced CSC pattern and a gather map reduced nnz -> full nnz are
	 * built once on the device. n == 0 is allowed (identity projection). */
	int pfa_set_constrained_dofs(pfa_handle *h, const int32_t *dofs, int64_t n);
	int pfa_reduced_sizes(const pfa_handle *h, int64_t *ndof_reduced, int64_t *nnz_reduced);
	/* reduced CSC pattern, host copies owned by the handle: outer[ndof_reduced+1], inner[nnz_reduced]
	 * (== the matrix project_hessian returns, :167-213) */
	int pfa_reduced_pattern(pfa_handle *h, const int32_t **outer, const int32_t **inner);
	int pfa_reduced_pattern_device(pfa_handle *h, const int32_t **outer_dev, const int32_t **inner_dev);
	/* project_gradient (:149-155): grad_reduced[i] = scale * grad_full[not_constraints[i]] */
	int pfa_project_gradient(pfa_handle *h, const double *grad_full, double scale, double *grad_reduced);
	/* project_hessian (:167-213): values_reduced[k] = scale * values_full[map[k]]; scale carries
	 * Form::second_derivative's weight() / scale_ (solver/forms/Form.hpp:42-56). Host or device pointers. */
	int pfa_project_hessian(pfa_handle *h, const double *values_full, double scale, double *values_reduced);

	/* Fused form of pfa_grad_hess + pfa_project_gradient + pfa_project_hessian (NeoHookean P1/P2 tets):
	 * the kernels scatter straight into the reduced CSC values / reduced gradient (contributions to
	 * constrained rows and columns are dropped at the source) and multiply energy, gradient and
	 * values by `scale` (Form::value / first_derivative / second_derivative weight, Form.hpp:30-56).
	 * Neither the full matrix nor the gather pass exist in this mode. Outputs may be NULL. */
	int pfa_grad_hess_reduced(pfa_handle *h, const double *x, int project_to_psd, double scale, double *energy, double *grad_reduced, double *values_reduced);

	/* pfa_grad_hess in two launches (device pointers only): part = PFA_PART_FIRST clears the outputs and
	 * assembles elements [0, n_first_elements); PFA_PART_REST adds the remaining elements to the same
	 * outputs; PFA_PART_ALL is pfa_grad_hess. */
#define PFA_PART_ALL 0
#define PFA_PART_FIRST 1
#define PFA_PART_REST 2
	int pfa_grad_hess_part(pfa_handle *h, const double *x, int project_to_psd, double *energy, double *grad, double *values, int part);

	/* ---- InertiaForm (solver/forms/InertiaForm.cpp:17-34) on a PFA_MASS handle ----
	 * y = A x for the symmetric CSC matrix (pattern of the handle, `values` as returned by
	 * pfa_linear_stiffness / pfa_hessian); host or device pointers. */
	int pfa_symv(pfa_handle *h, const double *values, const double *x, double *y);
	/* value_unweighted and first_derivative_unweighted in one pass: d = x - x_tilde,
	 * grad = M d (may be NULL), *energy = 0.5 d^T M d (may be NULL). second_derivative is `mass_values`
	 * itself. x_tilde = x_prev + dt v_prev is the caller's (ImplicitEuler.cpp:13-16). */
	int pfa_inertia(pfa_handle *h, const double *mass_values, const double *x, const double *x_tilde, double *energy, double *grad);
	/* y[k] += a * x[k], k < n: e.g. Hessian of the time-stepping problem = dt^2-weighted elastic values
	 * + mass values (same pattern when both handles come from the same connectivity) */
	int pfa_axpy(pfa_handle *h, int64_t n, double a, const double *x, double *y);

	/* Page-locked host memory for the host-pointer mode: the device-to-host copy of values[] (5.5 GB at BASELINE cfg 3) runs at
	 * the PCIe rate only into pinned memory; ordinary (pageable) buffers work too, at a fraction of it. pfa_host_alloc returns
	 * NULL when there is no device or not enough lockable memory - the caller then falls back to ordinary memory (the shim's
	 * HostValues does). Free with pfa_host_free only. */
	void *pfa_host_alloc(size_t bytes);
	void pfa_host_free(void *p);

	/* waits for all work enqueued on the handle's stream */
	int pfa_synchronize(pfa_handle *h);
	/* cudaStream_t the handle launches on (as void*), so callers can order their own work */
	void *pfa_stream(pfa_handle *h);

	/* makes the handle launch on a caller-owned cudaStream_t (e.g. the framework's current
	 * stream, so the caller's events bracket this library's kernels). NULL = legacy stream. */
	int pfa_set_stream(pfa_handle *h, void *stream);

	/* Instrumentation for bench.py: when enabled every kernel launch / fill of the next calls
	 * is bracketed by CUDA events on the handle's stream; records accumulate across calls.
	 * pfa_profile_read synchronizes the stream, returns the number of accumulated records,
	 * writes up to `cap` (name, ms) of them and clears the list. */
	int pfa_profile_enable(pfa_handle *h, int on);
	int pfa_profile_read(pfa_handle *h, int cap, const char **names, float *ms);
	/* kernels of this library launched by this handle since creation (bench.py's gpu_launches;
	 * driver memsets are not counted) */
	int64_t pfa_launch_count(const pfa_handle *h);
	/* seconds spent in pfa_create building pattern + slot map + precompute (one-off setup) */
	double pfa_setup_seconds(const pfa_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* PFA_H */
