/* CPU ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain C++17 (no Eigen, no TBB) restatement of PolyFEM's assembly hot path, used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
 * checker and the timed CPU baseline. The product (libpfa.so) never links or calls this.
 *
 * Parity status ("pinned" / "unpinned"), see DESIGN.md §Oracle:
 *   - quadrature + basis tables: PINNED against the reference's own generated sources
 *     (oracle/_ref, tests/golden/ref_tables.npz).
 *   - SparseMatrixCache semantics: PINNED by the reference's self-contained known-answer test
 *     tests/test_matrix.cpp:202-249 ("cache"), restated in tests/test_oracle_cache.py, and bit for bit against the
 *     reference's own utils/MatrixCache.cpp compiled unmodified (oracle/_ref/libcacheref.so,
 *     tests/test_oracle_cache_vs_reference.py, tests/golden/cache_sequences.npz).
 *   - NeoHookean local energy, gradient and Hessian: PINNED against the reference's own function bodies
 *     (NeoHookeanElasticity.cpp:338-658 compiled verbatim from /root/reference against oracle/refmath/mini_eigen.hpp
 *     into oracle/_ref/libnhref.so; outputs committed as tests/golden/nh_local.npz; tests/test_oracle_reference_math.py).
 *   - LinearElasticity / Laplacian / Mass local blocks: PINNED the same way (LinearElasticity.cpp:29-63,
 *     Laplacian.cpp:13-26, Mass.cpp:5-23 compiled verbatim; golden blocks in tests/golden/nh_local.npz).
 *   - the global nonlinear loops NLAssembler::assemble_energy / _per_element / assemble_gradient / assemble_hessian with their
 *     per-thread storages (Assembler.cpp:16-94, 495-771): PINNED against the reference's own loop bodies compiled verbatim
 *     over its own local functions and its unmodified MatrixCache.cpp (oracle/_ref/libloopref.so, multi-element meshes,
 *     1 and 3 thread storages, tests/golden/nl_loops.npz, tests/test_oracle_loops_vs_reference.py).
 *   - the global linear loop LinearAssembler::assemble (Assembler.cpp:157-384): PINNED the same way over the reference's
 *     own LinearElasticity / Laplacian / Mass local functions (tests/golden/linear_loops.npz).
 *   - the NL path of LinearElasticity: energy PINNED against the reference's own compute_energy_aux<double>
 *     (LinearElasticity.cpp:103-132, tests/golden/le_energy.npz); gradient / Hessian are autodiff of that function in the
 *     reference: PINNED since the end of round 2 against the reference's own assemble_gradient / assemble_hessian over its own
 *     utils/autodiff.h (oracle/_ref/libsvref.so::ref_le_nl_local, tests/golden/le_nl_local.npz,
 *     tests/test_oracle_saint_venant_reference.py, 1e-13), besides the earlier property checks.
 *   - ViscousDamping energy / gradient / Hessian: PINNED against the reference's own function bodies (ViscousDamping.cpp:5-62,
 *     122-229, 297-342 compiled verbatim into oracle/_ref/libvdref.so; tests/golden/vd_local.npz,
 *     tests/test_oracle_viscous_reference.py, 1e-13).
 *   - FixedCorotational energy / gradient / Hessian: PINNED against the reference's own function bodies AND its own 3 x 3 SVD
 *     (FixedCorotational.cpp:293-436, 592-827, utils/svd.hpp:134-317 compiled verbatim into oracle/_ref/libfcref.so;
 *     tests/golden/fc_local.npz, tests/test_oracle_corotational_reference.py, 1e-12; the oracle's own SVD is a Jacobi one).
 *   - SaintVenant energy / gradient / Hessian: PINNED against the reference's own code path - compute_energy_aux<T>, stress<T, N>,
 *     strain_from_disp_grad (SaintVenantElasticity.cpp:9-20, 61-70, 206-266) differentiated by the reference's OWN utils/autodiff.h
 *     (included unmodified) through its own gradient_from_energy / hessian_from_energy (utils/ElasticityUtils.cpp:81-270), over its
 *     own ElasticityTensor::set_from_lambda_mu (MatParams.cpp:211-253): oracle/_ref/libsvref.so, tests/golden/sv_local.npz,
 *     tests/test_oracle_saint_venant_reference.py, 1e-13.
 *   - MooneyRivlin energy / gradient / Hessian: PINNED the same way - MooneyRivlinElasticity::elastic_energy<T>
 *     (MooneyRivlinElasticity.hpp:25-47) through the reference's own autodiff.h inside GenericElastic's compute_energy_aux,
 *     compute_gradient_from_stress, compute_hessian_from_stress (GenericElastic.hpp:92-212, 268-351): oracle/_ref/libmrref.so,
 *     tests/golden/mr_local.npz, tests/test_oracle_mooney_reference.py, 1e-13.
 *   - isoparametric (curved P2) geometry: det, jac_it, grad_t_m PINNED against the reference's own finalize3d on curved elements
 *     (oracle/_ref/libgeomref.so::ref_finalize3d_iso, tests/golden/geom_iso.npz, P2 and P3 bases over P2 geometry, 1e-13).
 *   - project_to_psd (ipc-toolkit, source absent): UNPINNED, documented behaviour restated.
 *   - Mass (assembler/Mass.cpp:5-23): pinned by closed forms (P1 local mass rho*V/20*(1+delta_ij), total mass,
 *     stored zeros off the block diagonal; tests/test_oracle_mass.py) - the reference has no unit test for it.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src/polyfem/).
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

	enum
	{
		ORACLE_NEOHOOKEAN = 0,
		ORACLE_LINEAR_ELASTICITY = 1,
		ORACLE_LAPLACIAN = 2,
		ORACLE_MASS = 3, /* assembler/Mass.cpp: LinearAssembler with rho * phi_i * phi_j on the block diagonal */
		ORACLE_SAINT_VENANT = 4, /* assembler/SaintVenantElasticity.cpp with the isotropic tensor of (lambda, mu) */
		/* assembler/MooneyRivlinElasticity.hpp through GenericElastic (full autodiff): parameters c1 = lambda[], c2 = mu[], k = param3[] */
		ORACLE_MOONEY_RIVLIN = 5,
		/* assembler/ViscousDamping.cpp (NLAssembler that depends on the previous displacement and dt): parameters
		 * (psi, phi) = (lambda[], mu[]); oracle_set_previous supplies x_prev and dt, without it every result is zero
		 * (the reference's `data.x_prev.size() != data.x.size()` branch) */
		ORACLE_VISCOUS_DAMPING = 6,
		/* assembler/FixedCorotational.cpp: psi = mu sum (sigma_i - 1)^2 + lambda/2 (prod sigma - 1)^2 on the signed singular
		 * values of F (utils/svd.hpp AutoFlipSVD); stress and 9 x 9 stiffness from the SVD */
		ORACLE_FIXED_COROTATIONAL = 7
	};

	typedef struct
	{
		int32_t material;   /* ORACLE_* */
		int32_t n_elements; /* bases.size() */
		int32_t n_loc;      /* local bases per element */
		int32_t n_bases;    /* global bases (n_basis argument of the assembler) */
		int32_t n_qp;
		int32_t basis_order;         /* p of the Lagrange basis (used only when use_cache == 0) */
		const int32_t *node_lattice; /* [n_loc][3] lattice coordinates (x,y,z)*p of the local nodes */
		const int32_t *conn;         /* [n_elements][n_loc] global basis index of local basis j */
		const double *vertices;      /* [n_elements][4][3] P1 geometric nodes (gbases) */
		const double *quad_points;   /* [n_qp][3] */
		const double *quad_weights;  /* [n_qp], already divided by 6 */
		const double *ref_grads;     /* [n_qp][n_loc][3] reference gradients (used when use_cache == 1) */
		const double *lambda;        /* [n_elements] */
		const double *mu;            /* [n_elements] */
		int32_t use_cache;           /* 1: AssemblyValsCache::init, 0: init_empty (recompute per call) */
		int32_t n_threads;           /* stand-in for TBB's thread count */
		const double *ref_vals;      /* [n_qp][n_loc] basis values basis_values[j].val(q) (ORACLE_MASS only) */
		const double *density;       /* [n_elements] rho (ORACLE_MASS only) */
		/* isoparametric geometry (optional): geometric bases of order geom_order > 1 with n_geom_loc nodes per element,
		 * geom_nodes[n_elements][n_geom_loc][3] (gbases[e].bases[j].global()[0].node) and their lattice coordinates;
		 * finalize3d sums over the geometric bases (ElementAssemblyValues.cpp:79-93). 0 / NULL: P1 geometry from `vertices`. */
		int32_t geom_order;
		int32_t n_geom_loc;
		const int32_t *geom_lattice; /* [n_geom_loc][3] */
		const double *geom_nodes;
		const double *param3; /* [n_elements] third material parameter (ORACLE_MOONEY_RIVLIN: k); NULL otherwise */
	} oracle_desc;

	typedef struct oracle_problem oracle_problem;

	oracle_problem *oracle_create(const oracle_desc *desc);
	void oracle_destroy(oracle_problem *p);
	int oracle_size(const oracle_problem *p); /* Assembler::size(): 3, or 1 for Laplacian */

	/* displacement_prev and dt of the NLAssembler entry points (Assembler.hpp:79-125), used by ORACLE_VISCOUS_DAMPING only;
	 * x_prev NULL: no previous displacement */
	void oracle_set_previous(oracle_problem *p, const double *x_prev, double dt);

	/* NLAssembler entry points (assembler/Assembler.cpp:495-771) */
	double oracle_assemble_energy(oracle_problem *p, const double *x);
	void oracle_assemble_energy_per_element(oracle_problem *p, const double *x, double *out);
	void oracle_assemble_gradient(oracle_problem *p, const double *x, double *rhs);
	/* Runs assemble_hessian through the problem's persistent SparseMatrixCache (first call builds
	 * the pattern + slot map, later calls use it). Returns nnz; read the result with oracle_csc_*. */
	int64_t oracle_assemble_hessian(oracle_problem *p, const double *x, int project_to_psd);
	/* LinearAssembler::assemble (assembler/Assembler.cpp:157-384). Returns nnz. */
	int64_t oracle_assemble_linear(oracle_problem *p);
	/* the last matrix produced by oracle_assemble_hessian / oracle_assemble_linear (CSC, int32) */
	int64_t oracle_csc_nnz(const oracle_problem *p);
	const int32_t *oracle_csc_outer(const oracle_problem *p);
	const int32_t *oracle_csc_inner(const oracle_problem *p);
	const double *oracle_csc_values(const oracle_problem *p);
	/* seconds spent in the element loop / in merge+get_matrix of the last matrix assembly */
	double oracle_last_loop_seconds(const oracle_problem *p);
	double oracle_last_merge_seconds(const oracle_problem *p);

	/* local (per element) quantities, for unit tests */
	double oracle_local_energy(oracle_problem *p, int e, const double *x, int autodiff);
	void oracle_local_gradient(oracle_problem *p, int e, const double *x, int autodiff, double *g /*[n_loc*size]*/);
	void oracle_local_hessian(oracle_problem *p, int e, const double *x, int autodiff, double *h /*[N*N] row-major*/);
	void oracle_local_stiffness(oracle_problem *p, int e, int i, int j, double *blk /*[size*size], index n*size+m*/);

	/* ElementAssemblyValues of element e as the assemblers see them (ElementAssemblyValues.cpp:65-104):
	 * det[n_qp], jac_it[n_qp][9] row-major, grad_t_m[n_qp][n_loc][3] */
	void oracle_assembly_values(oracle_problem *p, int e, double *det, double *jac_it, double *grad_t_m);

	/* ipc::project_to_psd restatement on a dense symmetric n x n matrix (row-major, in place) */
	void oracle_project_to_psd(int n, double *a);

	/* SparseMatrixCache restatement exposed for the reference's "cache" known-answer test */
	typedef struct oracle_cache oracle_cache;
	oracle_cache *oracle_cache_new(int size);
	oracle_cache *oracle_cache_copy(const oracle_cache *other); /* SparseMatrixCache(const MatrixCache &) */
	void oracle_cache_free(oracle_cache *c);
	void oracle_cache_add_value(oracle_cache *c, int e, int i, int j, double v);
	void oracle_cache_prune(oracle_cache *c);
	void oracle_cache_set_zero(oracle_cache *c);
	void oracle_cache_add(oracle_cache *dst, const oracle_cache *src); /* operator+= */
	int64_t oracle_cache_get_matrix(oracle_cache *c); /* returns nnz; then read with the getters */
	const int32_t *oracle_cache_outer(const oracle_cache *c);
	const int32_t *oracle_cache_inner(const oracle_cache *c);
	const double *oracle_cache_values(const oracle_cache *c);

#ifdef __cplusplus
}
#endif
