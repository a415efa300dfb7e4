// Internal declarations shared by the C-ABI layer (pfa_api.cu), the host-side pattern /
// slot-map builder (pfa_pattern.cu) and the kernels (pfa_kernels.cu).
#pragma once
#include "../../include/pfa.h"

#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>

namespace pfa
{
	// ---- what the kernels see (device pointers, SoA as described in DESIGN.md §Data layout) ----
	struct DeviceMesh
	{
		int32_t material = 0;
		int32_t size = 3; // dofs per node: 3 (elasticity) or 1 (Laplacian)
		int32_t n_el = 0, n_loc = 0, n_bases = 0, n_qp = 0;
		int32_t n_first = 0;     // elements [0, n_first) can be assembled on their own (pfa_grad_hess_part)
		int32_t geom_per_qp = 0; // 0: affine, one J^-T/det per element; 1: one per (element, qp)
		int32_t mat_stride = 1;  // 1 or n_qp

		const int32_t *conn = nullptr;     // [n_el][n_loc]
		const double *jit = nullptr;       // [n_el][gq][9] row-major J^-T
		const double *detj = nullptr;      // [n_el][gq]   affine: det(J) ; per-qp: da = det*w
		const double *lambda = nullptr;    // [n_el][mat_stride]
		const double *mu = nullptr;        // [n_el][mat_stride]
		const double *param3 = nullptr;    // [n_el][mat_stride] (PFA_MOONEY_RIVLIN: k; lambda, mu hold c1, c2)
		const double *ref_vals = nullptr;  // [n_qp][n_loc] basis values (PFA_MASS)
		// [9][n_loc][n_loc] reference moment matrices S^{cd}_{ij} = sum_q w_q ghat_i[c](q) ghat_j[d](q) (linear
		// assemblers on affine elements, assemble_affine_linear_kernel)
		const double *ref_moments = nullptr;
		const double *ref_grads = nullptr; // [n_qp][n_loc][3]
		const double *qweights = nullptr;  // [n_qp]
		const double *ref_grads_host = nullptr; // host copy of ref_grads (owned by the handle): source of the __constant__ table
		int32_t p2_structured = 0;              // ref_grads has the structural zeros / equal components of the P2 tet basis

		// node-block CSR/CSC pattern (symmetric): adj_off[n_bases+1], adj[n_pairs] ascending
		const int32_t *adj_off = nullptr;
		const int32_t *adj = nullptr;
		int64_t n_pairs = 0;
		// slot map: pair index (into adj) of block (i,j) of element e: slot[e][i*n_loc+j],
		// where the pair is (row node g_i) inside the column list of node g_j
		const int32_t *slot = nullptr;
		// row-lane kernels (size 3): entry[e][i*n_loc+j] = 9*adj_off[g_j] + 3*k, k = position of g_i in
		// adj(g_j): values index of H[(g_i,0),(g_j,0)]; cstride[e][j] = 3*deg(g_j), the distance between
		// the three scalar columns of node g_j, in bits 0..27, and in bits 28..30 the mask of existing
		// components of node g_j (7 for the full matrix). When these are set, `slot` is not uploaded.
		const int32_t *entry = nullptr;
		const int32_t *cstride = nullptr;
		// in-kernel zero fill of values[] (row-lane kernels): the column blocks first touched by warp
		// batch b are the runs zruns[zoff[b] .. zoff[b+1]) = (first double, number of doubles); they are
		// cleared by batch b - kZeroLookahead, which then publishes zflag[b] = epoch
		const int32_t *zoff = nullptr;
		const int2 *zruns = nullptr;
		int32_t *zflag = nullptr;
		int32_t n_batches = 0;
		// elements are stored in an internal (spatially sorted, L2-friendly) order; elem_id[e] is the
		// caller's index of internal element e (nullptr = identity)
		const int32_t *elem_id = nullptr;
	};

	struct AssembleArgs
	{
		const double *x = nullptr;       // [ndof] device
		double *energy = nullptr;        // device scalar (accumulated, must be zeroed by caller)
		double *energy_per_el = nullptr; // [n_el]
		double *grad = nullptr;          // [ndof] (accumulated, zeroed by caller)
		double *values = nullptr;        // [nnz]  (accumulated, zeroed by caller)
		int project_to_psd = 0;
		int32_t e_begin = 0, e_end = 0; // (internal) element range of this launch
		int32_t batch_quota = 0;        // row-lane kernels: warp batches per warp before it retires (0 = persistent)
		// row-lane / psd kernels only: every output is multiplied by `scale` (Form weight); with
		// old_to_new set, gradient entries go to their Dirichlet-reduced position (or are dropped) and
		// DeviceMesh::entry / cstride must be the tables built for the reduced matrix
		double scale = 1.0;
		const int32_t *old_to_new = nullptr;
		int *work_counter = nullptr; // device int, zeroed before the launch (dynamic batch hand-out)
		int32_t epoch = 0;           // > 0: values[] is zero-filled inside the kernel (see DeviceMesh::zoff)
		const double *x_prev = nullptr; // [ndof] device, PFA_VISCOUS_DAMPING
		double inv_dt = 0.0;
	};

	// Owner-computes (column-lane) tables of a handle (pfa_collane2.h / pfa_collane2.cu), built at create time for NeoHookean
	// P1 / P2 handles on affine elements unless PFA_FLAG_ROW_LANE is given; all pointers are device memory
	struct ColumnLane2Tables
	{
		int32_t enabled = 0;
		int32_t n_chunks[2] = {0, 0}; // class 0: small strips, class 1: large strips (chunks of class 0 come first)
		int32_t rows_max[2] = {0, 0}; // strip rows of the two launches
		int32_t n_record_elements = 0; // own + ghost elements: every element incident to a scheduled node
		const int32_t *grp_info = nullptr;  // [G][5][4]: node, 9*adj_off, 3*deg, -
		const int32_t *grp_off = nullptr;   // [G+1] first step of each group
		const int32_t *grp_rows = nullptr;  // [G]
		const int32_t *chunk_off = nullptr; // [C+1] first group of each chunk
		const uint32_t *inc = nullptr;      // [total_steps][10][4]
		double *records = nullptr;          // [n_record_elements][n_qp*12 + 6]
		double *block_energy = nullptr;     // [ceil(n_record_elements / 128)] partial energy sums of the records kernel
		int *counters = nullptr;            // [2] chunk hand-out of the two launches
		const double *rg_padded = nullptr;  // [n_loc][n_qp][4] own-node reference gradients, padded rows
		int64_t n_steps[2] = {0, 0}; // steps of the two classes (their share of the work)
		int32_t p2z = 0;             // the reference table is the P2 basis on the symmetric 4-point rule (cl2::p2_rule_weights)
		double z4b = 0.0, zbeta = 0.0; // 4 zb, 4 (za - zb)
	};
	bool column_lane2_applies(int material, int n_loc, int n_qp);
	size_t column_lane2_record_doubles(int n_qp);
	// records kernel (+ energy sum) + one column kernel per strip class; writes every entry of values[] / grad[] of the
	// scheduled nodes exactly once (no zero fill needed) and stores a.energy (fixed summation order as well)
	cudaError_t launch_column_lane2(const DeviceMesh &m, const AssembleArgs &a, const ColumnLane2Tables &t, int sm_count, cudaStream_t st, int *launches);

	// kernel launchers (pfa_kernels.cu). Return cudaError_t of the launch.
	cudaError_t launch_geometry_precompute(const double *vertices_dev, int n_el, double *jit, double *detj, cudaStream_t st);
	cudaError_t launch_expand_inner(const DeviceMesh &m, int32_t *outer, int32_t *inner, cudaStream_t st);
	cudaError_t launch_expand_inner64(const DeviceMesh &m, int64_t *outer, int64_t *inner, cudaStream_t st);
	// fused per-element energy / gradient / Hessian with scatter (NLAssembler entry points);
	// `linear` selects LinearAssembler::assemble semantics (x ignored).
	cudaError_t launch_assemble(const DeviceMesh &m, const AssembleArgs &a, bool linear, int sm_count, cudaStream_t st, const char **kernel_name);
	bool assemble_supported(const DeviceMesh &m);
	// true when launch_assemble uses the row-lane kernels (which read entry/cstride instead of slot)
	bool rowlane_applies(int material, int n_loc, int n_qp);
	// exact structural zeros / equal components of the P2 tet basis gradients in a [n_qp][10][3] table
	bool p2_table_structured(const double *ref_grads, int n_loc, int n_qp);
	bool affine_linear_applies(const DeviceMesh &m);
	// S^{cd}_{ij} = sum_q w_q ghat_i[c](q) ghat_j[d](q), layout [c*3+d][i][j]
	void reference_moments(const double *ref_grads, const double *weights, int n_loc, int n_qp, std::vector<double> &out);
	// elements per warp batch of the row-lane kernel for this element type
	int rowlane_batch_elements(int n_loc, int n_qp);
#ifndef PFA_ZERO_LOOKAHEAD
#define PFA_ZERO_LOOKAHEAD 64
#endif
	constexpr int kZeroLookahead = PFA_ZERO_LOOKAHEAD; // batches between clearing a column block and its first use
	// zero-fill schedule of the row-lane kernels (host, once per mesh): zoff[n_batches+1], runs (start, length)
	void build_zero_schedule(const int32_t *conn, int n_el, int n_loc, int n_bases, const std::vector<int32_t> &adj_off, int size, int batch_elements,
							 std::vector<int32_t> &zoff, std::vector<int32_t> &zruns);

	// ---- Dirichlet projection and NaN scan (pfa_project.cu) ----
	cudaError_t exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, void **scratch, size_t *scratch_bytes, cudaStream_t st);
	// keep[d] = 1, then 0 for every listed dof; *bad = 1 when a listed dof is outside [0, ndof)
	cudaError_t launch_mark_constrained(const int32_t *dofs_dev, int64_t n, int32_t ndof, int32_t *keep, int *bad, cudaStream_t st);
	cudaError_t launch_finish_maps(const int32_t *keep, const int32_t *rank, int32_t ndof, int32_t *old_to_new, int32_t *not_constraints, cudaStream_t st);
	cudaError_t launch_count_kept(const int32_t *outer, const int32_t *inner, const int32_t *old_to_new, int32_t ndof, int32_t *col_count, cudaStream_t st);
	cudaError_t launch_fill_reduced(const int32_t *outer, const int32_t *inner, const int32_t *old_to_new, int32_t ndof, const int32_t *outer_red, int32_t *inner_red, int32_t *map, cudaStream_t st);
	// entry / cstride tables of the row-lane kernels for the Dirichlet-reduced matrix (pfa_grad_hess_reduced)
	cudaError_t launch_reduced_tables(const DeviceMesh &m, const int32_t *keep, const int32_t *old_to_new, const int32_t *outer_red,
									  int32_t *node_mask, int32_t *rowprefix, int32_t *cs_red, int32_t *cbase_red, int32_t *entry_red, int32_t *cstride_red, cudaStream_t st);
	// dst[t] = scale * src[map[t]]
	cudaError_t launch_gather_scale(const double *src, const int32_t *map, int64_t n, double scale, double *dst, int sm_count, cudaStream_t st);
	// y = A (x - x_tilde) for the symmetric CSC matrix (x_tilde may be NULL), *energy += 0.5 (x - x_tilde)^T y;
	// y and energy may be NULL
	cudaError_t launch_symv(const int32_t *outer, const int32_t *inner, const double *values, const double *x, const double *x_tilde, int32_t ndof,
							double *y, double *energy, cudaStream_t st);
	cudaError_t launch_axpy(int64_t n, double a, const double *x, double *y, int sm_count, cudaStream_t st);
	// *flag = 1 when any entry is NaN (flag must be zeroed by the caller)
	cudaError_t launch_any_nan(const double *v, int64_t n, int *flag, int sm_count, cudaStream_t st);

	// ---- host-side pattern + slot map (pfa_pattern.cu) ----
	struct HostPattern
	{
		std::vector<int32_t> adj_off; // [n_bases+1]
		std::vector<int32_t> adj;     // [n_pairs]
		std::vector<int32_t> slot;    // [n_el][n_loc*n_loc]
	};
	// throws std::runtime_error on invalid connectivity
	void build_pattern(const int32_t *conn, int n_el, int n_loc, int n_bases, HostPattern &out);
	// Morton order of the element centroids (vertices[e][4][3]): perm[internal] = caller's element index
	void spatial_element_order(const double *vertices, int n_el, std::vector<int32_t> &perm);
	// dst[i*stride+s] = src[perm[i]*stride+s] (device pointers)
	cudaError_t launch_gather_rows(const double *src, const int32_t *perm, int n, int stride, double *dst, cudaStream_t st);
} // namespace pfa
