// C-ABI layer of libpfa.so (include/pfa.h): handle lifetime, host<->device staging,
// error translation. No CPU fallback: every compute entry point runs the CUDA kernels or
// fails with PFA_ERR_NO_DEVICE / PFA_ERR_CUDA.
#include "pfa_collane2.h"

#include <nvtx3/nvToolsExt.h> // header-only NVTX v3: ranges show up in Nsight Systems / ncu --nvtx (SURVEY.md §5 tracing)
#include "pfa_internal.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

using namespace pfa;

namespace
{
	thread_local std::string g_create_error;

	struct ProfRecord
	{
		const char *name;
		cudaEvent_t start, stop;
		bool stop_recorded;
	};
} // namespace

struct pfa_handle
{
	int device = 0;
	int sm_count = 148;
	cudaStream_t stream = nullptr;
	bool own_stream = true;
	DeviceMesh dm;
	int64_t ndof = 0, nnz = 0;

	// owned device memory
	std::vector<void *> owned;
	int32_t *d_outer = nullptr, *d_inner = nullptr;
	bool large_index = false; // PFA_FLAG_LARGE_INDEX: int64 pattern (built on first use), no int32 pattern
	int64_t *d_outer64 = nullptr, *d_inner64 = nullptr;
	std::vector<int64_t> h_outer64, h_inner64;
	double *d_lambda = nullptr, *d_mu = nullptr, *d_param3 = nullptr;
	double *d_x_prev = nullptr; // pfa_set_previous (PFA_VISCOUS_DAMPING)
	bool has_prev = false;
	double dt = 1.0;
	int32_t *d_elem_id = nullptr; // internal -> caller element index (nullptr: identity)
	double *s_mat = nullptr;      // staging for pfa_set_materials when elements are re-ordered
	// staging for host-pointer calls (allocated on first use)
	double *s_x = nullptr, *s_grad = nullptr, *s_values = nullptr, *s_epe = nullptr;
	double *d_energy = nullptr;
	int *d_counter = nullptr;
	int32_t epoch = 0; // in-kernel zero-fill generation (row-lane kernels)
	bool cl_partial = false;    // pfa_mesh_desc.owned_nodes was given: the column-lane path writes the owned nodes only
	int32_t n_geo_elements = 0; // elements with geometry / material on the device: n_el (+ ghost elements with PFA_FLAG_GHOST_GEOMETRY)
	ColumnLane2Tables cl; // owner-computes path (default for NeoHookean P1 / P2 on affine elements)

	// Dirichlet projection (pfa_set_constrained_dofs)
	bool has_constraints = false;
	int64_t ndof_red = 0, nnz_red = 0;
	int32_t *d_old_to_new = nullptr, *d_not_constraints = nullptr, *d_outer_red = nullptr, *d_inner_red = nullptr, *d_map = nullptr;
	int32_t *d_entry_red = nullptr, *d_cstride_red = nullptr; // row-lane tables of the reduced matrix (pfa_grad_hess_reduced)
	std::vector<void *> owned_proj; // freed when the constraint set changes
	std::vector<int32_t> h_outer_red, h_inner_red;
	double *s_vec_in = nullptr, *s_vec_out = nullptr, *s_val_out = nullptr; // staging for host-pointer projection calls
	int *d_flag = nullptr;
	void *scan_scratch = nullptr;
	size_t scan_scratch_bytes = 0;

	std::vector<int32_t> h_outer, h_inner;
	std::vector<int32_t> h_adj_off, h_adj;
	std::vector<double> h_ref_grads;
	std::string err;

	bool profiling = false;
	std::vector<ProfRecord> prof;
	std::vector<const char *> prof_names;
	std::vector<float> prof_ms;
	int64_t launches = 0;
	double setup_seconds = 0.0;
};

namespace
{
	int fail(pfa_handle *h, int code, const std::string &msg)
	{
		if (h)
			h->err = msg;
		else
			g_create_error = msg;
		return code;
	}

#define PFA_CUDA(h, call)                                                                                      \
	do                                                                                                         \
	{                                                                                                          \
		cudaError_t e_ = (call);                                                                               \
		if (e_ != cudaSuccess)                                                                                 \
			return fail(h, e_ == cudaErrorMemoryAllocation ? PFA_ERR_NOMEM : PFA_ERR_CUDA,                     \
						std::string(#call) + ": " + cudaGetErrorString(e_));                                   \
	} while (0)

	template <typename T>
	int dev_alloc(pfa_handle *h, T **p, size_t count)
	{
		void *q = nullptr;
		PFA_CUDA(h, cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
		h->owned.push_back(q);
		*p = static_cast<T *>(q);
		return PFA_OK;
	}

	template <typename T>
	int dev_upload(pfa_handle *h, T **p, const T *src, size_t count)
	{
		int rc = dev_alloc(h, p, count);
		if (rc != PFA_OK)
			return rc;
		PFA_CUDA(h, cudaMemcpyAsync(*p, src, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
		return PFA_OK;
	}

	bool is_device_ptr(const void *p)
	{
		cudaPointerAttributes attr;
		if (cudaPointerGetAttributes(&attr, p) != cudaSuccess)
		{
			cudaGetLastError();
			return false;
		}
		return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
	}

	constexpr size_t kMaxProfRecords = 8192;
#ifndef PFA_REST_BATCH_QUOTA
#define PFA_REST_BATCH_QUOTA 4
#endif
	// owner-computes schedule (pfa_collane2.h): strip rows of the small class and steps per chunk; the environment
	// variables are for tuning experiments only
	inline int env_int(const char *name, int dflt)
	{
		const char *v = std::getenv(name);
		return v ? std::atoi(v) : dflt;
	}
	const int kSmallRows = env_int("PFA_CL_SMALL_ROWS", 96);
	const int kChunkSteps = env_int("PFA_CL_CHUNK_STEPS", 24);
	const int kBucketElements = env_int("PFA_CL_BUCKET", 16384); // 7 MB of P2 records per bucket
	constexpr int kRestBatchQuota = PFA_REST_BATCH_QUOTA; // pfa_grad_hess_part(PFA_PART_REST): warp batches per warp

	// NVTX range for the phases of a C-ABI call (H2D staging, zero fill, kernels, D2H); a no-op without a profiler attached
	struct NvtxRange
	{
		explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
		~NvtxRange() { nvtxRangePop(); }
		NvtxRange(const NvtxRange &) = delete;
		NvtxRange &operator=(const NvtxRange &) = delete;
	};

	void prof_begin(pfa_handle *h, const char *name, bool is_kernel = true)
	{
		if (is_kernel)
			++h->launches;
		if (!h->profiling || h->prof.size() >= kMaxProfRecords)
			return;
		ProfRecord r;
		r.name = name;
		r.stop_recorded = false;
		cudaEventCreate(&r.start);
		cudaEventCreate(&r.stop);
		cudaEventRecord(r.start, h->stream);
		h->prof.push_back(r);
	}
	void prof_end(pfa_handle *h)
	{
		if (!h->profiling || h->prof.empty() || h->prof.back().stop_recorded)
			return;
		cudaEventRecord(h->prof.back().stop, h->stream);
		h->prof.back().stop_recorded = true;
	}
	void prof_reset(pfa_handle *h)
	{
		for (auto &r : h->prof)
		{
			cudaEventDestroy(r.start);
			cudaEventDestroy(r.stop);
		}
		h->prof.clear();
	}

	int ensure_staging(pfa_handle *h, double **buf, size_t count)
	{
		if (*buf)
			return PFA_OK;
		return dev_alloc(h, buf, count);
	}

	// resolves an input vector to a device pointer (copying from the host when needed)
	int stage_input(pfa_handle *h, const double *x, const double **x_dev)
	{
		if (x == nullptr)
			return fail(h, PFA_ERR_INVALID, "displacement pointer is NULL");
		if (is_device_ptr(x))
		{
			*x_dev = x;
			return PFA_OK;
		}
		NvtxRange range("pfa: H2D displacement");
		int rc = ensure_staging(h, &h->s_x, size_t(h->ndof));
		if (rc != PFA_OK)
			return rc;
		PFA_CUDA(h, cudaMemcpyAsync(h->s_x, x, size_t(h->ndof) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
		*x_dev = h->s_x;
		return PFA_OK;
	}

	struct OutBuf
	{
		double *user = nullptr; // caller pointer (host or device) or NULL
		double *dev = nullptr;  // where the kernels write
		bool to_host = false;
		size_t count = 0;
	};

	// `capacity` (>= count) is what the staging buffer is allocated with on first use: the same buffer
	// serves the full-size and the Dirichlet-reduced outputs
	int stage_output(pfa_handle *h, double *user, size_t count, double **staging, OutBuf &o, size_t capacity = 0)
	{
		o.user = user;
		o.count = count;
		if (user == nullptr)
			return PFA_OK;
		if (is_device_ptr(user))
		{
			o.dev = user;
			return PFA_OK;
		}
		int rc = ensure_staging(h, staging, std::max(count, capacity));
		if (rc != PFA_OK)
			return rc;
		o.dev = *staging;
		o.to_host = true;
		return PFA_OK;
	}

	int finish_output(pfa_handle *h, const OutBuf &o)
	{
		if (o.to_host)
		{
			NvtxRange range("pfa: D2H result");
			PFA_CUDA(h, cudaMemcpyAsync(o.user, o.dev, o.count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
		}
		return PFA_OK;
	}

	// the one place that launches the assembly kernels
	int run_assemble(pfa_handle *h, bool linear, const double *x, int project_to_psd,
					 double *energy, double *energy_per_el, double *grad, double *values, double scale = 1.0, bool reduced = false, int part = PFA_PART_ALL)
	{
		NvtxRange call_range(linear ? "pfa: linear assembly" : "pfa: energy/gradient/Hessian assembly");
		h->err.clear();
		PFA_CUDA(h, cudaSetDevice(h->device));
		if (reduced && (!h->has_constraints || h->d_entry_red == nullptr))
			return fail(h, h->has_constraints ? PFA_ERR_UNSUPPORTED : PFA_ERR_INVALID,
						h->has_constraints ? "the fused Dirichlet-reduced assembly exists for NeoHookean P1/P2 tets only (use pfa_project_*)"
										   : "call pfa_set_constrained_dofs first");
		if (project_to_psd && values == nullptr)
			project_to_psd = 0; // no Hessian requested: nothing to project
		// The element stiffness of LinearElasticity is positive semi-definite by construction: ipc::project_to_psd returns such a
		// matrix unchanged (its six rigid-body eigenvalues are zero up to rounding), so the flag has no effect there.
		if (project_to_psd && h->dm.material == PFA_LINEAR_ELASTICITY)
			project_to_psd = 0;
		if (project_to_psd && h->dm.material != PFA_NEOHOOKEAN && h->dm.material != PFA_SAINT_VENANT && h->dm.material != PFA_MOONEY_RIVLIN && h->dm.material != PFA_VISCOUS_DAMPING && h->dm.material != PFA_FIXED_COROTATIONAL)
			return fail(h, PFA_ERR_UNSUPPORTED, "project_to_psd applies to the NLAssembler materials only");
		if (scale != 1.0 && !rowlane_applies(h->dm.material, h->dm.n_loc, h->dm.n_qp))
			return fail(h, PFA_ERR_UNSUPPORTED, "a Form weight other than 1 is fused for NeoHookean P1/P2 tets only");
		if ((h->dm.material == PFA_LAPLACIAN || h->dm.material == PFA_MASS) && !linear)
			return fail(h, PFA_ERR_UNSUPPORTED, "Laplacian and Mass are LinearAssemblers: only pfa_linear_stiffness applies");
		if ((h->dm.material == PFA_NEOHOOKEAN || h->dm.material == PFA_SAINT_VENANT || h->dm.material == PFA_MOONEY_RIVLIN || h->dm.material == PFA_VISCOUS_DAMPING || h->dm.material == PFA_FIXED_COROTATIONAL) && linear)
			return fail(h, PFA_ERR_UNSUPPORTED, "NeoHookean / SaintVenant are NLAssemblers: pfa_linear_stiffness does not apply");

		AssembleArgs a;
		int rc;
		if (!linear)
		{
			rc = stage_input(h, x, &a.x);
			if (rc != PFA_OK)
				return rc;
		}
		OutBuf oe, op, og, ov;
		if ((rc = stage_output(h, energy, 1, &h->d_energy, oe)) != PFA_OK)
			return rc;
		if ((rc = stage_output(h, energy_per_el, size_t(h->dm.n_el), &h->s_epe, op)) != PFA_OK)
			return rc;
		const size_t n_grad = reduced ? size_t(h->ndof_red) : size_t(h->ndof);
		const size_t n_val = reduced ? size_t(h->nnz_red) : size_t(h->nnz);
		// the full-size staging buffers are large enough for the reduced outputs as well
		if ((rc = stage_output(h, grad, n_grad, &h->s_grad, og, size_t(h->ndof))) != PFA_OK)
			return rc;
		if ((rc = stage_output(h, values, n_val, &h->s_values, ov, size_t(h->nnz))) != PFA_OK)
			return rc;
		a.energy = oe.dev;
		a.energy_per_el = op.dev;
		a.grad = og.dev;
		a.values = ov.dev;
		a.project_to_psd = project_to_psd;
		a.scale = scale;
		DeviceMesh dm = h->dm;
		if (reduced)
		{
			a.old_to_new = h->d_old_to_new;
			dm.entry = h->d_entry_red;
			dm.cstride = h->d_cstride_red;
			dm.zoff = nullptr; // values[] of the reduced matrix is cleared by a memset
		}
		a.work_counter = h->d_counter;
		a.e_begin = part == PFA_PART_REST ? h->dm.n_first : 0;
		a.e_end = part == PFA_PART_FIRST ? h->dm.n_first : h->dm.n_el;
		if (part != PFA_PART_ALL)
		{
			if (oe.to_host || op.to_host || og.to_host || ov.to_host || (x != nullptr && a.x != x))
				return fail(h, PFA_ERR_INVALID, "pfa_grad_hess_part works on device pointers only");
			dm.zoff = nullptr;
			// the second part is meant to run next to the interface exchange of another stream
			if (part == PFA_PART_REST)
			{
				// PFA_REST_QUOTA (environment, experiments): 0 = persistent warps
				static const int quota = [] { const char *v = std::getenv("PFA_REST_QUOTA"); return v ? std::atoi(v) : kRestBatchQuota; }();
				a.batch_quota = quota;
			}
		}
		PFA_CUDA(h, cudaMemsetAsync(h->d_counter, 0, 4 * sizeof(int), h->stream));
		// owner-computes path: full matrix of a NeoHookean P1/P2 handle that opted in; writes every entry once
		const bool use_cl = h->cl.enabled && !linear && !reduced && !project_to_psd && part == PFA_PART_ALL && a.values != nullptr;

		// outputs are accumulated with atomics: zero them first (rhs.setZero / set_zero,
		// Assembler.cpp:586-587, 666-667)
		if ((a.energy || a.grad || a.values) && part != PFA_PART_REST)
		{
			NvtxRange range("pfa: zero fill");
			prof_begin(h, "zero_fill(cudaMemsetAsync)", false);
			if (a.energy)
				PFA_CUDA(h, cudaMemsetAsync(a.energy, 0, sizeof(double), h->stream));
			// (owner-computes partition: the gradient entries of nodes owned by other ranks stay untouched, like their columns)
			if (a.grad && !(use_cl && h->cl_partial))
				PFA_CUDA(h, cudaMemsetAsync(a.grad, 0, n_grad * sizeof(double), h->stream));
			if (a.values && dm.zoff != nullptr && !linear && !project_to_psd && !use_cl)
				a.epoch = ++h->epoch; // the row-lane kernel clears values[] itself, block by block, just ahead of the scatter
			else if (a.values && !use_cl)
				PFA_CUDA(h, cudaMemsetAsync(a.values, 0, n_val * sizeof(double), h->stream));
			prof_end(h);
		}

		if (a.e_end <= a.e_begin)
			return PFA_OK; // empty part: outputs are cleared (or left), nothing to launch
		// ViscousDamping without a previous displacement: every output is zero (ViscousDamping.cpp:125-126, 176-179, 299-300)
		const bool no_prev = h->dm.material == PFA_VISCOUS_DAMPING && !h->has_prev;
		if (no_prev && a.energy_per_el)
			PFA_CUDA(h, cudaMemsetAsync(a.energy_per_el, 0, size_t(h->dm.n_el) * sizeof(double), h->stream));
		a.x_prev = h->d_x_prev;
		a.inv_dt = 1.0 / h->dt;
		const char *kname = use_cl ? "assemble_nh_column_lane(records+columns)" : "assemble";
		NvtxRange kernel_range(use_cl ? "pfa: owner-computes kernels (records + columns)" : "pfa: assembly kernel");
		prof_begin(h, kname);
		int cl_launches = 0;
		cudaError_t ce = no_prev ? cudaSuccess
						 : use_cl ? launch_column_lane2(dm, a, h->cl, h->sm_count, h->stream, &cl_launches)
								  : launch_assemble(dm, a, linear, h->sm_count, h->stream, &kname);
		h->launches += cl_launches; // records (+ energy sum) + one column kernel per strip class
		if (h->profiling && !h->prof.empty() && !h->prof.back().stop_recorded)
			h->prof.back().name = kname;
		prof_end(h);
		if (ce == cudaErrorNotSupported)
		{
			cudaGetLastError();
			return fail(h, PFA_ERR_UNSUPPORTED, project_to_psd ? "project_to_psd needs 2 N (N|1) doubles of shared memory per warp, N = 3 n_loc (rounded up to even) <= 128: P1..P4 tets, Q1/Q2 hexes"
																: "this material / element type combination is not implemented");
		}
		if (ce != cudaSuccess)
			return fail(h, PFA_ERR_CUDA, std::string("assembly kernel launch: ") + cudaGetErrorString(ce));

		if ((rc = finish_output(h, oe)) != PFA_OK || (rc = finish_output(h, op)) != PFA_OK || (rc = finish_output(h, og)) != PFA_OK || (rc = finish_output(h, ov)) != PFA_OK)
			return rc;
		if (oe.to_host || op.to_host || og.to_host || ov.to_host)
			PFA_CUDA(h, cudaStreamSynchronize(h->stream));
		return PFA_OK;
	}
} // namespace

extern "C"
{
	int pfa_create(const pfa_mesh_desc *d, pfa_handle **out)
	{
		NvtxRange create_range("pfa_create: pattern, schedule, geometry");
		g_create_error.clear();
		if (!d || !out)
			return fail(nullptr, PFA_ERR_INVALID, "pfa_create: NULL argument");
		*out = nullptr;
		if (d->struct_size != int32_t(sizeof(pfa_mesh_desc)))
			return fail(nullptr, PFA_ERR_INVALID, "pfa_create: struct_size does not match this library's pfa_mesh_desc");
		if (d->material < PFA_NEOHOOKEAN || d->material > PFA_FIXED_COROTATIONAL)
			return fail(nullptr, PFA_ERR_INVALID, "pfa_create: unknown material");
		if (d->n_elements <= 0 || d->n_loc <= 0 || d->n_bases <= 0 || d->n_qp <= 0 || d->n_ghost_elements < 0)
			return fail(nullptr, PFA_ERR_INVALID, "pfa_create: n_elements, n_loc, n_bases and n_qp must be positive");
		if (d->n_first_elements < 0 || d->n_first_elements > d->n_elements)
			return fail(nullptr, PFA_ERR_INVALID, "pfa_create: n_first_elements must lie in [0, n_elements]");
		const bool is_mass = d->material == PFA_MASS;
		if (!d->conn || !d->quad_weights || (!is_mass && !d->ref_grads))
			return fail(nullptr, PFA_ERR_INVALID, "pfa_create: conn, quad_weights and ref_grads are required");
		if (is_mass && (!d->ref_vals || !d->density))
			return fail(nullptr, PFA_ERR_INVALID, "pfa_create: PFA_MASS needs ref_vals and density");
		if (is_mass && d->n_loc != 4 && d->n_loc != 10 && d->n_loc != 20 && d->n_loc != 35)
			return fail(nullptr, PFA_ERR_UNSUPPORTED, "pfa_create: PFA_MASS is implemented for P1..P4 tets");
		// Lame parameters, or the density in both slots
		const double *lam_src = is_mass ? d->density : d->lambda, *mu_src = is_mass ? d->density : d->mu;
		const bool affine = d->vertices != nullptr;
		if (!affine && !(d->jac_it && d->da))
			return fail(nullptr, PFA_ERR_INVALID, "pfa_create: give either vertices (affine) or jac_it + da");
		if (d->material != PFA_LAPLACIAN && (!lam_src || !mu_src))
			return fail(nullptr, PFA_ERR_INVALID, "pfa_create: lambda and mu are required for elastic materials");
		if (d->material != PFA_LAPLACIAN && d->material_stride != 1 && d->material_stride != d->n_qp)
			return fail(nullptr, PFA_ERR_INVALID, "pfa_create: material_stride must be 1 or n_qp");
		const bool three_params = d->material == PFA_MOONEY_RIVLIN;
		if (three_params && !d->param3)
			return fail(nullptr, PFA_ERR_INVALID, "pfa_create: PFA_MOONEY_RIVLIN needs param3 (k) besides lambda (c1) and mu (c2)");

		// owned_nodes is honoured by the owner-computes kernels only: any other path would hand back partial sums of the
		// interface columns without saying so
		if (d->owned_nodes != nullptr
			&& (!column_lane2_applies(d->material, d->n_loc, d->n_qp) || !affine || (d->flags & PFA_FLAG_ROW_LANE) != 0
				|| (d->n_ghost_elements > 0 && (d->flags & PFA_FLAG_GHOST_GEOMETRY) == 0)))
			return fail(nullptr, PFA_ERR_UNSUPPORTED, "pfa_create: owned_nodes needs the owner-computes path (NeoHookean P1/P2 on affine tets, no PFA_FLAG_ROW_LANE, ghost elements with PFA_FLAG_GHOST_GEOMETRY)");

		int n_dev = 0;
		if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0)
		{
			cudaGetLastError();
			return fail(nullptr, PFA_ERR_NO_DEVICE, "pfa_create: no CUDA device available (this path has no CPU fallback)");
		}
		if (d->device < 0 || d->device >= n_dev)
			return fail(nullptr, PFA_ERR_INVALID, "pfa_create: device ordinal out of range");

		// multi-GPU owner-computes form: ghost elements carry geometry and material too
		const bool ghost_geom = (d->flags & PFA_FLAG_GHOST_GEOMETRY) != 0 && d->n_ghost_elements > 0;
		if (ghost_geom && !affine)
			return fail(nullptr, PFA_ERR_UNSUPPORTED, "pfa_create: PFA_FLAG_GHOST_GEOMETRY needs affine elements (vertices)");
		const auto t0 = std::chrono::steady_clock::now();
		pfa_handle *h = new (std::nothrow) pfa_handle();
		if (!h)
			return fail(nullptr, PFA_ERR_NOMEM, "pfa_create: out of host memory");
		auto bail = [&](int rc) {
			g_create_error = h->err;
			pfa_destroy(h);
			return rc;
		};
// inside pfa_create a CUDA failure must release the half-built handle and keep the message for pfa_last_error(NULL)
#define PFA_CREATE_CUDA(call)                                                                              \
	do                                                                                                     \
	{                                                                                                      \
		cudaError_t e_ = (call);                                                                           \
		if (e_ != cudaSuccess)                                                                             \
		{                                                                                                  \
			h->err = std::string("pfa_create: " #call ": ") + cudaGetErrorString(e_);                      \
			return bail(e_ == cudaErrorMemoryAllocation ? PFA_ERR_NOMEM : PFA_ERR_CUDA);                   \
		}                                                                                                  \
	} while (0)
		h->device = d->device;
		{
			cudaError_t e = cudaSetDevice(d->device);
			cudaDeviceProp prop;
			if (e == cudaSuccess)
				e = cudaGetDeviceProperties(&prop, d->device);
			if (e != cudaSuccess)
			{
				h->err = std::string("pfa_create: ") + cudaGetErrorString(e);
				return bail(PFA_ERR_CUDA);
			}
			if (prop.major < 10)
			{
				h->err = "pfa_create: device is not sm_100 class (library is built for sm_100a only)";
				return bail(PFA_ERR_NO_DEVICE);
			}
			h->sm_count = prop.multiProcessorCount;
			e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
			if (e != cudaSuccess)
			{
				h->err = std::string("pfa_create: ") + cudaGetErrorString(e);
				return bail(PFA_ERR_CUDA);
			}
		}

		DeviceMesh &m = h->dm;
		m.material = d->material;
		m.size = d->material == PFA_LAPLACIAN ? 1 : 3;
		m.n_el = d->n_elements;
		m.n_first = d->n_first_elements;
		m.n_loc = d->n_loc;
		m.n_bases = d->n_bases;
		m.n_qp = d->n_qp;
		m.geom_per_qp = affine ? 0 : 1;
		m.mat_stride = d->material == PFA_LAPLACIAN ? 1 : d->material_stride;
		if (!is_mass && !assemble_supported(m))
		{
			h->err = "pfa_create: n_loc x n_qp too large for the shared-memory staging of this build";
			return bail(PFA_ERR_UNSUPPORTED);
		}

		// internal element order (once per mesh): Morton order of the centroids when the vertices
		// are known; the caller's order otherwise
		std::vector<int32_t> perm;
		std::vector<int32_t> conn_p;
		std::vector<double> vert_p, lam_p, mu_p, p3_p;
		const double *p3_in = three_params ? d->param3 : nullptr;
		const int32_t *conn_in = d->conn;
		const double *vert_in = d->vertices, *lam_in = lam_src, *mu_in = mu_src;
		try
		{
			if (affine && !(d->flags & PFA_FLAG_KEEP_ELEMENT_ORDER) && d->n_elements > 1)
			{
				// the two element groups of pfa_grad_hess_part are ordered separately
				const int nf = d->n_first_elements;
				std::vector<int32_t> p2;
				if (nf > 1)
					spatial_element_order(d->vertices, nf, perm);
				else
					for (int e = 0; e < nf; ++e)
						perm.push_back(e);
				if (d->n_elements - nf > 1)
					spatial_element_order(d->vertices + size_t(nf) * 12, d->n_elements - nf, p2);
				else
					for (int e = 0; e < d->n_elements - nf; ++e)
						p2.push_back(e);
				for (int32_t v : p2)
					perm.push_back(v + nf);
				const size_t ne_ = size_t(d->n_elements), nl_ = size_t(d->n_loc), ng_ = size_t(d->n_ghost_elements);
				const size_t ngeo_ = ne_ + (ghost_geom ? ng_ : 0); // elements with geometry / material
				for (size_t e = ne_; e < ngeo_; ++e)
					perm.push_back(int32_t(e)); // ghost elements keep their place after the own elements
				conn_p.resize((ne_ + ng_) * nl_);
				vert_p.resize(ngeo_ * 12);
				for (size_t e = 0; e < ngeo_; ++e)
				{
					const size_t o = size_t(perm[e]);
					std::memcpy(&conn_p[e * nl_], d->conn + o * nl_, nl_ * sizeof(int32_t));
					std::memcpy(&vert_p[e * 12], d->vertices + o * 12, 12 * sizeof(double));
				}
				if (!ghost_geom)
					std::memcpy(conn_p.data() + ne_ * nl_, d->conn + ne_ * nl_, ng_ * nl_ * sizeof(int32_t));
				conn_in = conn_p.data();
				vert_in = vert_p.data();
				if (d->material != PFA_LAPLACIAN)
				{
					const size_t st_ = size_t(d->material_stride);
					lam_p.resize(ngeo_ * st_);
					mu_p.resize(ngeo_ * st_);
					for (size_t e = 0; e < ngeo_; ++e)
						for (size_t k = 0; k < st_; ++k)
						{
							lam_p[e * st_ + k] = lam_src[size_t(perm[e]) * st_ + k];
							mu_p[e * st_ + k] = mu_src[size_t(perm[e]) * st_ + k];
						}
					lam_in = lam_p.data();
					mu_in = mu_p.data();
					if (three_params)
					{
						p3_p.resize(ngeo_ * st_);
						for (size_t e = 0; e < ngeo_; ++e)
							for (size_t k = 0; k < st_; ++k)
								p3_p[e * st_ + k] = d->param3[size_t(perm[e]) * st_ + k];
						p3_in = p3_p.data();
					}
				}
			}
		}
		catch (const std::bad_alloc &)
		{
			h->err = "pfa_create: out of host memory while re-ordering the elements";
			return bail(PFA_ERR_NOMEM);
		}

		// pattern + slot map on the host (once per mesh)
		HostPattern hp;
		try
		{
			build_pattern(conn_in, d->n_elements + d->n_ghost_elements, d->n_loc, d->n_bases, hp);
		}
		catch (const std::bad_alloc &)
		{
			h->err = "pfa_create: out of host memory while building the pattern";
			return bail(PFA_ERR_NOMEM);
		}
		catch (const std::exception &ex)
		{
			h->err = ex.what();
			return bail(PFA_ERR_INVALID);
		}
		m.n_pairs = int64_t(hp.adj.size());
		h->ndof = int64_t(m.n_bases) * m.size;
		h->nnz = m.n_pairs * m.size * m.size;
		h->large_index = (d->flags & PFA_FLAG_LARGE_INDEX) != 0;
		if (h->nnz >= (int64_t(1) << 31) && !h->large_index)
		{
			h->err = "pfa_create: nnz exceeds int32 (StiffnessMatrix uses int indices unless POLYSOLVE_LARGE_INDEX): pass PFA_FLAG_LARGE_INDEX";
			return bail(PFA_ERR_UNSUPPORTED);
		}
		if (h->nnz >= (int64_t(1) << 31) && rowlane_applies(d->material, d->n_loc, d->n_qp))
		{
			h->err = "pfa_create: the NeoHookean P1/P2 kernels keep int32 entry tables: nnz must stay below 2^31";
			return bail(PFA_ERR_UNSUPPORTED);
		}

		int rc = PFA_OK;
#define UP(dst, src, count, T)                                             \
	do                                                                     \
	{                                                                      \
		T *tmp_ = nullptr;                                                 \
		rc = dev_upload<T>(h, &tmp_, (const T *)(src), size_t(count));     \
		if (rc != PFA_OK)                                                  \
			return bail(rc);                                               \
		dst = tmp_;                                                        \
	} while (0)
		const size_t ne = size_t(m.n_el), nl = size_t(m.n_loc), nq = size_t(m.n_qp);
		const size_t ngeo = ne + (ghost_geom ? size_t(d->n_ghost_elements) : 0); // elements with geometry / material on the device
		h->n_geo_elements = int32_t(ngeo);
		UP(m.conn, conn_in, ngeo * nl, int32_t);
		if (!perm.empty())
		{
			UP(h->d_elem_id, perm.data(), ngeo, int32_t);
			m.elem_id = h->d_elem_id;
		}
		if (d->ref_grads)
		{
			UP(m.ref_grads, d->ref_grads, nq * nl * 3, double);
			h->h_ref_grads.assign(d->ref_grads, d->ref_grads + nq * nl * 3);
			m.ref_grads_host = h->h_ref_grads.data();
			m.p2_structured = p2_table_structured(m.ref_grads_host, m.n_loc, m.n_qp) ? 1 : 0;
		}
		if (is_mass)
			UP(m.ref_vals, d->ref_vals, nq * nl, double);
		if ((m.material == PFA_LAPLACIAN || m.material == PFA_LINEAR_ELASTICITY) && affine && m.mat_stride == 1)
		{
			std::vector<double> mom;
			reference_moments(d->ref_grads, d->quad_weights, m.n_loc, m.n_qp, mom);
			UP(m.ref_moments, mom.data(), mom.size(), double);
			PFA_CREATE_CUDA(cudaStreamSynchronize(h->stream)); // mom is a local
		}
		UP(m.qweights, d->quad_weights, nq, double);
		UP(m.adj_off, hp.adj_off.data(), hp.adj_off.size(), int32_t);
		UP(m.adj, hp.adj.data(), hp.adj.size(), int32_t);
		if (rowlane_applies(m.material, m.n_loc, m.n_qp))
		{
			// values index of the (i, j) block's first entry and the column stride, per element
			std::vector<int32_t> cstride(ne * nl);
			for (size_t e = 0; e < ne; ++e)
				for (size_t j = 0; j < nl; ++j)
				{
					const int32_t gj = conn_in[e * nl + j];
					const int32_t off = hp.adj_off[size_t(gj)], deg = hp.adj_off[size_t(gj) + 1] - off;
					cstride[e * nl + j] = (3 * deg) | (7 << 28); // all three components of the node exist
					for (size_t i = 0; i < nl; ++i)
					{
						int32_t &sl = hp.slot[e * nl * nl + i * nl + j];
						sl = 9 * off + 3 * (sl - off);
					}
				}
			UP(m.entry, hp.slot.data(), size_t(ne) * nl * nl, int32_t);
			UP(m.cstride, cstride.data(), ne * nl, int32_t);
			// in-kernel zero fill of values[] (opt-in): who clears which column block
			std::vector<int32_t> zoff, zruns;
			if (d->flags & PFA_FLAG_INKERNEL_ZERO)
			{
			build_zero_schedule(conn_in, m.n_el, m.n_loc, m.n_bases, hp.adj_off, m.size, rowlane_batch_elements(m.n_loc, m.n_qp), zoff, zruns);
			m.n_batches = int32_t(zoff.size()) - 1;
			UP(m.zoff, zoff.data(), zoff.size(), int32_t);
			{
				int32_t *zr = nullptr;
				if ((rc = dev_upload<int32_t>(h, &zr, zruns.data(), zruns.size())) != PFA_OK)
					return bail(rc);
				m.zruns = reinterpret_cast<const int2 *>(zr);
			}
			if ((rc = dev_alloc<int32_t>(h, &m.zflag, size_t(m.n_batches))) != PFA_OK)
				return bail(rc);
			PFA_CREATE_CUDA(cudaMemsetAsync(m.zflag, 0, size_t(m.n_batches) * sizeof(int32_t), h->stream));
			}
			PFA_CREATE_CUDA(cudaStreamSynchronize(h->stream)); // cstride, zoff, zruns are locals
			// owner-computes path (default): schedule of (element, node) incidences + record buffer
			static const bool rl_env = [] { const char *v = std::getenv("PFA_ROW_LANE"); return v && std::atoi(v) != 0; }();
			int max_deg = 0;
			for (size_t b = 0; b + 1 < hp.adj_off.size(); ++b)
				max_deg = std::max(max_deg, hp.adj_off[b + 1] - hp.adj_off[b]);
			if (!(d->flags & PFA_FLAG_ROW_LANE) && !rl_env && affine && column_lane2_applies(m.material, m.n_loc, m.n_qp) && max_deg < 128 && (d->n_ghost_elements == 0 || ghost_geom))
			{
				try
				{
					// columns of at most kSmallRows strip rows form the first launch, the rest (P2 vertex nodes) the second;
					// chunks of about kChunkSteps steps are handed out to the warps
					const cl2::Schedule S = cl2::build_schedule(int(ngeo), m.n_loc, m.n_bases, conn_in, hp.adj_off.data(), hp.adj.data(), kSmallRows, kChunkSteps, d->owned_nodes, kBucketElements);
					UP(h->cl.grp_info, S.grp_info.data(), S.grp_info.size(), int32_t);
					UP(h->cl.grp_off, S.grp_off.data(), S.grp_off.size(), int32_t);
					UP(h->cl.grp_rows, S.grp_rows.data(), S.grp_rows.size(), int32_t);
					UP(h->cl.chunk_off, S.chunk_off.data(), S.chunk_off.size(), int32_t);
					UP(h->cl.inc, S.inc.data(), S.inc.size(), uint32_t);
					std::vector<double> rgp(nl * nq * 4, 0.0);
					for (size_t i = 0; i < nl; ++i)
						for (size_t q = 0; q < nq; ++q)
							for (size_t c = 0; c < 3; ++c)
								rgp[(i * nq + q) * 4 + c] = d->ref_grads[(q * nl + i) * 3 + c];
					UP(h->cl.rg_padded, rgp.data(), rgp.size(), double);
					if ((rc = dev_alloc<double>(h, &h->cl.records, ngeo * column_lane2_record_doubles(m.n_qp))) != PFA_OK || (rc = dev_alloc<double>(h, &h->cl.block_energy, (ngeo + 127) / 128)) != PFA_OK || (rc = dev_alloc<int>(h, &h->cl.counters, 2)) != PFA_OK)
						return bail(rc);
					for (int c = 0; c < 2; ++c)
					{
						h->cl.n_chunks[c] = S.n_chunks[c];
						h->cl.rows_max[c] = S.rows_max[c];
						h->cl.n_steps[c] = S.n_steps[c];
					}
					h->cl.n_record_elements = int32_t(ngeo);
					double za = 0.0, zb = 0.0;
					if (cl2::p2_rule_weights(d->ref_grads, m.n_loc, m.n_qp, za, zb))
					{
						h->cl.p2z = 1;
						h->cl.z4b = 4.0 * zb;
						h->cl.zbeta = 4.0 * (za - zb);
					}
					h->cl_partial = d->owned_nodes != nullptr;
					h->cl.enabled = 1;
					PFA_CREATE_CUDA(cudaStreamSynchronize(h->stream)); // S is a local
				}
				catch (const std::bad_alloc &)
				{
					h->err = "pfa_create: out of host memory while building the column-lane schedule";
					return bail(PFA_ERR_NOMEM);
				}
			}
		}
		else
			UP(m.slot, hp.slot.data(), size_t(ne) * nl * nl, int32_t);
		if (d->owned_nodes != nullptr && !h->cl.enabled)
		{
			h->err = "pfa_create: owned_nodes was given but the owner-computes path is off for this mesh (PFA_ROW_LANE=1 in the environment, or a node with 128 or more neighbours)";
			return bail(PFA_ERR_UNSUPPORTED);
		}
		if (m.material != PFA_LAPLACIAN)
		{
			const size_t cnt = ngeo * size_t(m.mat_stride);
			if ((rc = dev_upload<double>(h, &h->d_lambda, lam_in, cnt)) != PFA_OK || (rc = dev_upload<double>(h, &h->d_mu, mu_in, cnt)) != PFA_OK)
				return bail(rc);
			m.lambda = h->d_lambda;
			m.mu = h->d_mu;
			if (three_params)
			{
				if ((rc = dev_upload<double>(h, &h->d_param3, p3_in, cnt)) != PFA_OK)
					return bail(rc);
				m.param3 = h->d_param3;
			}
		}
		if (affine)
		{
			double *d_vert = nullptr, *jit = nullptr, *detj = nullptr;
			if ((rc = dev_upload<double>(h, &d_vert, vert_in, ngeo * 12)) != PFA_OK || (rc = dev_alloc<double>(h, &jit, ngeo * 9)) != PFA_OK || (rc = dev_alloc<double>(h, &detj, ngeo)) != PFA_OK)
				return bail(rc);
			++h->launches;
			cudaError_t e = launch_geometry_precompute(d_vert, int(ngeo), jit, detj, h->stream);
			if (e != cudaSuccess)
			{
				h->err = std::string("geometry precompute: ") + cudaGetErrorString(e);
				return bail(PFA_ERR_CUDA);
			}
			m.jit = jit;
			m.detj = detj;
		}
		else
		{
			UP(m.jit, d->jac_it, ne * nq * 9, double);
			UP(m.detj, d->da, ne * nq, double);
		}
#undef UP
		if ((rc = dev_alloc<double>(h, &h->d_energy, 1)) != PFA_OK || (rc = dev_alloc<int>(h, &h->d_counter, 4)) != PFA_OK)
			return bail(rc);
		if (!h->large_index)
		{
			if ((rc = dev_alloc<int32_t>(h, &h->d_outer, size_t(h->ndof) + 1)) != PFA_OK || (rc = dev_alloc<int32_t>(h, &h->d_inner, size_t(h->nnz))) != PFA_OK)
				return bail(rc);
			++h->launches;
			cudaError_t e = launch_expand_inner(m, h->d_outer, h->d_inner, h->stream);
			if (e == cudaSuccess)
				e = cudaStreamSynchronize(h->stream);
			if (e != cudaSuccess)
			{
				h->err = std::string("pattern expansion: ") + cudaGetErrorString(e);
				return bail(PFA_ERR_CUDA);
			}
		}
		else
			PFA_CREATE_CUDA(cudaStreamSynchronize(h->stream));
#undef PFA_CREATE_CUDA
		h->h_adj_off.swap(hp.adj_off);
		h->h_adj.swap(hp.adj);
		h->setup_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		*out = h;
		return PFA_OK;
	}

	void pfa_destroy(pfa_handle *h)
	{
		if (!h)
			return;
		cudaSetDevice(h->device);
		if (h->stream)
			cudaStreamSynchronize(h->stream);
		prof_reset(h);
		for (void *p : h->owned)
			cudaFree(p);
		for (void *p : h->owned_proj)
			cudaFree(p);
		if (h->scan_scratch)
			cudaFree(h->scan_scratch);
		if (h->stream && h->own_stream)
			cudaStreamDestroy(h->stream);
		delete h;
	}

	const char *pfa_last_error(const pfa_handle *h)
	{
		return h ? h->err.c_str() : g_create_error.c_str();
	}

	int pfa_sizes(const pfa_handle *h, int32_t *size, int64_t *ndof, int64_t *nnz)
	{
		if (!h)
			return PFA_ERR_INVALID;
		if (size)
			*size = h->dm.size;
		if (ndof)
			*ndof = h->ndof;
		if (nnz)
			*nnz = h->nnz;
		return PFA_OK;
	}

	// the int64 pattern of a PFA_FLAG_LARGE_INDEX handle on the device, built on first use
	static int ensure_pattern64(pfa_handle *h)
	{
		if (!h->large_index)
			return fail(h, PFA_ERR_UNSUPPORTED, "pfa_pattern_wide: the handle was not created with PFA_FLAG_LARGE_INDEX");
		if (h->d_outer64)
			return PFA_OK;
		PFA_CUDA(h, cudaSetDevice(h->device));
		int rc;
		if ((rc = dev_alloc<int64_t>(h, &h->d_outer64, size_t(h->ndof) + 1)) != PFA_OK || (rc = dev_alloc<int64_t>(h, &h->d_inner64, size_t(h->nnz))) != PFA_OK)
			return rc;
		++h->launches;
		PFA_CUDA(h, launch_expand_inner64(h->dm, h->d_outer64, h->d_inner64, h->stream));
		PFA_CUDA(h, cudaStreamSynchronize(h->stream));
		return PFA_OK;
	}

	int pfa_pattern_wide(pfa_handle *h, int64_t *nnz, const int64_t **outer, const int64_t **inner)
	{
		if (!h)
			return PFA_ERR_INVALID;
		int rc = ensure_pattern64(h);
		if (rc != PFA_OK)
			return rc;
		if (h->h_outer64.empty())
		{
			try
			{
				h->h_outer64.resize(size_t(h->ndof) + 1);
				h->h_inner64.resize(size_t(h->nnz));
			}
			catch (const std::bad_alloc &)
			{
				h->h_outer64.clear();
				return fail(h, PFA_ERR_NOMEM, "pfa_pattern_wide: out of host memory");
			}
			PFA_CUDA(h, cudaMemcpyAsync(h->h_outer64.data(), h->d_outer64, h->h_outer64.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
			PFA_CUDA(h, cudaMemcpyAsync(h->h_inner64.data(), h->d_inner64, h->h_inner64.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
			PFA_CUDA(h, cudaStreamSynchronize(h->stream));
		}
		if (nnz)
			*nnz = h->nnz;
		if (outer)
			*outer = h->h_outer64.data();
		if (inner)
			*inner = h->h_inner64.data();
		return PFA_OK;
	}

	int pfa_pattern_wide_device(pfa_handle *h, const int64_t **outer_dev, const int64_t **inner_dev)
	{
		if (!h)
			return PFA_ERR_INVALID;
		int rc = ensure_pattern64(h);
		if (rc != PFA_OK)
			return rc;
		if (outer_dev)
			*outer_dev = h->d_outer64;
		if (inner_dev)
			*inner_dev = h->d_inner64;
		return PFA_OK;
	}

	int pfa_pattern(pfa_handle *h, int64_t *nnz, const int32_t **outer, const int32_t **inner)
	{
		if (!h)
			return PFA_ERR_INVALID;
		if (h->large_index)
			return fail(h, PFA_ERR_UNSUPPORTED, "pfa_pattern: PFA_FLAG_LARGE_INDEX handle, use pfa_pattern_wide");
		PFA_CUDA(h, cudaSetDevice(h->device));
		if (h->h_outer.empty())
		{
			try
			{
				h->h_outer.resize(size_t(h->ndof) + 1);
				h->h_inner.resize(size_t(h->nnz));
			}
			catch (const std::bad_alloc &)
			{
				h->h_outer.clear();
				return fail(h, PFA_ERR_NOMEM, "pfa_pattern: out of host memory");
			}
			PFA_CUDA(h, cudaMemcpyAsync(h->h_outer.data(), h->d_outer, h->h_outer.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
			PFA_CUDA(h, cudaMemcpyAsync(h->h_inner.data(), h->d_inner, h->h_inner.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
			PFA_CUDA(h, cudaStreamSynchronize(h->stream));
		}
		if (nnz)
			*nnz = h->nnz;
		if (outer)
			*outer = h->h_outer.data();
		if (inner)
			*inner = h->h_inner.data();
		return PFA_OK;
	}

	int pfa_block_pattern(pfa_handle *h, int64_t *n_pairs, const int32_t **adj_off, const int32_t **adj)
	{
		if (!h)
			return PFA_ERR_INVALID;
		if (n_pairs)
			*n_pairs = int64_t(h->h_adj.size());
		if (adj_off)
			*adj_off = h->h_adj_off.data();
		if (adj)
			*adj = h->h_adj.data();
		return PFA_OK;
	}

	int pfa_pattern_device(pfa_handle *h, const int32_t **outer_dev, const int32_t **inner_dev)
	{
		if (!h)
			return PFA_ERR_INVALID;
		if (h->large_index)
			return fail(h, PFA_ERR_UNSUPPORTED, "pfa_pattern_device: PFA_FLAG_LARGE_INDEX handle, use pfa_pattern_wide_device");
		if (outer_dev)
			*outer_dev = h->d_outer;
		if (inner_dev)
			*inner_dev = h->d_inner;
		return PFA_OK;
	}

	int pfa_set_material_params(pfa_handle *h, const double *lambda, const double *mu, const double *p3, int32_t material_stride)
	{
		if (!h)
			return PFA_ERR_INVALID;
		if (h->dm.material == PFA_LAPLACIAN)
			return PFA_OK;
		if (h->dm.material == PFA_MASS && lambda && !mu)
			mu = lambda; // the density lives in both slots
		if (!lambda || !mu || material_stride != h->dm.mat_stride)
			return fail(h, PFA_ERR_INVALID, "pfa_set_materials: NULL array or material_stride differs from pfa_create");
		const bool three = h->dm.material == PFA_MOONEY_RIVLIN;
		if (three && !p3)
			return fail(h, PFA_ERR_INVALID, "pfa_set_material_params: this material has three parameters (use pfa_set_material_params with p3)");
		PFA_CUDA(h, cudaSetDevice(h->device));
		// (with PFA_FLAG_GHOST_GEOMETRY the arrays cover the ghost elements as well, like at pfa_create)
		const size_t cnt = size_t(h->n_geo_elements) * size_t(h->dm.mat_stride) * sizeof(double);
		const double *src[3] = {lambda, mu, p3};
		double *dst[3] = {h->d_lambda, h->d_mu, h->d_param3};
		const int n_arrays = three ? 3 : 2;
		if (h->d_elem_id == nullptr)
		{
			for (int k = 0; k < n_arrays; ++k)
				PFA_CUDA(h, cudaMemcpyAsync(dst[k], src[k], cnt, cudaMemcpyDefault, h->stream));
		}
		else
		{
			// elements live in the internal order: stage, then gather rows through elem_id
			int rc = ensure_staging(h, &h->s_mat, cnt / sizeof(double));
			if (rc != PFA_OK)
				return rc;
			for (int k = 0; k < n_arrays; ++k)
			{
				PFA_CUDA(h, cudaMemcpyAsync(h->s_mat, src[k], cnt, cudaMemcpyDefault, h->stream));
				++h->launches;
				PFA_CUDA(h, launch_gather_rows(h->s_mat, h->d_elem_id, h->n_geo_elements, h->dm.mat_stride, dst[k], h->stream));
			}
		}
		PFA_CUDA(h, cudaStreamSynchronize(h->stream));
		return PFA_OK;
	}

	int pfa_set_materials(pfa_handle *h, const double *lambda, const double *mu, int32_t material_stride)
	{
		return pfa_set_material_params(h, lambda, mu, nullptr, material_stride);
	}

	int pfa_set_previous(pfa_handle *h, const double *x_prev, double dt)
	{
		if (!h)
			return PFA_ERR_INVALID;
		if (!(dt > 0.0))
			return fail(h, PFA_ERR_INVALID, "pfa_set_previous: dt must be positive");
		PFA_CUDA(h, cudaSetDevice(h->device));
		h->has_prev = false; // stays off if the copy below fails
		if (x_prev)
		{
			if (h->d_x_prev == nullptr)
			{
				int rc = dev_alloc<double>(h, &h->d_x_prev, size_t(h->ndof));
				if (rc != PFA_OK)
					return rc;
			}
			PFA_CUDA(h, cudaMemcpyAsync(h->d_x_prev, x_prev, size_t(h->ndof) * sizeof(double), cudaMemcpyDefault, h->stream));
			PFA_CUDA(h, cudaStreamSynchronize(h->stream)); // the caller may reuse x_prev
		}
		h->dt = dt;
		h->has_prev = x_prev != nullptr;
		return PFA_OK;
	}

	int pfa_energy(pfa_handle *h, const double *x, double *energy)
	{
		if (!h || !energy)
			return fail(h, PFA_ERR_INVALID, "pfa_energy: NULL argument");
		return run_assemble(h, false, x, 0, energy, nullptr, nullptr, nullptr);
	}

	int pfa_energy_per_element(pfa_handle *h, const double *x, double *out)
	{
		if (!h || !out)
			return fail(h, PFA_ERR_INVALID, "pfa_energy_per_element: NULL argument");
		return run_assemble(h, false, x, 0, nullptr, out, nullptr, nullptr);
	}

	int pfa_gradient(pfa_handle *h, const double *x, double *grad)
	{
		if (!h || !grad)
			return fail(h, PFA_ERR_INVALID, "pfa_gradient: NULL argument");
		return run_assemble(h, false, x, 0, nullptr, nullptr, grad, nullptr);
	}

	int pfa_hessian(pfa_handle *h, const double *x, int project_to_psd, double *values)
	{
		if (!h || !values)
			return fail(h, PFA_ERR_INVALID, "pfa_hessian: NULL argument");
		return run_assemble(h, false, x, project_to_psd, nullptr, nullptr, nullptr, values);
	}

	int pfa_linear_stiffness(pfa_handle *h, double *values)
	{
		if (!h || !values)
			return fail(h, PFA_ERR_INVALID, "pfa_linear_stiffness: NULL argument");
		return run_assemble(h, true, nullptr, 0, nullptr, nullptr, nullptr, values);
	}

	int pfa_grad_hess(pfa_handle *h, const double *x, int project_to_psd, double *energy, double *grad, double *values)
	{
		if (!h)
			return PFA_ERR_INVALID;
		if (!energy && !grad && !values)
			return fail(h, PFA_ERR_INVALID, "pfa_grad_hess: all outputs are NULL");
		return run_assemble(h, false, x, project_to_psd, energy, nullptr, grad, values);
	}

	int pfa_grad_hess_weighted(pfa_handle *h, const double *x, int project_to_psd, double weight, double *energy, double *grad, double *values)
	{
		if (!h)
			return PFA_ERR_INVALID;
		if (!energy && !grad && !values)
			return fail(h, PFA_ERR_INVALID, "pfa_grad_hess_weighted: all outputs are NULL");
		return run_assemble(h, false, x, project_to_psd, energy, nullptr, grad, values, weight);
	}

	int pfa_is_step_valid(pfa_handle *h, const double *x, int32_t *valid, double *energy)
	{
		if (!h || !valid)
			return fail(h, PFA_ERR_INVALID, "pfa_is_step_valid: NULL argument");
		if (h->dm.material == PFA_LAPLACIAN)
			return fail(h, PFA_ERR_UNSUPPORTED, "pfa_is_step_valid: Laplacian has no nonlinear gradient");
		PFA_CUDA(h, cudaSetDevice(h->device));
		int rc = ensure_staging(h, &h->s_grad, size_t(h->ndof));
		if (rc != PFA_OK)
			return rc;
		if (!h->d_flag && (rc = dev_alloc(h, &h->d_flag, 1)) != PFA_OK)
			return rc;
		// gradient into device scratch (a device pointer: nothing is copied back), energy as asked
		rc = run_assemble(h, false, x, 0, energy, nullptr, h->s_grad, nullptr);
		if (rc != PFA_OK)
			return rc;
		PFA_CUDA(h, cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream));
		prof_begin(h, "any_nan_kernel");
		cudaError_t ce = launch_any_nan(h->s_grad, h->ndof, h->d_flag, h->sm_count, h->stream);
		prof_end(h);
		if (ce != cudaSuccess)
			return fail(h, PFA_ERR_CUDA, std::string("any_nan_kernel: ") + cudaGetErrorString(ce));
		int flag = 0;
		PFA_CUDA(h, cudaMemcpyAsync(&flag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
		PFA_CUDA(h, cudaStreamSynchronize(h->stream));
		*valid = flag ? 0 : 1;
		return PFA_OK;
	}

	int pfa_set_constrained_dofs(pfa_handle *h, const int32_t *dofs, int64_t n)
	{
		if (h && h->large_index)
			return fail(h, PFA_ERR_UNSUPPORTED, "pfa_set_constrained_dofs: not available on a PFA_FLAG_LARGE_INDEX handle (int32 pattern arrays)");
		if (!h || n < 0 || (n > 0 && !dofs))
			return fail(h, PFA_ERR_INVALID, "pfa_set_constrained_dofs: NULL list or negative count");
		if (h->ndof >= (int64_t(1) << 31) - 1)
			return fail(h, PFA_ERR_UNSUPPORTED, "pfa_set_constrained_dofs: ndof exceeds int32");
		h->err.clear();
		PFA_CUDA(h, cudaSetDevice(h->device));
		PFA_CUDA(h, cudaStreamSynchronize(h->stream));
		for (void *p : h->owned_proj)
			cudaFree(p);
		h->owned_proj.clear();
		h->has_constraints = false;
		h->h_outer_red.clear();
		h->h_inner_red.clear();
		h->s_vec_out = h->s_val_out = nullptr;
		h->d_entry_red = h->d_cstride_red = nullptr;
		auto alloc = [&](auto **p, size_t count) -> int {
			void *q = nullptr;
			PFA_CUDA(h, cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(**p)));
			h->owned_proj.push_back(q);
			*p = static_cast<std::remove_reference_t<decltype(*p)>>(q);
			return PFA_OK;
		};
		const int32_t ndof = int32_t(h->ndof);
		int32_t *d_dofs = nullptr, *d_keep = nullptr, *d_rank = nullptr, *d_cnt = nullptr;
		int rc;
		if (!h->d_flag && (rc = dev_alloc(h, &h->d_flag, 1)) != PFA_OK)
			return rc;
		if ((rc = alloc(&d_keep, size_t(ndof) + 1)) != PFA_OK || (rc = alloc(&d_rank, size_t(ndof) + 1)) != PFA_OK
			|| (rc = alloc(&h->d_old_to_new, size_t(ndof))) != PFA_OK)
			return rc;
		const int32_t *dofs_dev = dofs;
		if (n > 0 && !is_device_ptr(dofs))
		{
			if ((rc = alloc(&d_dofs, size_t(n))) != PFA_OK)
				return rc;
			PFA_CUDA(h, cudaMemcpyAsync(d_dofs, dofs, size_t(n) * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
			dofs_dev = d_dofs;
		}
		PFA_CUDA(h, cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream));
		h->launches += 2;
		PFA_CUDA(h, launch_mark_constrained(dofs_dev, n, ndof, d_keep, h->d_flag, h->stream));
		PFA_CUDA(h, cudaMemsetAsync(d_keep + ndof, 0, sizeof(int32_t), h->stream));
		// rank of every dof among the kept ones; rank[ndof] = number of kept dofs
		PFA_CUDA(h, exclusive_scan_i32(d_keep, d_rank, int64_t(ndof) + 1, &h->scan_scratch, &h->scan_scratch_bytes, h->stream));
		int32_t n_red = 0;
		int bad = 0;
		PFA_CUDA(h, cudaMemcpyAsync(&n_red, d_rank + ndof, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
		PFA_CUDA(h, cudaMemcpyAsync(&bad, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
		PFA_CUDA(h, cudaStreamSynchronize(h->stream));
		if (bad)
			return fail(h, PFA_ERR_INVALID, "pfa_set_constrained_dofs: a constrained dof is outside [0, ndof)");
		if ((rc = alloc(&h->d_not_constraints, size_t(n_red))) != PFA_OK || (rc = alloc(&d_cnt, size_t(n_red) + 1)) != PFA_OK
			|| (rc = alloc(&h->d_outer_red, size_t(n_red) + 1)) != PFA_OK)
			return rc;
		h->launches += 2;
		PFA_CUDA(h, launch_finish_maps(d_keep, d_rank, ndof, h->d_old_to_new, h->d_not_constraints, h->stream));
		PFA_CUDA(h, cudaMemsetAsync(d_cnt, 0, (size_t(n_red) + 1) * sizeof(int32_t), h->stream));
		PFA_CUDA(h, launch_count_kept(h->d_outer, h->d_inner, h->d_old_to_new, ndof, d_cnt, h->stream));
		PFA_CUDA(h, exclusive_scan_i32(d_cnt, h->d_outer_red, int64_t(n_red) + 1, &h->scan_scratch, &h->scan_scratch_bytes, h->stream));
		int32_t nnz_red = 0;
		PFA_CUDA(h, cudaMemcpyAsync(&nnz_red, h->d_outer_red + n_red, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
		PFA_CUDA(h, cudaStreamSynchronize(h->stream));
		if ((rc = alloc(&h->d_inner_red, size_t(nnz_red))) != PFA_OK || (rc = alloc(&h->d_map, size_t(nnz_red))) != PFA_OK)
			return rc;
		++h->launches;
		PFA_CUDA(h, launch_fill_reduced(h->d_outer, h->d_inner, h->d_old_to_new, ndof, h->d_outer_red, h->d_inner_red, h->d_map, h->stream));
		PFA_CUDA(h, cudaStreamSynchronize(h->stream));
		h->d_entry_red = h->d_cstride_red = nullptr;
		if (rowlane_applies(h->dm.material, h->dm.n_loc, h->dm.n_qp) && h->dm.entry != nullptr)
		{
			const size_t nb = size_t(h->dm.n_bases), ne = size_t(h->dm.n_el), nl = size_t(h->dm.n_loc);
			int32_t *node_mask = nullptr, *rowprefix = nullptr, *cs_red = nullptr, *cbase_red = nullptr;
			if ((rc = alloc(&node_mask, nb)) != PFA_OK || (rc = alloc(&rowprefix, size_t(h->h_adj.size()))) != PFA_OK || (rc = alloc(&cs_red, nb)) != PFA_OK
				|| (rc = alloc(&cbase_red, nb)) != PFA_OK || (rc = alloc(&h->d_entry_red, ne * nl * nl)) != PFA_OK || (rc = alloc(&h->d_cstride_red, ne * nl)) != PFA_OK)
				return rc;
			h->launches += 3;
			PFA_CUDA(h, launch_reduced_tables(h->dm, d_keep, h->d_old_to_new, h->d_outer_red, node_mask, rowprefix, cs_red, cbase_red, h->d_entry_red, h->d_cstride_red, h->stream));
			PFA_CUDA(h, cudaStreamSynchronize(h->stream));
		}
		// d_keep / d_rank / d_cnt / d_dofs stay in owned_proj until the next call (small next to the map)
		h->ndof_red = n_red;
		h->nnz_red = nnz_red;
		h->has_constraints = true;
		return PFA_OK;
	}

	int pfa_reduced_sizes(const pfa_handle *h, int64_t *ndof_reduced, int64_t *nnz_reduced)
	{
		if (!h || !h->has_constraints)
			return PFA_ERR_INVALID;
		if (ndof_reduced)
			*ndof_reduced = h->ndof_red;
		if (nnz_reduced)
			*nnz_reduced = h->nnz_red;
		return PFA_OK;
	}

	int pfa_reduced_pattern(pfa_handle *h, const int32_t **outer, const int32_t **inner)
	{
		if (!h)
			return PFA_ERR_INVALID;
		if (!h->has_constraints)
			return fail(h, PFA_ERR_INVALID, "pfa_reduced_pattern: call pfa_set_constrained_dofs first");
		PFA_CUDA(h, cudaSetDevice(h->device));
		if (h->h_outer_red.empty())
		{
			try
			{
				h->h_outer_red.resize(size_t(h->ndof_red) + 1);
				h->h_inner_red.resize(size_t(h->nnz_red));
			}
			catch (const std::bad_alloc &)
			{
				h->h_outer_red.clear();
				return fail(h, PFA_ERR_NOMEM, "pfa_reduced_pattern: out of host memory");
			}
			PFA_CUDA(h, cudaMemcpyAsync(h->h_outer_red.data(), h->d_outer_red, h->h_outer_red.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
			PFA_CUDA(h, cudaMemcpyAsync(h->h_inner_red.data(), h->d_inner_red, h->h_inner_red.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
			PFA_CUDA(h, cudaStreamSynchronize(h->stream));
		}
		if (outer)
			*outer = h->h_outer_red.data();
		if (inner)
			*inner = h->h_inner_red.data();
		return PFA_OK;
	}

	int pfa_reduced_pattern_device(pfa_handle *h, const int32_t **outer_dev, const int32_t **inner_dev)
	{
		if (!h || !h->has_constraints)
			return PFA_ERR_INVALID;
		if (outer_dev)
			*outer_dev = h->d_outer_red;
		if (inner_dev)
			*inner_dev = h->d_inner_red;
		return PFA_OK;
	}

	namespace
	{
		// shared by pfa_project_gradient / pfa_project_hessian: dst[t] = scale * src[map[t]]
		int project_common(pfa_handle *h, const char *what, const double *src, size_t n_src, const int32_t *map, size_t n_dst, double scale,
						   double *dst, double **stage_in, double **stage_out, bool out_in_proj)
		{
			if (!h)
				return PFA_ERR_INVALID;
			if (!h->has_constraints)
				return fail(h, PFA_ERR_INVALID, std::string(what) + ": call pfa_set_constrained_dofs first");
			if (!src || !dst)
				return fail(h, PFA_ERR_INVALID, std::string(what) + ": NULL argument");
			h->err.clear();
			PFA_CUDA(h, cudaSetDevice(h->device));
			const double *src_dev = src;
			int rc;
			if (!is_device_ptr(src))
			{
				if ((rc = ensure_staging(h, stage_in, n_src)) != PFA_OK)
					return rc;
				PFA_CUDA(h, cudaMemcpyAsync(*stage_in, src, n_src * sizeof(double), cudaMemcpyHostToDevice, h->stream));
				src_dev = *stage_in;
			}
			double *dst_dev = dst;
			const bool to_host = !is_device_ptr(dst);
			if (to_host)
			{
				if (!*stage_out)
				{
					void *q = nullptr;
					PFA_CUDA(h, cudaMalloc(&q, std::max<size_t>(n_dst, 1) * sizeof(double)));
					(out_in_proj ? h->owned_proj : h->owned).push_back(q);
					*stage_out = static_cast<double *>(q);
				}
				dst_dev = *stage_out;
			}
			prof_begin(h, what);
			cudaError_t ce = launch_gather_scale(src_dev, map, int64_t(n_dst), scale, dst_dev, h->sm_count, h->stream);
			prof_end(h);
			if (ce != cudaSuccess)
				return fail(h, PFA_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(ce));
			if (to_host)
			{
				PFA_CUDA(h, cudaMemcpyAsync(dst, dst_dev, n_dst * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
				PFA_CUDA(h, cudaStreamSynchronize(h->stream));
			}
			return PFA_OK;
		}
	} // namespace

	int pfa_project_gradient(pfa_handle *h, const double *grad_full, double scale, double *grad_reduced)
	{
		if (!h)
			return PFA_ERR_INVALID;
		return project_common(h, "project_gradient(gather)", grad_full, size_t(h->ndof), h->d_not_constraints, size_t(h->ndof_red), scale, grad_reduced,
							  &h->s_vec_in, &h->s_vec_out, true);
	}

	int pfa_project_hessian(pfa_handle *h, const double *values_full, double scale, double *values_reduced)
	{
		if (!h)
			return PFA_ERR_INVALID;
		return project_common(h, "project_hessian(gather)", values_full, size_t(h->nnz), h->d_map, size_t(h->nnz_red), scale, values_reduced,
							  &h->s_values, &h->s_val_out, true);
	}

	namespace
	{
		// device view of a read-only vector argument (host data is staged into *staging)
		int in_dev(pfa_handle *h, const double *p, size_t n, double **staging, const double **out)
		{
			if (is_device_ptr(p))
			{
				*out = p;
				return PFA_OK;
			}
			int rc = ensure_staging(h, staging, n);
			if (rc != PFA_OK)
				return rc;
			PFA_CUDA(h, cudaMemcpyAsync(*staging, p, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
			*out = *staging;
			return PFA_OK;
		}
	} // namespace

	int pfa_inertia(pfa_handle *h, const double *mass_values, const double *x, const double *x_tilde, double *energy, double *grad)
	{
		if (h && h->large_index)
			return fail(h, PFA_ERR_UNSUPPORTED, "pfa_inertia: not available on a PFA_FLAG_LARGE_INDEX handle (int32 pattern arrays)");
		if (!h || !mass_values || !x)
			return fail(h, PFA_ERR_INVALID, "pfa_inertia / pfa_symv: NULL argument");
		if (!energy && !grad)
			return fail(h, PFA_ERR_INVALID, "pfa_inertia: all outputs are NULL");
		h->err.clear();
		PFA_CUDA(h, cudaSetDevice(h->device));
		const double *v_dev, *x_dev, *xt_dev = nullptr;
		int rc;
		if ((rc = in_dev(h, mass_values, size_t(h->nnz), &h->s_values, &v_dev)) != PFA_OK || (rc = in_dev(h, x, size_t(h->ndof), &h->s_x, &x_dev)) != PFA_OK)
			return rc;
		if (x_tilde && (rc = in_dev(h, x_tilde, size_t(h->ndof), &h->s_vec_in, &xt_dev)) != PFA_OK)
			return rc;
		OutBuf oe, og;
		if ((rc = stage_output(h, energy, 1, &h->d_energy, oe)) != PFA_OK || (rc = stage_output(h, grad, size_t(h->ndof), &h->s_grad, og)) != PFA_OK)
			return rc;
		if (oe.dev)
			PFA_CUDA(h, cudaMemsetAsync(oe.dev, 0, sizeof(double), h->stream));
		prof_begin(h, "symv_kernel");
		cudaError_t ce = launch_symv(h->d_outer, h->d_inner, v_dev, x_dev, xt_dev, int32_t(h->ndof), og.dev, oe.dev, h->stream);
		prof_end(h);
		if (ce != cudaSuccess)
			return fail(h, PFA_ERR_CUDA, std::string("symv_kernel: ") + cudaGetErrorString(ce));
		if ((rc = finish_output(h, oe)) != PFA_OK || (rc = finish_output(h, og)) != PFA_OK)
			return rc;
		if (oe.to_host || og.to_host)
			PFA_CUDA(h, cudaStreamSynchronize(h->stream));
		return PFA_OK;
	}

	int pfa_symv(pfa_handle *h, const double *values, const double *x, double *y)
	{
		if (h && h->large_index)
			return fail(h, PFA_ERR_UNSUPPORTED, "pfa_symv: not available on a PFA_FLAG_LARGE_INDEX handle (int32 pattern arrays)");
		if (!y)
			return fail(h, PFA_ERR_INVALID, "pfa_symv: NULL argument");
		return pfa_inertia(h, values, x, nullptr, nullptr, y);
	}

	int pfa_axpy(pfa_handle *h, int64_t n, double a, const double *x, double *y)
	{
		if (!h || !x || !y || n < 0)
			return fail(h, PFA_ERR_INVALID, "pfa_axpy: NULL argument");
		if (!is_device_ptr(x) || !is_device_ptr(y))
			return fail(h, PFA_ERR_INVALID, "pfa_axpy works on device pointers");
		PFA_CUDA(h, cudaSetDevice(h->device));
		prof_begin(h, "axpy_kernel");
		cudaError_t ce = launch_axpy(n, a, x, y, h->sm_count, h->stream);
		prof_end(h);
		if (ce != cudaSuccess)
			return fail(h, PFA_ERR_CUDA, std::string("axpy_kernel: ") + cudaGetErrorString(ce));
		return PFA_OK;
	}

	int pfa_grad_hess_part(pfa_handle *h, const double *x, int project_to_psd, double *energy, double *grad, double *values, int part)
	{
		if (!h)
			return PFA_ERR_INVALID;
		if (part < PFA_PART_ALL || part > PFA_PART_REST)
			return fail(h, PFA_ERR_INVALID, "pfa_grad_hess_part: unknown part");
		if (!energy && !grad && !values)
			return fail(h, PFA_ERR_INVALID, "pfa_grad_hess_part: all outputs are NULL");
		return run_assemble(h, false, x, project_to_psd, energy, nullptr, grad, values, 1.0, false, part);
	}

	int pfa_grad_hess_reduced(pfa_handle *h, const double *x, int project_to_psd, double scale, double *energy, double *grad_reduced, double *values_reduced)
	{
		if (!h)
			return PFA_ERR_INVALID;
		if (!energy && !grad_reduced && !values_reduced)
			return fail(h, PFA_ERR_INVALID, "pfa_grad_hess_reduced: all outputs are NULL");
		return run_assemble(h, false, x, project_to_psd, energy, nullptr, grad_reduced, values_reduced, scale, true);
	}

	void *pfa_host_alloc(size_t bytes)
	{
		void *p = nullptr;
		if (bytes == 0)
			return nullptr;
		if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess)
		{
			cudaGetLastError(); // no device / not enough lockable memory: the caller falls back to ordinary memory
			return nullptr;
		}
		return p;
	}

	void pfa_host_free(void *p)
	{
		if (p)
			cudaFreeHost(p);
	}

	int pfa_synchronize(pfa_handle *h)
	{
		if (!h)
			return PFA_ERR_INVALID;
		PFA_CUDA(h, cudaSetDevice(h->device));
		PFA_CUDA(h, cudaStreamSynchronize(h->stream));
		return PFA_OK;
	}

	void *pfa_stream(pfa_handle *h) { return h ? (void *)h->stream : nullptr; }

	int pfa_set_stream(pfa_handle *h, void *stream)
	{
		if (!h)
			return PFA_ERR_INVALID;
		PFA_CUDA(h, cudaSetDevice(h->device));
		PFA_CUDA(h, cudaStreamSynchronize(h->stream));
		if (h->own_stream && h->stream)
			cudaStreamDestroy(h->stream);
		h->stream = (cudaStream_t)stream;
		h->own_stream = false;
		return PFA_OK;
	}

	int pfa_profile_enable(pfa_handle *h, int on)
	{
		if (!h)
			return PFA_ERR_INVALID;
		h->profiling = on != 0;
		return PFA_OK;
	}

	int pfa_profile_read(pfa_handle *h, int cap, const char **names, float *ms)
	{
		if (!h)
			return PFA_ERR_INVALID;
		PFA_CUDA(h, cudaSetDevice(h->device));
		PFA_CUDA(h, cudaStreamSynchronize(h->stream));
		h->prof_names.clear();
		h->prof_ms.clear();
		for (auto &r : h->prof)
		{
			float t = 0.f;
			if (r.stop_recorded)
				cudaEventElapsedTime(&t, r.start, r.stop);
			h->prof_names.push_back(r.name);
			h->prof_ms.push_back(t);
		}
		prof_reset(h);
		const int n = int(h->prof_ms.size());
		for (int i = 0; i < n && i < cap; ++i)
		{
			if (names)
				names[i] = h->prof_names[i];
			if (ms)
				ms[i] = h->prof_ms[i];
		}
		return n;
	}

	int64_t pfa_launch_count(const pfa_handle *h) { return h ? h->launches : 0; }
	double pfa_setup_seconds(const pfa_handle *h) { return h ? h->setup_seconds : 0.0; }
}
