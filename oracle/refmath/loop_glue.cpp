// TEST INFRASTRUCTURE: C entry points over the reference's OWN global loops NLAssembler::assemble_energy /
// assemble_gradient / assemble_hessian, LinearAssembler::assemble and their per-thread storage classes
// (assembler/Assembler.cpp:16-94, 157-384, 495-531, 574-643, 645-771), extracted at build time into ../_ref/loop_extracted.inc and compiled verbatim, driving the
// reference's OWN NeoHookean / LinearElasticity / Laplacian / Mass local functions (../_ref/nh_extracted.inc) and the reference's utils/MatrixCache.cpp
// compiled unmodified (shadow/ supplies <Eigen/...>, Types.hpp, Logger.hpp, MaybeParallelFor.hpp)
// -> oracle/_ref/libloopref.so. The element values (reference gradients, jac_it, det, weights) come in from the
// caller: they are oracle_assembly_values' outputs, themselves pinned against finalize3d by libgeomref.so.
// maybe_parallel_for runs `threads` consecutive chunks (shadow/polyfem/utils/MaybeParallelFor.hpp), so the
// per-thread storages, MatrixCache::copy / init(main) and the serial merge all run.
#include <polyfem/utils/MatrixCache.hpp>
#include <polyfem/utils/Logger.hpp>
#include <polyfem/utils/MaybeParallelFor.hpp>

#include <memory>

#define PFREF_LOOPS
#include "nh_harness.hpp" // opens namespace polyfem::assembler

#include "../_ref/nh_extracted.inc"

	using namespace basis;
	using namespace quadrature;
	using namespace utils;

#include "../_ref/loop_extracted.inc"
} // namespace polyfem::assembler

Eigen::MatrixXd ipc::project_to_psd(const Eigen::MatrixXd &m) { return m; }

using namespace polyfem::assembler;

struct refloop
{
	NeoHookeanElasticity nh;
	LinearElasticity le;
	Laplacian lap;
	Mass mass;
	int material = -1;
	AssemblyValsCache vals;
	std::vector<polyfem::basis::ElementBases> bases;
	int n_bases = 0;
	polyfem::utils::SparseMatrixCache mat_cache; // lives across assemble_hessian calls like ElasticForm::mat_cache_
	polyfem::StiffnessMatrix hess;
};

extern "C"
{
	// conn[n_el][n_loc]; ref_grads[n_qp][n_loc][3]; jac_it[n_el][n_qp][9] row-major; det[n_el][n_qp]; weights[n_qp]
	refloop *refloop_new(int n_el, int n_loc, int n_qp, int n_bases, const int *conn, const double *ref_grads, const double *jac_it,
						 const double *det, const double *weights, double lambda, double mu)
	{
		auto *r = new refloop();
		r->nh.params_.lambda = lambda;
		r->nh.params_.mu = mu;
		r->n_bases = n_bases;
		r->bases.resize(size_t(n_el));
		r->vals.cache.resize(size_t(n_el));
		for (int e = 0; e < n_el; ++e)
		{
			ElementAssemblyValues &v = r->vals.cache[size_t(e)];
			v.element_id = e;
			v.quadrature.points.resize(n_qp, 3);
			v.quadrature.weights.resize(n_qp, 1);
			v.val.resize(n_qp, 3);
			v.det.resize(n_qp, 1);
			v.jac_it.resize(size_t(n_qp));
			for (int q = 0; q < n_qp; ++q)
			{
				v.quadrature.weights(q) = weights[q];
				v.det(q) = det[size_t(e) * n_qp + q];
				v.jac_it[size_t(q)].resize(3, 3);
				for (int a = 0; a < 3; ++a)
					for (int b = 0; b < 3; ++b)
						v.jac_it[size_t(q)](a, b) = jac_it[(size_t(e) * n_qp + q) * 9 + a * 3 + b];
			}
			v.basis_values.resize(size_t(n_loc));
			for (int i = 0; i < n_loc; ++i)
			{
				v.basis_values[size_t(i)].global = {Local2Global{conn[size_t(e) * n_loc + i], 1.0}};
				v.basis_values[size_t(i)].grad.resize(n_qp, 3);
				for (int q = 0; q < n_qp; ++q)
					for (int c = 0; c < 3; ++c)
						v.basis_values[size_t(i)].grad(q, c) = ref_grads[(size_t(q) * n_loc + i) * 3 + c];
			}
		}
		return r;
	}
	void refloop_free(refloop *r) { delete r; }

	static Eigen::MatrixXd column(const double *x, long n)
	{
		Eigen::MatrixXd m(n, 1);
		for (long k = 0; k < n; ++k)
			m(k) = x[k];
		return m;
	}

	double refloop_energy(refloop *r, const double *x, int threads)
	{
		polyfem::utils::ref_thread_count() = threads;
		const Eigen::MatrixXd d = column(x, long(r->n_bases) * 3), prev;
		const double e = r->nh.assemble_energy(true, r->bases, r->bases, r->vals, 0.0, 1.0, d, prev);
		polyfem::utils::ref_thread_count() = 1;
		return e;
	}
	void refloop_energy_per_element(refloop *r, const double *x, int threads, double *out)
	{
		polyfem::utils::ref_thread_count() = threads;
		const Eigen::MatrixXd d = column(x, long(r->n_bases) * 3), prev;
		const Eigen::VectorXd v = r->nh.assemble_energy_per_element(true, r->bases, r->bases, r->vals, 0.0, 1.0, d, prev);
		polyfem::utils::ref_thread_count() = 1;
		for (long k = 0; k < v.size(); ++k)
			out[k] = v(k);
	}
	void refloop_gradient(refloop *r, const double *x, int threads, double *out)
	{
		polyfem::utils::ref_thread_count() = threads;
		const Eigen::MatrixXd d = column(x, long(r->n_bases) * 3), prev;
		Eigen::MatrixXd rhs;
		r->nh.assemble_gradient(true, r->n_bases, r->bases, r->bases, r->vals, 0.0, 1.0, d, prev, rhs);
		polyfem::utils::ref_thread_count() = 1;
		for (long k = 0; k < rhs.size(); ++k)
			out[k] = rhs(k);
	}
	// returns nnz; the matrix cache persists in r, so a second call takes the cached-pattern path of SparseMatrixCache
	long refloop_hessian(refloop *r, const double *x, int threads)
	{
		polyfem::utils::ref_thread_count() = threads;
		const Eigen::MatrixXd d = column(x, long(r->n_bases) * 3), prev;
		r->nh.assemble_hessian(true, r->n_bases, false, r->bases, r->bases, r->vals, 0.0, 1.0, d, prev, r->mat_cache, r->hess);
		polyfem::utils::ref_thread_count() = 1;
		return r->hess.nonZeros();
	}
	// ---- LinearAssembler::assemble (Assembler.cpp:157-384) ----
	// material: 0 LinearElasticity, 1 Laplacian, 2 Mass. grad_t_m[n_el][n_qp][n_loc][3] (physical gradients),
	// ref_vals[n_qp][n_loc] (basis values, Mass only, at the mass quadrature), det[n_el][n_qp], weights[n_qp]
	refloop *refloop_linear_new(int material, int n_el, int n_loc, int n_qp, int n_bases, const int *conn, const double *grad_t_m,
								const double *ref_vals, const double *det, const double *weights, double lambda, double mu, double rho)
	{
		auto *r = new refloop();
		r->material = material;
		r->le.params_.lambda = lambda;
		r->le.params_.mu = mu;
		r->mass.density_.rho = rho;
		r->n_bases = n_bases;
		r->bases.resize(size_t(n_el));
		r->vals.cache.resize(size_t(n_el));
		r->vals.is_mass_ = material == 2;
		for (int e = 0; e < n_el; ++e)
		{
			ElementAssemblyValues &v = r->vals.cache[size_t(e)];
			v.element_id = e;
			v.quadrature.points.resize(n_qp, 3);
			v.quadrature.weights.resize(n_qp, 1);
			v.val.resize(n_qp, 3);
			v.det.resize(n_qp, 1);
			for (int q = 0; q < n_qp; ++q)
			{
				v.quadrature.weights(q) = weights[q];
				v.det(q) = det[size_t(e) * n_qp + q];
			}
			v.basis_values.resize(size_t(n_loc));
			for (int i = 0; i < n_loc; ++i)
			{
				AssemblyValues &b = v.basis_values[size_t(i)];
				b.global = {Local2Global{conn[size_t(e) * n_loc + i], 1.0}};
				b.grad_t_m.resize(n_qp, 3);
				b.val.resize(n_qp, 1);
				for (int q = 0; q < n_qp; ++q)
				{
					for (int c = 0; c < 3; ++c)
						b.grad_t_m(q, c) = grad_t_m ? grad_t_m[((size_t(e) * n_qp + q) * n_loc + i) * 3 + c] : 0.0;
					b.val(q) = ref_vals ? ref_vals[size_t(q) * n_loc + i] : 0.0;
				}
			}
		}
		return r;
	}
	long refloop_linear_assemble(refloop *r, int threads)
	{
		polyfem::utils::ref_thread_count() = threads;
		const LinearAssembler *a = r->material == 0 ? static_cast<const LinearAssembler *>(&r->le)
								   : r->material == 1 ? static_cast<const LinearAssembler *>(&r->lap)
													  : static_cast<const LinearAssembler *>(&r->mass);
		a->assemble(true, r->n_bases, r->bases, r->bases, r->vals, 0.0, r->hess, r->material == 2);
		polyfem::utils::ref_thread_count() = 1;
		return r->hess.nonZeros();
	}
	int refloop_cols(const refloop *r) { return int(r->hess.cols()); }
	const int *refloop_outer(const refloop *r) { return r->hess.outerIndexPtr(); }
	const int *refloop_inner(const refloop *r) { return r->hess.innerIndexPtr(); }
	const double *refloop_values(const refloop *r) { return r->hess.valuePtr(); }
}
