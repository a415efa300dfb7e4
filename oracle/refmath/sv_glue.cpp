// TEST INFRASTRUCTURE: C entry points over the reference's OWN SaintVenant code path: compute_energy_aux<T> with
// strain_from_disp_grad and stress<T, N> (assembler/SaintVenantElasticity.cpp:9-20, 61-70, 206-266), differentiated by the reference's
// OWN forward-mode scalars (utils/autodiff.h, included UNMODIFIED) through its own gradient_from_energy / hessian_from_energy
// dispatch (utils/ElasticityUtils.cpp:81-270, SMALL_N = 80 and BIG_N = 1000 as in CMakeLists.txt:66-67), over the reference's own
// ElasticityTensor::set_from_lambda_mu / operator() (assembler/MatParams.cpp:91-123, 211-253). All extracted at build time into
// ../_ref/ and compiled verbatim against mini_eigen.hpp. Used by tools/make_golden.py (tests/golden/sv_local.npz) and
// tests/test_oracle_saint_venant_reference.py.
#include "mini_eigen.hpp"

#include <polyfem/utils/autodiff.h> // the reference's own file; its <Eigen/Core> resolves to shadow_core/Eigen/Core
DECLARE_DIFFSCALAR_BASE();

#include <array>
#include <functional>
#include <string>

#include "nh_harness.hpp" // opens namespace polyfem::assembler
} // namespace polyfem::assembler

namespace polyfem
{
	constexpr int SMALL_N = 80;  // POLYFEM_SMALL_N, CMakeLists.txt:66
	constexpr int BIG_N = 1000;  // POLYFEM_BIG_N, CMakeLists.txt:67
	struct LoggerStub
	{
		template <typename... A>
		void debug(A &&...) {}
	};
	inline LoggerStub &logger()
	{
		static LoggerStub l;
		return l;
	}
	template <class T>
	class AutoDiffAllocator // utils/AutodiffTypes.hpp:19-37
	{
	public:
		T operator()(const int i, double v) const { return T(i, v); }
	};
	template <>
	class AutoDiffAllocator<double>
	{
	public:
		double operator()(const int, double v) const { return v; }
	};
#include "../_ref/elutil_extracted.inc"
#include "../_ref/sv_dispatch_extracted.inc"
} // namespace polyfem

namespace polyfem::assembler
{
	class ElasticityTensor // assembler/MatParams.hpp:41-73: the members the extracted bodies use
	{
	public:
		void resize(const int size);
		double operator()(int i, int j) const;
		double &operator()(int i, int j);
		void set_from_lambda_mu(const double lambda, const double mu, const std::string &stress_unit, const std::string &root_path);

	private:
		Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, 0, 6, 6> stiffness_tensor_;
		int size_;
	};
#include "../_ref/sv_tensor_extracted.inc"

	class SaintVenantElasticity // assembler/SaintVenantElasticity.hpp:11-53
	{
	public:
		int size() const { return 3; }
		Eigen::VectorXd assemble_gradient(const NonLinearAssemblerData &data) const;
		Eigen::MatrixXd assemble_hessian(const NonLinearAssemblerData &data) const;
		template <typename T>
		T compute_energy_aux(const NonLinearAssemblerData &data) const;
		ElasticityTensor elasticity_tensor_;

	private:
		template <typename T, unsigned long N>
		T stress(const ElasticityTensor &elasticity_tensor, const std::array<T, N> &strain, const int j) const;
	};
	namespace
	{
#include "../_ref/sv_strain_extracted.inc"
	}
#include "../_ref/sv_extracted.inc"
	// LinearElasticity inside a nonlinear solve: compute_energy_aux<T> (LinearElasticity.cpp:103-132) and its autodiff gradient /
	// Hessian (:70-101), the class itself is declared in nh_harness.hpp
#include "../_ref/le_energy_extracted.inc"
#include "../_ref/le_ad_extracted.inc"
} // namespace polyfem::assembler

using namespace polyfem::assembler;

extern "C"
{
	static void fill_values(ElementAssemblyValues &vals, Eigen::VectorXd &dav, int n_basis, int n_qp, const double *grads, const double *jac_it, const double *da)
	{
		dav.resize(n_qp, 1);
		vals.quadrature.points.resize(n_qp, 3);
		vals.val.resize(n_qp, 3);
		vals.basis_values.resize(n_basis);
		for (int i = 0; i < n_basis; ++i)
		{
			vals.basis_values[i].global = {Local2Global{i, 1.0}};
			vals.basis_values[i].grad.resize(n_qp, 3);
			for (int q = 0; q < n_qp; ++q)
				for (int c = 0; c < 3; ++c)
					vals.basis_values[i].grad(q, c) = grads[(size_t(q) * n_basis + i) * 3 + c];
		}
		vals.jac_it.resize(n_qp);
		for (int q = 0; q < n_qp; ++q)
		{
			dav(q) = da[q];
			vals.jac_it[q].resize(3, 3);
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					vals.jac_it[q](r, c) = jac_it[size_t(q) * 9 + r * 3 + c];
		}
	}

	// LinearElasticity as an NLAssembler (LinearElasticity.cpp:65-132): energy and its autodiff gradient / Hessian
	int ref_le_nl_local(int n_basis, int n_qp, const double *u, const double *grads, const double *jac_it, const double *da, double lambda, double mu,
						double *energy, double *gradient, double *hessian)
	{
		ElementAssemblyValues vals;
		Eigen::MatrixXd x(long(n_basis) * 3, 1), x_prev;
		Eigen::VectorXd dav;
		for (int i = 0; i < n_basis * 3; ++i)
			x(i) = u[i];
		fill_values(vals, dav, n_basis, n_qp, grads, jac_it, da);
		LinearElasticity le;
		le.params_.lambda = lambda;
		le.params_.mu = mu;
		const NonLinearAssemblerData data{vals, 0.0, 1.0, x, x_prev, dav};
		*energy = le.compute_energy_aux<double>(data);
		const Eigen::VectorXd g = le.assemble_gradient(data);
		const Eigen::MatrixXd H = le.assemble_hessian(data);
		const long N = long(n_basis) * 3;
		if (g.size() != N || H.rows() != N || H.cols() != N)
			return -1;
		for (long r = 0; r < N; ++r)
		{
			gradient[r] = g(r);
			for (long c = 0; c < N; ++c)
				hessian[r * N + c] = H(r, c);
		}
		return 0;
	}

	// u [n_basis][3], grads [n_qp][n_basis][3], jac_it [n_qp][9] row-major, da [n_qp]; out: energy, gradient [N] node-major,
	// hessian [N][N] row-major
	int ref_sv_local(int n_basis, int n_qp, const double *u, const double *grads, const double *jac_it, const double *da, double lambda, double mu,
					 double *energy, double *gradient, double *hessian)
	{
		ElementAssemblyValues vals;
		Eigen::MatrixXd x(long(n_basis) * 3, 1), x_prev;
		Eigen::VectorXd dav(n_qp, 1);
		for (int i = 0; i < n_basis * 3; ++i)
			x(i) = u[i];
		vals.quadrature.points.resize(n_qp, 3);
		vals.val.resize(n_qp, 3);
		vals.basis_values.resize(n_basis);
		for (int i = 0; i < n_basis; ++i)
		{
			vals.basis_values[i].global = {Local2Global{i, 1.0}};
			vals.basis_values[i].grad.resize(n_qp, 3);
			for (int q = 0; q < n_qp; ++q)
				for (int c = 0; c < 3; ++c)
					vals.basis_values[i].grad(q, c) = grads[(size_t(q) * n_basis + i) * 3 + c];
		}
		vals.jac_it.resize(n_qp);
		for (int q = 0; q < n_qp; ++q)
		{
			dav(q) = da[q];
			vals.jac_it[q].resize(3, 3);
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					vals.jac_it[q](r, c) = jac_it[size_t(q) * 9 + r * 3 + c];
		}
		SaintVenantElasticity sv;
		sv.elasticity_tensor_.resize(3);
		sv.elasticity_tensor_.set_from_lambda_mu(lambda, mu, "", "");
		const NonLinearAssemblerData data{vals, 0.0, 1.0, x, x_prev, dav};
		*energy = sv.compute_energy_aux<double>(data);
		const Eigen::VectorXd g = sv.assemble_gradient(data);
		const Eigen::MatrixXd H = sv.assemble_hessian(data);
		const long N = long(n_basis) * 3;
		if (g.size() != N || H.rows() != N || H.cols() != N)
			return -1;
		for (long r = 0; r < N; ++r)
		{
			gradient[r] = g(r);
			for (long c = 0; c < N; ++c)
				hessian[r * N + c] = H(r, c);
		}
		return 0;
	}
}
