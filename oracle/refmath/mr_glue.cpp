// TEST INFRASTRUCTURE: C entry points over the reference's OWN MooneyRivlin code path: MooneyRivlinElasticity::elastic_energy<T>
// (MooneyRivlinElasticity.hpp:25-47) differentiated by the reference's OWN forward-mode scalars (utils/autodiff.h, included
// UNMODIFIED from the reference tree) inside GenericElastic's compute_energy_aux<double>, compute_gradient_from_stress and
// compute_hessian_from_stress (GenericElastic.hpp:92-212, 268-351; AutodiffType::STRESS is the default, :70), with
// compute_B_block (GenericElastic.cpp:19-33), first / second_invariant (utils/ElasticityUtils.hpp:138-150) and determinant
// (utils/MatrixUtils.hpp:21-35). All extracted at build time into ../_ref/ and compiled verbatim against mini_eigen.hpp.
// Used by tools/make_golden.py (tests/golden/mr_local.npz) and tests/test_oracle_mooney_reference.py.
#include "mini_eigen.hpp"

#include <polyfem/utils/autodiff.h> // the reference's own file; its <Eigen/Core> resolves to shadow_core/Eigen/Core
DECLARE_DIFFSCALAR_BASE();

#include "nh_harness.hpp" // opens namespace polyfem::assembler
} // namespace polyfem::assembler

namespace polyfem
{
	typedef Eigen::MatrixXd RowVectorNd; // utils/Types.hpp: a row vector of doubles; only handed through here
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
#include "../_ref/invariants_extracted.inc"
	namespace utils
	{
#include "../_ref/determinant_extracted.inc"
	}
} // namespace polyfem

namespace polyfem::assembler
{
	template <typename T>
	using DefGradMatrix = Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic, 0, 3, 3>; // GenericElastic.hpp:18-19
	struct GenericMatParam // assembler/MatParams.hpp: call form used by elastic_energy
	{
		double value = 0;
		double operator()(const RowVectorNd &, double, int) const { return value; }
	};

	template <typename Derived>
	class GenericElastic // GenericElastic.hpp:21-140: the members the extracted bodies use
	{
	public:
		int size() const { return 3; }
		const Derived &derived() const { return static_cast<const Derived &>(*this); }
		bool real_def_grad() const { return true; }

#include "../_ref/generic_members_extracted.inc"

		template <int dim>
		Eigen::Matrix<double, dim * dim, dim> compute_B_block(const Eigen::Matrix<double, 1, dim> &g) const;
	};
#include "../_ref/generic_bblock_extracted.inc"

	class MooneyRivlinElasticity : public GenericElastic<MooneyRivlinElasticity>
	{
	public:
		GenericMatParam c1_, c2_, k_;
#include "../_ref/mr_energy_extracted.inc"
	};
} // namespace polyfem::assembler

using namespace polyfem::assembler;

extern "C"
{
	// u [n_basis][3], grads [n_qp][n_basis][3], jac_it [n_qp][9] row-major, da [n_qp]; out: energy, gradient [N] node-major,
	// hessian [N][N] row-major; the instantiations follow assemble_gradient_stress_ad / assemble_hessian_stress_ad
	int ref_mr_local(int n_basis, int n_qp, const double *u, const double *grads, const double *jac_it, const double *da, double c1, double c2, double k,
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
		MooneyRivlinElasticity mr;
		mr.c1_.value = c1;
		mr.c2_.value = c2;
		mr.k_.value = k;
		const NonLinearAssemblerData data{vals, 0.0, 1.0, x, x_prev, dav};
		*energy = mr.compute_energy_aux<double>(data);
		const long N = long(n_basis) * 3;
		Eigen::VectorXd g;
		Eigen::MatrixXd H;
		if (n_basis == 4)
		{
			mr.compute_gradient_from_stress<4, 3>(data, g);
			mr.compute_hessian_from_stress<4, 3>(data, H);
		}
		else if (n_basis == 10)
		{
			mr.compute_gradient_from_stress<10, 3>(data, g);
			mr.compute_hessian_from_stress<10, 3>(data, H);
		}
		else if (n_basis == 20)
		{
			mr.compute_gradient_from_stress<20, 3>(data, g);
			mr.compute_hessian_from_stress<20, 3>(data, H);
		}
		else
		{
			mr.compute_gradient_from_stress<Eigen::Dynamic, 3>(data, g);
			mr.compute_hessian_from_stress<Eigen::Dynamic, 3>(data, H);
		}
		if (g.size() != N || H.rows() != N)
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
