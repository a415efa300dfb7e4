// TEST INFRASTRUCTURE: C entry points over the reference's OWN FixedCorotational function bodies
// (assembler/FixedCorotational.cpp:293-436, 592-827) and its OWN 3 x 3 SVD (utils/svd.hpp:134-317: analytic eigenvalues of
// A^T A, eigenvectors by cofactors, U from A V), extracted at build time into ../_ref/fc_extracted.inc / svd_extracted.inc and
// compiled verbatim against mini_eigen.hpp. Used by tools/make_golden.py (tests/golden/fc_local.npz) and by
// tests/test_oracle_corotational_reference.py.
#include "nh_harness.hpp" // opens namespace polyfem::assembler
} // namespace polyfem::assembler

namespace polyfem
{
	struct DiffScalarBase
	{
		static void setVariableCount(long) {}
	};
	template <typename T>
	class AutoDiffAllocator;
	template <>
	class AutoDiffAllocator<double> // utils/AutodiffTypes.hpp: the double specialisation returns the value
	{
	public:
		double operator()(const int, double v) const { return v; }
	};
#include "../_ref/elutil_extracted.inc"

	namespace utils
	{
		// utils/svd.hpp:14-51, 336-368: AutoFlipSVD<Matrix3d> as its 3-D compute() runs it - fastSVD3d when singular vectors are
		// asked for - with the extracted member functions below; the Eigen::JacobiSVD base it only uses in 2-D is left out
		template <typename MatrixType>
		class AutoFlipSVD
		{
		public:
			AutoFlipSVD(const MatrixType &mtr, unsigned int computationOptions = 0)
			{
				if ((computationOptions & Eigen::ComputeFullU) || (computationOptions & Eigen::ComputeFullV))
					fastSVD3d(mtr, matrixU_flipped, singularValues_flipped, matrixV_flipped);
				else
					fastComputeSingularValues3d(mtr, singularValues_flipped);
			}
			const Eigen::Vector3d &singularValues() const { return singularValues_flipped; }
			const MatrixType &matrixU() const { return matrixU_flipped; }
			const MatrixType &matrixV() const { return matrixV_flipped; }

		protected:
			Eigen::Vector3d singularValues_flipped;
			MatrixType matrixU_flipped, matrixV_flipped;
			void fastComputeSingularValues3d(const Eigen::Matrix3d &A, Eigen::Vector3d &singular_values) // svd.hpp:319-333
			{
				Eigen::Vector3d lambda;
				fastEigenvalues(A.transpose() * A, lambda);
				if (lambda(2) < 0)
					lambda = (lambda.array() >= 0.0).select(lambda, 0.0);
				singular_values = lambda.array().sqrt();
				if (A.determinant() < 0)
					singular_values(2) = -singular_values(2);
			}
#include "../_ref/svd_extracted.inc"
		};
		template <int dim>
		Eigen::Vector<double, dim> singular_values(const Eigen::Matrix<double, dim, dim> &A) // svd.hpp:382-387
		{
			AutoFlipSVD<Eigen::Matrix<double, dim, dim>> svd(A, Eigen::ComputeFullU | Eigen::ComputeFullV);
			return svd.singularValues();
		}
	} // namespace utils
} // namespace polyfem

namespace polyfem::assembler
{
	class FixedCorotational // assembler/FixedCorotational.hpp:11-100: the members the extracted bodies use
	{
	public:
		int size() const { return 3; }
		LameParameters params_;
		template <int dim>
		double compute_energy_aux(const NonLinearAssemblerData &data) const;
		template <int n_basis, int dim>
		void compute_energy_hessian_aux_fast(const NonLinearAssemblerData &data, Eigen::MatrixXd &H) const;
		template <int n_basis, int dim>
		void compute_energy_aux_gradient_fast(const NonLinearAssemblerData &data, Eigen::VectorXd &G_flattened) const;
		template <int dim>
		static double compute_energy_from_singular_values(const Eigen::Vector<double, dim> &sigmas, const double lambda, const double mu);
		template <int dim>
		static Eigen::Vector<double, dim> compute_stress_from_singular_values(const Eigen::Vector<double, dim> &sigmas, const double lambda, const double mu);
		template <int dim>
		static Eigen::Matrix<double, dim, dim> compute_stiffness_from_singular_values(const Eigen::Vector<double, dim> &sigmas, const double lambda, const double mu);
		template <int dim>
		static double compute_energy_from_def_grad(const Eigen::Matrix<double, dim, dim> &F, const double lambda, const double mu);
		template <int dim>
		static Eigen::Matrix<double, dim, dim> compute_stress_from_def_grad(const Eigen::Matrix<double, dim, dim> &F, const double lambda, const double mu);
		template <int dim>
		static Eigen::Matrix<double, dim * dim, dim * dim> compute_stiffness_from_def_grad(const Eigen::Matrix<double, dim, dim> &F, const double lambda, const double mu);
	};
#include "../_ref/fc_extracted.inc"
} // namespace polyfem::assembler

using namespace polyfem::assembler;

extern "C"
{
	// u [n_basis][3], grads [n_qp][n_basis][3], jac_it [n_qp][9] row-major, da [n_qp]; out: energy, gradient [N] node-major,
	// hessian [N][N] row-major. The instantiation follows the reference's dispatch on the number of bases (4, 10, 20 fixed).
	int ref_fc_local(int n_basis, int n_qp, const double *u, const double *grads, const double *jac_it, const double *da, double lambda, double mu,
					 double *energy, double *gradient, double *hessian)
	{
		ElementAssemblyValues vals;
		Eigen::MatrixXd x(long(n_basis) * 3, 1), x_prev;
		Eigen::VectorXd dav(n_qp, 1);
		for (int k = 0; k < n_basis * 3; ++k)
			x(k) = u[k];
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
		FixedCorotational fc;
		fc.params_.lambda = lambda;
		fc.params_.mu = mu;
		const NonLinearAssemblerData data{vals, 0.0, 1.0, x, x_prev, dav};
		*energy = fc.compute_energy_aux<3>(data);
		const long N = long(n_basis) * 3;
		Eigen::VectorXd g;
		Eigen::MatrixXd H(N, N);
		if (n_basis == 4)
		{
			fc.compute_energy_aux_gradient_fast<4, 3>(data, g);
			fc.compute_energy_hessian_aux_fast<4, 3>(data, H);
		}
		else if (n_basis == 10)
		{
			fc.compute_energy_aux_gradient_fast<10, 3>(data, g);
			fc.compute_energy_hessian_aux_fast<10, 3>(data, H);
		}
		else if (n_basis == 20)
		{
			fc.compute_energy_aux_gradient_fast<20, 3>(data, g);
			fc.compute_energy_hessian_aux_fast<20, 3>(data, H);
		}
		else
		{
			fc.compute_energy_aux_gradient_fast<Eigen::Dynamic, 3>(data, g);
			fc.compute_energy_hessian_aux_fast<Eigen::Dynamic, 3>(data, H);
		}
		if (g.size() != N)
			return -1;
		for (long r = 0; r < N; ++r)
		{
			gradient[r] = g(r);
			for (long c = 0; c < N; ++c)
				hessian[r * N + c] = H(r, c);
		}
		return 0;
	}

	// the reference's signed SVD alone: U, V row-major 3 x 3, sigma[3]
	void ref_svd3(const double *A, double *U, double *sigma, double *V)
	{
		Eigen::Matrix3d a;
		for (int r = 0; r < 3; ++r)
			for (int c = 0; c < 3; ++c)
				a(r, c) = A[r * 3 + c];
		polyfem::utils::AutoFlipSVD<Eigen::Matrix3d> svd(a, Eigen::ComputeFullU | Eigen::ComputeFullV);
		for (int r = 0; r < 3; ++r)
		{
			sigma[r] = svd.singularValues()(r);
			for (int c = 0; c < 3; ++c)
			{
				U[r * 3 + c] = svd.matrixU()(r, c);
				V[r * 3 + c] = svd.matrixV()(r, c);
			}
		}
	}
}
