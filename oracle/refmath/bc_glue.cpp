// TEST INFRASTRUCTURE: the reference's OWN BCLagrangianForm::project_gradient / project_hessian
// (solver/forms/lagrangian/BCLagrangianForm.cpp:149-155, 167-213) and the constructor loops that build not_constraints_ /
// old_to_new_ (:121-137), extracted at build time into ../_ref/bc_extracted.inc and compiled verbatim against
// mini_eigen.hpp / mini_sparse.hpp -> oracle/_ref/libbcref.so.
#include "mini_sparse.hpp"

#include <cassert>
#include <vector>

namespace polyfem
{
	typedef Eigen::SparseMatrix<double, Eigen::ColMajor> StiffnessMatrix; // utils/Types.hpp:24

	namespace solver
	{
		class BCLagrangianForm // only the members the extracted code touches (BCLagrangianForm.hpp)
		{
		public:
			int n_dofs_ = 0;
			std::vector<int> boundary_nodes_;
			std::vector<int> not_constraints_;
			std::vector<int> old_to_new_;
			std::vector<Eigen::Triplet<double>> A_triplets;
			void build_maps();
			void project_gradient(Eigen::VectorXd &grad) const;
			void project_hessian(StiffnessMatrix &hessian) const;
		};

#include "../_ref/bc_extracted.inc"
	} // namespace solver
} // namespace polyfem

using polyfem::StiffnessMatrix;
using polyfem::solver::BCLagrangianForm;

extern "C"
{
	// constrained[n_c] -> not_constraints[n_dofs - n_c], old_to_new[n_dofs]; returns the reduced size
	int ref_bc_maps(int n_dofs, int n_c, const int *constrained, int *not_constraints, int *old_to_new)
	{
		BCLagrangianForm f;
		f.n_dofs_ = n_dofs;
		f.boundary_nodes_.assign(constrained, constrained + n_c);
		f.build_maps();
		for (size_t k = 0; k < f.not_constraints_.size(); ++k)
			not_constraints[k] = f.not_constraints_[k];
		for (int k = 0; k < n_dofs; ++k)
			old_to_new[k] = f.old_to_new_[size_t(k)];
		return int(f.not_constraints_.size());
	}

	// grad[n_dofs] in, out[n_red]
	int ref_bc_project_gradient(int n_dofs, int n_c, const int *constrained, const double *grad, double *out)
	{
		BCLagrangianForm f;
		f.n_dofs_ = n_dofs;
		f.boundary_nodes_.assign(constrained, constrained + n_c);
		f.build_maps();
		Eigen::VectorXd g(n_dofs);
		for (int k = 0; k < n_dofs; ++k)
			g[k] = grad[k];
		f.project_gradient(g);
		for (long k = 0; k < g.size(); ++k)
			out[k] = g[k];
		return int(g.size());
	}

	// CSC in (n_dofs x n_dofs, nnz), CSC out in caller buffers sized for the input; returns the reduced nnz
	long ref_bc_project_hessian(int n_dofs, int n_c, const int *constrained, long nnz, const int *outer, const int *inner, const double *values,
								int *outer_red, int *inner_red, double *values_red)
	{
		BCLagrangianForm f;
		f.n_dofs_ = n_dofs;
		f.boundary_nodes_.assign(constrained, constrained + n_c);
		f.build_maps();
		StiffnessMatrix H = Eigen::Map<const StiffnessMatrix>(n_dofs, n_dofs, nnz, outer, inner, values);
		f.project_hessian(H);
		for (long c = 0; c <= H.cols(); ++c)
			outer_red[c] = H.outerIndexPtr()[c];
		for (long k = 0; k < H.nonZeros(); ++k)
		{
			inner_red[k] = H.innerIndexPtr()[k];
			values_red[k] = H.valuePtr()[k];
		}
		return H.nonZeros();
	}
}
