// TEST INFRASTRUCTURE: the reference's OWN ElementAssemblyValues::finalize3d (assembler/ElementAssemblyValues.cpp:65-104:
// J = sum grad N_g (x) node, det, jac_it = J^-T, grad_t_m = grad * jac_it), extracted at build time into
// ../_ref/geom_extracted.inc and compiled verbatim against mini_eigen.hpp -> oracle/_ref/libgeomref.so.
// (mini_eigen's 3x3 inverse is the cofactor formula; Eigen's own differs in rounding only.)
#include "mini_eigen.hpp"

#include <vector>

namespace polyfem
{
	namespace basis
	{
		struct Local2Global // basis/Basis.hpp:21-38
		{
			int index;
			double val;
			Eigen::Dense node; // 1 x 3
		};
		struct Basis
		{
			std::vector<Local2Global> g;
			const std::vector<Local2Global> &global() const { return g; }
		};
		struct ElementBases
		{
			std::vector<Basis> bases;
			bool has_parameterization = true;
		};
	} // namespace basis

	namespace assembler
	{
		using basis::Basis;
		using basis::ElementBases;
		struct AssemblyValues // assembler/AssemblyValues.hpp
		{
			Eigen::MatrixXd grad, grad_t_m;
			void finalize() { grad_t_m.resize(grad.rows(), grad.cols()); } // AssemblyValues.hpp:31-34
		};
		class ElementAssemblyValues // assembler/ElementAssemblyValues.hpp
		{
		public:
			std::vector<AssemblyValues> basis_values;
			std::vector<Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, 0, 3, 3>> jac_it;
			Eigen::MatrixXd val;
			Eigen::VectorXd det;
			void finalize3d(const ElementBases &gbasis, const std::vector<AssemblyValues> &gbasis_values);
		};

#include "../_ref/geom_extracted.inc"
	} // namespace assembler
} // namespace polyfem

using namespace polyfem;

extern "C"
{
	// vertices[4][3] (P1 geometry: gbases of a straight tet), ref_grads[n_qp][n_loc][3] ->
	// det[n_qp], jac_it[n_qp][9] row-major, grad_t_m[n_qp][n_loc][3]
	int ref_finalize3d(int n_loc, int n_qp, const double *vertices, const double *ref_grads, double *det, double *jac_it, double *grad_t_m)
	{
		static const double gg[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}}; // P1 reference gradients (auto_p_bases.cpp:1102-1124)
		basis::ElementBases gbasis;
		std::vector<assembler::AssemblyValues> gvals(4);
		gbasis.bases.resize(4);
		for (int j = 0; j < 4; ++j)
		{
			basis::Local2Global l2g{j, 1.0, Eigen::Dense(1, 3)};
			for (int c = 0; c < 3; ++c)
				l2g.node(0, c) = vertices[j * 3 + c];
			gbasis.bases[j].g = {l2g};
			gvals[j].grad.resize(n_qp, 3);
			for (int q = 0; q < n_qp; ++q)
				for (int c = 0; c < 3; ++c)
					gvals[j].grad(q, c) = gg[j][c];
		}
		assembler::ElementAssemblyValues vals;
		vals.val.resize(n_qp, 3);
		vals.basis_values.resize(n_loc);
		for (int i = 0; i < n_loc; ++i)
		{
			vals.basis_values[i].grad.resize(n_qp, 3);
			for (int q = 0; q < n_qp; ++q)
				for (int c = 0; c < 3; ++c)
					vals.basis_values[i].grad(q, c) = ref_grads[(size_t(q) * n_loc + i) * 3 + c];
		}
		vals.finalize3d(gbasis, gvals);
		for (int q = 0; q < n_qp; ++q)
		{
			det[q] = vals.det(q);
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					jac_it[size_t(q) * 9 + r * 3 + c] = vals.jac_it[q](r, c);
			for (int i = 0; i < n_loc; ++i)
				for (int c = 0; c < 3; ++c)
					grad_t_m[(size_t(q) * n_loc + i) * 3 + c] = vals.basis_values[i].grad_t_m(q, c);
		}
		return 0;
	}

	// isoparametric geometry: n_geom geometric bases with nodes geom_nodes[n_geom][3] and reference gradients
	// geom_grads[n_qp][n_geom][3] (e.g. the P2 basis of a curved tet) -> the same outputs
	int ref_finalize3d_iso(int n_loc, int n_qp, int n_geom, const double *geom_nodes, const double *geom_grads, const double *ref_grads, double *det,
						   double *jac_it, double *grad_t_m)
	{
		basis::ElementBases gbasis;
		std::vector<assembler::AssemblyValues> gvals(n_geom);
		gbasis.bases.resize(n_geom);
		for (int j = 0; j < n_geom; ++j)
		{
			basis::Local2Global l2g{j, 1.0, Eigen::Dense(1, 3)};
			for (int c = 0; c < 3; ++c)
				l2g.node(0, c) = geom_nodes[j * 3 + c];
			gbasis.bases[j].g = {l2g};
			gvals[j].grad.resize(n_qp, 3);
			for (int q = 0; q < n_qp; ++q)
				for (int c = 0; c < 3; ++c)
					gvals[j].grad(q, c) = geom_grads[(size_t(q) * n_geom + j) * 3 + c];
		}
		assembler::ElementAssemblyValues vals;
		vals.val.resize(n_qp, 3);
		vals.basis_values.resize(n_loc);
		for (int i = 0; i < n_loc; ++i)
		{
			vals.basis_values[i].grad.resize(n_qp, 3);
			for (int q = 0; q < n_qp; ++q)
				for (int c = 0; c < 3; ++c)
					vals.basis_values[i].grad(q, c) = ref_grads[(size_t(q) * n_loc + i) * 3 + c];
		}
		vals.finalize3d(gbasis, gvals);
		for (int q = 0; q < n_qp; ++q)
		{
			det[q] = vals.det(q);
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					jac_it[size_t(q) * 9 + r * 3 + c] = vals.jac_it[q](r, c);
			for (int i = 0; i < n_loc; ++i)
				for (int c = 0; c < 3; ++c)
					grad_t_m[(size_t(q) * n_loc + i) * 3 + c] = vals.basis_values[i].grad_t_m(q, c);
		}
		return 0;
	}
}
