// C entry points over the reference's own data-only sources (see refshim/Eigen/Dense).
// TEST INFRASTRUCTURE ONLY: built into oracle/_ref/libpfref.so by oracle/Makefile,
// used by tools/make_golden.py (here) and by tests to pin polyfem_b200/tables.py.
//
// Reference entry points wrapped (declared in the reference headers, compiled unmodified):
//   polyfem::autogen::p_nodes_3d / p_basis_value_3d / p_grad_basis_value_3d
//        (/root/reference/src/polyfem/autogen/auto_p_bases.hpp:14-18)
//   polyfem::quadrature::TetQuadrature::get_quadrature
//        (/root/reference/src/polyfem/quadrature/TetQuadrature.cpp:43-55)
#include <polyfem/autogen/auto_p_bases.hpp>
#include <polyfem/autogen/auto_b_bases.hpp>
#include <polyfem/autogen/p_n_bases.hpp>
#include <polyfem/quadrature/TetQuadrature.hpp>
#include <cstdlib>
#include <cstdio>

// The dispatchers reference Bernstein / arbitrary-order bases that live in other
// translation units; they are never reached for Lagrange P0..P4, so trap if they are.
namespace polyfem::autogen
{
	static void unreachable(const char *what)
	{
		std::fprintf(stderr, "pfref: %s is not part of the hot path\n", what);
		std::abort();
	}
	void b_basis_value_2d(const int, const int, const Eigen::MatrixXd &, Eigen::MatrixXd &) { unreachable("b_basis_value_2d"); }
	void b_grad_basis_value_2d(const int, const int, const Eigen::MatrixXd &, Eigen::MatrixXd &) { unreachable("b_grad_basis_value_2d"); }
	void b_basis_value_3d(const int, const int, const Eigen::MatrixXd &, Eigen::MatrixXd &) { unreachable("b_basis_value_3d"); }
	void b_grad_basis_value_3d(const int, const int, const Eigen::MatrixXd &, Eigen::MatrixXd &) { unreachable("b_grad_basis_value_3d"); }
	void p_n_nodes_2d(const int, Eigen::MatrixXd &) { unreachable("p_n_nodes_2d"); }
	void p_n_basis_value_2d(const int, const int, const Eigen::MatrixXd &, Eigen::MatrixXd &) { unreachable("p_n_basis_value_2d"); }
	void p_n_basis_grad_value_2d(const int, const int, const Eigen::MatrixXd &, Eigen::MatrixXd &) { unreachable("p_n_basis_grad_value_2d"); }
	void p_n_nodes_3d(const int, Eigen::MatrixXd &) { unreachable("p_n_nodes_3d"); }
	void p_n_basis_value_3d(const int, const int, const Eigen::MatrixXd &, Eigen::MatrixXd &) { unreachable("p_n_basis_value_3d"); }
	void p_n_basis_grad_value_3d(const int, const int, const Eigen::MatrixXd &, Eigen::MatrixXd &) { unreachable("p_n_basis_grad_value_3d"); }
} // namespace polyfem::autogen

extern "C"
{
	// number of tet quadrature points of `order`; fills pts[n*3] (row-major), w[n] (already /6)
	int pfref_tet_quadrature(int order, double *pts, double *w, int cap)
	{
		polyfem::quadrature::Quadrature q;
		polyfem::quadrature::TetQuadrature tq;
		tq.get_quadrature(order, q);
		const int n = int(q.points.rows());
		if (n > cap)
			return -n;
		for (int i = 0; i < n; ++i)
		{
			for (int d = 0; d < 3; ++d)
				pts[i * 3 + d] = q.points(i, d);
			w[i] = q.weights(i);
		}
		return n;
	}

	// reference-element node positions of the P_p tet, nodes[n*3] row-major
	int pfref_p_nodes_3d(int p, double *nodes, int cap)
	{
		Eigen::MatrixXd v;
		polyfem::autogen::p_nodes_3d(p, v);
		const int n = int(v.rows());
		if (n > cap)
			return -n;
		for (int i = 0; i < n; ++i)
			for (int d = 0; d < 3; ++d)
				nodes[i * 3 + d] = v(i, d);
		return n;
	}

	// val[n_pts], grad[n_pts*3] of local basis `li` of order p at pts[n_pts*3]
	void pfref_p_basis_3d(int p, int li, int n_pts, const double *pts, double *val, double *grad)
	{
		Eigen::MatrixXd uv(n_pts, 3), v, g;
		for (int i = 0; i < n_pts; ++i)
			for (int d = 0; d < 3; ++d)
				uv(i, d) = pts[i * 3 + d];
		polyfem::autogen::p_basis_value_3d(false, p, li, uv, v);
		polyfem::autogen::p_grad_basis_value_3d(false, p, li, uv, g);
		for (int i = 0; i < n_pts; ++i)
		{
			val[i] = v(i, 0);
			for (int d = 0; d < 3; ++d)
				grad[i * 3 + d] = g(i, d);
		}
	}
}
