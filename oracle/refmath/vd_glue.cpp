// TEST INFRASTRUCTURE: C entry points over the reference's OWN ViscousDamping function bodies
// (assembler/ViscousDamping.cpp:5-62, 122-229, 297-342), extracted at build time into ../_ref/vd_extracted.inc and compiled
// verbatim against mini_eigen.hpp. Used by tools/make_golden.py to write tests/golden/vd_local.npz and by
// tests/test_oracle_viscous_reference.py to pin oracle/oracle.cpp's restatement against the reference itself.
#include "nh_harness.hpp" // opens namespace polyfem::assembler

	struct DampingParameters // assembler/MatParams.hpp: two doubles (psi, phi) behind operator[]
	{
		double v[2] = {0, 0};
		double &operator[](int i) { return v[i]; }
		double operator[](int i) const { return v[i]; }
	};
	class ViscousDamping // assembler/ViscousDamping.hpp:10-66: the members the extracted bodies use
	{
	public:
		int size() const { return 3; }
		double compute_energy(const NonLinearAssemblerData &data) const;
		Eigen::MatrixXd assemble_hessian(const NonLinearAssemblerData &data) const;
		Eigen::VectorXd assemble_gradient(const NonLinearAssemblerData &data) const;
		DampingParameters damping_params_;

	protected:
		void compute_stress_aux(const Eigen::MatrixXd &F, const Eigen::MatrixXd &dFdt, Eigen::MatrixXd &dRdF, Eigen::MatrixXd &dRdFdot) const;
		void compute_stress_grad_aux(const Eigen::MatrixXd &F, const Eigen::MatrixXd &dFdt, Eigen::MatrixXd &d2RdF2, Eigen::MatrixXd &d2RdFdFdot, Eigen::MatrixXd &d2RdFdot2) const;
	};

#include "../_ref/vd_extracted.inc"
} // namespace polyfem::assembler

using namespace polyfem::assembler;

extern "C"
{
	// u / u_prev [n_basis][3] nodal displacements (u_prev NULL: a previous displacement of another size, the first step),
	// grads [n_qp][n_basis][3] reference gradients, jac_it [n_qp][9] row-major, da [n_qp]; out: energy, gradient [N] node-major,
	// hessian [N][N] row-major, N = 3 n_basis
	int ref_vd_local(int n_basis, int n_qp, const double *u, const double *u_prev, const double *grads, const double *jac_it, const double *da,
					 double dt, double psi, double phi, double *energy, double *gradient, double *hessian)
	{
		ElementAssemblyValues vals;
		Eigen::MatrixXd x(long(n_basis) * 3, 1), x_prev(u_prev ? long(n_basis) * 3 : 0, u_prev ? 1 : 0);
		Eigen::VectorXd dav(n_qp, 1);
		for (int k = 0; k < n_basis * 3; ++k)
		{
			x(k) = u[k];
			if (u_prev)
				x_prev(k) = u_prev[k];
		}
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
		ViscousDamping vd;
		vd.damping_params_[0] = psi;
		vd.damping_params_[1] = phi;
		const NonLinearAssemblerData data{vals, 0.0, dt, x, x_prev, dav};
		*energy = vd.compute_energy(data);
		const Eigen::VectorXd g = vd.assemble_gradient(data);
		const Eigen::MatrixXd H = vd.assemble_hessian(data);
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
