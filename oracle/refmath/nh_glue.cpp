// TEST INFRASTRUCTURE: C entry points over the reference's OWN NeoHookean gradient / Hessian function bodies
// (assembler/NeoHookeanElasticity.cpp:419-658), extracted at build time into ../_ref/nh_extracted.inc and compiled
// verbatim against mini_eigen.hpp. Used by tools/make_golden.py to write tests/golden/nh_local.npz and by
// tests/test_oracle_reference_math.py to pin oracle/oracle.cpp's local math against the reference itself.
#include "nh_harness.hpp" // opens namespace polyfem::assembler

#include "../_ref/nh_extracted.inc"
} // namespace polyfem::assembler

namespace polyfem
{
	// what utils/ElasticityUtils.hpp:70-136 names from utils/AutodiffTypes.hpp, for T = double
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
} // namespace polyfem

namespace polyfem::assembler
{
#include "../_ref/le_energy_extracted.inc"
} // namespace polyfem::assembler

using namespace polyfem::assembler;

namespace
{
	struct Inputs
	{
		ElementAssemblyValues vals;
		Eigen::MatrixXd x, x_prev;
		Eigen::VectorXd da;
	};

	// u[n_basis][3] nodal displacements, grads[n_qp][n_basis][3], jac_it[n_qp][9] row-major, da[n_qp]
	void fill(Inputs &in, int n_basis, int n_qp, const double *u, const double *grads, const double *jac_it, const double *da)
	{
		in.x.resize(long(n_basis) * 3, 1);
		for (int k = 0; k < n_basis * 3; ++k)
			in.x(k) = u[k];
		in.x_prev.resize(long(n_basis) * 3, 1);
		in.da.resize(n_qp, 1);
		in.vals.quadrature.points.resize(n_qp, 3);
		in.vals.val.resize(n_qp, 3);
		in.vals.basis_values.resize(n_basis);
		for (int i = 0; i < n_basis; ++i)
		{
			in.vals.basis_values[i].global = {Local2Global{i, 1.0}};
			in.vals.basis_values[i].grad.resize(n_qp, 3);
			for (int q = 0; q < n_qp; ++q)
				for (int c = 0; c < 3; ++c)
					in.vals.basis_values[i].grad(q, c) = grads[(size_t(q) * n_basis + i) * 3 + c];
		}
		in.vals.jac_it.resize(n_qp);
		for (int q = 0; q < n_qp; ++q)
		{
			in.da(q) = da[q];
			in.vals.jac_it[q].resize(3, 3);
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					in.vals.jac_it[q](r, c) = jac_it[size_t(q) * 9 + r * 3 + c];
		}
	}
} // namespace

extern "C"
{
	// NeoHookeanElasticity::compute_energy -> compute_energy_aux<double, n_basis, dim> (NeoHookeanElasticity.cpp:304-388)
	int ref_nh_energy(int n_basis, int n_qp, const double *u, const double *grads, const double *jac_it, const double *da, double lambda, double mu, double *out)
	{
		Inputs in;
		fill(in, n_basis, n_qp, u, grads, jac_it, da);
		NeoHookeanElasticity nh;
		nh.params_.lambda = lambda;
		nh.params_.mu = mu;
		const NonLinearAssemblerData data{in.vals, 0.0, 1.0, in.x, in.x_prev, in.da};
		if (n_basis == 4)
			*out = nh.compute_energy_aux<double, 4, 3>(data);
		else if (n_basis == 10)
			*out = nh.compute_energy_aux<double, 10, 3>(data);
		else if (n_basis == 20)
			*out = nh.compute_energy_aux<double, 20, 3>(data);
		else
			*out = nh.compute_energy_aux<double, Eigen::Dynamic, 3>(data);
		return 0;
	}

	// out[n_basis*3], node-major (NeoHookeanElasticity.cpp:540-544). The instantiation follows the reference's own
	// dispatch on the number of bases (assemble_gradient: 4, 10, 20 fixed, anything else dynamic).
	int ref_nh_gradient(int n_basis, int n_qp, const double *u, const double *grads, const double *jac_it, const double *da, double lambda, double mu, double *out)
	{
		Inputs in;
		fill(in, n_basis, n_qp, u, grads, jac_it, da);
		NeoHookeanElasticity nh;
		nh.params_.lambda = lambda;
		nh.params_.mu = mu;
		const NonLinearAssemblerData data{in.vals, 0.0, 1.0, in.x, in.x_prev, in.da};
		Eigen::Matrix<double, Eigen::Dynamic, 1> g;
		if (n_basis == 4)
			nh.compute_energy_aux_gradient_fast<4, 3>(data, g);
		else if (n_basis == 10)
			nh.compute_energy_aux_gradient_fast<10, 3>(data, g);
		else if (n_basis == 20)
			nh.compute_energy_aux_gradient_fast<20, 3>(data, g);
		else
			nh.compute_energy_aux_gradient_fast<Eigen::Dynamic, 3>(data, g);
		if (g.size() != long(n_basis) * 3)
			return -1;
		for (int k = 0; k < n_basis * 3; ++k)
			out[k] = g(k);
		return 0;
	}

	// out[N*N] row-major, N = n_basis*3, H(i*3+a, j*3+b) (NeoHookeanElasticity.cpp:652-656)
	int ref_nh_hessian(int n_basis, int n_qp, const double *u, const double *grads, const double *jac_it, const double *da, double lambda, double mu, double *out)
	{
		Inputs in;
		fill(in, n_basis, n_qp, u, grads, jac_it, da);
		NeoHookeanElasticity nh;
		nh.params_.lambda = lambda;
		nh.params_.mu = mu;
		const NonLinearAssemblerData data{in.vals, 0.0, 1.0, in.x, in.x_prev, in.da};
		const long N = long(n_basis) * 3;
		Eigen::MatrixXd H(N, N); // the caller zero-initialises it (Assembler.cpp / assemble_hessian: hessian.setZero())
		if (n_basis == 4)
			nh.compute_energy_hessian_aux_fast<4, 3>(data, H);
		else if (n_basis == 10)
			nh.compute_energy_hessian_aux_fast<10, 3>(data, H);
		else if (n_basis == 20)
			nh.compute_energy_hessian_aux_fast<20, 3>(data, H);
		else
			nh.compute_energy_hessian_aux_fast<Eigen::Dynamic, 3>(data, H);
		for (long r = 0; r < N; ++r)
			for (long c = 0; c < N; ++c)
				out[r * N + c] = H(r, c);
		return 0;
	}

	// LinearElasticity::compute_energy -> compute_energy_aux<double> (LinearElasticity.cpp:65-68, 103-132): the energy of a
	// linear material inside a nonlinear solve, same inputs as ref_nh_energy
	int ref_le_energy(int n_basis, int n_qp, const double *u, const double *grads, const double *jac_it, const double *da, double lambda, double mu, double *out)
	{
		Inputs in;
		fill(in, n_basis, n_qp, u, grads, jac_it, da);
		LinearElasticity le;
		le.params_.lambda = lambda;
		le.params_.mu = mu;
		const NonLinearAssemblerData data{in.vals, 0.0, 1.0, in.x, in.x_prev, in.da};
		*out = le.compute_energy_aux<double>(data);
		return 0;
	}

	// ---- linear local blocks: res[size*size] with index n*size + m (LinearAssembler::assemble, Assembler.cpp:228-236) ----
	// gi / gj: grad_t_m of bases i and j, [n_qp][3]; vi / vj: basis values [n_qp]
	static void fill_pair(ElementAssemblyValues &vals, Eigen::VectorXd &dav, int n_qp, const double *gi, const double *gj, const double *vi, const double *vj, const double *da)
	{
		vals.basis_values.resize(2);
		vals.quadrature.points.resize(n_qp, 3);
		vals.val.resize(n_qp, 3);
		dav.resize(n_qp, 1);
		for (int b = 0; b < 2; ++b)
		{
			vals.basis_values[b].grad_t_m.resize(n_qp, 3);
			vals.basis_values[b].val.resize(n_qp, 1);
		}
		for (int q = 0; q < n_qp; ++q)
		{
			dav(q) = da[q];
			for (int c = 0; c < 3; ++c)
			{
				vals.basis_values[0].grad_t_m(q, c) = gi ? gi[q * 3 + c] : 0.0;
				vals.basis_values[1].grad_t_m(q, c) = gj ? gj[q * 3 + c] : 0.0;
			}
			vals.basis_values[0].val(q) = vi ? vi[q] : 0.0;
			vals.basis_values[1].val(q) = vj ? vj[q] : 0.0;
		}
	}

	int ref_linear_elasticity_block(int n_qp, const double *gi, const double *gj, const double *da, double lambda, double mu, double *out9)
	{
		ElementAssemblyValues vals;
		Eigen::VectorXd dav;
		fill_pair(vals, dav, n_qp, gi, gj, nullptr, nullptr, da);
		LinearElasticity le;
		le.params_.lambda = lambda;
		le.params_.mu = mu;
		const auto res = le.assemble(LinearAssemblerData{vals, 0.0, 0, 1, dav});
		if (res.size() != 9)
			return -1;
		for (int k = 0; k < 9; ++k)
			out9[k] = res(k);
		return 0;
	}

	int ref_laplacian_block(int n_qp, const double *gi, const double *gj, const double *da, double *out1)
	{
		ElementAssemblyValues vals;
		Eigen::VectorXd dav;
		fill_pair(vals, dav, n_qp, gi, gj, nullptr, nullptr, da);
		Laplacian lap;
		const auto res = lap.assemble(LinearAssemblerData{vals, 0.0, 0, 1, dav});
		if (res.size() != 1)
			return -1;
		out1[0] = res(0);
		return 0;
	}

	int ref_mass_block(int n_qp, const double *vi, const double *vj, const double *da, double rho, double *out9)
	{
		ElementAssemblyValues vals;
		Eigen::VectorXd dav;
		fill_pair(vals, dav, n_qp, nullptr, nullptr, vi, vj, da);
		Mass mass;
		mass.density_.rho = rho;
		const auto res = mass.assemble(LinearAssemblerData{vals, 0.0, 0, 1, dav});
		if (res.size() != 9)
			return -1;
		for (int k = 0; k < 9; ++k)
			out9[k] = res(k);
		return 0;
	}
}
