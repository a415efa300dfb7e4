// TEST INFRASTRUCTURE: the few members of the reference's assembler types that the function bodies extracted by
// extract_nh.py touch, shared by nh_glue.cpp (local functions) and loop_glue.cpp (global loops). Ends inside
// namespace polyfem::assembler so that the including file can #include the extracted bodies there.
#pragma once
#include "mini_eigen.hpp"

#include <cmath>
#include <vector>

using std::log;

namespace polyfem::utils
{
	// named (qualified) in the autodiff branch of compute_energy_aux, which is never instantiated here
	template <typename M>
	double determinant(const M &m) { return m.determinant(); }
} // namespace polyfem::utils

namespace polyfem::assembler
{
	// the few members of the reference types that the extracted functions touch
	struct Local2Global // basis/Basis.hpp:21-38
	{
		int index;
		double val;
	};
	struct AssemblyValues // assembler/AssemblyValues.hpp
	{
		std::vector<Local2Global> global;
		Eigen::MatrixXd grad;     // n_qp x dim reference gradients
		Eigen::MatrixXd grad_t_m; // n_qp x dim physical gradients (grad * jac_it)
		Eigen::VectorXd val;      // n_qp basis values
	};
	struct Quadrature // quadrature/Quadrature.hpp
	{
		Eigen::MatrixXd points;
		Eigen::VectorXd weights;
	};
	struct ElementAssemblyValues // assembler/ElementAssemblyValues.hpp:12-61
	{
		std::vector<AssemblyValues> basis_values;
		std::vector<Eigen::MatrixXd> jac_it;
		Quadrature quadrature;
		Eigen::MatrixXd val;
		Eigen::VectorXd det; // n_qp Jacobian determinants
		int element_id = 0;
		Eigen::VectorXd eval_deformed_jacobian_determinant(const Eigen::MatrixXd &) const { return Eigen::VectorXd(); }
	};
	struct NonLinearAssemblerData // assembler/AssemblerData.hpp
	{
		const ElementAssemblyValues &vals;
		double t;
		double dt;
		const Eigen::MatrixXd &x;
		const Eigen::MatrixXd &x_prev;
		const Eigen::VectorXd &da;
		NonLinearAssemblerData(const ElementAssemblyValues &vals_, double t_, double dt_, const Eigen::MatrixXd &x_,
							   const Eigen::MatrixXd &x_prev_, const Eigen::VectorXd &da_)
			: vals(vals_), t(t_), dt(dt_), x(x_), x_prev(x_prev_), da(da_) {}
	};
	struct LameParameters // assembler/MatParams.hpp:83
	{
		double lambda = 0, mu = 0;
		void lambda_mu(const Eigen::Dense &, const Eigen::Dense &, double, int, double &l, double &m) const
		{
			l = lambda;
			m = mu;
		}
	};

	struct LinearAssemblerData // assembler/AssemblerData.hpp
	{
		const ElementAssemblyValues &vals;
		double t;
		int i, j;
		const Eigen::VectorXd &da;
		LinearAssemblerData(const ElementAssemblyValues &vals_, double t_, int i_, int j_, const Eigen::VectorXd &da_)
			: vals(vals_), t(t_), i(i_), j(j_), da(da_) {}
	};
	struct Density // assembler/MatParams.hpp (call form of Mass.cpp:13)
	{
		double rho = 1;
		double operator()(const Eigen::Dense &, const Eigen::Dense &, double, int) const { return rho; }
	};
#ifdef PFREF_LOOPS
	// loop_glue.cpp (which has included the reference's utils/MatrixCache.hpp): what the global loops of
	// NLAssembler (Assembler.cpp:495-771) name besides the types above
} // namespace polyfem::assembler
namespace igl
{
	struct Timer
	{
		void start() {}
		void stop() {}
		double getElapsedTime() const { return 0.0; }
	};
} // namespace igl
namespace ipc
{
	Eigen::MatrixXd project_to_psd(const Eigen::MatrixXd &m); // loop_glue.cpp; only reached with project_to_psd = true
}
namespace polyfem
{
	constexpr int MAX_QUAD_POINTS = -1;                                           // the default build (utils/Types.hpp)
	typedef Eigen::Matrix<double, Eigen::Dynamic, 1, 0, MAX_QUAD_POINTS, 1> QuadratureVector; // utils/Types.hpp
	namespace basis
	{
		struct ElementBases // the loops only hand bases[e] / gbases[e] to AssemblyValsCache::compute
		{
		};
	} // namespace basis
	namespace quadrature
	{
	}
} // namespace polyfem
namespace polyfem::assembler
{
	// AssemblyValsCache::compute (assembler/AssemblyValsCache.cpp:33-40) with a filled cache: vals = cache[el_index]
	class AssemblyValsCache
	{
	public:
		std::vector<ElementAssemblyValues> cache;
		bool is_mass_ = false;
		bool is_mass() const { return is_mass_; }
		void compute(const int el_index, const bool, const basis::ElementBases &, const basis::ElementBases &, ElementAssemblyValues &vals) const
		{
			vals = cache[size_t(el_index)];
		}
	};
	class NLAssembler // assembler/Assembler.hpp:247-330: the members the three loops use
	{
	public:
		virtual ~NLAssembler() = default;
		virtual int size() const = 0;
		double assemble_energy(const bool is_volume, const std::vector<basis::ElementBases> &bases, const std::vector<basis::ElementBases> &gbases,
							   const AssemblyValsCache &cache, const double t, const double dt, const Eigen::MatrixXd &displacement,
							   const Eigen::MatrixXd &displacement_prev) const;
		Eigen::VectorXd assemble_energy_per_element(const bool is_volume, const std::vector<basis::ElementBases> &bases,
													const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const double t,
													const double dt, const Eigen::MatrixXd &displacement, const Eigen::MatrixXd &displacement_prev) const;
		void assemble_gradient(const bool is_volume, const int n_basis, const std::vector<basis::ElementBases> &bases,
							   const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const double t, const double dt,
							   const Eigen::MatrixXd &displacement, const Eigen::MatrixXd &displacement_prev, Eigen::MatrixXd &rhs) const;
		void assemble_hessian(const bool is_volume, const int n_basis, const bool project_to_psd, const std::vector<basis::ElementBases> &bases,
							  const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const double t, const double dt,
							  const Eigen::MatrixXd &displacement, const Eigen::MatrixXd &displacement_prev, utils::MatrixCache &mat_cache,
							  StiffnessMatrix &hess) const;

	protected:
		virtual double compute_energy(const NonLinearAssemblerData &data) const = 0;
		virtual Eigen::VectorXd assemble_gradient(const NonLinearAssemblerData &data) const = 0;
		virtual Eigen::MatrixXd assemble_hessian(const NonLinearAssemblerData &data) const = 0;
	};
	class LinearAssembler // assembler/Assembler.hpp:120-170: the members the linear loop uses
	{
	public:
		virtual ~LinearAssembler() = default;
		virtual int size() const = 0;
		void assemble(const bool is_volume, const int n_basis, const std::vector<basis::ElementBases> &bases,
					  const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const double t,
					  StiffnessMatrix &stiffness, const bool is_mass = false) const;
		virtual Eigen::Matrix<double, Eigen::Dynamic, 1, 0, 9, 1> assemble(const LinearAssemblerData &data) const = 0;
	};
#define PFREF_NL_BASE : public NLAssembler
#define PFREF_LIN_BASE : public LinearAssembler
#define PFREF_LIN_USING using LinearAssembler::assemble;
#define PFREF_OVERRIDE override
#else
#define PFREF_NL_BASE
#define PFREF_LIN_BASE
#define PFREF_LIN_USING
#define PFREF_OVERRIDE
#endif
	class LinearElasticity PFREF_LIN_BASE
	{
	public:
		PFREF_LIN_USING
		int size() const { return 3; }
		LameParameters params_;
		Eigen::Matrix<double, Eigen::Dynamic, 1, 0, 9, 1> assemble(const LinearAssemblerData &data) const PFREF_OVERRIDE;
		template <typename T>
		T compute_energy_aux(const NonLinearAssemblerData &data) const; // LinearElasticity.cpp:103-132
		// the autodiff gradient / Hessian of that energy (LinearElasticity.cpp:70-101), defined only by sv_glue.cpp
		Eigen::VectorXd assemble_gradient(const NonLinearAssemblerData &data) const;
		Eigen::MatrixXd assemble_hessian(const NonLinearAssemblerData &data) const;
	};
	class Laplacian PFREF_LIN_BASE
	{
	public:
		PFREF_LIN_USING
		int size() const { return 1; }
		Eigen::Matrix<double, Eigen::Dynamic, 1, 0, 9, 1> assemble(const LinearAssemblerData &data) const PFREF_OVERRIDE;
	};
	class Mass PFREF_LIN_BASE
	{
	public:
		PFREF_LIN_USING
		int size() const { return 3; }
		Density density_;
		Eigen::Matrix<double, Eigen::Dynamic, 1, 0, 9, 1> assemble(const LinearAssemblerData &data) const PFREF_OVERRIDE;
	};

	class NeoHookeanElasticity PFREF_NL_BASE
	{
	public:
		int size() const { return 3; }
		bool use_robust_jacobian = false;
		LameParameters params_;
		template <typename T, int n_basis, int dim>
		T compute_energy_aux(const NonLinearAssemblerData &data) const;
		template <int n_basis, int dim>
		void compute_energy_aux_gradient_fast(const NonLinearAssemblerData &data, Eigen::Matrix<double, Eigen::Dynamic, 1> &G_flattened) const;
		template <int n_basis, int dim>
		void compute_energy_hessian_aux_fast(const NonLinearAssemblerData &data, Eigen::MatrixXd &H) const;
#ifdef PFREF_LOOPS
		using NLAssembler::assemble_gradient;
		using NLAssembler::assemble_hessian;
		// the reference's dispatch on the number of bases (NeoHookeanElasticity.cpp:50-113, 177-252, 304-336), 3-D cases
		double compute_energy(const NonLinearAssemblerData &data) const override
		{
			switch (data.vals.basis_values.size())
			{
			case 4: return compute_energy_aux<double, 4, 3>(data);
			case 10: return compute_energy_aux<double, 10, 3>(data);
			case 20: return compute_energy_aux<double, 20, 3>(data);
			default: return compute_energy_aux<double, Eigen::Dynamic, 3>(data);
			}
		}
		Eigen::VectorXd assemble_gradient(const NonLinearAssemblerData &data) const override
		{
			Eigen::Matrix<double, Eigen::Dynamic, 1> g;
			switch (data.vals.basis_values.size())
			{
			case 4: g.resize(12, 1); compute_energy_aux_gradient_fast<4, 3>(data, g); break;
			case 10: g.resize(30, 1); compute_energy_aux_gradient_fast<10, 3>(data, g); break;
			case 20: g.resize(60, 1); compute_energy_aux_gradient_fast<20, 3>(data, g); break;
			default: g.resize(long(data.vals.basis_values.size()) * 3, 1); compute_energy_aux_gradient_fast<Eigen::Dynamic, 3>(data, g);
			}
			return g;
		}
		Eigen::MatrixXd assemble_hessian(const NonLinearAssemblerData &data) const override
		{
			const long N = long(data.vals.basis_values.size()) * 3;
			Eigen::MatrixXd H(N, N); // resize + setZero
			switch (data.vals.basis_values.size())
			{
			case 4: compute_energy_hessian_aux_fast<4, 3>(data, H); break;
			case 10: compute_energy_hessian_aux_fast<10, 3>(data, H); break;
			case 20: compute_energy_hessian_aux_fast<20, 3>(data, H); break;
			default: compute_energy_hessian_aux_fast<Eigen::Dynamic, 3>(data, H);
			}
			return H;
		}
#endif
	};

