// TEST INFRASTRUCTURE: declaration-only stand-ins for the few PolyFEM / Eigen types that
// polyfem_b200/host/assembler_shim.hpp touches, so that the shim can be syntax- and type-checked
// (g++ -fsyntax-only, tests/test_shim_syntax.py) in an image that has neither Eigen nor PolyFEM.
// Signatures are transcribed from the reference headers cited next to each declaration; nothing
// here has a body and nothing here ships.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace Eigen
{
	struct RowStub
	{
		double operator()(int) const;
	};
	struct MatrixXd
	{
		MatrixXd();
		MatrixXd(long, long);
		double *data();
		const double *data() const;
		long size() const;
		void resize(long, long);
		double operator()(long, long) const;
		RowStub row(long) const;
	};
	struct VectorXd
	{
		VectorXd();
		explicit VectorXd(long);
		double *data();
		const double *data() const;
		long size() const;
		void resize(long);
		double operator()(long) const;
	};
	struct Matrix3d
	{
		double operator()(int, int) const;
	};
	template <typename T>
	struct Map;
} // namespace Eigen

namespace polyfem
{
	struct json; // nlohmann::json (utils/Types.hpp / Common.hpp), only passed through by reference
	class Units; // polyfem/Units.hpp
	// utils/Types.hpp:24  typedef Eigen::SparseMatrix<double, Eigen::ColMajor> StiffnessMatrix;
	struct StiffnessMatrix
	{
		StiffnessMatrix &operator=(const Eigen::Map<const StiffnessMatrix> &);
	};
	// utils/Logger.hpp:42-49
	template <typename... Args>
	[[noreturn]] void log_and_throw_error(const std::string &msg, const Args &...args);

	namespace utils
	{
		class MatrixCache // utils/MatrixCache.hpp
		{
		};
	} // namespace utils

	namespace basis
	{
		struct Local2Global // basis/Basis.hpp:21-38
		{
			int index;
			double val;
			Eigen::RowStub node;
		};
		struct Basis // basis/Basis.hpp:43-...
		{
			const std::vector<Local2Global> &global() const;
		};
		struct ElementBases // basis/ElementBases.hpp:16-114
		{
			std::vector<Basis> bases;
		};
	} // namespace basis

	namespace quadrature
	{
		struct Quadrature
		{
			Eigen::MatrixXd points;
			Eigen::VectorXd weights;
		};
	} // namespace quadrature

	namespace assembler
	{
		struct AssemblyValues // assembler/AssemblyValues.hpp
		{
			std::vector<basis::Local2Global> global;
			Eigen::VectorXd val;
			Eigen::MatrixXd grad;
		};
		struct ElementAssemblyValues // assembler/ElementAssemblyValues.hpp:12-61
		{
			std::vector<AssemblyValues> basis_values;
			std::vector<Eigen::Matrix3d> jac_it;
			quadrature::Quadrature quadrature;
			int element_id;
			Eigen::MatrixXd val;
			Eigen::VectorXd det;
		};
		class AssemblyValsCache // assembler/AssemblyValsCache.hpp:25
		{
		public:
			void compute(const int el_index, const bool is_volume, const basis::ElementBases &basis, const basis::ElementBases &gbasis, ElementAssemblyValues &vals) const;
		};
		struct LameParameters // assembler/MatParams.hpp:83
		{
			void lambda_mu(const Eigen::RowStub &param, const Eigen::RowStub &p, double t, int el_id, double &lambda, double &mu) const;
		};
		struct Density // assembler/MatParams.hpp (Mass.cpp:13 call form)
		{
			double operator()(const Eigen::RowStub &param, const Eigen::RowStub &p, double t, int el_id) const;
		};

		class Assembler // assembler/Assembler.hpp:53-204
		{
		public:
			virtual ~Assembler() = default;
			virtual std::string name() const = 0;
			int size() const;
			virtual void set_size(const int size);                                                                                       // Assembler.hpp:64
			virtual void add_multimaterial(const int index, const json &params, const Units &units, const std::string &root_path);       // Assembler.hpp:190
			virtual void assemble(const bool is_volume, const int n_basis, const std::vector<basis::ElementBases> &bases,
								  const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const double t,
								  StiffnessMatrix &stiffness, const bool is_mass = false) const;
			virtual double assemble_energy(const bool is_volume, const std::vector<basis::ElementBases> &bases,
										   const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const double t,
										   const double dt, const Eigen::MatrixXd &displacement, const Eigen::MatrixXd &displacement_prev) const;
			virtual Eigen::VectorXd assemble_energy_per_element(const bool is_volume, const std::vector<basis::ElementBases> &bases,
																const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache,
																const double t, const double dt, const Eigen::MatrixXd &displacement,
																const Eigen::MatrixXd &displacement_prev) const;
			virtual void assemble_gradient(const bool is_volume, const int n_basis, const std::vector<basis::ElementBases> &bases,
										   const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const double t,
										   const double dt, const Eigen::MatrixXd &displacement, const Eigen::MatrixXd &displacement_prev,
										   Eigen::MatrixXd &rhs) const;
			virtual void assemble_hessian(const bool is_volume, const int n_basis, const bool project_to_psd,
										  const std::vector<basis::ElementBases> &bases, const std::vector<basis::ElementBases> &gbases,
										  const AssemblyValsCache &cache, const double t, const double dt, const Eigen::MatrixXd &displacement,
										  const Eigen::MatrixXd &displacement_prev, utils::MatrixCache &mat_cache, StiffnessMatrix &grad) const;
		};
		class LinearAssembler : virtual public Assembler // Assembler.hpp:208
		{
		public:
			void assemble(const bool is_volume, const int n_basis, const std::vector<basis::ElementBases> &bases,
						  const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const double t,
						  StiffnessMatrix &stiffness, const bool is_mass = false) const override;
		};
		class NLAssembler : virtual public Assembler // Assembler.hpp:238
		{
		public:
			double assemble_energy(const bool is_volume, const std::vector<basis::ElementBases> &bases,
								   const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const double t,
								   const double dt, const Eigen::MatrixXd &displacement, const Eigen::MatrixXd &displacement_prev) const override;
			Eigen::VectorXd assemble_energy_per_element(const bool is_volume, const std::vector<basis::ElementBases> &bases,
														const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache,
														const double t, const double dt, const Eigen::MatrixXd &displacement,
														const Eigen::MatrixXd &displacement_prev) const override;
			void assemble_gradient(const bool is_volume, const int n_basis, const std::vector<basis::ElementBases> &bases,
								   const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const double t,
								   const double dt, const Eigen::MatrixXd &displacement, const Eigen::MatrixXd &displacement_prev,
								   Eigen::MatrixXd &rhs) const override;
			void assemble_hessian(const bool is_volume, const int n_basis, const bool project_to_psd,
								  const std::vector<basis::ElementBases> &bases, const std::vector<basis::ElementBases> &gbases,
								  const AssemblyValsCache &cache, const double t, const double dt, const Eigen::MatrixXd &displacement,
								  const Eigen::MatrixXd &displacement_prev, utils::MatrixCache &mat_cache, StiffnessMatrix &grad) const override;
		};
		class ElasticityAssembler : virtual public Assembler // Assembler.hpp:301-372
		{
		protected:
			bool use_robust_jacobian = false; // :371
		};
		class ElasticityNLAssembler : virtual public ElasticityAssembler, virtual public NLAssembler // :374
		{
		};
		class NeoHookeanElasticity : public ElasticityNLAssembler // NeoHookeanElasticity.hpp:11
		{
		public:
			std::string name() const override;
			void add_multimaterial(const int index, const json &params, const Units &units, const std::string &root_path) override; // NeoHookeanElasticity.hpp:53
			const LameParameters &lame_params() const; // :56
		};
		class SaintVenantElasticity : public ElasticityNLAssembler // SaintVenantElasticity.hpp:11-53
		{
		public:
			std::string name() const override;
			void set_size(const int size) override;                                                                                // :26
			double stifness_tensor(int i, int j) const;                                                                            // :29
			void add_multimaterial(const int index, const json &params, const Units &units, const std::string &root_path) override; // :31
		};
		struct GenericMatParam // assembler/MatParams.hpp (call form of MooneyRivlinElasticity.hpp:32-34)
		{
			double operator()(const Eigen::RowStub &p, double t, int el_id) const;
		};
		class MooneyRivlinElasticity : public ElasticityNLAssembler // MooneyRivlinElasticity.hpp:9-54 (GenericElastic<...>)
		{
		public:
			std::string name() const override;
			void add_multimaterial(const int index, const json &params, const Units &units, const std::string &root_path) override; // :16
			const GenericMatParam &c1() const;                                                                                     // :18
			const GenericMatParam &c2() const;
			const GenericMatParam &k() const;
		};
		class FixedCorotational : public ElasticityNLAssembler // FixedCorotational.hpp:11-100
		{
		public:
			std::string name() const override;
			void add_multimaterial(const int index, const json &params, const Units &units, const std::string &root_path) override; // :40
			const LameParameters &lame_params() const;                                                                             // :43
		};
		class ViscousDamping : public NLAssembler // ViscousDamping.hpp:10-66
		{
		public:
			std::string name() const override;
			void add_multimaterial(const int index, const json &params, const Units &units, const std::string &root_path) override; // :28
			double get_psi() const;                                                                                                // :49
			double get_phi() const;
		};
		class LinearElasticity : public LinearAssembler, public ElasticityNLAssembler // LinearElasticity.hpp
		{
		public:
			std::string name() const override;
			const LameParameters &lame_params() const; // :63
		};
		class Laplacian : public LinearAssembler // Laplacian.hpp
		{
		public:
			std::string name() const override;
		};
		class Mass : public LinearAssembler // Mass.hpp:9-37
		{
		public:
			std::string name() const override;
			const Density &density() const; // :28
		};
	} // namespace assembler
} // namespace polyfem

namespace Eigen
{
	template <>
	struct Map<const polyfem::StiffnessMatrix> // Eigen::Map<const SparseMatrix>(rows, cols, nnz, outer, inner, values)
	{
		Map(long rows, long cols, long nnz, const int *outer, const int *inner, const double *values);
	};
} // namespace Eigen
