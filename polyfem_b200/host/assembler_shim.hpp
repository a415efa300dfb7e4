// assembler_shim.hpp — PolyFEM-side binding of libpfa.so (header-only, C++17).
//
// This file is what a PolyFEM maintainer adds to the reference tree. It
// compiles only where PolyFEM + Eigen are available (they are not in this repo's offline image,
// where the same interface is mirrored by polyfem_b200/assembler.py and tested through the C ABI).
//
// It subclasses the reference assemblers and overrides the five virtuals of
// polyfem::assembler::Assembler (src/polyfem/assembler/Assembler.hpp:68-125) that
// solver::ElasticForm calls (src/polyfem/solver/forms/ElasticForm.cpp:287-328,411-418). Nothing
// above this boundary changes: name(), size(), add_multimaterial, the JSON material spec and the
// returned Eigen::SparseMatrix pattern are the reference's.
#pragma once

#include <pfa.h>

#include <polyfem/assembler/AssemblyValsCache.hpp>
#include <polyfem/assembler/Laplacian.hpp>
#include <polyfem/assembler/LinearElasticity.hpp>
#include <polyfem/assembler/NeoHookeanElasticity.hpp>
#include <polyfem/basis/ElementBases.hpp>
#include <polyfem/utils/Logger.hpp>
#include <polyfem/utils/MatrixCache.hpp>

#include <memory>
#include <vector>

namespace polyfem::assembler::b200
{
	/// Owns one pfa_handle; rebuilt when the FE space (bases pointer / size) changes.
	/// The reference's methods are const and called from one host thread, so the
	/// device state is `mutable` in the assemblers below.
	class DeviceAssembly
	{
	public:
		~DeviceAssembly() { reset(); }

		void reset()
		{
			if (h_)
				pfa_destroy(h_);
			h_ = nullptr;
			key_ = nullptr;
		}

		/// Reads what the hot path needs out of `bases` / `gbases` once per mesh.
		/// `lame(e, q, lambda, mu)` evaluates LameParameters::lambda_mu for element e at qp q.
		template <typename LameFn>
		pfa_handle *get(const pfa_material material, const bool is_volume, const int n_basis,
						const std::vector<basis::ElementBases> &bases,
						const std::vector<basis::ElementBases> &gbases,
						const AssemblyValsCache &cache, const LameFn &lame) const
		{
			if (h_ && key_ == bases.data() && n_elements_ == int(bases.size()))
				return h_;
			const_cast<DeviceAssembly *>(this)->reset();
			if (!is_volume)
				log_and_throw_error("B200 assembly path: only volumetric (tetrahedral) meshes are supported");

			ElementAssemblyValues vals;
			cache.compute(0, is_volume, bases[0], gbases[0], vals);
			const int n_el = int(bases.size());
			const int n_loc = int(vals.basis_values.size());
			const int n_qp = int(vals.quadrature.weights.size());

			std::vector<int32_t> conn(size_t(n_el) * n_loc);
			std::vector<double> jac_it(size_t(n_el) * n_qp * 9), da(size_t(n_el) * n_qp);
			std::vector<double> lambda(size_t(n_el) * n_qp), mu(size_t(n_el) * n_qp);
			std::vector<double> ref_grads(size_t(n_qp) * n_loc * 3), weights(n_qp);
			for (int q = 0; q < n_qp; ++q)
			{
				weights[q] = vals.quadrature.weights(q);
				for (int j = 0; j < n_loc; ++j)
					for (int c = 0; c < 3; ++c)
						ref_grads[(size_t(q) * n_loc + j) * 3 + c] = vals.basis_values[j].grad(q, c);
			}
			for (int e = 0; e < n_el; ++e)
			{
				cache.compute(e, is_volume, bases[e], gbases[e], vals);
				if (int(vals.basis_values.size()) != n_loc || int(vals.quadrature.weights.size()) != n_qp)
					log_and_throw_error("B200 assembly path: mixed element types / orders are not supported");
				for (int j = 0; j < n_loc; ++j)
				{
					const auto &g = vals.basis_values[j].global;
					if (g.size() != 1 || g[0].val != 1.0)
						log_and_throw_error("B200 assembly path: non-conforming bases (Local2Global lists) are not supported");
					conn[size_t(e) * n_loc + j] = g[0].index;
				}
				for (int q = 0; q < n_qp; ++q)
				{
					const Eigen::Matrix3d jit = vals.jac_it[q];
					for (int r = 0; r < 3; ++r)
						for (int c = 0; c < 3; ++c)
							jac_it[(size_t(e) * n_qp + q) * 9 + r * 3 + c] = jit(r, c);
					da[size_t(e) * n_qp + q] = vals.det(q) * vals.quadrature.weights(q);
					lame(vals, q, lambda[size_t(e) * n_qp + q], mu[size_t(e) * n_qp + q]);
				}
			}

			pfa_mesh_desc d{};
			d.struct_size = sizeof(pfa_mesh_desc);
			d.material = material;
			d.n_elements = n_el;
			d.n_loc = n_loc;
			d.n_bases = n_basis;
			d.n_qp = n_qp;
			d.conn = conn.data();
			d.quad_weights = weights.data();
			d.ref_grads = ref_grads.data();
			d.jac_it = jac_it.data(); // general form: works for affine and isoparametric geometry
			d.da = da.data();
			d.lambda = lambda.data();
			d.mu = mu.data();
			d.material_stride = n_qp;
			d.device = 0;
			if (pfa_create(&d, &h_) != PFA_OK)
				log_and_throw_error("B200 assembly path: {}", pfa_last_error(nullptr));
			key_ = bases.data();
			n_elements_ = n_el;
			return h_;
		}

		/// Wraps values[] in the reference's matrix type (pattern identical to SparseMatrixCache's).
		static void to_eigen(pfa_handle *h, const std::vector<double> &values, StiffnessMatrix &out)
		{
			int32_t size;
			int64_t ndof, nnz;
			const int32_t *outer, *inner;
			pfa_sizes(h, &size, &ndof, &nnz);
			if (pfa_pattern(h, &nnz, &outer, &inner) != PFA_OK)
				log_and_throw_error("B200 assembly path: {}", pfa_last_error(h));
			out = Eigen::Map<const StiffnessMatrix>(ndof, ndof, nnz, outer, inner, values.data());
		}

		static void check(pfa_handle *h, const int rc)
		{
			if (rc == PFA_ERR_NOMEM)
				log_and_throw_error("bad alloc {}", pfa_last_error(h)); // Assembler.cpp:377-380
			if (rc != PFA_OK)
				log_and_throw_error("B200 assembly path: {}", pfa_last_error(h));
		}

	private:
		mutable pfa_handle *h_ = nullptr;
		mutable const void *key_ = nullptr;
		mutable int n_elements_ = 0;
	};

	/// Drop-in for NeoHookeanElasticity ("NeoHookean" in AssemblerUtils::make_assembler).
	class NeoHookeanElasticityB200 : public NeoHookeanElasticity
	{
	public:
		double assemble_energy(const bool is_volume, const std::vector<basis::ElementBases> &bases,
							   const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache,
							   const double t, const double dt, const Eigen::MatrixXd &displacement,
							   const Eigen::MatrixXd &displacement_prev) const override
		{
			if (use_robust_jacobian) // Bezier evaluator (NeoHookeanElasticity.cpp:357-359): CPU path
				return NeoHookeanElasticity::assemble_energy(is_volume, bases, gbases, cache, t, dt, displacement, displacement_prev);
			pfa_handle *h = handle(is_volume, int(displacement.size() / size()), bases, gbases, cache, t);
			double e = 0;
			DeviceAssembly::check(h, pfa_energy(h, displacement.data(), &e));
			return e;
		}

		Eigen::VectorXd assemble_energy_per_element(const bool is_volume, const std::vector<basis::ElementBases> &bases,
													const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache,
													const double t, const double dt, const Eigen::MatrixXd &displacement,
													const Eigen::MatrixXd &displacement_prev) const override
		{
			pfa_handle *h = handle(is_volume, int(displacement.size() / size()), bases, gbases, cache, t);
			Eigen::VectorXd out(bases.size());
			DeviceAssembly::check(h, pfa_energy_per_element(h, displacement.data(), out.data()));
			return out;
		}

		void assemble_gradient(const bool is_volume, const int n_basis, const std::vector<basis::ElementBases> &bases,
							   const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache,
							   const double t, const double dt, const Eigen::MatrixXd &displacement,
							   const Eigen::MatrixXd &displacement_prev, Eigen::MatrixXd &rhs) const override
		{
			pfa_handle *h = handle(is_volume, n_basis, bases, gbases, cache, t);
			rhs.resize(n_basis * size(), 1);
			DeviceAssembly::check(h, pfa_gradient(h, displacement.data(), rhs.data()));
		}

		void assemble_hessian(const bool is_volume, const int n_basis, const bool project_to_psd,
							  const std::vector<basis::ElementBases> &bases, const std::vector<basis::ElementBases> &gbases,
							  const AssemblyValsCache &cache, const double t, const double dt,
							  const Eigen::MatrixXd &displacement, const Eigen::MatrixXd &displacement_prev,
							  utils::MatrixCache &mat_cache, StiffnessMatrix &hess) const override
		{
			pfa_handle *h = handle(is_volume, n_basis, bases, gbases, cache, t);
			int64_t nnz;
			pfa_sizes(h, nullptr, nullptr, &nnz);
			values_.resize(nnz);
			DeviceAssembly::check(h, pfa_hessian(h, displacement.data(), project_to_psd ? 1 : 0, values_.data()));
			DeviceAssembly::to_eigen(h, values_, hess);
			// mat_cache is caller-owned scratch (ElasticForm.hpp:116); it is left untouched and valid.
		}

	private:
		pfa_handle *handle(const bool is_volume, const int n_basis, const std::vector<basis::ElementBases> &bases,
						   const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const double t) const
		{
			return dev_.get(PFA_NEOHOOKEAN, is_volume, n_basis, bases, gbases, cache,
							[&](const ElementAssemblyValues &vals, const int q, double &lambda, double &mu) {
								lame_params().lambda_mu(vals.quadrature.points.row(q), vals.val.row(q), t, vals.element_id, lambda, mu);
							});
		}
		DeviceAssembly dev_;
		mutable std::vector<double> values_;
	};

	/// Drop-in for Laplacian ("Laplacian"): LinearAssembler::assemble only.
	class LaplacianB200 : public Laplacian
	{
	public:
		void assemble(const bool is_volume, const int n_basis, const std::vector<basis::ElementBases> &bases,
					  const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const double t,
					  StiffnessMatrix &stiffness, const bool is_mass = false) const override
		{
			if (is_mass)
				return Laplacian::assemble(is_volume, n_basis, bases, gbases, cache, t, stiffness, is_mass);
			pfa_handle *h = dev_.get(PFA_LAPLACIAN, is_volume, n_basis, bases, gbases, cache,
									 [](const ElementAssemblyValues &, const int, double &lambda, double &mu) { lambda = mu = 0; });
			int64_t nnz;
			pfa_sizes(h, nullptr, nullptr, &nnz);
			values_.resize(nnz);
			DeviceAssembly::check(h, pfa_linear_stiffness(h, values_.data()));
			DeviceAssembly::to_eigen(h, values_, stiffness);
		}

	private:
		DeviceAssembly dev_;
		mutable std::vector<double> values_;
	};

	/// Drop-in for LinearElasticity ("LinearElasticity"): linear `assemble` plus the NL
	/// energy / gradient used when the linear material sits inside a nonlinear solve.
	class LinearElasticityB200 : public LinearElasticity
	{
	public:
		void assemble(const bool is_volume, const int n_basis, const std::vector<basis::ElementBases> &bases,
					  const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const double t,
					  StiffnessMatrix &stiffness, const bool is_mass = false) const override
		{
			pfa_handle *h = handle(is_volume, n_basis, bases, gbases, cache, t);
			int64_t nnz;
			pfa_sizes(h, nullptr, nullptr, &nnz);
			values_.resize(nnz);
			DeviceAssembly::check(h, pfa_linear_stiffness(h, values_.data()));
			DeviceAssembly::to_eigen(h, values_, stiffness);
		}

		double assemble_energy(const bool is_volume, const std::vector<basis::ElementBases> &bases,
							   const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache,
							   const double t, const double dt, const Eigen::MatrixXd &displacement,
							   const Eigen::MatrixXd &displacement_prev) const override
		{
			pfa_handle *h = handle(is_volume, int(displacement.size() / size()), bases, gbases, cache, t);
			double e = 0;
			DeviceAssembly::check(h, pfa_energy(h, displacement.data(), &e));
			return e;
		}

		void assemble_gradient(const bool is_volume, const int n_basis, const std::vector<basis::ElementBases> &bases,
							   const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache,
							   const double t, const double dt, const Eigen::MatrixXd &displacement,
							   const Eigen::MatrixXd &displacement_prev, Eigen::MatrixXd &rhs) const override
		{
			pfa_handle *h = handle(is_volume, n_basis, bases, gbases, cache, t);
			rhs.resize(n_basis * size(), 1);
			DeviceAssembly::check(h, pfa_gradient(h, displacement.data(), rhs.data()));
		}

	private:
		pfa_handle *handle(const bool is_volume, const int n_basis, const std::vector<basis::ElementBases> &bases,
						   const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const double t) const
		{
			return dev_.get(PFA_LINEAR_ELASTICITY, is_volume, n_basis, bases, gbases, cache,
							[&](const ElementAssemblyValues &vals, const int q, double &lambda, double &mu) {
								lame_params().lambda_mu(vals.quadrature.points.row(q), vals.val.row(q), t, vals.element_id, lambda, mu);
							});
		}
		DeviceAssembly dev_;
		mutable std::vector<double> values_;
	};
} // namespace polyfem::assembler::b200
