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
#include <polyfem/assembler/FixedCorotational.hpp>
#include <polyfem/assembler/Laplacian.hpp>
#include <polyfem/assembler/LinearElasticity.hpp>
#include <polyfem/assembler/Mass.hpp>
#include <polyfem/assembler/MooneyRivlinElasticity.hpp>
#include <polyfem/assembler/NeoHookeanElasticity.hpp>
#include <polyfem/assembler/SaintVenantElasticity.hpp>
#include <polyfem/assembler/ViscousDamping.hpp>
#include <polyfem/basis/ElementBases.hpp>
#include <polyfem/utils/Logger.hpp>
#include <polyfem/utils/MatrixCache.hpp>

#include <cmath>
#include <memory>
#include <vector>

namespace polyfem::assembler::b200
{
	/// values[] buffer of the host-pointer calls, page-locked when the driver grants it (pfa_host_alloc): the device-to-host copy
	/// of the matrix values reaches the PCIe rate only into pinned memory. Falls back to ordinary memory. Copies start empty.
	class HostValues
	{
	public:
		HostValues() = default;
		HostValues(const HostValues &) {}
		HostValues &operator=(const HostValues &)
		{
			release();
			return *this;
		}
		~HostValues() { release(); }

		double *resize(const size_t n)
		{
			if (n > cap_)
			{
				release();
				p_ = static_cast<double *>(pfa_host_alloc(n * sizeof(double)));
				pinned_ = p_ != nullptr;
				if (!pinned_)
				{
					fallback_.resize(n);
					p_ = fallback_.data();
				}
				cap_ = n;
			}
			return p_;
		}
		const double *data() const { return p_; }

	private:
		void release()
		{
			if (pinned_)
				pfa_host_free(p_);
			std::vector<double>().swap(fallback_);
			p_ = nullptr;
			cap_ = 0;
			pinned_ = false;
		}
		double *p_ = nullptr;
		size_t cap_ = 0;
		bool pinned_ = false;
		std::vector<double> fallback_;
	};

	/// Owns one pfa_handle; rebuilt when the FE space (bases pointer / size) changes.
	/// The reference's methods are const and called from one host thread, so the
	/// device state is `mutable` in the assemblers below.
	class DeviceAssembly
	{
	public:
		DeviceAssembly() = default;
		// a copied assembler starts without device state (its handle is built on its first assemble_* call)
		DeviceAssembly(const DeviceAssembly &) {}
		DeviceAssembly &operator=(const DeviceAssembly &)
		{
			reset();
			return *this;
		}
		~DeviceAssembly() { reset(); }

		void reset()
		{
			if (h_)
				pfa_destroy(h_);
			h_ = nullptr;
			key_ = nullptr;
		}

		/// Call from add_multimaterial / set_materials / set_size overrides: the next assemble_* re-evaluates lambda, mu
		/// (or the density) and uploads them with pfa_set_materials.
		void invalidate_materials() const { ++material_version_; }

		/// Reads what the hot path needs out of `bases` / `gbases` once per mesh.
		/// `lame(vals, q, p1, p2, p3)` evaluates the material parameters of the element behind `vals` at quadrature point q:
		/// (lambda, mu) from LameParameters::lambda_mu, the density for PFA_MASS, (c1, c2, k) for PFA_MOONEY_RIVLIN.
		template <typename LameFn>
		pfa_handle *get(const pfa_material material, const bool is_volume, const int n_basis,
						const std::vector<basis::ElementBases> &bases,
						const std::vector<basis::ElementBases> &gbases,
						const AssemblyValsCache &cache, const double t, const LameFn &lame) const
		{
			// The handle is keyed on the FE space (bases pointer, element count and the first element's basis count / first
			// global index, which change when a vector is re-allocated at the same address for another space). The reference
			// evaluates lambda_mu(..., t, ...) on every assemble_* call (MatParams.cpp:368-401): when t or the material
			// version changed since the last upload, the parameters are re-evaluated and sent with pfa_set_materials.
			const int first_global = bases.empty() || bases[0].bases.empty() ? -1 : bases[0].bases[0].global()[0].index;
			const int first_count = bases.empty() ? 0 : int(bases[0].bases.size());
			if (h_ && key_ == bases.data() && n_elements_ == int(bases.size()) && n_basis_ == n_basis && first_global_ == first_global && first_count_ == first_count)
			{
				// (a refresh that changes the layout - parameters that were constant per element now vary inside an
				// element, or the other way round - rebuilds the handle below)
				if ((t == t_ && material_version_ == uploaded_version_) || refresh_materials(material, is_volume, bases, gbases, cache, lame, t))
					return h_;
			}
			const_cast<DeviceAssembly *>(this)->reset();
			if (!is_volume)
				log_and_throw_error("B200 assembly path: only volumetric (tetrahedral) meshes are supported");

			ElementAssemblyValues vals;
			cache.compute(0, is_volume, bases[0], gbases[0], vals);
			const int n_el = int(bases.size());
			const int n_loc = int(vals.basis_values.size());
			const int n_qp = int(vals.quadrature.weights.size());

			// straight meshes keep P1 geometry (VarForm.cpp:88-93, 358-366): J^-T and det are constant
			// per element, the library computes them from the 4 vertices, stores 10 doubles per element
			// instead of 10 per quadrature point and may re-order the elements along a space-filling curve
			bool affine = true;
			for (int e = 0; e < n_el && affine; ++e)
				affine = gbases[e].bases.size() == 4;
			std::vector<double> vertices(affine ? size_t(n_el) * 12 : 0);

			std::vector<int32_t> conn(size_t(n_el) * n_loc);
			std::vector<double> jac_it(affine ? 0 : size_t(n_el) * n_qp * 9), da(affine ? 0 : size_t(n_el) * n_qp);
			std::vector<double> ref_vals(material == PFA_MASS ? size_t(n_qp) * n_loc : 0);
			for (size_t k = 0; k < ref_vals.size(); ++k)
				ref_vals[k] = vals.basis_values[k % n_loc].val(k / n_loc);
			std::vector<double> lambda(size_t(n_el) * n_qp), mu(size_t(n_el) * n_qp), third(material == PFA_MOONEY_RIVLIN ? size_t(n_el) * n_qp : 0);
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
				if (affine)
					for (int k = 0; k < 4; ++k)
						for (int c = 0; c < 3; ++c)
							vertices[(size_t(e) * 4 + k) * 3 + c] = gbases[e].bases[k].global()[0].node(c);
				for (int q = 0; q < n_qp; ++q)
				{
					if (!affine)
					{
						const Eigen::Matrix3d jit = vals.jac_it[q];
						for (int r = 0; r < 3; ++r)
							for (int c = 0; c < 3; ++c)
								jac_it[(size_t(e) * n_qp + q) * 9 + r * 3 + c] = jit(r, c);
						da[size_t(e) * n_qp + q] = vals.det(q) * vals.quadrature.weights(q);
					}
					// (lambda, mu), or for PFA_MASS the density in `lambda`
					double p3 = 0;
					lame(vals, q, lambda[size_t(e) * n_qp + q], mu[size_t(e) * n_qp + q], p3);
					if (!third.empty())
						third[size_t(e) * n_qp + q] = p3;
				}
			}

			// One (lambda, mu) per element when the parameters do not vary inside any element (the common case: constant or
			// per-body materials): the library then takes its per-element paths (LinearElasticity / Laplacian: the
			// reference-moment kernel instead of the quadrature loop)
			const int mat_stride = compress_if_uniform(n_el, n_qp, lambda, mu, third);

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
			if (affine)
				d.vertices = vertices.data();
			else
			{
				d.jac_it = jac_it.data(); // isoparametric geometry: J^-T and det*w per quadrature point
				d.da = da.data();
			}
			d.lambda = lambda.data();
			d.mu = mu.data();
			if (material == PFA_MASS)
			{
				d.ref_vals = ref_vals.data();
				d.density = lambda.data();
			}
			if (!third.empty())
				d.param3 = third.data();
			d.material_stride = mat_stride;
			d.device = 0;
#ifdef POLYSOLVE_LARGE_INDEX
			d.flags |= PFA_FLAG_LARGE_INDEX;
#endif
			if (pfa_create(&d, &h_) != PFA_OK)
				log_and_throw_error("B200 assembly path: {}", pfa_last_error(nullptr));
			key_ = bases.data();
			n_elements_ = n_el;
			n_basis_ = n_basis;
			first_global_ = first_global;
			first_count_ = first_count;
			n_qp_ = n_qp;
			mat_stride_ = mat_stride;
			t_ = t;
			uploaded_version_ = material_version_;
			return h_;
		}

		/// [n_el][n_qp] parameter arrays -> [n_el] when every element carries one value per array; returns the stride (1 or n_qp)
		static int compress_if_uniform(const int n_el, const int n_qp, std::vector<double> &p1, std::vector<double> &p2, std::vector<double> &p3)
		{
			if (n_qp == 1)
				return 1;
			for (std::vector<double> *a : {&p1, &p2, &p3})
				for (size_t e = 0; e < a->size() / size_t(n_qp); ++e)
					for (int q = 1; q < n_qp; ++q)
						if ((*a)[e * n_qp + q] != (*a)[e * n_qp])
							return n_qp;
			for (std::vector<double> *a : {&p1, &p2, &p3})
				if (!a->empty())
				{
					for (int e = 1; e < n_el; ++e)
						(*a)[size_t(e)] = (*a)[size_t(e) * n_qp];
					a->resize(size_t(n_el));
				}
			return 1;
		}

		/// lambda / mu (or the density) of every (element, quadrature point) again, for a new t or after a material change
		/// Returns false when the new parameters need another layout than the handle was created with (then nothing is uploaded
		/// and the caller rebuilds the handle).
		template <typename LameFn>
		bool refresh_materials(const pfa_material material, const bool is_volume, const std::vector<basis::ElementBases> &bases,
							   const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const LameFn &lame, const double t) const
		{
			if (material != PFA_LAPLACIAN)
			{
				const int n_el = int(bases.size());
				std::vector<double> lambda(size_t(n_el) * n_qp_), mu(size_t(n_el) * n_qp_), third(material == PFA_MOONEY_RIVLIN ? size_t(n_el) * n_qp_ : 0);
				ElementAssemblyValues vals;
				for (int e = 0; e < n_el; ++e)
				{
					cache.compute(e, is_volume, bases[e], gbases[e], vals);
					for (int q = 0; q < n_qp_; ++q)
					{
						double p3 = 0;
						lame(vals, q, lambda[size_t(e) * n_qp_ + q], mu[size_t(e) * n_qp_ + q], p3);
						if (!third.empty())
							third[size_t(e) * n_qp_ + q] = p3;
					}
				}
				if (compress_if_uniform(n_el, n_qp_, lambda, mu, third) != mat_stride_)
					return false;
				if (!third.empty())
					check(h_, pfa_set_material_params(h_, lambda.data(), mu.data(), third.data(), mat_stride_));
				else
					check(h_, pfa_set_materials(h_, lambda.data(), material == PFA_MASS ? nullptr : mu.data(), mat_stride_));
			}
			t_ = t;
			uploaded_version_ = material_version_;
			return true;
		}

		/// Wraps values[] in the reference's matrix type (pattern identical to SparseMatrixCache's). With POLYSOLVE_LARGE_INDEX
		/// (utils/Types.hpp:21-25) StiffnessMatrix stores std::ptrdiff_t indices: the handle is then created with
		/// PFA_FLAG_LARGE_INDEX and hands out the int64 pattern.
		static void to_eigen(pfa_handle *h, const double *values, StiffnessMatrix &out)
		{
			int32_t size;
			int64_t ndof, nnz;
			pfa_sizes(h, &size, &ndof, &nnz);
#ifdef POLYSOLVE_LARGE_INDEX
			const int64_t *outer, *inner;
			if (pfa_pattern_wide(h, &nnz, &outer, &inner) != PFA_OK)
				log_and_throw_error("B200 assembly path: {}", pfa_last_error(h));
			out = Eigen::Map<const StiffnessMatrix>(ndof, ndof, nnz, reinterpret_cast<const std::ptrdiff_t *>(outer), reinterpret_cast<const std::ptrdiff_t *>(inner), values);
#else
			const int32_t *outer, *inner;
			if (pfa_pattern(h, &nnz, &outer, &inner) != PFA_OK)
				log_and_throw_error("B200 assembly path: {}", pfa_last_error(h));
			out = Eigen::Map<const StiffnessMatrix>(ndof, ndof, nnz, outer, inner, values);
#endif
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
		mutable int n_elements_ = 0, n_basis_ = 0, first_global_ = -1, first_count_ = 0, n_qp_ = 0, mat_stride_ = 1;
		mutable double t_ = 0;
		mutable unsigned material_version_ = 0, uploaded_version_ = 0;
	};

	/// The four NLAssembler virtuals (Assembler.cpp:495-771) through the C ABI, for the reference assembler class `Base` whose
	/// local math the library implements as MATERIAL. A drop-in derives from this, evaluates its material parameters in
	/// material_params() and may veto the device path per object state in on_device() (then every virtual forwards to Base:
	/// energy, gradient and Hessian must come from the same implementation).
	template <class Base, pfa_material MATERIAL>
	class NLAssemblerB200 : public Base
	{
	public:
		double assemble_energy(const bool is_volume, const std::vector<basis::ElementBases> &bases,
							   const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache,
							   const double t, const double dt, const Eigen::MatrixXd &displacement,
							   const Eigen::MatrixXd &displacement_prev) const override
		{
			if (!on_device())
				return Base::assemble_energy(is_volume, bases, gbases, cache, t, dt, displacement, displacement_prev);
			pfa_handle *h = handle(is_volume, int(displacement.size() / this->size()), bases, gbases, cache, t);
			before_call(h, dt, displacement, displacement_prev);
			double e = 0;
			DeviceAssembly::check(h, pfa_energy(h, displacement.data(), &e));
			return e;
		}

		Eigen::VectorXd assemble_energy_per_element(const bool is_volume, const std::vector<basis::ElementBases> &bases,
													const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache,
													const double t, const double dt, const Eigen::MatrixXd &displacement,
													const Eigen::MatrixXd &displacement_prev) const override
		{
			if (!on_device())
				return Base::assemble_energy_per_element(is_volume, bases, gbases, cache, t, dt, displacement, displacement_prev);
			pfa_handle *h = handle(is_volume, int(displacement.size() / this->size()), bases, gbases, cache, t);
			before_call(h, dt, displacement, displacement_prev);
			Eigen::VectorXd out(bases.size());
			DeviceAssembly::check(h, pfa_energy_per_element(h, displacement.data(), out.data()));
			return out;
		}

		void assemble_gradient(const bool is_volume, const int n_basis, const std::vector<basis::ElementBases> &bases,
							   const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache,
							   const double t, const double dt, const Eigen::MatrixXd &displacement,
							   const Eigen::MatrixXd &displacement_prev, Eigen::MatrixXd &rhs) const override
		{
			if (!on_device())
				return Base::assemble_gradient(is_volume, n_basis, bases, gbases, cache, t, dt, displacement, displacement_prev, rhs);
			pfa_handle *h = handle(is_volume, n_basis, bases, gbases, cache, t);
			before_call(h, dt, displacement, displacement_prev);
			rhs.resize(n_basis * this->size(), 1);
			DeviceAssembly::check(h, pfa_gradient(h, displacement.data(), rhs.data()));
		}

		void assemble_hessian(const bool is_volume, const int n_basis, const bool project_to_psd,
							  const std::vector<basis::ElementBases> &bases, const std::vector<basis::ElementBases> &gbases,
							  const AssemblyValsCache &cache, const double t, const double dt,
							  const Eigen::MatrixXd &displacement, const Eigen::MatrixXd &displacement_prev,
							  utils::MatrixCache &mat_cache, StiffnessMatrix &hess) const override
		{
			if (!on_device())
				return Base::assemble_hessian(is_volume, n_basis, project_to_psd, bases, gbases, cache, t, dt, displacement, displacement_prev, mat_cache, hess);
			pfa_handle *h = handle(is_volume, n_basis, bases, gbases, cache, t);
			before_call(h, dt, displacement, displacement_prev);
			int64_t nnz;
			pfa_sizes(h, nullptr, nullptr, &nnz);
			double *values = values_.resize(size_t(nnz));
			DeviceAssembly::check(h, pfa_hessian(h, displacement.data(), project_to_psd ? 1 : 0, values));
			DeviceAssembly::to_eigen(h, values, hess);
			// mat_cache is caller-owned scratch (ElasticForm.hpp:116); it is left untouched and valid.
		}

		// material changes after the first assembly must reach the device (the handle caches the parameters)
		void add_multimaterial(const int index, const json &params, const Units &units, const std::string &root_path) override
		{
			Base::add_multimaterial(index, params, units, root_path);
			dev_.invalidate_materials();
		}
		void set_size(const int size) override
		{
			Base::set_size(size);
			dev_.invalidate_materials();
		}

	protected:
		virtual bool on_device() const { return true; }
		/// assemblers that read dt / displacement_prev hand them to the handle here (ViscousDamping)
		virtual void before_call(pfa_handle *, const double, const Eigen::MatrixXd &, const Eigen::MatrixXd &) const {}
		/// parameters of the element behind `vals` at quadrature point q, evaluated like the reference does inside its local
		/// assembly (p3 only for three-parameter materials)
		virtual void material_params(const ElementAssemblyValues &vals, const int q, const double t, double &p1, double &p2, double &p3) const = 0;

		pfa_handle *handle(const bool is_volume, const int n_basis, const std::vector<basis::ElementBases> &bases,
						   const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const double t) const
		{
			return dev_.get(MATERIAL, is_volume, n_basis, bases, gbases, cache, t,
							[&](const ElementAssemblyValues &vals, const int q, double &p1, double &p2, double &p3) { material_params(vals, q, t, p1, p2, p3); });
		}
		DeviceAssembly dev_;
		mutable HostValues values_;
	};

	/// Drop-in for NeoHookeanElasticity ("NeoHookean" in AssemblerUtils::make_assembler).
	class NeoHookeanElasticityB200 : public NLAssemblerB200<NeoHookeanElasticity, PFA_NEOHOOKEAN>
	{
	protected:
		// Bezier evaluator of the Jacobian (NeoHookeanElasticity.cpp:357-359, 475-498, 567-587): energy, gradient and Hessian
		// all use it, so the whole assembler stays on the CPU path when it is enabled
		bool on_device() const override { return !use_robust_jacobian; }
		void material_params(const ElementAssemblyValues &vals, const int q, const double t, double &lambda, double &mu, double &) const override
		{
			lame_params().lambda_mu(vals.quadrature.points.row(q), vals.val.row(q), t, vals.element_id, lambda, mu);
		}
	};

	/// Drop-in for FixedCorotational ("FixedCorotational"): Lame parameters like NeoHookean; inverted elements stay on the
	/// device path (allow_inversion() is true, the signed SVD handles det F < 0).
	class FixedCorotationalB200 : public NLAssemblerB200<FixedCorotational, PFA_FIXED_COROTATIONAL>
	{
	protected:
		bool on_device() const override { return size() == 3; }
		void material_params(const ElementAssemblyValues &vals, const int q, const double t, double &lambda, double &mu, double &) const override
		{
			lame_params().lambda_mu(vals.quadrature.points.row(q), vals.val.row(q), t, vals.element_id, lambda, mu);
		}
	};

	/// Drop-in for SaintVenantElasticity ("SaintVenant") when its elasticity tensor is the isotropic one of (lambda, mu)
	/// (ElasticityTensor::set_from_lambda_mu / set_from_young_poisson, MatParams.cpp:211-276: C_ii = lambda + 2 mu, C_ij = lambda
	/// for i != j < 3, C_kk = mu for k >= 3, zero otherwise). Any other tensor (set_from_entries, orthotropic, rotated by a
	/// fiber direction) keeps the CPU path.
	class SaintVenantElasticityB200 : public NLAssemblerB200<SaintVenantElasticity, PFA_SAINT_VENANT>
	{
	protected:
		bool on_device() const override
		{
			if (size() != 3)
				return false;
			const double lambda = stifness_tensor(0, 1), mu = stifness_tensor(3, 3);
			const double tol = 1e-14 * (std::abs(lambda) + std::abs(mu));
			for (int i = 0; i < 6; ++i)
				for (int j = 0; j < 6; ++j)
				{
					const double want = i == j ? (i < 3 ? lambda + 2 * mu : mu) : (i < 3 && j < 3 ? lambda : 0.0);
					if (std::abs(stifness_tensor(i, j) - want) > tol)
						return false;
				}
			return true;
		}
		void material_params(const ElementAssemblyValues &, const int, const double, double &lambda, double &mu, double &) const override
		{
			lambda = stifness_tensor(0, 1);
			mu = stifness_tensor(3, 3);
		}
	};

	/// Drop-in for MooneyRivlinElasticity ("MooneyRivlin", a GenericElastic): c1, c2, k are evaluated per element as in
	/// MooneyRivlinElasticity::elastic_energy (MooneyRivlinElasticity.hpp:32-34: GenericMatParam(p, t, el_id) with
	/// p = vals.val.row(q)).
	class MooneyRivlinElasticityB200 : public NLAssemblerB200<MooneyRivlinElasticity, PFA_MOONEY_RIVLIN>
	{
	protected:
		bool on_device() const override { return size() == 3; }
		void material_params(const ElementAssemblyValues &vals, const int q, const double t, double &p1, double &p2, double &p3) const override
		{
			p1 = c1()(vals.val.row(q), t, vals.element_id);
			p2 = c2()(vals.val.row(q), t, vals.element_id);
			p3 = k()(vals.val.row(q), t, vals.element_id);
		}
	};

	/// Drop-in for ViscousDamping ("ViscousDamping"; the damping form of a transient elastic solve): psi, phi are global
	/// (DampingParameters); displacement_prev and dt of the virtuals go through pfa_set_previous, and a displacement_prev of
	/// another size - the first step - gives zeros exactly like ViscousDamping.cpp:125-126, 176-179, 299-300.
	class ViscousDampingB200 : public NLAssemblerB200<ViscousDamping, PFA_VISCOUS_DAMPING>
	{
	protected:
		bool on_device() const override { return size() == 3; }
		void material_params(const ElementAssemblyValues &, const int, const double, double &psi, double &phi, double &) const override
		{
			psi = get_psi();
			phi = get_phi();
		}
		void before_call(pfa_handle *h, const double dt, const Eigen::MatrixXd &displacement, const Eigen::MatrixXd &displacement_prev) const override
		{
			DeviceAssembly::check(h, pfa_set_previous(h, displacement_prev.size() == displacement.size() ? displacement_prev.data() : nullptr, dt));
		}
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
			pfa_handle *h = dev_.get(PFA_LAPLACIAN, is_volume, n_basis, bases, gbases, cache, t,
									 [](const ElementAssemblyValues &, const int, double &lambda, double &mu, double &) { lambda = mu = 0; });
			int64_t nnz;
			pfa_sizes(h, nullptr, nullptr, &nnz);
			double *values = values_.resize(size_t(nnz));
			DeviceAssembly::check(h, pfa_linear_stiffness(h, values));
			DeviceAssembly::to_eigen(h, values, stiffness);
		}

	private:
		DeviceAssembly dev_;
		mutable HostValues values_;
	};

	/// Drop-in for Mass ("Mass"): the matrix InertiaForm is built with (State::build_mass_matrix /
	/// SolveData::init_forms). The caller's AssemblyValsCache must be the mass one (is_mass() quadrature).
	class MassB200 : public Mass
	{
	public:
		void assemble(const bool is_volume, const int n_basis, const std::vector<basis::ElementBases> &bases,
					  const std::vector<basis::ElementBases> &gbases, const AssemblyValsCache &cache, const double t,
					  StiffnessMatrix &stiffness, const bool is_mass = false) const override
		{
			if (size() != 3)
				return Mass::assemble(is_volume, n_basis, bases, gbases, cache, t, stiffness, is_mass);
			pfa_handle *h = dev_.get(PFA_MASS, is_volume, n_basis, bases, gbases, cache, t,
									 [&](const ElementAssemblyValues &vals, const int q, double &rho, double &unused, double &) {
										 rho = density()(vals.quadrature.points.row(q), vals.val.row(q), t, vals.element_id);
										 unused = 0;
									 });
			int64_t nnz;
			pfa_sizes(h, nullptr, nullptr, &nnz);
			double *values = values_.resize(size_t(nnz));
			DeviceAssembly::check(h, pfa_linear_stiffness(h, values));
			DeviceAssembly::to_eigen(h, values, stiffness);
		}

	private:
		DeviceAssembly dev_;
		mutable HostValues values_;
	};

	/// Optional: the Newton system of NLProblem straight from the device (fourth edit).
	/// Replaces FullNLProblem::hessian + BCLagrangianForm::project_hessian / project_gradient
	/// (solver/NLProblem.cpp:596-640, 735-751) for a problem whose only assembled form is the elastic one.
	struct ReducedNewtonSystem
	{
		/// boundary_nodes = BCLagrangianForm::boundary_nodes_ (constrained dofs, any order)
		static void set_constraints(pfa_handle *h, const std::vector<int> &boundary_nodes)
		{
			std::vector<int32_t> dofs(boundary_nodes.begin(), boundary_nodes.end());
			DeviceAssembly::check(h, pfa_set_constrained_dofs(h, dofs.data(), int64_t(dofs.size())));
		}

		/// weight = Form::weight() / scale_ (solver/forms/Form.hpp:30-56); x is the FULL vector
		static double assemble(pfa_handle *h, const Eigen::VectorXd &x_full, const bool project_to_psd, const double weight,
							   Eigen::VectorXd &grad_reduced, StiffnessMatrix &hessian_reduced, std::vector<double> &values)
		{
			int64_t ndof_r, nnz_r;
			DeviceAssembly::check(h, pfa_reduced_sizes(h, &ndof_r, &nnz_r));
			grad_reduced.resize(ndof_r);
			values.resize(nnz_r);
			double energy = 0;
			DeviceAssembly::check(h, pfa_grad_hess_reduced(h, x_full.data(), project_to_psd ? 1 : 0, weight, &energy, grad_reduced.data(), values.data()));
			const int32_t *outer, *inner;
			DeviceAssembly::check(h, pfa_reduced_pattern(h, &outer, &inner));
			hessian_reduced = Eigen::Map<const StiffnessMatrix>(ndof_r, ndof_r, nnz_r, outer, inner, values.data());
			return energy;
		}

		/// ElasticForm::is_step_valid (solver/forms/ElasticForm.cpp:388-396) + the energy of the same pass
		static bool is_step_valid(pfa_handle *h, const Eigen::VectorXd &x1, double &energy)
		{
			int32_t valid = 0;
			DeviceAssembly::check(h, pfa_is_step_valid(h, x1.data(), &valid, &energy));
			return valid != 0;
		}
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
			double *values = values_.resize(size_t(nnz));
			DeviceAssembly::check(h, pfa_linear_stiffness(h, values));
			DeviceAssembly::to_eigen(h, values, stiffness);
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
			return dev_.get(PFA_LINEAR_ELASTICITY, is_volume, n_basis, bases, gbases, cache, t,
							[&](const ElementAssemblyValues &vals, const int q, double &lambda, double &mu, double &) {
								lame_params().lambda_mu(vals.quadrature.points.row(q), vals.val.row(q), t, vals.element_id, lambda, mu);
							});
		}
		DeviceAssembly dev_;
		mutable HostValues values_;
	};
} // namespace polyfem::assembler::b200
