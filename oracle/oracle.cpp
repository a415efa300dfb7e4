// CPU ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.h for the parity status).
//
// Restates, without Eigen/TBB, the algorithm of PolyFEM's assembly hot path. Reference
// locations are cited per function as file:line relative to /root/reference/src/polyfem/.
// The restatement keeps the reference's data flow on purpose (per-element assembly values,
// dense B^T H_F B, per-thread value buffers + serial merge, slot map by (element, call#)):
// it is both the numerical checker and the "reference-algorithm CPU baseline" bench.py times.
#include "oracle.h"

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <utility>
#include <vector>

namespace
{
	using std::size_t;

	double now_seconds()
	{
		return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
	}

	// ---------------------------------------------------------------------------------------
	// utils/MaybeParallelFor.tpp:18-68 — parallel_for over a blocked range with a thread id.
	// TBB's blocked_range is replaced by one contiguous block per thread.
	// ---------------------------------------------------------------------------------------
	void maybe_parallel_for(int n, int n_threads, const std::function<void(int, int, int)> &body)
	{
		n_threads = std::max(1, std::min(n_threads, n));
		if (n_threads == 1)
		{
			body(0, n, 0);
			return;
		}
		std::vector<std::thread> pool;
		pool.reserve(n_threads);
		for (int t = 0; t < n_threads; ++t)
		{
			const int start = int((long long)n * t / n_threads);
			const int end = int((long long)n * (t + 1) / n_threads);
			pool.emplace_back([=, &body]() { body(start, end, t); });
		}
		for (auto &th : pool)
			th.join();
	}

	// ---------------------------------------------------------------------------------------
	// Forward-mode autodiff scalars standing in for utils/autodiff.h DScalar1 / DScalar2
	// (value + gradient [+ Hessian] w.r.t. the local dofs), dynamic size.
	// ---------------------------------------------------------------------------------------
	thread_local int g_nvars = 0;

	struct D1
	{
		double v = 0;
		std::vector<double> g;
		D1() : g(g_nvars, 0.0) {}
		D1(double c) : v(c), g(g_nvars, 0.0) {}
		D1(int idx, double c) : v(c), g(g_nvars, 0.0) { g[idx] = 1.0; }
	};
	D1 operator+(const D1 &a, const D1 &b)
	{
		D1 r(a.v + b.v);
		for (int i = 0; i < g_nvars; ++i)
			r.g[i] = a.g[i] + b.g[i];
		return r;
	}
	D1 operator-(const D1 &a, const D1 &b)
	{
		D1 r(a.v - b.v);
		for (int i = 0; i < g_nvars; ++i)
			r.g[i] = a.g[i] - b.g[i];
		return r;
	}
	D1 operator*(const D1 &a, const D1 &b)
	{
		D1 r(a.v * b.v);
		for (int i = 0; i < g_nvars; ++i)
			r.g[i] = a.v * b.g[i] + b.v * a.g[i];
		return r;
	}
	D1 operator/(const D1 &a, const D1 &b)
	{
		D1 r(a.v / b.v);
		for (int i = 0; i < g_nvars; ++i)
			r.g[i] = (a.g[i] - r.v * b.g[i]) / b.v;
		return r;
	}
	D1 log(const D1 &a)
	{
		D1 r(std::log(a.v));
		for (int i = 0; i < g_nvars; ++i)
			r.g[i] = a.g[i] / a.v;
		return r;
	}

	D1 pow(const D1 &a, double p)
	{
		D1 r(std::pow(a.v, p));
		const double d1 = p * std::pow(a.v, p - 1.0);
		for (int i = 0; i < g_nvars; ++i)
			r.g[i] = d1 * a.g[i];
		return r;
	}

	struct D2
	{
		double v = 0;
		std::vector<double> g, h;
		D2() : g(g_nvars, 0.0), h(size_t(g_nvars) * g_nvars, 0.0) {}
		D2(double c) : v(c), g(g_nvars, 0.0), h(size_t(g_nvars) * g_nvars, 0.0) {}
		D2(int idx, double c) : v(c), g(g_nvars, 0.0), h(size_t(g_nvars) * g_nvars, 0.0) { g[idx] = 1.0; }
	};
	D2 operator+(const D2 &a, const D2 &b)
	{
		D2 r(a.v + b.v);
		for (int i = 0; i < g_nvars; ++i)
			r.g[i] = a.g[i] + b.g[i];
		for (size_t i = 0; i < r.h.size(); ++i)
			r.h[i] = a.h[i] + b.h[i];
		return r;
	}
	D2 operator-(const D2 &a, const D2 &b)
	{
		D2 r(a.v - b.v);
		for (int i = 0; i < g_nvars; ++i)
			r.g[i] = a.g[i] - b.g[i];
		for (size_t i = 0; i < r.h.size(); ++i)
			r.h[i] = a.h[i] - b.h[i];
		return r;
	}
	D2 operator*(const D2 &a, const D2 &b)
	{
		const int n = g_nvars;
		D2 r(a.v * b.v);
		for (int i = 0; i < n; ++i)
			r.g[i] = a.v * b.g[i] + b.v * a.g[i];
		for (int i = 0; i < n; ++i)
			for (int j = 0; j < n; ++j)
				r.h[size_t(i) * n + j] = a.v * b.h[size_t(i) * n + j] + b.v * a.h[size_t(i) * n + j] + a.g[i] * b.g[j] + b.g[i] * a.g[j];
		return r;
	}
	D2 operator/(const D2 &a, const D2 &b)
	{
		// a * (1/b)
		const int n = g_nvars;
		D2 inv(1.0 / b.v);
		for (int i = 0; i < n; ++i)
			inv.g[i] = -b.g[i] / (b.v * b.v);
		for (int i = 0; i < n; ++i)
			for (int j = 0; j < n; ++j)
				inv.h[size_t(i) * n + j] = 2.0 * b.g[i] * b.g[j] / (b.v * b.v * b.v) - b.h[size_t(i) * n + j] / (b.v * b.v);
		return a * inv;
	}
	D2 log(const D2 &a)
	{
		const int n = g_nvars;
		D2 r(std::log(a.v));
		for (int i = 0; i < n; ++i)
			r.g[i] = a.g[i] / a.v;
		for (int i = 0; i < n; ++i)
			for (int j = 0; j < n; ++j)
				r.h[size_t(i) * n + j] = a.h[size_t(i) * n + j] / a.v - a.g[i] * a.g[j] / (a.v * a.v);
		return r;
	}

	D2 pow(const D2 &a, double p)
	{
		const int n = g_nvars;
		D2 r(std::pow(a.v, p));
		const double d1 = p * std::pow(a.v, p - 1.0), d2 = p * (p - 1.0) * std::pow(a.v, p - 2.0);
		for (int i = 0; i < n; ++i)
			r.g[i] = d1 * a.g[i];
		for (int i = 0; i < n; ++i)
			for (int j = 0; j < n; ++j)
				r.h[size_t(i) * n + j] = d1 * a.h[size_t(i) * n + j] + d2 * a.g[i] * a.g[j];
		return r;
	}

	template <typename T>
	struct Alloc
	{
		static T var(int i, double v) { return T(i, v); }
	};
	template <>
	struct Alloc<double>
	{
		static double var(int, double v) { return v; }
	};

	// ---------------------------------------------------------------------------------------
	// 3x3 helpers following Eigen's fixed-size formulas (cofactor determinant / inverse).
	// Matrices are row-major double[9]: m[r*3+c].
	// ---------------------------------------------------------------------------------------
	template <typename T>
	T det3(const T *m)
	{
		auto h = [&](int a, int b, int c) { return m[0 * 3 + a] * (m[1 * 3 + b] * m[2 * 3 + c] - m[1 * 3 + c] * m[2 * 3 + b]); };
		return h(0, 1, 2) - h(1, 0, 2) + h(2, 0, 1);
	}

	void inverse3(const double *m, double *inv)
	{
		auto cof = [&](int i, int j) {
			const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
			return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
		};
		const double invdet = 1.0 / det3(m);
		for (int i = 0; i < 3; ++i)
			for (int j = 0; j < 3; ++j)
				inv[i * 3 + j] = cof(j, i) * invdet;
	}

	// ---------------------------------------------------------------------------------------
	// Lagrange P_p basis on the reference tet (used when the assembly-values cache is empty and
	// the reference re-evaluates the bases for every element, ElementAssemblyValues.cpp:181-183).
	// ---------------------------------------------------------------------------------------
	void lagrange_grads(int p, int n_loc, const int32_t *lattice, const double *pt, double *grads /*[n_loc][3]*/)
	{
		const double lam[4] = {1.0 - pt[0] - pt[1] - pt[2], pt[0], pt[1], pt[2]};
		static const double dlam[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
		for (int n = 0; n < n_loc; ++n)
		{
			const int m[4] = {p - lattice[n * 3] - lattice[n * 3 + 1] - lattice[n * 3 + 2], lattice[n * 3], lattice[n * 3 + 1], lattice[n * 3 + 2]};
			double f[4], df[4], c = 1.0;
			for (int v = 0; v < 4; ++v)
			{
				f[v] = 1.0;
				df[v] = 0.0;
				for (int k = 0; k < m[v]; ++k)
				{
					const double term = p * lam[v] - k;
					df[v] = df[v] * term + f[v] * p;
					f[v] *= term;
					c /= double(k + 1);
				}
			}
			double g[3] = {0, 0, 0};
			for (int v = 0; v < 4; ++v)
			{
				double others = 1.0;
				for (int u = 0; u < 4; ++u)
					if (u != v)
						others *= f[u];
				for (int d = 0; d < 3; ++d)
					g[d] += df[v] * others * dlam[v][d];
			}
			for (int d = 0; d < 3; ++d)
				grads[n * 3 + d] = c * g[d];
		}
	}

	// ---------------------------------------------------------------------------------------
	// assembler/ElementAssemblyValues.{hpp,cpp}, AssemblyValues.hpp:12-35
	// ---------------------------------------------------------------------------------------
	struct ElementAssemblyValues
	{
		int element_id = -1;
		int n_loc = 0, n_qp = 0;
		std::vector<int> global;      // [n_loc]  basis_values[j].global[0].index  (weight 1)
		std::vector<double> grad;     // [n_loc][n_qp][3]  basis_values[j].grad
		std::vector<double> grad_t_m; // [n_loc][n_qp][3]  basis_values[j].grad_t_m
		std::vector<double> val;      // [n_loc][n_qp]     basis_values[j].val (only when the table is given)
		std::vector<double> jac_it;   // [n_qp][9] row-major
		std::vector<double> det;      // [n_qp]
		std::vector<double> weights;  // [n_qp] quadrature.weights

		const double *g(int j, int q) const { return &grad[(size_t(j) * n_qp + q) * 3]; }
		const double *gt(int j, int q) const { return &grad_t_m[(size_t(j) * n_qp + q) * 3]; }
	};

	struct Problem;

	// ElementAssemblyValues::compute + finalize3d (ElementAssemblyValues.cpp:164-237, 65-104)
	void compute_assembly_values(const Problem &pb, int e, ElementAssemblyValues &vals);

	struct Problem
	{
		oracle_desc d;
		int size = 3; // Assembler::size()
		std::vector<int32_t> conn, lattice, geom_lattice;
		std::vector<double> vertices, qpts, qw, ref_grads, lambda, mu, param3, ref_vals, density, geom_nodes;
		// AssemblyValsCache (AssemblyValsCache.cpp:11-67)
		std::vector<ElementAssemblyValues> cache;

		// displacement_prev / dt of the NL entry points (ViscousDamping)
		std::vector<double> x_prev;
		double dt = 1.0;

		// result of the last matrix assembly
		std::vector<int32_t> outer, inner;
		std::vector<double> values;
		double loop_seconds = 0, merge_seconds = 0;

		// AssemblyValsCache::compute (AssemblyValsCache.cpp:52-67)
		void cache_compute(int e, ElementAssemblyValues &vals) const
		{
			if (cache.empty())
				compute_assembly_values(*this, e, vals);
			else
				vals = cache[e]; // deep copy, as the reference does
		}
	};

	void compute_assembly_values(const Problem &pb, int e, ElementAssemblyValues &vals)
	{
		const int n_loc = pb.d.n_loc, n_qp = pb.d.n_qp;
		vals.element_id = e;
		vals.n_loc = n_loc;
		vals.n_qp = n_qp;
		vals.global.resize(n_loc);
		vals.grad.resize(size_t(n_loc) * n_qp * 3);
		vals.grad_t_m.resize(size_t(n_loc) * n_qp * 3);
		vals.jac_it.resize(size_t(n_qp) * 9);
		vals.det.resize(n_qp);
		vals.weights = pb.qw;
		for (int j = 0; j < n_loc; ++j)
			vals.global[j] = pb.conn[size_t(e) * n_loc + j];

		// basis.evaluate_grads(pts, basis_values)  (ElementAssemblyValues.cpp:181-183)
		if (pb.d.use_cache == 0 && !pb.lattice.empty())
		{
			std::vector<double> tmp(size_t(n_loc) * 3);
			for (int q = 0; q < n_qp; ++q)
			{
				lagrange_grads(pb.d.basis_order, n_loc, pb.lattice.data(), &pb.qpts[size_t(q) * 3], tmp.data());
				for (int j = 0; j < n_loc; ++j)
					for (int c = 0; c < 3; ++c)
						vals.grad[(size_t(j) * n_qp + q) * 3 + c] = tmp[size_t(j) * 3 + c];
			}
		}
		else
		{
			for (int q = 0; q < n_qp; ++q)
				for (int j = 0; j < n_loc; ++j)
					for (int c = 0; c < 3; ++c)
						vals.grad[(size_t(j) * n_qp + q) * 3 + c] = pb.ref_grads[(size_t(q) * n_loc + j) * 3 + c];
		}

		// basis.evaluate_bases(pts, basis_values) (ElementAssemblyValues.cpp:178-180): from the table
		if (!pb.ref_vals.empty())
		{
			vals.val.resize(size_t(n_loc) * n_qp);
			for (int q = 0; q < n_qp; ++q)
				for (int j = 0; j < n_loc; ++j)
					vals.val[size_t(j) * n_qp + q] = pb.ref_vals[size_t(q) * n_loc + j];
		}

		// finalize3d (ElementAssemblyValues.cpp:65-104): tmp.row(c) += dN_j/dxi_c * node_j over the geometric bases; P1
		// geometry (gradients (-1,-1,-1), e_x, e_y, e_z) unless isoparametric nodes were given
		static const double gg[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
		const bool iso = pb.d.geom_order > 1 && !pb.geom_nodes.empty();
		const int ngl = iso ? pb.d.n_geom_loc : 4;
		const double *vx = iso ? &pb.geom_nodes[size_t(e) * ngl * 3] : &pb.vertices[size_t(e) * 12];
		std::vector<double> ggrad(iso ? size_t(ngl) * 3 : 0);
		for (int k = 0; k < n_qp; ++k)
		{
			double tmp[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
			if (iso)
				lagrange_grads(pb.d.geom_order, ngl, pb.geom_lattice.data(), &pb.qpts[size_t(k) * 3], ggrad.data());
			for (int j = 0; j < ngl; ++j)
				for (int c = 0; c < 3; ++c)
					for (int d = 0; d < 3; ++d)
						tmp[c * 3 + d] += (iso ? ggrad[size_t(j) * 3 + c] : gg[j][c]) * vx[j * 3 + d];
			vals.det[k] = det3(tmp);
			double inv[9];
			inverse3(tmp, inv);
			double *jit = &vals.jac_it[size_t(k) * 9];
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					jit[r * 3 + c] = inv[c * 3 + r]; // tmp.inverse().transpose()
			for (int j = 0; j < n_loc; ++j)
			{
				const double *g = vals.g(j, k);
				double *gt = &vals.grad_t_m[(size_t(j) * n_qp + k) * 3];
				for (int c = 0; c < 3; ++c)
					gt[c] = g[0] * jit[0 * 3 + c] + g[1] * jit[1 * 3 + c] + g[2] * jit[2 * 3 + c];
			}
		}
	}

	// assembler/AssemblerData.hpp:7-51 — what the local assemblers see
	struct NLData
	{
		const ElementAssemblyValues &vals;
		const double *x;
		const std::vector<double> &da;
		double lambda, mu; // params_.lambda_mu(...) for a per-element constant material
		double param3 = 0.0; // MooneyRivlin: (c1, c2, k) = (lambda, mu, param3)
		const double *x_prev = nullptr; // displacement_prev (ViscousDamping); nullptr: sizes differ in the reference's terms
		double dt = 1.0;
	};

	// =======================================================================================
	// NeoHookeanElasticity (assembler/NeoHookeanElasticity.cpp)
	// =======================================================================================

	// utils/ElasticityUtils.hpp:70-101 get_local_disp
	template <typename T>
	void get_local_disp(const NLData &data, int size, std::vector<T> &local_disp)
	{
		const int n_loc = data.vals.n_loc;
		g_nvars = n_loc * size;
		local_disp.clear();
		local_disp.reserve(size_t(n_loc) * size);
		for (int i = 0; i < n_loc; ++i)
			for (int d = 0; d < size; ++d)
				local_disp.push_back(Alloc<T>::var(i * size + d, data.x[size_t(data.vals.global[i]) * size + d]));
	}

	// utils/ElasticityUtils.hpp:103-136 compute_disp_grad_at_quad (3D)
	template <typename T>
	void compute_disp_grad_at_quad(const NLData &data, const std::vector<T> &local_disp, int p, T *def_grad /*[9]*/)
	{
		T acc[9];
		for (int k = 0; k < 9; ++k)
			acc[k] = T(0.0);
		for (int i = 0; i < data.vals.n_loc; ++i)
		{
			const double *grad = data.vals.g(i, p);
			for (int d = 0; d < 3; ++d)
				for (int c = 0; c < 3; ++c)
					acc[d * 3 + c] = acc[d * 3 + c] + T(grad[c]) * local_disp[size_t(i) * 3 + d];
		}
		const double *jit = &data.vals.jac_it[size_t(p) * 9];
		for (int r = 0; r < 3; ++r)
			for (int c = 0; c < 3; ++c)
			{
				T s = T(0.0);
				for (int k = 0; k < 3; ++k)
					s = s + acc[r * 3 + k] * T(jit[k * 3 + c]);
				def_grad[r * 3 + c] = s;
			}
	}

	// NeoHookeanElasticity.cpp:338-419, autodiff branch (T != double): used as the
	// "NeoHookeanAutodiff" cross-check of tests/test_assembler.cpp:148-315.
	template <typename T>
	T neohookean_energy_autodiff(const NLData &data)
	{
		std::vector<T> local_disp;
		get_local_disp<T>(data, 3, local_disp);
		T energy = T(0.0);
		T F[9];
		for (int p = 0; p < data.vals.n_qp; ++p)
		{
			compute_disp_grad_at_quad<T>(data, local_disp, p, F);
			for (int d = 0; d < 3; ++d)
				F[d * 3 + d] = F[d * 3 + d] + T(1.0);
			const T log_det_j = log(det3<T>(F));
			T tr = T(0.0); // (F^T F).trace()
			for (int k = 0; k < 9; ++k)
				tr = tr + F[k] * F[k];
			const T val = T(data.mu / 2) * (tr - T(3.0) - T(2.0) * log_det_j) + T(data.lambda / 2) * log_det_j * log_det_j;
			energy = energy + val * T(data.da[p]);
		}
		return energy;
	}

	// gather of the local displacement (NeoHookeanElasticity.cpp:343-355, 461-473, 555-564)
	void gather_local_disp(const NLData &data, std::vector<double> &local_disp /*[n_loc][3]*/)
	{
		const int n_loc = data.vals.n_loc;
		local_disp.assign(size_t(n_loc) * 3, 0.0);
		for (int i = 0; i < n_loc; ++i)
			for (int d = 0; d < 3; ++d)
				local_disp[size_t(i) * 3 + d] += 1.0 * data.x[size_t(data.vals.global[i]) * 3 + d];
	}

	// NeoHookeanElasticity.cpp:338-388 compute_energy_aux<double>
	double neohookean_energy(const NLData &data)
	{
		const int n_loc = data.vals.n_loc;
		std::vector<double> u;
		gather_local_disp(data, u);
		double energy = 0.0;
		for (int p = 0; p < data.vals.n_qp; ++p)
		{
			// (local_disp^T * grad) * jac_it + Id     (:377)
			double A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
			for (int i = 0; i < n_loc; ++i)
			{
				const double *g = data.vals.g(i, p);
				for (int d = 0; d < 3; ++d)
					for (int c = 0; c < 3; ++c)
						A[d * 3 + c] += u[size_t(i) * 3 + d] * g[c];
			}
			const double *jit = &data.vals.jac_it[size_t(p) * 9];
			double F[9];
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					F[r * 3 + c] = A[r * 3 + 0] * jit[0 * 3 + c] + A[r * 3 + 1] * jit[1 * 3 + c] + A[r * 3 + 2] * jit[2 * 3 + c] + (r == c ? 1.0 : 0.0);
			const double J = det3(F);
			const double log_det_j = std::log(J);
			double sq = 0.0;
			for (int k = 0; k < 9; ++k)
				sq += F[k] * F[k];
			const double val = data.mu / 2 * (sq - 3 - 2 * log_det_j) + data.lambda / 2 * log_det_j * log_det_j; // (:384)
			energy += val * data.da[p];
		}
		return energy;
	}

	void cross3(const double *x, const double *y, double *z)
	{
		z[0] = x[1] * y[2] - x[2] * y[1];
		z[1] = x[2] * y[0] - x[0] * y[2];
		z[2] = x[0] * y[1] - x[1] * y[0];
	}

	// F = local_disp^T * delF_delU + Id with delF_delU = grad * jac_it (:493-496, :582-585)
	void neohookean_def_grad(const NLData &data, const std::vector<double> &u, int p, std::vector<double> &delF_delU /*[n_loc][3]*/, double *F)
	{
		const int n_loc = data.vals.n_loc;
		const double *jit = &data.vals.jac_it[size_t(p) * 9];
		delF_delU.resize(size_t(n_loc) * 3);
		for (int i = 0; i < n_loc; ++i)
		{
			const double *g = data.vals.g(i, p);
			for (int c = 0; c < 3; ++c)
				delF_delU[size_t(i) * 3 + c] = g[0] * jit[0 * 3 + c] + g[1] * jit[1 * 3 + c] + g[2] * jit[2 * 3 + c];
		}
		for (int r = 0; r < 3; ++r)
			for (int c = 0; c < 3; ++c)
			{
				double s = 0.0;
				for (int i = 0; i < n_loc; ++i)
					s += u[size_t(i) * 3 + r] * delF_delU[size_t(i) * 3 + c];
				F[r * 3 + c] = s + (r == c ? 1.0 : 0.0);
			}
	}

	// delJ_delF columns by cross products of the columns of F (:513-527, :613-619); row-major C[r*3+c]
	void neohookean_cofactor(const double *F, double *C, double *u, double *v, double *w)
	{
		for (int r = 0; r < 3; ++r)
		{
			u[r] = F[r * 3 + 0];
			v[r] = F[r * 3 + 1];
			w[r] = F[r * 3 + 2];
		}
		double c0[3], c1[3], c2[3];
		cross3(v, w, c0);
		cross3(w, u, c1);
		cross3(u, v, c2);
		for (int r = 0; r < 3; ++r)
		{
			C[r * 3 + 0] = c0[r];
			C[r * 3 + 1] = c1[r];
			C[r * 3 + 2] = c2[r];
		}
	}

	// NeoHookeanElasticity.cpp:453-545 compute_energy_aux_gradient_fast; output G_flat[i*3+d]
	void neohookean_gradient(const NLData &data, std::vector<double> &G_flat)
	{
		const int n_loc = data.vals.n_loc;
		std::vector<double> u, delF_delU;
		gather_local_disp(data, u);
		std::vector<double> G(size_t(n_loc) * 3, 0.0);
		for (int p = 0; p < data.vals.n_qp; ++p)
		{
			double F[9], C[9], cu[3], cv[3], cw[3];
			neohookean_def_grad(data, u, p, delF_delU, F);
			const double J = det3(F);
			const double log_det_j = std::log(J);
			neohookean_cofactor(F, C, cu, cv, cw);
			// gradient_temp = mu F - mu (1/J) dJ/dF + lambda log(J) (1/J) dJ/dF    (:532)
			double P[9];
			for (int k = 0; k < 9; ++k)
				P[k] = data.mu * F[k] - data.mu * (1 / J) * C[k] + data.lambda * log_det_j * (1 / J) * C[k];
			// gradient = delF_delU * gradient_temp^T ; G += gradient * da(p)       (:533-537)
			for (int i = 0; i < n_loc; ++i)
				for (int a = 0; a < 3; ++a)
				{
					double s = 0.0;
					for (int k = 0; k < 3; ++k)
						s += delF_delU[size_t(i) * 3 + k] * P[a * 3 + k];
					G[size_t(i) * 3 + a] += s * data.da[p];
				}
		}
		G_flat = G; // G^T flattened column-major == node-major i*dim+d (:540-544)
	}

	void hat3(const double *x, double *m /*row-major 3x3*/)
	{
		for (int k = 0; k < 9; ++k)
			m[k] = 0.0;
		m[0 * 3 + 1] = -x[2];
		m[0 * 3 + 2] = x[1];
		m[1 * 3 + 0] = x[2];
		m[1 * 3 + 2] = -x[0];
		m[2 * 3 + 0] = -x[1];
		m[2 * 3 + 1] = x[0];
	}

	// NeoHookeanElasticity.cpp:547-658 compute_energy_hessian_aux_fast; H row-major [N][N]
	void neohookean_hessian(const NLData &data, std::vector<double> &H)
	{
		const int n_loc = data.vals.n_loc, N = n_loc * 3;
		std::vector<double> u, delF_delU;
		gather_local_disp(data, u);
		H.assign(size_t(N) * N, 0.0);
		std::vector<double> B(size_t(9) * N), T(size_t(N) * 9);
		for (int p = 0; p < data.vals.n_qp; ++p)
		{
			double F[9], C[9], cu[3], cv[3], cw[3];
			neohookean_def_grad(data, u, p, delF_delU, F);
			const double J = det3(F);
			const double log_det_j = std::log(J);
			neohookean_cofactor(F, C, cu, cv, cw);

			// del2J_delF2 (9x9, vec(F) column-major: index r + 3c), blocks of hat() (:621-626)
			double d2J[81];
			for (int k = 0; k < 81; ++k)
				d2J[k] = 0.0;
			double hu[9], hv[9], hw[9];
			hat3(cu, hu);
			hat3(cv, hv);
			hat3(cw, hw);
			auto set_block = [&](int r0, int c0, const double *m, double sgn) {
				for (int r = 0; r < 3; ++r)
					for (int c = 0; c < 3; ++c)
						d2J[(r0 + r) * 9 + (c0 + c)] = sgn * m[r * 3 + c];
			};
			set_block(0, 6, hv, 1.0);
			set_block(6, 0, hv, -1.0);
			set_block(0, 3, hw, -1.0);
			set_block(3, 0, hw, 1.0);
			set_block(3, 6, hu, -1.0);
			set_block(6, 3, hu, 1.0);

			// g_j = vec(delJ_delF) column-major (:630)
			double gj[9];
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					gj[r + 3 * c] = C[r * 3 + c];

			// hessian_temp (:632-636)
			double HF[81];
			const double c1 = (data.mu + data.lambda * (1 - log_det_j)) / (J * J);
			const double c2 = (data.lambda * log_det_j - data.mu) / J;
			for (int r = 0; r < 9; ++r)
				for (int c = 0; c < 9; ++c)
					HF[r * 9 + c] = (r == c ? data.mu : 0.0) + c1 * (gj[r] * gj[c]) + c2 * d2J[r * 9 + c];

			// delF_delU_tensor: column i*3+j = vec of the matrix whose row j is delF_delU.row(i) (:638-650)
			std::fill(B.begin(), B.end(), 0.0);
			for (int i = 0; i < n_loc; ++i)
				for (int j = 0; j < 3; ++j)
					for (int k = 0; k < 3; ++k)
						B[size_t(j + 3 * k) * N + (i * 3 + j)] = delF_delU[size_t(i) * 3 + k];

			// hessian = B^T * hessian_temp * B (two dense products, :652) ; H += hessian * da(p)
			for (int r = 0; r < N; ++r)
				for (int c = 0; c < 9; ++c)
				{
					double s = 0.0;
					for (int k = 0; k < 9; ++k)
						s += B[size_t(k) * N + r] * HF[k * 9 + c];
					T[size_t(r) * 9 + c] = s;
				}
			for (int r = 0; r < N; ++r)
				for (int c = 0; c < N; ++c)
				{
					double s = 0.0;
					for (int k = 0; k < 9; ++k)
						s += T[size_t(r) * 9 + k] * B[size_t(k) * N + c];
					H[size_t(r) * N + c] += s * data.da[p];
				}
		}
	}

	// =======================================================================================
	// LinearElasticity (assembler/LinearElasticity.cpp) and Laplacian (assembler/Laplacian.cpp)
	// =======================================================================================

	// LinearElasticity.cpp:30-63 ; res index jj*size+ii
	void linear_elasticity_local(const ElementAssemblyValues &vals, const std::vector<double> &da, int i, int j, double lambda, double mu, double *res /*[9]*/)
	{
		for (int k = 0; k < 9; ++k)
			res[k] = 0.0;
		for (int k = 0; k < vals.n_qp; ++k)
		{
			const double *gi = vals.gt(i, k), *gj = vals.gt(j, k);
			// outer = gradi^T gradj, column-major linear index r + 3c
			double outer[9];
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					outer[r + 3 * c] = gi[r] * gj[c];
			const double dot = gi[0] * gj[0] + gi[1] * gj[1] + gi[2] * gj[2];
			double res_k[9];
			for (int ii = 0; ii < 3; ++ii)
				for (int jj = 0; jj < 3; ++jj)
				{
					res_k[jj * 3 + ii] = outer[ii * 3 + jj] * mu + outer[jj * 3 + ii] * lambda;
					if (ii == jj)
						res_k[jj * 3 + ii] += mu * dot;
				}
			for (int q = 0; q < 9; ++q)
				res[q] += res_k[q] * da[k];
		}
	}

	// Laplacian.cpp:13-26
	double laplacian_local(const ElementAssemblyValues &vals, const std::vector<double> &da, int i, int j)
	{
		double res = 0;
		for (int k = 0; k < vals.n_qp; ++k)
		{
			const double *gi = vals.gt(i, k), *gj = vals.gt(j, k);
			res += (gi[0] * gj[0] + gi[1] * gj[1] + gi[2] * gj[2]) * da[k];
		}
		return res;
	}

	// Mass.cpp:5-23: tmp = sum_q rho * phi_i(q) * phi_j(q) * da(q) on the diagonal of the size x size block
	void mass_local(const ElementAssemblyValues &vals, const std::vector<double> &da, int i, int j, double rho, int size, double *blk)
	{
		double tmp = 0;
		for (int q = 0; q < vals.n_qp; ++q)
			tmp += rho * vals.val[size_t(i) * vals.n_qp + q] * vals.val[size_t(j) * vals.n_qp + q] * da[q];
		for (int k = 0; k < size * size; ++k)
			blk[k] = 0.0;
		for (int k = 0; k < size; ++k)
			blk[k * size + k] = tmp;
	}

	// LinearElasticity.cpp:106-134 compute_energy_aux<T>
	template <typename T>
	T linear_elasticity_energy(const NLData &data)
	{
		std::vector<T> local_disp;
		get_local_disp<T>(data, 3, local_disp);
		T energy = T(0.0);
		T G[9];
		for (int p = 0; p < data.vals.n_qp; ++p)
		{
			compute_disp_grad_at_quad<T>(data, local_disp, p, G);
			T strain[9];
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					strain[r * 3 + c] = (G[r * 3 + c] + G[c * 3 + r]) / T(2.0);
			T tr2 = T(0.0); // (strain^T strain).trace()
			for (int k = 0; k < 9; ++k)
				tr2 = tr2 + strain[k] * strain[k];
			const T tr = strain[0] + strain[4] + strain[8];
			const T val = T(data.mu) * tr2 + T(data.lambda / 2) * tr * tr;
			energy = energy + val * T(data.da[p]);
		}
		return energy;
	}

	// SaintVenantElasticity.cpp:9-20, 219-266 compute_energy_aux<T> with the isotropic elasticity tensor of
	// MatParams.cpp:211-253 (set_from_lambda_mu: C = lambda 1 (x) 1 + 2 mu I_sym in Voigt form with engineering shear):
	// strain = (G^T G + G + G^T) / 2, stress = C : strain, energy = 1/2 sum_p (stress * strain).trace() da_p.
	// Gradient and Hessian are the forward-mode derivatives of this function, as in the reference
	// (gradient_from_energy / hessian_from_energy, SaintVenantElasticity.cpp:88-129).
	template <typename T>
	T saint_venant_energy(const NLData &data)
	{
		std::vector<T> local_disp;
		get_local_disp<T>(data, 3, local_disp);
		T energy = T(0.0);
		T G[9];
		const double C00 = 2.0 * data.mu + data.lambda, C01 = data.lambda, C33 = data.mu;
		for (int p = 0; p < data.vals.n_qp; ++p)
		{
			compute_disp_grad_at_quad<T>(data, local_disp, p, G);
			T strain[9]; // strain_from_disp_grad
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
				{
					T gtg = T(0.0);
					for (int k = 0; k < 3; ++k)
						gtg = gtg + G[k * 3 + r] * G[k * 3 + c];
					strain[r * 3 + c] = (gtg + G[r * 3 + c] + G[c * 3 + r]) * T(0.5);
				}
			const T eps[6] = {strain[0], strain[4], strain[8], T(2.0) * strain[5], T(2.0) * strain[2], T(2.0) * strain[1]};
			T sig[6]; // stress(elasticity_tensor, eps, j)
			sig[0] = T(C00) * eps[0] + T(C01) * eps[1] + T(C01) * eps[2];
			sig[1] = T(C01) * eps[0] + T(C00) * eps[1] + T(C01) * eps[2];
			sig[2] = T(C01) * eps[0] + T(C01) * eps[1] + T(C00) * eps[2];
			sig[3] = T(C33) * eps[3];
			sig[4] = T(C33) * eps[4];
			sig[5] = T(C33) * eps[5];
			const T st[9] = {sig[0], sig[5], sig[4], sig[5], sig[1], sig[3], sig[4], sig[3], sig[2]};
			T tr = T(0.0); // (stress_tensor * strain).trace()
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					tr = tr + st[r * 3 + c] * strain[c * 3 + r];
			energy = energy + tr * T(data.da[p]);
		}
		return energy * T(0.5);
	}

	// MooneyRivlinElasticity.hpp:26-47 elastic_energy inside GenericElastic::compute_energy_aux (GenericElastic.hpp:93-132):
	// def_grad = I + grad u, J = det, F~ = def_grad / J^(1/3), C~ = F~ F~^T, I1~ = tr C~, I2~ = (tr^2 - tr(C~ C~)) / 2
	// (utils/ElasticityUtils.hpp:139-150), val = c1 (I1~ - 3) + c2 (I2~ - 3) + k/2 ln^2 J; energy = sum_p val da_p.
	// Gradient and Hessian by forward-mode autodiff over the local dofs (AutodiffType::FULL, GenericElastic.cpp:137-175).
	template <typename T>
	T mooney_rivlin_energy(const NLData &data)
	{
		using std::log;
		using std::pow;
		std::vector<T> local_disp;
		get_local_disp<T>(data, 3, local_disp);
		T energy = T(0.0);
		T F[9];
		const double c1 = data.lambda, c2 = data.mu, k = data.param3;
		for (int p = 0; p < data.vals.n_qp; ++p)
		{
			compute_disp_grad_at_quad<T>(data, local_disp, p, F);
			F[0] = F[0] + T(1.0);
			F[4] = F[4] + T(1.0);
			F[8] = F[8] + T(1.0);
			const T J = det3(F);
			const T log_J = log(J);
			const T scale = pow(J, 1.0 / 3.0);
			T Ft[9], C[9];
			for (int i = 0; i < 9; ++i)
				Ft[i] = F[i] / scale;
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
				{
					T sum = T(0.0);
					for (int m = 0; m < 3; ++m)
						sum = sum + Ft[r * 3 + m] * Ft[c * 3 + m];
					C[r * 3 + c] = sum;
				}
			const T I1 = C[0] + C[4] + C[8];
			T trCC = T(0.0);
			for (int r = 0; r < 3; ++r)
				for (int m = 0; m < 3; ++m)
					trCC = trCC + C[r * 3 + m] * C[m * 3 + r];
			const T I2 = T(0.5) * (I1 * I1 - trCC);
			const T val = T(c1) * (I1 - T(3.0)) + T(c2) * (I2 - T(3.0)) + T(k / 2) * (log_J * log_J);
			energy = energy + val * T(data.da[p]);
		}
		return energy;
	}

	// =======================================================================================
	// ViscousDamping (assembler/ViscousDamping.cpp): R = psi |dE/dt|^2 + phi/2 tr(dE/dt)^2 with
	// dE/dt = sym(dF/dt^T F), dF/dt = (F - F_prev) / dt. (psi, phi) = (data.lambda, data.mu).
	// =======================================================================================
	struct VdPoint
	{
		double F[9], Fd[9], Ed[9]; // def_grad, dFdt, dEdt (row-major)
	};

	// local_disp / local_prev_disp / local_vel and, per point, def_grad = local_disp^T delF_delU + I, dFdt = local_vel^T delF_delU
	// (ViscousDamping.cpp:127-160, 304-334), delF_delU = grad * jac_it = grad_t_m
	void vd_point(const NLData &data, const std::vector<double> &u, const std::vector<double> &vel, int p, VdPoint &pt)
	{
		for (int k = 0; k < 9; ++k)
			pt.F[k] = pt.Fd[k] = 0.0;
		for (int i = 0; i < data.vals.n_loc; ++i)
		{
			const double *gt = data.vals.gt(i, p);
			for (int a = 0; a < 3; ++a)
				for (int d = 0; d < 3; ++d)
				{
					pt.F[a * 3 + d] += u[size_t(i) * 3 + a] * gt[d];
					pt.Fd[a * 3 + d] += vel[size_t(i) * 3 + a] * gt[d];
				}
		}
		pt.F[0] += 1.0;
		pt.F[4] += 1.0;
		pt.F[8] += 1.0;
		double M[9]; // dFdt^T F
		for (int r = 0; r < 3; ++r)
			for (int c = 0; c < 3; ++c)
				M[r * 3 + c] = pt.Fd[0 + r] * pt.F[0 + c] + pt.Fd[3 + r] * pt.F[3 + c] + pt.Fd[6 + r] * pt.F[6 + c];
		for (int r = 0; r < 3; ++r)
			for (int c = 0; c < 3; ++c)
				pt.Ed[r * 3 + c] = (M[r * 3 + c] + M[c * 3 + r]) / 2.;
	}

	void vd_local_fields(const NLData &data, std::vector<double> &u, std::vector<double> &vel)
	{
		const int n_loc = data.vals.n_loc;
		u.assign(size_t(n_loc) * 3, 0.0);
		vel.assign(size_t(n_loc) * 3, 0.0);
		for (int i = 0; i < n_loc; ++i)
			for (int d = 0; d < 3; ++d)
			{
				const size_t g = size_t(data.vals.global[i]) * 3 + d;
				u[size_t(i) * 3 + d] = data.x[g];
				vel[size_t(i) * 3 + d] = (data.x[g] - data.x_prev[g]) / data.dt;
			}
	}

	// ViscousDamping.cpp:297-342 compute_energy
	double viscous_damping_energy(const NLData &data)
	{
		if (data.x_prev == nullptr)
			return 0.0;
		std::vector<double> u, vel;
		vd_local_fields(data, u, vel);
		double energy = 0.0;
		for (int p = 0; p < data.vals.n_qp; ++p)
		{
			VdPoint pt;
			vd_point(data, u, vel, p, pt);
			double sq = 0.0;
			for (int k = 0; k < 9; ++k)
				sq += pt.Ed[k] * pt.Ed[k];
			const double tr = pt.Ed[0] + pt.Ed[4] + pt.Ed[8];
			energy += (data.lambda * sq + 0.5 * data.mu * tr * tr) * data.da[p];
		}
		return energy;
	}

	// ViscousDamping.cpp:122-170 assemble_gradient: G += delF_delU ((dFdt + F / dt) tmp)^T da, tmp = 2 psi dEdt + phi tr(dEdt) I
	void viscous_damping_gradient(const NLData &data, std::vector<double> &g)
	{
		const int n_loc = data.vals.n_loc;
		g.assign(size_t(n_loc) * 3, 0.0);
		if (data.x_prev == nullptr)
			return;
		std::vector<double> u, vel;
		vd_local_fields(data, u, vel);
		for (int p = 0; p < data.vals.n_qp; ++p)
		{
			VdPoint pt;
			vd_point(data, u, vel, p, pt);
			const double tr = pt.Ed[0] + pt.Ed[4] + pt.Ed[8];
			double tmp[9], lhs[9], S[9];
			for (int k = 0; k < 9; ++k)
			{
				tmp[k] = 2 * data.lambda * pt.Ed[k] + (k % 4 == 0 ? data.mu * tr : 0.0);
				lhs[k] = pt.Fd[k] + pt.F[k] / data.dt;
			}
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					S[r * 3 + c] = lhs[r * 3 + 0] * tmp[0 + c] + lhs[r * 3 + 1] * tmp[3 + c] + lhs[r * 3 + 2] * tmp[6 + c];
			for (int i = 0; i < n_loc; ++i)
			{
				const double *gt = data.vals.gt(i, p);
				for (int a = 0; a < 3; ++a)
					g[size_t(i) * 3 + a] += (gt[0] * S[a * 3 + 0] + gt[1] * S[a * 3 + 1] + gt[2] * S[a * 3 + 2]) * data.da[p];
			}
		}
	}

	// ViscousDamping.cpp:16-62 compute_stress_grad_aux: the three 9 x 9 second derivatives of R with respect to (F, dF/dt),
	// F(i, j) at index i * 3 + j
	void vd_second_derivatives(const VdPoint &pt, double psi, double phi, double *RFF, double *RFD, double *RDD)
	{
		double dEdF[81], dEdD[81]; // d dEdt(i,j) / d F(p,q), / d dFdt(p,q)
		for (int k = 0; k < 81; ++k)
			dEdF[k] = dEdD[k] = RFF[k] = RFD[k] = RDD[k] = 0.0;
		for (int i = 0; i < 3; ++i)
			for (int j = 0; j < 3; ++j)
				for (int p = 0; p < 3; ++p)
				{
					dEdF[(i * 3 + j) * 9 + p * 3 + j] += pt.Fd[p * 3 + i] / 2.;
					dEdD[(i * 3 + j) * 9 + p * 3 + j] += pt.F[p * 3 + i] / 2.;
					dEdF[(i * 3 + j) * 9 + p * 3 + i] += pt.Fd[p * 3 + j] / 2.;
					dEdD[(i * 3 + j) * 9 + p * 3 + i] += pt.F[p * 3 + j] / 2.;
				}
		for (int i = 0; i < 3; ++i)
			for (int j = 0; j < 3; ++j)
			{
				const int idx = i * 3 + j;
				for (int k = 0; k < 3; ++k)
				{
					for (int c = 0; c < 9; ++c)
					{
						RFF[idx * 9 + c] += (2 * psi) * pt.Fd[i * 3 + k] * dEdF[(k * 3 + j) * 9 + c] + phi * pt.Fd[i * 3 + j] * dEdF[(k * 3 + k) * 9 + c];
						RDD[idx * 9 + c] += (2 * psi) * pt.F[i * 3 + k] * dEdD[(k * 3 + j) * 9 + c] + phi * pt.F[i * 3 + j] * dEdD[(k * 3 + k) * 9 + c];
						RFD[idx * 9 + c] += (2 * psi) * (pt.Fd[i * 3 + k] * dEdD[(k * 3 + j) * 9 + c]) + phi * (dEdD[(k * 3 + k) * 9 + c] * pt.Fd[i * 3 + j]);
					}
					RFD[idx * 9 + i * 3 + k] += 2 * psi * pt.Ed[k * 3 + j];
					RFD[idx * 9 + idx] += phi * pt.Ed[k * 3 + k];
				}
			}
	}

	// ViscousDamping.cpp:173-229 assemble_hessian: sum_p B^T (RFF + (RFD + RFD^T) / dt + RDD / dt^2) B da with
	// B(3 j + d, 3 i + j) = delF_delU(i, d)
	void viscous_damping_hessian(const NLData &data, std::vector<double> &h)
	{
		const int n_loc = data.vals.n_loc, N = 3 * n_loc;
		h.assign(size_t(N) * N, 0.0);
		if (data.x_prev == nullptr)
			return;
		std::vector<double> u, vel;
		vd_local_fields(data, u, vel);
		for (int p = 0; p < data.vals.n_qp; ++p)
		{
			VdPoint pt;
			vd_point(data, u, vel, p, pt);
			double RFF[81], RFD[81], RDD[81], T[81];
			vd_second_derivatives(pt, data.lambda, data.mu, RFF, RFD, RDD);
			for (int r = 0; r < 9; ++r)
				for (int c = 0; c < 9; ++c)
					T[r * 9 + c] = RFF[r * 9 + c] + (1. / data.dt) * (RFD[r * 9 + c] + RFD[c * 9 + r]) + (1. / data.dt / data.dt) * RDD[r * 9 + c];
			for (int i = 0; i < n_loc; ++i)
				for (int j = 0; j < n_loc; ++j)
				{
					const double *gi = data.vals.gt(i, p), *gj = data.vals.gt(j, p);
					for (int a = 0; a < 3; ++a)
						for (int b = 0; b < 3; ++b)
						{
							double s = 0.0;
							for (int d = 0; d < 3; ++d)
								for (int d2 = 0; d2 < 3; ++d2)
									s += gi[d] * T[(a * 3 + d) * 9 + b * 3 + d2] * gj[d2];
							h[size_t(i * 3 + a) * N + j * 3 + b] += s * data.da[p];
						}
				}
		}
	}

	// =======================================================================================
	// ipc::project_to_psd (ipc-toolkit, source not in /root/reference; called at
	// assembler/Assembler.cpp:693-694). Documented behaviour: symmetric eigendecomposition,
	// return A unchanged if the smallest eigenvalue is >= 0, else clamp negative eigenvalues
	// to 0 and rebuild. Eigen-solver: cyclic Jacobi.
	// =======================================================================================
	void jacobi_eigen(int n, std::vector<double> &A, std::vector<double> &V, std::vector<double> &w)
	{
		V.assign(size_t(n) * n, 0.0);
		for (int i = 0; i < n; ++i)
			V[size_t(i) * n + i] = 1.0;
		for (int sweep = 0; sweep < 100; ++sweep)
		{
			double off = 0.0, diag = 0.0;
			for (int i = 0; i < n; ++i)
				for (int j = 0; j < n; ++j)
					(i == j ? diag : off) += A[size_t(i) * n + j] * A[size_t(i) * n + j];
			if (off <= 1e-30 * diag || off == 0.0)
				break;
			for (int p = 0; p < n - 1; ++p)
				for (int q = p + 1; q < n; ++q)
				{
					const double apq = A[size_t(p) * n + q];
					if (apq == 0.0)
						continue;
					const double app = A[size_t(p) * n + p], aqq = A[size_t(q) * n + q];
					const double theta = (aqq - app) / (2.0 * apq);
					const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
					const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
					for (int k = 0; k < n; ++k)
					{
						const double akp = A[size_t(k) * n + p], akq = A[size_t(k) * n + q];
						A[size_t(k) * n + p] = c * akp - s * akq;
						A[size_t(k) * n + q] = s * akp + c * akq;
					}
					for (int k = 0; k < n; ++k)
					{
						const double apk = A[size_t(p) * n + k], aqk = A[size_t(q) * n + k];
						A[size_t(p) * n + k] = c * apk - s * aqk;
						A[size_t(q) * n + k] = s * apk + c * aqk;
					}
					for (int k = 0; k < n; ++k)
					{
						const double vkp = V[size_t(k) * n + p], vkq = V[size_t(k) * n + q];
						V[size_t(k) * n + p] = c * vkp - s * vkq;
						V[size_t(k) * n + q] = s * vkp + c * vkq;
					}
				}
		}
		w.resize(n);
		for (int i = 0; i < n; ++i)
			w[i] = A[size_t(i) * n + i];
	}

	void project_to_psd(int n, std::vector<double> &A)
	{
		bool zero = true;
		for (double v : A)
			if (v != 0.0)
				zero = false;
		if (zero)
			return;
		std::vector<double> M = A, V, w;
		// symmetrise the working copy (SelfAdjointEigenSolver reads one triangle)
		for (int i = 0; i < n; ++i)
			for (int j = 0; j < i; ++j)
				M[size_t(j) * n + i] = M[size_t(i) * n + j];
		for (double v : M)
			if (!std::isfinite(v))
				return; // eigensolver failure -> matrix returned unchanged
		jacobi_eigen(n, M, V, w);
		double wmin = w[0];
		for (double v : w)
			wmin = std::min(wmin, v);
		if (wmin >= 0.0)
			return;
		for (double &v : w)
			if (v < 0.0)
				v = 0.0;
		for (int i = 0; i < n; ++i)
			for (int j = 0; j < n; ++j)
			{
				double s = 0.0;
				for (int k = 0; k < n; ++k)
					s += V[size_t(i) * n + k] * w[k] * V[size_t(j) * n + k];
				A[size_t(i) * n + j] = s;
			}
	}

	// =======================================================================================
	// Sparse matrix pieces of Eigen that define the result pattern (third-party, Eigen 3.4.0):
	// setFromTriplets (duplicates summed, explicit zeros kept, inner indices ascending),
	// sparse += (pattern union), compressed column-major storage with int indices.
	// =======================================================================================
	struct Triplet
	{
		int row, col;
		double value;
	};

	struct SparseCSC
	{
		int rows = 0, cols = 0;
		std::vector<int> outer; // cols + 1
		std::vector<int> inner;
		std::vector<double> val;
		size_t nnz() const { return inner.size(); }
		void resize(int r, int c)
		{
			rows = r;
			cols = c;
			outer.assign(size_t(c) + 1, 0);
			inner.clear();
			val.clear();
		}
	};

	SparseCSC from_triplets(int rows, int cols, const std::vector<Triplet> &t)
	{
		SparseCSC m;
		m.resize(rows, cols);
		if (t.empty())
			return m;
		// counting sort by column, then stable sort by row inside each column; duplicates are
		// accumulated in insertion order like Eigen's collapse of the row-major temporary.
		std::vector<size_t> count(size_t(cols) + 1, 0);
		for (const Triplet &e : t)
			++count[size_t(e.col) + 1];
		for (int c = 0; c < cols; ++c)
			count[size_t(c) + 1] += count[c];
		std::vector<uint32_t> order(t.size());
		{
			std::vector<size_t> pos(count.begin(), count.end() - 1);
			for (size_t k = 0; k < t.size(); ++k)
				order[pos[t[k].col]++] = uint32_t(k);
		}
		m.inner.reserve(t.size() / 2);
		m.val.reserve(t.size() / 2);
		for (int c = 0; c < cols; ++c)
		{
			auto b = order.begin() + count[c], e = order.begin() + count[size_t(c) + 1];
			std::stable_sort(b, e, [&](uint32_t a, uint32_t bb) { return t[a].row < t[bb].row; });
			int last = -1;
			for (auto it = b; it != e; ++it)
			{
				const Triplet &tr = t[*it];
				if (tr.row == last)
					m.val.back() += tr.value;
				else
				{
					m.inner.push_back(tr.row);
					m.val.push_back(tr.value);
					last = tr.row;
				}
			}
			m.outer[size_t(c) + 1] = int(m.inner.size());
		}
		return m;
	}

	void add_in_place(SparseCSC &a, const SparseCSC &b)
	{
		if (b.nnz() == 0)
			return;
		if (a.nnz() == 0)
		{
			const int r = a.rows, c = a.cols;
			a = b;
			a.rows = r;
			a.cols = c;
			return;
		}
		SparseCSC out;
		out.resize(a.rows, a.cols);
		out.inner.reserve(std::max(a.nnz(), b.nnz()));
		out.val.reserve(std::max(a.nnz(), b.nnz()));
		for (int c = 0; c < a.cols; ++c)
		{
			int ia = a.outer[c], ea = a.outer[size_t(c) + 1], ib = b.outer[c], eb = b.outer[size_t(c) + 1];
			while (ia < ea || ib < eb)
			{
				if (ib >= eb || (ia < ea && a.inner[ia] < b.inner[ib]))
				{
					out.inner.push_back(a.inner[ia]);
					out.val.push_back(a.val[ia++]);
				}
				else if (ia >= ea || b.inner[ib] < a.inner[ia])
				{
					out.inner.push_back(b.inner[ib]);
					out.val.push_back(b.val[ib++]);
				}
				else
				{
					out.inner.push_back(a.inner[ia]);
					out.val.push_back(a.val[ia++] + b.val[ib++]);
				}
			}
			out.outer[size_t(c) + 1] = int(out.inner.size());
		}
		a = std::move(out);
	}

	// =======================================================================================
	// utils/MatrixCache.{hpp,cpp} — SparseMatrixCache
	// =======================================================================================
	class SparseMatrixCache
	{
	public:
		SparseMatrixCache() {}
		explicit SparseMatrixCache(size_t size) { init(size); }
		// SparseMatrixCache(const SparseMatrixCache &other, bool copy_main_cache_ptr) (MatrixCache.cpp:23-27)
		SparseMatrixCache(const SparseMatrixCache &other, bool copy_main_cache_ptr) { init(other, copy_main_cache_ptr); }

		// MatrixCache.cpp:29-37
		void init(size_t size)
		{
			assert(mapping().empty() || size_ == size);
			size_ = size;
			tmp_.resize(int(size_), int(size_));
			mat_.resize(int(size_), int(size_));
		}
		// MatrixCache.cpp:39-47
		void init(size_t rows, size_t cols)
		{
			size_ = rows == cols ? rows : 0;
			tmp_.resize(int(rows), int(cols));
			mat_.resize(int(rows), int(cols));
		}
		// MatrixCache.cpp:56-78
		void init(const SparseMatrixCache &other, bool copy_main_cache_ptr = false)
		{
			if (copy_main_cache_ptr)
				main_cache_ = other.main_cache_;
			else if (main_cache_ == nullptr)
				main_cache_ = other.main_cache();
			size_ = other.size_;
			values_.assign(other.values_.size(), 0.0);
			tmp_.resize(other.mat_.rows, other.mat_.cols);
			mat_.resize(other.mat_.rows, other.mat_.cols);
		}
		// MatrixCache.hpp:53-57
		std::unique_ptr<SparseMatrixCache> copy() const { return std::make_unique<SparseMatrixCache>(*this, true); }

		// MatrixCache.cpp:80-86
		void set_zero()
		{
			tmp_.resize(tmp_.rows, tmp_.cols);
			mat_.resize(mat_.rows, mat_.cols);
			std::fill(values_.begin(), values_.end(), 0.0);
		}

		void reserve(size_t n) { entries_.reserve(n); }
		size_t entries_size() const { return entries_.size(); }
		size_t triplet_count() const { return entries_.size() + mat_.nnz(); }

		// MatrixCache.cpp:88-113
		void add_value(int e, int i, int j, double value)
		{
			if (mapping().empty())
			{
				entries_.push_back({i, j, value});
				if (int(second_cache_entries_.size()) <= e)
					second_cache_entries_.resize(size_t(e) + 1);
				second_cache_entries_[e].emplace_back(i, j);
			}
			else
			{
				if (e != current_e_)
				{
					current_e_ = e;
					current_e_index_ = 0;
				}
				values_[second_cache()[e][current_e_index_]] += value;
				current_e_index_++;
			}
		}

		// MatrixCache.cpp:115-132
		void prune()
		{
			if (mapping().empty())
			{
				tmp_ = from_triplets(tmp_.rows, tmp_.cols, entries_);
				add_in_place(mat_, tmp_);
				tmp_.resize(tmp_.rows, tmp_.cols);
				entries_.clear();
			}
		}

		// MatrixCache.cpp:134-228 ; the returned matrix is left in mat_
		const SparseCSC &get_matrix(bool compute_mapping = true)
		{
			prune();
			if (mapping().empty())
			{
				if (compute_mapping && size_ > 0)
				{
					assert(main_cache_ == nullptr);
					values_.resize(mat_.nnz());
					inner_index_ = mat_.inner;
					outer_index_ = mat_.outer;
					mapping_.assign(size_t(mat_.rows), {});
					size_t index = 0;
					for (int i = 0; i < mat_.cols; ++i)
						for (int ii = outer_index_[i]; ii < outer_index_[size_t(i) + 1]; ++ii)
						{
							const int j = inner_index_[ii];
							mapping_[j].emplace_back(i, index);
							++index;
						}
					second_cache_.clear();
					second_cache_.resize(second_cache_entries_.size());
					for (size_t e = 0; e < second_cache_entries_.size(); ++e)
						for (const auto &p : second_cache_entries_[e])
						{
							const int i = p.first, j = p.second;
							const auto &map = mapping_[i];
							int index2 = -1;
							for (const auto &q : map)
								if (q.first == j)
								{
									index2 = int(q.second);
									break;
								}
							assert(index2 >= 0);
							second_cache_[e].emplace_back(index2);
						}
					second_cache_entries_.clear();
				}
			}
			else
			{
				const auto &outer_index = main_cache()->outer_index_;
				const auto &inner_index = main_cache()->inner_index_;
				mat_.rows = mat_.cols = int(size_);
				mat_.outer = outer_index;
				mat_.inner = inner_index;
				mat_.val = values_;
				current_e_ = -1;
				current_e_index_ = -1;
			}
			std::fill(values_.begin(), values_.end(), 0.0);
			return mat_;
		}

		// MatrixCache.cpp:289-322
		void operator+=(const SparseMatrixCache &o)
		{
			if (mapping().empty() || o.mapping().empty())
			{
				add_in_place(mat_, o.mat_);
				const size_t this_e_size = second_cache_entries_.size();
				const size_t o_e_size = o.second_cache_entries_.size();
				second_cache_entries_.resize(std::max(this_e_size, o_e_size));
				for (size_t e = 0; e < o_e_size; ++e)
				{
					assert(second_cache_entries_[e].empty() || o.second_cache_entries_[e].empty());
					second_cache_entries_[e].insert(second_cache_entries_[e].end(), o.second_cache_entries_[e].begin(), o.second_cache_entries_[e].end());
				}
			}
			else
			{
				assert(values_.size() == o.values_.size());
				for (size_t i = 0; i < o.values_.size(); ++i)
					values_[i] += o.values_[i];
			}
		}

		const SparseCSC &mat() const { return mat_; }
		const std::vector<Triplet> &entries() const { return entries_; }
		bool has_mapping() const { return !mapping().empty(); }

	private:
		size_t size_ = 0;
		SparseCSC tmp_, mat_;
		std::vector<Triplet> entries_;
		std::vector<std::vector<std::pair<int, size_t>>> mapping_;
		std::vector<int> inner_index_, outer_index_;
		std::vector<double> values_;
		const SparseMatrixCache *main_cache_ = nullptr;
		std::vector<std::vector<int>> second_cache_;
		std::vector<std::vector<std::pair<int, int>>> second_cache_entries_;
		int current_e_ = -1, current_e_index_ = -1;

		const SparseMatrixCache *main_cache() const { return main_cache_ == nullptr ? this : main_cache_; }
		const std::vector<std::vector<std::pair<int, size_t>>> &mapping() const { return main_cache()->mapping_; }
		const std::vector<std::vector<int>> &second_cache() const { return main_cache()->second_cache_; }
	};

	// assembler/Assembler.cpp:18-94 — per-thread storages
	struct LocalThreadMatStorage
	{
		std::unique_ptr<SparseMatrixCache> cache;
		ElementAssemblyValues vals;
		std::vector<double> da;
	};

	// =======================================================================================
	// FixedCorotational (assembler/FixedCorotational.cpp) on the signed SVD of utils/svd.hpp
	// =======================================================================================

	// utils/svd.hpp:270-317 fastSVD3d as AutoFlipSVD<Matrix3d> presents it: eigen-decomposition of A^T A (here: cyclic Jacobi,
	// eigenvalues in decreasing order), sigma = sqrt(max(lambda, 0)) with sigma_2 negated when det A < 0, U from A V:
	// u_0 = A v_0 normalised, u_1 = the part of A v_1 orthogonal to u_0 normalised, u_2 = u_0 x u_1; V a rotation.
	// Matrices row-major [r*3+c]; columns of U / V are the singular vectors.
	void svd3_signed(const double *A, double *U, double *sig, double *V)
	{
		std::vector<double> C(9), Vv, w;
		for (int r = 0; r < 3; ++r)
			for (int c = 0; c < 3; ++c)
				C[r * 3 + c] = A[0 + r] * A[0 + c] + A[3 + r] * A[3 + c] + A[6 + r] * A[6 + c];
		jacobi_eigen(3, C, Vv, w);
		int order[3] = {0, 1, 2};
		std::sort(order, order + 3, [&](int a, int b) { return w[a] > w[b]; });
		for (int k = 0; k < 3; ++k)
		{
			sig[k] = std::sqrt(std::max(w[order[k]], 0.0));
			for (int r = 0; r < 3; ++r)
				V[r * 3 + k] = Vv[size_t(r) * 3 + order[k]];
		}
		if (det3(V) < 0) // keep V a rotation
			for (int r = 0; r < 3; ++r)
				V[r * 3 + 2] = -V[r * 3 + 2];
		if (det3(A) < 0)
			sig[2] = -sig[2];
		double u0[3], u1[3], av1[3];
		for (int r = 0; r < 3; ++r)
		{
			u0[r] = A[r * 3 + 0] * V[0] + A[r * 3 + 1] * V[3] + A[r * 3 + 2] * V[6];
			av1[r] = A[r * 3 + 0] * V[1] + A[r * 3 + 1] * V[4] + A[r * 3 + 2] * V[7];
		}
		double n0 = std::sqrt(u0[0] * u0[0] + u0[1] * u0[1] + u0[2] * u0[2]);
		if (n0 != 0)
			for (int r = 0; r < 3; ++r)
				u0[r] /= n0;
		else
			u0[0] = 1, u0[1] = u0[2] = 0;
		const double d01 = av1[0] * u0[0] + av1[1] * u0[1] + av1[2] * u0[2];
		for (int r = 0; r < 3; ++r)
			u1[r] = av1[r] - d01 * u0[r];
		double n1 = std::sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
		if (n1 != 0)
			for (int r = 0; r < 3; ++r)
				u1[r] /= n1;
		else
		{
			// any unit vector orthogonal to u0
			const int k = std::fabs(u0[0]) < std::fabs(u0[1]) ? (std::fabs(u0[0]) < std::fabs(u0[2]) ? 0 : 2) : (std::fabs(u0[1]) < std::fabs(u0[2]) ? 1 : 2);
			double e[3] = {0, 0, 0};
			e[k] = 1;
			const double d = u0[k];
			for (int r = 0; r < 3; ++r)
				u1[r] = e[r] - d * u0[r];
			n1 = std::sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
			for (int r = 0; r < 3; ++r)
				u1[r] /= n1;
		}
		const double u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
		for (int r = 0; r < 3; ++r)
		{
			U[r * 3 + 0] = u0[r];
			U[r * 3 + 1] = u1[r];
			U[r * 3 + 2] = u2[r];
		}
	}

	void fc_def_grad(const NLData &data, const std::vector<double> &u, int p, double *F)
	{
		for (int k = 0; k < 9; ++k)
			F[k] = 0.0;
		for (int i = 0; i < data.vals.n_loc; ++i)
		{
			const double *gt = data.vals.gt(i, p);
			for (int a = 0; a < 3; ++a)
				for (int d = 0; d < 3; ++d)
					F[a * 3 + d] += u[size_t(i) * 3 + a] * gt[d];
		}
		F[0] += 1.0;
		F[4] += 1.0;
		F[8] += 1.0;
	}

	// FixedCorotational.cpp:601-628 compute_stress_from_singular_values
	void fc_dE_dsigma(const double *s, double lambda, double mu, double *dE)
	{
		const double pl = lambda * (s[0] * s[1] * s[2] - 1.0);
		const double other[3] = {s[1] * s[2], s[2] * s[0], s[0] * s[1]};
		for (int k = 0; k < 3; ++k)
			dE[k] = 2 * mu * (s[k] - 1.0) + other[k] * pl;
	}

	// FixedCorotational.cpp:293-319, 592-599, 671-676 compute_energy
	double fixed_corotational_energy(const NLData &data)
	{
		std::vector<double> u;
		gather_local_disp(data, u);
		double energy = 0.0;
		for (int p = 0; p < data.vals.n_qp; ++p)
		{
			double F[9], U[9], s[3], V[9];
			fc_def_grad(data, u, p, F);
			svd3_signed(F, U, s, V);
			const double sq = (s[0] - 1) * (s[0] - 1) + (s[1] - 1) * (s[1] - 1) + (s[2] - 1) * (s[2] - 1);
			const double pm1 = s[0] * s[1] * s[2] - 1.0;
			energy += (data.mu * sq + data.lambda / 2.0 * pm1 * pm1) * data.da[p];
		}
		return energy;
	}

	// FixedCorotational.cpp:321-377, 678-706: stress = lambda (prod sigma - 1) dJ/dF + 2 mu (F - U V^T), G += delF_delU stress^T da
	void fixed_corotational_gradient(const NLData &data, std::vector<double> &g)
	{
		const int n_loc = data.vals.n_loc;
		g.assign(size_t(n_loc) * 3, 0.0);
		std::vector<double> u;
		gather_local_disp(data, u);
		for (int p = 0; p < data.vals.n_qp; ++p)
		{
			double F[9], U[9], s[3], V[9], P[9];
			fc_def_grad(data, u, p, F);
			svd3_signed(F, U, s, V);
			// dJ/dF: columns are the cross products of the other two columns of F
			double cof[9];
			for (int c = 0; c < 3; ++c)
			{
				const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
				cof[0 * 3 + c] = F[1 * 3 + c1] * F[2 * 3 + c2] - F[2 * 3 + c1] * F[1 * 3 + c2];
				cof[1 * 3 + c] = F[2 * 3 + c1] * F[0 * 3 + c2] - F[0 * 3 + c1] * F[2 * 3 + c2];
				cof[2 * 3 + c] = F[0 * 3 + c1] * F[1 * 3 + c2] - F[1 * 3 + c1] * F[0 * 3 + c2];
			}
			const double pm1 = s[0] * s[1] * s[2] - 1.0;
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
				{
					const double R = U[r * 3 + 0] * V[c * 3 + 0] + U[r * 3 + 1] * V[c * 3 + 1] + U[r * 3 + 2] * V[c * 3 + 2];
					P[r * 3 + c] = data.lambda * pm1 * cof[r * 3 + c] + data.mu * 2 * (F[r * 3 + c] - R);
				}
			for (int i = 0; i < n_loc; ++i)
			{
				const double *gt = data.vals.gt(i, p);
				for (int a = 0; a < 3; ++a)
					g[size_t(i) * 3 + a] += (gt[0] * P[a * 3 + 0] + gt[1] * P[a * 3 + 1] + gt[2] * P[a * 3 + 2]) * data.da[p];
			}
		}
	}

	// FixedCorotational.cpp:630-669, 708-827 compute_stiffness_from_def_grad: d2psi / dF(i,j) dF(r,s) =
	//   sum_kl A_kl U_ik V_jk U_rl V_sl  (A = d2E / dsigma2)
	// + sum over the pairs {k,l} of (L + R) (U_ik V_jl U_rk V_sl + U_il V_jk U_rl V_sk) + (L - R) (U_ik V_jl U_rl V_sk + U_il V_jk U_rk V_sl)
	// with L = mu - lambda/2 (prod sigma - 1) sigma_m (m the third index), R = (dE_k + dE_l) / (2 max(sigma_k + sigma_l, 1e-12)).
	// T[(i*3+j)*9 + r*3+s].
	void fc_stiffness(const double *F, double lambda, double mu, double *T)
	{
		double U[9], s[3], V[9], dE[3], A[9];
		svd3_signed(F, U, s, V);
		fc_dE_dsigma(s, lambda, mu, dE);
		const double prod = s[0] * s[1] * s[2];
		const double other[3] = {s[1] * s[2], s[2] * s[0], s[0] * s[1]};
		for (int k = 0; k < 3; ++k)
			for (int l = 0; l < 3; ++l)
			{
				if (k == l)
					A[k * 3 + l] = 2 * mu + lambda * other[k] * other[k];
				else
					A[k * 3 + l] = lambda * (s[3 - k - l] * (prod - 1.0) + other[k] * other[l]);
			}
		for (int k = 0; k < 81; ++k)
			T[k] = 0.0;
		for (int i = 0; i < 3; ++i)
			for (int j = 0; j < 3; ++j)
				for (int r = 0; r < 3; ++r)
					for (int q = 0; q < 3; ++q)
					{
						double sum = 0.0;
						for (int k = 0; k < 3; ++k)
							for (int l = 0; l < 3; ++l)
								sum += A[k * 3 + l] * U[i * 3 + k] * V[j * 3 + k] * U[r * 3 + l] * V[q * 3 + l];
						for (int k = 0; k < 3; ++k)
						{
							const int l = (k + 1) % 3, m = 3 - k - l;
							const double left = mu - lambda / 2.0 * (prod - 1.0) * s[m];
							const double right = (dE[k] + dE[l]) / (2.0 * std::max(s[k] + s[l], 1.0e-12));
							sum += (left + right) * (U[i * 3 + k] * V[j * 3 + l] * U[r * 3 + k] * V[q * 3 + l] + U[i * 3 + l] * V[j * 3 + k] * U[r * 3 + l] * V[q * 3 + k]);
							sum += (left - right) * (U[i * 3 + k] * V[j * 3 + l] * U[r * 3 + l] * V[q * 3 + k] + U[i * 3 + l] * V[j * 3 + k] * U[r * 3 + k] * V[q * 3 + l]);
						}
						T[(i * 3 + j) * 9 + r * 3 + q] = sum;
					}
	}

	// FixedCorotational.cpp:379-436: H += B^T stiffness B da, B(F(j,k), 3 i + j) = delF_delU(i, k)
	void fixed_corotational_hessian(const NLData &data, std::vector<double> &h)
	{
		const int n_loc = data.vals.n_loc, N = 3 * n_loc;
		h.assign(size_t(N) * N, 0.0);
		std::vector<double> u;
		gather_local_disp(data, u);
		for (int p = 0; p < data.vals.n_qp; ++p)
		{
			double F[9], T[81];
			fc_def_grad(data, u, p, F);
			fc_stiffness(F, data.lambda, data.mu, T);
			for (int i = 0; i < n_loc; ++i)
				for (int j = 0; j < n_loc; ++j)
				{
					const double *gi = data.vals.gt(i, p), *gj = data.vals.gt(j, p);
					for (int a = 0; a < 3; ++a)
						for (int b = 0; b < 3; ++b)
						{
							double sum = 0.0;
							for (int d = 0; d < 3; ++d)
								for (int d2 = 0; d2 < 3; ++d2)
									sum += gi[d] * T[(a * 3 + d) * 9 + b * 3 + d2] * gj[d2];
							h[size_t(i * 3 + a) * N + j * 3 + b] += sum * data.da[p];
						}
				}
		}
	}

	void compute_da(const ElementAssemblyValues &vals, std::vector<double> &da)
	{
		da.resize(vals.n_qp);
		for (int q = 0; q < vals.n_qp; ++q)
			da[q] = vals.det[q] * vals.weights[q]; // Assembler.cpp:206, 518, 686
	}

	// local dispatch by material (the virtual calls of NLAssembler)
	double local_energy(const Problem &pb, const NLData &data)
	{
		if (pb.d.material == ORACLE_NEOHOOKEAN)
			return neohookean_energy(data);
		if (pb.d.material == ORACLE_SAINT_VENANT)
			return saint_venant_energy<double>(data);
		if (pb.d.material == ORACLE_MOONEY_RIVLIN)
			return mooney_rivlin_energy<double>(data);
		if (pb.d.material == ORACLE_VISCOUS_DAMPING)
			return viscous_damping_energy(data);
		if (pb.d.material == ORACLE_FIXED_COROTATIONAL)
			return fixed_corotational_energy(data);
		return linear_elasticity_energy<double>(data);
	}
	void local_gradient(const Problem &pb, const NLData &data, std::vector<double> &g)
	{
		if (pb.d.material == ORACLE_NEOHOOKEAN)
		{
			neohookean_gradient(data, g);
			return;
		}
		if (pb.d.material == ORACLE_VISCOUS_DAMPING)
		{
			viscous_damping_gradient(data, g);
			return;
		}
		if (pb.d.material == ORACLE_FIXED_COROTATIONAL)
		{
			fixed_corotational_gradient(data, g);
			return;
		}
		// utils/ElasticityUtils.cpp:81-... gradient_from_energy: autodiff gradient
		const D1 e = pb.d.material == ORACLE_SAINT_VENANT   ? saint_venant_energy<D1>(data)
					 : pb.d.material == ORACLE_MOONEY_RIVLIN ? mooney_rivlin_energy<D1>(data)
															 : linear_elasticity_energy<D1>(data);
		g = e.g;
	}
	void local_hessian(const Problem &pb, const NLData &data, std::vector<double> &h)
	{
		if (pb.d.material == ORACLE_NEOHOOKEAN)
		{
			neohookean_hessian(data, h);
			return;
		}
		if (pb.d.material == ORACLE_VISCOUS_DAMPING)
		{
			viscous_damping_hessian(data, h);
			return;
		}
		if (pb.d.material == ORACLE_FIXED_COROTATIONAL)
		{
			fixed_corotational_hessian(data, h);
			return;
		}
		const D2 e = pb.d.material == ORACLE_SAINT_VENANT   ? saint_venant_energy<D2>(data)
					 : pb.d.material == ORACLE_MOONEY_RIVLIN ? mooney_rivlin_energy<D2>(data)
															 : linear_elasticity_energy<D2>(data);
		h = e.h;
	}
} // namespace

struct oracle_problem
{
	Problem pb;
	SparseMatrixCache mat_cache; // the caller-owned utils::MatrixCache of ElasticForm (ElasticForm.hpp:116)
	bool mat_cache_inited = false;
};

struct oracle_cache
{
	SparseMatrixCache c;
	SparseCSC last;
	std::vector<int32_t> outer, inner;
};

extern "C"
{
	void oracle_set_previous(oracle_problem *op, const double *x_prev, double dt)
	{
		Problem &pb = op->pb;
		pb.x_prev.clear();
		if (x_prev)
			pb.x_prev.assign(x_prev, x_prev + size_t(pb.d.n_bases) * pb.size);
		pb.dt = dt;
	}

	oracle_problem *oracle_create(const oracle_desc *desc)
	{
		auto *op = new oracle_problem();
		Problem &pb = op->pb;
		pb.d = *desc;
		pb.size = desc->material == ORACLE_LAPLACIAN ? 1 : 3;
		const size_t ne = desc->n_elements, nl = desc->n_loc, nq = desc->n_qp;
		pb.conn.assign(desc->conn, desc->conn + ne * nl);
		pb.vertices.assign(desc->vertices, desc->vertices + ne * 12);
		pb.qw.assign(desc->quad_weights, desc->quad_weights + nq);
		if (desc->quad_points)
			pb.qpts.assign(desc->quad_points, desc->quad_points + nq * 3);
		pb.ref_grads.assign(desc->ref_grads, desc->ref_grads + nq * nl * 3);
		if (desc->node_lattice)
			pb.lattice.assign(desc->node_lattice, desc->node_lattice + nl * 3);
		if (desc->lambda)
			pb.lambda.assign(desc->lambda, desc->lambda + ne);
		if (desc->mu)
			pb.mu.assign(desc->mu, desc->mu + ne);
		if (desc->material == ORACLE_MOONEY_RIVLIN && desc->param3)
			pb.param3.assign(desc->param3, desc->param3 + ne);
		if (desc->ref_vals)
			pb.ref_vals.assign(desc->ref_vals, desc->ref_vals + nq * nl);
		if (desc->density)
			pb.density.assign(desc->density, desc->density + ne);
		if (desc->geom_order > 1 && desc->geom_nodes && desc->geom_lattice)
		{
			pb.geom_nodes.assign(desc->geom_nodes, desc->geom_nodes + ne * size_t(desc->n_geom_loc) * 3);
			pb.geom_lattice.assign(desc->geom_lattice, desc->geom_lattice + size_t(desc->n_geom_loc) * 3);
		}
		if (pb.density.empty())
			pb.density.assign(ne, 1.0);
		if (pb.lambda.empty())
			pb.lambda.assign(ne, 0.0);
		if (pb.mu.empty())
			pb.mu.assign(ne, 0.0);
		if (pb.param3.empty())
			pb.param3.assign(ne, 0.0);
		pb.d.conn = nullptr;
		if (desc->use_cache)
		{
			// AssemblyValsCache::init (AssemblyValsCache.cpp:11-32)
			pb.cache.resize(ne);
			maybe_parallel_for(int(ne), desc->n_threads, [&](int s, int e, int) {
				for (int k = s; k < e; ++k)
					compute_assembly_values(pb, k, pb.cache[k]);
			});
		}
		return op;
	}

	void oracle_destroy(oracle_problem *p) { delete p; }
	int oracle_size(const oracle_problem *p) { return p->pb.size; }

	// NLAssembler::assemble_energy (assembler/Assembler.cpp:495-531)
	double oracle_assemble_energy(oracle_problem *op, const double *x)
	{
		Problem &pb = op->pb;
		const int nt = std::max(1, pb.d.n_threads);
		std::vector<double> partial(nt, 0.0);
		maybe_parallel_for(pb.d.n_elements, nt, [&](int start, int end, int tid) {
			ElementAssemblyValues vals;
			std::vector<double> da;
			double local = 0.0;
			for (int e = start; e < end; ++e)
			{
				pb.cache_compute(e, vals);
				compute_da(vals, da);
				local += local_energy(pb, NLData{vals, x, da, pb.lambda[e], pb.mu[e], pb.param3[e], pb.x_prev.empty() ? nullptr : pb.x_prev.data(), pb.dt});
			}
			partial[tid] = local;
		});
		double res = 0;
		for (double v : partial)
			res += v; // serial merge (:526-530)
		return res;
	}

	// NLAssembler::assemble_energy_per_element (assembler/Assembler.cpp:533-572)
	void oracle_assemble_energy_per_element(oracle_problem *op, const double *x, double *out)
	{
		Problem &pb = op->pb;
		maybe_parallel_for(pb.d.n_elements, pb.d.n_threads, [&](int start, int end, int) {
			ElementAssemblyValues vals;
			std::vector<double> da;
			for (int e = start; e < end; ++e)
			{
				pb.cache_compute(e, vals);
				compute_da(vals, da);
				out[e] = local_energy(pb, NLData{vals, x, da, pb.lambda[e], pb.mu[e], pb.param3[e], pb.x_prev.empty() ? nullptr : pb.x_prev.data(), pb.dt});
			}
		});
	}

	// NLAssembler::assemble_gradient (assembler/Assembler.cpp:574-643)
	void oracle_assemble_gradient(oracle_problem *op, const double *x, double *rhs)
	{
		Problem &pb = op->pb;
		const int size = pb.size;
		const size_t ndof = size_t(pb.d.n_bases) * size;
		const int nt = std::max(1, std::min(pb.d.n_threads, pb.d.n_elements));
		std::vector<std::vector<double>> vecs(nt, std::vector<double>(ndof, 0.0)); // LocalThreadVecStorage
		maybe_parallel_for(pb.d.n_elements, nt, [&](int start, int end, int tid) {
			ElementAssemblyValues vals;
			std::vector<double> da, val;
			std::vector<double> &vec = vecs[tid];
			for (int e = start; e < end; ++e)
			{
				pb.cache_compute(e, vals);
				compute_da(vals, da);
				local_gradient(pb, NLData{vals, x, da, pb.lambda[e], pb.mu[e], pb.param3[e], pb.x_prev.empty() ? nullptr : pb.x_prev.data(), pb.dt}, val);
				for (int j = 0; j < vals.n_loc; ++j)
					for (int m = 0; m < size; ++m)
						vec[size_t(vals.global[j]) * size + m] += val[size_t(j) * size + m] * 1.0; // (:616-629)
			}
		});
		std::fill(rhs, rhs + ndof, 0.0);
		for (const auto &v : vecs) // serial merge (:641-642)
			for (size_t i = 0; i < ndof; ++i)
				rhs[i] += v[i];
	}

	static void publish(Problem &pb, const SparseCSC &m)
	{
		pb.outer.assign(m.outer.begin(), m.outer.end());
		pb.inner.assign(m.inner.begin(), m.inner.end());
		pb.values = m.val;
	}

	// NLAssembler::assemble_hessian (assembler/Assembler.cpp:645-771)
	int64_t oracle_assemble_hessian(oracle_problem *op, const double *x, int psd)
	{
		Problem &pb = op->pb;
		const int size = pb.size;
		const long max_triplets_size = long(1e7);
		const long buffer_size = std::min(max_triplets_size, long(pb.d.n_bases) * size);

		op->mat_cache.init(size_t(pb.d.n_bases) * size); // (:666)
		op->mat_cache.set_zero();                        // (:667)

		const int nt = std::max(1, std::min(pb.d.n_threads, pb.d.n_elements));
		std::vector<LocalThreadMatStorage> storage(nt); // create_thread_storage (:669)
		for (auto &s : storage)
		{
			s.cache = op->mat_cache.copy();
			s.cache->reserve(size_t(buffer_size));
			s.cache->init(op->mat_cache);
		}

		const double t0 = now_seconds();
		maybe_parallel_for(pb.d.n_elements, nt, [&](int start, int end, int tid) {
			LocalThreadMatStorage &ls = storage[tid];
			std::vector<double> H;
			for (int e = start; e < end; ++e)
			{
				ElementAssemblyValues &vals = ls.vals;
				pb.cache_compute(e, vals);
				compute_da(vals, ls.da);
				const int n_loc = vals.n_loc, N = n_loc * size;
				local_hessian(pb, NLData{vals, x, ls.da, pb.lambda[e], pb.mu[e], pb.param3[e], pb.x_prev.empty() ? nullptr : pb.x_prev.data(), pb.dt}, H);
				if (psd)
					project_to_psd(N, H); // (:693-694)
				for (int i = 0; i < n_loc; ++i)
					for (int j = 0; j < n_loc; ++j)
						for (int n = 0; n < size; ++n)
							for (int m = 0; m < size; ++m)
							{
								const double local_value = H[size_t(i * size + m) * N + (j * size + n)];
								const int gi = vals.global[i] * size + m;
								const int gj = vals.global[j] * size + n;
								ls.cache->add_value(e, gi, gj, local_value * 1.0 * 1.0); // (:737)
								if (long(ls.cache->entries_size()) >= max_triplets_size)
									ls.cache->prune(); // (:742-746)
							}
			}
		});
		const double t1 = now_seconds();
		for (auto &ls : storage) // serial merge (:762-766)
		{
			ls.cache->prune();
			op->mat_cache += *ls.cache;
		}
		const SparseCSC &hess = op->mat_cache.get_matrix(); // (:767)
		publish(pb, hess);
		const double t2 = now_seconds();
		pb.loop_seconds = t1 - t0;
		pb.merge_seconds = t2 - t1;
		return int64_t(pb.values.size());
	}

	// LinearAssembler::assemble (assembler/Assembler.cpp:157-384)
	int64_t oracle_assemble_linear(oracle_problem *op)
	{
		Problem &pb = op->pb;
		const int size = pb.size;
		const int rows = pb.d.n_bases * size;
		const long max_triplets_size = long(1e7);
		const long buffer_size = std::min(max_triplets_size, long(pb.d.n_bases) * size);
		const int nt = std::max(1, std::min(pb.d.n_threads, pb.d.n_elements));
		std::vector<LocalThreadMatStorage> storage(nt);
		for (auto &s : storage)
		{
			s.cache = std::make_unique<SparseMatrixCache>();
			s.cache->reserve(size_t(buffer_size));
			s.cache->init(size_t(rows), size_t(rows));
		}
		const double t0 = now_seconds();
		maybe_parallel_for(pb.d.n_elements, nt, [&](int start, int end, int tid) {
			LocalThreadMatStorage &ls = storage[tid];
			for (int e = start; e < end; ++e)
			{
				ElementAssemblyValues &vals = ls.vals;
				pb.cache_compute(e, vals);
				compute_da(vals, ls.da);
				const int n_loc = vals.n_loc;
				for (int i = 0; i < n_loc; ++i)
					for (int j = 0; j <= i; ++j) // symmetry (:217)
					{
						double blk[9];
						if (pb.d.material == ORACLE_LAPLACIAN)
							blk[0] = laplacian_local(vals, ls.da, i, j);
						else if (pb.d.material == ORACLE_MASS)
							mass_local(vals, ls.da, i, j, pb.density[e], size, blk);
						else
							linear_elasticity_local(vals, ls.da, i, j, pb.lambda[e], pb.mu[e], blk);
						for (int n = 0; n < size; ++n)
							for (int m = 0; m < size; ++m)
							{
								const double local_value = blk[n * size + m];
								const int gi = vals.global[i] * size + m;
								const int gj = vals.global[j] * size + n;
								ls.cache->add_value(e, gi, gj, local_value * 1.0 * 1.0);
								if (j < i)
									ls.cache->add_value(e, gj, gi, local_value * 1.0 * 1.0);
								if (long(ls.cache->entries_size()) >= max_triplets_size)
									ls.cache->prune();
							}
					}
			}
		});
		const double t1 = now_seconds();
		// prune, concatenate all thread triplets, one global setFromTriplets (:286-371)
		std::vector<Triplet> triplets;
		size_t total = 0;
		for (auto &ls : storage)
		{
			ls.cache->prune();
			total += ls.cache->triplet_count();
		}
		triplets.reserve(total);
		for (auto &ls : storage)
		{
			const auto &ent = ls.cache->entries();
			triplets.insert(triplets.end(), ent.begin(), ent.end());
			const SparseCSC &m = ls.cache->mat();
			for (int c = 0; c < m.cols; ++c)
				for (int k = m.outer[c]; k < m.outer[size_t(c) + 1]; ++k)
					triplets.push_back({m.inner[k], c, m.val[k]});
		}
		const SparseCSC stiffness = from_triplets(rows, rows, triplets);
		publish(pb, stiffness);
		const double t2 = now_seconds();
		pb.loop_seconds = t1 - t0;
		pb.merge_seconds = t2 - t1;
		return int64_t(pb.values.size());
	}

	int64_t oracle_csc_nnz(const oracle_problem *p) { return int64_t(p->pb.values.size()); }
	const int32_t *oracle_csc_outer(const oracle_problem *p) { return p->pb.outer.data(); }
	const int32_t *oracle_csc_inner(const oracle_problem *p) { return p->pb.inner.data(); }
	const double *oracle_csc_values(const oracle_problem *p) { return p->pb.values.data(); }
	double oracle_last_loop_seconds(const oracle_problem *p) { return p->pb.loop_seconds; }
	double oracle_last_merge_seconds(const oracle_problem *p) { return p->pb.merge_seconds; }

	// ---- local quantities for unit tests ----
	static NLData make_data(Problem &pb, int e, const double *x, ElementAssemblyValues &vals, std::vector<double> &da)
	{
		pb.cache_compute(e, vals);
		compute_da(vals, da);
		return NLData{vals, x, da, pb.lambda[e], pb.mu[e], pb.param3[e], pb.x_prev.empty() ? nullptr : pb.x_prev.data(), pb.dt};
	}

	double oracle_local_energy(oracle_problem *op, int e, const double *x, int autodiff)
	{
		ElementAssemblyValues vals;
		std::vector<double> da;
		const NLData data = make_data(op->pb, e, x, vals, da);
		if (op->pb.d.material == ORACLE_NEOHOOKEAN)
			return autodiff ? neohookean_energy_autodiff<D1>(data).v : neohookean_energy(data);
		return local_energy(op->pb, data); // the material's own compute_energy
	}

	void oracle_local_gradient(oracle_problem *op, int e, const double *x, int autodiff, double *g)
	{
		ElementAssemblyValues vals;
		std::vector<double> da, out;
		const NLData data = make_data(op->pb, e, x, vals, da);
		if (op->pb.d.material == ORACLE_NEOHOOKEAN && autodiff)
			out = neohookean_energy_autodiff<D1>(data).g;
		else
			local_gradient(op->pb, data, out);
		std::copy(out.begin(), out.end(), g);
	}

	void oracle_local_hessian(oracle_problem *op, int e, const double *x, int autodiff, double *h)
	{
		ElementAssemblyValues vals;
		std::vector<double> da, out;
		const NLData data = make_data(op->pb, e, x, vals, da);
		if (op->pb.d.material == ORACLE_NEOHOOKEAN && autodiff)
			out = neohookean_energy_autodiff<D2>(data).h;
		else
			local_hessian(op->pb, data, out);
		std::copy(out.begin(), out.end(), h);
	}

	void oracle_local_stiffness(oracle_problem *op, int e, int i, int j, double *blk)
	{
		ElementAssemblyValues vals;
		std::vector<double> da;
		op->pb.cache_compute(e, vals);
		compute_da(vals, da);
		if (op->pb.d.material == ORACLE_LAPLACIAN)
			blk[0] = laplacian_local(vals, da, i, j);
		else if (op->pb.d.material == ORACLE_MASS)
			mass_local(vals, da, i, j, op->pb.density[e], op->pb.size, blk);
		else
			linear_elasticity_local(vals, da, i, j, op->pb.lambda[e], op->pb.mu[e], blk);
	}

	void oracle_assembly_values(oracle_problem *op, int e, double *det, double *jac_it, double *grad_t_m)
	{
		ElementAssemblyValues vals;
		op->pb.cache_compute(e, vals);
		for (int q = 0; q < vals.n_qp; ++q)
		{
			det[q] = vals.det[q];
			for (int k = 0; k < 9; ++k)
				jac_it[size_t(q) * 9 + k] = vals.jac_it[size_t(q) * 9 + k];
			for (int j = 0; j < vals.n_loc; ++j)
				for (int c = 0; c < 3; ++c)
					grad_t_m[(size_t(q) * vals.n_loc + j) * 3 + c] = vals.gt(j, q)[c];
		}
	}

	void oracle_project_to_psd(int n, double *a)
	{
		std::vector<double> A(a, a + size_t(n) * n);
		project_to_psd(n, A);
		std::copy(A.begin(), A.end(), a);
	}

	// ---- SparseMatrixCache exposed for the reference's "cache" known-answer test ----
	oracle_cache *oracle_cache_new(int size)
	{
		auto *c = new oracle_cache();
		c->c.init(size_t(size));
		return c;
	}
	oracle_cache *oracle_cache_copy(const oracle_cache *other)
	{
		auto *c = new oracle_cache();
		c->c.init(other->c); // SparseMatrixCache(const MatrixCache &other) -> init(other) (MatrixCache.cpp:18-21,49-54)
		return c;
	}
	void oracle_cache_free(oracle_cache *c) { delete c; }
	void oracle_cache_add_value(oracle_cache *c, int e, int i, int j, double v) { c->c.add_value(e, i, j, v); }
	void oracle_cache_prune(oracle_cache *c) { c->c.prune(); }
	void oracle_cache_set_zero(oracle_cache *c) { c->c.set_zero(); }
	void oracle_cache_add(oracle_cache *dst, const oracle_cache *src) { dst->c += src->c; } // SparseMatrixCache::operator+= (MatrixCache.cpp:289-322)
	int64_t oracle_cache_get_matrix(oracle_cache *c)
	{
		c->last = c->c.get_matrix();
		c->outer.assign(c->last.outer.begin(), c->last.outer.end());
		c->inner.assign(c->last.inner.begin(), c->last.inner.end());
		return int64_t(c->last.val.size());
	}
	const int32_t *oracle_cache_outer(const oracle_cache *c) { return c->outer.data(); }
	const int32_t *oracle_cache_inner(const oracle_cache *c) { return c->inner.data(); }
	const double *oracle_cache_values(const oracle_cache *c) { return c->last.val.data(); }
}
