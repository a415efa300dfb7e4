// Host-side, once-per-mesh construction of the sparsity pattern and the slot map.
//
// Replaces the first-call path of SparseMatrixCache (reference utils/MatrixCache.cpp:88-100
// triplet buffering, :115-132 prune/setFromTriplets, :134-213 mapping + second_cache): the
// pattern is derived from connectivity alone. Because every element contributes the full
// (g_i*size+m, g_j*size+n) block for every local pair (i,j) — explicit zeros included,
// Assembler.cpp:712-737 — the scalar pattern is the node-adjacency pattern expanded by
// size x size, which is what `adj_off/adj` encode (one entry per node pair instead of
// size^2 scalar entries), and the slot map stores one int per local node pair instead of
// the reference's one int per scalar contribution (MatrixCache.hpp:118).
#include "pfa_internal.h"

#include <algorithm>
#include <new>
#include <stdexcept>
#include <thread>

namespace pfa
{
	namespace
	{
		template <typename F>
		void parallel_ranges(int64_t n, F &&body)
		{
			unsigned hw = std::thread::hardware_concurrency();
			int nt = int(std::max(1u, std::min(hw ? hw : 1u, 64u)));
			if (n < 4096)
				nt = 1;
			std::vector<std::thread> pool;
			for (int t = 0; t < nt; ++t)
			{
				const int64_t s = n * t / nt, e = n * (t + 1) / nt;
				pool.emplace_back([=, &body]() { body(t, s, e); });
			}
			for (auto &th : pool)
				th.join();
		}
	} // namespace

	void build_pattern(const int32_t *conn, int n_el, int n_loc, int n_bases, HostPattern &out)
	{
		// node -> incident elements (counting sort)
		std::vector<int64_t> ne_off(size_t(n_bases) + 1, 0);
		for (int64_t k = 0; k < int64_t(n_el) * n_loc; ++k)
		{
			const int32_t g = conn[k];
			if (g < 0 || g >= n_bases)
				throw std::runtime_error("pfa: connectivity index out of range [0, n_bases)");
			++ne_off[size_t(g) + 1];
		}
		for (int b = 0; b < n_bases; ++b)
			ne_off[size_t(b) + 1] += ne_off[b];
		std::vector<int32_t> node_el;
		node_el.resize(size_t(ne_off[size_t(n_bases)]));
		{
			std::vector<int64_t> pos(ne_off.begin(), ne_off.end() - 1);
			for (int e = 0; e < n_el; ++e)
				for (int j = 0; j < n_loc; ++j)
					node_el[size_t(pos[conn[size_t(e) * n_loc + j]]++)] = e;
		}

		// adjacency of every node = sorted union of the nodes of its incident elements
		const int max_threads = 64; // parallel_ranges never uses more
		std::vector<std::vector<int32_t>> part(max_threads);
		std::vector<int32_t> deg(size_t(n_bases), 0);
		parallel_ranges(n_bases, [&](int t, int64_t s, int64_t e) {
			std::vector<int32_t> buf;
			std::vector<int32_t> &mine = part[t];
			for (int64_t b = s; b < e; ++b)
			{
				buf.clear();
				for (int64_t k = ne_off[b]; k < ne_off[b + 1]; ++k)
				{
					const int32_t *c = conn + size_t(node_el[size_t(k)]) * n_loc;
					buf.insert(buf.end(), c, c + n_loc);
				}
				std::sort(buf.begin(), buf.end());
				buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
				deg[size_t(b)] = int32_t(buf.size());
				mine.insert(mine.end(), buf.begin(), buf.end());
			}
		});
		int64_t total = 0;
		for (int b = 0; b < n_bases; ++b)
			total += deg[b];
		if (total >= (int64_t(1) << 31))
			throw std::runtime_error("pfa: more than 2^31 node pairs; int32 StiffnessMatrix indices cannot hold this pattern");
		out.adj_off.resize(size_t(n_bases) + 1);
		out.adj_off[0] = 0;
		for (int b = 0; b < n_bases; ++b)
			out.adj_off[size_t(b) + 1] = out.adj_off[b] + deg[b];
		out.adj.resize(size_t(total));
		{
			// parts are in node order (thread t handled a contiguous node range)
			size_t at = 0;
			for (auto &p : part)
			{
				std::copy(p.begin(), p.end(), out.adj.begin() + at);
				at += p.size();
				std::vector<int32_t>().swap(p);
			}
		}

		// slot map: pair index of (row node g_i) in the column list of node g_j
		const int nl2 = n_loc * n_loc;
		out.slot.resize(size_t(n_el) * nl2);
		parallel_ranges(n_el, [&](int, int64_t s, int64_t e) {
			for (int64_t el = s; el < e; ++el)
			{
				const int32_t *c = conn + size_t(el) * n_loc;
				for (int j = 0; j < n_loc; ++j)
				{
					const int32_t *lb = out.adj.data() + out.adj_off[c[j]];
					const int32_t *le = out.adj.data() + out.adj_off[size_t(c[j]) + 1];
					for (int i = 0; i < n_loc; ++i)
					{
						const int32_t *it = std::lower_bound(lb, le, c[i]);
						out.slot[size_t(el) * nl2 + i * n_loc + j] = int32_t(it - out.adj.data());
					}
				}
			}
		});
	}

	// Internal element order: Morton (Z-curve) order of the element centroids. The scatter of
	// consecutive elements then lands in a window of CSC columns that stays resident in the
	// 126 MB L2 until the neighbouring elements have added their share, instead of being
	// evicted and re-fetched once per mesh layer (caller order: x-slowest sweeps).
	void spatial_element_order(const double *vertices, int n_el, std::vector<int32_t> &perm)
	{
		double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
		for (int64_t k = 0; k < int64_t(n_el) * 4; ++k)
			for (int c = 0; c < 3; ++c)
			{
				const double v = vertices[k * 3 + c];
				if (v < lo[c])
					lo[c] = v;
				if (v > hi[c])
					hi[c] = v;
			}
		double inv[3];
		for (int c = 0; c < 3; ++c)
			inv[c] = hi[c] > lo[c] ? double((1 << 20) - 1) / (hi[c] - lo[c]) : 0.0;
		auto spread = [](uint64_t v) { // 21 bits -> every third bit
			v &= 0x1fffff;
			v = (v | v << 32) & 0x1f00000000ffffull;
			v = (v | v << 16) & 0x1f0000ff0000ffull;
			v = (v | v << 8) & 0x100f00f00f00f00full;
			v = (v | v << 4) & 0x10c30c30c30c30c3ull;
			v = (v | v << 2) & 0x1249249249249249ull;
			return v;
		};
		std::vector<std::pair<uint64_t, int32_t>> keyed{size_t(n_el)};
		parallel_ranges(n_el, [&](int, int64_t s, int64_t e) {
			for (int64_t el = s; el < e; ++el)
			{
				const double *v = vertices + el * 12;
				uint64_t key = 0;
				for (int c = 0; c < 3; ++c)
				{
					const double cen = 0.25 * (v[c] + v[3 + c] + v[6 + c] + v[9 + c]);
					double q = (cen - lo[c]) * inv[c];
					if (!(q >= 0.0)) // NaN coordinates: keep them together at the front
						q = 0.0;
					key |= spread(uint64_t(q)) << c;
				}
				keyed[size_t(el)] = {key, int32_t(el)};
			}
		});
		std::sort(keyed.begin(), keyed.end()); // ties broken by the caller's index: deterministic
		perm.resize(size_t(n_el));
		for (int64_t el = 0; el < n_el; ++el)
			perm[size_t(el)] = keyed[size_t(el)].second;
	}

	// Which warp batch touches a node's column block first decides who clears it: the block of
	// node b (values[size^2 adj_off[b] .. size^2 adj_off[b+1])) goes to the list of batch
	// first(b) = min element index containing b / batch_elements. Nodes no computed element touches
	// (pattern-only ghost connectivity) are spread round-robin. Neighbouring blocks are merged.
	void build_zero_schedule(const int32_t *conn, int n_el, int n_loc, int n_bases, const std::vector<int32_t> &adj_off, int size, int batch_elements,
							 std::vector<int32_t> &zoff, std::vector<int32_t> &zruns)
	{
		const int n_batches = (n_el + batch_elements - 1) / batch_elements;
		std::vector<int32_t> first;
		first.assign(size_t(n_bases), -1);
		for (int e = 0; e < n_el; ++e)
			for (int j = 0; j < n_loc; ++j)
			{
				int32_t &f = first[size_t(conn[size_t(e) * n_loc + j])];
				if (f < 0)
					f = e / batch_elements; // elements are visited in increasing order
			}
		std::vector<int32_t> count;
		count.assign(size_t(n_batches) + 1, 0);
		int rr = 0;
		for (int b = 0; b < n_bases; ++b)
		{
			if (first[b] < 0)
				first[b] = (rr++) % n_batches;
			++count[size_t(first[b]) + 1];
		}
		for (int k = 0; k < n_batches; ++k)
			count[size_t(k) + 1] += count[k];
		std::vector<int32_t> nodes;
		nodes.resize(size_t(n_bases));
		std::vector<int32_t> pos(count.begin(), count.end() - 1);
		for (int b = 0; b < n_bases; ++b) // increasing b inside every list
			nodes[size_t(pos[first[b]]++)] = b;
		zoff.assign(size_t(n_batches) + 1, 0);
		zruns.clear();
		zruns.reserve(size_t(n_bases) * 2);
		const int64_t s2 = int64_t(size) * size;
		for (int k = 0; k < n_batches; ++k)
		{
			int64_t run_start = -1, run_end = -1;
			for (int32_t t = count[k]; t < count[size_t(k) + 1]; ++t)
			{
				const int b = nodes[size_t(t)];
				const int64_t s = s2 * adj_off[size_t(b)], e = s2 * adj_off[size_t(b) + 1];
				if (e == s)
					continue;
				if (s == run_end)
					run_end = e;
				else
				{
					if (run_end > run_start)
					{
						zruns.push_back(int32_t(run_start));
						zruns.push_back(int32_t(run_end - run_start));
					}
					run_start = s;
					run_end = e;
				}
			}
			if (run_end > run_start)
			{
				zruns.push_back(int32_t(run_start));
				zruns.push_back(int32_t(run_end - run_start));
			}
			zoff[size_t(k) + 1] = int32_t(zruns.size() / 2);
		}
	}
} // namespace pfa

// ---- host-only entry points of include/pfa.h: the pattern builder and the element order without a device ----
struct pfa_host_pattern
{
	pfa::HostPattern hp;
};

extern "C"
{
	int pfa_host_pattern_create(int32_t n_elements, int32_t n_loc, int32_t n_bases, const int32_t *conn, pfa_host_pattern **out)
	{
		if (!conn || !out || n_elements <= 0 || n_loc <= 0 || n_bases <= 0)
			return PFA_ERR_INVALID;
		*out = nullptr;
		pfa_host_pattern *p = new (std::nothrow) pfa_host_pattern();
		if (!p)
			return PFA_ERR_NOMEM;
		try
		{
			pfa::build_pattern(conn, n_elements, n_loc, n_bases, p->hp);
		}
		catch (const std::bad_alloc &)
		{
			delete p;
			return PFA_ERR_NOMEM;
		}
		catch (const std::exception &)
		{
			delete p;
			return PFA_ERR_INVALID; // connectivity index out of range, or more than 2^31 node pairs
		}
		*out = p;
		return PFA_OK;
	}

	int pfa_host_pattern_arrays(const pfa_host_pattern *p, int64_t *n_pairs, const int32_t **adj_off, const int32_t **adj, const int32_t **slot)
	{
		if (!p)
			return PFA_ERR_INVALID;
		if (n_pairs)
			*n_pairs = int64_t(p->hp.adj.size());
		if (adj_off)
			*adj_off = p->hp.adj_off.data();
		if (adj)
			*adj = p->hp.adj.data();
		if (slot)
			*slot = p->hp.slot.data();
		return PFA_OK;
	}

	void pfa_host_pattern_destroy(pfa_host_pattern *p) { delete p; }

	int pfa_host_element_order(int32_t n_elements, const double *vertices, int32_t *perm)
	{
		if (!vertices || !perm || n_elements <= 0)
			return PFA_ERR_INVALID;
		try
		{
			std::vector<int32_t> order;
			pfa::spatial_element_order(vertices, n_elements, order);
			std::copy(order.begin(), order.end(), perm);
		}
		catch (const std::bad_alloc &)
		{
			return PFA_ERR_NOMEM;
		}
		return PFA_OK;
	}
}
