// Element partition for the multi-GPU owner-computes path (host code only): pfa_partition_* of include/pfa.h.
// Reference counterpart: the element ranges maybe_parallel_for hands to the threads (utils/MaybeParallelFor.tpp:18-68);
// here a range is what one GPU assembles, and because every rank also gets the ghost elements around its owned nodes,
// the columns it produces are finished without an exchange.
#include "../../include/pfa.h"

#include <algorithm>
#include <cstdint>
#include <memory>
#include <new>
#include <vector>

struct pfa_partition
{
	int32_t n_own = 0, n_ghost = 0, n_local = 0, n_owned = 0;
	std::vector<int32_t> elements, conn, l2g;
	std::vector<uint8_t> owned;
};

extern "C"
{
	int pfa_partition_create(int32_t n_elements, int32_t n_loc, int32_t n_bases, const int32_t *conn, int32_t world, int32_t rank, pfa_partition **out)
	{
		if (!conn || !out || n_elements <= 0 || n_loc <= 0 || n_bases <= 0 || world <= 0 || rank < 0 || rank >= world)
			return PFA_ERR_INVALID;
		*out = nullptr;
		try
		{
			const size_t ne = size_t(n_elements), nl = size_t(n_loc);
			// first element touching each node (its owner is the rank of that element) and incidences per node
			std::vector<int32_t> first(size_t(n_bases), -1), ninc(size_t(n_bases), 0);
			for (size_t e = 0; e < ne; ++e)
				for (size_t j = 0; j < nl; ++j)
				{
					const int32_t g = conn[e * nl + j];
					if (g < 0 || g >= n_bases)
						return PFA_ERR_INVALID;
					if (first[size_t(g)] < 0)
						first[size_t(g)] = int32_t(e);
					++ninc[size_t(g)];
				}
			// work carried by element e = incidences of the nodes it is the first to touch; cuts at equal cumulative work
			std::vector<int64_t> cum(ne + 1, 0);
			for (size_t e = 0; e < ne; ++e)
			{
				int64_t w = 0;
				for (size_t j = 0; j < nl; ++j)
				{
					const int32_t g = conn[e * nl + j];
					if (first[size_t(g)] == int32_t(e))
					{
						bool seen = false; // a node listed twice in one element counts once
						for (size_t k = 0; k < j; ++k)
							seen = seen || conn[e * nl + k] == g;
						if (!seen)
							w += ninc[size_t(g)];
					}
				}
				cum[e + 1] = cum[e] + w;
			}
			std::vector<int32_t> cut(size_t(world) + 1, 0);
			cut[size_t(world)] = n_elements;
			for (int r = 1; r < world; ++r)
			{
				const int64_t target = cum[ne] * r / world;
				cut[size_t(r)] = int32_t(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
				cut[size_t(r)] = std::max(cut[size_t(r)], cut[size_t(r) - 1]);
			}
			auto rank_of = [&](int32_t e) { return int32_t(std::upper_bound(cut.begin() + 1, cut.end(), e) - (cut.begin() + 1)); };
			std::unique_ptr<pfa_partition> p(new pfa_partition()); // released only when everything below succeeded
			const int32_t e0 = cut[size_t(rank)], e1 = cut[size_t(rank) + 1];
			for (int32_t e = e0; e < e1; ++e)
				p->elements.push_back(e);
			p->n_own = e1 - e0;
			// ghost elements: outside [e0, e1) but touching a node whose first element lies inside
			for (size_t e = 0; e < ne; ++e)
			{
				if (int32_t(e) >= e0 && int32_t(e) < e1)
					continue;
				bool touches = false;
				for (size_t j = 0; j < nl && !touches; ++j)
				{
					const int32_t f = first[size_t(conn[e * nl + j])];
					touches = f >= e0 && f < e1;
				}
				if (touches)
					p->elements.push_back(int32_t(e));
			}
			p->n_ghost = int32_t(p->elements.size()) - p->n_own;
			// local numbering: first touch over own, then ghost elements
			std::vector<int32_t> g2l(size_t(n_bases), -1);
			p->conn.resize(p->elements.size() * nl);
			for (size_t t = 0; t < p->elements.size(); ++t)
				for (size_t j = 0; j < nl; ++j)
				{
					const int32_t g = conn[size_t(p->elements[t]) * nl + j];
					if (g2l[size_t(g)] < 0)
					{
						g2l[size_t(g)] = int32_t(p->l2g.size());
						p->l2g.push_back(g);
						const bool mine = rank_of(first[size_t(g)]) == rank;
						p->owned.push_back(mine ? 1 : 0);
						p->n_owned += mine ? 1 : 0;
					}
					p->conn[t * nl + j] = g2l[size_t(g)];
				}
			p->n_local = int32_t(p->l2g.size());
			*out = p.release();
			return PFA_OK;
		}
		catch (const std::bad_alloc &)
		{
			return PFA_ERR_NOMEM;
		}
	}

	int pfa_partition_sizes(const pfa_partition *p, int32_t *n_own_elements, int32_t *n_ghost_elements, int32_t *n_local_bases, int32_t *n_owned_bases)
	{
		if (!p)
			return PFA_ERR_INVALID;
		if (n_own_elements)
			*n_own_elements = p->n_own;
		if (n_ghost_elements)
			*n_ghost_elements = p->n_ghost;
		if (n_local_bases)
			*n_local_bases = p->n_local;
		if (n_owned_bases)
			*n_owned_bases = p->n_owned;
		return PFA_OK;
	}
	const int32_t *pfa_partition_elements(const pfa_partition *p) { return p ? p->elements.data() : nullptr; }
	const int32_t *pfa_partition_conn(const pfa_partition *p) { return p ? p->conn.data() : nullptr; }
	const int32_t *pfa_partition_local_to_global(const pfa_partition *p) { return p ? p->l2g.data() : nullptr; }
	const uint8_t *pfa_partition_owned(const pfa_partition *p) { return p ? p->owned.data() : nullptr; }
	void pfa_partition_destroy(pfa_partition *p) { delete p; }
}
