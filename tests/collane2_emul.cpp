// TEST INFRASTRUCTURE: CPU emulation of the owner-computes kernels (polyfem_b200/csrc/pfa_collane2.cu) built from the SAME
// header (pfa_collane2.h: element record, per-lane column math, host schedule). It walks chunks, groups, steps, triples
// and lanes the way the kernel does - one strip column per (node slot, component) shared by the two triples of the slot,
// half-warp 0 updating before half-warp 1, first contributions stored instead of added (the strips start as NaN, not 0),
// the same table words, the same address arithmetic, the 32 x 15 transposition of the flush - so that
// tests/test_collane2_emulation.py can compare the data flow with the oracle without a GPU. What it cannot check:
// launch configuration, shared-memory sizing, the TMA / mbarrier pipeline and synchronisation of the real kernel.
#include "../polyfem_b200/csrc/pfa_collane2.h"

#include <cstring>
#include <limits>

using namespace pfa::cl2;

namespace
{
	struct HostTable
	{
		const double *p;
		double operator[](int i) const { return p[i]; }
	};

	template <int NL, int NQ, int MODE>
	int emulate(int n_el, int n_bases, const int32_t *conn, const int32_t *adj_off, const int32_t *adj, const double *jit, const double *detj,
				const double *qw, const double *ref_grads, const double *lam, const double *mu, int mstride, const double *x, int small_rows,
				int chunk_steps, const uint8_t *owned, double scale, double *energy, double *grad, double *values, int64_t *stats)
	{
		double za = 0.0, zb = 0.0;
		if (MODE >= 2 && !p2_rule_weights(ref_grads, NL, NQ, za, zb))
			return -9;
		const int bucket_elements = chunk_steps % 7 == 3 ? 5 : (1 << 30); // some test cases exercise the spatial buckets of the schedule
		constexpr int RECD = Rec<NQ>::D;
		const HostTable G{ref_grads};
		// (A) records
		std::vector<double> rec(size_t(n_el) * RECD);
		double e_sum = 0.0;
		for (int e = 0; e < n_el; ++e)
		{
			double u[NL * 3];
			for (int i = 0; i < NL; ++i)
				for (int c = 0; c < 3; ++c)
					u[i * 3 + c] = x[size_t(conn[size_t(e) * NL + i]) * 3 + c];
			e_sum += element_record<NL, NQ>(jit + size_t(e) * 9, detj[e], qw, lam + size_t(e) * mstride, mu + size_t(e) * mstride, mstride, u, G,
											rec.data() + size_t(e) * RECD);
		}
		*energy = e_sum * scale;
		// own-node reference gradients, padded rows [ri][q][4] (the kernel's shared-memory table)
		std::vector<double> rgp(size_t(NL * NQ * 4), 0.0);
		for (int i = 0; i < NL; ++i)
			for (int q = 0; q < NQ; ++q)
				for (int c = 0; c < 3; ++c)
					rgp[(size_t(i) * NQ + q) * 4 + c] = ref_grads[(size_t(q) * NL + i) * 3 + c];
		// (B) column lanes
		const Schedule S = build_schedule(n_el, NL, n_bases, conn, adj_off, adj, small_rows, chunk_steps, owned, bucket_elements);
		const int n_groups = S.n_groups[0] + S.n_groups[1];
		if (int(S.chunk_off.size()) != S.n_chunks[0] + S.n_chunks[1] + 1 || S.chunk_off.back() != n_groups)
			return -6;
		const double nan = std::numeric_limits<double>::quiet_NaN();
		for (int ch = 0; ch + 1 < int(S.chunk_off.size()); ++ch)
			for (int g = S.chunk_off[size_t(ch)]; g < S.chunk_off[size_t(ch) + 1]; ++g)
			{
				const int cls = ch < S.n_chunks[0] ? 0 : 1;
				if ((g < S.n_groups[0] ? 0 : 1) != cls)
					return -7; // a chunk straddles the two launches
				const int rows = S.grp_rows[size_t(g)], s0 = S.grp_off[size_t(g)], s1 = S.grp_off[size_t(g) + 1];
				if (rows > S.rows_max[cls])
					return -2;
				std::vector<double> strip(size_t(rows) * kStripLd, nan); // [row][column], as in shared memory; NOT cleared
				double g_acc[32];
				for (int l = 0; l < 32; ++l)
					g_acc[l] = 0.0;
				for (int s = s0; s < s1; ++s)
					for (int half = 0; half < 2; ++half) // half-warp 0 updates the strips before half-warp 1
						for (int within = 0; within < 15; ++within)
						{
							const int lane = half * 16 + within, ns = within / 3, mm = within - 3 * ns, tr = half * kNodes + ns;
							const uint32_t *w = S.inc.data() + (size_t(s) * kTriples + tr) * 4;
							if (w[0] == kIdle)
								continue;
							const int b = S.grp_info[(size_t(g) * kNodes + ns) * 4];
							if (b < 0)
								return -3;
							const int e = int(w[0]);
							const int ri = (w[3] >> 16) & 0xff;
							if (e >= n_el || conn[size_t(e) * NL + ri] != b)
								return -4;
							// half-warp 0 starts its sums from the strip and stores them back; half-warp 1 sums from zero and adds afterwards;
							// rows touched for the first time start from zero (the strips are never cleared)
							int krow[NL];
							bool first[NL];
							double acc[NL][3];
							for (int j = 0; j < NL; ++j)
							{
								const int kb = (w[1 + j / 4] >> (8 * (j % 4))) & 0xff;
								krow[j] = 3 * (kb & 0x7f);
								first[j] = (kb & 0x80) != 0;
								for (int sft = 0; sft < 3; ++sft)
								{
									const int n = (mm + sft) % 3;
									if (krow[j] + n >= rows)
										return -5;
									acc[j][sft] = (half == 0 && !first[j]) ? strip[size_t(krow[j] + n) * kStripLd + within] : 0.0;
								}
							}
							column_of_element<NL, NQ, MODE>(rec.data() + size_t(e) * RECD, rgp.data() + size_t(ri) * NQ * 4, mm, G, acc, g_acc[lane], 4.0 * zb, 4.0 * (za - zb));
							for (int j = 0; j < NL; ++j)
								for (int sft = 0; sft < 3; ++sft)
								{
									double &dst = strip[size_t(krow[j] + (mm + sft) % 3) * kStripLd + within];
									dst = half == 0 ? acc[j][sft] : (first[j] ? 0.0 : dst) + acc[j][sft];
								}
						}
				// gradient: the two lanes of a column add their partial sums
				for (int within = 0; within < 15; ++within)
				{
					const int b = S.grp_info[(size_t(g) * kNodes + within / 3) * 4];
					if (b >= 0)
						grad[size_t(b) * 3 + within % 3] = scale * (g_acc[within] + g_acc[16 + within]);
				}
				// flush: blocks of 32 strip rows x 15 columns through a transposition buffer with leading dimension 15; column 3b+m
				// starts at 9*adj_off[b] + m*3*deg(b)
				double tb[32 * 15];
				for (int r0 = 0; r0 < rows; r0 += 32)
				{
					for (int lane = 0; lane < 32; ++lane)
						for (int i = 0; i < 16; ++i)
						{
							const int rr = 2 * i + (lane >> 4), c = lane & 15;
							if (c < 15)
								tb[rr * 15 + c] = r0 + rr < rows ? strip[size_t(r0 + rr) * kStripLd + c] : 0.0;
						}
					for (int lane = 0; lane < 32; ++lane)
						for (int c = 0; c < 15; ++c)
						{
							const int32_t *info = S.grp_info.data() + (size_t(g) * kNodes + c / 3) * 4;
							const int r = r0 + lane;
							if (info[0] >= 0 && r < info[2])
							{
								if (info[1] != 9 * adj_off[info[0]] || info[2] != 3 * (adj_off[info[0] + 1] - adj_off[info[0]]))
									return -8;
								values[size_t(info[1]) + size_t(c % 3) * info[2] + r] = scale * tb[lane * 15 + c];
							}
						}
				}
			}
		stats[0] = S.n_groups[0];
		stats[1] = S.n_groups[1];
		stats[2] = S.rows_max[0];
		stats[3] = S.rows_max[1];
		stats[4] = S.total_steps;
		stats[5] = S.busy;
		stats[6] = S.n_chunks[0];
		stats[7] = S.n_chunks[1];
		return 0;
	}
} // namespace

extern "C" int collane2_emulate(int n_loc, int n_qp, int n_el, int n_bases, const int32_t *conn, const int32_t *adj_off, const int32_t *adj,
								  const double *jit, const double *detj, const double *qw, const double *ref_grads, const double *lam, const double *mu,
								  int mstride, const double *x, int small_rows, int chunk_steps, int structured, const uint8_t *owned, double scale,
								  double *energy, double *grad, double *values, int64_t *stats)
{
	if (n_loc == 4 && n_qp == 1)
		return emulate<4, 1, 0>(n_el, n_bases, conn, adj_off, adj, jit, detj, qw, ref_grads, lam, mu, mstride, x, small_rows, chunk_steps, owned, scale, energy,
									grad, values, stats);
	if (n_loc == 10 && n_qp == 4 && structured == 3)
		return emulate<10, 4, 3>(n_el, n_bases, conn, adj_off, adj, jit, detj, qw, ref_grads, lam, mu, mstride, x, small_rows, chunk_steps, owned, scale, energy,
									grad, values, stats);
	if (n_loc == 10 && n_qp == 4 && structured == 2)
		return emulate<10, 4, 2>(n_el, n_bases, conn, adj_off, adj, jit, detj, qw, ref_grads, lam, mu, mstride, x, small_rows, chunk_steps, owned, scale, energy,
									grad, values, stats);
	if (n_loc == 10 && n_qp == 4 && structured)
		return emulate<10, 4, 1>(n_el, n_bases, conn, adj_off, adj, jit, detj, qw, ref_grads, lam, mu, mstride, x, small_rows, chunk_steps, owned, scale, energy,
									grad, values, stats);
	if (n_loc == 10 && n_qp == 4)
		return emulate<10, 4, 0>(n_el, n_bases, conn, adj_off, adj, jit, detj, qw, ref_grads, lam, mu, mstride, x, small_rows, chunk_steps, owned, scale, energy,
									 grad, values, stats);
	return -1;
}

// schedule statistics only (no math): n_groups[2], rows_max[2], total_steps, busy, n_chunks[2], bytes of the incidence table
extern "C" int collane2_schedule_stats(int n_loc, int n_el, int n_bases, const int32_t *conn, const int32_t *adj_off, const int32_t *adj, int small_rows,
										 int chunk_steps, int64_t *stats)
{
	const Schedule S = build_schedule(n_el, n_loc, n_bases, conn, adj_off, adj, small_rows, chunk_steps);
	stats[0] = S.n_groups[0];
	stats[1] = S.n_groups[1];
	stats[2] = S.rows_max[0];
	stats[3] = S.rows_max[1];
	stats[4] = S.total_steps;
	stats[5] = S.busy;
	stats[6] = S.n_chunks[0];
	stats[7] = S.n_chunks[1];
	stats[8] = int64_t(S.inc.size() * sizeof(uint32_t));
	return 0;
}
