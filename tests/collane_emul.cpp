// TEST INFRASTRUCTURE: CPU emulation of the column-lane kernels (polyfem_b200/csrc/pfa_collane.cu) built from the SAME
// header (pfa_collane.h: record math, per-lane column math, host schedule). It walks groups, steps, slots and lanes the
// way the kernel does - a lane-private strip per (slot, component), the same table words, the same address arithmetic -
// so that tests/test_collane_emulation.py can compare the data flow with the oracle without a GPU. What it cannot check:
// launch configuration, shared-memory sizing and synchronisation of the real kernel.
#include "../polyfem_b200/csrc/pfa_collane.h"

#include <cstring>

using namespace pfa::collane;

namespace
{
	struct HostTable
	{
		const double *p;
		double operator[](int i) const { return p[i]; }
	};

	template <int NL, int NQ, bool P2S>
	int emulate(int n_el, int n_bases, const int32_t *conn, const int32_t *adj_off, const int32_t *adj, const double *jit, const double *detj,
				const double *qw, const double *ref_grads, double lam, double mu, const double *x, int small_rows, double *energy, double *grad,
				double *values, int64_t *stats)
	{
		// (A) records
		std::vector<double> rec(size_t(n_el) * NQ * kRec);
		double e_sum = 0.0;
		for (int e = 0; e < n_el; ++e)
		{
			double u[NL * 3];
			for (int i = 0; i < NL; ++i)
				for (int c = 0; c < 3; ++c)
					u[i * 3 + c] = x[size_t(conn[size_t(e) * NL + i]) * 3 + c];
			for (int q = 0; q < NQ; ++q)
				e_sum += qp_record<NL>(jit + size_t(e) * 9, detj[e] * qw[q], lam, mu, u, ref_grads + size_t(q) * NL * 3, rec.data() + (size_t(e) * NQ + q) * kRec);
		}
		*energy = e_sum;
		// (B) column lanes
		const Schedule S = build_schedule(n_el, NL, n_bases, conn, adj_off, adj, small_rows);
		const HostTable G{ref_grads};
		const int n_groups = S.n_groups[0] + S.n_groups[1];
		for (int g = 0; g < n_groups; ++g)
		{
			const int rows = S.grp_rows[size_t(g)], s0 = S.grp_off[size_t(g)], s1 = S.grp_off[size_t(g) + 1];
			if (rows > S.rows_max[g < S.n_groups[0] ? 0 : 1])
				return -2;
			// the kernel stages the records of a step in shared memory (ColLayout): 32 lanes x PER_LANE doubles
			using CL = ColLayout<NQ>;
			std::vector<double> stages(size_t(s1 - s0) * CL::STAGE, -12345.0);
			for (int s = s0; s < s1; ++s)
				for (int lane = 0; lane < 32; ++lane)
					for (int i = 0; i < CL::PER_LANE; ++i)
					{
						int tt, kk;
						CL::staged(lane, i, tt, kk);
						if (tt >= kSlots)
							continue;
						const uint32_t e_t = S.inc[(size_t(s) * kSlots + tt) * 4];
						if (e_t != 0xffffffffu)
							stages[size_t(s - s0) * CL::STAGE + size_t(tt) * CL::SSTR + kk] = rec[size_t(e_t) * CL::RECQ + kk];
					}
			std::vector<double> strips(size_t(rows) * 32, 0.0); // [row][lane], as in shared memory
			for (int lane = 0; lane < 3 * kSlots; ++lane)
			{
				const int slot = lane / 3, m = lane - slot * 3;
				const int b = S.grp_node[size_t(g) * kSlots + slot];
				double g_acc = 0.0;
				for (int s = s0; s < s1; ++s)
				{
					const uint32_t *w = S.inc.data() + (size_t(s) * kSlots + slot) * 4;
					if (w[0] == 0xffffffffu)
						continue;
					if (b < 0)
						return -3;
					const int e = int(w[0]);
					const int ri = (w[3] >> 16) & 0xff;
					if (conn[size_t(e) * NL + ri] != b)
						return -4;
					double acc[NL][3];
					for (int j = 0; j < NL; ++j)
						acc[j][0] = acc[j][1] = acc[j][2] = 0.0;
					column_of_element<NL, NQ, P2S, P2S>(stages.data() + size_t(s - s0) * CL::STAGE + size_t(slot) * CL::SSTR, ref_grads, ri, m, G, acc, g_acc);
					for (int j = 0; j < NL; ++j)
					{
						const int k = (w[1 + j / 4] >> (8 * (j % 4))) & 0xff;
						for (int sft = 0; sft < 3; ++sft)
						{
							const int n = (m + sft) % 3;
							if (3 * k + n >= rows)
								return -5;
							strips[size_t(3 * k + n) * 32 + lane] += acc[j][sft];
						}
					}
				}
				if (b >= 0)
					grad[size_t(b) * 3 + m] = g_acc;
			}
			// flush. Column 3b+m starts at 9*adj_off[b] + m*3*deg(b) and has 3*deg(b) rows.
			constexpr int kFlushLd = 33;
			if (CL::STAGE >= 32 * kFlushLd)
			{
				// the kernel's cooperative flush: 32-row blocks through a 32 x 33 transposition buffer (the record stage)
				std::vector<double> tb(size_t(CL::STAGE), -777.0);
				for (int r0 = 0; r0 < rows; r0 += 32)
				{
					for (int lane = 0; lane < 32; ++lane)
						for (int rr = 0; rr < 32; ++rr)
							if (r0 + rr < rows)
								tb[size_t(rr) * kFlushLd + lane] = strips[size_t(r0 + rr) * 32 + lane];
					for (int c = 0; c < 3 * kSlots; ++c)
					{
						const int bc = S.grp_node[size_t(g) * kSlots + c / 3];
						const int off_c = bc >= 0 ? adj_off[bc] : 0, deg_c = bc >= 0 ? adj_off[bc + 1] - off_c : 0;
						for (int lane = 0; lane < 32; ++lane)
						{
							const int r = r0 + lane;
							if (r < 3 * deg_c)
								values[size_t(off_c) * 9 + size_t(c % 3) * 3 * deg_c + r] = tb[size_t(lane) * kFlushLd + c];
						}
					}
				}
			}
			else
				for (int lane = 0; lane < 3 * kSlots; ++lane)
				{
					const int b = S.grp_node[size_t(g) * kSlots + lane / 3], m = lane % 3;
					if (b < 0)
						continue;
					const int deg = adj_off[b + 1] - adj_off[b];
					double *dst = values + size_t(adj_off[b]) * 9 + size_t(m) * 3 * deg;
					for (int r = 0; r < 3 * deg; ++r)
						dst[r] = strips[size_t(r) * 32 + lane];
				}
		}
		stats[0] = S.n_groups[0];
		stats[1] = S.n_groups[1];
		stats[2] = S.rows_max[0];
		stats[3] = S.rows_max[1];
		stats[4] = S.total_steps;
		stats[5] = S.busy_slots;
		return 0;
	}
} // namespace

extern "C" int collane_emulate(int n_loc, int n_qp, int n_el, int n_bases, const int32_t *conn, const int32_t *adj_off, const int32_t *adj,
								 const double *jit, const double *detj, const double *qw, const double *ref_grads, double lam, double mu, const double *x,
								 int small_rows, int structured, double *energy, double *grad, double *values, int64_t *stats)
{
	if (n_loc == 4 && n_qp == 1)
		return emulate<4, 1, false>(n_el, n_bases, conn, adj_off, adj, jit, detj, qw, ref_grads, lam, mu, x, small_rows, energy, grad, values, stats);
	if (n_loc == 10 && n_qp == 4 && structured)
		return emulate<10, 4, true>(n_el, n_bases, conn, adj_off, adj, jit, detj, qw, ref_grads, lam, mu, x, small_rows, energy, grad, values, stats);
	if (n_loc == 10 && n_qp == 4)
		return emulate<10, 4, false>(n_el, n_bases, conn, adj_off, adj, jit, detj, qw, ref_grads, lam, mu, x, small_rows, energy, grad, values, stats);
	return -1;
}

// schedule statistics only (no math): n_groups[2], rows_max[2], total_steps, busy_slots, bytes of the incidence table
extern "C" int collane_schedule_stats(int n_loc, int n_el, int n_bases, const int32_t *conn, const int32_t *adj_off, const int32_t *adj, int small_rows,
										int64_t *stats)
{
	const Schedule S = build_schedule(n_el, n_loc, n_bases, conn, adj_off, adj, small_rows);
	stats[0] = S.n_groups[0];
	stats[1] = S.n_groups[1];
	stats[2] = S.rows_max[0];
	stats[3] = S.rows_max[1];
	stats[4] = S.total_steps;
	stats[5] = S.busy_slots;
	stats[6] = int64_t(S.inc.size() * sizeof(uint32_t));
	return 0;
}
