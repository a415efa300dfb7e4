// "Column lane" (owner-computes) formulation of the NeoHookean Hessian on P1 / P2 tets: the per-lane math
// and the host-side schedule, shared by the CUDA kernels (pfa_collane.cu) and by a CPU emulation of the same data flow
// (tests/collane_emul.cpp) that checks tables and indexing against the oracle without a GPU.
//
// Reference semantics are those of NLAssembler::assemble_hessian / assemble_gradient (assembler/Assembler.cpp:574-771) with
// NeoHookeanElasticity::compute_energy_hessian_aux_fast / _gradient_fast (NeoHookeanElasticity.cpp:453-658); the closed
// form and the reference-coordinate rewriting are the ones documented at the row-lane kernel in pfa_kernels.cu.
//
// Data flow: (A) one record of 34 doubles per (element, quadrature point), same layout as the row-lane kernel's
// shared-memory record, written to global memory; (B) a lane owns one CSC column (node b, component m) at a time,
// walks the elements incident to b, and adds the 3*NL entries each of them contributes to a lane-private strip of
// shared memory (address = row*32 + lane); when the node is finished the strip IS the column and is streamed to
// values[] - no atomics, no zero fill, a fixed summation order.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <vector>

#if defined(__CUDACC__)
#define PFA_HD __host__ __device__ __forceinline__
#else
#define PFA_HD inline
#endif

namespace pfa
{
	namespace collane
	{
		constexpr int kRec = 34;   // doubles per (element, qp) record
		constexpr int kSlots = 10; // nodes a warp works on at a time (3 lanes each)

		PFA_HD double det3(const double *F)
		{
			return F[0] * (F[4] * F[8] - F[5] * F[7]) - F[1] * (F[3] * F[8] - F[5] * F[6]) + F[2] * (F[3] * F[7] - F[4] * F[6]);
		}
		PFA_HD void cofactor3(const double *F, double *C)
		{
			C[0] = F[4] * F[8] - F[5] * F[7];
			C[1] = F[5] * F[6] - F[3] * F[8];
			C[2] = F[3] * F[7] - F[4] * F[6];
			C[3] = F[2] * F[7] - F[1] * F[8];
			C[4] = F[0] * F[8] - F[2] * F[6];
			C[5] = F[1] * F[6] - F[0] * F[7];
			C[6] = F[1] * F[5] - F[2] * F[4];
			C[7] = F[2] * F[3] - F[0] * F[5];
			C[8] = F[0] * F[4] - F[1] * F[3];
		}

		// Record of one (element, quadrature point). J = J^-T row-major, da = det * weight, u[NL][3] nodal displacements,
		// rg[NL][3] reference gradients at this point. Returns the energy density times da.
		//   [0..5]   mu da K,  K = J^-T J^-1 (00 01 02 11 12 22)        [6..14]  T = (c2 da F) cof(J^-T)^T
		//   [15..23] CJ = cof(F) J^-T^T                                   [24..32] PJ = (P da) J^-T^T      [33] c1 da
		template <int NL>
		PFA_HD double qp_record(const double *J, double da, double lam, double mu, const double *u, const double *rg, double *rec)
		{
			double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
			for (int i = 0; i < NL; ++i)
			{
				const double u0 = u[i * 3 + 0], u1 = u[i * 3 + 1], u2 = u[i * 3 + 2];
				const double *rr = rg + i * 3;
				const double d0 = rr[0] * J[0] + rr[1] * J[3] + rr[2] * J[6];
				const double d1 = rr[0] * J[1] + rr[1] * J[4] + rr[2] * J[7];
				const double d2 = rr[0] * J[2] + rr[1] * J[5] + rr[2] * J[8];
				F[0] += u0 * d0;
				F[1] += u0 * d1;
				F[2] += u0 * d2;
				F[3] += u1 * d0;
				F[4] += u1 * d1;
				F[5] += u1 * d2;
				F[6] += u2 * d0;
				F[7] += u2 * d1;
				F[8] += u2 * d2;
			}
			const double Jd = det3(F);
			const double lJ = log(Jd); // NaN for J <= 0, propagates like the reference
			double C[9];
			cofactor3(F, C);
			const double invJ = 1.0 / Jd;
			const double pc = (lam * lJ - mu) * invJ; // P = mu F + pc C ; c2 = pc
			double sq = 0.0;
			for (int k = 0; k < 9; ++k)
				sq += F[k] * F[k];
			for (int aa = 0; aa < 3; ++aa)
			{
				const double p0 = (mu * F[aa * 3 + 0] + pc * C[aa * 3 + 0]) * da;
				const double p1 = (mu * F[aa * 3 + 1] + pc * C[aa * 3 + 1]) * da;
				const double p2 = (mu * F[aa * 3 + 2] + pc * C[aa * 3 + 2]) * da;
				for (int c = 0; c < 3; ++c)
					rec[24 + aa * 3 + c] = p0 * J[c * 3 + 0] + p1 * J[c * 3 + 1] + p2 * J[c * 3 + 2];
			}
			const double mu_da = mu * da;
			rec[0] = mu_da * (J[0] * J[0] + J[1] * J[1] + J[2] * J[2]);
			rec[1] = mu_da * (J[0] * J[3] + J[1] * J[4] + J[2] * J[5]);
			rec[2] = mu_da * (J[0] * J[6] + J[1] * J[7] + J[2] * J[8]);
			rec[3] = mu_da * (J[3] * J[3] + J[4] * J[4] + J[5] * J[5]);
			rec[4] = mu_da * (J[3] * J[6] + J[4] * J[7] + J[5] * J[8]);
			rec[5] = mu_da * (J[6] * J[6] + J[7] * J[7] + J[8] * J[8]);
			double R[9];
			cofactor3(J, R); // rows of cof(J^-T)
			const double c2da = pc * da;
			for (int rr = 0; rr < 3; ++rr)
			{
				const double f0 = c2da * F[rr * 3 + 0], f1 = c2da * F[rr * 3 + 1], f2 = c2da * F[rr * 3 + 2];
				for (int k = 0; k < 3; ++k)
					rec[6 + rr * 3 + k] = f0 * R[k * 3 + 0] + f1 * R[k * 3 + 1] + f2 * R[k * 3 + 2];
				for (int c = 0; c < 3; ++c)
					rec[15 + rr * 3 + c] = C[rr * 3 + 0] * J[c * 3 + 0] + C[rr * 3 + 1] * J[c * 3 + 1] + C[rr * 3 + 2] * J[c * 3 + 2];
			}
			rec[33] = (mu + lam * (1.0 - lJ)) * invJ * invJ * da;
			return (0.5 * mu * (sq - 3.0 - 2.0 * lJ) + 0.5 * lam * lJ * lJ) * da;
		}

		// The 25 entries of a record that the lane of component mm needs (lane-dependent ADDRESSES, fixed register names)
		struct LaneRecord
		{
			double K[6], ta[3], tb[3], c0[3], c1[3], c2[3], pj[3], c1da;
		};
		PFA_HD void load_lane_record(const double *rec, int mm, int ra, int rb, LaneRecord &r)
		{
			for (int k = 0; k < 6; ++k)
				r.K[k] = rec[k];
			for (int k = 0; k < 3; ++k)
			{
				r.ta[k] = rec[6 + ra * 3 + k];
				r.tb[k] = rec[6 + rb * 3 + k];
				r.c0[k] = rec[15 + mm * 3 + k];
				r.c1[k] = rec[15 + ra * 3 + k];
				r.c2[k] = rec[15 + rb * 3 + k];
				r.pj[k] = rec[24 + mm * 3 + k];
			}
			r.c1da = rec[33];
		}

		// One incident element's contribution to the column of dof (its local node ri, component mm):
		//   acc[j][s] += H[(ri, mm), (j, (mm + s) % 3)]   (rotated by mm, like the row-lane kernel)      g_row += G[(ri, mm)]
		// rec: [NQ][kRec] of the element; rg: reference gradients [NQ][NL][3] (row side, index depends on the lane);
		// G: the same table as the column operand (device: __constant__ memory, uniform index).
		// P2S: the table has the structural zeros / equal components of the P2 tet basis (checked on the host by
		// p2_table_structured, pfa_kernels.cu): 20 instead of 30 DFMA-pipe operations per (column component, quadrature point)
		// PIPE: load the next point's record entries before the current point's math (for records read from global memory;
		// the kernel stages records in shared memory and uses PIPE = false)
		template <int NL, int NQ, bool P2S, bool PIPE, class ColTable>
		PFA_HD void column_of_element(const double *rec_e, const double *rg, int ri, int mm, const ColTable &G, double (*acc)[3], double &g_row)
		{
			const int ra = (mm + 1) % 3, rb = (mm + 2) % 3;
			LaneRecord nxt;
			if (PIPE)
				load_lane_record(rec_e, mm, ra, rb, nxt);
			// not unrolled on the device: keeps the register count of the column kernel near the row-lane kernel's (the column
			// table is then read with a uniform run-time index, LDC)
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
			for (int qq = 0; qq < NQ; ++qq)
			{
				LaneRecord rec;
				if (PIPE)
				{
					rec = nxt;
					if (qq + 1 < NQ)
						load_lane_record(rec_e + (qq + 1) * kRec, mm, ra, rb, nxt);
				}
				else
					load_lane_record(rec_e + qq * kRec, mm, ra, rb, rec);
				const double *gr = rg + (qq * NL + ri) * 3;
				const double g0 = gr[0], g1 = gr[1], g2 = gr[2];
				g_row = fma(g0, rec.pj[0], fma(g1, rec.pj[1], fma(g2, rec.pj[2], g_row)));
				const double K00 = rec.K[0], K01 = rec.K[1], K02 = rec.K[2], K11 = rec.K[3], K12 = rec.K[4], K22 = rec.K[5];
				const double v0 = K00 * g0 + K01 * g1 + K02 * g2;
				const double v1 = K01 * g0 + K11 * g1 + K12 * g2;
				const double v2 = K02 * g0 + K12 * g1 + K22 * g2;
				const double *ta = rec.ta, *tb = rec.tb;
				const double b0 = tb[1] * g2 - tb[2] * g1, b1 = tb[2] * g0 - tb[0] * g2, b2 = tb[0] * g1 - tb[1] * g0; //  t_b x g
				const double a0 = ta[2] * g1 - ta[1] * g2, a1 = ta[0] * g2 - ta[2] * g0, a2 = ta[1] * g0 - ta[0] * g1; // -t_a x g
				const double cA = rec.c1da * (rec.c0[0] * g0 + rec.c0[1] * g1 + rec.c0[2] * g2);
				const double *c0r = rec.c0, *c1r = rec.c1, *c2r = rec.c2;
				double Y[3][3];
				Y[0][0] = fma(cA, c0r[0], v0);
				Y[0][1] = fma(cA, c0r[1], v1);
				Y[0][2] = fma(cA, c0r[2], v2);
				Y[1][0] = fma(cA, c1r[0], b0);
				Y[1][1] = fma(cA, c1r[1], b1);
				Y[1][2] = fma(cA, c1r[2], b2);
				Y[2][0] = fma(cA, c2r[0], a0);
				Y[2][1] = fma(cA, c2r[1], a1);
				Y[2][2] = fma(cA, c2r[2], a2);
				if constexpr (P2S && NL == 10)
				{
					const int o = qq * NL * 3;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
					for (int n = 0; n < 3; ++n)
					{
						const double y0 = Y[n][0], y1 = Y[n][1], y2 = Y[n][2];
						const double u12 = y1 + y2, u02 = y0 + y2, u01 = y0 + y1, sy = y0 + u12;
						acc[0][n] = fma(sy, G[o + 0], acc[0][n]);
						acc[1][n] = fma(y0, G[o + 3], acc[1][n]);
						acc[2][n] = fma(y1, G[o + 7], acc[2][n]);
						acc[3][n] = fma(y2, G[o + 11], acc[3][n]);
						acc[4][n] = fma(y0, G[o + 12], fma(u12, G[o + 13], acc[4][n]));
						acc[5][n] = fma(y0, G[o + 15], fma(y1, G[o + 16], acc[5][n]));
						acc[6][n] = fma(y1, G[o + 19], fma(u02, G[o + 18], acc[6][n]));
						acc[7][n] = fma(y2, G[o + 23], fma(u01, G[o + 21], acc[7][n]));
						acc[8][n] = fma(y0, G[o + 24], fma(y2, G[o + 26], acc[8][n]));
						acc[9][n] = fma(y1, G[o + 28], fma(y2, G[o + 29], acc[9][n]));
					}
				}
				else
				{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
					for (int j = 0; j < NL; ++j)
					{
						const double c0 = G[(qq * NL + j) * 3 + 0], c1 = G[(qq * NL + j) * 3 + 1], c2 = G[(qq * NL + j) * 3 + 2];
						acc[j][0] = fma(Y[0][0], c0, fma(Y[0][1], c1, fma(Y[0][2], c2, acc[j][0])));
						acc[j][1] = fma(Y[1][0], c0, fma(Y[1][1], c1, fma(Y[1][2], c2, acc[j][1])));
						acc[j][2] = fma(Y[2][0], c0, fma(Y[2][1], c1, fma(Y[2][2], c2, acc[j][2])));
					}
				}
			}
		}

		// Shared memory of one warp of the column kernel: the records of the (up to) 10 elements of the current step, staged
		// cooperatively (a lane-private read of 10 different records costs 10 L1 wavefronts per load instruction; coalesced
		// copies cost 2), followed by the strips. Slot stride odd: the 10 slots read their records conflict-free.
		template <int NQ>
		struct ColLayout
		{
			static constexpr int RECQ = NQ * kRec;                       // doubles per element record
			static constexpr int SSTR = RECQ | 1;                        // slot stride in the stage
			static constexpr int STAGE = (kSlots * SSTR + 1) & ~1;       // doubles per warp
			static constexpr int PER_LANE = (kSlots * RECQ + 31) / 32;   // staged doubles per lane and step
			// the i-th double lane `lane` stages: slot tt (>= kSlots: nothing), offset kk inside that slot's element record
			static PFA_HD void staged(int lane, int i, int &tt, int &kk)
			{
				const int idx = lane + 32 * i;
				tt = idx / RECQ;
				kk = idx - tt * RECQ;
			}
			static PFA_HD size_t warp_bytes(int strip_rows) { return sizeof(double) * (size_t(STAGE) + size_t(strip_rows) * 32); }
		};

		// ---- host side: which lane works on which (element, node) incidence, in which order ----
		// Nodes are put into groups of kSlots nodes (one per warp slot); a warp processes a group from its first to its last
		// step and then flushes the kSlots*3 finished columns. Class 0: groups whose strips need <= small_rows rows (3*deg),
		// class 1: the rest (their own launch with larger strips). Within a class nodes are ordered by their number of
		// incident elements (descending), then by node id, so that the slots of a group finish together and stay local.
		struct Schedule
		{
			int n_groups[2] = {0, 0};
			int rows_max[2] = {0, 0};
			int64_t total_steps = 0;
			int64_t busy_slots = 0;          // (element, node) incidences = non-idle (step, slot) pairs
			std::vector<int32_t> grp_node; // [G][kSlots] node of each slot, -1 = unused slot
			std::vector<int32_t> grp_off;  // [G+1] first step of each group
			std::vector<int32_t> grp_rows; // [G] strip rows the group needs (3 * max degree of its nodes)
			// [total_steps][kSlots][4] words: element (-1 = idle), then 12 bytes: k_j (position of local node j of that element in
			// the adjacency of the slot's node) for j < NL, byte 10 = local index of the slot's node in the element
			std::vector<uint32_t> inc;
		};

		inline Schedule build_schedule(int n_el, int NL, int n_bases, const int32_t *conn, const int32_t *adj_off, const int32_t *adj, int small_rows)
		{
			Schedule S;
			std::vector<int32_t> cnt(size_t(n_bases) + 1, 0);
			for (int64_t t = 0; t < int64_t(n_el) * NL; ++t)
				++cnt[size_t(conn[t]) + 1];
			for (int b = 0; b < n_bases; ++b)
				cnt[size_t(b) + 1] += cnt[size_t(b)];
			std::vector<int32_t> inc_e(size_t(n_el) * NL), fill(cnt.begin(), cnt.end() - 1); // incidence (e*NL + i) lists per node, element order
			for (int e = 0; e < n_el; ++e)
				for (int i = 0; i < NL; ++i)
					inc_e[size_t(fill[size_t(conn[size_t(e) * NL + i])]++)] = e * NL + i;
			auto n_inc = [&](int b) { return cnt[size_t(b) + 1] - cnt[size_t(b)]; };
			auto rows_of = [&](int b) { return 3 * (adj_off[b + 1] - adj_off[b]); };
			std::vector<int32_t> order[2];
			for (int b = 0; b < n_bases; ++b)
				if (n_inc(b) > 0)
					order[rows_of(b) <= small_rows ? 0 : 1].push_back(b);
			for (int c = 0; c < 2; ++c)
				std::stable_sort(order[c].begin(), order[c].end(), [&](int a, int b) { return n_inc(a) > n_inc(b); });
			S.grp_off.push_back(0);
			for (int c = 0; c < 2; ++c)
			{
				const std::vector<int32_t> &o = order[c];
				for (size_t first = 0; first < o.size(); first += kSlots)
				{
					const size_t last = std::min(o.size(), first + kSlots);
					int steps = 0, rows = 0;
					for (size_t t = first; t < last; ++t)
					{
						steps = std::max(steps, n_inc(o[t]));
						rows = std::max(rows, rows_of(o[t]));
					}
					for (int s = 0; s < kSlots; ++s)
						S.grp_node.push_back(first + s < last ? o[first + s] : -1);
					const size_t base = S.inc.size();
					S.inc.resize(base + size_t(steps) * kSlots * 4, 0u);
					for (int st = 0; st < steps; ++st)
						for (int s = 0; s < kSlots; ++s)
						{
							uint32_t *w = S.inc.data() + base + (size_t(st) * kSlots + s) * 4;
							w[0] = 0xffffffffu;
							if (first + s >= last)
								continue;
							const int b = o[first + s];
							if (st >= n_inc(b))
								continue;
							const int ei = inc_e[size_t(cnt[size_t(b)]) + st];
							const int e = ei / NL, i = ei - e * NL;
							w[0] = uint32_t(e);
							uint8_t bytes[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
							const int32_t *lo = adj + adj_off[b], *hi = adj + adj_off[b + 1];
							for (int j = 0; j < NL; ++j)
								bytes[j] = uint8_t(std::lower_bound(lo, hi, conn[size_t(e) * NL + j]) - lo);
							bytes[10] = uint8_t(i);
							for (int k = 0; k < 12; ++k)
								w[1 + k / 4] |= uint32_t(bytes[k]) << (8 * (k % 4));
							++S.busy_slots;
						}
					S.total_steps += steps;
					S.grp_off.push_back(int32_t(S.total_steps));
					S.grp_rows.push_back(rows);
					S.rows_max[c] = std::max(S.rows_max[c], rows);
					++S.n_groups[c];
				}
			}
			return S;
		}
	} // namespace collane
} // namespace pfa
