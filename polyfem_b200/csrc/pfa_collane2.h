// Owner-computes ("column lane") NeoHookean assembly on affine P1 / P2 tets, second generation: the per-lane math, the
// element record and the host-side schedule, shared by the CUDA kernels (pfa_collane2.cu) and by a CPU emulation of the
// same data flow (tests/collane2_emul.cpp) that is checked against the oracle without a GPU.
//
// Reference semantics: NLAssembler::assemble_hessian / assemble_gradient / assemble_energy (assembler/Assembler.cpp:495-771)
// over NeoHookeanElasticity::compute_energy_hessian_aux_fast / _gradient_fast / compute_energy_aux
// (assembler/NeoHookeanElasticity.cpp:338-658), scattered as SparseMatrixCache does (utils/MatrixCache.cpp:88-113).
//
// Everything is written in REFERENCE coordinates so that the per-element state is small (54 doubles for P2):
//   J = J^-T (row-major), dJ = det J, K = J J^T, ghat_i = reference gradient of basis i at the quadrature point,
//   Fhat = sum_i u_i (x) ghat_i,   A = F J^-1 = J^-1 + Fhat   (F = I + Fhat J is the deformation gradient),
//   cof(F) J^T = dJ cof(A),   F cof(J)^T = dJ A,   det F = det(A) dJ.
// With c1 = (mu + lambda (1 - ln det F)) / det F^2, c2 = (lambda ln det F - mu) / det F, da = det(J_geom) w_q:
//   c1t = c1 da dJ^2, c2t = c2 da dJ, muda = mu da   and the Hessian block of local nodes (i, j) is
//   H_ij = sum_q [ muda (ghat_i . K ghat_j) I + c1t (cof(A) ghat_i)(cof(A) ghat_j)^T - c2t hat(A (ghat_i x ghat_j)) ]
// (closed form of B^T H_F B, NeoHookeanElasticity.cpp:613-656; tests/test_gpu_parity.py checks it against the dense form).
//
// Data flow: (A) one record per element (A, c1t, c2t, muda per quadrature point + K); (B) a CSC column (node b,
// component m) is owned by one lane PAIR: the two lanes walk the elements incident to b (even / odd positions of its
// incidence list), and add the 3*NL entries each element contributes to ONE strip column of shared memory
// (address = row*16 + column: bank = column, no conflicts, no atomics; the two lanes sit in different half-warps and
// update one after the other); when the node is finished the strip IS the column and is streamed to values[] - every
// output is written exactly once, in a fixed summation order, no zero fill.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <vector>

#if defined(__CUDACC__)
#define PFA2_HD __host__ __device__ __forceinline__
#else
#define PFA2_HD inline
#endif

namespace pfa
{
	namespace cl2
	{
		constexpr int kNodes = 5;    // nodes a warp works on at a time
		constexpr int kTriples = 10; // lane triples (one per column component): two per node
		constexpr int kQpRec = 12;   // doubles per quadrature point in a record: A (9), c1t, c2t, muda
		constexpr int kStripLd = 16; // strip columns per warp (15 used)
		constexpr uint32_t kIdle = 0xffffffffu;

		template <int NQ>
		struct Rec
		{
			static constexpr int D = NQ * kQpRec + 6; // doubles per element record (K at the end): P1 18, P2 54
		};

		PFA2_HD double det3(const double *F)
		{
			return F[0] * (F[4] * F[8] - F[5] * F[7]) - F[1] * (F[3] * F[8] - F[5] * F[6]) + F[2] * (F[3] * F[7] - F[4] * F[6]);
		}
		// c = a x b
		PFA2_HD void cross3(const double *a, const double *b, double *c)
		{
			c[0] = a[1] * b[2] - a[2] * b[1];
			c[1] = a[2] * b[0] - a[0] * b[2];
			c[2] = a[0] * b[1] - a[1] * b[0];
		}

		// Record of one element. J = J^-T row-major, detg = det of the geometric Jacobian (da = detg * w_q), u[NL][3] nodal
		// displacements, G[(q*NL + i)*3 + c] reference gradients, lam / mu [mstride] (1 or NQ values). Returns the energy.
		template <int NL, int NQ, class Tab>
		PFA2_HD double element_record(const double *J, double detg, const double *qw, const double *lam, const double *mu, int mstride, const double *u,
									  const Tab &G, double *rec)
		{
			double R[9]; // rows of cof(J)
			cross3(J + 3, J + 6, R);
			cross3(J + 6, J, R + 3);
			cross3(J, J + 3, R + 6);
			const double dJ = J[0] * R[0] + J[1] * R[1] + J[2] * R[2];
			const double idJ = 1.0 / dJ;
			double Ji[9]; // J^-1 = cof(J)^T / dJ
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					Ji[r * 3 + c] = R[c * 3 + r] * idJ;
			double K[6];
			K[0] = J[0] * J[0] + J[1] * J[1] + J[2] * J[2];
			K[1] = J[0] * J[3] + J[1] * J[4] + J[2] * J[5];
			K[2] = J[0] * J[6] + J[1] * J[7] + J[2] * J[8];
			K[3] = J[3] * J[3] + J[4] * J[4] + J[5] * J[5];
			K[4] = J[3] * J[6] + J[4] * J[7] + J[5] * J[8];
			K[5] = J[6] * J[6] + J[7] * J[7] + J[8] * J[8];
			for (int k = 0; k < 6; ++k)
				rec[NQ * kQpRec + k] = K[k];
			double energy = 0.0;
			for (int q = 0; q < NQ; ++q)
			{
				double A[9];
				for (int k = 0; k < 9; ++k)
					A[k] = Ji[k];
				for (int i = 0; i < NL; ++i)
				{
					const double g0 = G[(q * NL + i) * 3 + 0], g1 = G[(q * NL + i) * 3 + 1], g2 = G[(q * NL + i) * 3 + 2];
					for (int aa = 0; aa < 3; ++aa)
					{
						const double ua = u[i * 3 + aa];
						A[aa * 3 + 0] = fma(ua, g0, A[aa * 3 + 0]);
						A[aa * 3 + 1] = fma(ua, g1, A[aa * 3 + 1]);
						A[aa * 3 + 2] = fma(ua, g2, A[aa * 3 + 2]);
					}
				}
				const double JF = det3(A) * dJ;
				const double lJ = log(JF); // NaN for det F <= 0, propagates like the reference
				const double invJ = 1.0 / JF;
				const double l = lam[mstride == 1 ? 0 : q], m_ = mu[mstride == 1 ? 0 : q];
				const double da = detg * qw[q];
				double *rq = rec + q * kQpRec;
				for (int k = 0; k < 9; ++k)
					rq[k] = A[k];
				rq[9] = (m_ + l * (1.0 - lJ)) * invJ * invJ * da * dJ * dJ;
				rq[10] = (l * lJ - m_) * invJ * da * dJ;
				rq[11] = m_ * da;
				double sq = 0.0; // |F|_F^2 = sum_a A[a] K A[a]^T
				for (int aa = 0; aa < 3; ++aa)
				{
					const double a0 = A[aa * 3 + 0], a1 = A[aa * 3 + 1], a2 = A[aa * 3 + 2];
					sq += a0 * (K[0] * a0 + K[1] * a1 + K[2] * a2) + a1 * (K[1] * a0 + K[3] * a1 + K[4] * a2) + a2 * (K[2] * a0 + K[4] * a1 + K[5] * a2);
				}
				energy += (0.5 * m_ * (sq - 3.0 - 2.0 * lJ) + 0.5 * l * lJ * lJ) * da;
			}
			return energy;
		}

		// One incident element's contribution to the column of dof (its local node ri, component mm):
		//   acc[j][s] += H[(j, (mm + s) % 3), (ri, mm)]   (the column, by symmetry the row; rotated by mm)     g_row += G[(ri, mm)]
		// rec: the element record; gri[q*4 + c]: reference gradient of the lane's own local node ri at point q (row of a padded
		// table); G: reference gradients of all nodes with a uniform index (device: __constant__ memory).
		// MODE 0: any table. MODE 1 (P2S): the table has the structural zeros / equal components of the P2 tet basis
		// (p2_table_structured): 20 instead of 30 FP64 operations per (component, point). MODE 2 (P2Z): the table is the P2 Lagrange
		// basis on the symmetric 4-point rule (p2_rule_weights): the barycentric coordinates of point q are zb everywhere except za
		// at vertex 3 - q, so with S = sum_q Y_q and V_v = 4 zb S + 4 (za - zb) Y_{3-v} (= 4 sum_q lambda_v(q) Y_q)
		//   vertex node v:  grad lambda_v . (V_v - S),     edge node (a, b):  grad lambda_b . V_a + grad lambda_a . V_b,
		// and grad lambda = (-1,-1,-1), e_x, e_y, e_z turns every dot product into a pick or a sum: 54 operations per component
		// for all four points instead of 80, and no table loads. z4b = 4 zb, zbeta = 4 (za - zb).
		// MODE 3 (P2Z, streamed): the same sums without holding the four Y_q: the point coefficients are scaled by zbeta, so that
		// Y'_q = zbeta Y_q goes straight into the entries that involve vertex 3 - q, and S' = sum_q Y'_q supplies the rest at the
		// end (4 zb S = (z4b / zbeta) S', S = S' / zbeta): 27 instead of 81 live doubles, about 4% more operations.
		template <int NL, int NQ, int MODE, class CTab>
		PFA2_HD void column_of_element(const double *rec, const double *gri, int mm, const CTab &G, double (*acc)[3], double &g_row, double z4b = 0.0,
									   double zbeta = 0.0)
		{
			constexpr bool P2S = MODE == 1;
			constexpr bool P2Z = MODE == 2 && NL == 10 && NQ == 4;
			constexpr bool P2Y = MODE == 3 && NL == 10 && NQ == 4;
			double Yq[P2Z ? NQ : 1][3][3];
			double Ss[P2Y ? 3 : 1][3];
			double g_loc = 0.0;
			const double K00 = rec[NQ * kQpRec + 0], K01 = rec[NQ * kQpRec + 1], K02 = rec[NQ * kQpRec + 2];
			const double K11 = rec[NQ * kQpRec + 3], K12 = rec[NQ * kQpRec + 4], K22 = rec[NQ * kQpRec + 5];
			// rows mm, mm+1, mm+2 (mod 3) of A: lane-dependent offsets into the record (the three lanes of a triple read the three
			// rows in rotated order)
			const int ro0 = 3 * mm, ro1 = mm == 2 ? 0 : ro0 + 3, ro2 = mm == 0 ? 6 : ro0 - 3;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
			for (int qq = 0; qq < NQ; ++qq)
			{
				const double *a = rec + qq * kQpRec;
				double r0[3], r1[3], r2[3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
				for (int c = 0; c < 3; ++c)
				{
					r0[c] = a[ro0 + c];
					r1[c] = a[ro1 + c];
					r2[c] = a[ro2 + c];
				}
				const double c1t = P2Y ? zbeta * a[9] : a[9], c2t = P2Y ? zbeta * a[10] : a[10], muda = P2Y ? zbeta * a[11] : a[11];
				double c0[3], c1[3], c2[3]; // rows mm, mm+1, mm+2 of cof(A)
				cross3(r1, r2, c0);
				cross3(r2, r0, c1);
				cross3(r0, r1, c2);
				const double g0 = gri[qq * 4 + 0], g1 = gri[qq * 4 + 1], g2 = gri[qq * 4 + 2];
				const double v0 = muda * (K00 * g0 + K01 * g1 + K02 * g2);
				const double v1 = muda * (K01 * g0 + K11 * g1 + K12 * g2);
				const double v2 = muda * (K02 * g0 + K12 * g1 + K22 * g2);
				const double cdot = c0[0] * g0 + c0[1] * g1 + c0[2] * g2;
				const double cA = c1t * cdot;
				if constexpr (P2Y)
					g_loc = fma(r0[0], v0, fma(r0[1], v1, fma(r0[2], v2, fma(c2t, cdot, g_loc))));
				else
					g_row = fma(r0[0], v0, fma(r0[1], v1, fma(r0[2], v2, fma(c2t, cdot, g_row))));
				const double s0 = c2t * g0, s1 = c2t * g1, s2 = c2t * g2;
				double Y[3][3];
				Y[0][0] = fma(cA, c0[0], v0);
				Y[0][1] = fma(cA, c0[1], v1);
				Y[0][2] = fma(cA, c0[2], v2);
				Y[1][0] = fma(cA, c1[0], r2[1] * s2 - r2[2] * s1); //  A[mm+2] x (c2t g)
				Y[1][1] = fma(cA, c1[1], r2[2] * s0 - r2[0] * s2);
				Y[1][2] = fma(cA, c1[2], r2[0] * s1 - r2[1] * s0);
				Y[2][0] = fma(cA, c2[0], r1[2] * s1 - r1[1] * s2); // -A[mm+1] x (c2t g)
				Y[2][1] = fma(cA, c2[1], r1[0] * s2 - r1[2] * s0);
				Y[2][2] = fma(cA, c2[2], r1[1] * s0 - r1[0] * s1);
				if constexpr (P2Z)
				{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
					for (int n = 0; n < 3; ++n)
						for (int c = 0; c < 3; ++c)
							Yq[qq][n][c] = Y[n][c];
				}
				else if constexpr (P2Y)
				{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
					for (int n = 0; n < 3; ++n)
					{
						const double y0 = Y[n][0], y1 = Y[n][1], y2 = Y[n][2];
						const double sy = y0 + y1 + y2;
						if (qq == 0)
						{
							Ss[n][0] = y0;
							Ss[n][1] = y1;
							Ss[n][2] = y2;
							acc[3][n] += y2; // point 0 carries za at vertex 3
							acc[7][n] -= sy;
							acc[8][n] += y0;
							acc[9][n] += y1;
						}
						else
						{
							Ss[n][0] += y0;
							Ss[n][1] += y1;
							Ss[n][2] += y2;
							if (qq == 1) // vertex 2
							{
								acc[2][n] += y1;
								acc[5][n] += y0;
								acc[6][n] -= sy;
								acc[9][n] += y2;
							}
							else if (qq == 2) // vertex 1
							{
								acc[1][n] += y0;
								acc[4][n] -= sy;
								acc[5][n] += y1;
								acc[8][n] += y2;
							}
							else // vertex 0
							{
								acc[0][n] -= sy;
								acc[4][n] += y0;
								acc[6][n] += y1;
								acc[7][n] += y2;
							}
						}
					}
				}
				else if constexpr (P2S && NL == 10)
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
						const double h0 = G[(qq * NL + j) * 3 + 0], h1 = G[(qq * NL + j) * 3 + 1], h2 = G[(qq * NL + j) * 3 + 2];
						acc[j][0] = fma(Y[0][0], h0, fma(Y[0][1], h1, fma(Y[0][2], h2, acc[j][0])));
						acc[j][1] = fma(Y[1][0], h0, fma(Y[1][1], h1, fma(Y[1][2], h2, acc[j][1])));
						acc[j][2] = fma(Y[2][0], h0, fma(Y[2][1], h1, fma(Y[2][2], h2, acc[j][2])));
					}
				}
			}
			if constexpr (P2Y)
			{
				const double ib = 1.0 / zbeta, k1 = z4b * ib, k2 = k1 - ib;
				g_row = fma(ib, g_loc, g_row);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
				for (int n = 0; n < 3; ++n)
				{
					const double s0 = Ss[n][0], s1 = Ss[n][1], s2 = Ss[n][2];
					const double ss = s0 + s1 + s2;
					acc[0][n] = fma(-k2, ss, acc[0][n]);
					acc[1][n] = fma(k2, s0, acc[1][n]);
					acc[2][n] = fma(k2, s1, acc[2][n]);
					acc[3][n] = fma(k2, s2, acc[3][n]);
					acc[4][n] = fma(k1, s0 - ss, acc[4][n]);
					acc[5][n] = fma(k1, s1 + s0, acc[5][n]);
					acc[6][n] = fma(k1, s1 - ss, acc[6][n]);
					acc[7][n] = fma(k1, s2 - ss, acc[7][n]);
					acc[8][n] = fma(k1, s2 + s0, acc[8][n]);
					acc[9][n] = fma(k1, s2 + s1, acc[9][n]);
				}
			}
			if constexpr (P2Z)
			{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
				for (int n = 0; n < 3; ++n)
				{
					double S[3], V[4][3], sV[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
					for (int c = 0; c < 3; ++c)
					{
						S[c] = (Yq[0][n][c] + Yq[1][n][c]) + (Yq[2][n][c] + Yq[3][n][c]);
						const double b4 = z4b * S[c];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
						for (int v = 0; v < 4; ++v)
							V[v][c] = fma(zbeta, Yq[3 - v][n][c], b4); // point 3 - v carries the weight za at vertex v
					}
					for (int v = 0; v < 4; ++v)
						sV[v] = V[v][0] + V[v][1] + V[v][2];
					const double sS = S[0] + S[1] + S[2];
					acc[0][n] += sS - sV[0];            // vertex 0: grad lambda_0 = -(1, 1, 1)
					acc[1][n] += V[1][0] - S[0];        // vertex 1: e_x
					acc[2][n] += V[2][1] - S[1];        // vertex 2: e_y
					acc[3][n] += V[3][2] - S[2];        // vertex 3: e_z
					acc[4][n] += V[0][0] - sV[1];       // edge (0, 1)
					acc[5][n] += V[1][1] + V[2][0];     // edge (1, 2)
					acc[6][n] += V[0][1] - sV[2];       // edge (2, 0)
					acc[7][n] += V[0][2] - sV[3];       // edge (0, 3)
					acc[8][n] += V[1][2] + V[3][0];     // edge (1, 3)
					acc[9][n] += V[2][2] + V[3][1];     // edge (2, 3)
				}
			}
		}

		// The P2 Lagrange basis on the symmetric 4-point tet rule: true when ref_grads[4][10][3] equals, to 1e-14, the gradients
		// generated from barycentric coordinates that are zb at every vertex except za at vertex 3 - q (the reference's point order,
		// quadrature/TetQuadrature.cpp with auto_tetrahedron.ipp order 2); returns za, zb.
		inline bool p2_rule_weights(const double *g, int n_loc, int n_qp, double &za, double &zb)
		{
			if (n_loc != 10 || n_qp != 4 || g == nullptr)
				return false;
			// lambda_1 = (d phi_1 / dx + 1) / 4 at point 0 (where vertex 3 carries za) is zb; at point 2 it is za
			zb = (g[(0 * 10 + 1) * 3 + 0] + 1.0) / 4.0;
			za = (g[(2 * 10 + 1) * 3 + 0] + 1.0) / 4.0;
			static const int ev[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
			static const double gl[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
			for (int q = 0; q < 4; ++q)
			{
				double lam[4];
				for (int v = 0; v < 4; ++v)
					lam[v] = v == 3 - q ? za : zb;
				if (std::fabs(lam[0] + lam[1] + lam[2] + lam[3] - 1.0) > 1e-14)
					return false;
				for (int j = 0; j < 10; ++j)
					for (int c = 0; c < 3; ++c)
					{
						const double want = j < 4 ? (4.0 * lam[j] - 1.0) * gl[j][c] : 4.0 * (lam[ev[j - 4][0]] * gl[ev[j - 4][1]][c] + lam[ev[j - 4][1]] * gl[ev[j - 4][0]][c]);
						if (std::fabs(g[(q * 10 + j) * 3 + c] - want) > 1e-14)
							return false;
					}
			}
			return true;
		}

		// ---- host side: which lane triple works on which (element, node) incidence, in which order ----
		// Nodes are put into groups of kNodes; node slot s of a group is served by triples s (even incidences of the node) and
		// kNodes + s (odd incidences). A warp processes a group from its first to its last step and then flushes the 15 finished
		// columns. Class 0: nodes whose column needs <= small_rows strip rows (3*deg), class 1: the rest (a launch with larger
		// strips). Within a class nodes are ordered by their number of incident elements (descending), then by node id, so that
		// the slots of a group finish together and stay local. Groups are handed out to warps in chunks of consecutive groups.
		struct Schedule
		{
			int n_groups[2] = {0, 0};
			int rows_max[2] = {0, 0};
			int n_chunks[2] = {0, 0};
			int64_t n_steps[2] = {0, 0};
			int64_t total_steps = 0;
			int64_t busy = 0;                // (element, node) incidences = non-idle (step, triple) pairs
			std::vector<int32_t> grp_info;  // [G][kNodes][4]: node (-1 = unused slot), 9*adj_off[node], 3*deg(node), 0
			std::vector<int32_t> grp_off;   // [G+1] first step of each group
			std::vector<int32_t> grp_rows;  // [G] strip rows the group needs
			std::vector<int32_t> chunk_off; // [C+1] first group of each chunk (chunks of class 0 first)
			// [total_steps][kTriples][4] words: element (kIdle = no work), then 12 bytes: byte j < NL = position k_j of local node j of
			// that element in the adjacency of the slot's node, bit 7 set when this is the FIRST contribution to the strip rows
			// of that position (the lane stores instead of adding: no zero fill); byte 10 = local index of the slot's node
			std::vector<uint32_t> inc;
		};

		// owned: optional per-node flag (multi-GPU: only the columns of owned nodes are produced); n_el counts every element
		// whose record exists on this device (own + ghost elements)
		inline Schedule build_schedule(int n_el, int NL, int n_bases, const int32_t *conn, const int32_t *adj_off, const int32_t *adj, int small_rows,
									   int chunk_steps, const uint8_t *owned = nullptr, int bucket_elements = 1 << 30)
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
				if (n_inc(b) > 0 && (owned == nullptr || owned[b]))
					order[rows_of(b) <= small_rows ? 0 : 1].push_back(b);
			// (the incidence lists are in element order: the first entry is the node's first incident element)
			auto bucket = [&](int b) { return (inc_e[size_t(cnt[size_t(b)])] / NL) / std::max(bucket_elements, 1); };
			for (int c = 0; c < 2; ++c)
				std::stable_sort(order[c].begin(), order[c].end(), [&](int a, int b) {
					const int ba = bucket(a), bb = bucket(b);
					return ba != bb ? ba < bb : n_inc(a) > n_inc(b);
				});
			// chunk size: the requested number of steps, but never so large that a launch has fewer than ~16 chunks per resident
			// warp (148 SMs x 8 warps): short schedules (a rank of an 8-GPU run) would otherwise end in an unbalanced tail
			{
				int64_t est_steps = 0;
				for (int c = 0; c < 2; ++c)
					for (int b : order[c])
						est_steps += (n_inc(b) + 1) / 2;
				est_steps /= kNodes;
				chunk_steps = int(std::max<int64_t>(8, std::min<int64_t>(chunk_steps, est_steps / (148 * 8 * 16))));
			}
			S.grp_off.push_back(0);
			S.chunk_off.push_back(0);
			std::vector<uint8_t> seen;
			for (int c = 0; c < 2; ++c)
			{
				const std::vector<int32_t> &o = order[c];
				int chunk_acc = 0;
				for (size_t first = 0; first < o.size(); first += kNodes)
				{
					const size_t last = std::min(o.size(), first + kNodes);
					int steps = 0, rows = 0;
					for (size_t t = first; t < last; ++t)
					{
						steps = std::max(steps, (n_inc(o[t]) + 1) / 2);
						rows = std::max(rows, rows_of(o[t]));
					}
					const size_t base = S.inc.size();
					S.inc.resize(base + size_t(steps) * kTriples * 4, 0u);
					for (int st = 0; st < steps; ++st)
						for (int tr = 0; tr < kTriples; ++tr)
							S.inc[base + (size_t(st) * kTriples + tr) * 4] = kIdle;
					for (int s = 0; s < kNodes; ++s)
					{
						const int b = first + s < last ? o[first + s] : -1;
						S.grp_info.push_back(b);
						S.grp_info.push_back(b >= 0 ? 9 * adj_off[b] : 0);
						S.grp_info.push_back(b >= 0 ? rows_of(b) : 0);
						S.grp_info.push_back(0);
						if (b < 0)
							continue;
						const int32_t *lo = adj + adj_off[b], *hi = adj + adj_off[b + 1];
						seen.assign(size_t(hi - lo), 0);
						for (int t = 0; t < n_inc(b); ++t) // incidence t: step t/2, triple s (t even) or kNodes + s (t odd)
						{
							const int ei = inc_e[size_t(cnt[size_t(b)]) + t];
							const int e = ei / NL, i = ei - e * NL;
							uint32_t *w = S.inc.data() + base + (size_t(t / 2) * kTriples + (t & 1) * kNodes + s) * 4;
							w[0] = uint32_t(e);
							uint8_t bytes[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
							for (int j = 0; j < NL; ++j)
							{
								const int k = int(std::lower_bound(lo, hi, conn[size_t(e) * NL + j]) - lo);
								bytes[j] = uint8_t(k | (seen[size_t(k)] ? 0 : 0x80));
								seen[size_t(k)] = 1;
							}
							bytes[10] = uint8_t(i);
							for (int k = 0; k < 12; ++k)
								w[1 + k / 4] |= uint32_t(bytes[k]) << (8 * (k % 4));
							++S.busy;
						}
					}
					S.total_steps += steps;
					S.n_steps[c] += steps;
					S.grp_off.push_back(int32_t(S.total_steps));
					S.grp_rows.push_back(rows);
					S.rows_max[c] = std::max(S.rows_max[c], rows);
					++S.n_groups[c];
					chunk_acc += steps;
					if (chunk_acc >= chunk_steps)
					{
						S.chunk_off.push_back(int32_t(S.grp_rows.size()));
						++S.n_chunks[c];
						chunk_acc = 0;
					}
				}
				if (chunk_acc > 0)
				{
					S.chunk_off.push_back(int32_t(S.grp_rows.size()));
					++S.n_chunks[c];
				}
			}
			return S;
		}
	} // namespace cl2
} // namespace pfa
