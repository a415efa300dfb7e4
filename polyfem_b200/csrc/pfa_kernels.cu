// sm_100a kernels of the assembly hot path.
//
// Reference semantics reproduced (file:line relative to /root/reference/src/polyfem/):
//   geometry precompute  assembler/ElementAssemblyValues.cpp:65-104 (J, det, J^-T)
//   NeoHookean           assembler/NeoHookeanElasticity.cpp:338-388 (energy), :453-545
//                        (gradient), :547-658 (Hessian)
//   LinearElasticity     assembler/LinearElasticity.cpp:30-63 (stiffness block), :106-134 (energy)
//   Laplacian            assembler/Laplacian.cpp:13-26
//   global loops/scatter assembler/Assembler.cpp:157-384, 495-771; utils/MatrixCache.cpp:88-113
//
// The math is the closed form of the reference's dense B^T H_F B:
//   H_e[(i,a),(j,b)] = sum_q da_q [ mu (D_i.D_j) d_ab + c1 (C D_i)_a (C D_j)_b - c2 hat(F (D_i x D_j))_ab ]
// with D = grad * J^-T, C = cof(F), c1 = (mu + lambda (1 - ln J)) / J^2, c2 = (lambda ln J - mu) / J.
#include "pfa_internal.h"

#include <cstdio>

namespace pfa
{
	namespace
	{
		
		__device__ __forceinline__ double det3(const double *F)
		{
			return F[0] * (F[4] * F[8] - F[5] * F[7]) - F[1] * (F[3] * F[8] - F[5] * F[6]) + F[2] * (F[3] * F[7] - F[4] * F[6]);
		}

		// cofactor matrix C = dJ/dF (row-major), columns are cross products of the columns of F
		__device__ __forceinline__ void cofactor3(const double *F, double *C)
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

		// ------------------------------------------------------------------------------------
		// once-per-mesh geometry precompute for affine (P1-geometry) tets: one thread per element
		// ------------------------------------------------------------------------------------
		__global__ void geometry_precompute_kernel(const double *__restrict__ vertices, int n_el, double *__restrict__ jit, double *__restrict__ detj)
		{
			const int e = blockIdx.x * blockDim.x + threadIdx.x;
			if (e >= n_el)
				return;
			const double *v = vertices + size_t(e) * 12;
			// rows of J are v1-v0, v2-v0, v3-v0 (ElementAssemblyValues.cpp:81-94 with P1 gradients)
			double J[9];
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					J[r * 3 + c] = v[(r + 1) * 3 + c] - v[c];
			const double det = det3(J);
			double C[9];
			cofactor3(J, C);
			// J^-1 = C^T / det  =>  J^-T = C / det
			const double inv = 1.0 / det;
			for (int k = 0; k < 9; ++k)
				jit[size_t(e) * 9 + k] = C[k] * inv;
			detj[e] = det;
		}

		// scalar CSC arrays from the node-block pattern: column (b,n) lists rows (a,m), a in adj(b)
		__global__ void expand_inner_kernel(const int32_t *__restrict__ adj_off, const int32_t *__restrict__ adj, int n_bases, int size, int32_t *__restrict__ outer, int32_t *__restrict__ inner)
		{
			const int b = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32; // one warp per node column-block
			const int lane = threadIdx.x & 31;
			if (b >= n_bases)
				return;
			const int off = adj_off[b], deg = adj_off[b + 1] - off;
			const int64_t base = int64_t(off) * size * size;
			for (int n = lane; n < size; n += 32)
				outer[b * size + n] = int32_t(base + int64_t(n) * size * deg);
			if (b == n_bases - 1 && lane == 0)
				outer[n_bases * size] = int32_t(int64_t(adj_off[n_bases]) * size * size);
			const int per_col = deg * size;
			for (int t = lane; t < per_col * size; t += 32)
			{
				const int n = t / per_col, r = t - n * per_col;
				const int k = r / size, m = r - k * size;
				inner[base + t] = adj[off + k] * size + m;
			}
		}

		// ------------------------------------------------------------------------------------
		// Generic fused assembly kernel: one warp per element, any n_loc / n_qp (runtime),
		// per-warp staging in shared memory, scatter with RED.ADD.F64 into the CSC values.
		// ------------------------------------------------------------------------------------
		struct WarpLayout
		{
			int U, D, A, Q, J, DA, I; // offsets in doubles (I: start of the int region, in doubles)
			int total;                // doubles per warp
		};
		constexpr int kQRec = 30; // per-qp record: C[9] | P*da[9] | c2*da*F[9] | c1*da | mu*da | lambda*da

		__host__ __device__ inline WarpLayout warp_layout(int n_loc, int n_qp)
		{
			WarpLayout L;
			int o = 0;
			L.U = o;
			o += n_loc * 3;
			L.D = o;
			o += n_qp * n_loc * 3;
			L.A = o;
			o += n_qp * n_loc * 3;
			L.Q = o;
			o += n_qp * kQRec;
			L.J = o;
			o += n_qp * 9;
			L.DA = o;
			o += n_qp;
			L.I = o;
			o += (3 * n_loc + 1) / 2;
			L.total = o;
			return L;
		}

		template <int MAT, bool LINEAR, int kWarps>
		__global__ void __launch_bounds__(kWarps * 32) assemble_generic_kernel(const DeviceMesh m, const AssembleArgs a)
		{
			extern __shared__ double smem[];
			const int n_loc = m.n_loc, n_qp = m.n_qp;
			const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
			const WarpLayout L = warp_layout(n_loc, n_qp);

			// CTA-shared reference tables
			double *s_rg = smem;                   // [n_qp][n_loc][3]
			double *s_w = s_rg + n_qp * n_loc * 3; // [n_qp]
			for (int t = threadIdx.x; t < n_qp * n_loc * 3; t += blockDim.x)
				s_rg[t] = m.ref_grads[t];
			for (int t = threadIdx.x; t < n_qp; t += blockDim.x)
				s_w[t] = m.qweights[t];
			__syncthreads();

			double *ws = s_w + n_qp + warp * L.total;
			double *sU = ws + L.U, *sD = ws + L.D, *sA = ws + L.A, *sQ = ws + L.Q, *sJ = ws + L.J, *sDA = ws + L.DA;
			int *sG = reinterpret_cast<int *>(ws + L.I); // [n_loc] global node
			int *sOff = sG + n_loc;                      // [n_loc] adj_off of that node
			int *sDeg = sOff + n_loc;                    // [n_loc] degree of that node

			double energy_acc = 0.0;
			const int warps_total = gridDim.x * kWarps;
			const bool want_h = a.values != nullptr;
			const bool want_g = !LINEAR && a.grad != nullptr;
			const bool want_e = !LINEAR && (a.energy != nullptr || a.energy_per_el != nullptr);

			for (int e = blockIdx.x * kWarps + warp; e < m.n_el; e += warps_total)
			{
				// ---- 1. gather connectivity, displacement, geometry ----
				for (int j = lane; j < n_loc; j += 32)
				{
					const int g = m.conn[size_t(e) * n_loc + j];
					sG[j] = g;
					const int o = m.adj_off[g];
					sOff[j] = o;
					sDeg[j] = m.adj_off[g + 1] - o;
					if (!LINEAR && MAT != PFA_LAPLACIAN)
					{
						sU[j * 3 + 0] = a.x[size_t(g) * 3 + 0];
						sU[j * 3 + 1] = a.x[size_t(g) * 3 + 1];
						sU[j * 3 + 2] = a.x[size_t(g) * 3 + 2];
					}
				}
				const int gq = m.geom_per_qp ? n_qp : 1;
				for (int t = lane; t < gq * 9; t += 32)
					sJ[t] = m.jit[size_t(e) * gq * 9 + t];
				for (int q = lane; q < n_qp; q += 32)
					sDA[q] = m.geom_per_qp ? m.detj[size_t(e) * n_qp + q] : m.detj[e] * s_w[q];
				__syncwarp();

				// ---- 2. physical gradients D[q][i][:] = grad[q][i][:] * J^-T ----
				for (int t = lane; t < n_qp * n_loc; t += 32)
				{
					const int q = t / n_loc;
					const double *J = sJ + (m.geom_per_qp ? q * 9 : 0);
					const double g0 = s_rg[t * 3 + 0], g1 = s_rg[t * 3 + 1], g2 = s_rg[t * 3 + 2];
					sD[t * 3 + 0] = g0 * J[0] + g1 * J[3] + g2 * J[6];
					sD[t * 3 + 1] = g0 * J[1] + g1 * J[4] + g2 * J[7];
					sD[t * 3 + 2] = g0 * J[2] + g1 * J[5] + g2 * J[8];
				}
				__syncwarp();

				// ---- 3. per quadrature point: F, stress, Hessian coefficients ----
				double e_loc = 0.0;
				for (int q = lane; q < n_qp; q += 32)
				{
					const double da = sDA[q];
					const int ms = m.mat_stride == 1 ? 0 : q;
					double lam = 0.0, mu = 0.0;
					if (MAT != PFA_LAPLACIAN)
					{
						lam = m.lambda[size_t(e) * m.mat_stride + ms];
						mu = m.mu[size_t(e) * m.mat_stride + ms];
					}
					double *rec = sQ + q * kQRec;
					rec[28] = mu * da;
					rec[29] = lam * da;
					if (!LINEAR && MAT != PFA_LAPLACIAN)
					{
						double F[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
						for (int i = 0; i < n_loc; ++i)
						{
							const double *Di = sD + (q * n_loc + i) * 3;
							const double u0 = sU[i * 3 + 0], u1 = sU[i * 3 + 1], u2 = sU[i * 3 + 2];
							F[0] += u0 * Di[0];
							F[1] += u0 * Di[1];
							F[2] += u0 * Di[2];
							F[3] += u1 * Di[0];
							F[4] += u1 * Di[1];
							F[5] += u1 * Di[2];
							F[6] += u2 * Di[0];
							F[7] += u2 * Di[1];
							F[8] += u2 * Di[2];
						}
						if (MAT == PFA_NEOHOOKEAN)
						{
							F[0] += 1.0;
							F[4] += 1.0;
							F[8] += 1.0;
							const double J = det3(F);
							const double lJ = log(J); // NaN for J <= 0, propagates like the reference
							double C[9];
							cofactor3(F, C);
							const double invJ = 1.0 / J;
							const double pc = (lam * lJ - mu) * invJ; // P = mu F + pc C
							double sq = 0.0;
							for (int k = 0; k < 9; ++k)
							{
								sq += F[k] * F[k];
								rec[k] = C[k];
								rec[9 + k] = (mu * F[k] + pc * C[k]) * da;
								rec[18 + k] = pc * da * F[k]; // c2 * da * F
							}
							rec[27] = (mu + lam * (1.0 - lJ)) * invJ * invJ * da; // c1 * da
							e_loc += (0.5 * mu * (sq - 3.0 - 2.0 * lJ) + 0.5 * lam * lJ * lJ) * da;
						}
						else // LinearElasticity: F holds grad u
						{
							const double tr = F[0] + F[4] + F[8];
							double eps[9], tr2 = 0.0;
							for (int r = 0; r < 3; ++r)
								for (int c = 0; c < 3; ++c)
								{
									eps[r * 3 + c] = 0.5 * (F[r * 3 + c] + F[c * 3 + r]);
									tr2 += eps[r * 3 + c] * eps[r * 3 + c];
								}
							for (int k = 0; k < 9; ++k)
								rec[9 + k] = 2.0 * mu * eps[k] * da;
							rec[9 + 0] += lam * tr * da;
							rec[9 + 4] += lam * tr * da;
							rec[9 + 8] += lam * tr * da;
							e_loc += (mu * tr2 + 0.5 * lam * tr * tr) * da;
						}
					}
				}
				__syncwarp();

				// ---- 4. energy, gradient, A = C D ----
				if (want_e)
				{
					for (int o = 16; o > 0; o >>= 1)
						e_loc += __shfl_xor_sync(0xffffffffu, e_loc, o);
					energy_acc += e_loc; // identical on all lanes
					if (a.energy_per_el != nullptr && lane == 0)
						a.energy_per_el[e] = e_loc;
				}
				if (want_g)
				{
					for (int t = lane; t < n_loc * 3; t += 32)
					{
						const int i = t / 3, c = t - i * 3;
						double g = 0.0;
						for (int q = 0; q < n_qp; ++q)
						{
							const double *Di = sD + (q * n_loc + i) * 3;
							const double *P = sQ + q * kQRec + 9 + c * 3;
							g += Di[0] * P[0] + Di[1] * P[1] + Di[2] * P[2];
						}
						atomicAdd(a.grad + size_t(sG[i]) * 3 + c, g);
					}
				}
				if (want_h && MAT == PFA_NEOHOOKEAN && !LINEAR)
				{
					for (int t = lane; t < n_qp * n_loc; t += 32)
					{
						const int q = t / n_loc;
						const double *C = sQ + q * kQRec;
						const double *Di = sD + t * 3;
						sA[t * 3 + 0] = C[0] * Di[0] + C[1] * Di[1] + C[2] * Di[2];
						sA[t * 3 + 1] = C[3] * Di[0] + C[4] * Di[1] + C[5] * Di[2];
						sA[t * 3 + 2] = C[6] * Di[0] + C[7] * Di[1] + C[8] * Di[2];
					}
				}
				__syncwarp();

				// ---- 5. local Hessian / stiffness blocks and scatter ----
				if (want_h)
				{
					for (int b = lane; b < n_loc * n_loc; b += 32)
					{
						const int i = b / n_loc, j = b - i * n_loc;
						const int slot = m.slot[size_t(e) * n_loc * n_loc + b];
						if (MAT == PFA_LAPLACIAN)
						{
							double s = 0.0;
							for (int q = 0; q < n_qp; ++q)
							{
								const double *Di = sD + (q * n_loc + i) * 3, *Dj = sD + (q * n_loc + j) * 3;
								s += (Di[0] * Dj[0] + Di[1] * Dj[1] + Di[2] * Dj[2]) * sDA[q];
							}
							atomicAdd(a.values + slot, s);
							continue;
						}
						double blk[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; // blk[a*3+b] = H[(i,a),(j,b)]
						if (MAT == PFA_NEOHOOKEAN && !LINEAR)
						{
							double s = 0.0, W0 = 0.0, W1 = 0.0, W2 = 0.0;
							for (int q = 0; q < n_qp; ++q)
							{
								const double *rec = sQ + q * kQRec;
								const double *Di = sD + (q * n_loc + i) * 3, *Dj = sD + (q * n_loc + j) * 3;
								const double *Ai = sA + (q * n_loc + i) * 3, *Aj = sA + (q * n_loc + j) * 3;
								s += rec[28] * (Di[0] * Dj[0] + Di[1] * Dj[1] + Di[2] * Dj[2]);
								const double c1 = rec[27];
								const double a0 = c1 * Ai[0], a1 = c1 * Ai[1], a2 = c1 * Ai[2];
								blk[0] += a0 * Aj[0];
								blk[1] += a0 * Aj[1];
								blk[2] += a0 * Aj[2];
								blk[3] += a1 * Aj[0];
								blk[4] += a1 * Aj[1];
								blk[5] += a1 * Aj[2];
								blk[6] += a2 * Aj[0];
								blk[7] += a2 * Aj[1];
								blk[8] += a2 * Aj[2];
								const double z0 = Di[1] * Dj[2] - Di[2] * Dj[1];
								const double z1 = Di[2] * Dj[0] - Di[0] * Dj[2];
								const double z2 = Di[0] * Dj[1] - Di[1] * Dj[0];
								const double *Fs = rec + 18; // c2 * da * F
								W0 += Fs[0] * z0 + Fs[1] * z1 + Fs[2] * z2;
								W1 += Fs[3] * z0 + Fs[4] * z1 + Fs[5] * z2;
								W2 += Fs[6] * z0 + Fs[7] * z1 + Fs[8] * z2;
							}
							blk[0] += s;
							blk[4] += s;
							blk[8] += s;
							// - hat(W)
							blk[1] += W2;
							blk[2] -= W1;
							blk[3] -= W2;
							blk[5] += W0;
							blk[6] += W1;
							blk[7] -= W0;
						}
						else // LinearElasticity stiffness block (LinearElasticity.cpp:40-60)
						{
							for (int q = 0; q < n_qp; ++q)
							{
								const double *rec = sQ + q * kQRec;
								const double *Di = sD + (q * n_loc + i) * 3, *Dj = sD + (q * n_loc + j) * 3;
								const double mu = rec[28], lam = rec[29];
								const double dot = Di[0] * Dj[0] + Di[1] * Dj[1] + Di[2] * Dj[2];
								for (int r = 0; r < 3; ++r)
									for (int c = 0; c < 3; ++c)
										blk[r * 3 + c] += mu * (Di[c] * Dj[r]) + lam * (Di[r] * Dj[c]);
								blk[0] += mu * dot;
								blk[4] += mu * dot;
								blk[8] += mu * dot;
							}
						}
						// values index of (row (g_i,m), col (g_j,n)) = 9*off_j + n*3*deg_j + 3*k + m
						const int off = sOff[j], deg = sDeg[j];
						double *dst = a.values + (size_t(off) * 9 + size_t(slot - off) * 3);
						for (int n = 0; n < 3; ++n)
							for (int mm = 0; mm < 3; ++mm)
								atomicAdd(dst + size_t(n) * 3 * deg + mm, blk[mm * 3 + n]);
					}
				}
				__syncwarp();
			}

			if (want_e && a.energy != nullptr)
			{
				__shared__ double s_e[kWarps];
				if (lane == 0)
					s_e[warp] = energy_acc;
				__syncthreads();
				if (threadIdx.x == 0)
				{
					double t = 0.0;
					for (int w = 0; w < kWarps; ++w)
						t += s_e[w];
					atomicAdd(a.energy, t);
				}
			}
		}

		// ------------------------------------------------------------------------------------
		// NeoHookean kernel for n_qp in {1, 4} (P1, P2 tets): "row-lane" mapping.
		//  phase 1  lane <-> (element, quadrature point): 32/NQ elements per warp batch. Gathers
		//           x, builds F, stress and the Hessian coefficients in registers, leaves per
		//           (element, qp) records {D_i, C D_i, c2 da F, mu da, c1 da} in shared memory;
		//           gradient and energy are reduced over the qp lanes with shuffles.
		//  phase 2  lane <-> (local node i, component m) of one element: the lane keeps its row's
		//           data in registers, walks the column nodes j (records broadcast from shared
		//           memory) and produces H[(i,m),(j,0..2)]. The three m-lanes of a node write three
		//           consecutive doubles and the NL nodes of the element hit one CSC column segment
		//           per RED instruction, which minimises L2 sector operations.
		// ------------------------------------------------------------------------------------
#ifndef PFA_RL_PAD
#define PFA_RL_PAD 14
#endif
#ifndef PFA_RL_WARPS_P2
#define PFA_RL_WARPS_P2 4
#endif
		template <int NL, int NQ>
		struct RowLane
		{
			static constexpr int EB = 32 / NQ;      // elements per phase-1 batch
			static constexpr int REC = NL * 6 + PFA_RL_PAD; // doubles per (element, qp) record; (REC mod 16) = 6/10 keeps the
			                                        // lane-strided phase-1 stores at 2-way bank conflicts, 16-byte aligned
			static constexpr int ROWL = 3 * NL;     // phase-2 lanes per element
			static constexpr int EPW = 32 / ROWL;   // elements per phase-2 round
			static constexpr int WARP_DOUBLES = EB * NQ * REC;
			static constexpr int WARP_INTS = EB * NL * 3 + EB * NL * NL; // g, off, deg per node + slot per pair
			static size_t smem_bytes(int warps)
			{
				return sizeof(double) * (size_t(NQ) * NL * 3 + ((NQ + 1) & ~1) + size_t(warps) * WARP_DOUBLES) + sizeof(int) * size_t(warps) * WARP_INTS;
			}
		};

		template <int NL, int NQ, int WARPS>
		__global__ void __launch_bounds__(WARPS * 32) assemble_nh_rowlane_kernel(const DeviceMesh m, const AssembleArgs a)
		{
			using RL = RowLane<NL, NQ>;
			constexpr int EB = RL::EB, REC = RL::REC, ROWL = RL::ROWL, EPW = RL::EPW;
			extern __shared__ double smem[];
			const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
			double *s_rg = smem;               // [NQ][NL][3]
			double *s_w = s_rg + NQ * NL * 3;  // [NQ]
			double *s_rec = s_w + ((NQ + 1) & ~1) + warp * RL::WARP_DOUBLES;
			int *s_int = reinterpret_cast<int *>(s_w + ((NQ + 1) & ~1) + WARPS * RL::WARP_DOUBLES) + warp * RL::WARP_INTS;
			int *sG = s_int, *sOff = s_int + EB * NL, *sDeg = s_int + 2 * EB * NL, *sSlot = s_int + 3 * EB * NL;
			for (int t = threadIdx.x; t < NQ * NL * 3; t += WARPS * 32)
				s_rg[t] = m.ref_grads[t];
			for (int t = threadIdx.x; t < NQ; t += WARPS * 32)
				s_w[t] = m.qweights[t];
			__syncthreads();

			const bool want_h = a.values != nullptr;
			const bool want_g = a.grad != nullptr;
			const bool want_e = a.energy != nullptr || a.energy_per_el != nullptr;
			double energy_acc = 0.0;
			const int el = lane / NQ, q = lane % NQ;

			// batches are handed out dynamically (one atomic per warp and batch) so that SMs slowed
			// down by L2 contention do not leave a tail
			for (;;)
			{
				int batch = 0;
				if (lane == 0)
					batch = atomicAdd(a.work_counter, EB);
				batch = __shfl_sync(0xffffffffu, batch, 0);
				if (batch >= m.n_el)
					break;
				// ---- connectivity and pattern offsets of the batch ----
				for (int t = lane; t < EB * NL; t += 32)
				{
					const int ee = batch + t / NL;
					if (ee < m.n_el)
					{
						const int g = m.conn[size_t(ee) * NL + (t % NL)];
						const int o = m.adj_off[g];
						sG[t] = g;
						sOff[t] = o;
						sDeg[t] = m.adj_off[g + 1] - o;
					}
				}
				if (want_h)
				{
					const int n_batch = min(EB, m.n_el - batch);
					const int32_t *src = m.slot + size_t(batch) * NL * NL;
					for (int t = lane; t < n_batch * NL * NL; t += 32)
						sSlot[t] = src[t];
				}
				__syncwarp();

				// ---- phase 1 ----
				const int e = batch + el;
				const bool valid = e < m.n_el;
				double G[NL * 3];
#pragma unroll
				for (int t = 0; t < NL * 3; ++t)
					G[t] = 0.0;
				double e_q = 0.0;
				if (valid)
				{
					double *rec = s_rec + (el * NQ + q) * REC;
					const size_t gi = m.geom_per_qp ? size_t(e) * NQ + q : size_t(e);
					double J[9];
#pragma unroll
					for (int k = 0; k < 9; ++k)
						J[k] = m.jit[gi * 9 + k];
					const double da = m.geom_per_qp ? m.detj[gi] : m.detj[e] * s_w[q];
					const size_t mi = size_t(e) * m.mat_stride + (m.mat_stride == 1 ? 0 : q);
					const double lam = m.lambda[mi], mu = m.mu[mi];
					double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
					for (int i = 0; i < NL; ++i)
					{
						const int g = sG[el * NL + i];
						const double u0 = a.x[size_t(g) * 3 + 0], u1 = a.x[size_t(g) * 3 + 1], u2 = a.x[size_t(g) * 3 + 2];
						const double *r = s_rg + (q * NL + i) * 3;
						const double d0 = r[0] * J[0] + r[1] * J[3] + r[2] * J[6];
						const double d1 = r[0] * J[1] + r[1] * J[4] + r[2] * J[7];
						const double d2 = r[0] * J[2] + r[1] * J[5] + r[2] * J[8];
						rec[i * 6 + 0] = d0;
						rec[i * 6 + 1] = d1;
						rec[i * 6 + 2] = d2;
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
					double P[9], sq = 0.0;
#pragma unroll
					for (int k = 0; k < 9; ++k)
					{
						sq += F[k] * F[k];
						P[k] = (mu * F[k] + pc * C[k]) * da;
						rec[NL * 6 + k] = pc * da * F[k];
					}
					rec[NL * 6 + 9] = mu * da;
					rec[NL * 6 + 10] = (mu + lam * (1.0 - lJ)) * invJ * invJ * da;
					e_q = (0.5 * mu * (sq - 3.0 - 2.0 * lJ) + 0.5 * lam * lJ * lJ) * da;
#pragma unroll
					for (int i = 0; i < NL; ++i)
					{
						const double d0 = rec[i * 6 + 0], d1 = rec[i * 6 + 1], d2 = rec[i * 6 + 2];
						rec[i * 6 + 3] = C[0] * d0 + C[1] * d1 + C[2] * d2;
						rec[i * 6 + 4] = C[3] * d0 + C[4] * d1 + C[5] * d2;
						rec[i * 6 + 5] = C[6] * d0 + C[7] * d1 + C[8] * d2;
						G[i * 3 + 0] = d0 * P[0] + d1 * P[1] + d2 * P[2];
						G[i * 3 + 1] = d0 * P[3] + d1 * P[4] + d2 * P[5];
						G[i * 3 + 2] = d0 * P[6] + d1 * P[7] + d2 * P[8];
					}
				}
				if (want_e)
				{
#pragma unroll
					for (int k = 1; k < NQ; k <<= 1)
						e_q += __shfl_xor_sync(0xffffffffu, e_q, k);
					if (valid && q == 0)
					{
						energy_acc += e_q;
						if (a.energy_per_el != nullptr)
							a.energy_per_el[e] = e_q;
					}
				}
				if (want_g)
				{
#pragma unroll
					for (int t = 0; t < NL * 3; ++t)
					{
#pragma unroll
						for (int k = 1; k < NQ; k <<= 1)
							G[t] += __shfl_xor_sync(0xffffffffu, G[t], k);
					}
#pragma unroll
					for (int i = 0; i < NL; ++i)
						if (valid && (i % NQ) == q)
						{
							double *dst = a.grad + size_t(sG[el * NL + i]) * 3;
							atomicAdd(dst + 0, G[i * 3 + 0]);
							atomicAdd(dst + 1, G[i * 3 + 1]);
							atomicAdd(dst + 2, G[i * 3 + 2]);
						}
				}
				__syncwarp();

				// ---- phase 2 ----
				if (want_h)
				{
					const int sub = lane / ROWL, r = lane % ROWL;
					const int i = r / 3, mm = r % 3;
					const int ra = (mm + 1) % 3, rb = (mm + 2) % 3;
					for (int eb = 0; eb < EB; eb += EPW)
					{
						const int el2 = eb + sub;
						const int e2 = batch + el2;
						if (lane < EPW * ROWL && el2 < EB && e2 < m.n_el)
						{
							// row-side registers: mu da D_i, c1 da (C D_i)_m, rows (m+1)%3,(m+2)%3 of c2 da F x D_i
							double Dp[NQ][3], cA[NQ], Ma[NQ][3], Mb[NQ][3];
#pragma unroll
							for (int qq = 0; qq < NQ; ++qq)
							{
								const double *rec = s_rec + (el2 * NQ + qq) * REC;
								const double d0 = rec[i * 6 + 0], d1 = rec[i * 6 + 1], d2 = rec[i * 6 + 2];
								const double mu_da = rec[NL * 6 + 9], c1_da = rec[NL * 6 + 10];
								const double *fa = rec + NL * 6 + ra * 3, *fb = rec + NL * 6 + rb * 3;
								Dp[qq][0] = mu_da * d0;
								Dp[qq][1] = mu_da * d1;
								Dp[qq][2] = mu_da * d2;
								cA[qq] = c1_da * rec[i * 6 + 3 + mm];
								Ma[qq][0] = fa[1] * d2 - fa[2] * d1;
								Ma[qq][1] = fa[2] * d0 - fa[0] * d2;
								Ma[qq][2] = fa[0] * d1 - fa[1] * d0;
								Mb[qq][0] = fb[1] * d2 - fb[2] * d1;
								Mb[qq][1] = fb[2] * d0 - fb[0] * d2;
								Mb[qq][2] = fb[0] * d1 - fb[1] * d0;
							}
#ifndef PFA_RL_UNROLL_J
#define PFA_RL_UNROLL_J 10
#endif
							const int *sl = sSlot + el2 * NL * NL + i * NL;
							constexpr int kUnrollJ = PFA_RL_UNROLL_J;
#pragma unroll kUnrollJ
							for (int j = 0; j < NL; ++j)
							{
								double s = 0.0, R0 = 0.0, R1 = 0.0, R2 = 0.0, wa = 0.0, wb = 0.0;
#pragma unroll
								for (int qq = 0; qq < NQ; ++qq)
								{
									const double2 *nj = reinterpret_cast<const double2 *>(s_rec + (el2 * NQ + qq) * REC + j * 6);
									const double2 v0 = nj[0], v1 = nj[1], v2 = nj[2]; // D0 D1 | D2 A0 | A1 A2
									s = fma(Dp[qq][0], v0.x, s);
									s = fma(Dp[qq][1], v0.y, s);
									s = fma(Dp[qq][2], v1.x, s);
									R0 = fma(cA[qq], v1.y, R0);
									R1 = fma(cA[qq], v2.x, R1);
									R2 = fma(cA[qq], v2.y, R2);
									wa = fma(Ma[qq][0], v0.x, wa);
									wa = fma(Ma[qq][1], v0.y, wa);
									wa = fma(Ma[qq][2], v1.x, wa);
									wb = fma(Mb[qq][0], v0.x, wb);
									wb = fma(Mb[qq][1], v0.y, wb);
									wb = fma(Mb[qq][2], v1.x, wb);
								}
								// row m of  R + s I - hat(w):  out[m] += s, out[(m+1)%3] += w_{(m+2)%3}, out[(m+2)%3] -= w_{(m+1)%3}
								if (mm == 0)
								{
									R0 += s;
									R1 += wb;
									R2 -= wa;
								}
								else if (mm == 1)
								{
									R1 += s;
									R2 += wb;
									R0 -= wa;
								}
								else
								{
									R2 += s;
									R0 += wb;
									R1 -= wa;
								}
								const int off = sOff[el2 * NL + j], deg = sDeg[el2 * NL + j];
								double *dst = a.values + (size_t(off) * 9 + size_t(sl[j] - off) * 3 + mm);
								atomicAdd(dst, R0);
								atomicAdd(dst + size_t(3) * deg, R1);
								atomicAdd(dst + size_t(6) * deg, R2);
							}
						}
					}
				}
				__syncwarp();
			}

			if (want_e && a.energy != nullptr)
			{
				__shared__ double s_e[WARPS];
#pragma unroll
				for (int o = 16; o > 0; o >>= 1)
					energy_acc += __shfl_xor_sync(0xffffffffu, energy_acc, o);
				if (lane == 0)
					s_e[warp] = energy_acc;
				__syncthreads();
				if (threadIdx.x == 0)
				{
					double t = 0.0;
					for (int w = 0; w < WARPS; ++w)
						t += s_e[w];
					atomicAdd(a.energy, t);
				}
			}
		}

		template <int NL, int NQ, int WARPS>
		cudaError_t launch_rowlane(const DeviceMesh &m, const AssembleArgs &a, int sm_count, cudaStream_t st)
		{
			using RL = RowLane<NL, NQ>;
			const size_t smem = RL::smem_bytes(WARPS);
			auto kern = assemble_nh_rowlane_kernel<NL, NQ, WARPS>;
			cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
			if (err != cudaSuccess)
				return err;
			int per_sm = 1;
			err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem);
			if (err != cudaSuccess)
				return err;
			if (per_sm < 1)
				per_sm = 1;
			const int64_t need = (int64_t(m.n_el) + WARPS * RL::EB - 1) / (WARPS * RL::EB);
			const int grid = int(std::max<int64_t>(1, std::min<int64_t>(need, int64_t(sm_count) * per_sm)));
			kern<<<grid, WARPS * 32, smem, st>>>(m, a);
			return cudaGetLastError();
		}

		constexpr size_t kMaxSmem = 227 * 1024;

		size_t generic_smem_bytes(int n_loc, int n_qp, int warps)
		{
			const WarpLayout L = warp_layout(n_loc, n_qp);
			return sizeof(double) * (size_t(n_qp) * n_loc * 3 + n_qp + size_t(warps) * L.total);
		}

		// warps per CTA: the largest of 8/4/2/1 whose staging fits in shared memory
		int pick_warps(int n_loc, int n_qp)
		{
			for (int w = 8; w >= 1; w >>= 1)
				if (generic_smem_bytes(n_loc, n_qp, w) <= kMaxSmem)
					return w;
			return 0;
		}

		template <int MAT, bool LINEAR, int kWarps>
		cudaError_t launch_generic_w(const DeviceMesh &m, const AssembleArgs &a, int sm_count, cudaStream_t st)
		{
			const size_t smem = generic_smem_bytes(m.n_loc, m.n_qp, kWarps);
			auto kern = assemble_generic_kernel<MAT, LINEAR, kWarps>;
			cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
			if (err != cudaSuccess)
				return err;
			int per_sm = 1;
			err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarps * 32, smem);
			if (err != cudaSuccess)
				return err;
			if (per_sm < 1)
				per_sm = 1;
			const int64_t need = (int64_t(m.n_el) + kWarps - 1) / kWarps;
			const int grid = int(std::max<int64_t>(1, std::min<int64_t>(need, int64_t(sm_count) * per_sm)));
			kern<<<grid, kWarps * 32, smem, st>>>(m, a);
			return cudaGetLastError();
		}

		template <int MAT, bool LINEAR>
		cudaError_t launch_generic(const DeviceMesh &m, const AssembleArgs &a, int sm_count, cudaStream_t st)
		{
			switch (pick_warps(m.n_loc, m.n_qp))
			{
			case 8:
				return launch_generic_w<MAT, LINEAR, 8>(m, a, sm_count, st);
			case 4:
				return launch_generic_w<MAT, LINEAR, 4>(m, a, sm_count, st);
			case 2:
				return launch_generic_w<MAT, LINEAR, 2>(m, a, sm_count, st);
			case 1:
				return launch_generic_w<MAT, LINEAR, 1>(m, a, sm_count, st);
			default:
				return cudaErrorInvalidConfiguration;
			}
		}
	} // namespace

	bool assemble_supported(const DeviceMesh &m)
	{
		return pick_warps(m.n_loc, m.n_qp) > 0;
	}

	cudaError_t launch_geometry_precompute(const double *vertices_dev, int n_el, double *jit, double *detj, cudaStream_t st)
	{
		const int threads = 256;
		geometry_precompute_kernel<<<(n_el + threads - 1) / threads, threads, 0, st>>>(vertices_dev, n_el, jit, detj);
		return cudaGetLastError();
	}

	cudaError_t launch_expand_inner(const DeviceMesh &m, int32_t *outer, int32_t *inner, cudaStream_t st)
	{
		const int warps = 8;
		expand_inner_kernel<<<(m.n_bases + warps - 1) / warps, warps * 32, 0, st>>>(m.adj_off, m.adj, m.n_bases, m.size, outer, inner);
		return cudaGetLastError();
	}

	cudaError_t launch_assemble(const DeviceMesh &m, const AssembleArgs &a, bool linear, int sm_count, cudaStream_t st, const char **kernel_name)
	{
		if (kernel_name)
			*kernel_name = "assemble_generic_kernel";
		switch (m.material)
		{
		case PFA_NEOHOOKEAN:
			if (m.n_loc == 10 && m.n_qp == 4)
			{
				if (kernel_name)
					*kernel_name = "assemble_nh_rowlane_kernel<10,4>";
				return launch_rowlane<10, 4, PFA_RL_WARPS_P2>(m, a, sm_count, st);
			}
			if (m.n_loc == 4 && m.n_qp == 1)
			{
				if (kernel_name)
					*kernel_name = "assemble_nh_rowlane_kernel<4,1>";
				return launch_rowlane<4, 1, 8>(m, a, sm_count, st);
			}
			return launch_generic<PFA_NEOHOOKEAN, false>(m, a, sm_count, st);
		case PFA_LINEAR_ELASTICITY:
			return linear ? launch_generic<PFA_LINEAR_ELASTICITY, true>(m, a, sm_count, st)
						  : launch_generic<PFA_LINEAR_ELASTICITY, false>(m, a, sm_count, st);
		case PFA_LAPLACIAN:
			return launch_generic<PFA_LAPLACIAN, true>(m, a, sm_count, st);
		default:
			return cudaErrorInvalidValue;
		}
	}
} // namespace pfa
