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
#include <cstring>
#include <mutex>
#include <vector>

#ifndef PFA_RL_UNROLL_Q
#define PFA_RL_UNROLL_Q 1 // qp loop of phase 2: 1 keeps register pressure low (c_refgrad is read with LDC)
#endif
#ifndef PFA_RL_WARPS_P2
#define PFA_RL_WARPS_P2 6
#endif
#ifndef PFA_RL_MINB_P2
#define PFA_RL_MINB_P2 2
#endif
// timing experiments (tools/kbench.py, never in the product build): drop the column FMAs, drop
// the scatter, or scatter with plain stores
#ifndef PFA_RL_DENSE_COLUMNS // 1: never use the structured (P2S) column loop
#define PFA_RL_DENSE_COLUMNS 0
#endif
#ifndef PFA_NO_AFFINE_LINEAR // 1: never use the reference-moment kernel for linear stiffness on affine elements
#define PFA_NO_AFFINE_LINEAR 0
#endif
#ifndef PFA_NO_LAPLACIAN_TILE // 1: Laplacian always through the generic kernel
#define PFA_NO_LAPLACIAN_TILE 0
#endif
#ifndef PFA_EXP_MODE // bit 0: no phase-1 math, bit 1: no phase-2 math, bit 2: no scatter, bit 3: plain stores, bit 4: lean scatter
#define PFA_EXP_MODE 0
#endif

namespace pfa
{
	namespace
	{
		
		__device__ __forceinline__ double det3(const double *F)
		{
			return F[0] * (F[4] * F[8] - F[5] * F[7]) - F[1] * (F[3] * F[8] - F[5] * F[6]) + F[2] * (F[3] * F[7] - F[4] * F[6]);
		}

		// fire-and-forget reduction (REDG): spelled in PTX so that fences elsewhere in a kernel do not
		// make the compiler fall back to the returning form (ATOMG)
		__device__ __forceinline__ void red_add(double *p, double v)
		{
			asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
		}

		// p ? a : b as SEL instructions (a ternary on doubles may be compiled into a divergent branch)
		__device__ __forceinline__ double sel(int p, double a, double b)
		{
			double r;
			asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; selp.f64 %0, %1, %2, q; }" : "=d"(r) : "d"(a), "d"(b), "r"(p));
			return r;
		}

		// Scatter of the three column components of one (row lane, column node) pair. `stride_word` is
		// DeviceMesh::cstride: bits 0..27 the distance between the scalar columns of the column node,
		// bits 28..30 which of its three columns exist (all of them unless the tables were built for a
		// Dirichlet-reduced matrix, pfa_grad_hess_reduced); row_kept says whether this lane's row exists.
		__device__ __forceinline__ void scatter3(double *dst, unsigned stride_word, bool row_kept, double o0, double o1, double o2)
		{
			const unsigned mj = stride_word >> 28;
			const size_t cs = size_t(stride_word & 0x0fffffffu);
			if (row_kept && (mj & 1u))
				red_add(dst, o0);
			if (row_kept && (mj & 2u))
				red_add(dst + (mj & 1u) * cs, o1);
			if (row_kept && (mj & 4u))
				red_add(dst + __popc(mj & 3u) * cs, o2);
		}

		// gradient entry of dof `dof` (full numbering); with a Dirichlet map the entry goes to its
		// reduced position or is dropped
		__device__ __forceinline__ void add_gradient(const AssembleArgs &a, size_t dof, double g)
		{
			if (a.old_to_new != nullptr)
			{
				const int32_t r = a.old_to_new[dof];
				if (r >= 0)
					atomicAdd(a.grad + r, g * a.scale);
			}
			else
				atomicAdd(a.grad + dof, g * a.scale);
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

		__global__ void gather_rows_kernel(const double *__restrict__ src, const int32_t *__restrict__ perm, int n, int stride, double *__restrict__ dst)
		{
			const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
			if (t < int64_t(n) * stride)
				dst[t] = src[int64_t(perm[t / stride]) * stride + t % stride];
		}

		// scalar CSC arrays from the node-block pattern: column (b,n) lists rows (a,m), a in adj(b)
		template <typename Index>
		__global__ void expand_inner_kernel(const int32_t *__restrict__ adj_off, const int32_t *__restrict__ adj, int n_bases, int size, Index *__restrict__ outer, Index *__restrict__ inner)
		{
			const int b = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32; // one warp per node column-block
			const int lane = threadIdx.x & 31;
			if (b >= n_bases)
				return;
			const int off = adj_off[b], deg = adj_off[b + 1] - off;
			const int64_t base = int64_t(off) * size * size;
			for (int n = lane; n < size; n += 32)
				outer[b * size + n] = Index(base + int64_t(n) * size * deg);
			if (b == n_bases - 1 && lane == 0)
				outer[n_bases * size] = Index(int64_t(adj_off[n_bases]) * size * size);
			const int per_col = deg * size;
			for (int t = lane; t < per_col * size; t += 32)
			{
				const int n = t / per_col, r = t - n * per_col;
				const int k = r / size, m = r - k * size;
				inner[base + t] = adj[off + k] * size + m;
			}
		}

		// Signed SVD of a 3 x 3 matrix the way utils/svd.hpp (fastSVD3d behind AutoFlipSVD) defines it: eigenvectors of A^T A (cyclic
		// Jacobi here), eigenvalues in decreasing order, sigma = sqrt(max(lambda, 0)) with sigma_2 negated when det A < 0, V a
		// rotation, u_0 = A v_0 normalised, u_1 = the normalised part of A v_1 orthogonal to u_0, u_2 = u_0 x u_1. Row-major
		// matrices, singular vectors in the columns.
		__device__ inline void svd3_signed(const double *A, double *U, double *sig, double *V)
		{
			double C[9], W[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					C[r * 3 + c] = A[0 + r] * A[0 + c] + A[3 + r] * A[3 + c] + A[6 + r] * A[6 + c];
			for (int sweep = 0; sweep < 30; ++sweep)
			{
				const double off = C[1] * C[1] + C[2] * C[2] + C[5] * C[5], diag = C[0] * C[0] + C[4] * C[4] + C[8] * C[8];
				if (off <= 1e-32 * diag || off == 0.0)
					break;
				for (int p = 0; p < 2; ++p)
					for (int q = p + 1; q < 3; ++q)
					{
						const double apq = C[p * 3 + q];
						if (apq == 0.0)
							continue;
						const double theta = (C[q * 3 + q] - C[p * 3 + p]) / (2.0 * apq);
						const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
						const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
						for (int k = 0; k < 3; ++k)
						{
							const double akp = C[k * 3 + p], akq = C[k * 3 + q];
							C[k * 3 + p] = c * akp - sn * akq;
							C[k * 3 + q] = sn * akp + c * akq;
						}
						for (int k = 0; k < 3; ++k)
						{
							const double apk = C[p * 3 + k], aqk = C[q * 3 + k];
							C[p * 3 + k] = c * apk - sn * aqk;
							C[q * 3 + k] = sn * apk + c * aqk;
						}
						for (int k = 0; k < 3; ++k)
						{
							const double vkp = W[k * 3 + p], vkq = W[k * 3 + q];
							W[k * 3 + p] = c * vkp - sn * vkq;
							W[k * 3 + q] = sn * vkp + c * vkq;
						}
					}
			}
			// decreasing order (3-element sort of column indices)
			int o0 = 0, o1 = 1, o2 = 2;
			double w0 = C[0], w1 = C[4], w2 = C[8];
			if (w0 < w1)
			{
				const double tw = w0;
				w0 = w1;
				w1 = tw;
				const int to = o0;
				o0 = o1;
				o1 = to;
			}
			if (w1 < w2)
			{
				const double tw = w1;
				w1 = w2;
				w2 = tw;
				const int to = o1;
				o1 = o2;
				o2 = to;
			}
			if (w0 < w1)
			{
				const double tw = w0;
				w0 = w1;
				w1 = tw;
				const int to = o0;
				o0 = o1;
				o1 = to;
			}
			sig[0] = sqrt(fmax(w0, 0.0));
			sig[1] = sqrt(fmax(w1, 0.0));
			sig[2] = sqrt(fmax(w2, 0.0));
			for (int r = 0; r < 3; ++r)
			{
				V[r * 3 + 0] = W[r * 3 + o0];
				V[r * 3 + 1] = W[r * 3 + o1];
				V[r * 3 + 2] = W[r * 3 + o2];
			}
			if (det3(V) < 0)
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
			const double n0 = sqrt(u0[0] * u0[0] + u0[1] * u0[1] + u0[2] * u0[2]);
			if (n0 != 0)
				for (int r = 0; r < 3; ++r)
					u0[r] /= n0;
			else
			{
				u0[0] = 1;
				u0[1] = u0[2] = 0;
			}
			const double d01 = av1[0] * u0[0] + av1[1] * u0[1] + av1[2] * u0[2];
			for (int r = 0; r < 3; ++r)
				u1[r] = av1[r] - d01 * u0[r];
			double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
			if (n1 == 0)
			{
				// any unit vector orthogonal to u0: start from the axis u0 is least aligned with
				const int k = fabs(u0[0]) < fabs(u0[1]) ? (fabs(u0[0]) < fabs(u0[2]) ? 0 : 2) : (fabs(u0[1]) < fabs(u0[2]) ? 1 : 2);
				const double d = u0[k];
				for (int r = 0; r < 3; ++r)
					u1[r] = (r == k ? 1.0 : 0.0) - d * u0[r];
				n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
			}
			for (int r = 0; r < 3; ++r)
				u1[r] /= n1;
			const double u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
			for (int r = 0; r < 3; ++r)
			{
				U[r * 3 + 0] = u0[r];
				U[r * 3 + 1] = u1[r];
				U[r * 3 + 2] = u2[r];
			}
		}

		// ------------------------------------------------------------------------------------
		// Generic fused assembly kernel: one warp per element, any n_loc / n_qp (runtime),
		// per-warp staging in shared memory, scatter with RED.ADD.F64 into the CSC values.
		// ------------------------------------------------------------------------------------
		struct WarpLayout
		{
			int U, D, A, Q, J, DA, I; // offsets in doubles (I: start of the int region, in doubles)
			int UP = 0;               // previous local displacement (ViscousDamping)
			int H = 0, V = 0, CS = 0, PQ = 0; // project_to_psd variant only
			int total;                // doubles per warp
		};
		// per-qp record: C[9] | P*da[9] | c2*da*F[9] | c1*da | mu*da | lambda*da | (SaintVenant) mu*da*F F^T [6]
		// SaintVenant uses the slots as: F[9] | P*da[9] | S*da[9] | - | mu*da | lambda*da | mu*da*F F^T (00 01 02 11 12 22)
		constexpr int kQRec = 36;
		// MooneyRivlin: F | P da | F M | cof F | F F^T (6) | M (6) | psi_1, psi_2, psi_J, psi_1J, psi_2J, psi_JJ (x da)
		// FixedCorotational: U | P da | stiffness da (81) | V | d2E/dsigma2 (9) | pair coefficients L + R (3), L - R (3)
		__host__ __device__ constexpr int qrec_of(int material) { return material == PFA_MOONEY_RIVLIN ? 54 : (material == PFA_FIXED_COROTATIONAL ? 124 : kQRec); }

		// dimension of the local matrix as project_to_psd sees it: 3 n_loc, plus one zero row and column when that is odd
		__host__ __device__ inline int psd_dim(int n_loc) { return 3 * n_loc + ((3 * n_loc) & 1); }

		__host__ __device__ constexpr bool needs_prev(int material) { return material == PFA_VISCOUS_DAMPING; }

		__host__ __device__ inline WarpLayout warp_layout(int n_loc, int n_qp, bool psd = false, int qrec = kQRec, bool prev = false)
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
			o += n_qp * qrec;
			L.J = o;
			o += n_qp * 9;
			L.DA = o;
			o += n_qp;
			L.I = o;
			o += (3 * n_loc + 1) / 2;
			if (prev)
			{
				L.UP = o;
				o += n_loc * 3;
			}
			if (psd)
			{
				// local matrix H and eigenvector matrix V [Np][Np|1] (Np = N rounded up to even: the round-robin ordering pairs all
				// indices), rotation parameters c, s [Np/2] and pairs p, q [Np/2] ints
				const int N = psd_dim(n_loc), LD = N | 1, HALF = N / 2;
				L.H = o;
				o += N * LD;
				L.V = o;
				o += N * LD;
				L.CS = o;
				o += 2 * HALF;
				L.PQ = o;
				o += HALF + 1;
			}
			L.total = o;
			return L;
		}

		// project_to_psd for the generic kernel (ipc::project_to_psd on the local Hessian, Assembler.cpp:693-694), any
		// NLAssembler material whose 2 Np (Np|1) doubles fit the shared memory of one warp (Np = psd_dim: P1..P4 tets, Q1/Q2
		// hexes). n = 3 n_loc is the size of the matrix; when n is odd the caller's N = n + 1 and row/column n is a zero
		// padding (its rotations are identities, its eigenvalue 0 never counts as negative). The local
		// matrix sH (lower triangle mirrored, as the eigen-solver reads one triangle) is diagonalised by a parallel cyclic
		// Jacobi method (round-robin ordering: N/2 disjoint rotations per step; lanes loop over rows) and, if its smallest
		// eigenvalue is negative, rebuilt as V max(D, 0) V^T. Non-finite and already-PSD matrices are left unchanged.
		// Returns true when the matrix was changed (then sH holds the projected matrix); false: the caller keeps the original.
		__device__ inline bool psd_project_warp(double *sH, double *sV, int n, int N, int LD, int *sP, int *sQ, double *sC, double *sS, int lane)
		{
			const int HALF = N / 2;
			if (N > n)
			{
				for (int k = lane; k < N; k += 32)
					sH[n * LD + k] = sH[k * LD + n] = 0.0;
				__syncwarp();
			}
			for (int r = lane; r < N; r += 32)
				for (int c = r + 1; c < N; ++c)
					sH[r * LD + c] = sH[c * LD + r];
			__syncwarp();
			bool finite = true, nonzero = false;
			for (int r = lane; r < N; r += 32)
				for (int c = 0; c < N; ++c)
				{
					const double v = sH[r * LD + c];
					finite = finite && isfinite(v);
					nonzero = nonzero || v != 0.0;
					sV[r * LD + c] = c == r ? 1.0 : 0.0;
				}
			finite = __all_sync(0xffffffffu, finite);
			nonzero = __any_sync(0xffffffffu, nonzero);
			__syncwarp();
			if (!finite || !nonzero)
				return false;
			for (int sweep = 0; sweep < 100; ++sweep)
			{
				double off = 0.0, diag = 0.0;
				for (int r = lane; r < N; r += 32)
					for (int c = 0; c < N; ++c)
					{
						const double v = sH[r * LD + c];
						if (c == r)
							diag += v * v;
						else
							off += v * v;
					}
				for (int o = 16; o > 0; o >>= 1)
				{
					off += __shfl_xor_sync(0xffffffffu, off, o);
					diag += __shfl_xor_sync(0xffffffffu, diag, o);
				}
				if (off <= 1e-30 * diag || off == 0.0)
					break;
				for (int step = 0; step < N - 1; ++step)
				{
					for (int k = lane; k < HALF; k += 32)
					{
						int pp, qq;
						if (k == 0)
						{
							pp = step;
							qq = N - 1;
						}
						else
						{
							pp = (step + k) % (N - 1);
							qq = (step - k + (N - 1)) % (N - 1);
						}
						if (pp > qq)
						{
							const int t = pp;
							pp = qq;
							qq = t;
						}
						const double apq = sH[pp * LD + qq];
						double c = 1.0, sn = 0.0;
						if (apq != 0.0)
						{
							const double app = sH[pp * LD + pp], aqq = sH[qq * LD + qq];
							const double theta = (aqq - app) / (2.0 * apq);
							const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
							c = 1.0 / sqrt(t * t + 1.0);
							sn = t * c;
						}
						sP[k] = pp;
						sQ[k] = qq;
						sC[k] = c;
						sS[k] = sn;
					}
					__syncwarp();
					for (int r = lane; r < N; r += 32) // H <- H J and V <- V J, row by row
						for (int k = 0; k < HALF; ++k)
						{
							const int pp = sP[k], qq = sQ[k];
							const double c = sC[k], sn = sS[k];
							const double akp = sH[r * LD + pp], akq = sH[r * LD + qq];
							sH[r * LD + pp] = c * akp - sn * akq;
							sH[r * LD + qq] = sn * akp + c * akq;
							const double vkp = sV[r * LD + pp], vkq = sV[r * LD + qq];
							sV[r * LD + pp] = c * vkp - sn * vkq;
							sV[r * LD + qq] = sn * vkp + c * vkq;
						}
					__syncwarp();
					for (int cc = lane; cc < N; cc += 32) // H <- J^T H, column by column
						for (int k = 0; k < HALF; ++k)
						{
							const int pp = sP[k], qq = sQ[k];
							const double c = sC[k], sn = sS[k];
							const double apk = sH[pp * LD + cc], aqk = sH[qq * LD + cc];
							sH[pp * LD + cc] = c * apk - sn * aqk;
							sH[qq * LD + cc] = sn * apk + c * aqk;
						}
					__syncwarp();
				}
			}
			double wmin = 1e300;
			for (int r = lane; r < N; r += 32)
				wmin = fmin(wmin, sH[r * LD + r]);
			for (int o = 16; o > 0; o >>= 1)
				wmin = fmin(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
			if (!(wmin < 0.0))
				return false; // already PSD: the reference returns the matrix unchanged
			// H = V max(D, 0) V^T. The (clamped) eigenvalues stay on the diagonal while the strictly lower triangle - which carries
			// no information after convergence - is overwritten with the rebuilt entries; the diagonal follows in a second pass.
			for (int r = lane; r < N; r += 32)
				if (sH[r * LD + r] < 0.0)
					sH[r * LD + r] = 0.0;
			__syncwarp();
			for (int r = lane; r < N; r += 32)
				for (int c = 0; c < r; ++c)
				{
					double sum = 0.0;
					for (int k = 0; k < N; ++k)
						sum += sV[r * LD + k] * sH[k * LD + k] * sV[c * LD + k];
					sH[r * LD + c] = sum;
				}
			__syncwarp();
			// diagonal entries last (they hold the eigenvalues until here): each into a register, synchronise, then store
			double dnew[4]; // N <= 128
			int cnt = 0;
			for (int r = lane; r < N; r += 32)
			{
				double sum = 0.0;
				for (int k = 0; k < N; ++k)
					sum += sV[r * LD + k] * sH[k * LD + k] * sV[r * LD + k];
				dnew[cnt++] = sum;
			}
			__syncwarp();
			cnt = 0;
			for (int r = lane; r < N; r += 32)
				sH[r * LD + r] = dnew[cnt++];
			__syncwarp();
			for (int r = lane; r < N; r += 32)
				for (int c = r + 1; c < N; ++c)
					sH[r * LD + c] = sH[c * LD + r];
			__syncwarp();
			return true;
		}

		template <int MAT, bool LINEAR, int kWarps, bool PSD = false, bool TABLES_SHARED = true>
		__global__ void __launch_bounds__(kWarps * 32) assemble_generic_kernel(const DeviceMesh m, const AssembleArgs a)
		{
			extern __shared__ double smem[];
			const int n_loc = m.n_loc, n_qp = m.n_qp;
			const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
			constexpr int QR = qrec_of(MAT);
			const WarpLayout L = warp_layout(n_loc, n_qp, PSD, QR, needs_prev(MAT));

			// CTA-shared reference tables
			// (TABLES_SHARED false: read through the cache from global memory instead - the P4 projection needs the space)
			const double *s_rg = m.ref_grads; // [n_qp][n_loc][3]
			const double *s_w = m.qweights;   // [n_qp]
			double *ws = smem + warp * L.total;
			if (TABLES_SHARED)
			{
				double *t_rg = smem, *t_w = smem + n_qp * n_loc * 3;
				for (int t = threadIdx.x; t < n_qp * n_loc * 3; t += blockDim.x)
					t_rg[t] = m.ref_grads[t];
				for (int t = threadIdx.x; t < n_qp; t += blockDim.x)
					t_w[t] = m.qweights[t];
				__syncthreads();
				s_rg = t_rg;
				s_w = t_w;
				ws = t_w + n_qp + warp * L.total;
			}
			double *sU = ws + L.U, *sD = ws + L.D, *sA = ws + L.A, *sQ = ws + L.Q, *sJ = ws + L.J, *sDA = ws + L.DA;
			int *sG = reinterpret_cast<int *>(ws + L.I); // [n_loc] global node
			int *sOff = sG + n_loc;                      // [n_loc] adj_off of that node
			int *sDeg = sOff + n_loc;                    // [n_loc] degree of that node

			double energy_acc = 0.0;
			const int warps_total = gridDim.x * kWarps;
			const bool want_h = a.values != nullptr;
			const bool want_g = !LINEAR && a.grad != nullptr;
			const bool want_e = !LINEAR && (a.energy != nullptr || a.energy_per_el != nullptr);

			for (int e = a.e_begin + blockIdx.x * kWarps + warp; e < a.e_end; e += warps_total)
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
						if (needs_prev(MAT))
						{
							double *sUP = ws + L.UP;
							sUP[j * 3 + 0] = a.x_prev[size_t(g) * 3 + 0];
							sUP[j * 3 + 1] = a.x_prev[size_t(g) * 3 + 1];
							sUP[j * 3 + 2] = a.x_prev[size_t(g) * 3 + 2];
						}
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
					double *rec = sQ + q * QR;
					if (MAT != PFA_MOONEY_RIVLIN)
					{
						rec[28] = mu * da;
						rec[29] = lam * da;
					}
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
						else if (MAT == PFA_FIXED_COROTATIONAL)
						{
							// FixedCorotational.cpp:293-319 (energy), :321-377 + :678-706 (stress), :379-436 + :708-827 (stiffness). This lane:
							// SVD, energy, stress, and the coefficients of the stiffness in the singular basis; the 81 entries follow below,
							// spread over the warp.
							F[0] += 1.0;
							F[4] += 1.0;
							F[8] += 1.0;
							double Us[9], Vs[9], sg[3], G[9];
							svd3_signed(F, Us, sg, Vs);
							cofactor3(F, G);
							const double prod = sg[0] * sg[1] * sg[2], pm1 = prod - 1.0;
							const double other[3] = {sg[1] * sg[2], sg[2] * sg[0], sg[0] * sg[1]};
							double dE[3];
							for (int k = 0; k < 3; ++k)
								dE[k] = 2.0 * mu * (sg[k] - 1.0) + other[k] * lam * pm1;
							for (int r = 0; r < 3; ++r)
								for (int c = 0; c < 3; ++c)
								{
									const double R = Us[r * 3 + 0] * Vs[c * 3 + 0] + Us[r * 3 + 1] * Vs[c * 3 + 1] + Us[r * 3 + 2] * Vs[c * 3 + 2];
									rec[r * 3 + c] = Us[r * 3 + c];
									rec[99 + r * 3 + c] = Vs[r * 3 + c];
									rec[9 + r * 3 + c] = (lam * pm1 * G[r * 3 + c] + 2.0 * mu * (F[r * 3 + c] - R)) * da;
									// d2E / dsigma_r dsigma_c
									rec[108 + r * 3 + c] = (r == c ? 2.0 * mu + lam * other[r] * other[r] : lam * (sg[3 - r - c] * pm1 + other[r] * other[c])) * da;
								}
							for (int k = 0; k < 3; ++k)
							{
								const int l = (k + 1) % 3, m3 = 3 - k - l;
								const double left = mu - 0.5 * lam * pm1 * sg[m3];
								const double right = (dE[k] + dE[l]) / (2.0 * fmax(sg[k] + sg[l], 1.0e-12));
								rec[117 + k] = (left + right) * da;
								rec[120 + k] = (left - right) * da;
							}
							e_loc += (mu * ((sg[0] - 1.0) * (sg[0] - 1.0) + (sg[1] - 1.0) * (sg[1] - 1.0) + (sg[2] - 1.0) * (sg[2] - 1.0)) + 0.5 * lam * pm1 * pm1) * da;
						}
						else if (MAT == PFA_VISCOUS_DAMPING)
						{
							// ViscousDamping.cpp:297-342 (energy), :122-170 (gradient), :16-62 + :173-229 (Hessian) in closed form. With
							// Fp = I + grad u_prev, dF/dt = (F - Fp) / dt, dE/dt = sym(dF/dt^T F), T = 2 psi dE/dt + phi tr(dE/dt) I and
							// A = 2 F - Fp (the variation of dE/dt is sym(A^T dF) / dt, its second variation 2 sym(dF1^T dF2) / dt):
							//   P = A T / dt,   tangent = SaintVenant's with F -> A, S -> 2 T / dt, mu -> psi / dt^2, lambda -> phi / dt^2
							// (psi, phi) = (lambda, mu) arrays. The record uses the SaintVenant slots.
							const double *sUP = ws + L.UP;
							double Fp[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
							for (int i = 0; i < n_loc; ++i)
							{
								const double *Di = sD + (q * n_loc + i) * 3;
								for (int r = 0; r < 3; ++r)
									for (int c = 0; c < 3; ++c)
										Fp[r * 3 + c] += sUP[i * 3 + r] * Di[c];
							}
							F[0] += 1.0;
							F[4] += 1.0;
							F[8] += 1.0;
							Fp[0] += 1.0;
							Fp[4] += 1.0;
							Fp[8] += 1.0;
							const double psi = lam, phi = mu, idt = a.inv_dt;
							double Fd[9], A[9], Ed[9], T[9];
							for (int k = 0; k < 9; ++k)
							{
								Fd[k] = (F[k] - Fp[k]) * idt;
								A[k] = 2.0 * F[k] - Fp[k];
							}
							for (int r = 0; r < 3; ++r)
								for (int c = 0; c < 3; ++c)
								{
									const double m1 = Fd[0 + r] * F[0 + c] + Fd[3 + r] * F[3 + c] + Fd[6 + r] * F[6 + c];
									const double m2 = Fd[0 + c] * F[0 + r] + Fd[3 + c] * F[3 + r] + Fd[6 + c] * F[6 + r];
									Ed[r * 3 + c] = 0.5 * (m1 + m2);
								}
							const double trE = Ed[0] + Ed[4] + Ed[8];
							double EE = 0.0;
							for (int k = 0; k < 9; ++k)
							{
								EE += Ed[k] * Ed[k];
								T[k] = 2.0 * psi * Ed[k];
							}
							T[0] += phi * trE;
							T[4] += phi * trE;
							T[8] += phi * trE;
							for (int r = 0; r < 3; ++r)
								for (int c = 0; c < 3; ++c)
								{
									rec[r * 3 + c] = A[r * 3 + c];
									rec[9 + r * 3 + c] = (A[r * 3 + 0] * T[0 + c] + A[r * 3 + 1] * T[3 + c] + A[r * 3 + 2] * T[6 + c]) * idt * da;
									rec[18 + r * 3 + c] = 2.0 * T[r * 3 + c] * idt * da;
								}
							rec[28] = psi * idt * idt * da;
							rec[29] = phi * idt * idt * da;
							int k6 = 0;
							for (int r = 0; r < 3; ++r)
								for (int c = r; c < 3; ++c)
									rec[30 + k6++] = rec[28] * (A[r * 3 + 0] * A[c * 3 + 0] + A[r * 3 + 1] * A[c * 3 + 1] + A[r * 3 + 2] * A[c * 3 + 2]);
							e_loc += (psi * EE + 0.5 * phi * trE * trE) * da;
						}
						else if (MAT == PFA_MOONEY_RIVLIN)
						{
							// MooneyRivlinElasticity.hpp:26-47: psi = c1 (J^-2/3 I1 - 3) + c2 (J^-4/3 I2 - 3) + k/2 ln^2 J with I1 = tr C,
							// I2 = (I1^2 - tr C^2) / 2, C = F^T F (the invariants of F~ F~^T are those of C scaled by powers of J). The
							// reference differentiates this by autodiff; here the chain rule over (I1, I2, J):
							//   dI1 = 2 F, dI2 = 2 F M (M = I1 I - C), dJ = cof F;  P = 2 psi_1 F + 2 psi_2 F M + psi_J cof F.
							// (c1, c2, k) = (lambda, mu, param3). Record: F | P da | F M | cof F | F F^T (6) | M (6) | 9 scalars x da.
							F[0] += 1.0;
							F[4] += 1.0;
							F[8] += 1.0;
							const double c1 = lam, c2 = mu, kk = m.param3[size_t(e) * m.mat_stride + ms];
							double C[9], G[9];
							cofactor3(F, G);
							const double J = det3(F);
							const double lJ = log(J); // NaN for J <= 0, propagates like the reference
							for (int r = 0; r < 3; ++r)
								for (int c = 0; c < 3; ++c)
									C[r * 3 + c] = F[0 + r] * F[0 + c] + F[3 + r] * F[3 + c] + F[6 + r] * F[6 + c];
							const double I1 = C[0] + C[4] + C[8];
							double trCC = 0.0;
							for (int k9 = 0; k9 < 9; ++k9)
								trCC += C[k9] * C[k9];
							const double I2 = 0.5 * (I1 * I1 - trCC);
							const double ja = pow(J, -2.0 / 3.0), jb = ja * ja, iJ = 1.0 / J;
							const double p1 = c1 * ja, p2 = c2 * jb;
							const double p1J = c1 * (-2.0 / 3.0) * ja * iJ, p2J = c2 * (-4.0 / 3.0) * jb * iJ;
							const double pJ = I1 * p1J + I2 * p2J + kk * lJ * iJ;
							const double pJJ = (c1 * I1 * (10.0 / 9.0) * ja + c2 * I2 * (28.0 / 9.0) * jb + kk * (1.0 - lJ)) * iJ * iJ;
							double Mm[9];
							for (int k9 = 0; k9 < 9; ++k9)
								Mm[k9] = -C[k9];
							Mm[0] += I1;
							Mm[4] += I1;
							Mm[8] += I1;
							for (int r = 0; r < 3; ++r)
								for (int c = 0; c < 3; ++c)
								{
									const double fm = F[r * 3 + 0] * Mm[0 + c] + F[r * 3 + 1] * Mm[3 + c] + F[r * 3 + 2] * Mm[6 + c];
									rec[r * 3 + c] = F[r * 3 + c];
									rec[18 + r * 3 + c] = fm;
									rec[27 + r * 3 + c] = G[r * 3 + c];
									rec[9 + r * 3 + c] = (2.0 * p1 * F[r * 3 + c] + 2.0 * p2 * fm + pJ * G[r * 3 + c]) * da;
								}
							int k6 = 0;
							for (int r = 0; r < 3; ++r)
								for (int c = r; c < 3; ++c)
								{
									rec[36 + k6] = F[r * 3 + 0] * F[c * 3 + 0] + F[r * 3 + 1] * F[c * 3 + 1] + F[r * 3 + 2] * F[c * 3 + 2];
									rec[42 + k6] = Mm[r * 3 + c];
									++k6;
								}
							rec[48] = p1 * da;
							rec[49] = p2 * da;
							rec[50] = pJ * da;
							rec[51] = p1J * da;
							rec[52] = p2J * da;
							rec[53] = pJJ * da;
							e_loc += (c1 * (ja * I1 - 3.0) + c2 * (jb * I2 - 3.0) + 0.5 * kk * lJ * lJ) * da;
						}
						else if (MAT == PFA_SAINT_VENANT)
						{
							// SaintVenantElasticity.cpp:219-266 in closed form: E = (F^T F - I) / 2, S = 2 mu E + lambda tr(E) I, P = F S,
							// psi = mu E:E + lambda/2 tr(E)^2
							F[0] += 1.0;
							F[4] += 1.0;
							F[8] += 1.0;
							double E[9], S[9], trE = 0.0, EE = 0.0;
							for (int r = 0; r < 3; ++r)
								for (int c = 0; c < 3; ++c)
									E[r * 3 + c] = 0.5 * (F[0 + r] * F[0 + c] + F[3 + r] * F[3 + c] + F[6 + r] * F[6 + c] - (r == c ? 1.0 : 0.0));
							trE = E[0] + E[4] + E[8];
							for (int k = 0; k < 9; ++k)
							{
								EE += E[k] * E[k];
								S[k] = 2.0 * mu * E[k];
							}
							S[0] += lam * trE;
							S[4] += lam * trE;
							S[8] += lam * trE;
							for (int r = 0; r < 3; ++r)
								for (int c = 0; c < 3; ++c)
								{
									rec[r * 3 + c] = F[r * 3 + c];
									rec[9 + r * 3 + c] = (F[r * 3 + 0] * S[0 + c] + F[r * 3 + 1] * S[3 + c] + F[r * 3 + 2] * S[6 + c]) * da;
									rec[18 + r * 3 + c] = S[r * 3 + c] * da;
								}
							int k6 = 0;
							for (int r = 0; r < 3; ++r)
								for (int c = r; c < 3; ++c)
									rec[30 + k6++] = mu * da * (F[r * 3 + 0] * F[c * 3 + 0] + F[r * 3 + 1] * F[c * 3 + 1] + F[r * 3 + 2] * F[c * 3 + 2]);
							e_loc += (mu * EE + 0.5 * lam * trE * trE) * da;
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

				if (MAT == PFA_FIXED_COROTATIONAL && want_h && !LINEAR)
				{
					// stiffness entry (i, j | r, s) of point q (x da): sum_kl A_kl U_ik V_jk U_rl V_sl + for the pairs {k, l}:
					// (L + R) (U_ik V_jl U_rk V_sl + U_il V_jk U_rl V_sk) + (L - R) (U_ik V_jl U_rl V_sk + U_il V_jk U_rk V_sl)
					for (int t = lane; t < n_qp * 81; t += 32)
					{
						const int q = t / 81, idx = t - q * 81;
						const int ij = idx / 9, rs = idx - ij * 9;
						const int i = ij / 3, j = ij - i * 3, r = rs / 3, sI = rs - r * 3;
						double *rec = sQ + q * QR;
						const double *Us = rec, *Vs = rec + 99, *Ak = rec + 108;
						double sum = 0.0;
						for (int k = 0; k < 3; ++k)
						{
							const double uv = Us[i * 3 + k] * Vs[j * 3 + k];
							sum += uv * (Ak[k * 3 + 0] * Us[r * 3 + 0] * Vs[sI * 3 + 0] + Ak[k * 3 + 1] * Us[r * 3 + 1] * Vs[sI * 3 + 1] + Ak[k * 3 + 2] * Us[r * 3 + 2] * Vs[sI * 3 + 2]);
							const int l = (k + 1) % 3;
							const double a_kl = Us[i * 3 + k] * Vs[j * 3 + l], a_lk = Us[i * 3 + l] * Vs[j * 3 + k];
							const double b_kl = Us[r * 3 + k] * Vs[sI * 3 + l], b_lk = Us[r * 3 + l] * Vs[sI * 3 + k];
							sum += rec[117 + k] * (a_kl * b_kl + a_lk * b_lk) + rec[120 + k] * (a_kl * b_lk + a_lk * b_kl);
						}
						rec[18 + idx] = sum;
					}
					__syncwarp();
				}

				// ---- 4. energy, gradient, A = C D ----
				if (want_e)
				{
					for (int o = 16; o > 0; o >>= 1)
						e_loc += __shfl_xor_sync(0xffffffffu, e_loc, o);
					energy_acc += e_loc; // identical on all lanes
					if (a.energy_per_el != nullptr && lane == 0)
						a.energy_per_el[m.elem_id ? m.elem_id[e] : e] = e_loc;
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
							const double *P = sQ + q * QR + 9 + c * 3;
							g += Di[0] * P[0] + Di[1] * P[1] + Di[2] * P[2];
						}
						atomicAdd(a.grad + size_t(sG[i]) * 3 + c, g);
					}
				}
				if (want_h && (MAT == PFA_NEOHOOKEAN || MAT == PFA_SAINT_VENANT || MAT == PFA_VISCOUS_DAMPING) && !LINEAR)
				{
					// A_i = C D_i (NeoHookean: C = cof F) or F D_i (SaintVenant: the first record slot holds F)
					for (int t = lane; t < n_qp * n_loc; t += 32)
					{
						const int q = t / n_loc;
						const double *C = sQ + q * QR;
						const double *Di = sD + t * 3;
						sA[t * 3 + 0] = C[0] * Di[0] + C[1] * Di[1] + C[2] * Di[2];
						sA[t * 3 + 1] = C[3] * Di[0] + C[4] * Di[1] + C[5] * Di[2];
						sA[t * 3 + 2] = C[6] * Di[0] + C[7] * Di[1] + C[8] * Di[2];
					}
				}
				__syncwarp();

				// ---- 5. local Hessian / stiffness blocks and scatter ----
				// project_to_psd: pass 0 puts the blocks into the local matrix, which is then projected; pass 1 scatters the projected
				// blocks, or - when the projection left the matrix unchanged - recomputes and scatters them like the plain path
				bool psd_changed = false;
				if (want_h)
				for (int pass = PSD ? 0 : 1; pass < 2; ++pass)
				{
					if (PSD && pass == 1)
					{
						__syncwarp();
						psd_changed = psd_project_warp(ws + L.H, ws + L.V, 3 * n_loc, psd_dim(n_loc), psd_dim(n_loc) | 1, reinterpret_cast<int *>(ws + L.PQ),
													   reinterpret_cast<int *>(ws + L.PQ) + psd_dim(n_loc) / 2, ws + L.CS, ws + L.CS + psd_dim(n_loc) / 2, lane);
					}
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
						if (PSD && pass == 1 && psd_changed)
						{
							const double *sH = ws + L.H;
							const int LDh = psd_dim(n_loc) | 1;
							for (int r = 0; r < 3; ++r)
								for (int c = 0; c < 3; ++c)
									blk[r * 3 + c] = sH[(i * 3 + r) * LDh + j * 3 + c];
						}
						else if (MAT == PFA_NEOHOOKEAN && !LINEAR)
						{
							double s = 0.0, W0 = 0.0, W1 = 0.0, W2 = 0.0;
							for (int q = 0; q < n_qp; ++q)
							{
								const double *rec = sQ + q * QR;
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
						else if (MAT == PFA_FIXED_COROTATIONAL && !LINEAR)
						{
							// H[(i,a),(j,b)] = sum_q sum_dd' D_i[d] T_q[(a,d),(b,d')] D_j[d']   (B^T T B of FixedCorotational.cpp:420-434)
							for (int q = 0; q < n_qp; ++q)
							{
								const double *T = sQ + q * QR + 18;
								const double *Di = sD + (q * n_loc + i) * 3, *Dj = sD + (q * n_loc + j) * 3;
								for (int r = 0; r < 3; ++r)
									for (int c = 0; c < 3; ++c)
									{
										double acc = 0.0;
										for (int d = 0; d < 3; ++d)
										{
											const double *row = T + (r * 3 + d) * 9 + c * 3;
											acc += Di[d] * (row[0] * Dj[0] + row[1] * Dj[1] + row[2] * Dj[2]);
										}
										blk[r * 3 + c] += acc;
									}
							}
						}
						else if (MAT == PFA_MOONEY_RIVLIN && !LINEAR)
						{
							// with f = F D, n = F M D, c = cof(F) D (per node), z = D_i x D_j:
							//   H[(i,a),(j,b)] = sum_q da [ p_i[a]^T Psi'' p_j[b]   (p = (2 f, 2 n, c); Psi'' has only the J row / column)
							//        + delta_ab (2 psi_1 D_i.D_j + 2 psi_2 D_i.M D_j) + psi_2 (4 f_i[a] f_j[b] - 2 f_j[a] f_i[b] - 2 (F F^T)_ab D_i.D_j)
							//        - psi_J hat(F z)_ab ]
							// (d2I1 = 2 I, d2I2[ij,kl] = 4 F_ij F_kl + 2 I1 d_ik d_jl - 2 (d_ik C_lj + F_il F_kj + d_jl B_ik), d2J = the hat blocks)
							for (int q = 0; q < n_qp; ++q)
							{
								const double *rec = sQ + q * QR;
								const double *Di = sD + (q * n_loc + i) * 3, *Dj = sD + (q * n_loc + j) * 3;
								const double *F = rec, *FM = rec + 18, *G = rec + 27, *B = rec + 36, *Mm = rec + 42;
								const double w1 = rec[48], w2 = rec[49], wJ = rec[50], w1J = rec[51], w2J = rec[52], wJJ = rec[53];
								double fi[3], fj[3], ni[3], nj[3], ci[3], cj[3];
								for (int r = 0; r < 3; ++r)
								{
									fi[r] = F[r * 3] * Di[0] + F[r * 3 + 1] * Di[1] + F[r * 3 + 2] * Di[2];
									fj[r] = F[r * 3] * Dj[0] + F[r * 3 + 1] * Dj[1] + F[r * 3 + 2] * Dj[2];
									ni[r] = FM[r * 3] * Di[0] + FM[r * 3 + 1] * Di[1] + FM[r * 3 + 2] * Di[2];
									nj[r] = FM[r * 3] * Dj[0] + FM[r * 3 + 1] * Dj[1] + FM[r * 3 + 2] * Dj[2];
									ci[r] = G[r * 3] * Di[0] + G[r * 3 + 1] * Di[1] + G[r * 3 + 2] * Di[2];
									cj[r] = G[r * 3] * Dj[0] + G[r * 3 + 1] * Dj[1] + G[r * 3 + 2] * Dj[2];
								}
								const double dot = Di[0] * Dj[0] + Di[1] * Dj[1] + Di[2] * Dj[2];
								const double dMd = Di[0] * (Mm[0] * Dj[0] + Mm[1] * Dj[1] + Mm[2] * Dj[2]) + Di[1] * (Mm[1] * Dj[0] + Mm[3] * Dj[1] + Mm[4] * Dj[2])
												   + Di[2] * (Mm[2] * Dj[0] + Mm[4] * Dj[1] + Mm[5] * Dj[2]);
								for (int r = 0; r < 3; ++r)
								{
									// column operand of the rank part: Psi'' p_j (psi_11 = psi_12 = psi_22 = 0)
									const double pi1 = 2.0 * fi[r], pi2 = 2.0 * ni[r], pi3 = ci[r];
									for (int c = 0; c < 3; ++c)
									{
										const double q1 = w1J * cj[c], q2 = w2J * cj[c], q3 = w1J * 2.0 * fj[c] + w2J * 2.0 * nj[c] + wJJ * cj[c];
										blk[r * 3 + c] += pi1 * q1 + pi2 * q2 + pi3 * q3 + w2 * (4.0 * fi[r] * fj[c] - 2.0 * fj[r] * fi[c]);
									}
								}
								const double dg = 2.0 * w1 * dot + 2.0 * w2 * dMd, bd = 2.0 * w2 * dot;
								blk[0] += dg - bd * B[0];
								blk[1] -= bd * B[1];
								blk[2] -= bd * B[2];
								blk[3] -= bd * B[1];
								blk[4] += dg - bd * B[3];
								blk[5] -= bd * B[4];
								blk[6] -= bd * B[2];
								blk[7] -= bd * B[4];
								blk[8] += dg - bd * B[5];
								const double z0 = Di[1] * Dj[2] - Di[2] * Dj[1], z1 = Di[2] * Dj[0] - Di[0] * Dj[2], z2 = Di[0] * Dj[1] - Di[1] * Dj[0];
								const double W0 = wJ * (F[0] * z0 + F[1] * z1 + F[2] * z2), W1 = wJ * (F[3] * z0 + F[4] * z1 + F[5] * z2),
											 W2 = wJ * (F[6] * z0 + F[7] * z1 + F[8] * z2);
								blk[1] += W2;
								blk[2] -= W1;
								blk[3] -= W2;
								blk[5] += W0;
								blk[6] += W1;
								blk[7] -= W0;
							}
						}
						else if ((MAT == PFA_SAINT_VENANT || MAT == PFA_VISCOUS_DAMPING) && !LINEAR)
						{
							// (ViscousDamping: the same form with the substitutions of its record)
							// H[(i,a),(j,b)] = sum_q [ (D_i . S D_j) delta_ab + mu (F D_j)_a (F D_i)_b + lambda (F D_i)_a (F D_j)_b
							//                          + mu (F F^T)_ab (D_i . D_j) ] da     (tangent of P = F S(E))
							for (int q = 0; q < n_qp; ++q)
							{
								const double *rec = sQ + q * QR;
								const double *Di = sD + (q * n_loc + i) * 3, *Dj = sD + (q * n_loc + j) * 3;
								const double *Ai = sA + (q * n_loc + i) * 3, *Aj = sA + (q * n_loc + j) * 3;
								const double *S = rec + 18, *B = rec + 30;
								const double mu = rec[28], lam = rec[29];
								const double dsd = Di[0] * (S[0] * Dj[0] + S[1] * Dj[1] + S[2] * Dj[2]) + Di[1] * (S[3] * Dj[0] + S[4] * Dj[1] + S[5] * Dj[2])
												   + Di[2] * (S[6] * Dj[0] + S[7] * Dj[1] + S[8] * Dj[2]);
								const double dot = Di[0] * Dj[0] + Di[1] * Dj[1] + Di[2] * Dj[2];
								for (int r = 0; r < 3; ++r)
									for (int c = 0; c < 3; ++c)
										blk[r * 3 + c] += mu * (Aj[r] * Ai[c]) + lam * (Ai[r] * Aj[c]);
								blk[0] += dsd + B[0] * dot;
								blk[1] += B[1] * dot;
								blk[2] += B[2] * dot;
								blk[3] += B[1] * dot;
								blk[4] += dsd + B[3] * dot;
								blk[5] += B[4] * dot;
								blk[6] += B[2] * dot;
								blk[7] += B[4] * dot;
								blk[8] += dsd + B[5] * dot;
							}
						}
						else // LinearElasticity stiffness block (LinearElasticity.cpp:40-60)
						{
							for (int q = 0; q < n_qp; ++q)
							{
								const double *rec = sQ + q * QR;
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
						if (PSD && pass == 0)
						{
							double *sH = ws + L.H;
							const int LDh = psd_dim(n_loc) | 1;
							for (int r = 0; r < 3; ++r)
								for (int c = 0; c < 3; ++c)
									sH[(i * 3 + r) * LDh + j * 3 + c] = blk[r * 3 + c];
							continue;
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
		// NeoHookean kernel for P1 (n_loc 4, 1 qp) and P2 (n_loc 10, 4 qp) tets: "row-lane" mapping
		// in reference coordinates.
		//
		// With D_j = J^-T^T g_j (g_j = reference gradient of basis j at the quadrature point, the
		// same table for every element) the closed form (top of this file) becomes
		//     H[(i,m),(j,n)] = sum_q  Y_q[(i,m)][n] . g_j(q)
		// where the 3-vector Y[(i,m)][n] depends on the row (i,m) only. The column operand g_j(q)
		// is a mesh-independent constant: it lives in __constant__ memory and enters the DFMAs as
		// a constant-bank operand, so the inner loop (9 DFMA per column node) loads nothing.
		//
		//  phase 1  lane <-> (element, quadrature point). Gathers x, builds F, stress and
		//           coefficients and leaves a 34-double record per (element, qp) in shared memory:
		//             [0..5]   mu da K,  K = J^-T J^-1 (symmetric: 00 01 02 11 12 22)
		//             [6..14]  T = (c2 da F) cof(J^-T)^T   (row r: t_r = cof(J^-T) (c2 da F)_r)
		//             [15..23] CJ = cof(F) J^-T^T           (A_j = cof(F) D_j = CJ g_j)
		//             [24..32] PJ = (P da) J^-T^T           (gradient: G[(i,m)] = PJ_m . g_i)
		//             [33]     c1 da
		//  phase 2  lane <-> (local node i, component m) of one element; per quadrature point
		//             Y[m]       = mu da K g_i
		//             Y[(m+1)%3] =  t_{(m+2)%3} x g_i
		//             Y[(m+2)%3] = -t_{(m+1)%3} x g_i
		//             Y[n]      += c1 da (CJ_m . g_i) CJ_n        for n = 0,1,2
		//           then acc[j][n] += Y[n] . g_j(q) for all column nodes j (registers only), and
		//           after the qp loop 3 REDs per column node: for a fixed (j, n) the 3 NL lanes of an
		//           element write NL runs of 3 consecutive doubles inside one CSC column segment,
		//           the densest pattern the CSC layout allows.
		// ------------------------------------------------------------------------------------
		constexpr int kConstSlotDoubles = 4 * 10 * 3;
		__constant__ double c_refgrad[2][kConstSlotDoubles]; // slot 0: P1 [1][4][3], slot 1: P2 [4][10][3]

		template <int NL, int NQ>
		struct RowLane
		{
			static constexpr int SLOT = NL == 4 ? 0 : 1;
			static constexpr int ROWL = 3 * NL;       // phase-2 lanes per element
			static constexpr int EPW = 32 / ROWL;     // elements per phase-2 round
			static constexpr int EB = (32 / NQ) / EPW * EPW; // elements per warp batch (whole rounds)
			static constexpr int REC = 35;            // doubles per record (34 used), odd: conflict-free lane-strided stores
			static constexpr int WARP_DOUBLES = EB * NQ * REC;
			static constexpr int WARP_INTS = EB * NL * 2 + EB * NL * NL; // global node, column stride per (element, node); entry per (element, i, j)
			static_assert((EB * NL * 2) % 4 == 0 && (NL * NL) % 4 == 0, "16-byte cp.async staging of the entry table");
			static size_t smem_bytes(int warps)
			{
				return sizeof(double) * (size_t(NQ) * NL * 3 + ((NQ + 1) & ~1) + size_t(warps) * WARP_DOUBLES) + sizeof(int) * size_t(warps) * WARP_INTS;
			}
		};

		// In-kernel zero fill of values[] (DeviceMesh::zoff): the warp that draws batch index bi clears
		// the column blocks first touched by batch bi + kZeroLookahead (the first kZeroLookahead
		// batches also clear their own) and publishes zflag = epoch for them.
		__device__ __forceinline__ void zero_duty(const DeviceMesh &m, const AssembleArgs &a, int bi, int lane)
		{
			if (bi >= m.n_batches)
				return;
			for (int pass = (bi < kZeroLookahead ? 0 : 1); pass < 2; ++pass)
			{
				const int zb = pass == 0 ? bi : bi + kZeroLookahead;
				if (zb < m.n_batches)
				{
					const int r0 = m.zoff[zb], r1 = m.zoff[zb + 1];
					for (int rr = r0; rr < r1; ++rr)
					{
						const int2 run = m.zruns[rr];
						double *dst = a.values + size_t(run.x);
						for (int t = lane; t < run.y; t += 32)
							dst[t] = 0.0;
					}
					__threadfence();
					__syncwarp();
					if (lane == 0)
						*reinterpret_cast<volatile int32_t *>(m.zflag + zb) = a.epoch;
				}
			}
		}

		// P2S: the reference-gradient table has the structural zeros / equal components of the P2 tet basis
		// (auto_p_bases.cpp:1286-1344: grad phi_0 = c (1,1,1), grad phi_1..3 along one axis, phi_4 = (p,r,r),
		// phi_5 = (a,b,0), phi_6 = (r,p,r), phi_7 = (r,r,p), phi_8 = (a,0,b), phi_9 = (0,a,b)) at every quadrature
		// point, checked on the host (p2_structured); the column loop then needs 20 instead of 30 DFMA-pipe
		// operations per (row lane, column component, quadrature point).
		template <int NL, int NQ, int WARPS, int MINB, bool P2S>
		__global__ void __launch_bounds__(WARPS * 32, MINB) assemble_nh_rowlane_kernel(const DeviceMesh m, const AssembleArgs a)
		{
			using RL = RowLane<NL, NQ>;
			constexpr int EB = RL::EB, REC = RL::REC, ROWL = RL::ROWL, EPW = RL::EPW, SLOT = RL::SLOT;
			extern __shared__ double smem[];
			const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
			double *s_rg = smem;              // [NQ][NL][3] (lane-indexed reads; the column side uses c_refgrad)
			double *s_w = s_rg + NQ * NL * 3; // [NQ]
			double *s_rec = s_w + ((NQ + 1) & ~1) + warp * RL::WARP_DOUBLES;
			int *s_int = reinterpret_cast<int *>(s_w + ((NQ + 1) & ~1) + WARPS * RL::WARP_DOUBLES) + warp * RL::WARP_INTS;
			int *sG = s_int, *sStride = s_int + EB * NL, *sEnt = s_int + 2 * EB * NL;
			for (int t = threadIdx.x; t < NQ * NL * 3; t += WARPS * 32)
				s_rg[t] = m.ref_grads[t];
			for (int t = threadIdx.x; t < NQ; t += WARPS * 32)
				s_w[t] = m.qweights[t];
			__syncthreads();

			const bool want_h = a.values != nullptr;
			const bool want_g = a.grad != nullptr;
			const bool want_e = a.energy != nullptr || a.energy_per_el != nullptr;
			double energy_acc = 0.0;
			// phase-1 role
			const int el = lane / NQ, q = lane % NQ;
			// phase-2 role
			const int sub = lane / ROWL, r = lane % ROWL;
			const int ri = r / 3, mm = r % 3;
			const int ra = (mm + 1) % 3, rb = (mm + 2) % 3;
			const int is0 = mm == 0, is1 = mm == 1;
			const bool row_lane = lane < EPW * ROWL;

			// warp batches are handed out dynamically (one atomic per batch); the next index is
			// fetched before the current batch is processed so that its latency is hidden
			int batch = 0;
			if (lane == 0)
				batch = a.e_begin + atomicAdd(a.work_counter, EB);
			batch = __shfl_sync(0xffffffffu, batch, 0);
			if (a.epoch > 0)
				zero_duty(m, a, batch / EB, lane);
			// batch_quota > 0: a warp retires after that many batches (the grid is sized to cover the range),
			// so that CTAs of kernels on other streams - the interface exchange of the multi-GPU path -
			// get SM slots while this kernel is running; 0: persistent warps
			int left = a.batch_quota > 0 ? a.batch_quota : 0x7fffffff;
			while (batch < a.e_end)
			{
				int next = a.e_end; // a warp that has used up its quota draws nothing more
				--left;
				if (lane == 0 && left > 0)
					next = a.e_begin + atomicAdd(a.work_counter, EB);
				// ---- connectivity and entry offsets of the batch ----
				const int n_batch = min(EB, a.e_end - batch);
				for (int t = lane; t < n_batch * NL; t += 32)
				{
					sG[t] = m.conn[size_t(batch) * NL + t];
					if (want_h)
						sStride[t] = m.cstride[size_t(batch) * NL + t];
				}
				if (want_h)
				{
					// asynchronous 16-byte copies that land during phase 1
					const int32_t *src = m.entry + size_t(batch) * NL * NL;
					const uint32_t dst = uint32_t(__cvta_generic_to_shared(sEnt));
					for (int t = lane; t < n_batch * NL * NL / 4; t += 32)
						asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * t), "l"(src + 4 * t) : "memory");
					asm volatile("cp.async.commit_group;" ::: "memory");
				}
				__syncwarp();

				// ---- phase 1 ----
				const int e = batch + el;
				const bool valid = lane < EB * NQ && e < a.e_end;
				double e_q = 0.0;
#if PFA_EXP_MODE & 1
				if (valid)
					for (int k = 0; k < 34; ++k)
						s_rec[(el * NQ + q) * REC + k] = 1.0 + k;
				if (false)
#else
				if (valid)
#endif
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
						const double *rr = s_rg + (q * NL + i) * 3;
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
#pragma unroll
					for (int k = 0; k < 9; ++k)
						sq += F[k] * F[k];
					e_q = (0.5 * mu * (sq - 3.0 - 2.0 * lJ) + 0.5 * lam * lJ * lJ) * da;
					// PJ[a][c] = sum_k (P da)[a][k] J[c][k]
#pragma unroll
					for (int aa = 0; aa < 3; ++aa)
					{
						const double p0 = (mu * F[aa * 3 + 0] + pc * C[aa * 3 + 0]) * da;
						const double p1 = (mu * F[aa * 3 + 1] + pc * C[aa * 3 + 1]) * da;
						const double p2 = (mu * F[aa * 3 + 2] + pc * C[aa * 3 + 2]) * da;
#pragma unroll
						for (int c = 0; c < 3; ++c)
							rec[24 + aa * 3 + c] = p0 * J[c * 3 + 0] + p1 * J[c * 3 + 1] + p2 * J[c * 3 + 2];
					}
					if (want_h)
					{
						const double mu_da = mu * da;
						rec[0] = mu_da * (J[0] * J[0] + J[1] * J[1] + J[2] * J[2]);
						rec[1] = mu_da * (J[0] * J[3] + J[1] * J[4] + J[2] * J[5]);
						rec[2] = mu_da * (J[0] * J[6] + J[1] * J[7] + J[2] * J[8]);
						rec[3] = mu_da * (J[3] * J[3] + J[4] * J[4] + J[5] * J[5]);
						rec[4] = mu_da * (J[3] * J[6] + J[4] * J[7] + J[5] * J[8]);
						rec[5] = mu_da * (J[6] * J[6] + J[7] * J[7] + J[8] * J[8]);
						// rows of cof(J^-T): R_k = r_{k+1} x r_{k+2}; cofactor3 returns exactly these rows
						double R[9];
						cofactor3(J, R);
						const double c2da = pc * da;
#pragma unroll
						for (int rr = 0; rr < 3; ++rr)
						{
							const double f0 = c2da * F[rr * 3 + 0], f1 = c2da * F[rr * 3 + 1], f2 = c2da * F[rr * 3 + 2];
#pragma unroll
							for (int k = 0; k < 3; ++k)
								rec[6 + rr * 3 + k] = f0 * R[k * 3 + 0] + f1 * R[k * 3 + 1] + f2 * R[k * 3 + 2];
#pragma unroll
							for (int c = 0; c < 3; ++c)
								rec[15 + rr * 3 + c] = C[rr * 3 + 0] * J[c * 3 + 0] + C[rr * 3 + 1] * J[c * 3 + 1] + C[rr * 3 + 2] * J[c * 3 + 2];
						}
						rec[33] = (mu + lam * (1.0 - lJ)) * invJ * invJ * da;
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
							a.energy_per_el[m.elem_id ? m.elem_id[e] : e] = e_q;
					}
				}
				if (want_h)
					asm volatile("cp.async.wait_group 0;" ::: "memory");
				const int next_batch = __shfl_sync(0xffffffffu, next, 0);
				if (a.epoch > 0)
				{
					// zero-fill duty of the batch index just drawn: clear the column blocks first touched
					// kZeroLookahead batches after it and publish their flag. The duty is tied to the
					// DRAW of an index (draws are ordered in time), not to the start of its processing.
					zero_duty(m, a, next_batch / EB, lane);
					// the blocks this batch is the first to touch were cleared by the warp that drew batch
					// index - kZeroLookahead; blocks first touched by earlier batches transitively earlier
					if (lane == 0)
					{
						const volatile int32_t *flag = m.zflag + batch / EB;
						while (*flag != a.epoch)
							__nanosleep(64);
					}
					__syncwarp();
					__threadfence();
				}
				__syncwarp();

				// ---- phase 2 ----
				if (want_h || want_g)
				{
#pragma unroll 1
					for (int eb = 0; eb < EB; eb += EPW)
					{
						const int el2 = eb + sub;
						const int e2 = batch + el2;
						if (row_lane && e2 < a.e_end)
						{
							double acc[NL][3];
#pragma unroll
							for (int j = 0; j < NL; ++j)
								acc[j][0] = acc[j][1] = acc[j][2] = 0.0;
							double g_row = 0.0;
							constexpr int kUnrollQ = PFA_RL_UNROLL_Q;
#if PFA_EXP_MODE & 2
#pragma unroll
							for (int j = 0; j < NL; ++j)
								acc[j][0] = acc[j][1] = acc[j][2] = double(j + lane);
#endif
#pragma unroll kUnrollQ
							for (int qq = 0; qq < ((PFA_EXP_MODE & 2) ? 0 : NQ); ++qq)
							{
								const double *rec = s_rec + (el2 * NQ + qq) * REC;
								const double *gr = s_rg + (qq * NL + ri) * 3;
								const double g0 = gr[0], g1 = gr[1], g2 = gr[2];
								const double *pj = rec + 24 + mm * 3;
								g_row = fma(g0, pj[0], fma(g1, pj[1], fma(g2, pj[2], g_row)));
								if (want_h)
								{
									// Y rows before the rotation by m
									const double K00 = rec[0], K01 = rec[1], K02 = rec[2], K11 = rec[3], K12 = rec[4], K22 = rec[5];
									const double v0 = K00 * g0 + K01 * g1 + K02 * g2;
									const double v1 = K01 * g0 + K11 * g1 + K12 * g2;
									const double v2 = K02 * g0 + K12 * g1 + K22 * g2;
									const double *ta = rec + 6 + ra * 3, *tb = rec + 6 + rb * 3;
									const double b0 = tb[1] * g2 - tb[2] * g1, b1 = tb[2] * g0 - tb[0] * g2, b2 = tb[0] * g1 - tb[1] * g0; //  t_b x g
									const double a0 = ta[2] * g1 - ta[1] * g2, a1 = ta[0] * g2 - ta[2] * g0, a2 = ta[1] * g0 - ta[0] * g1; // -t_a x g
									const double *cj = rec + 15;
									const double cA = rec[33] * (cj[mm * 3 + 0] * g0 + cj[mm * 3 + 1] * g1 + cj[mm * 3 + 2] * g2);
									// rotated rows: Y[s] is row n = (m + s) % 3 of the header's Y (no lane-dependent selects)
									const double *c0r = cj + mm * 3, *c1r = cj + ra * 3, *c2r = cj + rb * 3;
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
										const double *G = &c_refgrad[SLOT][qq * NL * 3];
#pragma unroll
										for (int n = 0; n < 3; ++n)
										{
											const double y0 = Y[n][0], y1 = Y[n][1], y2 = Y[n][2];
											const double u12 = y1 + y2, u02 = y0 + y2, u01 = y0 + y1, sy = y0 + u12;
											acc[0][n] = fma(sy, G[0], acc[0][n]);
											acc[1][n] = fma(y0, G[3], acc[1][n]);
											acc[2][n] = fma(y1, G[7], acc[2][n]);
											acc[3][n] = fma(y2, G[11], acc[3][n]);
											acc[4][n] = fma(y0, G[12], fma(u12, G[13], acc[4][n]));
											acc[5][n] = fma(y0, G[15], fma(y1, G[16], acc[5][n]));
											acc[6][n] = fma(y1, G[19], fma(u02, G[18], acc[6][n]));
											acc[7][n] = fma(y2, G[23], fma(u01, G[21], acc[7][n]));
											acc[8][n] = fma(y0, G[24], fma(y2, G[26], acc[8][n]));
											acc[9][n] = fma(y1, G[28], fma(y2, G[29], acc[9][n]));
										}
									}
									else
									{
#pragma unroll
										for (int j = 0; j < NL; ++j)
										{
											const double c0 = c_refgrad[SLOT][(qq * NL + j) * 3 + 0], c1 = c_refgrad[SLOT][(qq * NL + j) * 3 + 1], c2 = c_refgrad[SLOT][(qq * NL + j) * 3 + 2];
											acc[j][0] = fma(Y[0][0], c0, fma(Y[0][1], c1, fma(Y[0][2], c2, acc[j][0])));
											acc[j][1] = fma(Y[1][0], c0, fma(Y[1][1], c1, fma(Y[1][2], c2, acc[j][1])));
											acc[j][2] = fma(Y[2][0], c0, fma(Y[2][1], c1, fma(Y[2][2], c2, acc[j][2])));
										}
									}
								}
							}
							if (want_g)
								add_gradient(a, size_t(sG[el2 * NL + ri]) * 3 + mm, g_row);
							if (want_h)
							{
								const int *ent = sEnt + el2 * NL * NL + ri * NL;
								const int *st = sStride + el2 * NL;
								// this lane's row (node ri, component mm) inside the (possibly reduced) column
								const unsigned mi = unsigned(st[ri]) >> 28;
								const bool row_kept = (mi >> mm) & 1u;
								const int row_off = __popc(mi & ((1u << mm) - 1u));
								const double sc = a.scale;
#if PFA_EXP_MODE & 4 // timing experiment: no scatter (the condition is never true)
								double ssum = 0.0;
#pragma unroll
								for (int j = 0; j < NL; ++j)
									ssum += acc[j][0] + acc[j][1] + acc[j][2];
								if (ssum == 12345.6789)
									red_add(a.values + ent[0] + (st[0] & 0x0fffffff), ssum);
#else
#pragma unroll
								for (int j = 0; j < NL; ++j)
								{
									double *dst = a.values + (size_t(ent[j]) + row_off);
									// undo the rotation with selects so that one instruction still writes one column
									// component n for all lanes (runs of 3 consecutive doubles per node)
									const double o0 = sc * sel(is0, acc[j][0], sel(is1, acc[j][2], acc[j][1]));
									const double o1 = sc * sel(is0, acc[j][1], sel(is1, acc[j][0], acc[j][2]));
									const double o2 = sc * sel(is0, acc[j][2], sel(is1, acc[j][1], acc[j][0]));
#if PFA_EXP_MODE & 16 // timing experiment: scatter without un-rotation, masks and scale (wrong columns, same addresses)
									{
										const size_t cs16 = size_t(unsigned(st[j]) & 0x0fffffffu);
										double *d16 = a.values + (size_t(ent[j]) + mm);
										red_add(d16, acc[j][0]);
										red_add(d16 + cs16, acc[j][1]);
										red_add(d16 + 2 * cs16, acc[j][2]);
										continue;
									}
#endif
#if PFA_EXP_MODE & 8 // timing experiment: plain stores instead of reductions (wrong values)
									const size_t cs = size_t(unsigned(st[j]) & 0x0fffffffu);
									dst[0] = o0;
									dst[cs] = o1;
									dst[2 * cs] = o2;
#else
									scatter3(dst, unsigned(st[j]), row_kept, o0, o1, o2);
#endif
								}
#endif
							}
						}
					}
				}
				// lanes without a row (lanes 30-31 for P2, 24-31 for P1, tail elements) must not run ahead into the next batch and
				// overwrite the connectivity / stride / entry tables the row lanes are still reading
				__syncwarp();
				batch = next_batch;
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
					atomicAdd(a.energy, t * a.scale);
				}
			}
		}

		// the reference-gradient table of a handle goes to its __constant__ slot when it differs
		// from what the slot holds (one table per basis order and device at a time: two handles
		// with DIFFERENT quadrature tables for the same order must not run concurrently)
		std::mutex g_const_mutex;
		double g_const_shadow[16][2][kConstSlotDoubles];
		bool g_const_valid[16][2] = {};

		cudaError_t ensure_const_table(const DeviceMesh &m, int slot, cudaStream_t st)
		{
			int dev = 0;
			cudaError_t err = cudaGetDevice(&dev);
			if (err != cudaSuccess)
				return err;
			if (dev < 0 || dev >= 16 || m.ref_grads_host == nullptr)
				return cudaErrorInvalidValue;
			const size_t bytes = sizeof(double) * size_t(m.n_qp) * m.n_loc * 3;
			std::lock_guard<std::mutex> lock(g_const_mutex);
			if (g_const_valid[dev][slot] && std::memcmp(g_const_shadow[dev][slot], m.ref_grads_host, bytes) == 0)
				return cudaSuccess;
			std::memcpy(g_const_shadow[dev][slot], m.ref_grads_host, bytes);
			g_const_valid[dev][slot] = true;
			return cudaMemcpyToSymbolAsync(c_refgrad, g_const_shadow[dev][slot], bytes, sizeof(double) * size_t(slot) * kConstSlotDoubles, cudaMemcpyHostToDevice, st);
		}

		template <int NL, int NQ, int WARPS, int MINB, bool P2S = false>
		cudaError_t launch_rowlane(const DeviceMesh &m, const AssembleArgs &a, int sm_count, cudaStream_t st)
		{
			using RL = RowLane<NL, NQ>;
			const size_t smem = RL::smem_bytes(WARPS);
			auto kern = assemble_nh_rowlane_kernel<NL, NQ, WARPS, MINB, P2S>;
			cudaError_t err = ensure_const_table(m, RL::SLOT, st);
			if (err != cudaSuccess)
				return err;
			err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
			if (err != cudaSuccess)
				return err;
			int per_sm = 1;
			err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem);
			if (err != cudaSuccess)
				return err;
			if (per_sm < 1)
				per_sm = 1;
			const int64_t n_range = std::max<int64_t>(1, int64_t(a.e_end) - a.e_begin);
			const int64_t need = (n_range + WARPS * RL::EB - 1) / (WARPS * RL::EB);
			int grid = int(std::max<int64_t>(1, std::min<int64_t>(need, int64_t(sm_count) * per_sm)));
			if (a.batch_quota > 0 && a.epoch == 0) // retiring warps: enough CTAs for every batch of the range
				grid = int(std::max<int64_t>(grid, (need + a.batch_quota - 1) / a.batch_quota));
			kern<<<grid, WARPS * 32, smem, st>>>(m, a);
			return cudaGetLastError();
		}

		// ------------------------------------------------------------------------------------
		// NeoHookean P1/P2 with project_to_psd (Assembler.cpp:693-694, ipc::project_to_psd): one
		// warp per element. The local Hessian is formed with the same row-lane math as above, put
		// in shared memory, diagonalised by a parallel cyclic Jacobi method (round-robin ordering:
		// N/2 disjoint rotations per step, lanes <-> rows), and, if its smallest eigenvalue is
		// negative, rebuilt as V max(D, 0) V^T before the scatter. Matrices with non-finite
		// entries and matrices that are already PSD are scattered unchanged (as the oracle does).
		// ------------------------------------------------------------------------------------
		template <int NL, int NQ, int WARPS>
		struct PsdLayout
		{
			static constexpr int N = 3 * NL;
			static constexpr int LD = N | 1; // odd leading dimension: conflict-free row- and column-wise access
			static constexpr int REC = 35;
			static constexpr int WARP_DOUBLES = NQ * REC + 2 * N * LD + 2 * (N / 2);
			static constexpr int WARP_INTS = 2 * (N / 2) + 2 * NL + NL * NL;
			static size_t smem_bytes()
			{
				return sizeof(double) * (size_t(NQ) * NL * 3 + ((NQ + 1) & ~1) + size_t(WARPS) * WARP_DOUBLES) + sizeof(int) * size_t(WARPS) * WARP_INTS;
			}
		};

		template <int NL, int NQ, int WARPS>
		__global__ void __launch_bounds__(WARPS * 32) assemble_nh_psd_kernel(const DeviceMesh m, const AssembleArgs a)
		{
			using PL = PsdLayout<NL, NQ, WARPS>;
			constexpr int N = PL::N, LD = PL::LD, REC = PL::REC, SLOT = RowLane<NL, NQ>::SLOT, HALF = N / 2;
			extern __shared__ double smem[];
			const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
			double *s_rg = smem;
			double *s_w = s_rg + NQ * NL * 3;
			double *ws = s_w + ((NQ + 1) & ~1) + warp * PL::WARP_DOUBLES;
			double *s_rec = ws, *sA = ws + NQ * REC, *sV = sA + N * LD, *sC = sV + N * LD, *sS = sC + HALF;
			int *wi = reinterpret_cast<int *>(s_w + ((NQ + 1) & ~1) + WARPS * PL::WARP_DOUBLES) + warp * PL::WARP_INTS;
			int *sP = wi, *sQ = wi + HALF, *sG = wi + 2 * HALF, *sStride = sG + NL, *sEnt = sStride + NL;
			for (int t = threadIdx.x; t < NQ * NL * 3; t += WARPS * 32)
				s_rg[t] = m.ref_grads[t];
			for (int t = threadIdx.x; t < NQ; t += WARPS * 32)
				s_w[t] = m.qweights[t];
			__syncthreads();

			const bool want_g = a.grad != nullptr;
			const bool want_e = a.energy != nullptr || a.energy_per_el != nullptr;
			double energy_acc = 0.0;
			const int ri = lane / 3, mm = lane % 3;
			const int ra = (mm + 1) % 3, rb = (mm + 2) % 3;
			const bool row_lane = lane < N;

			for (int e = a.e_begin + blockIdx.x * WARPS + warp; e < a.e_end; e += gridDim.x * WARPS)
			{
				for (int t = lane; t < NL; t += 32)
				{
					sG[t] = m.conn[size_t(e) * NL + t];
					sStride[t] = m.cstride[size_t(e) * NL + t];
				}
				for (int t = lane; t < NL * NL; t += 32)
					sEnt[t] = m.entry[size_t(e) * NL * NL + t];
				__syncwarp();
				// ---- phase 1: lane <-> quadrature point ----
				double e_q = 0.0;
				if (lane < NQ)
				{
					const int q = lane;
					double *rec = s_rec + q * REC;
					const size_t gi = m.geom_per_qp ? size_t(e) * NQ + q : size_t(e);
					double J[9];
#pragma unroll
					for (int k = 0; k < 9; ++k)
						J[k] = m.jit[gi * 9 + k];
					const double da = m.geom_per_qp ? m.detj[gi] : m.detj[e] * s_w[q];
					const size_t mi = size_t(e) * m.mat_stride + (m.mat_stride == 1 ? 0 : q);
					const double lam = m.lambda[mi], mu = m.mu[mi];
					double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
					for (int i = 0; i < NL; ++i)
					{
						const int g = sG[i];
						const double u0 = a.x[size_t(g) * 3 + 0], u1 = a.x[size_t(g) * 3 + 1], u2 = a.x[size_t(g) * 3 + 2];
						const double *rr = s_rg + (q * NL + i) * 3;
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
					const double lJ = log(Jd);
					double C[9];
					cofactor3(F, C);
					const double invJ = 1.0 / Jd;
					const double pc = (lam * lJ - mu) * invJ;
					double sq = 0.0;
					for (int k = 0; k < 9; ++k)
						sq += F[k] * F[k];
					e_q = (0.5 * mu * (sq - 3.0 - 2.0 * lJ) + 0.5 * lam * lJ * lJ) * da;
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
					cofactor3(J, R);
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
				}
				if (want_e)
				{
#pragma unroll
					for (int k = 1; k < NQ; k <<= 1)
						e_q += __shfl_xor_sync(0xffffffffu, e_q, k);
					if (lane == 0)
					{
						energy_acc += e_q;
						if (a.energy_per_el != nullptr)
							a.energy_per_el[m.elem_id ? m.elem_id[e] : e] = e_q;
					}
				}
				__syncwarp();

				// ---- phase 2: row (i, m) of the local Hessian in registers ----
				double acc[NL][3];
#pragma unroll
				for (int j = 0; j < NL; ++j)
					acc[j][0] = acc[j][1] = acc[j][2] = 0.0;
				double g_row = 0.0;
				if (row_lane)
				{
#pragma unroll 1
					for (int qq = 0; qq < NQ; ++qq)
					{
						const double *rec = s_rec + qq * REC;
						const double *gr = s_rg + (qq * NL + ri) * 3;
						const double g0 = gr[0], g1 = gr[1], g2 = gr[2];
						const double *pj = rec + 24 + mm * 3;
						g_row = fma(g0, pj[0], fma(g1, pj[1], fma(g2, pj[2], g_row)));
						const double K00 = rec[0], K01 = rec[1], K02 = rec[2], K11 = rec[3], K12 = rec[4], K22 = rec[5];
						const double v0 = K00 * g0 + K01 * g1 + K02 * g2;
						const double v1 = K01 * g0 + K11 * g1 + K12 * g2;
						const double v2 = K02 * g0 + K12 * g1 + K22 * g2;
						const double *ta = rec + 6 + ra * 3, *tb = rec + 6 + rb * 3;
						const double b0 = tb[1] * g2 - tb[2] * g1, b1 = tb[2] * g0 - tb[0] * g2, b2 = tb[0] * g1 - tb[1] * g0;
						const double a0 = ta[2] * g1 - ta[1] * g2, a1 = ta[0] * g2 - ta[2] * g0, a2 = ta[1] * g0 - ta[0] * g1;
						const double *cj = rec + 15;
						const double cA = rec[33] * (cj[mm * 3 + 0] * g0 + cj[mm * 3 + 1] * g1 + cj[mm * 3 + 2] * g2);
						double Y[3][3];
						Y[0][0] = fma(cA, cj[0], mm == 0 ? v0 : (mm == 1 ? a0 : b0));
						Y[0][1] = fma(cA, cj[1], mm == 0 ? v1 : (mm == 1 ? a1 : b1));
						Y[0][2] = fma(cA, cj[2], mm == 0 ? v2 : (mm == 1 ? a2 : b2));
						Y[1][0] = fma(cA, cj[3], mm == 1 ? v0 : (mm == 2 ? a0 : b0));
						Y[1][1] = fma(cA, cj[4], mm == 1 ? v1 : (mm == 2 ? a1 : b1));
						Y[1][2] = fma(cA, cj[5], mm == 1 ? v2 : (mm == 2 ? a2 : b2));
						Y[2][0] = fma(cA, cj[6], mm == 2 ? v0 : (mm == 0 ? a0 : b0));
						Y[2][1] = fma(cA, cj[7], mm == 2 ? v1 : (mm == 0 ? a1 : b1));
						Y[2][2] = fma(cA, cj[8], mm == 2 ? v2 : (mm == 0 ? a2 : b2));
#pragma unroll
						for (int j = 0; j < NL; ++j)
						{
							const double *cg = &c_refgrad[SLOT][(qq * NL + j) * 3];
							const double c0 = cg[0], c1 = cg[1], c2 = cg[2];
							acc[j][0] = fma(Y[0][0], c0, fma(Y[0][1], c1, fma(Y[0][2], c2, acc[j][0])));
							acc[j][1] = fma(Y[1][0], c0, fma(Y[1][1], c1, fma(Y[1][2], c2, acc[j][1])));
							acc[j][2] = fma(Y[2][0], c0, fma(Y[2][1], c1, fma(Y[2][2], c2, acc[j][2])));
						}
					}
					// local Hessian row -> shared memory; V = identity
#pragma unroll
					for (int j = 0; j < NL; ++j)
					{
						sA[lane * LD + j * 3 + 0] = acc[j][0];
						sA[lane * LD + j * 3 + 1] = acc[j][1];
						sA[lane * LD + j * 3 + 2] = acc[j][2];
					}
					for (int c = 0; c < N; ++c)
						sV[lane * LD + c] = c == lane ? 1.0 : 0.0;
				}
				__syncwarp();
				// the solver reads one triangle (SelfAdjointEigenSolver): mirror the lower one
				bool finite = true;
				if (row_lane)
				{
					for (int c = lane + 1; c < N; ++c)
						sA[lane * LD + c] = sA[c * LD + lane];
				}
				__syncwarp();
				if (row_lane)
				{
					for (int c = 0; c < N; ++c)
						finite = finite && isfinite(sA[lane * LD + c]);
				}
				finite = __all_sync(0xffffffffu, finite);
				bool nonzero = false;
				if (row_lane)
					for (int c = 0; c < N; ++c)
						nonzero = nonzero || sA[lane * LD + c] != 0.0;
				nonzero = __any_sync(0xffffffffu, nonzero);

				if (finite && nonzero)
				{
					for (int sweep = 0; sweep < 100; ++sweep)
					{
						double off = 0.0, diag = 0.0;
						if (row_lane)
							for (int c = 0; c < N; ++c)
							{
								const double v = sA[lane * LD + c];
								if (c == lane)
									diag += v * v;
								else
									off += v * v;
							}
#pragma unroll
						for (int o = 16; o > 0; o >>= 1)
						{
							off += __shfl_xor_sync(0xffffffffu, off, o);
							diag += __shfl_xor_sync(0xffffffffu, diag, o);
						}
						if (off <= 1e-30 * diag || off == 0.0)
							break;
						for (int step = 0; step < N - 1; ++step)
						{
							// round-robin pairing: (N-1, step), ((step+k) mod (N-1), (step-k) mod (N-1))
							if (lane < HALF)
							{
								int p, q;
								if (lane == 0)
								{
									p = step;
									q = N - 1;
								}
								else
								{
									p = (step + lane) % (N - 1);
									q = (step - lane + (N - 1)) % (N - 1);
								}
								if (p > q)
								{
									const int t = p;
									p = q;
									q = t;
								}
								const double apq = sA[p * LD + q];
								double c = 1.0, sn = 0.0;
								if (apq != 0.0)
								{
									const double app = sA[p * LD + p], aqq = sA[q * LD + q];
									const double theta = (aqq - app) / (2.0 * apq);
									const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
									c = 1.0 / sqrt(t * t + 1.0);
									sn = t * c;
								}
								sP[lane] = p;
								sQ[lane] = q;
								sC[lane] = c;
								sS[lane] = sn;
							}
							__syncwarp();
							if (row_lane) // A <- A J and V <- V J: lane = row
							{
								for (int k = 0; k < HALF; ++k)
								{
									const int p = sP[k], q = sQ[k];
									const double c = sC[k], sn = sS[k];
									const double akp = sA[lane * LD + p], akq = sA[lane * LD + q];
									sA[lane * LD + p] = c * akp - sn * akq;
									sA[lane * LD + q] = sn * akp + c * akq;
									const double vkp = sV[lane * LD + p], vkq = sV[lane * LD + q];
									sV[lane * LD + p] = c * vkp - sn * vkq;
									sV[lane * LD + q] = sn * vkp + c * vkq;
								}
							}
							__syncwarp();
							if (row_lane) // A <- J^T A: lane = column
							{
								for (int k = 0; k < HALF; ++k)
								{
									const int p = sP[k], q = sQ[k];
									const double c = sC[k], sn = sS[k];
									const double apk = sA[p * LD + lane], aqk = sA[q * LD + lane];
									sA[p * LD + lane] = c * apk - sn * aqk;
									sA[q * LD + lane] = sn * apk + c * aqk;
								}
							}
							__syncwarp();
						}
					}
					double w = row_lane ? sA[lane * LD + lane] : 0.0;
					double wmin = row_lane ? w : 1e300;
#pragma unroll
					for (int o = 16; o > 0; o >>= 1)
						wmin = fmin(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
					if (wmin < 0.0)
					{
						// rebuild row `lane` of V max(D,0) V^T; the clamped eigenvalues go to sC-like scratch in sA's diagonal
						__syncwarp();
						if (row_lane)
							sA[lane * LD + lane] = w < 0.0 ? 0.0 : w;
						__syncwarp();
						if (row_lane)
						{
#pragma unroll
							for (int j = 0; j < NL; ++j)
#pragma unroll
								for (int n = 0; n < 3; ++n)
								{
									const int c = j * 3 + n;
									double sum = 0.0;
									for (int k = 0; k < N; ++k)
										sum += sV[lane * LD + k] * sA[k * LD + k] * sV[c * LD + k];
									acc[j][n] = sum;
								}
						}
					}
				}
				__syncwarp();

				// ---- scatter ----
				if (row_lane)
				{
					if (want_g)
						add_gradient(a, size_t(sG[ri]) * 3 + mm, g_row);
					if (a.values != nullptr)
					{
						const unsigned mi = unsigned(sStride[ri]) >> 28;
						const bool row_kept = (mi >> mm) & 1u;
						const int row_off = __popc(mi & ((1u << mm) - 1u));
						const double sc = a.scale;
#pragma unroll
						for (int j = 0; j < NL; ++j)
						{
							double *dst = a.values + (size_t(sEnt[ri * NL + j]) + row_off);
							scatter3(dst, unsigned(sStride[j]), row_kept, sc * acc[j][0], sc * acc[j][1], sc * acc[j][2]);
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
					atomicAdd(a.energy, t * a.scale);
				}
			}
		}

		template <int NL, int NQ, int WARPS>
		cudaError_t launch_psd(const DeviceMesh &m, const AssembleArgs &a, int sm_count, cudaStream_t st)
		{
			using PL = PsdLayout<NL, NQ, WARPS>;
			const size_t smem = PL::smem_bytes();
			auto kern = assemble_nh_psd_kernel<NL, NQ, WARPS>;
			cudaError_t err = ensure_const_table(m, RowLane<NL, NQ>::SLOT, st);
			if (err != cudaSuccess)
				return err;
			err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
			if (err != cudaSuccess)
				return err;
			int per_sm = 1;
			err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem);
			if (err != cudaSuccess)
				return err;
			if (per_sm < 1)
				per_sm = 1;
			const int64_t need = (int64_t(m.n_el) + WARPS - 1) / WARPS;
			const int grid = int(std::max<int64_t>(1, std::min<int64_t>(need, int64_t(sm_count) * per_sm)));
			kern<<<grid, WARPS * 32, smem, st>>>(m, a);
			return cudaGetLastError();
		}

		// ------------------------------------------------------------------------------------
		// Laplacian stiffness (Laplacian.cpp:13-26, LinearAssembler::assemble Assembler.cpp:157-384)
		// for P1..P4 tets: one warp per element, K_e = sum_q da_q D_q D_q^T as a register-tiled
		// rank-3 update. 25 lanes hold a T x T tile each (T = ceil(NL / 5): 7 for P4, 4 for P3);
		// per quadrature point a lane reads T row and T column gradients (broadcast loads from the
		// warp's staging area) for 3 T^2 DFMAs. FP64-bound for P4 (84.5 kFLOP per element).
		// ------------------------------------------------------------------------------------
		template <int NL>
		struct LapTile
		{
			static constexpr int T = (NL + 4) / 5;
			static constexpr int NLP = 5 * T; // padded basis count (rows beyond NL hold zeros)
			__host__ __device__ static size_t warp_doubles(int n_qp) { return size_t(n_qp) * NLP * 3 + size_t(n_qp) * 9 + n_qp; }
			static size_t smem_bytes(int n_qp, int warps)
			{
				return sizeof(double) * (size_t(n_qp) * NL * 3 + n_qp + size_t(warps) * warp_doubles(n_qp));
			}
		};

		template <int NL, int WARPS>
		__global__ void __launch_bounds__(WARPS * 32) assemble_laplacian_tile_kernel(const DeviceMesh m, const AssembleArgs a)
		{
			using LT = LapTile<NL>;
			constexpr int T = LT::T, NLP = LT::NLP;
			extern __shared__ double smem[];
			const int n_qp = m.n_qp;
			const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
			double *s_rg = smem;               // [n_qp][NL][3]
			double *s_w = s_rg + n_qp * NL * 3; // [n_qp]
			double *ws = s_w + n_qp + warp * LT::warp_doubles(n_qp);
			double *sD = ws;                   // [n_qp][NLP][3] physical gradients
			double *sJ = sD + n_qp * NLP * 3;  // [gq][9]
			double *sDA = sJ + n_qp * 9;       // [n_qp]
			for (int t = threadIdx.x; t < n_qp * NL * 3; t += WARPS * 32)
				s_rg[t] = m.ref_grads[t];
			for (int t = threadIdx.x; t < n_qp; t += WARPS * 32)
				s_w[t] = m.qweights[t];
			__syncthreads();
			const int ti = lane / 5, tj = lane % 5;
			const bool tile_lane = lane < 25;
			const int gq = m.geom_per_qp ? n_qp : 1;

			for (int e = a.e_begin + blockIdx.x * WARPS + warp; e < a.e_end; e += gridDim.x * WARPS)
			{
				for (int t = lane; t < gq * 9; t += 32)
					sJ[t] = m.jit[size_t(e) * gq * 9 + t];
				for (int q = lane; q < n_qp; q += 32)
					sDA[q] = m.geom_per_qp ? m.detj[size_t(e) * n_qp + q] : m.detj[e] * s_w[q];
				__syncwarp();
				for (int t = lane; t < n_qp * NLP; t += 32)
				{
					const int q = t / NLP, i = t - q * NLP;
					double d0 = 0.0, d1 = 0.0, d2 = 0.0;
					if (i < NL)
					{
						const double *J = sJ + (m.geom_per_qp ? q * 9 : 0);
						const double *g = s_rg + (q * NL + i) * 3;
						d0 = g[0] * J[0] + g[1] * J[3] + g[2] * J[6];
						d1 = g[0] * J[1] + g[1] * J[4] + g[2] * J[7];
						d2 = g[0] * J[2] + g[1] * J[5] + g[2] * J[8];
					}
					sD[t * 3 + 0] = d0;
					sD[t * 3 + 1] = d1;
					sD[t * 3 + 2] = d2;
				}
				__syncwarp();
				if (tile_lane)
				{
					double acc[T][T];
#pragma unroll
					for (int r = 0; r < T; ++r)
#pragma unroll
						for (int c = 0; c < T; ++c)
							acc[r][c] = 0.0;
#pragma unroll 1
					for (int q = 0; q < n_qp; ++q)
					{
						const double da = sDA[q];
						const double *Dq = sD + q * NLP * 3;
						double cj[T][3];
#pragma unroll
						for (int c = 0; c < T; ++c)
						{
							cj[c][0] = da * Dq[(tj * T + c) * 3 + 0];
							cj[c][1] = da * Dq[(tj * T + c) * 3 + 1];
							cj[c][2] = da * Dq[(tj * T + c) * 3 + 2];
						}
#pragma unroll
						for (int r = 0; r < T; ++r)
						{
							const double r0 = Dq[(ti * T + r) * 3 + 0], r1 = Dq[(ti * T + r) * 3 + 1], r2 = Dq[(ti * T + r) * 3 + 2];
#pragma unroll
							for (int c = 0; c < T; ++c)
								acc[r][c] = fma(r0, cj[c][0], fma(r1, cj[c][1], fma(r2, cj[c][2], acc[r][c])));
						}
					}
					const int32_t *slot = m.slot + size_t(e) * NL * NL;
#pragma unroll
					for (int r = 0; r < T; ++r)
					{
						const int i = ti * T + r;
#pragma unroll
						for (int c = 0; c < T; ++c)
						{
							const int j = tj * T + c;
							if (i < NL && j < NL)
								red_add(a.values + slot[i * NL + j], acc[r][c]);
						}
					}
				}
				__syncwarp();
			}
		}

		template <int NL>
		cudaError_t launch_laplacian_tile(const DeviceMesh &m, const AssembleArgs &a, int sm_count, cudaStream_t st)
		{
			constexpr int WARPS = 8;
			const size_t smem = LapTile<NL>::smem_bytes(m.n_qp, WARPS);
			if (smem > 227 * 1024)
				return cudaErrorInvalidConfiguration;
			auto kern = assemble_laplacian_tile_kernel<NL, WARPS>;
			cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
			if (err != cudaSuccess)
				return err;
			int per_sm = 1;
			err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem);
			if (err != cudaSuccess)
				return err;
			if (per_sm < 1)
				per_sm = 1;
			const int64_t need = (int64_t(m.n_el) + WARPS - 1) / WARPS;
			const int grid = int(std::max<int64_t>(1, std::min<int64_t>(need, int64_t(sm_count) * per_sm)));
			kern<<<grid, WARPS * 32, smem, st>>>(m, a);
			return cudaGetLastError();
		}

		// ------------------------------------------------------------------------------------
		// Mass matrix (Mass.cpp:5-23 through LinearAssembler::assemble, Assembler.cpp:157-384), size 3:
		// M[(i,m),(j,m)] = sum_q rho phi_i(q) phi_j(q) da_q, the other entries of the 3x3 block are
		// stored zeros. Same register tiling as the Laplacian kernel with scalar operands; the basis
		// values are one CTA-wide table (they do not depend on the element).
		// ------------------------------------------------------------------------------------
		template <int NL, int WARPS>
		__global__ void __launch_bounds__(WARPS * 32) assemble_mass_tile_kernel(const DeviceMesh m, const AssembleArgs a)
		{
			constexpr int T = LapTile<NL>::T, NLP = LapTile<NL>::NLP;
			extern __shared__ double smem[];
			const int n_qp = m.n_qp;
			const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
			double *s_phi = smem;             // [n_qp][NLP], zero beyond NL
			double *s_w = s_phi + n_qp * NLP; // [n_qp]
			double *sC = s_w + n_qp + warp * n_qp; // [n_qp] rho * da of the warp's element
			for (int t = threadIdx.x; t < n_qp * NLP; t += WARPS * 32)
			{
				const int q = t / NLP, i = t - q * NLP;
				s_phi[t] = i < NL ? m.ref_vals[q * NL + i] : 0.0;
			}
			for (int t = threadIdx.x; t < n_qp; t += WARPS * 32)
				s_w[t] = m.qweights[t];
			__syncthreads();
			const int ti = lane / 5, tj = lane % 5;
			const bool tile_lane = lane < 25;

			for (int e = a.e_begin + blockIdx.x * WARPS + warp; e < a.e_end; e += gridDim.x * WARPS)
			{
				for (int q = lane; q < n_qp; q += 32)
				{
					const double da = m.geom_per_qp ? m.detj[size_t(e) * n_qp + q] : m.detj[e] * s_w[q];
					sC[q] = m.lambda[size_t(e) * m.mat_stride + (m.mat_stride == 1 ? 0 : q)] * da;
				}
				__syncwarp();
				if (tile_lane)
				{
					double acc[T][T];
#pragma unroll
					for (int r = 0; r < T; ++r)
#pragma unroll
						for (int c = 0; c < T; ++c)
							acc[r][c] = 0.0;
#pragma unroll 1
					for (int q = 0; q < n_qp; ++q)
					{
						const double *ph = s_phi + q * NLP;
						const double cq = sC[q];
						double cj[T];
#pragma unroll
						for (int c = 0; c < T; ++c)
							cj[c] = cq * ph[tj * T + c];
#pragma unroll
						for (int r = 0; r < T; ++r)
						{
							const double pr = ph[ti * T + r];
#pragma unroll
							for (int c = 0; c < T; ++c)
								acc[r][c] = fma(pr, cj[c], acc[r][c]);
						}
					}
					const int32_t *slot = m.slot + size_t(e) * NL * NL;
					const int32_t *conn = m.conn + size_t(e) * NL;
#pragma unroll
					for (int c = 0; c < T; ++c)
					{
						const int j = tj * T + c;
						if (j >= NL)
							continue;
						const int gj = conn[j];
						const int off = m.adj_off[gj], deg = m.adj_off[gj + 1] - off;
#pragma unroll
						for (int r = 0; r < T; ++r)
						{
							const int i = ti * T + r;
							if (i < NL)
							{
								// (row (g_i, mm), column (g_j, mm)) for mm = 0, 1, 2
								double *dst = a.values + (size_t(off) * 9 + size_t(slot[i * NL + j] - off) * 3);
								red_add(dst, acc[r][c]);
								red_add(dst + 3 * deg + 1, acc[r][c]);
								red_add(dst + 6 * deg + 2, acc[r][c]);
							}
						}
					}
				}
				__syncwarp();
			}
		}

		template <int NL>
		cudaError_t launch_mass_tile(const DeviceMesh &m, const AssembleArgs &a, int sm_count, cudaStream_t st)
		{
			constexpr int WARPS = 8;
			const size_t smem = sizeof(double) * (size_t(m.n_qp) * LapTile<NL>::NLP + m.n_qp + size_t(WARPS) * m.n_qp);
			if (smem > 227 * 1024)
				return cudaErrorInvalidConfiguration;
			auto kern = assemble_mass_tile_kernel<NL, WARPS>;
			cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
			if (err != cudaSuccess)
				return err;
			int per_sm = 1;
			err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem);
			if (err != cudaSuccess)
				return err;
			if (per_sm < 1)
				per_sm = 1;
			const int64_t need = (int64_t(m.n_el) + WARPS - 1) / WARPS;
			const int grid = int(std::max<int64_t>(1, std::min<int64_t>(need, int64_t(sm_count) * per_sm)));
			kern<<<grid, WARPS * 32, smem, st>>>(m, a);
			return cudaGetLastError();
		}

		// ------------------------------------------------------------------------------------
		// Linear stiffness on AFFINE elements with per-element constant coefficients (Laplacian.cpp:13-26,
		// LinearElasticity.cpp:30-63 through LinearAssembler::assemble): the quadrature sum factors out
		// of the element loop. With g_i = ghat_i J^-T,
		//     sum_q w_q g_i[a] g_j[b] = (J^-T^T S_ij J^-T)[a][b],   S_ij[c][d] = sum_q w_q ghat_i[c](q) ghat_j[d](q),
		// and the nine NL x NL reference moment matrices S^{cd} do not depend on the element: they are
		// built once per handle (DeviceMesh::ref_moments) and live in shared memory. Per node pair the
		// element then costs 9 (Laplacian) or ~70 (elasticity) DFMAs instead of n_qp times that
		// (P4: 23 quadrature points) - the kernel is bound by the scatter, not by the FP64 pipe.
		// One warp per element, lanes over the NL^2 node pairs.
		// ------------------------------------------------------------------------------------
		template <int MAT, int WARPS>
		__global__ void __launch_bounds__(WARPS * 32) assemble_affine_linear_kernel(const DeviceMesh m, const AssembleArgs a)
		{
			extern __shared__ double smem[];
			const int NL = m.n_loc, NP = NL * NL;
			const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
			double *sS = smem; // [9][NP]
			for (int t = threadIdx.x; t < 9 * NP; t += WARPS * 32)
				sS[t] = m.ref_moments[t];
			__syncthreads();

			for (int e = a.e_begin + blockIdx.x * WARPS + warp; e < a.e_end; e += gridDim.x * WARPS)
			{
				double J[9];
#pragma unroll
				for (int k = 0; k < 9; ++k)
					J[k] = m.jit[size_t(e) * 9 + k];
				const double det = m.detj[e]; // quadrature weights are inside the moments
				const int32_t *slot = m.slot + size_t(e) * NP;
				if (MAT == PFA_LAPLACIAN)
				{
					// M = det * J^-T J^-T^T  (D_i . D_j = ghat_i^T M ghat_j)
					double M[9];
#pragma unroll
					for (int c = 0; c < 3; ++c)
#pragma unroll
						for (int d = 0; d < 3; ++d)
							M[c * 3 + d] = det * (J[c * 3 + 0] * J[d * 3 + 0] + J[c * 3 + 1] * J[d * 3 + 1] + J[c * 3 + 2] * J[d * 3 + 2]);
					for (int p = lane; p < NP; p += 32)
					{
						double v = 0.0;
#pragma unroll
						for (int k = 0; k < 9; ++k)
							v = fma(M[k], sS[k * NP + p], v);
						red_add(a.values + slot[p], v);
					}
				}
				else
				{
					// lane <-> (node pair, row component aa): 10 pairs per warp pass; one RED instruction then
					// writes, for a fixed column component b, runs of 3 consecutive doubles (aa = 0, 1, 2)
					const double mu = m.mu[e] * det, lam = m.lambda[e] * det;
					const int32_t *conn = m.conn + size_t(e) * NL;
					const int aa = lane % 3, sub = lane / 3;
					// column aa of J^-T (lane-dependent, built once per element with selects)
					const double ja0 = aa == 0 ? J[0] : (aa == 1 ? J[1] : J[2]);
					const double ja1 = aa == 0 ? J[3] : (aa == 1 ? J[4] : J[5]);
					const double ja2 = aa == 0 ? J[6] : (aa == 1 ? J[7] : J[8]);
					// M = J^-T J^-T^T for the trace: tr P = sum_cd S[c][d] M[c][d]
					double M[9];
#pragma unroll
					for (int c = 0; c < 3; ++c)
#pragma unroll
						for (int d = 0; d < 3; ++d)
							M[c * 3 + d] = J[c * 3 + 0] * J[d * 3 + 0] + J[c * 3 + 1] * J[d * 3 + 1] + J[c * 3 + 2] * J[d * 3 + 2];
					for (int p0 = 0; p0 < NP; p0 += 10)
					{
						const int p = p0 + sub;
						if (lane >= 30 || p >= NP)
							continue;
						const int i = p / NL, j = p - i * NL;
						(void)i;
						double S[9];
#pragma unroll
						for (int k = 0; k < 9; ++k)
							S[k] = sS[k * NP + p];
						// P[a][b] = sum_cd J[c][a] S[c][d] J[d][b] = sum_q w g_i[a] g_j[b]
						// row aa of P:    U[d] = sum_c J[c][aa] S[c][d],   P[aa][b] = sum_d U[d] J[d][b]
						const double u0 = ja0 * S[0] + ja1 * S[3] + ja2 * S[6];
						const double u1 = ja0 * S[1] + ja1 * S[4] + ja2 * S[7];
						const double u2 = ja0 * S[2] + ja1 * S[5] + ja2 * S[8];
						// column aa of P: V[c] = sum_d S[c][d] J[d][aa],   P[b][aa] = sum_c J[c][b] V[c]
						const double v0 = S[0] * ja0 + S[1] * ja1 + S[2] * ja2;
						const double v1 = S[3] * ja0 + S[4] * ja1 + S[5] * ja2;
						const double v2 = S[6] * ja0 + S[7] * ja1 + S[8] * ja2;
						double tr = 0.0;
#pragma unroll
						for (int k = 0; k < 9; ++k)
							tr = fma(S[k], M[k], tr);
						tr *= mu;
						const int gj = conn[j];
						const int off = m.adj_off[gj], deg = m.adj_off[gj + 1] - off;
						double *dst = a.values + (size_t(off) * 9 + size_t(slot[p] - off) * 3 + aa);
						// K[(i,aa),(j,b)] = mu P[b][aa] + lambda P[aa][b] + mu tr(P) delta  (LinearElasticity.cpp:40-60)
#pragma unroll
						for (int b = 0; b < 3; ++b)
						{
							const double row = u0 * J[0 * 3 + b] + u1 * J[1 * 3 + b] + u2 * J[2 * 3 + b]; // P[aa][b]
							const double col = J[0 * 3 + b] * v0 + J[1 * 3 + b] * v1 + J[2 * 3 + b] * v2; // P[b][aa]
							red_add(dst + size_t(b) * 3 * deg, mu * col + lam * row + (aa == b ? tr : 0.0));
						}
					}
				}
			}
		}

		template <int MAT>
		cudaError_t launch_affine_linear(const DeviceMesh &m, const AssembleArgs &a, int sm_count, cudaStream_t st)
		{
			constexpr int WARPS = MAT == PFA_LAPLACIAN ? 16 : 8; // elasticity needs more than 64 registers per thread
			const size_t smem = sizeof(double) * 9 * size_t(m.n_loc) * m.n_loc;
			if (smem > 200 * 1024)
				return cudaErrorInvalidConfiguration;
			auto kern = assemble_affine_linear_kernel<MAT, WARPS>;
			cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
			if (err != cudaSuccess)
				return err;
			int per_sm = 1;
			err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem);
			if (err != cudaSuccess)
				return err;
			if (per_sm < 1)
				per_sm = 1;
			const int64_t need = (int64_t(m.n_el) + WARPS - 1) / WARPS;
			const int grid = int(std::max<int64_t>(1, std::min<int64_t>(need, int64_t(sm_count) * per_sm)));
			kern<<<grid, WARPS * 32, smem, st>>>(m, a);
			return cudaGetLastError();
		}

		constexpr size_t kMaxSmem = 227 * 1024;

		size_t generic_smem_bytes(int n_loc, int n_qp, int warps, bool psd = false, bool tables_shared = true, int qrec = kQRec, bool prev = false)
		{
			const WarpLayout L = warp_layout(n_loc, n_qp, psd, qrec, prev);
			return sizeof(double) * ((tables_shared ? size_t(n_qp) * n_loc * 3 + n_qp : size_t(0)) + size_t(warps) * L.total);
		}

		// warps per CTA: the largest of 8/4/2/1 whose staging fits in shared memory
		int pick_warps(int n_loc, int n_qp, int qrec = kQRec, bool prev = false)
		{
			for (int w = 8; w >= 1; w >>= 1)
				if (generic_smem_bytes(n_loc, n_qp, w, false, true, qrec, prev) <= kMaxSmem)
					return w;
			return 0;
		}

		template <int MAT, bool LINEAR, int kWarps>
		cudaError_t launch_generic_w(const DeviceMesh &m, const AssembleArgs &a, int sm_count, cudaStream_t st)
		{
			const size_t smem = generic_smem_bytes(m.n_loc, m.n_qp, kWarps, false, true, qrec_of(MAT), needs_prev(MAT));
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

		// project_to_psd through the generic kernel: two warps per CTA, one when the two local matrices (2 Np (Np|1) doubles per
		// warp) of two warps do not fit the shared memory
		template <int MAT, int kW, bool TABLES_SHARED>
		cudaError_t launch_generic_psd_w(const DeviceMesh &m, const AssembleArgs &a, int sm_count, cudaStream_t st)
		{
			const size_t smem = generic_smem_bytes(m.n_loc, m.n_qp, kW, true, TABLES_SHARED, qrec_of(MAT), needs_prev(MAT));
			if (smem > kMaxSmem)
				return cudaErrorNotSupported;
			auto kern = assemble_generic_kernel<MAT, false, kW, true, TABLES_SHARED>;
			cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
			if (err != cudaSuccess)
				return err;
			int per_sm = 1;
			if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kW * 32, smem)) != cudaSuccess)
				return err;
			const int64_t need = (int64_t(m.n_el) + kW - 1) / kW;
			const int grid = int(std::max<int64_t>(1, std::min<int64_t>(need, int64_t(sm_count) * std::max(per_sm, 1))));
			kern<<<grid, kW * 32, smem, st>>>(m, a);
			return cudaGetLastError();
		}

		// two warps per CTA; one when the local matrices (2 Np (Np|1) doubles per warp) of two do not fit the shared memory; P4
		// (23 points x 35 nodes) fits only with the reference tables left in global memory
		template <int MAT>
		cudaError_t launch_generic_psd(const DeviceMesh &m, const AssembleArgs &a, int sm_count, cudaStream_t st)
		{
			if (psd_dim(m.n_loc) > 128) // the diagonal of the rebuilt matrix is held in four registers per lane
				return cudaErrorNotSupported;
			if (generic_smem_bytes(m.n_loc, m.n_qp, 2, true, true, qrec_of(MAT), needs_prev(MAT)) <= kMaxSmem)
				return launch_generic_psd_w<MAT, 2, true>(m, a, sm_count, st);
			if (generic_smem_bytes(m.n_loc, m.n_qp, 1, true, true, qrec_of(MAT), needs_prev(MAT)) <= kMaxSmem)
				return launch_generic_psd_w<MAT, 1, true>(m, a, sm_count, st);
			return launch_generic_psd_w<MAT, 1, false>(m, a, sm_count, st);
		}

		template <int MAT, bool LINEAR>
		cudaError_t launch_generic(const DeviceMesh &m, const AssembleArgs &a, int sm_count, cudaStream_t st)
		{
			switch (pick_warps(m.n_loc, m.n_qp, qrec_of(MAT), needs_prev(MAT)))
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

	// affine geometry, one (lambda, mu) per element, reference moments present
	bool affine_linear_applies(const DeviceMesh &m)
	{
		return !PFA_NO_AFFINE_LINEAR && m.geom_per_qp == 0 && m.mat_stride == 1 && m.ref_moments != nullptr && m.slot != nullptr;
	}

	void reference_moments(const double *ref_grads, const double *weights, int n_loc, int n_qp, std::vector<double> &out)
	{
		const size_t np = size_t(n_loc) * n_loc;
		out.assign(9 * np, 0.0);
		for (int c = 0; c < 3; ++c)
			for (int d = 0; d < 3; ++d)
				for (int i = 0; i < n_loc; ++i)
					for (int j = 0; j < n_loc; ++j)
					{
						double s = 0.0;
						for (int q = 0; q < n_qp; ++q)
							s += weights[q] * (ref_grads[(size_t(q) * n_loc + i) * 3 + c] * ref_grads[(size_t(q) * n_loc + j) * 3 + d]);
						out[size_t(c * 3 + d) * np + size_t(i) * n_loc + j] = s;
					}
	}

	bool p2_table_structured(const double *g, int n_loc, int n_qp)
	{
		if (n_loc != 10 || g == nullptr)
			return false;
		for (int q = 0; q < n_qp; ++q)
		{
			const double *G = g + size_t(q) * 30;
			const bool ok = G[0] == G[1] && G[0] == G[2]        // phi_0: c (1,1,1)
							&& G[4] == 0 && G[5] == 0             // phi_1: (a,0,0)
							&& G[6] == 0 && G[8] == 0             // phi_2: (0,a,0)
							&& G[9] == 0 && G[10] == 0            // phi_3: (0,0,a)
							&& G[13] == G[14]                     // phi_4: (p,r,r)
							&& G[17] == 0                         // phi_5: (a,b,0)
							&& G[18] == G[20]                     // phi_6: (r,p,r)
							&& G[21] == G[22]                     // phi_7: (r,r,p)
							&& G[25] == 0                         // phi_8: (a,0,b)
							&& G[27] == 0;                        // phi_9: (0,a,b)
			if (!ok)
				return false;
		}
		return true;
	}

	bool rowlane_applies(int material, int n_loc, int n_qp)
	{
		return material == PFA_NEOHOOKEAN && ((n_loc == 10 && n_qp == 4) || (n_loc == 4 && n_qp == 1));
	}

	int rowlane_batch_elements(int n_loc, int n_qp)
	{
		return n_loc == 4 ? RowLane<4, 1>::EB : RowLane<10, 4>::EB;
	}

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

	cudaError_t launch_gather_rows(const double *src, const int32_t *perm, int n, int stride, double *dst, cudaStream_t st)
	{
		const int64_t total = int64_t(n) * stride;
		gather_rows_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(src, perm, n, stride, dst);
		return cudaGetLastError();
	}

	cudaError_t launch_expand_inner(const DeviceMesh &m, int32_t *outer, int32_t *inner, cudaStream_t st)
	{
		const int warps = 8;
		expand_inner_kernel<int32_t><<<(m.n_bases + warps - 1) / warps, warps * 32, 0, st>>>(m.adj_off, m.adj, m.n_bases, m.size, outer, inner);
		return cudaGetLastError();
	}

	cudaError_t launch_expand_inner64(const DeviceMesh &m, int64_t *outer, int64_t *inner, cudaStream_t st)
	{
		const int warps = 8;
		expand_inner_kernel<int64_t><<<(m.n_bases + warps - 1) / warps, warps * 32, 0, st>>>(m.adj_off, m.adj, m.n_bases, m.size, outer, inner);
		return cudaGetLastError();
	}

	cudaError_t launch_assemble(const DeviceMesh &m, const AssembleArgs &a, bool linear, int sm_count, cudaStream_t st, const char **kernel_name)
	{
		if (kernel_name)
			*kernel_name = "assemble_generic_kernel";
		switch (m.material)
		{
		case PFA_NEOHOOKEAN:
			if (a.project_to_psd)
			{
				if (kernel_name)
					*kernel_name = "assemble_nh_psd_kernel";
				if (m.n_loc == 10 && m.n_qp == 4 && m.entry != nullptr)
					return launch_psd<10, 4, 8>(m, a, sm_count, st);
				if (m.n_loc == 4 && m.n_qp == 1 && m.entry != nullptr)
					return launch_psd<4, 1, 8>(m, a, sm_count, st);
				if (kernel_name)
					*kernel_name = "assemble_generic_kernel<NeoHookean,psd>";
				return launch_generic_psd<PFA_NEOHOOKEAN>(m, a, sm_count, st);
			}
			if (m.n_loc == 10 && m.n_qp == 4)
			{
				if (kernel_name)
					*kernel_name = "assemble_nh_rowlane_kernel<10,4>";
				if (m.p2_structured && !PFA_RL_DENSE_COLUMNS)
					return launch_rowlane<10, 4, PFA_RL_WARPS_P2, PFA_RL_MINB_P2, true>(m, a, sm_count, st);
				return launch_rowlane<10, 4, PFA_RL_WARPS_P2, PFA_RL_MINB_P2>(m, a, sm_count, st);
			}
			if (m.n_loc == 4 && m.n_qp == 1)
			{
				if (kernel_name)
					*kernel_name = "assemble_nh_rowlane_kernel<4,1>";
				return launch_rowlane<4, 1, 8, 2>(m, a, sm_count, st);
			}
			return launch_generic<PFA_NEOHOOKEAN, false>(m, a, sm_count, st);
		case PFA_SAINT_VENANT:
			if (linear)
				return cudaErrorNotSupported;
			if (kernel_name)
				*kernel_name = a.project_to_psd ? "assemble_generic_kernel<SaintVenant,psd>" : "assemble_generic_kernel<SaintVenant>";
			if (a.project_to_psd)
				return launch_generic_psd<PFA_SAINT_VENANT>(m, a, sm_count, st);
			return launch_generic<PFA_SAINT_VENANT, false>(m, a, sm_count, st);
		case PFA_MOONEY_RIVLIN:
			if (linear)
				return cudaErrorNotSupported;
			if (kernel_name)
				*kernel_name = a.project_to_psd ? "assemble_generic_kernel<MooneyRivlin,psd>" : "assemble_generic_kernel<MooneyRivlin>";
			if (a.project_to_psd)
				return launch_generic_psd<PFA_MOONEY_RIVLIN>(m, a, sm_count, st);
			return launch_generic<PFA_MOONEY_RIVLIN, false>(m, a, sm_count, st);
		case PFA_FIXED_COROTATIONAL:
			if (linear)
				return cudaErrorNotSupported;
			if (kernel_name)
				*kernel_name = a.project_to_psd ? "assemble_generic_kernel<FixedCorotational,psd>" : "assemble_generic_kernel<FixedCorotational>";
			if (a.project_to_psd)
				return launch_generic_psd<PFA_FIXED_COROTATIONAL>(m, a, sm_count, st);
			return launch_generic<PFA_FIXED_COROTATIONAL, false>(m, a, sm_count, st);
		case PFA_VISCOUS_DAMPING:
			if (linear || a.x_prev == nullptr)
				return cudaErrorNotSupported;
			if (kernel_name)
				*kernel_name = a.project_to_psd ? "assemble_generic_kernel<ViscousDamping,psd>" : "assemble_generic_kernel<ViscousDamping>";
			if (a.project_to_psd)
				return launch_generic_psd<PFA_VISCOUS_DAMPING>(m, a, sm_count, st);
			return launch_generic<PFA_VISCOUS_DAMPING, false>(m, a, sm_count, st);
		case PFA_LINEAR_ELASTICITY:
			if (linear && affine_linear_applies(m) && a.values != nullptr)
			{
				cudaError_t err = launch_affine_linear<PFA_LINEAR_ELASTICITY>(m, a, sm_count, st);
				if (err != cudaErrorInvalidConfiguration)
				{
					if (kernel_name)
						*kernel_name = "assemble_affine_linear_kernel";
					return err;
				}
			}
			return linear ? launch_generic<PFA_LINEAR_ELASTICITY, true>(m, a, sm_count, st)
						  : launch_generic<PFA_LINEAR_ELASTICITY, false>(m, a, sm_count, st);
		case PFA_MASS:
			if (kernel_name)
				*kernel_name = "assemble_mass_tile_kernel";
			if (a.values == nullptr)
				return cudaErrorInvalidValue;
			switch (m.n_loc)
			{
			case 4:
				return launch_mass_tile<4>(m, a, sm_count, st);
			case 10:
				return launch_mass_tile<10>(m, a, sm_count, st);
			case 20:
				return launch_mass_tile<20>(m, a, sm_count, st);
			case 35:
				return launch_mass_tile<35>(m, a, sm_count, st);
			default:
				return cudaErrorNotSupported;
			}
		case PFA_LAPLACIAN:
			if (affine_linear_applies(m) && a.values != nullptr)
			{
				cudaError_t err = launch_affine_linear<PFA_LAPLACIAN>(m, a, sm_count, st);
				if (err != cudaErrorInvalidConfiguration)
				{
					if (kernel_name)
						*kernel_name = "assemble_affine_linear_kernel";
					return err;
				}
			}
			if (a.values != nullptr && !PFA_NO_LAPLACIAN_TILE)
			{
				cudaError_t err = cudaErrorInvalidConfiguration;
				if (m.n_loc == 35)
					err = launch_laplacian_tile<35>(m, a, sm_count, st);
				else if (m.n_loc == 20)
					err = launch_laplacian_tile<20>(m, a, sm_count, st);
				else if (m.n_loc == 10)
					err = launch_laplacian_tile<10>(m, a, sm_count, st);
				if (err != cudaErrorInvalidConfiguration)
				{
					if (kernel_name)
						*kernel_name = "assemble_laplacian_tile_kernel";
					return err;
				}
			}
			return launch_generic<PFA_LAPLACIAN, true>(m, a, sm_count, st);
		default:
			return cudaErrorInvalidValue;
		}
	}
} // namespace pfa
