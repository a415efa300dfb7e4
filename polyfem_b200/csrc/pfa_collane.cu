// Column-lane (owner-computes) NeoHookean assembly for P1 / P2 tets: opt-in alternative to the row-lane RED kernel
// (PFA_FLAG_COLUMN_LANE or PFA_COLUMN_LANE=1; pfa_grad_hess / pfa_hessian with the full matrix, affine elements).
// Math, record layout and schedule: pfa_collane.h (also compiled into the CPU emulation tests/collane_emul.cpp, which
// is checked against the oracle). Written after the round-1 GPU budget was spent: NOT yet run on a GPU; the default
// path does not use it.
//
//   collane_records_kernel   thread <-> (element, quadrature point): gathers x, writes the 34-double record to global
//                            memory, reduces the energy
//   collane_columns_kernel   warp <-> group of 10 nodes (3 lanes each); per step every slot takes one incident element
//                            of its node (the 10 records are staged in shared memory by coalesced copies, one step
//                            ahead through registers), adds its 3*NL entries to the lane's private strip (shared memory, address
//                            row*32 + lane: bank = lane, no conflicts, no atomics); after the last step the strip is the
//                            finished CSC column and is streamed to values[], the gradient entry is stored: every output
//                            is written exactly once, in a fixed summation order (bitwise reproducible), no zero fill
#include "pfa_collane.h"
#include "pfa_internal.h"

#include <cstring>
#include <mutex>

namespace pfa
{
	namespace
	{
		using namespace collane;
		constexpr int kSlotDoubles = 4 * 10 * 3;
		__constant__ double c_cl_refgrad[2][kSlotDoubles]; // slot 0: P1 [1][4][3], slot 1: P2 [4][10][3]

#ifndef PFA_CL_COOP_FLUSH
#define PFA_CL_COOP_FLUSH 1 // 0: every lane streams its own column (30 sectors per store instruction)
#endif
		constexpr int kFlushLd = 33; // leading dimension of the 32 x 32 transposition block of the cooperative flush

		template <int SLOT>
		struct ConstTable
		{
			__device__ __forceinline__ double operator[](int i) const { return c_cl_refgrad[SLOT][i]; }
		};

		template <int NL, int NQ>
		__global__ void __launch_bounds__(128) collane_records_kernel(const DeviceMesh m, const AssembleArgs a, double *__restrict__ rec_out, double *__restrict__ block_energy)
		{
			static_assert(32 % NQ == 0, "the quadrature points of an element sit in one warp");
			const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
			const int64_t e64 = t / NQ;
			const int q = int(t - e64 * NQ);
			const bool valid = e64 < m.n_el;
			constexpr int kRecLd = kRec + 1; // odd: conflict-free record-per-thread writes
			__shared__ double s_rec[128 * kRecLd];
			double e_q = 0.0;
			if (valid)
			{
				const int e = int(e64);
				double u[NL * 3];
#pragma unroll
				for (int i = 0; i < NL; ++i)
				{
					const size_t g = size_t(m.conn[size_t(e) * NL + i]);
					u[i * 3 + 0] = a.x[g * 3 + 0];
					u[i * 3 + 1] = a.x[g * 3 + 1];
					u[i * 3 + 2] = a.x[g * 3 + 2];
				}
				double J[9];
#pragma unroll
				for (int k = 0; k < 9; ++k)
					J[k] = m.jit[size_t(e) * 9 + k];
				const double da = m.detj[e] * m.qweights[q];
				const size_t mi = size_t(e) * m.mat_stride + (m.mat_stride == 1 ? 0 : q);
				e_q = qp_record<NL>(J, da, m.lambda[mi], m.mu[mi], u, m.ref_grads + size_t(q) * NL * 3, s_rec + size_t(threadIdx.x) * kRecLd);
			}
			// the 32 records of a warp are contiguous in global memory: write them with coalesced stores (a thread storing its
			// own record touches 32 lines per instruction)
			__syncwarp();
			{
				const int lane = threadIdx.x & 31;
				const int64_t t0 = t - lane; // first (element, qp) of this warp
				const int64_t n_valid = min(int64_t(32), int64_t(m.n_el) * NQ - t0);
				const double *src = s_rec + size_t(threadIdx.x - lane) * kRecLd;
				double *dst = rec_out + size_t(t0) * kRec;
				if (n_valid > 0)
					for (int idx = lane; idx < int(n_valid) * kRec; idx += 32)
						dst[idx] = src[(idx / kRec) * kRecLd + idx % kRec];
			}
			if (a.energy != nullptr || a.energy_per_el != nullptr)
			{
				double e_el = e_q;
#pragma unroll
				for (int k = 1; k < NQ; k <<= 1)
					e_el += __shfl_xor_sync(0xffffffffu, e_el, k);
				if (valid && q == 0 && a.energy_per_el != nullptr)
					a.energy_per_el[m.elem_id ? m.elem_id[e64] : e64] = e_el;
				if (a.energy != nullptr)
				{
					double w = e_q;
#pragma unroll
					for (int o = 16; o > 0; o >>= 1)
						w += __shfl_xor_sync(0xffffffffu, w, o);
					__shared__ double s_e[4];
					if ((threadIdx.x & 31) == 0)
						s_e[threadIdx.x >> 5] = w;
					__syncthreads();
					if (threadIdx.x == 0) // per-block partial sums, added up in a fixed order by collane_energy_kernel
						block_energy[blockIdx.x] = s_e[0] + s_e[1] + s_e[2] + s_e[3];
				}
			}
		}

		// energy = scale * sum of the per-block partial sums, in a fixed order (one block): like values[] and the gradient,
		// the energy of the column-lane path is the same bit pattern from call to call
		__global__ void __launch_bounds__(1024) collane_energy_kernel(const double *__restrict__ block_energy, int n_blocks, double scale, double *__restrict__ energy)
		{
			__shared__ double s[1024];
			double t = 0.0;
			for (int k = threadIdx.x; k < n_blocks; k += 1024)
				t += block_energy[k];
			s[threadIdx.x] = t;
			__syncthreads();
			for (int o = 512; o > 0; o >>= 1)
			{
				if (int(threadIdx.x) < o)
					s[threadIdx.x] += s[threadIdx.x + o];
				__syncthreads();
			}
			if (threadIdx.x == 0)
				*energy = s[0] * scale;
		}

		template <int NL, int NQ, int SLOT, bool P2S>
		__global__ void __launch_bounds__(256) collane_columns_kernel(const DeviceMesh m, const AssembleArgs a, const ColumnLaneTables t, const double *__restrict__ rec, int g_begin,
																	  int g_end, int strip_rows)
		{
			using CL = ColLayout<NQ>;
			constexpr int RECQ = CL::RECQ, SSTR = CL::SSTR, PER_LANE = CL::PER_LANE;
			constexpr unsigned kFull = 0xffffffffu;
			extern __shared__ double smem[];
			double *s_rg = smem; // [NQ][NL][3]: row-side reference gradients (lane-dependent index)
			for (int k = threadIdx.x; k < NQ * NL * 3; k += blockDim.x)
				s_rg[k] = m.ref_grads[k];
			__syncthreads();
			const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
			double *wbase = smem + ((NQ * NL * 3 + 1) & ~1) + size_t(warp) * (size_t(CL::STAGE) + size_t(strip_rows) * 32);
			double *stage = wbase;
			double *strip = wbase + CL::STAGE + lane;
			const int slot = lane / 3, mm = lane - slot * 3;
			const bool active = lane < 3 * kSlots;
			const uint4 *inc = reinterpret_cast<const uint4 *>(t.inc);
			const uint4 idle = make_uint4(0xffffffffu, 0u, 0u, 0u);
			for (int g = g_begin + blockIdx.x * warps + warp; g < g_end; g += gridDim.x * warps)
			{
				const int rows = t.grp_rows[g], s0 = t.grp_off[g], s1 = t.grp_off[g + 1];
				const int b = active ? t.grp_node[size_t(g) * kSlots + slot] : -1;
				for (int r = 0; r < rows; ++r)
					strip[r * 32] = 0.0;
				double g_acc = 0.0;
				// records of the first step -> registers (every lane takes PER_LANE consecutive-by-32 doubles of the 10 records)
				uint4 w_nxt = (active && s0 < s1) ? inc[size_t(s0) * kSlots + slot] : idle;
				double R[PER_LANE];
#pragma unroll
				for (int i = 0; i < PER_LANE; ++i)
				{
					int tt, kk;
					CL::staged(lane, i, tt, kk);
					const int et = __shfl_sync(kFull, int(w_nxt.x), min(3 * tt, 31));
					R[i] = (tt < kSlots && et >= 0) ? rec[size_t(et) * RECQ + kk] : 0.0;
				}
				for (int s = s0; s < s1; ++s)
				{
					const uint4 w = w_nxt;
#pragma unroll
					for (int i = 0; i < PER_LANE; ++i)
					{
						int tt, kk;
						CL::staged(lane, i, tt, kk);
						if (tt < kSlots)
							stage[tt * SSTR + kk] = R[i];
					}
					__syncwarp();
					if (s + 1 < s1) // warp-uniform: the next step's records travel while this step is computed
					{
						w_nxt = active ? inc[size_t(s + 1) * kSlots + slot] : idle;
#pragma unroll
						for (int i = 0; i < PER_LANE; ++i)
						{
							int tt, kk;
							CL::staged(lane, i, tt, kk);
							const int et = __shfl_sync(kFull, int(w_nxt.x), min(3 * tt, 31));
							R[i] = (tt < kSlots && et >= 0) ? rec[size_t(et) * RECQ + kk] : 0.0;
						}
					}
					if (active && w.x != 0xffffffffu)
					{
						const int ri = (w.w >> 16) & 0xff;
						double acc[NL][3];
#pragma unroll
						for (int j = 0; j < NL; ++j)
							acc[j][0] = acc[j][1] = acc[j][2] = 0.0;
						column_of_element<NL, NQ, P2S, false>(stage + slot * SSTR, s_rg, ri, mm, ConstTable<SLOT>(), acc, g_acc);
						// column component n = (mm + shift) % 3 of column node j goes to row 3*k_j + n of this lane's column
						const int n1 = mm == 2 ? 0 : mm + 1, n2 = mm == 0 ? 2 : mm - 1;
#pragma unroll
						for (int j = 0; j < NL; ++j)
						{
							const uint32_t word = j < 4 ? w.y : (j < 8 ? w.z : w.w);
							const int k3 = 3 * int((word >> (8 * (j & 3))) & 0xffu);
							strip[(k3 + mm) * 32] += acc[j][0];
							strip[(k3 + n1) * 32] += acc[j][1];
							strip[(k3 + n2) * 32] += acc[j][2];
						}
					}
					__syncwarp(); // the stage is overwritten at the top of the next step
				}
				// column 3b+mm of values[] starts at 9*adj_off[b] + mm*3*deg(b) and has 3*deg(b) rows
				const int off = b >= 0 ? m.adj_off[b] : 0, deg = b >= 0 ? m.adj_off[b + 1] - off : 0;
				if (b >= 0 && a.grad != nullptr)
					a.grad[size_t(b) * 3 + mm] = a.scale * g_acc;
				if (a.values != nullptr)
				{
					if constexpr (CL::STAGE >= 32 * kFlushLd && PFA_CL_COOP_FLUSH)
					{
						// cooperative flush: a lane-private store touches 30 sectors per instruction. Blocks of 32 rows go through
						// the (now idle) record stage with leading dimension 33 - written by columns, read by rows, both
						// conflict-free - so that one store instruction writes 32 consecutive rows of ONE column
						for (int r0 = 0; r0 < rows; r0 += 32)
						{
#pragma unroll 4
							for (int rr = 0; rr < 32; ++rr)
								if (r0 + rr < rows)
									stage[rr * kFlushLd + lane] = strip[(r0 + rr) * 32];
							__syncwarp();
							for (int c = 0; c < 3 * kSlots; ++c)
							{
								const int off_c = __shfl_sync(kFull, off, c), deg_c = __shfl_sync(kFull, deg, c);
								const int r = r0 + lane;
								if (r < 3 * deg_c)
									a.values[size_t(off_c) * 9 + size_t(c % 3) * 3 * deg_c + r] = a.scale * stage[lane * kFlushLd + c];
							}
							__syncwarp();
						}
					}
					else if (b >= 0)
					{
						double *dst = a.values + (size_t(off) * 9 + size_t(mm) * 3 * deg);
						for (int r = 0; r < 3 * deg; ++r)
							dst[r] = a.scale * strip[r * 32];
					}
				}
			}
		}

		std::mutex g_cl_mutex;
		double g_cl_shadow[16][2][kSlotDoubles];
		bool g_cl_valid[16][2] = {};

		cudaError_t ensure_cl_table(const DeviceMesh &m, int slot, cudaStream_t st)
		{
			int dev = 0;
			cudaError_t err = cudaGetDevice(&dev);
			if (err != cudaSuccess)
				return err;
			if (dev < 0 || dev >= 16 || m.ref_grads_host == nullptr)
				return cudaErrorInvalidValue;
			const size_t bytes = sizeof(double) * size_t(m.n_qp) * m.n_loc * 3;
			std::lock_guard<std::mutex> lock(g_cl_mutex);
			if (g_cl_valid[dev][slot] && std::memcmp(g_cl_shadow[dev][slot], m.ref_grads_host, bytes) == 0)
				return cudaSuccess;
			std::memcpy(g_cl_shadow[dev][slot], m.ref_grads_host, bytes);
			g_cl_valid[dev][slot] = true;
			return cudaMemcpyToSymbolAsync(c_cl_refgrad, g_cl_shadow[dev][slot], bytes, sizeof(double) * size_t(slot) * kSlotDoubles, cudaMemcpyHostToDevice, st);
		}

		template <int NL, int NQ, int SLOT, bool P2S>
		cudaError_t launch_cl(const DeviceMesh &m, const AssembleArgs &a, const ColumnLaneTables &t, int sm_count, cudaStream_t st)
		{
			cudaError_t err = ensure_cl_table(m, SLOT, st);
			if (err != cudaSuccess)
				return err;
			const int64_t threads = int64_t(m.n_el) * NQ;
			const unsigned rec_blocks = unsigned((threads + 127) / 128);
			collane_records_kernel<NL, NQ><<<rec_blocks, 128, 0, st>>>(m, a, t.records, t.block_energy);
			if ((err = cudaGetLastError()) != cudaSuccess)
				return err;
			if (a.energy != nullptr)
			{
				collane_energy_kernel<<<1, 1024, 0, st>>>(t.block_energy, int(rec_blocks), a.scale, a.energy);
				if ((err = cudaGetLastError()) != cudaSuccess)
					return err;
			}
			if (a.values == nullptr && a.grad == nullptr)
				return cudaSuccess;
			auto kern = collane_columns_kernel<NL, NQ, SLOT, P2S>;
			int dev = 0, smem_max = 0;
			if ((err = cudaGetDevice(&dev)) != cudaSuccess || (err = cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess)
				return err;
			if ((err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max)) != cudaSuccess)
				return err;
			const size_t table_bytes = sizeof(double) * size_t((NQ * NL * 3 + 1) & ~1);
			int g0 = 0;
			for (int c = 0; c < 2; ++c)
			{
				const int ng = t.n_groups[c];
				if (ng > 0)
				{
					const size_t warp_bytes = ColLayout<NQ>::warp_bytes(t.rows_max[c]); // record stage + strips
					int warps = int((size_t(smem_max) - table_bytes) / warp_bytes);
					if (warps < 1)
						return cudaErrorInvalidConfiguration;
					warps = std::min(warps, 8);
					const size_t smem = table_bytes + size_t(warps) * warp_bytes;
					int per_sm = 1; // small strips (P1) leave room for more than one CTA per SM
					if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem)) != cudaSuccess)
						return err;
					const int grid = std::max(1, std::min((ng + warps - 1) / warps, sm_count * std::max(per_sm, 1)));
					kern<<<grid, warps * 32, smem, st>>>(m, a, t, t.records, g0, g0 + ng, t.rows_max[c]);
					if ((err = cudaGetLastError()) != cudaSuccess)
						return err;
				}
				g0 += ng;
			}
			return cudaSuccess;
		}
	} // namespace

	bool column_lane_applies(int material, int n_loc, int n_qp)
	{
		return material == PFA_NEOHOOKEAN && ((n_loc == 4 && n_qp == 1) || (n_loc == 10 && n_qp == 4));
	}

	cudaError_t launch_column_lane(const DeviceMesh &m, const AssembleArgs &a, const ColumnLaneTables &t, int sm_count, cudaStream_t st)
	{
		if (m.n_loc == 4)
			return launch_cl<4, 1, 0, false>(m, a, t, sm_count, st);
		// the structured column step needs the structural zeros of the P2 reference gradients (DeviceMesh::p2_structured)
		return m.p2_structured ? launch_cl<10, 4, 1, true>(m, a, t, sm_count, st) : launch_cl<10, 4, 1, false>(m, a, t, sm_count, st);
	}
} // namespace pfa
