// Owner-computes (column-lane) NeoHookean assembly for affine P1 / P2 tets: the default path of pfa_grad_hess / pfa_hessian
// for the full matrix. Math, record layout and schedule: pfa_collane2.h (also compiled into the CPU emulation
// tests/collane2_emul.cpp, which is checked against the oracle).
//
//   cl2_records_kernel   thread <-> element: gathers x, writes the element record (54 doubles for P2: A = F J^-1, c1t, c2t,
//                        mu*da per quadrature point and K = J J^T) to global memory through shared memory (coalesced),
//                        reduces the energy in a fixed order
//   cl2_columns_kernel   one warp per CTA <-> chunk of node groups. A group is 5 nodes; node slot s is served by lane
//                        triples s (lanes 3s..3s+2) and 5+s (lanes 16+3s..), one lane per column component, each triple
//                        on its own incident element per step. The 10 element records of a step are fetched by TMA
//                        (cp.async.bulk, one copy per triple, completion on an mbarrier, two steps ahead in a
//                        double-buffered stage); every lane adds its 3*NL entries to the strip column of its dof
//                        (shared memory, address row*16 + column: bank = column; half-warp 0 updates before half-warp 1;
//                        the first contribution to a row is a store, so strips are never cleared). After the last step
//                        the strip is the finished CSC column: 16-row blocks are transposed through a 16 x 17 buffer
//                        and written with coalesced stores; the gradient entry is stored. Every output is written
//                        exactly once, in a fixed summation order (bitwise reproducible), no atomics, no zero fill.
#include "pfa_collane2.h"
#include "pfa_internal.h"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace pfa
{
	namespace
	{
		using namespace cl2;
		constexpr int kSlotDoubles = 4 * 10 * 3;
		__constant__ double c_cl2_refgrad[2][kSlotDoubles]; // slot 0: P1 [1][4][3], slot 1: P2 [4][10][3]
		constexpr unsigned kFull = 0xffffffffu;
		// flush: blocks of kFlushRows strip rows x 15 columns are transposed through shared memory; with the odd leading dimension
		// 15 the column-wise writes (16 lanes, consecutive columns) and the row-wise reads (32 lanes, consecutive rows) are both
		// conflict-free
		constexpr int kFlushRows = 32;
		constexpr int kTbLd = 15;

		template <int SLOT>
		struct ConstTable
		{
			__device__ __forceinline__ double operator[](int i) const { return c_cl2_refgrad[SLOT][i]; }
		};

		// ---------------------------------------------------------------- records
		template <int NL, int NQ, int SLOT>
		__global__ void __launch_bounds__(128) cl2_records_kernel(const DeviceMesh m, const AssembleArgs a, const int n_own, double *__restrict__ rec_out, double *__restrict__ block_energy)
		{
			constexpr int RECD = Rec<NQ>::D;
			constexpr int LD = RECD | 1; // odd: conflict-free record-per-thread writes
			extern __shared__ __align__(16) double s_rec[];
			const int64_t e64 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
			const bool valid = e64 < m.n_el;
			double e_el = 0.0;
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
				const size_t mi = size_t(e) * m.mat_stride;
				e_el = element_record<NL, NQ>(J, m.detj[e], m.qweights, m.lambda + mi, m.mu + mi, m.mat_stride, u, ConstTable<SLOT>(), s_rec + size_t(threadIdx.x) * LD);
				if (e >= n_own)
					e_el = 0.0; // ghost element (multi-GPU): its record is needed, its energy belongs to another rank
				else if (a.energy_per_el != nullptr)
					a.energy_per_el[m.elem_id ? m.elem_id[e] : e] = e_el;
			}
			// the 32 records of a warp are contiguous in global memory: coalesced stores
			__syncwarp();
			{
				const int lane = threadIdx.x & 31;
				const int64_t e0 = e64 - lane;
				const int n_valid = int(min(int64_t(32), int64_t(m.n_el) - e0));
				const double *src = s_rec + size_t(threadIdx.x - lane) * LD;
				double *dst = rec_out + size_t(e0) * RECD;
				for (int idx = lane; idx < n_valid * RECD; idx += 32)
					dst[idx] = src[(idx / RECD) * LD + idx % RECD];
			}
			if (a.energy != nullptr)
			{
				double w = e_el;
#pragma unroll
				for (int o = 16; o > 0; o >>= 1)
					w += __shfl_xor_sync(kFull, w, o);
				__shared__ double s_e[4];
				if ((threadIdx.x & 31) == 0)
					s_e[threadIdx.x >> 5] = w;
				__syncthreads();
				if (threadIdx.x == 0) // per-block partial sums, added up in a fixed order by cl2_energy_kernel
					block_energy[blockIdx.x] = s_e[0] + s_e[1] + s_e[2] + s_e[3];
			}
		}

		// energy = scale * sum of the per-block partial sums, in a fixed order (one block)
		__global__ void __launch_bounds__(1024) cl2_energy_kernel(const double *__restrict__ block_energy, int n_blocks, double scale, double *__restrict__ energy)
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

		// ---------------------------------------------------------------- TMA / mbarrier primitives (sm_90+ PTX)
		__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
		__device__ __forceinline__ void mbar_init(uint32_t bar, int count)
		{
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
		}
		__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
		{
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
		}
		__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
		{
			uint32_t ok;
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
						 : "=r"(ok)
						 : "r"(bar), "r"(parity)
						 : "memory");
			return ok != 0;
		}
		// global -> shared bulk copy (TMA), bytes a multiple of 16, both addresses 16-byte aligned; completion = complete_tx on the mbarrier
		[[maybe_unused]] __device__ __forceinline__ void tma_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
		{
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
						 : "memory");
		}

		// the same, predicated (no branch around the copy)
		[[maybe_unused]] __device__ __forceinline__ void tma_load_if(bool pred, uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
		{
			asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(dst),
						 "l"(src), "r"(bytes), "r"(bar), "r"(uint32_t(pred))
						 : "memory");
		}
		// true in exactly one (the lowest active) lane of the warp
		[[maybe_unused]] __device__ __forceinline__ bool elect_one()
		{
			uint32_t e;
			asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(e));
			return e != 0;
		}

		// Ampere-style asynchronous copy (LDGSTS): 16 bytes global -> shared without a register round trip; L2 only (.cg)
		__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
		{
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
		}
		__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
		template <int N>
		__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

		// Record staging, measured on B200 (profiles/clvar_r02f.jsonl): TMA bulk copies (one elected issue per triple, completion
		// on an mbarrier) and cp.async (16 bytes per lane and instruction) are equally fast for the 432-byte P2 records
		// (5.03 / 5.07 ms at cfg 3); for the 144-byte P1 records cp.async wins (0.224 / 0.263 ms at cfg 2). PFA_CL2_TMA
		// overrides the per-element-type choice (experiments).
#ifndef PFA_CL2_TMA_ELECT
#define PFA_CL2_TMA_ELECT 1 // 1: one elected lane issues all TMA copies of a step (see issue()); 0: every triple leader issues its own
#endif
#ifdef PFA_CL2_TMA
		template <int NQ>
		constexpr bool kUseTma = PFA_CL2_TMA != 0;
#else
		template <int NQ>
		constexpr bool kUseTma = NQ > 1;
#endif

#ifndef PFA_CL2_P1_STAGES
#define PFA_CL2_P1_STAGES 4 // record buffers of the P1 kernel (experiments: 6, 8)
#endif
		// shared memory of the one-warp CTA, in doubles
		template <int NL, int NQ>
		struct WarpLayout
		{
			static constexpr int RECD = Rec<NQ>::D;
			// slot stride of a record inside a buffer (16-byte aligned: RECD is even). (Padding the 144-byte P1 record to 176 bytes
			// so that the 16-byte cp.async writes of the ten triples spread over the bank groups - 4 instead of 20 wavefronts per
			// instruction in ncu - did not change the P1 time: 0.203 / 0.202 ms at cfg 2, profiles/clvar_r02s.jsonl.)
			static constexpr int SSTR = RECD;
			static constexpr int STAGE = kTriples * SSTR; // one buffer: 10 element records
			// record buffers = how many steps ahead the copies are issued: 2 for the 432-byte P2 records (shared memory is what
			// limits the resident warps there), 4 for the 144-byte P1 records, whose steps are too short to cover the latency of the
			// schedule-word load and of the copy with two
			static constexpr int STAGES = NQ > 1 ? 2 : PFA_CL2_P1_STAGES;
			static constexpr int TB = kFlushRows * kTbLd; // transposition block of the flush
			// the flush of a group runs after the last step of the group has consumed its records and before that buffer is
			// refilled: when a buffer is large enough (P2) the transposition block lives there
			static constexpr bool TB_ALIAS = STAGE >= TB;
			static constexpr int OFF_STAGE = (STAGES + 1) & ~1; // after the mbarriers (one per buffer)
			static constexpr int OFF_TB = OFF_STAGE + STAGES * STAGE;
			static constexpr int OFF_RG = OFF_TB + (TB_ALIAS ? 0 : ((TB + 1) & ~1));
			// rows of the own-node gradient table [ri][q][4]: with 4 points a row is 128 bytes = all 32 banks, so the triples of a
			// half-warp (different ri) would collide on every load; two doubles of padding shift consecutive rows by 4 banks
			// (ncu: 5.9 M excess wavefronts of 98.5 M in the edge launch at n = 40 without it)
#ifndef PFA_CL2_RG_PAD
#define PFA_CL2_RG_PAD 2
#endif
			static constexpr int RG_LD = NQ * 4 + (NQ > 1 ? PFA_CL2_RG_PAD : 0);
			static constexpr int OFF_INFO = OFF_RG + NL * RG_LD; // 5 x 4 ints = 10 doubles
			static constexpr int OFF_STRIP = OFF_INFO + 10;
			static_assert(RECD % 2 == 0 && OFF_STAGE % 2 == 0 && OFF_RG % 2 == 0 && OFF_STRIP % 2 == 0, "16-byte alignment");
			static size_t bytes(int strip_rows) { return sizeof(double) * (size_t(OFF_STRIP) + size_t(strip_rows) * kStripLd); }
		};

#ifndef PFA_CL2_MINBLOCKS
#define PFA_CL2_MINBLOCKS 1 // CTAs (= warps) per SM the register allocation must allow (experiments: 12 caps at 168 registers)
#endif
		template <int NL, int NQ, int SLOT, int MODE>
		__global__ void __launch_bounds__(32, PFA_CL2_MINBLOCKS) cl2_columns_kernel(const DeviceMesh m, const AssembleArgs a, const ColumnLane2Tables t, const int cls, const int chunk_begin,
																 const int chunk_end, const int strip_rows)
		{
			using L = WarpLayout<NL, NQ>;
			constexpr int RECD = L::RECD;
			constexpr uint32_t REC_BYTES = RECD * sizeof(double);
			constexpr bool TMA = kUseTma<NQ>;
			extern __shared__ __align__(16) double smem[];
			const int lane = threadIdx.x;
			const int half = lane >> 4, within = lane & 15;
			const bool active = within < 15;
			const int ns = active ? within / 3 : 0;
			const int mm = within - 3 * (within / 3);
			const int tr = half * kNodes + ns;
			const bool leader = active && mm == 0;
			// strip rows (times the leading dimension) of my three entries per row position: components mm, mm+1, mm+2 (mod 3)
			const int o0 = mm * kStripLd, o1 = (mm == 2 ? 0 : mm + 1) * kStripLd, o2 = (mm == 0 ? 2 : mm - 1) * kStripLd;
			double *stage = smem + L::OFF_STAGE;
			double *s_rg = smem + L::OFF_RG;
			int *s_info = reinterpret_cast<int *>(smem + L::OFF_INFO);
			double *strip = smem + L::OFF_STRIP + within;
			constexpr int D = L::STAGES;
			const uint32_t bar0 = smem_u32(smem); // mbarrier of buffer b at bar0 + 8 b
			if (lane == 0)
			{
				for (int b = 0; b < D; ++b)
					mbar_init(bar0 + 8 * b, 1);
				asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
			}
			// own-node reference gradients, padded rows [ri][q][4]
			for (int k = lane; k < NL * NQ * 4; k += 32)
				s_rg[(k / (NQ * 4)) * L::RG_LD + k % (NQ * 4)] = t.rg_padded[k];
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			__syncwarp();

			const uint4 *inc = reinterpret_cast<const uint4 *>(t.inc);
			const uint4 idle = make_uint4(kIdle, 0u, 0u, 0u);
			const uint32_t stage_u32 = smem_u32(stage);
			// records of the busy triples of a step -> buffer `buf`
			auto issue = [&](const uint4 &w, int buf, bool real) {
				if constexpr (TMA)
				{
					if (!real)
						return;
#if PFA_CL2_TMA_ELECT
					// Every lane learns the ten elements of the step by shuffles from the triple leaders (constant source lanes: the
					// results are warp-uniform, the compiler keeps them in uniform registers), then ONE elected lane arms the barrier
					// and issues the up to ten copies back to back. (When each leader issues its own copy the compiler has to
					// serialise the leaders in a loop - ELECT, three R2UR.BROADCAST, UBLKCP, BRA.U.ANY, about 58 cycles per copy in
					// the ncu source view of the edge launch at cfg 3, profiles/cl2_r02al_phases_edge.txt - because UBLKCP takes its
					// operands from uniform registers.) Measured at cfg 3: 4.881 -> 4.796 ms (profiles/clvar_r02am.jsonl). The P1 kernel
					// keeps cp.async: TMA with the elected issue is 0.224 against 0.203 ms at cfg 2 (same file).
					const unsigned mask = __ballot_sync(kFull, leader && w.x != kIdle);
					uint32_t el[kTriples];
#pragma unroll
					for (int k = 0; k < kTriples; ++k)
						el[k] = __shfl_sync(kFull, w.x, (k / kNodes) * 16 + (k % kNodes) * 3);
					if (elect_one())
					{
						const uint32_t bar = bar0 + 8 * buf;
						mbar_arrive_expect_tx(bar, uint32_t(__popc(mask)) * REC_BYTES);
#pragma unroll
						for (int k = 0; k < kTriples; ++k)
							tma_load_if(el[k] != kIdle, stage_u32 + uint32_t(buf * L::STAGE + k * L::SSTR) * 8u, t.records + size_t(el[k]) * RECD, REC_BYTES, bar); // (address unused when idle)
					}
					return;
#else
					// one TMA copy per busy triple; lane 0 arms the barrier with the byte count first
					const bool want = leader && w.x != kIdle;
					const unsigned mask = __ballot_sync(kFull, want);
					const uint32_t bar = bar0 + 8 * buf;
					if (lane == 0)
						mbar_arrive_expect_tx(bar, uint32_t(__popc(mask)) * REC_BYTES);
					__syncwarp();
					if (want)
						tma_load(stage_u32 + uint32_t(buf * L::STAGE + tr * L::SSTR) * 8u, t.records + size_t(w.x) * RECD, REC_BYTES, bar);
#endif
				}
				else
				{
					// the three lanes of a triple copy its record, 16 bytes per lane and instruction (interleaved: consecutive lanes,
					// consecutive chunks); every lane commits one group per step
					if (real && active && w.x != kIdle)
					{
						const char *src = reinterpret_cast<const char *>(t.records + size_t(w.x) * RECD) + mm * 16;
						const uint32_t dst = stage_u32 + uint32_t(buf * L::STAGE + tr * L::SSTR) * 8u + uint32_t(mm) * 16u;
#pragma unroll
						for (int i = 0; i < RECD / 6; ++i)
							cp_async16(dst + i * 48, src + i * 48);
					}
					cp_async_commit();
				}
			};

			unsigned it = 0; // steps this warp has consumed: buffer = it % D, barrier parity = (it / D) & 1
			for (;;)
			{
				int chunk = 0;
				if (lane == 0)
					chunk = chunk_begin + atomicAdd(t.counters + cls, 1);
				chunk = __shfl_sync(kFull, chunk, 0);
				if (chunk >= chunk_end)
					break;
				int g = t.chunk_off[chunk];
				const int g_end = t.chunk_off[chunk + 1];
				const int s_begin = t.grp_off[g], s_end = t.grp_off[g_end];
				const int4 *grp_info = reinterpret_cast<const int4 *>(t.grp_info);
				int g_last = t.grp_off[g + 1]; // first step after group g
				// node of my slot in the current group: (node, 9*adj_off, 3*deg, -); the next group's words are loaded one group ahead
				int4 info = active ? grp_info[size_t(g) * kNodes + ns] : make_int4(-1, 0, 0, 0);
				int rows_g = t.grp_rows[g];
				int4 info_n = (active && g + 1 < g_end) ? grp_info[size_t(g + 1) * kNodes + ns] : make_int4(-1, 0, 0, 0);
				int rows_n = g + 1 < g_end ? t.grp_rows[g + 1] : 0;
				int last_n = g + 1 < g_end ? t.grp_off[g + 2] : 0;
				// schedule words of this triple: W[0] = the current step ... W[D] = the step whose copies are issued during the current
				// one; the word of step s + D + 1 is loaded during step s, a whole step before it is needed
				uint4 W[D + 1];
#pragma unroll
				for (int k = 0; k <= D; ++k)
					W[k] = (active && s_begin + k < s_end) ? inc[size_t(s_begin + k) * kTriples + tr] : idle;
#pragma unroll
				for (int k = 0; k < D; ++k)
					issue(W[k], int((it + k) % D), s_begin + k < s_end);
				double g_acc = 0.0;
				for (int s = s_begin; s < s_end; ++s, ++it)
				{
					const int buf = int(it % D);
					const uint4 w_next = (active && s + D + 1 < s_end) ? inc[size_t(s + D + 1) * kTriples + tr] : idle;
					const uint4 w0 = W[0], w2 = W[D];
					const bool issue_real = s + D < s_end;
					const bool busy = active && w0.x != kIdle;
					// strip offsets of the NL row positions of this element (doubles, relative to my column); bit 7 of a position byte:
					// first contribution to these rows
					int ko[NL];
					unsigned fm;
#pragma unroll
					for (int j = 0; j < NL; ++j)
					{
						const uint32_t word = j < 4 ? w0.y : (j < 8 ? w0.z : w0.w);
						ko[j] = int(__byte_perm(word, 0u, 0x4440u + (j & 3)) & 0x7fu) * (3 * kStripLd);
					}
					// bit 7 of the four bytes of a word -> bits 0..3
					fm = (((w0.y >> 7) & 0x01010101u) * 0x01020408u) >> 24;
					if (NL > 4)
						fm |= ((((w0.z >> 7) & 0x01010101u) * 0x01020408u) >> 24) << 4 | ((((w0.w >> 7) & 0x00000101u) * 0x01020408u) >> 24) << 8;
					// half-warp 0 starts its sums from the strip (its update is then a plain store, issued before half-warp 1
					// reads); rows touched for the first time start from zero: the strips are never cleared
					const bool pre = busy && half == 0;
					double acc[NL][3];
#pragma unroll
					for (int j = 0; j < NL; ++j)
					{
						const bool ld = pre && !((fm >> j) & 1u);
						acc[j][0] = ld ? strip[ko[j] + o0] : 0.0;
						acc[j][1] = ld ? strip[ko[j] + o1] : 0.0;
						acc[j][2] = ld ? strip[ko[j] + o2] : 0.0;
					}
					if constexpr (TMA)
					{
						const uint32_t bar = bar0 + 8 * buf, parity = (it / D) & 1;
						while (!mbar_try_wait(bar, parity))
						{
						}
					}
					else
					{
						cp_async_wait<D - 1>(); // my copies of this step have landed (only the groups of later steps may be pending)
						__syncwarp();      // ... and so have the other lanes'
					}
					if (busy)
					{
						const int ri = (w0.w >> 16) & 0xff;
						column_of_element<NL, NQ, MODE>(stage + buf * L::STAGE + tr * L::SSTR, s_rg + ri * L::RG_LD, mm, ConstTable<SLOT>(), acc, g_acc, t.z4b, t.zbeta);
					}
					__syncwarp(); // every lane has read its record: the buffer can be refilled
					const bool group_ends = s + 1 == g_last;
					// (when the group ends here, the flush uses this buffer as its transposition block first)
					if (!(L::TB_ALIAS && group_ends))
						issue(w2, buf, issue_real);
					if (pre)
					{
#pragma unroll
						for (int j = 0; j < NL; ++j)
						{
							strip[ko[j] + o0] = acc[j][0];
							strip[ko[j] + o1] = acc[j][1];
							strip[ko[j] + o2] = acc[j][2];
						}
					}
					__syncwarp();
					if (busy && half == 1)
					{
#pragma unroll
						for (int jb = 0; jb < NL; jb += 5)
						{
							constexpr int JB = NL < 5 ? NL : 5;
							double old[JB][3];
#pragma unroll
							for (int jj = 0; jj < JB; ++jj)
							{
								const int j = jb + jj;
								const bool ld = !((fm >> j) & 1u);
								old[jj][0] = ld ? strip[ko[j] + o0] : 0.0;
								old[jj][1] = ld ? strip[ko[j] + o1] : 0.0;
								old[jj][2] = ld ? strip[ko[j] + o2] : 0.0;
							}
#pragma unroll
							for (int jj = 0; jj < JB; ++jj)
							{
								const int j = jb + jj;
								strip[ko[j] + o0] = old[jj][0] + acc[j][0];
								strip[ko[j] + o1] = old[jj][1] + acc[j][1];
								strip[ko[j] + o2] = old[jj][2] + acc[j][2];
							}
						}
					}
					__syncwarp();
#pragma unroll
					for (int k = 0; k < D; ++k)
						W[k] = W[k + 1];
					W[D] = w_next;
					if (group_ends)
					{
						// ---- group finished: gradient entries and the 15 columns ----
						const double g_tot = g_acc + __shfl_xor_sync(kFull, g_acc, 16);
						if (half == 0 && active && info.x >= 0 && a.grad != nullptr)
							a.grad[size_t(info.x) * 3 + mm] = a.scale * g_tot;
						if (half == 0 && active && mm == 0)
							reinterpret_cast<int4 *>(s_info)[ns] = info;
						__syncwarp();
						double *tb = L::TB_ALIAS ? stage + buf * L::STAGE : smem + L::OFF_TB;
						// phase 2 of the flush: this lane stores row `lane` of the 32-row block, for all 15 columns
						double *dst[15];
						int lim[15];
#pragma unroll
						for (int c = 0; c < 15; ++c)
						{
							const int4 nf = reinterpret_cast<const int4 *>(s_info)[c / 3];
							dst[c] = a.values + (size_t(nf.y) + size_t(c % 3) * nf.z + lane);
							lim[c] = nf.x >= 0 ? nf.z - lane : 0;
						}
						for (int r0 = 0; r0 < rows_g; r0 += kFlushRows)
						{
							// rows past the group's last row are never stored; rows past the allocation are not read. (Loading the strip
							// rows of the next block before the stores of this one, and the B-lane update in one batch of 30 instead of
							// two of 15, were measured together: +0.7 %, profiles/clvar_r02am.jsonl.)
							double v[16];
#pragma unroll
							for (int i = 0; i < 16; ++i)
								v[i] = r0 + 2 * i + half < strip_rows ? smem[L::OFF_STRIP + (r0 + 2 * i + half) * kStripLd + within] : 0.0;
							if (active)
							{
#pragma unroll
								for (int i = 0; i < 16; ++i)
									tb[(2 * i + half) * kTbLd + within] = v[i];
							}
							__syncwarp();
#pragma unroll
							for (int c = 0; c < 15; ++c)
								v[c] = tb[lane * kTbLd + c];
#pragma unroll
							for (int c = 0; c < 15; ++c)
								if (r0 < lim[c])
									__stcs(dst[c] + r0, a.scale * v[c]); // written once, never read here: keep the records in L2 instead
							__syncwarp();
						}
						if (L::TB_ALIAS)
						{
							if constexpr (TMA)
								asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic writes of the block before the TMA refill
							issue(w2, buf, issue_real);
						}
						g_acc = 0.0;
						++g;
						info = info_n;
						rows_g = rows_n;
						g_last = last_n;
						if (g + 1 < g_end)
						{
							info_n = active ? grp_info[size_t(g + 1) * kNodes + ns] : make_int4(-1, 0, 0, 0);
							rows_n = t.grp_rows[g + 1];
							last_n = t.grp_off[g + 2];
						}
					}
				}
			}
		}

		std::mutex g_cl2_mutex;
		double g_cl2_shadow[16][2][kSlotDoubles];
		bool g_cl2_valid[16][2] = {};

		cudaError_t ensure_table(const DeviceMesh &m, int slot, cudaStream_t st)
		{
			int dev = 0;
			cudaError_t err = cudaGetDevice(&dev);
			if (err != cudaSuccess)
				return err;
			if (dev < 0 || dev >= 16 || m.ref_grads_host == nullptr)
				return cudaErrorInvalidValue;
			const size_t bytes = sizeof(double) * size_t(m.n_qp) * m.n_loc * 3;
			std::lock_guard<std::mutex> lock(g_cl2_mutex);
			if (g_cl2_valid[dev][slot] && std::memcmp(g_cl2_shadow[dev][slot], m.ref_grads_host, bytes) == 0)
				return cudaSuccess;
			std::memcpy(g_cl2_shadow[dev][slot], m.ref_grads_host, bytes);
			g_cl2_valid[dev][slot] = true;
			return cudaMemcpyToSymbolAsync(c_cl2_refgrad, g_cl2_shadow[dev][slot], bytes, sizeof(double) * size_t(slot) * kSlotDoubles, cudaMemcpyHostToDevice, st);
		}

		template <int NL, int NQ, int SLOT, int MODE>
		cudaError_t launch_cl2(const DeviceMesh &m, const AssembleArgs &a, const ColumnLane2Tables &t, int sm_count, cudaStream_t st, int *launches)
		{
			cudaError_t err = ensure_table(m, SLOT, st);
			if (err != cudaSuccess)
				return err;
			constexpr int RECD = Rec<NQ>::D;
			const int n_rec = t.n_record_elements;
			const unsigned rec_blocks = unsigned((n_rec + 127) / 128);
			{
				auto rk = cl2_records_kernel<NL, NQ, SLOT>;
				const size_t rec_smem = sizeof(double) * 128 * size_t(RECD | 1);
				if ((err = cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, int(rec_smem))) != cudaSuccess)
					return err;
				DeviceMesh mr = m;
				mr.n_el = n_rec; // own + ghost elements
				rk<<<rec_blocks, 128, rec_smem, st>>>(mr, a, m.n_el, t.records, t.block_energy);
				if ((err = cudaGetLastError()) != cudaSuccess)
					return err;
				++*launches;
			}
			if (a.energy != nullptr)
			{
				cl2_energy_kernel<<<1, 1024, 0, st>>>(t.block_energy, int(rec_blocks), a.scale, a.energy);
				if ((err = cudaGetLastError()) != cudaSuccess)
					return err;
				++*launches;
			}
			if (a.values == nullptr)
				return cudaSuccess;
			if ((err = cudaMemsetAsync(t.counters, 0, 2 * sizeof(int), st)) != cudaSuccess)
				return err;
			auto kern = cl2_columns_kernel<NL, NQ, SLOT, MODE>;
			int dev = 0, smem_max = 0;
			if ((err = cudaGetDevice(&dev)) != cudaSuccess || (err = cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess)
				return err;
			if ((err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max)) != cudaSuccess)
				return err;
			// The two strip classes are two launches, one after the other. (Running them side by side on two streams with the SMs
			// shared 3 : 5 was measured slower on B200: 6.77 against 4.99 ms at cfg 3, profiles/clvar_r02m.jsonl.)
			static const int cap = [] { const char *v = std::getenv("PFA_CL_WARPS_PER_SM"); return v ? std::atoi(v) : 0; }(); // experiments
			int c0 = 0;
			for (int c = 0; c < 2; ++c)
			{
				const int nc = t.n_chunks[c];
				if (nc > 0)
				{
					const size_t smem = WarpLayout<NL, NQ>::bytes(t.rows_max[c]);
					if (smem > size_t(smem_max))
						return cudaErrorInvalidConfiguration;
					int per_sm = 1;
					if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem)) != cudaSuccess)
						return err;
					if (cap > 0)
						per_sm = std::min(per_sm, cap);
					const int grid = std::max(1, std::min(nc, sm_count * std::max(per_sm, 1)));
					kern<<<grid, 32, smem, st>>>(m, a, t, c, c0, c0 + nc, t.rows_max[c]);
					if ((err = cudaGetLastError()) != cudaSuccess)
						return err;
					++*launches;
				}
				c0 += nc;
			}
			return cudaSuccess;
		}
	} // namespace

	bool column_lane2_applies(int material, int n_loc, int n_qp)
	{
		return material == PFA_NEOHOOKEAN && ((n_loc == 4 && n_qp == 1) || (n_loc == 10 && n_qp == 4));
	}

	size_t column_lane2_record_doubles(int n_qp) { return size_t(n_qp) * kQpRec + 6; }

	cudaError_t launch_column_lane2(const DeviceMesh &m, const AssembleArgs &a, const ColumnLane2Tables &t, int sm_count, cudaStream_t st, int *launches)
	{
		if (m.n_loc == 4)
			return launch_cl2<4, 1, 0, 0>(m, a, t, sm_count, st, launches);
		// entry step: vertex-weighted sums when the table is the P2 basis on the symmetric 4-point rule (p2_rule_weights),
		// else the structural zeros of the P2 reference gradients (DeviceMesh::p2_structured), else the plain table
		static const bool no_z = [] { const char *v = std::getenv("PFA_CL_NO_P2Z"); return v && std::atoi(v) != 0; }(); // experiments
		// streamed form of the vertex-weighted step (MODE 3: 27 instead of 81 live doubles) is the default since the elected TMA issue:
		// 4.737 against 4.802 ms at cfg 3 (profiles/clvar_r02ap.jsonl); PFA_CL_P2Y=0 selects the held form (MODE 2)
		static const bool stream_z = [] { const char *v = std::getenv("PFA_CL_P2Y"); return v == nullptr || std::atoi(v) != 0; }();
		if (t.p2z && !no_z)
			return stream_z ? launch_cl2<10, 4, 1, 3>(m, a, t, sm_count, st, launches) : launch_cl2<10, 4, 1, 2>(m, a, t, sm_count, st, launches);
		return m.p2_structured ? launch_cl2<10, 4, 1, 1>(m, a, t, sm_count, st, launches) : launch_cl2<10, 4, 1, 0>(m, a, t, sm_count, st, launches);
	}
} // namespace pfa
