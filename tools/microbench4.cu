// Fourth set of machine micro-benchmarks: a MODEL of the inner loop of the "column lane" (owner-computes) design for
// the P2 NeoHookean Hessian, to be run before that kernel is written. Not run in round 1 (written after
// the GPU budget was spent); build and run next round:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench4 tools/microbench4.cu
//   ./tools/microbench4 > profiles/microbench4_rNN.jsonl
//
// What the model keeps of the real kernel, per warp step:
//   * 10 slots x 3 lanes (lane = slot*3 + m, 2 lanes idle): every slot works on ONE (element, node) incidence, so the warp
//     reads 10 different element records of 50 doubles (400 B: F~ per quadrature point 36, c1*da and c2*da per point 8,
//     pad 6) from global memory (index stream precomputed; records L2-resident or DRAM-sized);
//   * each lane builds its row operand Y[q][n][c] (4 points x 3 x 3) from the record (~30 DFMA per point);
//   * each lane produces the 30 entries (j, n) of its column with 12 DFMA each against reference gradients in
//     __constant__ memory (uniform index), and adds entry (j, n) to row 3*k_j + n of its lane-private strip in shared
//     memory (address = row*32 + lane: bank = lane, no atomics); k_j comes from a byte table (10 bytes per slot step);
//   * every `flush_every` steps the strip is streamed to global memory and cleared (the node group is finished).
// Reported: column entries per second for the whole GPU, against the 1.77 G entries one cfg-3 assembly needs, for strips
// of 72 rows (edge-midpoint nodes, 18 KB per warp) and 195 rows (vertex nodes, 49 KB per warp) and both record
// working sets. A variant without the strip update (entries summed into a register) and one without the entry DFMAs
// separate the two costs.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                              \
	do                                                                                     \
	{                                                                                      \
		cudaError_t e_ = (x);                                                              \
		if (e_ != cudaSuccess)                                                             \
		{                                                                                  \
			printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
			exit(1);                                                                       \
		}                                                                                  \
	} while (0)

template <typename F>
static float best_ms(F &&f, int reps = 5)
{
	cudaEvent_t a, b;
	CK(cudaEventCreate(&a));
	CK(cudaEventCreate(&b));
	f();
	CK(cudaDeviceSynchronize());
	float best = 1e30f;
	for (int r = 0; r < reps; ++r)
	{
		CK(cudaEventRecord(a));
		f();
		CK(cudaEventRecord(b));
		CK(cudaEventSynchronize(b));
		float ms;
		CK(cudaEventElapsedTime(&ms, a, b));
		best = ms < best ? ms : best;
	}
	return best;
}

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd()
{
	rng_state ^= rng_state << 13;
	rng_state ^= rng_state >> 7;
	rng_state ^= rng_state << 17;
	return uint32_t(rng_state >> 32);
}

constexpr int kRec = 50;  // doubles per element record
constexpr int kQ = 4;     // quadrature points (P2, order 2)
constexpr int kLoc = 10;  // local nodes
__constant__ double c_refgrad[kQ * kLoc * 3];

// MODE 0: full model; 1: no strip update (entries summed into a register); 2: strip update of a constant (no entry DFMAs)
template <int MODE>
__global__ void column_lane_model(const double *__restrict__ records, const int32_t *__restrict__ elem_of, // [warps_total][steps][10]
								  const uint32_t *__restrict__ kidx,                                          // [warps_total][steps][10][3] (10 bytes used)
								  int steps, int rows, int flush_every, double *__restrict__ out)
{
	extern __shared__ double strips[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int warps_per_cta = blockDim.x >> 5;
	const size_t gw = size_t(blockIdx.x) * warps_per_cta + warp;
	double *strip = strips + size_t(warp) * rows * 32;
	for (int r = 0; r < rows; ++r)
		strip[r * 32 + lane] = 0.0;
	const int slot = lane / 3, m = lane - slot * 3;
	const bool active = lane < 30;
	double sink = 0.0;
	for (int s = 0; s < steps; ++s)
	{
		if (active)
		{
			const size_t base = (gw * steps + s) * 10 + slot;
			const double *rec = records + size_t(elem_of[base]) * kRec;
			const uint32_t k0 = kidx[base * 3 + 0], k1 = kidx[base * 3 + 1], k2 = kidx[base * 3 + 2];
			// row operand of this lane: Y[q][n][c]
			double Y[kQ][3][3];
#pragma unroll
			for (int q = 0; q < kQ; ++q)
			{
				const double c1 = rec[36 + q], c2 = rec[40 + q];
				double F[9];
#pragma unroll
				for (int t = 0; t < 9; ++t)
					F[t] = rec[q * 9 + t];
				// row m of F and column m of F by selects (a runtime index would put F in local memory)
				const double a0 = m == 0 ? F[0] : (m == 1 ? F[3] : F[6]);
				const double a1 = m == 0 ? F[1] : (m == 1 ? F[4] : F[7]);
				const double a2 = m == 0 ? F[2] : (m == 1 ? F[5] : F[8]);
				const double b0 = m == 0 ? F[0] : (m == 1 ? F[1] : F[2]);
				const double b1 = m == 0 ? F[3] : (m == 1 ? F[4] : F[5]);
				const double b2 = m == 0 ? F[6] : (m == 1 ? F[7] : F[8]);
#pragma unroll
				for (int n = 0; n < 3; ++n)
				{
					const double w = c1 * (a0 * F[n * 3 + 0] + a1 * F[n * 3 + 1] + a2 * F[n * 3 + 2]);
#pragma unroll
					for (int c = 0; c < 3; ++c)
						Y[q][n][c] = w * F[c * 3 + n] + c2 * F[((n + 1) % 3) * 3 + c] * a0 + (n == m ? c1 : 0.0) * (c == 0 ? b0 : (c == 1 ? b1 : b2));
				}
			}
#pragma unroll
			for (int j = 0; j < kLoc; ++j)
			{
				const uint32_t word = j < 4 ? k0 : (j < 8 ? k1 : k2);
				const int k = (word >> (8 * (j & 3))) & 0xff;
#pragma unroll
				for (int n = 0; n < 3; ++n)
				{
					double v = 1.0;
					if (MODE != 2)
					{
						v = 0.0;
#pragma unroll
						for (int q = 0; q < kQ; ++q)
						{
							const double *g = c_refgrad + (q * kLoc + j) * 3;
							v += Y[q][n][0] * g[0] + Y[q][n][1] * g[1] + Y[q][n][2] * g[2];
						}
					}
					else
						v = Y[j & 3][n][0];
					if (MODE == 1)
						sink += v;
					else
						strip[(k * 3 + n) * 32 + lane] += v;
				}
			}
		}
		if ((s + 1) % flush_every == 0)
		{
			__syncwarp();
			double *dst = out + gw * size_t(rows) * 32;
			for (int r = 0; r < rows; ++r)
			{
				dst[r * 32 + lane] = strip[r * 32 + lane];
				strip[r * 32 + lane] = 0.0;
			}
			__syncwarp();
		}
	}
	if (sink == 123.456)
		out[gw] = sink;
}

template <int MODE>
static void run(const char *name, int sms, int rows, size_t n_records, int steps)
{
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, 0));
	const size_t strip_bytes = size_t(rows) * 32 * sizeof(double);
	const size_t smem_budget = size_t(prop.sharedMemPerMultiprocessor) - 4096;
	int warps_per_sm = int(smem_budget / strip_bytes);
	if (warps_per_sm > 16)
		warps_per_sm = 16;
	if (warps_per_sm < 1)
	{
		printf("{\"bench\": \"%s\", \"rows\": %d, \"skipped\": \"strip larger than shared memory\"}\n", name, rows);
		return;
	}
	// one CTA per SM holding all its warps (<= 227 KB dynamic shared memory per CTA)
	int warps_per_cta = warps_per_sm;
	while (size_t(warps_per_cta) * strip_bytes > size_t(prop.sharedMemPerBlockOptin))
		--warps_per_cta;
	const int ctas = sms;
	const size_t warps_total = size_t(ctas) * warps_per_cta;
	const int k_nodes = rows / 3;

	std::vector<double> rec(n_records * kRec);
	for (auto &v : rec)
		v = 0.5 + (rnd() & 0xffff) / 65536.0;
	std::vector<int32_t> elem(warps_total * steps * 10);
	// locality like a Morton-ordered mesh: a warp walks a window of ~4096 consecutive records
	for (size_t w = 0; w < warps_total; ++w)
	{
		const size_t window = (size_t(rnd()) % (n_records > 4096 ? n_records - 4096 : 1));
		for (size_t t = 0; t < size_t(steps) * 10; ++t)
			elem[w * steps * 10 + t] = int32_t(window + rnd() % (n_records > 4096 ? 4096 : n_records));
	}
	std::vector<uint32_t> kidx(warps_total * steps * 10 * 3);
	for (auto &v : kidx)
	{
		uint32_t word = 0;
		for (int b = 0; b < 4; ++b)
			word |= (rnd() % k_nodes) << (8 * b);
		v = word;
	}
	std::vector<double> rg(kQ * kLoc * 3);
	for (auto &v : rg)
		v = (rnd() & 0xffff) / 65536.0 - 0.5;
	CK(cudaMemcpyToSymbol(c_refgrad, rg.data(), rg.size() * sizeof(double)));

	double *d_rec, *d_out;
	int32_t *d_elem;
	uint32_t *d_kidx;
	CK(cudaMalloc(&d_rec, rec.size() * sizeof(double)));
	CK(cudaMalloc(&d_elem, elem.size() * sizeof(int32_t)));
	CK(cudaMalloc(&d_kidx, kidx.size() * sizeof(uint32_t)));
	CK(cudaMalloc(&d_out, warps_total * size_t(rows) * 32 * sizeof(double)));
	CK(cudaMemcpy(d_rec, rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice));
	CK(cudaMemcpy(d_elem, elem.data(), elem.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
	CK(cudaMemcpy(d_kidx, kidx.data(), kidx.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
	const size_t smem = size_t(warps_per_cta) * strip_bytes;
	CK(cudaFuncSetAttribute(column_lane_model<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
	// a node group of a warp finishes after ~24 incidences per slot (vertex nodes) / ~5 (edge nodes)
	const int flush_every = rows > 100 ? 24 : 6;
	const float ms = best_ms([&] {
		column_lane_model<MODE><<<ctas, warps_per_cta * 32, smem>>>(d_rec, d_elem, d_kidx, steps, rows, flush_every, d_out);
	});
	CK(cudaGetLastError());
	const double entries = double(warps_total) * steps * 30.0 * 30.0;
	printf("{\"bench\": \"%s\", \"rows\": %d, \"strip_kb\": %.1f, \"warps_per_sm\": %d, \"records_mb\": %.1f, \"steps\": %d, \"flush_every\": %d, "
		   "\"ms\": %.4f, \"g_entries_per_s\": %.1f, \"cfg3_ms_at_this_rate\": %.2f}\n",
		   name, rows, strip_bytes / 1024.0, warps_per_cta, n_records * kRec * 8 / 1e6, steps, flush_every, ms, entries / ms * 1e-6,
		   1.774e9 / (entries / ms * 1e-6 * 1e9) * 1e3);
	fflush(stdout);
	CK(cudaFree(d_rec));
	CK(cudaFree(d_elem));
	CK(cudaFree(d_kidx));
	CK(cudaFree(d_out));
}

int main()
{
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, 0));
	const int sms = prop.multiProcessorCount;
	for (int rows : {72, 195})
		for (size_t n_records : {size_t(1) << 16, size_t(2) << 20})
		{
			const int steps = 384;
			run<0>("column_lane_full", sms, rows, n_records, steps);
			run<1>("column_lane_no_strip", sms, rows, n_records, steps);
			run<2>("column_lane_no_entry_math", sms, rows, n_records, steps);
		}
	return 0;
}
