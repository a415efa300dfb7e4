// Third set of machine micro-benchmarks: the numbers a combining (shared-memory) scatter design
// for the P2 Hessian depends on. Not run in round 1 (written after the GPU budget
// was spent); build and run next round:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench3 tools/microbench3.cu
//   ./tools/microbench3 > profiles/microbench3_rNN.jsonl
//
//  red_pattern   RED.F64 sector-operation rate with a precomputed index stream (no index arithmetic
//                in the timed loop): runs of L doubles, aligned / unaligned to 32-byte sectors,
//                working set L2-resident (64 MB) or DRAM-sized (4 GB)
//  smem_rmw      plain shared-memory read-modify-write of doubles, throughput form (8 independent
//                updates in flight per lane): conflict-free (bank = lane), random, runs of 3
//  smem_atomic   atomicAdd(double) on shared memory with the same three patterns
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

// each thread issues `per_thread` REDs to idx[t + k * n_threads] (coalesced index loads)
__global__ void red_stream_kernel(double *__restrict__ dst, const uint32_t *__restrict__ idx, int per_thread, size_t n_threads)
{
	const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
	for (int k = 0; k < per_thread; ++k)
		asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst + idx[t + size_t(k) * n_threads]), "d"(1.0) : "memory");
}

// mode 0: plain RMW, 1: atomicAdd; pattern 0: own bank (conflict-free), 1: random, 2: runs of 3 at random row positions
__global__ void smem_update_kernel(double *out, int iters, int mode, int pattern, int n_acc)
{
	extern __shared__ double acc[];
	for (int i = threadIdx.x; i < n_acc; i += blockDim.x)
		acc[i] = 0.0;
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int warps = blockDim.x >> 5;
	const int per_warp = n_acc / warps; // warp-private slice: no cross-warp races for the plain RMW
	double *mine = acc + warp * per_warp;
	uint32_t s = (blockIdx.x * 1315423911u) ^ (threadIdx.x * 2654435761u) ^ 12345u;
	for (int it = 0; it < iters; ++it)
	{
		int off[8];
#pragma unroll
		for (int u = 0; u < 8; ++u)
		{
			s = s * 1664525u + 1013904223u;
			const uint32_t r = s >> 8;
			if (pattern == 0)
				off[u] = int((r % uint32_t(per_warp / 32)) * 32 + lane); // bank pair = lane
			else if (pattern == 1)
				off[u] = int(r % uint32_t(per_warp));
			else
				off[u] = int(((__shfl_sync(0xffffffffu, r, lane - lane % 3) % uint32_t(per_warp / 3 - 1)) * 3) + lane % 3);
		}
		if (mode == 0)
		{
			double v[8];
#pragma unroll
			for (int u = 0; u < 8; ++u)
				v[u] = mine[off[u]];
#pragma unroll
			for (int u = 0; u < 8; ++u)
				mine[off[u]] = v[u] + 1.0; // (duplicates inside one batch lose an update: timing only)
		}
		else
		{
#pragma unroll
			for (int u = 0; u < 8; ++u)
				atomicAdd(mine + off[u], 1.0);
		}
		__syncwarp();
	}
	__syncthreads();
	double sum = 0;
	for (int i = threadIdx.x; i < n_acc; i += blockDim.x)
		sum += acc[i];
	if (sum == -1.0)
		out[0] = sum;
}

int main(int argc, char **argv)
{
	const bool quick = argc > 1; // any argument: the short list (one working set, fewer run lengths)
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, 0));
	const int sms = prop.multiProcessorCount;
	printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);

	// ---- RED patterns ----
	const size_t n_threads = size_t(sms) * 8 * 256;
	const int per_thread = 64;
	uint32_t *d_idx;
	CK(cudaMalloc(&d_idx, n_threads * per_thread * sizeof(uint32_t)));
	std::vector<uint32_t> h_idx(n_threads * per_thread);
	for (size_t ws_mb : {size_t(64), size_t(4096)})
	{
		if (quick && ws_mb != 64)
			continue;
		const size_t n_doubles = ws_mb * 1024 * 1024 / 8;
		double *d_dst;
		CK(cudaMalloc(&d_dst, n_doubles * sizeof(double)));
		CK(cudaMemset(d_dst, 0, n_doubles * sizeof(double)));
		for (int aligned = 0; aligned < 2; ++aligned)
			for (int L : {1, 3, 4, 8, 32})
			{
				if (aligned && (L == 1 || L == 3))
					continue;
				if (quick && !((L == 3 && !aligned) || (L == 4 && aligned) || (L == 32 && aligned) || (L == 1)))
					continue;
				// consecutive lanes of a warp form runs of L doubles; runs start at random positions inside a
				// 2 KB window that moves with the warp (the locality of one column block)
				for (size_t k = 0; k < size_t(per_thread); ++k)
					for (size_t w = 0; w < n_threads / 32; ++w)
					{
						const size_t window = (size_t(rnd()) % (n_doubles / 256)) * 256;
						for (int l0 = 0; l0 < 32; l0 += L)
						{
							size_t start = window + rnd() % (256 - 32);
							if (aligned)
								start &= ~size_t(3);
							for (int l = l0; l < l0 + L && l < 32; ++l)
								h_idx[k * n_threads + w * 32 + l] = uint32_t(start + (l - l0));
						}
					}
				CK(cudaMemcpy(d_idx, h_idx.data(), h_idx.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
				const float ms = best_ms([&] { red_stream_kernel<<<unsigned(n_threads / 256), 256>>>(d_dst, d_idx, per_thread, n_threads); });
				const double ops = double(n_threads) * per_thread;
				printf("{\"bench\": \"red_pattern\", \"L\": %d, \"sector_aligned\": %d, \"working_set_MB\": %zu, \"ms\": %.3f, \"Gops\": %.1f}\n", L, aligned, ws_mb, ms,
					   ops / ms * 1e-6);
			}
		CK(cudaFree(d_dst));
	}
	CK(cudaFree(d_idx));

	// ---- shared-memory accumulation ----
	double *d_out;
	CK(cudaMalloc(&d_out, 8));
	const int n_acc = 12288; // 96 KB
	CK(cudaFuncSetAttribute(smem_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, n_acc * 8));
	for (int mode = 0; mode < 2; ++mode)
		for (int pattern = 0; pattern < 3; ++pattern)
			for (int warps : {4, 8, 16})
			{
				if (quick && warps == 4)
					continue;
				const int iters = 2000;
				const float ms = best_ms([&] { smem_update_kernel<<<sms * 2, warps * 32, n_acc * 8>>>(d_out, iters, mode, pattern, n_acc); });
				const double upd = double(sms) * 2 * warps * 32 * iters * 8;
				printf("{\"bench\": \"%s\", \"pattern\": \"%s\", \"warps_per_cta\": %d, \"ctas_per_sm\": 2, \"ms\": %.3f, \"updates_per_clk_per_sm\": %.2f, \"G_updates_per_s\": %.1f}\n",
					   mode == 0 ? "smem_rmw" : "smem_atomic", pattern == 0 ? "own_bank" : (pattern == 1 ? "random" : "runs_of_3"), warps, ms,
					   upd / (ms * 1e-3) / (double(prop.clockRate) * 1e3) / sms, upd / ms * 1e-6);
			}
	return 0;
}
