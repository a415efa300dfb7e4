// Second set of machine micro-benchmarks (DESIGN.md §Measured machine limits), answering the
// questions the tile-accumulation design depends on:
//   smem_atomic   atomicAdd(double) on shared memory (CAS loop in SASS) for the scatter shape
//   smem_rmw      the same updates as plain load/add/store (warp-private accumulators)
//   dmma          mma.sync m8n8k4 f64 throughput, alone and next to DFMA (separate pipe or not?)
//   runs          RED / ST throughput against run length (3..32 doubles) and SM count
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o microbench2 microbench2.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                          \
	do                                                                                 \
	{                                                                                  \
		cudaError_t e = (x);                                                           \
		if (e != cudaSuccess)                                                          \
		{                                                                              \
			printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
			exit(1);                                                                   \
		}                                                                              \
	} while (0)

template <typename F>
float time_ms(F &&f, int reps = 3)
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

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
	x ^= x >> 16;
	x *= 0x7feb352dU;
	x ^= x >> 15;
	x *= 0x846ca68bU;
	x ^= x >> 16;
	return x;
}

// ---------------- shared-memory accumulation ----------------
// mode 0: atomicAdd(double) ; 1: plain ld/add/st (racy across warps, timing only) ;
// 2: 9 independent CAS updates interleaved by hand
__global__ void smem_acc_kernel(double *out, int n_acc, int iters, int mode)
{
	extern __shared__ double acc[];
	for (int i = threadIdx.x; i < n_acc; i += blockDim.x)
		acc[i] = 0.0;
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int runs = n_acc / 3;
	if (lane < 30)
	{
		for (int it = 0; it < iters; ++it)
		{
			if (mode == 2)
			{
				unsigned long long *p[9];
				unsigned long long old[9];
				bool done[9];
#pragma unroll
				for (int k = 0; k < 9; ++k)
				{
					const uint32_t r = hash32(uint32_t((it * 9 + k) * 977 + warp * 131 + lane / 3 + blockIdx.x * 7919)) % runs;
					p[k] = reinterpret_cast<unsigned long long *>(acc + r * 3 + lane % 3);
					old[k] = *p[k];
					done[k] = false;
				}
				bool all = false;
				while (!all)
				{
					all = true;
#pragma unroll
					for (int k = 0; k < 9; ++k)
						if (!done[k])
						{
							const unsigned long long nv = __double_as_longlong(__longlong_as_double(old[k]) + 1.0);
							const unsigned long long got = atomicCAS(p[k], old[k], nv);
							done[k] = got == old[k];
							old[k] = got;
							all = all && done[k];
						}
				}
			}
			else
			{
				const uint32_t r = hash32(uint32_t(it * 977 + warp * 131 + lane / 3 + blockIdx.x * 7919)) % runs;
				double *p = acc + r * 3 + lane % 3;
				if (mode == 0)
					atomicAdd(p, 1.0);
				else
					*p = *p + 1.0;
			}
		}
	}
	__syncthreads();
	double s = 0;
	for (int i = threadIdx.x; i < n_acc; i += blockDim.x)
		s += acc[i];
	if (s == 12345.678)
		out[0] = s;
}

// ---------------- DFMA / DMMA ----------------
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// kind 0: DFMA only, 1: DMMA only, 2: both in every warp, 3: even warps DFMA, odd warps DMMA
__global__ void fp64_mix_kernel(double *out, int iters, int kind, double a, double b)
{
	double x[8], c[8][2];
#pragma unroll
	for (int k = 0; k < 8; ++k)
	{
		x[k] = threadIdx.x + k;
		c[k][0] = k;
		c[k][1] = -k;
	}
	const int warp = threadIdx.x >> 5;
	const bool do_fma = kind == 0 || kind == 2 || (kind == 3 && (warp & 1) == 0);
	const bool do_mma = kind == 1 || kind == 2 || (kind == 3 && (warp & 1) == 1);
	for (int i = 0; i < iters; ++i)
	{
		if (do_fma)
		{
#pragma unroll
			for (int k = 0; k < 8; ++k)
				x[k] = fma(x[k], a, b);
		}
		if (do_mma)
		{
#pragma unroll
			for (int k = 0; k < 8; ++k)
				dmma(c[k][0], c[k][1], a, b);
		}
	}
	double s = 0;
#pragma unroll
	for (int k = 0; k < 8; ++k)
		s += x[k] + c[k][0] + c[k][1];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------- RED / ST against run length ----------------
// every warp-iteration touches one 2 KB window; lanes form runs of L consecutive doubles that
// start at pseudo-random (8-byte aligned) places of the window. op 0: RED, 1: ST
__global__ void runs_kernel(double *buf, size_t n_doubles, int ops_per_thread, int L, int op)
{
	const size_t tid = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
	const size_t nthreads = size_t(gridDim.x) * blockDim.x;
	const int lane = threadIdx.x & 31;
	const size_t warp = tid >> 5;
	for (int it = 0; it < ops_per_thread; ++it)
	{
		const size_t window = ((size_t(it) * (nthreads >> 5) + warp) * 256) % (n_doubles - 512);
		const uint32_t r = hash32(uint32_t(it * 131071 + warp * 31 + lane / L)) % (256 / L);
		const size_t idx = window + size_t(r) * L + lane % L;
		if (op == 0)
			atomicAdd(buf + idx, 1.0);
		else
			buf[idx] = 1.0;
	}
}

int main()
{
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, 0));
	int clk_khz = 0;
	cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
	const int sms = prop.multiProcessorCount;
	printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, clk_khz);
	double *out;
	CK(cudaMalloc(&out, sizeof(double) * 1024 * sms * 8));

	{
		const char *names[3] = {"smem_atomicAdd_f64", "smem_plain_rmw_f64", "smem_cas_f64_9wide"};
		const int n_acc = 20480; // 160 KB of accumulators
		CK(cudaFuncSetAttribute(smem_acc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, n_acc * 8));
		for (int mode = 0; mode < 3; ++mode)
			for (int warps = 4; warps <= 16; warps *= 2)
			{
				const int iters = mode == 2 ? 1024 : 8192;
				float ms = time_ms([&] { smem_acc_kernel<<<sms, warps * 32, n_acc * 8>>>(out, n_acc, iters, mode); });
				const double total = double(iters) * (mode == 2 ? 9 : 1) * warps * 30;
				printf("{\"bench\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.3f, \"updates_per_clk_per_sm\": %.3f, \"G_updates_per_s_chip\": %.1f}\n", names[mode], warps, ms,
					   total / (ms * 1e-3) / (clk_khz * 1e3), total * sms / ms * 1e-6);
			}
	}
	{
		const char *names[4] = {"dfma_only", "dmma_only", "dfma_and_dmma_same_warp", "dfma_warps_and_dmma_warps"};
		const int threads = 512, blocks = sms * 4, iters = 1 << 14;
		for (int kind = 0; kind < 4; ++kind)
		{
			float ms = time_ms([&] { fp64_mix_kernel<<<blocks, threads>>>(out, iters, kind, 1.0000001, 1e-9); });
			const double nthreads = double(threads) * blocks;
			double fma_flops = 0, mma_flops = 0;
			const double fw = kind == 3 ? 0.5 : 1.0;
			if (kind == 0 || kind >= 2)
				fma_flops = 2.0 * 8 * iters * nthreads * fw;
			if (kind >= 1)
				mma_flops = 2.0 * 8 * 8 * 4 * 8 * double(iters) * (nthreads / 32) * fw; // m8n8k4 = 256 FMA per warp instruction
			printf("{\"bench\": \"%s\", \"ms\": %.3f, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"sum_tflops\": %.2f}\n", names[kind], ms, fma_flops / ms * 1e-9,
				   mma_flops / ms * 1e-9, (fma_flops + mma_flops) / ms * 1e-9);
		}
	}
	{
		const size_t big = size_t(4) << 30;
		double *a;
		CK(cudaMalloc(&a, big));
		CK(cudaMemset(a, 0, big));
		const int Ls[5] = {3, 4, 8, 16, 32};
		for (int ws = 0; ws < 2; ++ws)
			for (int op = 0; op < 2; ++op)
				for (int li = 0; li < 5; ++li)
					for (int frac = 1; frac <= 2; ++frac)
					{
						if (frac == 2 && (ws == 1 || (Ls[li] != 3 && Ls[li] != 32)))
							continue;
						const size_t bytes = ws == 0 ? (size_t(48) << 20) : big;
						const int blocks = sms * 8 / frac, threads = 256, ops = 2048;
						float ms = time_ms([&] { runs_kernel<<<blocks, threads>>>(a, bytes / 8, ops, Ls[li], op); });
						const double total = double(ops) * blocks * threads;
						printf("{\"bench\": \"%s_runs\", \"L\": %d, \"working_set_MB\": %zu, \"blocks\": %d, \"ms\": %.3f, \"Gops\": %.1f, \"GBs_payload\": %.1f}\n", op ? "st" : "red", Ls[li],
							   bytes >> 20, blocks, ms, total / ms * 1e-6, total * 8 / ms * 1e-6);
					}
		CK(cudaFree(a));
	}
	return 0;
}
