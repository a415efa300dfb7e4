// Machine micro-benchmarks that ground the kernel design (DESIGN.md §Measured machine limits):
//   dfma      FP64 FMA peak on all SMs (the FP64 roofline denominator, not in MEASURED_PEAKS.json)
//   copy      HBM read+write and write-only bandwidth with 16-byte accesses
//   red       RED.ADD.F64 throughput for the access shapes of the Hessian scatter
//   bulkred   cp.reduce.async.bulk (TMA) f64 add from shared memory
//   l2        read bandwidth of an L2-resident buffer
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

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
float time_ms(F &&f, int reps = 5)
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

// ---------------- DFMA peak ----------------
__global__ void dfma_kernel(double *out, int iters, double a, double b)
{
	double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
	for (int i = 0; i < iters; ++i)
	{
		x0 = fma(x0, a, b);
		x1 = fma(x1, a, b);
		x2 = fma(x2, a, b);
		x3 = fma(x3, a, b);
		x4 = fma(x4, a, b);
		x5 = fma(x5, a, b);
		x6 = fma(x6, a, b);
		x7 = fma(x7, a, b);
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// ---------------- copy / fill ----------------
__global__ void copy_kernel(const double2 *__restrict__ in, double2 *__restrict__ out, size_t n)
{
	for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
		out[i] = in[i];
}
__global__ void fill_kernel(double2 *__restrict__ out, size_t n, double v)
{
	for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
		out[i] = make_double2(v, v);
}
__global__ void read_kernel(const double2 *__restrict__ in, size_t n, int passes, double *sink)
{
	double acc = 0;
	for (int p = 0; p < passes; ++p)
		for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
		{
			double2 v = in[i];
			acc += v.x + v.y;
		}
	if (acc == 123.456)
		*sink = acc;
}

// ---------------- RED.F64 shapes ----------------
__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
	x ^= x >> 16;
	x *= 0x7feb352dU;
	x ^= x >> 15;
	x *= 0x846ca68bU;
	x ^= x >> 16;
	return x;
}

// mode 0: warp-contiguous (lane l -> base + l), consecutive warps consecutive chunks
// mode 1: runs of 3 doubles at pseudo-random places inside a 2 KB window that slides (element scatter)
// mode 2: fully random addresses over the buffer
// mode 3: like 1 but plain (non-atomic) stores, to see the cost of 24-byte partial writes
__global__ void red_kernel(double *buf, size_t n_doubles, size_t n_ops_per_thread, int mode)
{
	const size_t tid = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
	const size_t nthreads = size_t(gridDim.x) * blockDim.x;
	const int lane = threadIdx.x & 31;
	const size_t warp = tid >> 5;
	for (size_t it = 0; it < n_ops_per_thread; ++it)
	{
		size_t idx;
		if (mode == 0)
			idx = ((it * (nthreads >> 5) + warp) * 32 + lane) % n_doubles;
		else if (mode == 1 || mode == 3)
		{
			// each warp-iteration works inside one 256-double window; lane -> (run = lane/3, m = lane%3)
			const size_t window = ((it * (nthreads >> 5) + warp) * 256) % (n_doubles - 256);
			const uint32_t r = hash32(uint32_t(it * 131071 + warp * 31 + lane / 3)) % 84; // run start /3
			idx = window + size_t(r) * 3 + lane % 3;
		}
		else
			idx = (size_t(hash32(uint32_t(tid * 2654435761u + it))) * 2654435761ull + it) % n_doubles;
		if (mode == 3)
			buf[idx] = 1.0;
		else
			atomicAdd(buf + idx, 1.0);
	}
}

// ---------------- TMA bulk reduce ----------------
__global__ void bulkred_kernel(double *buf, size_t n_doubles, int chunk_doubles, int iters)
{
	extern __shared__ __align__(128) double sm[];
	for (int i = threadIdx.x; i < chunk_doubles; i += blockDim.x)
		sm[i] = 1.0;
	__syncthreads();
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	if (threadIdx.x == 0)
	{
		const size_t chunks = n_doubles / chunk_doubles;
		for (int it = 0; it < iters; ++it)
		{
			const size_t c = (size_t(it) * gridDim.x + blockIdx.x) % chunks;
			double *dst = buf + c * chunk_doubles;
			uint32_t src = (uint32_t)__cvta_generic_to_shared(sm);
			asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(chunk_doubles * 8)
						 : "memory");
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
			if ((it & 7) == 7)
				asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
		}
		asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
	}
}

int main(int argc, char **argv)
{
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, 0));
	int clk_khz = 0;
	cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
	printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"l2_bytes\": %d}\n", prop.name, prop.multiProcessorCount, clk_khz, prop.l2CacheSize);
	const int sms = prop.multiProcessorCount;

	{ // DFMA
		double *out;
		const int threads = 512, blocks = sms * 4, iters = 1 << 16;
		CK(cudaMalloc(&out, sizeof(double) * threads * blocks));
		float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
		const double flops = 2.0 * 8 * double(iters) * threads * blocks;
		printf("{\"bench\": \"dfma\", \"ms\": %.3f, \"tflops\": %.2f, \"dfma_per_clk_per_sm_at_nominal\": %.1f}\n", ms, flops / ms * 1e-9,
			   flops / 2 / (ms * 1e-3) / sms / (clk_khz * 1e3));
		CK(cudaFree(out));
	}

	const size_t big = size_t(4) << 30; // 4 GiB buffers
	double *a, *b;
	CK(cudaMalloc(&a, big));
	CK(cudaMalloc(&b, big));
	CK(cudaMemset(a, 0, big));
	CK(cudaMemset(b, 0, big));
	{
		const size_t n2 = big / 16;
		float ms = time_ms([&] { copy_kernel<<<sms * 16, 512>>>((const double2 *)a, (double2 *)b, n2); });
		printf("{\"bench\": \"copy\", \"ms\": %.3f, \"GBs_read_plus_write\": %.1f}\n", ms, 2.0 * big / ms * 1e-6);
		ms = time_ms([&] { fill_kernel<<<sms * 16, 512>>>((double2 *)b, n2, 0.0); });
		printf("{\"bench\": \"fill\", \"ms\": %.3f, \"GBs_write\": %.1f}\n", ms, 1.0 * big / ms * 1e-6);
		ms = time_ms([&] { CK(cudaMemsetAsync(b, 0, big)); });
		printf("{\"bench\": \"cudaMemset\", \"ms\": %.3f, \"GBs_write\": %.1f}\n", ms, 1.0 * big / ms * 1e-6);
		double *sink;
		CK(cudaMalloc(&sink, 8));
		ms = time_ms([&] { read_kernel<<<sms * 16, 512>>>((const double2 *)a, n2, 1, sink); });
		printf("{\"bench\": \"read_hbm\", \"ms\": %.3f, \"GBs\": %.1f}\n", ms, 1.0 * big / ms * 1e-6);
		const size_t small = size_t(48) << 20;
		ms = time_ms([&] { read_kernel<<<sms * 16, 512>>>((const double2 *)a, small / 16, 20, sink); });
		printf("{\"bench\": \"read_l2_48MB\", \"ms\": %.3f, \"GBs\": %.1f}\n", ms, 20.0 * small / ms * 1e-6);
	}
	{
		const char *names[4] = {"red_contig", "red_runs3_in_2KB_window", "red_random", "store_runs3_in_2KB_window"};
		for (int ws = 0; ws < 2; ++ws)
		{
			const size_t bytes = ws == 0 ? (size_t(48) << 20) : big;
			const size_t nd = bytes / 8;
			for (int mode = 0; mode < 4; ++mode)
			{
				const int blocks = sms * 8, threads = 256;
				const size_t ops = 2048;
				float ms = time_ms([&] { red_kernel<<<blocks, threads>>>(a, nd, ops, mode); }, 3);
				const double total = double(ops) * blocks * threads;
				printf("{\"bench\": \"%s\", \"working_set_MB\": %zu, \"ms\": %.3f, \"Gops\": %.2f, \"GBs_payload\": %.1f, \"ops_per_clk_per_sm_at_nominal\": %.3f}\n", names[mode], bytes >> 20, ms,
					   total / ms * 1e-6, total * 8 / ms * 1e-6, total / (ms * 1e-3) / sms / (clk_khz * 1e3));
			}
		}
	}
	{
		for (int chunk = 256; chunk <= 4096; chunk *= 4) // doubles: 2 KB, 8 KB, 32 KB
		{
			const int iters = 512;
			CK(cudaFuncSetAttribute(bulkred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk * 8));
			for (int ws = 0; ws < 2; ++ws)
			{
				const size_t bytes = ws == 0 ? (size_t(48) << 20) : big;
				float ms = time_ms([&] { bulkred_kernel<<<sms * 2, 128, chunk * 8>>>(a, bytes / 8, chunk, iters); }, 3);
				const double total = double(iters) * sms * 2 * chunk * 8;
				printf("{\"bench\": \"tma_bulk_reduce_add_f64\", \"chunk_bytes\": %d, \"working_set_MB\": %zu, \"ms\": %.3f, \"GBs_payload\": %.1f}\n", chunk * 8, bytes >> 20, ms, total / ms * 1e-6);
			}
		}
	}
	CK(cudaFree(a));
	CK(cudaFree(b));
	return 0;
}
