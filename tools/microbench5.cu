// Fifth micro-benchmark: what FP64 issue rate do W resident warps per SM reach with ILP independent DFMA chains per lane,
// alone and mixed with the same number of FP32-pipe / ALU instructions (FSEL, IMAD)? Decides how many warps per SM the
// owner-computes kernel (pfa_collane2.cu, 238 registers -> 8 warps per SM) needs to saturate the FP64 pipe.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench5 tools/microbench5.cu && ./tools/microbench5
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int ILP, int MIX>
__global__ void __launch_bounds__(32) k(double *out, int iters, double a, double b, int sel)
{
	extern __shared__ double pad[];
	double x[ILP];
	int y[ILP];
#pragma unroll
	for (int i = 0; i < ILP; ++i)
	{
		x[i] = a * (threadIdx.x + i);
		y[i] = threadIdx.x + i;
	}
	for (int it = 0; it < iters; ++it)
	{
#pragma unroll
		for (int u = 0; u < 8; ++u)
		{
#pragma unroll
			for (int i = 0; i < ILP; ++i)
			{
				x[i] = fma(x[i], a, b);
				if (MIX)
					y[i] = y[i] * sel + u; // IMAD: one non-FP64 instruction per DFMA
			}
		}
	}
	double s = 0;
	int t = 0;
#pragma unroll
	for (int i = 0; i < ILP; ++i)
	{
		s += x[i];
		t += y[i];
	}
	if (s == 12345.678 || t == 123456789)
		out[blockIdx.x] = s + t + pad[0];
}

template <int ILP, int MIX>
void run(int warps_per_sm, int sms, double *out)
{
	// occupancy is pinned by dynamic shared memory: 227 KB / warps_per_sm per one-warp CTA
	const int smem = (227 * 1024 / warps_per_sm - 1024) & ~127;
	CK(cudaFuncSetAttribute(k<ILP, MIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
	int per_sm = 0;
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k<ILP, MIX>, 32, smem));
	const int iters = 20000, grid = sms * warps_per_sm;
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	float best = 1e30f;
	for (int r = 0; r < 3; ++r)
	{
		CK(cudaEventRecord(e0));
		k<ILP, MIX><<<grid, 32, smem>>>(out, iters, 1.0000001, 1e-9, 3);
		CK(cudaEventRecord(e1));
		CK(cudaEventSynchronize(e1));
		float ms;
		CK(cudaEventElapsedTime(&ms, e0, e1));
		best = ms < best ? ms : best;
	}
	const double dfma = double(iters) * 8 * ILP * 32 * grid;
	printf("{\"bench\": \"dfma_occupancy\", \"warps_per_sm\": %d, \"resident\": %d, \"ilp\": %d, \"mix_imad\": %d, \"ms\": %.3f, \"tflops\": %.2f, \"dfma_lanes_per_clk_per_sm_at_1965\": %.1f}\n",
		   warps_per_sm, per_sm, ILP, MIX, best, 2 * dfma / best * 1e-9, dfma / (best * 1e-3) / sms / 1.965e9);
}

int main()
{
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, 0));
	double *out;
	CK(cudaMalloc(&out, 8 * 148 * 64));
	for (int w : {4, 8, 12, 16, 32})
	{
		run<1, 0>(w, prop.multiProcessorCount, out);
		run<2, 0>(w, prop.multiProcessorCount, out);
		run<4, 0>(w, prop.multiProcessorCount, out);
		run<8, 0>(w, prop.multiProcessorCount, out);
		run<4, 1>(w, prop.multiProcessorCount, out);
		run<8, 1>(w, prop.multiProcessorCount, out);
	}
	return 0;
}
