// Device side of the rows right after the assembly path:
//
//  * Dirichlet projection of gradient and Hessian, the job of
//    BCLagrangianForm::project_gradient / project_hessian
//    (solver/forms/lagrangian/BCLagrangianForm.cpp:149-155, 167-213) called from
//    NLProblem::gradient / full_hessian_to_reduced_hessian (solver/NLProblem.cpp:596-633, 735-751):
//    constrained dofs are dropped from the vector and from the rows and columns of the CSC
//    matrix, kept entries keep their relative order. Here the reduced pattern and a gather map
//    (reduced nnz -> full nnz) are built once per constraint set on the device; every Newton
//    iteration is then one streaming gather, with the Form weight (Form.hpp:42-56) folded in.
//  * NaN scan of a device vector for ElasticForm::is_step_valid (ElasticForm.cpp:388-396).
#include "pfa_internal.h"

#include <cub/cub.cuh>

namespace pfa
{
	namespace
	{
		__global__ void mark_constrained_kernel(const int32_t *__restrict__ dofs, int64_t n, int32_t ndof, int32_t *__restrict__ keep, int *__restrict__ bad)
		{
			const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
			if (t >= n)
				return;
			const int32_t d = dofs[t];
			if (d < 0 || d >= ndof)
				*bad = 1;
			else
				keep[d] = 0;
		}

		__global__ void fill_int_kernel(int32_t *__restrict__ p, int64_t n, int32_t v)
		{
			const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
			if (t < n)
				p[t] = v;
		}

		// old_to_new[d] = rank among kept dofs or -1 (BCLagrangianForm's old_to_new_); not_constraints[rank] = d
		__global__ void finish_maps_kernel(const int32_t *__restrict__ keep, const int32_t *__restrict__ rank, int32_t ndof, int32_t *__restrict__ old_to_new, int32_t *__restrict__ not_constraints)
		{
			const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
			if (t >= ndof)
				return;
			if (keep[t])
			{
				old_to_new[t] = rank[t];
				not_constraints[rank[t]] = int32_t(t);
			}
			else
				old_to_new[t] = -1;
		}

		// pass 1 of project_hessian (BCLagrangianForm.cpp:180-190): kept entries per kept column; one warp per column
		__global__ void count_kept_kernel(const int32_t *__restrict__ outer, const int32_t *__restrict__ inner, const int32_t *__restrict__ old_to_new, int32_t ndof, int32_t *__restrict__ col_count)
		{
			const int64_t c = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
			const int lane = threadIdx.x & 31;
			if (c >= ndof)
				return;
			const int32_t nc = old_to_new[c];
			if (nc < 0)
				return;
			int cnt = 0;
			for (int64_t k = outer[c] + lane; k < outer[c + 1]; k += 32)
				cnt += old_to_new[inner[k]] >= 0;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1)
				cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
			if (lane == 0)
				col_count[nc] = cnt;
		}

		// pass 2 (BCLagrangianForm.cpp:196-209): ordered compaction of each kept column
		__global__ void fill_reduced_kernel(const int32_t *__restrict__ outer, const int32_t *__restrict__ inner, const int32_t *__restrict__ old_to_new, int32_t ndof,
											const int32_t *__restrict__ outer_red, int32_t *__restrict__ inner_red, int32_t *__restrict__ map)
		{
			const int64_t c = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
			const int lane = threadIdx.x & 31;
			if (c >= ndof)
				return;
			const int32_t nc = old_to_new[c];
			if (nc < 0)
				return;
			int64_t pos = outer_red[nc];
			const int64_t k0 = outer[c], k1 = outer[c + 1];
			for (int64_t base = k0; base < k1; base += 32)
			{
				const int64_t k = base + lane;
				int32_t nr = -1;
				if (k < k1)
					nr = old_to_new[inner[k]];
				const unsigned ballot = __ballot_sync(0xffffffffu, nr >= 0);
				if (nr >= 0)
				{
					const int64_t p = pos + __popc(ballot & ((1u << lane) - 1u));
					inner_red[p] = nr;
					map[p] = int32_t(k);
				}
				pos += __popc(ballot);
			}
		}

		__global__ void gather_scale_kernel(const double *__restrict__ src, const int32_t *__restrict__ map, int64_t n, double scale, double *__restrict__ dst)
		{
			const int64_t stride = int64_t(gridDim.x) * blockDim.x;
			for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += stride)
				dst[t] = scale * src[map[t]];
		}

		__global__ void any_nan_kernel(const double *__restrict__ v, int64_t n, int *__restrict__ flag)
		{
			const int64_t stride = int64_t(gridDim.x) * blockDim.x;
			bool bad = false;
			for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += stride)
				bad |= isnan(v[t]);
			if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0)
				*flag = 1;
		}

		// y[c] = sum_k values[k] * d[inner[k]] over column c of the symmetric matrix (a CSC column read as
		// the CSR row): one warp per column, no atomics on y; energy by one atomic per CTA
		__global__ void symv_kernel(const int32_t *__restrict__ outer, const int32_t *__restrict__ inner, const double *__restrict__ values,
									const double *__restrict__ x, const double *__restrict__ x_tilde, int32_t ndof, double *__restrict__ y, double *__restrict__ energy)
		{
			const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
			const int64_t c = int64_t(blockIdx.x) * (blockDim.x >> 5) + warp;
			double part = 0.0;
			if (c < ndof)
			{
				double s = 0.0;
				for (int64_t k = outer[c] + lane; k < outer[c + 1]; k += 32)
				{
					const int32_t r = inner[k];
					s += values[k] * (x_tilde ? x[r] - x_tilde[r] : x[r]);
				}
#pragma unroll
				for (int o = 16; o > 0; o >>= 1)
					s += __shfl_xor_sync(0xffffffffu, s, o);
				if (lane == 0)
				{
					if (y)
						y[c] = s;
					part = 0.5 * s * (x_tilde ? x[c] - x_tilde[c] : x[c]);
				}
			}
			if (energy)
			{
				__shared__ double sh[32];
				if (lane == 0)
					sh[warp] = part;
				__syncthreads();
				if (threadIdx.x == 0)
				{
					double t = 0.0;
					for (int w = 0; w < int(blockDim.x >> 5); ++w)
						t += sh[w];
					atomicAdd(energy, t);
				}
			}
		}

		__global__ void axpy_kernel(int64_t n, double a, const double *__restrict__ x, double *__restrict__ y)
		{
			const int64_t stride = int64_t(gridDim.x) * blockDim.x;
			for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += stride)
				y[t] += a * x[t];
		}

		// ---- tables of the row-lane kernels for the reduced matrix ----
		// node_mask[b]: which of the 3 dofs of node b are kept
		__global__ void node_mask_kernel(const int32_t *__restrict__ keep, int32_t n_bases, int32_t *__restrict__ node_mask)
		{
			const int64_t b = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
			if (b < n_bases)
				node_mask[b] = (keep[3 * b] ? 1 : 0) | (keep[3 * b + 1] ? 2 : 0) | (keep[3 * b + 2] ? 4 : 0);
		}

		// one warp per column node b: rowprefix[p] = kept rows before row node adj[p] inside a column of b,
		// cs_red[b] = kept rows of such a column, cbase_red[b] = reduced values index of its first kept column
		__global__ void column_prefix_kernel(const int32_t *__restrict__ adj_off, const int32_t *__restrict__ adj, const int32_t *__restrict__ node_mask,
											 const int32_t *__restrict__ old_to_new, const int32_t *__restrict__ outer_red, int32_t n_bases,
											 int32_t *__restrict__ rowprefix, int32_t *__restrict__ cs_red, int32_t *__restrict__ cbase_red)
		{
			const int64_t b = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
			const int lane = threadIdx.x & 31;
			if (b >= n_bases)
				return;
			int run = 0;
			for (int p0 = adj_off[b]; p0 < adj_off[b + 1]; p0 += 32)
			{
				const int p = p0 + lane;
				const int c = p < adj_off[b + 1] ? __popc(node_mask[adj[p]]) : 0;
				int incl = c;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1)
				{
					const int v = __shfl_up_sync(0xffffffffu, incl, o);
					if (lane >= o)
						incl += v;
				}
				if (p < adj_off[b + 1])
					rowprefix[p] = run + incl - c;
				run += __shfl_sync(0xffffffffu, incl, 31);
			}
			if (lane == 0)
			{
				cs_red[b] = run;
				const int mb = node_mask[b];
				cbase_red[b] = mb ? outer_red[old_to_new[3 * b + (__ffs(mb) - 1)]] : 0;
			}
		}

		__global__ void reduced_entry_kernel(const int32_t *__restrict__ conn, const int32_t *__restrict__ entry, const int32_t *__restrict__ adj_off,
											 const int32_t *__restrict__ node_mask, const int32_t *__restrict__ rowprefix, const int32_t *__restrict__ cs_red,
											 const int32_t *__restrict__ cbase_red, int64_t n_el, int n_loc, int32_t *__restrict__ entry_red, int32_t *__restrict__ cstride_red)
		{
			const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
			const int64_t per_el = int64_t(n_loc) * n_loc;
			if (t >= n_el * per_el)
				return;
			const int64_t e = t / per_el;
			const int ij = int(t - e * per_el);
			const int i = ij / n_loc, j = ij - i * n_loc;
			const int32_t gj = conn[e * n_loc + j];
			const int32_t off = adj_off[gj];
			const int32_t p = off + (entry[t] - 9 * off) / 3; // position of the pair in adj
			entry_red[t] = cbase_red[gj] + rowprefix[p];
			if (i == 0)
				cstride_red[e * n_loc + j] = cs_red[gj] | (node_mask[gj] << 28);
		}

		inline unsigned blocks_for(int64_t n, int threads) { return unsigned(std::max<int64_t>(1, (n + threads - 1) / threads)); }
	} // namespace

	cudaError_t exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, void **scratch, size_t *scratch_bytes, cudaStream_t st)
	{
		size_t need = 0;
		cudaError_t err = cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, int(n), st);
		if (err != cudaSuccess)
			return err;
		if (need > *scratch_bytes)
		{
			if (*scratch)
				cudaFree(*scratch);
			*scratch = nullptr;
			*scratch_bytes = 0;
			err = cudaMalloc(scratch, need);
			if (err != cudaSuccess)
				return err;
			*scratch_bytes = need;
		}
		return cub::DeviceScan::ExclusiveSum(*scratch, need, in, out, int(n), st);
	}

	cudaError_t launch_mark_constrained(const int32_t *dofs_dev, int64_t n, int32_t ndof, int32_t *keep, int *bad, cudaStream_t st)
	{
		fill_int_kernel<<<blocks_for(ndof, 256), 256, 0, st>>>(keep, ndof, 1);
		if (n > 0)
			mark_constrained_kernel<<<blocks_for(n, 256), 256, 0, st>>>(dofs_dev, n, ndof, keep, bad);
		return cudaGetLastError();
	}

	cudaError_t launch_finish_maps(const int32_t *keep, const int32_t *rank, int32_t ndof, int32_t *old_to_new, int32_t *not_constraints, cudaStream_t st)
	{
		finish_maps_kernel<<<blocks_for(ndof, 256), 256, 0, st>>>(keep, rank, ndof, old_to_new, not_constraints);
		return cudaGetLastError();
	}

	cudaError_t launch_count_kept(const int32_t *outer, const int32_t *inner, const int32_t *old_to_new, int32_t ndof, int32_t *col_count, cudaStream_t st)
	{
		count_kept_kernel<<<blocks_for(int64_t(ndof) * 32, 256), 256, 0, st>>>(outer, inner, old_to_new, ndof, col_count);
		return cudaGetLastError();
	}

	cudaError_t launch_fill_reduced(const int32_t *outer, const int32_t *inner, const int32_t *old_to_new, int32_t ndof, const int32_t *outer_red, int32_t *inner_red, int32_t *map, cudaStream_t st)
	{
		fill_reduced_kernel<<<blocks_for(int64_t(ndof) * 32, 256), 256, 0, st>>>(outer, inner, old_to_new, ndof, outer_red, inner_red, map);
		return cudaGetLastError();
	}

	cudaError_t launch_reduced_tables(const DeviceMesh &m, const int32_t *keep, const int32_t *old_to_new, const int32_t *outer_red,
									  int32_t *node_mask, int32_t *rowprefix, int32_t *cs_red, int32_t *cbase_red, int32_t *entry_red, int32_t *cstride_red, cudaStream_t st)
	{
		node_mask_kernel<<<blocks_for(m.n_bases, 256), 256, 0, st>>>(keep, m.n_bases, node_mask);
		column_prefix_kernel<<<blocks_for(int64_t(m.n_bases) * 32, 256), 256, 0, st>>>(m.adj_off, m.adj, node_mask, old_to_new, outer_red, m.n_bases, rowprefix, cs_red, cbase_red);
		const int64_t total = int64_t(m.n_el) * m.n_loc * m.n_loc;
		reduced_entry_kernel<<<blocks_for(total, 256), 256, 0, st>>>(m.conn, m.entry, m.adj_off, node_mask, rowprefix, cs_red, cbase_red, m.n_el, m.n_loc, entry_red, cstride_red);
		return cudaGetLastError();
	}

	cudaError_t launch_gather_scale(const double *src, const int32_t *map, int64_t n, double scale, double *dst, int sm_count, cudaStream_t st)
	{
		if (n <= 0)
			return cudaSuccess;
		const unsigned grid = unsigned(std::min<int64_t>(blocks_for(n, 256), int64_t(sm_count) * 16));
		gather_scale_kernel<<<grid, 256, 0, st>>>(src, map, n, scale, dst);
		return cudaGetLastError();
	}

	cudaError_t launch_symv(const int32_t *outer, const int32_t *inner, const double *values, const double *x, const double *x_tilde, int32_t ndof,
							double *y, double *energy, cudaStream_t st)
	{
		const int warps = 8;
		symv_kernel<<<blocks_for(ndof, warps), warps * 32, 0, st>>>(outer, inner, values, x, x_tilde, ndof, y, energy);
		return cudaGetLastError();
	}

	cudaError_t launch_axpy(int64_t n, double a, const double *x, double *y, int sm_count, cudaStream_t st)
	{
		if (n <= 0)
			return cudaSuccess;
		const unsigned grid = unsigned(std::min<int64_t>(blocks_for(n, 256), int64_t(sm_count) * 16));
		axpy_kernel<<<grid, 256, 0, st>>>(n, a, x, y);
		return cudaGetLastError();
	}

	cudaError_t launch_any_nan(const double *v, int64_t n, int *flag, int sm_count, cudaStream_t st)
	{
		const unsigned grid = unsigned(std::min<int64_t>(blocks_for(n, 256), int64_t(sm_count) * 8));
		any_nan_kernel<<<grid, 256, 0, st>>>(v, n, flag);
		return cudaGetLastError();
	}
} // namespace pfa
