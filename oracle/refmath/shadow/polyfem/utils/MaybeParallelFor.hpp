// shadow of polyfem/utils/MaybeParallelFor.hpp for oracle/refmath: maybe_parallel_for without a thread pool
// (MaybeParallelFor.hpp:20-21, .tpp:18-42): the index range is cut into ref_thread_count() consecutive chunks that run
// one after the other with thread ids 0, 1, ... - the per-thread storages and their serial merge are exercised, the
// run stays deterministic. Default 1 chunk = the reference's serial build.
#pragma once
#include <algorithm>
#include <functional>
#include <vector>
namespace polyfem::utils
{
	inline int &ref_thread_count()
	{
		static int n = 1;
		return n;
	}
	inline void maybe_parallel_for(int size, const std::function<void(int, int, int)> &partial_for)
	{
		const int chunks = std::max(1, std::min(ref_thread_count(), size));
		for (int t = 0; t < chunks; ++t)
			partial_for(int(long(size) * t / chunks), int(long(size) * (t + 1) / chunks), t);
	}
	inline void maybe_parallel_for(int size, const std::function<void(int)> &body)
	{
		for (int i = 0; i < size; ++i)
			body(i);
	}
	// create_thread_storage / get_local_thread_storage (MaybeParallelFor.hpp:28-32): one copy of the exemplar per thread
	template <typename T>
	inline std::vector<T> create_thread_storage(const T &exemplar)
	{
		return std::vector<T>(size_t(std::max(1, ref_thread_count())), exemplar);
	}
	template <typename S>
	inline auto &get_local_thread_storage(S &storage, int thread_id)
	{
		return storage[size_t(thread_id)];
	}
} // namespace polyfem::utils
