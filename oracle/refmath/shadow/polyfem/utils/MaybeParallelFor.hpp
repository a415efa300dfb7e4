// shadow of polyfem/utils/MaybeParallelFor.hpp for oracle/refmath: the serial form of maybe_parallel_for
// (MaybeParallelFor.tpp:18-29 without TBB): one range, thread id 0
#pragma once
namespace polyfem::utils
{
	template <typename F>
	inline void maybe_parallel_for(int size, const F &f) { f(0, size, 0); }
} // namespace polyfem::utils
