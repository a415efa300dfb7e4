// TEST INFRASTRUCTURE: C entry points over the reference's OWN utils/MatrixCache.cpp, compiled UNMODIFIED from
// /root/reference (this file includes the reference's header; shadow/ supplies <Eigen/...>, Types.hpp, Logger.hpp and
// MaybeParallelFor.hpp) -> oracle/_ref/libcacheref.so. Same call surface as oracle_cache_* in oracle/oracle.h, so a
// test can drive the reference class and the oracle's restatement with one script.
#include <polyfem/utils/MatrixCache.hpp>

#include <memory>

using polyfem::utils::MatrixCache;
using polyfem::utils::SparseMatrixCache;

struct refcache
{
	std::unique_ptr<SparseMatrixCache> c;
	polyfem::StiffnessMatrix last;
};

extern "C"
{
	refcache *refcache_new(int size)
	{
		auto *r = new refcache();
		r->c = std::make_unique<SparseMatrixCache>(size_t(size));
		return r;
	}
	// A thread-local cache the way NLAssembler::assemble_hessian makes one (Assembler.cpp:31-41, 60-66, 669):
	// exemplar = c.copy(); exemplar->init(c);  then every thread storage is exemplar->copy()
	refcache *refcache_copy(const refcache *o)
	{
		std::unique_ptr<MatrixCache> exemplar = o->c->copy();
		exemplar->init(*o->c);
		auto *r = new refcache();
		std::unique_ptr<MatrixCache> p = exemplar->copy();
		r->c.reset(dynamic_cast<SparseMatrixCache *>(p.release()));
		return r;
	}
	// SparseMatrixCache(const MatrixCache &other) -> init(other) (MatrixCache.cpp:18-21, 49-54)
	refcache *refcache_copy_ctor(const refcache *o)
	{
		auto *r = new refcache();
		r->c = std::make_unique<SparseMatrixCache>(static_cast<const MatrixCache &>(*o->c));
		return r;
	}
	void refcache_free(refcache *r) { delete r; }
	void refcache_add_value(refcache *r, int e, int i, int j, double v) { r->c->add_value(e, i, j, v); }
	void refcache_prune(refcache *r) { r->c->prune(); }
	void refcache_set_zero(refcache *r) { r->c->set_zero(); }
	void refcache_add(refcache *dst, const refcache *src) { *dst->c += *src->c; }
	long refcache_get_matrix(refcache *r)
	{
		r->last = r->c->get_matrix();
		return r->last.nonZeros();
	}
	const int *refcache_outer(const refcache *r) { return r->last.outerIndexPtr(); }
	const int *refcache_inner(const refcache *r) { return r->last.innerIndexPtr(); }
	const double *refcache_values(const refcache *r) { return r->last.valuePtr(); }
}
