// TEST INFRASTRUCTURE: column-compressed sparse matrix with the handful of Eigen::SparseMatrix operations that the
// reference's utils/MatrixCache.cpp uses, so that file can be compiled UNMODIFIED here (oracle/refmath/shadow/ is
// put in front of the reference tree on the include path). Semantics restated from Eigen's documentation:
// setFromTriplets sums duplicates and keeps zero-valued entries as structural entries; += / + form the union of
// the two patterns (no pruning); inner indices ascending inside each column. Not Eigen, no shared code.
#pragma once
#include "mini_eigen.hpp"

#include <algorithm>
#include <vector>

namespace Eigen
{
	constexpr int ColMajor = 0;
	constexpr int RowMajor = 1;

	template <typename S>
	class Triplet
	{
	public:
		Triplet() = default;
		Triplet(int i, int j, S v = S(0)) : i_(i), j_(j), v_(v) {}
		int row() const { return i_; }
		int col() const { return j_; }
		S value() const { return v_; }

	private:
		int i_ = 0, j_ = 0;
		S v_ = S(0);
	};

	template <typename T>
	class Map;

	template <typename S, int Opt = ColMajor, typename I = int>
	class SparseMatrix;

	template <typename S, int Opt, typename I>
	class SparseMatrix
	{
	public:
		typedef I StorageIndex;
		SparseMatrix() : outer_(1, 0) {}
		SparseMatrix(long r, long c) { resize(r, c); }

		void resize(long r, long c)
		{
			rows_ = r;
			cols_ = c;
			filled_ = -1;
			outer_.assign(size_t(c) + 1, 0);
			inner_.clear();
			val_.clear();
		}
		void setZero()
		{
			outer_.assign(size_t(cols_) + 1, 0);
			inner_.clear();
			val_.clear();
		}
		long rows() const { return rows_; }
		long cols() const { return cols_; }
		long nonZeros() const { return long(val_.size()); }
		long outerSize() const { return cols_; }
		void makeCompressed() {}
		struct Storage
		{
			void squeeze() {}
		};
		Storage data() { return Storage(); }
		const I *outerIndexPtr() const { return outer_.data(); }
		const I *innerIndexPtr() const { return inner_.data(); }
		const S *valuePtr() const { return val_.data(); }

		// column-by-column construction (startVec / insertBack / finalize) and read-out (InnerIterator)
		void reserve(long n)
		{
			inner_.reserve(size_t(n));
			val_.reserve(size_t(n));
		}
		void startVec(long c)
		{
			for (long k = filled_ + 1; k <= c; ++k)
				outer_[size_t(k)] = I(inner_.size());
			filled_ = c;
		}
		S &insertBack(long r, long c)
		{
			assert(c == filled_);
			inner_.push_back(I(r));
			val_.push_back(S(0));
			return val_.back();
		}
		void finalize()
		{
			for (long k = filled_ + 1; k <= cols_; ++k)
				outer_[size_t(k)] = I(inner_.size());
			filled_ = cols_;
		}
		class InnerIterator
		{
		public:
			InnerIterator(const SparseMatrix &m, long c) : m_(m), k_(m.outer_[size_t(c)]), end_(m.outer_[size_t(c) + 1]), c_(c) {}
			explicit operator bool() const { return k_ < end_; }
			InnerIterator &operator++()
			{
				++k_;
				return *this;
			}
			I row() const { return m_.inner_[size_t(k_)]; }
			I col() const { return I(c_); }
			S value() const { return m_.val_[size_t(k_)]; }

		private:
			const SparseMatrix &m_;
			I k_, end_;
			long c_;
		};

		template <typename It>
		void setFromTriplets(It begin, It end)
		{
			std::vector<Triplet<S>> t(begin, end);
			std::stable_sort(t.begin(), t.end(), [](const Triplet<S> &a, const Triplet<S> &b) {
				return a.col() != b.col() ? a.col() < b.col() : a.row() < b.row();
			});
			outer_.assign(size_t(cols_) + 1, 0);
			inner_.clear();
			val_.clear();
			for (size_t k = 0; k < t.size(); ++k)
			{
				if (k > 0 && t[k].col() == t[k - 1].col() && t[k].row() == t[k - 1].row())
					val_.back() += t[k].value();
				else
				{
					inner_.push_back(I(t[k].row()));
					val_.push_back(t[k].value());
					outer_[size_t(t[k].col()) + 1]++;
				}
			}
			for (long c = 0; c < cols_; ++c)
				outer_[size_t(c) + 1] += outer_[size_t(c)];
		}

		SparseMatrix &operator+=(const SparseMatrix &o)
		{
			*this = *this + o;
			return *this;
		}
		friend SparseMatrix operator+(const SparseMatrix &a, const SparseMatrix &b)
		{
			assert(a.rows_ == b.rows_ && a.cols_ == b.cols_);
			SparseMatrix r(a.rows_, a.cols_);
			for (long c = 0; c < a.cols_; ++c)
			{
				I ka = a.outer_[size_t(c)], kb = b.outer_[size_t(c)];
				const I ea = a.outer_[size_t(c) + 1], eb = b.outer_[size_t(c) + 1];
				while (ka < ea || kb < eb)
				{
					if (kb >= eb || (ka < ea && a.inner_[size_t(ka)] < b.inner_[size_t(kb)]))
					{
						r.inner_.push_back(a.inner_[size_t(ka)]);
						r.val_.push_back(a.val_[size_t(ka)]);
						++ka;
					}
					else if (ka >= ea || b.inner_[size_t(kb)] < a.inner_[size_t(ka)])
					{
						r.inner_.push_back(b.inner_[size_t(kb)]);
						r.val_.push_back(b.val_[size_t(kb)]);
						++kb;
					}
					else
					{
						r.inner_.push_back(a.inner_[size_t(ka)]);
						r.val_.push_back(a.val_[size_t(ka)] + b.val_[size_t(kb)]);
						++ka;
						++kb;
					}
				}
				r.outer_[size_t(c) + 1] = I(r.inner_.size());
			}
			return r;
		}

		// from Map<const SparseMatrix>(rows, cols, nnz, outer, inner, values)
		SparseMatrix(const Map<const SparseMatrix> &m);
		SparseMatrix &operator=(const Map<const SparseMatrix> &m);

		std::vector<I> outer_, inner_;
		std::vector<S> val_;

	private:
		long rows_ = 0, cols_ = 0, filled_ = -1;
	};

	template <typename S, int Opt, typename I>
	class Map<const SparseMatrix<S, Opt, I>>
	{
	public:
		Map(long rows, long cols, long nnz, const I *outer, const I *inner, const S *values)
			: rows(rows), cols(cols), nnz(nnz), outer(outer), inner(inner), values(values) {}
		long rows, cols, nnz;
		const I *outer, *inner;
		const S *values;
	};

	template <typename S, int Opt, typename I>
	SparseMatrix<S, Opt, I> &SparseMatrix<S, Opt, I>::operator=(const Map<const SparseMatrix> &m)
	{
		rows_ = m.rows;
		cols_ = m.cols;
		outer_.assign(m.outer, m.outer + m.cols + 1);
		inner_.assign(m.inner, m.inner + m.nnz);
		val_.assign(m.values, m.values + m.nnz);
		return *this;
	}
	template <typename S, int Opt, typename I>
	SparseMatrix<S, Opt, I>::SparseMatrix(const Map<const SparseMatrix> &m) { *this = m; }

	template <int R, int C, int Opt, int MR, int MC>
	template <typename S2, int O2, typename I2>
	Matrix<double, R, C, Opt, MR, MC>::Matrix(const SparseMatrix<S2, O2, I2> &sp) : Dense(sp.rows(), sp.cols())
	{
		for (long c = 0; c < sp.outerSize(); ++c)
			for (typename SparseMatrix<S2, O2, I2>::InnerIterator it(sp, c); it; ++it)
				(*this)(it.row(), it.col()) += it.value();
	}

	// Dense::sparseView() (DenseMatrixCache::get_matrix): non-zero entries only
	inline SparseMatrix<double, 0, int> Dense::sparseView() const
	{
		std::vector<Triplet<double>> t;
		for (long j = 0; j < cols(); ++j)
			for (long i = 0; i < rows(); ++i)
				if ((*this)(i, j) != 0.0)
					t.emplace_back(int(i), int(j), (*this)(i, j));
		SparseMatrix<double, 0, int> s(rows(), cols());
		s.setFromTriplets(t.begin(), t.end());
		return s;
	}
} // namespace Eigen
