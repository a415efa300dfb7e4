// TEST INFRASTRUCTURE: a small dense-matrix stand-in for the part of Eigen's API that the reference's
// NeoHookean "fast" gradient / Hessian functions (assembler/NeoHookeanElasticity.cpp:419-658) and the linear local
// blocks (LinearElasticity.cpp:29-63, Laplacian.cpp:13-26, Mass.cpp:5-23) use, so that
// THOSE FUNCTION BODIES can be compiled from /root/reference and executed here, where Eigen is not installed
// (oracle/refmath/extract_nh.py + nh_glue.cpp -> oracle/_ref/libnhref.so). Every operation is eager and
// returns a plain column-major matrix (`Dense`); fixed-size template arguments only provide default shapes.
// This is not Eigen and shares no code with it.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <ostream>
#include <type_traits>
#include <vector>

namespace Eigen
{
	constexpr int Dynamic = -1;
	template <typename S, int Opt, typename I>
	class SparseMatrix; // mini_sparse.hpp

	struct IOFormat // only named by the stream operators of utils/autodiff.h, which nothing here calls
	{
		template <typename... A>
		IOFormat(A...) {}
	};

	class Dense
	{
	public:
		typedef double Scalar;
		Dense() = default;
		Dense(long r, long c) : r_(int(r)), c_(int(c)), d_(size_t(r) * size_t(c), 0.0) {}

		long rows() const { return r_; }
		long cols() const { return c_; }
		long size() const { return long(r_) * c_; }
		double *data() { return d_.data(); }
		const double *data() const { return d_.data(); }
		void resize(long r, long c)
		{
			r_ = int(r);
			c_ = int(c);
			d_.assign(size_t(r) * size_t(c), 0.0);
		}
		void resize(long n) // vectors
		{
			if (c_ == 1 || (r_ == 0 && c_ == 0))
				resize(n, 1);
			else
				resize(1, n);
		}
		const char *format(const IOFormat &) const { return ""; }
		void setZero() { d_.assign(d_.size(), 0.0); }
		void setZero(long r, long c) { resize(r, c); }

		double &operator()(long i, long j) { return d_[size_t(j) * r_ + i]; }
		double operator()(long i, long j) const { return d_[size_t(j) * r_ + i]; }
		double &operator[](long i) { return d_[size_t(i)]; }
		double operator[](long i) const { return d_[size_t(i)]; }
		void conservativeResize(long n) // vectors
		{
			d_.resize(size_t(n));
			if (c_ == 1 || r_ == 0)
			{
				r_ = int(n);
				c_ = 1;
			}
			else
				c_ = int(n);
		}
		double &operator()(long i) { return d_[size_t(i)]; } // vectors (and column-major linear access)
		double operator()(long i) const { return d_[size_t(i)]; }

		Dense transpose() const
		{
			Dense t(c_, r_);
			for (int i = 0; i < r_; ++i)
				for (int j = 0; j < c_; ++j)
					t(j, i) = (*this)(i, j);
			return t;
		}
		double trace() const
		{
			double s = 0;
			for (int i = 0; i < r_ && i < c_; ++i)
				s += (*this)(i, i);
			return s;
		}
		double determinant() const
		{
			const Dense &m = *this;
			assert(r_ == c_ && r_ >= 1 && r_ <= 3);
			if (r_ == 1)
				return m(0, 0);
			if (r_ == 2)
				return m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0);
			return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0))
				   + m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
		}
		Dense inverse() const // closed forms of utils/MatrixUtils.hpp:37-80 (Eigen's own 3x3 inverse differs only in rounding)
		{
			const Dense &m = *this;
			assert(r_ == c_ && r_ >= 1 && r_ <= 3);
			Dense inv(r_, c_);
			const double det = determinant();
			if (r_ == 1)
				inv(0, 0) = 1.0 / m(0, 0);
			else if (r_ == 2)
			{
				inv(0, 0) = m(1, 1) / det;
				inv(0, 1) = -m(0, 1) / det;
				inv(1, 0) = -m(1, 0) / det;
				inv(1, 1) = m(0, 0) / det;
			}
			else
			{
				inv(0, 0) = (-m(1, 2) * m(2, 1) + m(1, 1) * m(2, 2)) / det;
				inv(0, 1) = (m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2)) / det;
				inv(0, 2) = (-m(0, 2) * m(1, 1) + m(0, 1) * m(1, 2)) / det;
				inv(1, 0) = (m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2)) / det;
				inv(1, 1) = (-m(0, 2) * m(2, 0) + m(0, 0) * m(2, 2)) / det;
				inv(1, 2) = (m(0, 2) * m(1, 0) - m(0, 0) * m(1, 2)) / det;
				inv(2, 0) = (-m(1, 1) * m(2, 0) + m(1, 0) * m(2, 1)) / det;
				inv(2, 1) = (m(0, 1) * m(2, 0) - m(0, 0) * m(2, 1)) / det;
				inv(2, 2) = (-m(0, 1) * m(1, 0) + m(0, 0) * m(1, 1)) / det;
			}
			return inv;
		}
		double sum() const
		{
			double s = 0;
			for (double v : d_)
				s += v;
			return s;
		}
		double squaredNorm() const
		{
			double s = 0;
			for (double v : d_)
				s += v * v;
			return s;
		}
		double dot(const Dense &o) const
		{
			assert(size() == o.size());
			double s = 0;
			for (long k = 0; k < size(); ++k)
				s += d_[size_t(k)] * o(k);
			return s;
		}
		double norm() const { return std::sqrt(squaredNorm()); }
		double prod() const
		{
			double s = 1;
			for (double v : d_)
				s *= v;
			return s;
		}
		Dense cross(const Dense &o) const // 3-vectors
		{
			assert(size() == 3 && o.size() == 3);
			Dense r(r_, c_);
			r(0) = (*this)(1) * o(2) - (*this)(2) * o(1);
			r(1) = (*this)(2) * o(0) - (*this)(0) * o(2);
			r(2) = (*this)(0) * o(1) - (*this)(1) * o(0);
			return r;
		}
		// a unit vector orthogonal to this 3-vector, by the rule Eigen documents for unitOrthogonal(): rotate in the xy plane
		// unless x and y are both negligible against z (relative 1e-12), then in the yz plane
		Dense unitOrthogonal() const
		{
			assert(size() == 3);
			const double x = (*this)(0), y = (*this)(1), z = (*this)(2), prec = 1e-12;
			Dense r(r_, c_);
			if (!(std::fabs(x) <= std::fabs(z) * prec) || !(std::fabs(y) <= std::fabs(z) * prec))
			{
				const double invnm = 1.0 / std::sqrt(x * x + y * y);
				r(0) = -y * invnm;
				r(1) = x * invnm;
				r(2) = 0;
			}
			else
			{
				const double invnm = 1.0 / std::sqrt(y * y + z * z);
				r(0) = 0;
				r(1) = -z * invnm;
				r(2) = y * invnm;
			}
			return r;
		}
		template <typename I>
		double maxCoeff(I *where) const
		{
			long best = 0;
			for (long k = 1; k < size(); ++k)
				if (d_[size_t(k)] > d_[size_t(best)])
					best = k;
			*where = I(best);
			return d_[size_t(best)];
		}
		struct CommaInit // m << a, b, c, ...;  fills row by row, whatever the storage order (as Eigen's comma initialiser does)
		{
			Dense &m;
			long k;
			CommaInit &operator,(double v)
			{
				m(k / m.cols(), k % m.cols()) = v;
				++k;
				return *this;
			}
		};
		CommaInit operator<<(double v)
		{
			(*this)(0) = v;
			return CommaInit{*this, 1};
		}
		struct DiagArray // m.diagonal().array() += s
		{
			Dense &m;
			DiagArray &array() { return *this; }
			DiagArray &operator+=(double s)
			{
				for (long i = 0; i < m.rows() && i < m.cols(); ++i)
					m(i, i) += s;
				return *this;
			}
		};
		DiagArray diagonal() { return DiagArray{*this}; }
		struct Colwise // m.colwise().squaredNorm(): one entry per column
		{
			const Dense &m;
			Dense squaredNorm() const
			{
				Dense r(1, m.cols());
				for (long j = 0; j < m.cols(); ++j)
					for (long i = 0; i < m.rows(); ++i)
						r(0, j) += m(i, j) * m(i, j);
				return r;
			}
		};
		Colwise colwise() const { return Colwise{*this}; }
		Dense &noalias() { return *this; }
		Dense eval() const { return *this; }
		void transposeInPlace() { *this = transpose(); }
		struct Mask // (m.array() >= s): element-wise comparison result
		{
			std::vector<char> keep;
			int r, c;
			Dense select(const Dense &then, double otherwise) const
			{
				Dense out(r, c);
				for (size_t k = 0; k < keep.size(); ++k)
					out(long(k)) = keep[k] ? then(long(k)) : otherwise;
				return out;
			}
		};
		struct ArrayView // .array(): element-wise view (products of two views, Assembler.cpp:518; >=, sqrt: utils/svd.hpp)
		{
			const Dense &m;
			Mask operator>=(double s) const
			{
				Mask k{std::vector<char>(size_t(m.size())), int(m.rows()), int(m.cols())};
				for (long i = 0; i < m.size(); ++i)
					k.keep[size_t(i)] = m(i) >= s;
				return k;
			}
			Dense sqrt() const
			{
				Dense out(m.rows(), m.cols());
				for (long i = 0; i < m.size(); ++i)
					out(i) = std::sqrt(m(i));
				return out;
			}
		};
		ArrayView array() const { return ArrayView{*this}; }
		SparseMatrix<double, 0, int> sparseView() const; // defined in mini_sparse.hpp
		Dense &operator*=(const Dense &o); // matrix product, defined after operator*
		Dense &operator*=(double s)
		{
			for (double &v : d_)
				v *= s;
			return *this;
		}
		Dense &operator/=(double s)
		{
			for (double &v : d_)
				v /= s;
			return *this;
		}
		Dense &operator-=(const Dense &o)
		{
			assert(r_ == o.r_ && c_ == o.c_);
			for (size_t k = 0; k < d_.size(); ++k)
				d_[k] -= o.d_[k];
			return *this;
		}
		Dense &operator+=(const Dense &o)
		{
			assert(r_ == o.r_ && c_ == o.c_);
			for (size_t k = 0; k < d_.size(); ++k)
				d_[k] += o.d_[k];
			return *this;
		}

		// ---- writable views ----
		struct RowProxy
		{
			Dense &m;
			long i;
			RowProxy &operator=(const Dense &v)
			{
				assert(v.size() == m.cols());
				for (long j = 0; j < m.cols(); ++j)
					m(i, j) = v(j);
				return *this;
			}
			RowProxy &operator=(const RowProxy &o) { return *this = Dense(o); }
			RowProxy &operator+=(const Dense &v)
			{
				assert(v.size() == m.cols());
				for (long j = 0; j < m.cols(); ++j)
					m(i, j) += v(j);
				return *this;
			}
			operator Dense() const
			{
				Dense r(1, m.cols());
				for (long j = 0; j < m.cols(); ++j)
					r(0, j) = m(i, j);
				return r;
			}
			Dense transpose() const { return Dense(*this).transpose(); }
		};
		struct ColProxy
		{
			Dense &m;
			long j;
			ColProxy &operator=(const Dense &v)
			{
				assert(v.size() == m.rows());
				for (long i = 0; i < m.rows(); ++i)
					m(i, j) = v(i);
				return *this;
			}
			operator Dense() const
			{
				Dense r(m.rows(), 1);
				for (long i = 0; i < m.rows(); ++i)
					r(i, 0) = m(i, j);
				return r;
			}
			ColProxy &operator=(const ColProxy &o) { return *this = Dense(o); }
			double norm() const { return Dense(*this).norm(); }
			Dense cross(const Dense &o) const { return Dense(*this).cross(o); }
			Dense unitOrthogonal() const { return Dense(*this).unitOrthogonal(); }
			struct CommaCol
			{
				ColProxy &p;
				long k;
				CommaCol &operator,(double v)
				{
					p.m(k++, p.j) = v;
					return *this;
				}
			};
			CommaCol operator<<(double v)
			{
				m(0, j) = v;
				return CommaCol{*this, 1};
			}
		};
		struct BlockProxy
		{
			Dense &m;
			long i0, j0, br, bc;
			BlockProxy &operator=(const Dense &v)
			{
				assert(v.rows() == br && v.cols() == bc);
				for (long i = 0; i < br; ++i)
					for (long j = 0; j < bc; ++j)
						m(i0 + i, j0 + j) = v(i, j);
				return *this;
			}
			BlockProxy &operator+=(const Dense &v)
			{
				assert(v.rows() == br && v.cols() == bc);
				for (long i = 0; i < br; ++i)
					for (long j = 0; j < bc; ++j)
						m(i0 + i, j0 + j) += v(i, j);
				return *this;
			}
			operator Dense() const
			{
				Dense r(br, bc);
				for (long i = 0; i < br; ++i)
					for (long j = 0; j < bc; ++j)
						r(i, j) = m(i0 + i, j0 + j);
				return r;
			}
			Dense transpose() const { return Dense(*this).transpose(); }
		};
		RowProxy row(long i) { return RowProxy{*this, i}; }
		Dense row(long i) const { return Dense(RowProxy{const_cast<Dense &>(*this), i}); }
		ColProxy col(long j) { return ColProxy{*this, j}; }
		Dense col(long j) const { return Dense(ColProxy{const_cast<Dense &>(*this), j}); }
		// run-time sized block(i, j, rows, cols)
		BlockProxy block(long i, long j, long br, long bc) { return BlockProxy{*this, i, j, br, bc}; }
		Dense block(long i, long j, long br, long bc) const { return Dense(BlockProxy{const_cast<Dense &>(*this), i, j, br, bc}); }
		template <int BR, int BC>
		BlockProxy block(long i, long j) { return BlockProxy{*this, i, j, BR, BC}; }
		template <int BR, int BC>
		Dense block(long i, long j) const { return Dense(BlockProxy{const_cast<Dense &>(*this), i, j, BR, BC}); }

	protected:
		int r_ = 0, c_ = 0;
		std::vector<double> d_;
	};

	inline Dense operator+(const Dense &a, const Dense &b)
	{
		assert(a.rows() == b.rows() && a.cols() == b.cols());
		Dense r(a.rows(), a.cols());
		for (long k = 0; k < a.size(); ++k)
			r(k) = a(k) + b(k);
		return r;
	}
	inline Dense operator-(const Dense &a, const Dense &b)
	{
		assert(a.rows() == b.rows() && a.cols() == b.cols());
		Dense r(a.rows(), a.cols());
		for (long k = 0; k < a.size(); ++k)
			r(k) = a(k) - b(k);
		return r;
	}
	inline Dense operator-(const Dense &a)
	{
		Dense r(a.rows(), a.cols());
		for (long k = 0; k < a.size(); ++k)
			r(k) = -a(k);
		return r;
	}
	inline Dense operator*(const Dense &a, const Dense &b)
	{
		assert(a.cols() == b.rows());
		Dense r(a.rows(), b.cols());
		for (long i = 0; i < a.rows(); ++i)
			for (long j = 0; j < b.cols(); ++j)
			{
				double s = 0;
				for (long k = 0; k < a.cols(); ++k)
					s += a(i, k) * b(k, j);
				r(i, j) = s;
			}
		return r;
	}
	inline Dense &Dense::operator*=(const Dense &o) { return *this = *this * o; }
	inline Dense operator*(double s, const Dense &a)
	{
		Dense r(a.rows(), a.cols());
		for (long k = 0; k < a.size(); ++k)
			r(k) = s * a(k);
		return r;
	}
	inline Dense operator*(const Dense &a, double s) { return s * a; }
	inline Dense operator/(const Dense &a, double s)
	{
		Dense r(a.rows(), a.cols());
		for (long k = 0; k < r.size(); ++k)
			r(k) = a(k) / s;
		return r;
	}
	inline Dense operator*(const Dense::ArrayView &a, const Dense::ArrayView &b)
	{
		assert(a.m.size() == b.m.size());
		Dense r(a.m.rows(), a.m.cols());
		for (long k = 0; k < r.size(); ++k)
			r(k) = a.m(k) * b.m(k);
		return r;
	}

	// Primary template: a small dense column-major matrix of ANY scalar type (the autodiff scalars of utils/autodiff.h inside
	// GenericElastic: 3 x 3 deformation gradients of DScalar1 / DScalar2). Only what those function bodies use.
	template <typename S, int R, int C, int Opt = 0, int MR = R, int MC = C>
	class Matrix
	{
	public:
		typedef S Scalar;
		Matrix() : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C), d_(size_t(r_) * size_t(c_)) {}
		template <typename I, typename J, typename = std::enable_if_t<std::is_integral_v<I> && std::is_integral_v<J>>>
		Matrix(I r, J c) : r_(int(r)), c_(int(c)), d_(size_t(r) * size_t(c)) {}
		template <int R2, int C2, int O2, int MR2, int MC2>
		Matrix(const Matrix<S, R2, C2, O2, MR2, MC2> &o) : r_(int(o.rows())), c_(int(o.cols())), d_(size_t(o.size()))
		{
			for (long k = 0; k < o.size(); ++k)
				d_[size_t(k)] = o(k);
		}
		long rows() const { return r_; }
		long cols() const { return c_; }
		long size() const { return long(r_) * c_; }
		void resize(long r, long c)
		{
			r_ = int(r);
			c_ = int(c);
			d_.assign(size_t(r) * size_t(c), S());
		}
		S &operator()(long i, long j) { return d_[size_t(j) * r_ + i]; }
		const S &operator()(long i, long j) const { return d_[size_t(j) * r_ + i]; }
		S &operator()(long k) { return d_[size_t(k)]; }
		const S &operator()(long k) const { return d_[size_t(k)]; }
		Matrix transpose() const
		{
			Matrix t(c_, r_);
			for (int i = 0; i < r_; ++i)
				for (int j = 0; j < c_; ++j)
					t(j, i) = (*this)(i, j);
			return t;
		}
		S trace() const
		{
			S s = (*this)(0, 0);
			for (int i = 1; i < r_ && i < c_; ++i)
				s = s + (*this)(i, i);
			return s;
		}
		Matrix &operator*=(const Matrix &o) { return *this = *this * o; }
		Matrix operator+(const Matrix &o) const
		{
			assert(r_ == o.r_ && c_ == o.c_);
			Matrix r(r_, c_);
			for (size_t k = 0; k < d_.size(); ++k)
				r.d_[k] = d_[k] + o.d_[k];
			return r;
		}
		struct CommaInit // row by row, as Eigen's comma initialiser
		{
			Matrix &m;
			long k;
			CommaInit &operator,(const S &v)
			{
				m(k / m.cols(), k % m.cols()) = v;
				++k;
				return *this;
			}
		};
		CommaInit operator<<(const S &v)
		{
			(*this)(0, 0) = v;
			return CommaInit{*this, 1};
		}

	private:
		int r_, c_;
		std::vector<S> d_;
	};
	template <typename S, int R, int C, int O, int MR, int MC, int R2, int C2, int O2, int MR2, int MC2>
	std::enable_if_t<!std::is_same_v<S, double>, Matrix<S, R, C, O, MR, MC>> operator*(const Matrix<S, R, C, O, MR, MC> &a, const Matrix<S, R2, C2, O2, MR2, MC2> &b)
	{
		assert(a.cols() == b.rows());
		Matrix<S, R, C, O, MR, MC> r(a.rows(), b.cols());
		for (long i = 0; i < a.rows(); ++i)
			for (long j = 0; j < b.cols(); ++j)
			{
				S s = a(i, 0) * b(0, j);
				for (long k = 1; k < a.cols(); ++k)
					s = s + a(i, k) * b(k, j);
				r(i, j) = s;
			}
		return r;
	}
	template <typename S, int R, int C, int O, int MR, int MC>
	std::enable_if_t<!std::is_same_v<S, double>, Matrix<S, R, C, O, MR, MC>> operator/(const Matrix<S, R, C, O, MR, MC> &a, const S &s)
	{
		Matrix<S, R, C, O, MR, MC> r(a.rows(), a.cols());
		for (long k = 0; k < a.size(); ++k)
			r(k) = a(k) / s;
		return r;
	}

	// double: the eager `Dense` matrix above
	template <int R, int C, int Opt, int MR, int MC>
	class Matrix<double, R, C, Opt, MR, MC> : public Dense
	{
		typedef double S;

	public:
		typedef S Scalar;
		typedef long Index;
		Matrix() : Dense(R == Dynamic ? 0 : R, C == Dynamic ? 0 : C) {}
		template <typename I, typename J, typename = std::enable_if_t<std::is_integral_v<I> && std::is_integral_v<J>>>
		Matrix(I r, J c) : Dense(long(r), long(c)) {}
		template <typename I, typename = std::enable_if_t<std::is_integral_v<I>>>
		explicit Matrix(I n) : Dense(C == 1 ? long(n) : 1, C == 1 ? 1 : long(n)) {}
		Matrix(const Dense &o) : Dense(o) {}
		template <typename S2, int O2, typename I2>
		Matrix(const SparseMatrix<S2, O2, I2> &sp); // dense copy of a sparse matrix, defined in mini_sparse.hpp
		Matrix &operator=(const Dense &o)
		{
			Dense::operator=(o);
			return *this;
		}
		static Dense Identity(long r, long c)
		{
			Dense m(r, c);
			for (long i = 0; i < r && i < c; ++i)
				m(i, i) = 1.0;
			return m;
		}
		static Dense Identity() { return Identity(R, C); } // fixed-size form
		static Dense Ones()                                // fixed-size form
		{
			Dense m(R, C);
			for (long k = 0; k < m.size(); ++k)
				m(k) = 1.0;
			return m;
		}
		static Dense Zero(long r, long c) { return Dense(r, c); }
		static Dense Zero(long n) { return C == 1 ? Dense(n, 1) : Dense(1, n); } // vectors
		static Dense Constant(double v) // fixed-size form only
		{
			Dense m(R, C);
			for (long k = 0; k < m.size(); ++k)
				m(k) = v;
			return m;
		}
	};
	using MatrixXd = Matrix<double, Dynamic, Dynamic>;
	using VectorXd = Matrix<double, Dynamic, 1>;
	using Matrix3d = Matrix<double, 3, 3>;
	using Matrix2d = Matrix<double, 2, 2>;
	using Vector3d = Matrix<double, 3, 1>;
	using Vector2d = Matrix<double, 2, 1>;
	template <typename S, int N>
	using Vector = Matrix<S, N, 1>;
	constexpr unsigned ComputeFullU = 0x04, ComputeFullV = 0x10; // Eigen/src/Core/util/Constants.h values (only tested for non-zero)

	// Map<T>(ptr, n): the reference only reads through it (a column vector of n entries)
	template <typename T>
	class Map : public Dense
	{
	public:
		Map(const double *p, long n) : Dense(n, 1)
		{
			for (long k = 0; k < n; ++k)
				(*this)(k) = p[k];
		}
	};
} // namespace Eigen
