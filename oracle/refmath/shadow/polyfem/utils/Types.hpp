// shadow of polyfem/utils/Types.hpp for oracle/refmath: only the typedef utils/MatrixCache.* needs (Types.hpp:24)
#pragma once
#include <Eigen/Dense>
#include <Eigen/Sparse>
namespace polyfem
{
	typedef Eigen::SparseMatrix<double, Eigen::ColMajor> StiffnessMatrix;
}
