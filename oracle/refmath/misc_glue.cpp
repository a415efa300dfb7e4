// TEST INFRASTRUCTURE: the reference's OWN AssemblerUtils::quadrature_order (assembler/AssemblerUtils.cpp:201-245) and
// convert_to_lambda / convert_to_mu (utils/ElasticityUtils.cpp:12-23), extracted at build time into
// ../_ref/misc_extracted.inc and compiled verbatim -> oracle/_ref/libmiscref.so. Pins polyfem_b200/tables.py::
// quadrature_order and polyfem_b200/mesh.py::lame_from_E_nu.
#include <algorithm>
#include <string>

namespace polyfem
{
	namespace assembler
	{
		class AssemblerUtils // assembler/AssemblerUtils.hpp:14-40
		{
		public:
			enum class BasisType
			{
				SIMPLEX_LAGRANGE,
				CUBE_LAGRANGE,
				PRISM_LAGRANGE,
				PYRAMID_LAGRANGE,
				SPLINE,
				POLY
			};
			static int quadrature_order(const std::string &assembler, const int basis_degree, const BasisType &b_type, const int dim);
		};
	} // namespace assembler
	using namespace assembler;

#include "../_ref/misc_extracted.inc"
} // namespace polyfem

extern "C"
{
	// basis_type: 0 simplex Lagrange (the only one on the hot path), 1 cube Lagrange, ...
	int ref_quadrature_order(const char *assembler, int basis_degree, int basis_type, int dim)
	{
		return polyfem::assembler::AssemblerUtils::quadrature_order(assembler, basis_degree, polyfem::assembler::AssemblerUtils::BasisType(basis_type), dim);
	}
	double ref_convert_to_lambda(int is_volume, double E, double nu) { return polyfem::convert_to_lambda(is_volume != 0, E, nu); }
	double ref_convert_to_mu(double E, double nu) { return polyfem::convert_to_mu(E, nu); }
}
