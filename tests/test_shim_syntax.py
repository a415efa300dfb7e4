"""The PolyFEM-side C++ binding (polyfem_b200/host/assembler_shim.hpp) cannot be built here
(no Eigen, no PolyFEM), but it can be syntax- and type-checked against declaration-only
stand-ins of the reference types it touches (tests/stubs/, signatures transcribed from the
reference headers): every `override` must match a virtual of the base class, every pfa_* call
must match include/pfa.h."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TU = r"""
#include <assembler_shim.hpp>
using namespace polyfem;
using namespace polyfem::assembler;
int main()
{
	b200::NeoHookeanElasticityB200 nh;
	b200::LinearElasticityB200 le;
	b200::LaplacianB200 lap;
	b200::MassB200 mass;
	b200::SaintVenantElasticityB200 sv;
	b200::MooneyRivlinElasticityB200 mr;
	b200::ViscousDampingB200 vd;
	b200::FixedCorotationalB200 fc;
	std::vector<basis::ElementBases> bases, gbases;
	AssemblyValsCache cache;
	Eigen::MatrixXd x, rhs;
	StiffnessMatrix K;
	utils::MatrixCache mc;
	const Assembler &a = nh; // what ElasticForm holds (ElasticForm.hpp:107)
	double e = a.assemble_energy(true, bases, gbases, cache, 0.0, 1.0, x, x);
	Eigen::VectorXd epe = a.assemble_energy_per_element(true, bases, gbases, cache, 0.0, 1.0, x, x);
	a.assemble_gradient(true, 1, bases, gbases, cache, 0.0, 1.0, x, x, rhs);
	a.assemble_hessian(true, 1, false, bases, gbases, cache, 0.0, 1.0, x, x, mc, K);
	static_cast<const Assembler &>(le).assemble(true, 1, bases, gbases, cache, 0.0, K);
	for (const Assembler *nl : {static_cast<const Assembler *>(&sv), static_cast<const Assembler *>(&mr), static_cast<const Assembler *>(&vd), static_cast<const Assembler *>(&fc)})
	{
		e += nl->assemble_energy(true, bases, gbases, cache, 0.0, 1.0, x, x);
		nl->assemble_gradient(true, 1, bases, gbases, cache, 0.0, 1.0, x, x, rhs);
		nl->assemble_hessian(true, 1, true, bases, gbases, cache, 0.0, 1.0, x, x, mc, K);
	}
	static_cast<const Assembler &>(lap).assemble(true, 1, bases, gbases, cache, 0.0, K);
	static_cast<const Assembler &>(mass).assemble(true, 1, bases, gbases, cache, 0.0, K, true);
	pfa_handle *h = nullptr;
	std::vector<int> bn;
	std::vector<double> values;
	Eigen::VectorXd xf, g;
	b200::ReducedNewtonSystem::set_constraints(h, bn);
	e += b200::ReducedNewtonSystem::assemble(h, xf, false, 1.0, g, K, values);
	bool ok = b200::ReducedNewtonSystem::is_step_valid(h, xf, e);
	return ok && epe.size() ? 0 : 1;
}
"""


def test_shim_compiles_against_reference_signatures(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    src = tmp_path / "shim_check.cpp"
    src.write_text(TU)
    cmd = [gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror=overloaded-virtual",
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "polyfem_b200", "host"),
           "-I", os.path.join(ROOT, "tests", "stubs"), str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


RUNTIME_TU = r"""
#include <assembler_shim.hpp>
#include <cstdio>
using polyfem::assembler::b200::DeviceAssembly;
int main()
{
	// one value per element in every array: [n_el][n_qp] -> [n_el], stride 1
	std::vector<double> a = {1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3}, b = {5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7}, c;
	if (DeviceAssembly::compress_if_uniform(3, 4, a, b, c) != 1 || a != std::vector<double>{1, 2, 3} || b != std::vector<double>{5, 6, 7} || !c.empty())
		return 1;
	// one element varies inside: everything stays per quadrature point
	std::vector<double> d = {1, 1, 1, 1, 2, 2, 2.5, 2}, e = {5, 5, 5, 5, 6, 6, 6, 6}, f = e;
	if (DeviceAssembly::compress_if_uniform(2, 4, d, e, f) != 4 || d.size() != 8 || e.size() != 8 || f.size() != 8)
		return 2;
	std::vector<double> g = {1, 2}, h = {3, 4}, k;
	if (DeviceAssembly::compress_if_uniform(2, 1, g, h, k) != 1 || g.size() != 2)
		return 3;
	// values[] buffer: pinned when a device grants it, ordinary memory otherwise (here); grows, keeps its pointer while it fits,
	// and a copy starts empty instead of sharing (or double-freeing) the allocation
	polyfem::assembler::b200::HostValues hv;
	double *q = hv.resize(10);
	if (!q || hv.data() != q)
		return 4;
	for (int i = 0; i < 10; ++i)
		q[i] = i;
	if (hv.resize(5) != q)
		return 5;
	double *q2 = hv.resize(1000);
	if (!q2)
		return 6;
	q2[999] = 1.0;
	polyfem::assembler::b200::HostValues copy(hv);
	if (copy.data() != nullptr)
		return 7;
	copy = hv;
	if (copy.data() != nullptr || copy.resize(3) == nullptr)
		return 8;
	std::puts("ok");
	return 0;
}
"""


def test_shim_material_layout_helper_runs(tmp_path):
    """The one piece of host logic in the shim that does not need PolyFEM: per-element parameters are detected and sent with
    material_stride 1 (the library's per-element kernels), per-quadrature-point ones with stride n_qp."""
    gxx = shutil.which("g++")
    libdir = os.path.join(ROOT, "polyfem_b200")
    if gxx is None or not os.path.exists(os.path.join(libdir, "libpfa.so")):
        pytest.skip("g++ or libpfa.so not available")
    src = tmp_path / "shim_rt.cpp"
    src.write_text(RUNTIME_TU)
    exe = str(tmp_path / "shim_rt")
    subprocess.run([gxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "polyfem_b200", "host"),
                    "-I", os.path.join(ROOT, "tests", "stubs"), str(src), "-o", exe, "-L", libdir, "-lpfa", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
