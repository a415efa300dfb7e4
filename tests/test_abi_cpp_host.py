"""A C++ host driving libpfa through include/pfa.h alone (tests/abi_cpp_host.cpp): no Python or torch between the caller and the
C ABI, like the PolyFEM shim. The host-only half (partition) runs everywhere, the assembling half needs the GPU."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "polyfem_b200")


def _build(tmp_path):
    if not os.path.exists(os.path.join(LIBDIR, "libpfa.so")):
        pytest.skip("libpfa.so not built")
    exe = str(tmp_path / "abi_cpp_host")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_cpp_host.cpp"), "-o", exe, "-L", LIBDIR, "-lpfa", "-Wl,-rpath," + LIBDIR],
                   check=True)
    return exe


def test_cpp_host_partition(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, "partition"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank 1" in out.stdout
    # one Kuhn cell: nodes 0 and 7 see all 8 nodes, the six others 0, 7, themselves and their two tet neighbours... = 2*8 + 6*5 pairs
    assert "host pattern: 46 node pairs" in out.stdout


@pytest.mark.gpu
def test_cpp_host_assembles_and_two_rank_columns_match(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, "assemble"], capture_output=True, text=True, timeout=300)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "24 owned columns" in out.stdout
