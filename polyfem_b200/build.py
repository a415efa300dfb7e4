"""In-tree build of libpfa.so (CUDA kernels + C-ABI) for sm_100a with nvcc.

The shared library is git-ignored but travels to the GPU box with the gpurun snapshot.
`python -m polyfem_b200.build` or `__graft_entry__.build()` run this.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpfa.so")
# experiment hook: PFA_DEFS="-DFOO=1 ..." PFA_LIB_SUFFIX=_foo builds polyfem_b200/libpfa_foo.so
SOURCES = ["pfa_api.cu", "pfa_pattern.cu", "pfa_kernels.cu", "pfa_project.cu", "pfa_collane2.cu", "pfa_partition.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: libpfa.so cannot be built (there is no CPU fallback)")
    return cand


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "pfa.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    suffix = os.environ.get("PFA_LIB_SUFFIX", "")
    defs = os.environ.get("PFA_DEFS", "").split()
    lib = LIB if not suffix else os.path.join(HERE, f"libpfa{suffix}.so")
    if not force and not suffix and not _stale():
        return LIB
    nvcc = _nvcc()
    flags = [f for f in FLAGS if not f.startswith("--use_fast_math")] + defs
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", f"{suffix}.o"))
        cmd = [nvcc, *ARCH, *flags, "-Xptxas", "-v" if verbose else "-warn-spills", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.check_call([nvcc, *ARCH, "-shared", "-o", lib, *objs, "-lcudart"])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
