"""The C-ABI library loads and exports every symbol include/pfa.h declares; without a GPU the
product fails loudly instead of falling back to any CPU path."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pfa.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pfa_[a-z_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from polyfem_b200 import build, capi
    build.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in pfa.h but not exported by libpfa.so"
    assert sorted(capi.EXPORTS) == declared


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from polyfem_b200 import capi, mesh, tables
    m = mesh.kuhn_cube(1, 1)
    t = tables.reference_tables(1)
    with pytest.raises(capi.PfaError) as ei:
        capi.Handle("NeoHookean", m.conn, m.n_bases, t["weights"], t["grad"], vertices=m.vertices, lam=1.0, mu=1.0)
    assert ei.value.code == capi.PFA_ERR_NO_DEVICE


def test_create_rejects_bad_descriptions():
    from polyfem_b200 import capi
    L = capi.lib()
    d = capi.MeshDesc()
    h = ctypes.c_void_p()
    assert L.pfa_create(ctypes.byref(d), ctypes.byref(h)) == capi.PFA_ERR_INVALID  # struct_size == 0
    assert b"struct_size" in L.pfa_last_error(None)
    d.struct_size = ctypes.sizeof(capi.MeshDesc)
    d.material = 7
    assert L.pfa_create(ctypes.byref(d), ctypes.byref(h)) == capi.PFA_ERR_INVALID
    assert L.pfa_create(None, ctypes.byref(h)) == capi.PFA_ERR_INVALID


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (tier rule)."""
    pkg = os.path.join(ROOT, "polyfem_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".hpp", ".cpp", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in src and "liboracle" not in src and "oracle/" not in src.replace("oracle/_ref", ""), f


def test_owned_nodes_is_refused_off_the_owner_computes_path():
    """pfa_mesh_desc.owned_nodes only has a meaning for the owner-computes kernels (NeoHookean P1/P2, affine, no
    PFA_FLAG_ROW_LANE): every other combination is refused before a device is touched, so the check runs here too."""
    from polyfem_b200 import capi, mesh, tables
    m = mesh.kuhn_cube(2, 1)
    t = tables.reference_tables(1)
    own = np.ones(m.n_bases, np.uint8)
    for material, flags in (("LinearElasticity", 0), ("NeoHookean", capi.FLAG_ROW_LANE)):
        with pytest.raises(capi.PfaError) as ei:
            capi.Handle(material, m.conn, m.n_bases, t["weights"], t["grad"], vertices=m.vertices, lam=1.0, mu=1.0, owned_nodes=own, flags=flags)
        assert ei.value.code == capi.PFA_ERR_UNSUPPORTED and "owned_nodes" in str(ei.value)
    # ghost elements without geometry cannot serve owned columns
    with pytest.raises(capi.PfaError) as ei:
        capi.Handle("NeoHookean", np.vstack([m.conn, m.conn[:1]]), m.n_bases, t["weights"], t["grad"], vertices=m.vertices, lam=1.0, mu=1.0,
                    owned_nodes=own, n_ghost_elements=1)
    assert ei.value.code == capi.PFA_ERR_UNSUPPORTED


def test_pinned_host_allocation_fails_softly_without_a_device():
    """pfa_host_alloc returns NULL when no device can pin memory (the shim's HostValues then uses ordinary memory);
    pfa_host_free(NULL) is a no-op."""
    import torch
    from polyfem_b200 import capi
    L = capi.lib()
    L.pfa_host_alloc.restype = ctypes.c_void_p
    L.pfa_host_alloc.argtypes = [ctypes.c_size_t]
    L.pfa_host_free.argtypes = [ctypes.c_void_p]
    L.pfa_host_free(None)
    assert L.pfa_host_alloc(0) is None
    p = L.pfa_host_alloc(1 << 20)
    if torch.cuda.is_available():
        assert p is not None
        ctypes.memset(p, 0, 1 << 20)
        L.pfa_host_free(p)
    else:
        assert p is None

