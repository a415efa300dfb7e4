"""PFA_FLAG_LARGE_INDEX (utils/Types.hpp:21-25, POLYSOLVE_LARGE_INDEX: StiffnessMatrix with std::ptrdiff_t indices): the int64
pattern equals the int32 pattern entry by entry, the values are the same, and the int32 accessors refuse. The case the flag
exists for - BASELINE cfg 4, LinearElasticity P4 n = 32 with 2.5 G nnz - is `python bench.py --config 4le --n 32`
(profiles/), too large for the test suite."""
import numpy as np
import pytest

from helpers import gpu_handle, make_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("material,p,n", [("LinearElasticity", 2, 3), ("Laplacian", 3, 2), ("NeoHookean", 2, 2)])
def test_int64_pattern_equals_int32_pattern(material, p, n):
    from polyfem_b200 import capi
    mesh, x, t = make_case(n, p, jitter=0.1)
    h32 = gpu_handle(mesh, material, t)
    h64 = gpu_handle(mesh, material, t, flags=capi.FLAG_LARGE_INDEX)
    o32, i32 = h32.pattern()
    o64, i64 = h64.pattern_wide()
    assert o64.dtype == np.int64 and i64.dtype == np.int64
    assert np.array_equal(o32.astype(np.int64), o64) and np.array_equal(i32.astype(np.int64), i64)
    if material == "NeoHookean":
        xx = x[: mesh.n_bases * 3]
        assert np.array_equal(h32.hessian(xx), h64.hessian(xx))
    else:
        a, b = h32.linear_stiffness(), h64.linear_stiffness()  # RED accumulation: same sums in another order
        assert np.abs(a - b).max() <= 1e-13 * np.abs(a).max()
    with pytest.raises(capi.PfaError) as ei:
        h64.pattern()
    assert ei.value.code == capi.PFA_ERR_UNSUPPORTED
    with pytest.raises(capi.PfaError):
        h32.pattern_wide()
