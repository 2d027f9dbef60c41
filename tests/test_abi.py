"""The C-ABI library loads and exports every symbol include/lsx.h declares; plan queries (pure
host code) agree with the device model.  No GPU needed, no compute calls."""
import ctypes
import os
import re

import pytest

from tests import device_model as dm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "lsx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lsx_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from linalg_solver_b200 import _lib
    names = header_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(_lib.lib, name), "liblsx.so does not export %s" % name
        assert name in _lib.SIGNATURES, "no ctypes signature for %s" % name
    assert _lib.lib.lsx_abi_version() == 1


def test_plan_struct_layout_matches_header():
    from linalg_solver_b200 import _lib
    assert ctypes.sizeof(_lib.Plan) == 10 * 4 + 2 * 8 + 8


@pytest.mark.parametrize("n,a,exp_k,exp_l", [(4, 5, 1, 1), (8, 5, 2, 1), (64, 5, 12, 11)])
def test_plan_inverse_counts(n, a, exp_k, exp_l):
    from linalg_solver_b200 import _lib
    p = _lib.Plan()
    assert _lib.lib.lsx_plan_inverse(n, a, ctypes.byref(p)) == 0
    bits = dm.log2_minor_bound(n, n, True, a, 1, True, n)
    K, L = dm.plan_bits(bits)
    assert (p.n_primes, p.limbs) == (K, L) == (exp_k, exp_l)
    assert abs(p.log2_bound - bits) < 1e-9
    assert (p.m, p.n, p.bar_col, p.pivot_slots) == (n, 2 * n, n, n)


def test_plan_solve_and_rref_counts():
    from linalg_solver_b200 import _lib
    p = _lib.Plan()
    assert _lib.lib.lsx_plan_solve(16, 16, 250, 2000, 10, 16, ctypes.byref(p)) == 0
    bits = dm.log2_minor_bound(16, 16, True, 250, 2000, False, 10)
    assert (p.n_primes, p.limbs) == dm.plan_bits(bits)
    assert (p.m, p.n, p.bar_col, p.gen_cap) == (16, 17, 16, 16)
    assert _lib.lib.lsx_plan_rref(4, 4, 3, 5, 5, 0, ctypes.byref(p)) == 0
    assert (p.n_primes, p.limbs, p.pivot_slots) == (1, 1, 3)


def test_plan_shape_errors():
    from linalg_solver_b200 import _lib
    p = _lib.Plan()
    assert _lib.lib.lsx_plan_rref(0, 4, 3, 5, 5, 0, ctypes.byref(p)) == _lib.ERR_BAD_SHAPE
    assert _lib.lib.lsx_plan_rref(4, 4, 5, 5, 5, 0, ctypes.byref(p)) == _lib.ERR_BAD_SHAPE
    assert _lib.lib.lsx_plan_inverse(0, 5, ctypes.byref(p)) == _lib.ERR_BAD_SHAPE
    assert _lib.lib.lsx_plan_inverse(200, 2**31 - 1, ctypes.byref(p)) == _lib.ERR_BOUND


def test_large_det_prime_count():
    from linalg_solver_b200 import _lib
    k, bits = ctypes.c_int(), ctypes.c_double()
    assert _lib.lib.lsx_det_large_prime_count(4096, 5, ctypes.byref(k), ctypes.byref(bits)) == 0
    assert 34086 < bits.value < 34088          # 2^34087
    assert k.value == dm.plan_bits(bits.value)[0] == 1100


def test_no_device_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from linalg_solver_b200 import Engine, LsxError
    with pytest.raises(LsxError):
        Engine(0)
