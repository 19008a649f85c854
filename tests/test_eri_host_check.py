"""The product's device ERI code (openlowdin_b200/csrc/it_eri.cuh: McMurchie-Davidson + Boys function) compiled for the HOST by
tests/mock/build_mock.py and compared with the oracle (Rys-form 2-D recurrences + Gauss-Legendre): two independent algorithms
for every angular-momentum class up to (ff|ff).  CPU only; the -m gpu twin is tests/test_gpu_eri.py."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from eri_cases import h2_sto3g, nbf, water_like, SZABO_H2

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "mock"))


@pytest.fixture(scope="module")
def host_eri():
    import build_mock
    from openlowdin_b200 import capi
    L = C.CDLL(build_mock.build_eri())
    f64 = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    L.mock_eri_packed_intra.argtypes = [C.c_int, C.POINTER(capi.Shell), f64, f64, f64, f64]

    def run(shells):
        arr, ex, co, n = capi.pack_shells(shells)
        M = n * (n + 1) // 2
        packed, norma = np.zeros(M * (M + 1) // 2), np.zeros(n)
        assert L.mock_eri_packed_intra(len(shells), arr, ex, co, packed, norma) == 0
        return packed, norma
    return run


def test_device_algorithm_reproduces_textbook_h2(O, host_eri):
    packed, norma = host_eri(h2_sto3g(O))
    M = 3    # pairs (0,0) (0,1) (1,1)
    pid = {(0, 0): 0, (0, 1): 1, (1, 0): 1, (1, 1): 2}
    for (i, j, k, l), ref in SZABO_H2.items():
        lo, hi = sorted((pid[(i, j)], pid[(k, l)]))
        assert abs(packed[lo * M - lo * (lo + 1) // 2 + hi] - ref) < 1e-4
    assert np.allclose(norma, O.eri_norma(h2_sto3g(O)), rtol=1e-14)


@pytest.mark.parametrize("with_f", [False, True])
def test_device_algorithm_matches_oracle_spdf(O, host_eri, with_f):
    sh = water_like(with_f)
    got, norma = host_eri(sh)
    ref = O.eri_packed_intra(sh)
    assert np.allclose(norma, O.eri_norma(sh), rtol=1e-13)
    both = (got != 0) & (ref != 0)
    assert both.sum() > 0.3 * ref.size     # the rest are symmetry zeros (two centres in the x = 0 plane)
    assert np.abs(got - ref)[both].max() < 2e-12, np.abs(got - ref)[both].max()
    edge = (got != 0) != (ref != 0)           # dropped by one side only: raw value at the 1e-10 filter
    assert np.abs(got - ref)[edge].max(initial=0.0) < 1e-8
    assert np.abs(ref).max() > 1.0            # the (ss|ss) core integrals are of order 1..5
