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


@pytest.fixture(scope="module")
def host_fill():
    import build_mock
    from openlowdin_b200 import capi
    L = C.CDLL(build_mock.build_eri())
    f64 = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    PS = C.POINTER(capi.Shell)
    L.mock_eri_fill.argtypes = [C.c_int, C.c_int, PS, f64, f64, C.c_int, PS, f64, f64, C.c_int, C.c_int, C.c_int, f64, C.POINTER(C.c_int64)]

    def run(mode, sa, sb, count, logB=0, G=1, rank=0):
        aa, ea, ca, _ = capi.pack_shells(sa)
        ab, eb, cb, _ = capi.pack_shells(sb)
        dst = np.full(count, np.nan)
        nrows = C.c_int64()
        assert L.mock_eri_fill(mode, len(sa), aa, ea, ca, len(sb), ab, eb, cb, logB, G, rank, dst, C.byref(nrows)) == 0
        return dst, nrows.value
    return run


def test_kernel_grid_mapping_of_the_three_storage_modes(O, host_fill):
    """eri_fill_element over eri_fill_grid on the host: every element of the packed tensor / the inter-species rectangle / a rank's
    rows is written exactly where the transformer reads it."""
    from eri_cases import nuclear_like
    sa, sb = water_like()[:5], nuclear_like()     # 13 and 10 functions: 91 and 55 pairs
    na, nb = nbf(sa), nbf(sb)
    Ma, Mb = na * (na + 1) // 2, nb * (nb + 1) // 2
    ref = O.eri_packed_intra(sa)
    got, _ = host_fill(0, sa, sa, Ma * (Ma + 1) // 2)
    assert not np.isnan(got).any() and np.abs(got - ref).max() < 2e-12
    rect = O.eri_rect_inter(sa, sb)
    got, nrows = host_fill(1, sa, sb, Ma * Mb)
    assert nrows == Mb and not np.isnan(got).any() and np.abs(got.reshape(Mb, Ma) - rect).max() < 2e-12
    # rows of rank 1 of 3, blocks of 4 slabs: local row r is global slab (r // 4 * 3 + 1) * 4 + r % 4, the whole symmetric M-vector
    sq = np.zeros((Ma, Ma)); iu = np.triu_indices(Ma); sq[iu] = ref; sq = sq + np.triu(sq, 1).T
    owned = [s for s in range(Ma) if (s >> 2) % 3 == 1]
    got, nrows = host_fill(2, sa, sa, len(owned) * Ma, logB=2, G=3, rank=1)
    assert nrows == len(owned) and not np.isnan(got).any()
    assert np.abs(got.reshape(len(owned), Ma) - sq[owned]).max() < 2e-12
