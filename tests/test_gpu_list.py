"""List-driven first quarter: the canonical AO list stays on the device as uploaded
(LOWDIN_IT_OPT_AO_LIST) and every stored integral is scattered with its <= 4 permutational images into the quarter-transformed
slabs -- the reference's DIRECT first quarter (Libint2Iface.cpp:793-853, consumed by TransformIntegralsC.f90:545-558)."""
import numpy as np
import pytest

import openlowdin_b200 as ol
from helpers import assert_lists_match, dense_pairs, dense_quads


def test_restated_direct_first_quarter_is_the_dense_contraction(O):
    """CPU: the restated scatter rule against a dense einsum of the unpacked tensor."""
    n = 7
    packed = O.hash_packed_intra(2, n)
    M = O.npairs(n)
    sq = O.packed_to_square(packed, M)
    xy = O.pair_table(n)
    ao4 = O.dense4_from_square(sq, xy)
    lst = O.canonical_list_intra(packed, n)
    coef = np.random.default_rng(0).standard_normal(n)
    GG = O.direct_first_quarter(coef, *lst)
    ref = np.einsum("mnls,m->nls", ao4, coef)
    assert np.abs(GG - ref).max() <= 1e-12


@pytest.fixture
def list_mode(T):
    T.set_option(T.OPT_AO_LIST, 1)
    yield T
    T.set_option(T.OPT_AO_LIST, 0)


def _sparse(lst, keep_every=3):
    """A sparse list: real AO lists hold only |v| > 1e-10 (Libint2Iface.cpp:369)."""
    idx = np.arange(len(lst[4]))
    keep = (idx % keep_every) != 1
    return tuple(x[keep] for x in lst)


@pytest.mark.gpu
def test_list_scatter_is_bit_exact_with_identity_coefficients(O, list_mode):
    """C = identity: T1[f][slab][nu] = (f nu|slab) receives exactly ONE image of one stored integral, so the result must equal
    the dense tensor element bit for bit -- the index work of the scatter (which element each image lands in)."""
    T = list_mode
    n = 9
    M = O.npairs(n)
    packed = O.hash_packed_intra(17, n)
    lst = _sparse(O.canonical_list_intra(packed, n))
    T.set_species(0, np.eye(n))
    T.upload_ao(0, 0, *lst, stack=200)
    got = T.debug_first_quarter(0, 0, 1, n, 0, M)             # [f][slab][nu]
    sq = O.packed_to_square(O.scatter_intra(*lst, n), M)      # the same sparse tensor, densified by the oracle's loader
    i1, i2 = np.triu_indices(n)
    ref = np.zeros((M, n, n)); ref[:, i1, i2] = sq; ref[:, i2, i1] = sq     # ref[slab][mu][nu]
    assert np.array_equal(got, ref.transpose(1, 0, 2))


@pytest.mark.gpu
def test_list_first_quarter_matches_restated_direct(O, list_mode):
    T = list_mode
    n = 11
    M = O.npairs(n)
    packed = O.hash_packed_intra(23, n)
    lst = _sparse(O.canonical_list_intra(packed, n))
    Cm = O.random_orthonormal(n, n)
    T.set_species(0, Cm)
    T.upload_ao(0, 0, *lst, stack=128)
    got = T.debug_first_quarter(0, 0, 3, 4, 0, M)             # MO indices 3..6
    i1, i2 = np.triu_indices(n)
    for f in range(4):
        GG = O.direct_first_quarter(Cm[:, 2 + f], *lst)       # GG[nu][lam][sig]
        assert np.abs(got[f] - GG[:, i1, i2].T).max() <= 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("conv", ["C", "E"])
@pytest.mark.parametrize("mode", ["MP2", "ALL", "PT2"])
def test_list_mode_transform_intra_matches_restated_transformers(O, list_mode, conv, mode):
    T = list_mode
    n, occ = 13, 4
    packed_full = O.hash_packed_intra(61, n)
    lst = _sparse(O.canonical_list_intra(packed_full, n))
    packed = O.scatter_intra(*lst, n)
    Cm = O.random_orthonormal(n, n)
    T.set_species(0, Cm)
    T.upload_ao(0, 0, *lst, stack=300)
    M = O.npairs(n)
    if conv == "C":
        win, sym = O.windows_c_intra(mode, n, occ)
        got = T.transform(0, 0, win, ol.CONV_C, symmetric=sym)
        ref = O.transform_c_intra(Cm, packed, win, sym)
        assert np.abs(dense_quads(*got, n, n) - dense_quads(*ref, n, n)).max() <= 1e-10
        assert_lists_match(got[:4], got[4], ref[:4], ref[4])
    else:
        win = O.windows_e_intra(mode, n, occ)
        got = T.transform(0, 0, win, ol.CONV_E)
        ref = O.transform_e_intra(Cm, packed, win)
        assert np.abs(dense_pairs(*got, M, M) - dense_pairs(*ref, M, M)).max() <= 1e-10
        assert_lists_match(got[:2], got[2], ref[:2], ref[2])


@pytest.mark.gpu
@pytest.mark.parametrize("swapped", [False, True])
def test_list_mode_transform_inter_matches_restated_e(O, list_mode, swapped):
    T = list_mode
    na, nb, oa, ob = 10, 7, 3, 2
    rect = O.hash_rect_inter(5, na, nb)
    Ca, Cb = O.random_orthonormal(na, 1), O.random_orthonormal(nb, 2)
    T.set_species(0, Ca); T.set_species(1, Cb)
    p, q, r, s, v = O.canonical_list_inter(rect, na, nb)
    if swapped:
        T.upload_ao(0, 1, r, s, p, q, v, swapped=True, stack=64)     # the file of the reversed pair: (B B|A A)
    else:
        T.upload_ao(0, 1, p, q, r, s, v, stack=64)
    win = O.windows_e_inter("MP2", na, nb, oa, ob)
    got = T.transform(0, 1, win, ol.CONV_E)
    ref = O.transform_e_inter(Ca, Cb, rect, win)
    Ma, Mb = O.npairs(na), O.npairs(nb)
    assert np.abs(dense_pairs(*got, Ma, Mb) - dense_pairs(*ref, Ma, Mb)).max() <= 1e-10


@pytest.mark.gpu
def test_list_mode_streaming_energy_and_occupied_batches(O, list_mode):
    """Occupied batching with the list: T1 of every pass is rebuilt from the same resident list."""
    T = list_mode
    n, occ = 19, 5
    packed = O.hash_packed_intra(7, n)
    lst = O.canonical_list_intra(packed, n)
    Cm = O.random_orthonormal(n, n)
    eps = O.synthetic_eps(occ, n)
    T.set_species(0, Cm)
    T.upload_ao(0, 0, *lst, stack=1024)
    win = O.windows_e_intra("MP2", n, occ)
    rij, rkl, rv = O.transform_e_intra(Cm, packed, win)
    want = np.array([len(rv), rv.sum(), (rv * rv).sum(), O.mp2_intra_from_pairs(rij, rkl, rv, n, occ, eps)])
    T.set_option(T.OPT_CHUNK_COLS, 50)
    try:
        for qb in (0, 2, 3):
            got = T.transform_stream(0, 0, win, ol.CONV_E, occ_batch=qb, epsA=eps)
            assert got[0] == want[0] and np.abs(got[1:] - want[1:]).max() <= 1e-9
    finally:
        T.set_option(T.OPT_CHUNK_COLS, 0)
