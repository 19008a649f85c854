"""Kind-K (rank-8 separable) synthetic AO tensors: the one input whose MO integrals have a closed form
(p q|r s) = sum_k (C^T L^k C)[p,q] (C^T L^k C)[r,s]  (SURVEY.md 8d), i.e. an O(K N^3) oracle at ANY size.  The small cases
pin the closed form and the device generator to the restated transformer E; the large ones check the CUDA path at the
benchmark's own sizes (N_bf = 500 and 1500), where no CPU can run the transform."""
import numpy as np
import pytest

import openlowdin_b200 as ol
from helpers import dense_pairs, rankk_stream_sums


def _mp2_win(n, occ):  # E.f90:1938-1949
    return [occ + 1, n, 1, occ, occ + 1, n, 1, occ]


def test_closed_form_sums_equal_the_oracle_energy(O):
    """CPU: the closed-form stream sums (helpers.rankk_stream_sums) against restated transformer E + the MP2 reader."""
    n, occ = 12, 3
    L, lv = O.rankk_factors(5, n)
    sq = lv.T @ lv
    packed = O.square_to_packed(sq)
    Cm = O.random_orthonormal(n, n)
    eps = O.synthetic_eps(occ, n)
    win = _mp2_win(n, occ)
    rij, rkl, rv = O.transform_e_intra(Cm, packed, win)
    want = np.array([len(rv), rv.sum(), (rv * rv).sum(), O.mp2_intra_from_pairs(rij, rkl, rv, n, occ, eps)])
    Tvo = O.rankk_mo_factors(L, Cm, np.arange(occ, n), np.arange(occ))
    got = rankk_stream_sums(Tvo, eps, occ, range(occ))
    assert got[0] == want[0]
    assert np.abs(got[1:] - want[1:]).max() <= 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("materialize", [False, True])
def test_rankk_small_intra_matches_restated_e(O, T, materialize):
    n, occ = 19, 5
    L, lv = O.rankk_factors(11, n)
    packed = O.square_to_packed(lv.T @ lv)
    Cm = O.random_orthonormal(n, n)
    T.set_species(0, Cm)
    T.set_rankk(0, 0, lv)
    if materialize:
        T.materialize(0, 0)
    M = O.npairs(n)
    X = T.debug_expand(0, 0, 3, 4)
    sq = O.packed_to_square(packed, M)
    i1, i2 = np.triu_indices(n)
    for z in range(4):
        ref = np.zeros((n, n)); ref[i1, i2] = sq[3 + z]; ref[i2, i1] = sq[3 + z]
        assert np.abs(X[z] - ref).max() <= 1e-13
    for mode in ("MP2", "ALL"):
        win = O.windows_e_intra(mode, n, occ) if mode == "MP2" else [1, n] * 4
        ij, kl, v = T.transform(0, 0, win, ol.CONV_E)
        rij, rkl, rv = O.transform_e_intra(Cm, packed, win)
        assert np.abs(dense_pairs(ij, kl, v, M, M) - dense_pairs(rij, rkl, rv, M, M)).max() <= 1e-10


@pytest.mark.gpu
def test_rankk_small_inter_matches_restated_e(O, T):
    na, nb, oa, ob = 12, 9, 3, 2
    rect, La, Lb = O.rankk_square(21, na, nb)     # [M_a, M_b] pair matrix
    iu, iub = np.triu_indices(na), np.triu_indices(nb)
    Ca, Cb = O.random_orthonormal(na, 3), O.random_orthonormal(nb, 4)
    T.set_species(0, Ca); T.set_species(1, Cb)
    T.set_rankk(0, 1, La[:, iu[0], iu[1]], Lb[:, iub[0], iub[1]])
    win = O.windows_e_inter("MP2", na, nb, oa, ob)
    stored = np.ascontiguousarray(rect.T).ravel()  # AO storage (rs-1)*M_a + pq  (C.f90:882)
    rij, rkl, rv = O.transform_e_inter(Ca, Cb, stored, win)
    Ma, Mb = O.npairs(na), O.npairs(nb)
    for mat in (False, True):
        if mat:
            T.materialize(0, 1)
        ij, kl, v = T.transform(0, 1, win, ol.CONV_E)
        assert np.abs(dense_pairs(ij, kl, v, Ma, Mb) - dense_pairs(rij, rkl, rv, Ma, Mb)).max() <= 1e-10


def _large_case(O, T, n, seed):
    occ = n // 10
    L, lv = O.rankk_factors(seed, n)
    Cm = O.random_orthonormal(n, n)
    eps = O.synthetic_eps(occ, n)
    T.set_species(0, Cm)
    T.set_rankk(0, 0, lv)
    return occ, L, Cm, eps


@pytest.mark.gpu
def test_rankk_n500_downloaded_integrals_match_closed_form(O, T):
    """N_bf = 500: every MO integral of a window slice ((a i|b j), 6 virtuals x 3 occupied on the first pair, the whole MP2
    window on the second) against the closed form, max |delta| <= 1e-10."""
    n = 500
    occ, L, Cm, eps = _large_case(O, T, n, 500)
    win = [occ + 1, occ + 6, 2, 4, occ + 1, n, 1, occ]
    ij, kl, v = T.transform(0, 0, win, ol.CONV_E)
    a, i = np.arange(occ, occ + 6), np.arange(1, 4)
    b, j = np.arange(occ, n), np.arange(occ)
    ref = O.rankk_mo_block(L, Cm, a, i, rows_b=b, cols_b=j)       # [a, i, b, j]
    M = O.npairs(n)
    pid = lambda x, y: np.minimum(x, y) * n - np.minimum(x, y) * (np.minimum(x, y) - 1) // 2 + np.abs(x - y) + 1  # noqa: E731
    want = {}
    A, I, B, J = np.meshgrid(a, i, b, j, indexing="ij")
    keys = (pid(A, I).ravel().astype(np.int64) - 1) * M + (pid(B, J).ravel().astype(np.int64) - 1)
    got = np.zeros(keys.shape)
    order = np.argsort(keys)
    pos = np.searchsorted(keys[order], (ij - 1) * M + (kl - 1))
    assert np.array_equal(keys[order][pos], (ij - 1) * M + (kl - 1))
    got[order[pos]] = v
    d = np.abs(got - ref.ravel())
    assert d.max() <= 1e-10, d.max()
    assert len(v) >= 0.99 * len(keys)          # |x| <= 1e-10 is rare


@pytest.mark.gpu
@pytest.mark.parametrize("stored", [False, True])
def test_rankk_n500_stream_sums_match_closed_form(O, T, stored):
    """N_bf = 500, whole MP2 transform (two occupied batches, several AO-pair chunks): count, sum, sum of squares and the MP2
    pair energy against the closed form; `stored` runs the same tensor materialised in HBM (63 GB packed) through the stored-AO kernels."""
    n = 500
    occ, L, Cm, eps = _large_case(O, T, n, 501)
    try:
        if stored:
            T.materialize(0, 0)
        got = T.transform_stream(0, 0, _mp2_win(n, occ), ol.CONV_E, occ_batch=25, epsA=eps)
    finally:
        T.set_generator(0, 0, 1)                  # release the 63 GB tensor whatever happens
    Tvo = O.rankk_mo_factors(L, Cm, np.arange(occ, n), np.arange(occ))
    want = rankk_stream_sums(Tvo, eps, occ, range(occ))
    assert got[0] == want[0]
    assert abs(got[1] - want[1]) <= 1e-9 * max(1.0, abs(want[1]))
    assert abs(got[2] - want[2]) <= 1e-10 * want[2]
    assert abs(got[3] - want[3]) <= 1e-9 * max(1.0, abs(want[3])), (got[3], want[3])


@pytest.mark.gpu
def test_rankk_n1500_one_occupied_batch_matches_closed_form(O, T):
    """The benchmark's own size: one pass of 8 occupied orbitals of the N_bf = 1500 MP2 transform (all 1 125 750 AO-pair slabs,
    chunked second half) against the closed form."""
    n = 1500
    occ, L, Cm, eps = _large_case(O, T, n, 1500)
    qb = 8
    got = T.transform_stream(0, 0, _mp2_win(n, occ), ol.CONV_E, occ_batch=qb, first_pass=1, n_passes=1, epsA=eps)
    Tvo = O.rankk_mo_factors(L, Cm, np.arange(occ, n), np.arange(occ))
    want = rankk_stream_sums(Tvo, eps, occ, range(qb, 2 * qb))
    assert got[0] == want[0]
    assert abs(got[1] - want[1]) <= 1e-9 * max(1.0, abs(want[1]))
    assert abs(got[2] - want[2]) <= 1e-10 * want[2]
    assert abs(got[3] - want[3]) <= 1e-9 * max(1.0, abs(want[3])), (got[3], want[3])
