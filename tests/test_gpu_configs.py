"""GPU parity at the shapes of BASELINE.json's configs: H2O/6-311G (N=19, occ 5), H2O.APMO
(e- N=19 occ 5; H nuclei N=50 occ 1: intra + inter), HCN.e+ (e- N=53 occ 7; e+ N=30 occ 1), C6H6/cc-pVDZ (N=120, occ 21).
Inputs are synthetic (the reference's libint2 + SCF cannot run here); the oracle is the restated transformer for the
small shapes and the closed form of a rank-K separable AO tensor at N=120, where the CPU restatement would take minutes."""
import numpy as np
import pytest

import openlowdin_b200 as ol
from helpers import assert_lists_match, dense_pairs, dense_quads

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _upload_intra(O, T, slot, n, seed, cseed):
    packed = O.hash_packed_intra(seed, n)
    Cm = O.random_orthonormal(n, cseed)
    T.set_species(slot, Cm)
    T.upload_ao(slot, slot, *O.canonical_list_intra(packed, n), stack=30000)
    return packed, Cm


def test_h2o_apmo_shapes_intra_and_inter(O, T):
    """H2O.APMO.MP2: electrons (19, occ 5) and a proton species (50, occ 1); intra C and E, inter C and E."""
    ne, nh, oe, oh = 19, 50, 5, 1
    pe, Ce = _upload_intra(O, T, 0, ne, 1, 11)
    eps = O.synthetic_eps(oe, ne)
    win, sym = O.windows_c_intra("MP2", ne, oe)
    ref = O.transform_c_intra(Ce, pe, win, sym)
    got = T.transform(0, 0, win, ol.CONV_C, symmetric=sym)
    assert np.abs(dense_quads(*got, ne, ne) - dense_quads(*ref, ne, ne)).max() <= TOL
    assert abs(O.mp2_intra_from_quads(*got, ne, oe, eps) - O.mp2_intra_from_quads(*ref, ne, oe, eps)) <= 1e-9
    # proton species: generated tensor (N=50: 813k unique AO integrals), E convention
    Ch = O.random_orthonormal(nh, 12)
    ph = O.hash_packed_intra(2, nh)
    T.set_species(1, Ch)
    T.set_generator(1, 1, 2)
    wine = O.windows_e_intra("MP2", nh, oh)
    rij, rkl, rv = O.transform_e_intra(Ch, ph, wine)
    ij, kl, v = T.transform(1, 1, wine, ol.CONV_E)
    Mh = O.npairs(nh)
    assert np.abs(dense_pairs(ij, kl, v, Mh, Mh) - dense_pairs(rij, rkl, rv, Mh, Mh)).max() <= TOL
    assert_lists_match((ij, kl), v, (rij, rkl), rv)
    # inter e-/H: 190 x 1275 AO pair matrix
    rect = O.hash_rect_inter(3, ne, nh)
    T.set_generator(0, 1, 3)
    winc, symc = O.windows_c_inter("MP2", ne, nh, oe, oh)
    refc = O.transform_c_inter(Ce, Ch, rect, winc, symc)
    gotc = T.transform(0, 1, winc, ol.CONV_C, symmetric=symc)
    assert np.abs(dense_quads(*gotc, ne, nh) - dense_quads(*refc, ne, nh)).max() <= TOL
    ea, eb = O.synthetic_eps(oe, ne), O.synthetic_eps(oh, nh)
    e1 = O.mp2_inter_from_quads(*gotc, ne, nh, oe, oh, ea, eb)
    e2 = O.mp2_inter_from_quads(*refc, ne, nh, oe, oh, ea, eb)
    assert abs(e1 - e2) <= 1e-9
    wine2 = O.windows_e_inter("MP2", ne, nh, oe, oh)
    r2 = O.transform_e_inter(Ce, Ch, rect, wine2)
    g2 = T.transform(0, 1, wine2, ol.CONV_E)
    assert np.abs(dense_pairs(*g2, O.npairs(ne), Mh) - dense_pairs(*r2, O.npairs(ne), Mh)).max() <= TOL


def test_hcn_positron_shapes_pt2(O, T):
    """HCN.e+ with propagatorTheoryCorrection=2: e- N=53 occ 7, e+ N=30 occ 1; PT2 windows,
    intra (C, symmetric=.false.) and the electron-positron inter-species transform (E)."""
    ne, npos, oe, op = 53, 30, 7, 1
    Ce, Cp = O.random_orthonormal(ne, 21), O.random_orthonormal(npos, 22)
    pe = O.hash_packed_intra(5, ne)
    T.set_species(0, Ce); T.set_species(1, Cp)
    T.set_generator(0, 0, 5)
    win, sym = O.windows_c_intra("PT2", ne, oe)
    assert sym is False
    ref = O.transform_c_intra(Ce, pe, win, sym)
    got = T.transform(0, 0, win, ol.CONV_C, symmetric=sym)
    assert np.abs(dense_quads(*got, ne, ne) - dense_quads(*ref, ne, ne)).max() <= TOL
    assert_lists_match(got[:4], got[4], ref[:4], ref[4])
    rect = O.hash_rect_inter(6, ne, npos)
    T.set_generator(0, 1, 6)
    wine = O.windows_e_inter("PT2", ne, npos, oe, op, ionize_species=("POSITRON",), name_a="E-", name_b="POSITRON")
    r = O.transform_e_inter(Ce, Cp, rect, wine)
    g = T.transform(0, 1, wine, ol.CONV_E)
    Ma, Mb = O.npairs(ne), O.npairs(npos)
    assert np.abs(dense_pairs(*g, Ma, Mb) - dense_pairs(*r, Ma, Mb)).max() <= TOL
    assert_lists_match(g[:2], g[2], r[:2], r[2])


def test_c6h6_shape_rank_k_closed_form(O, T):
    """C6H6/cc-pVDZ shape: N=120 (Cartesian), 21 occupied, MP2 window.  AO tensor = rank-8 separable, uploaded through the
    reference's stack layout (26.4 M unique integrals); every (ia|jb) is compared with the closed form and the MP2
    pair energy with the value computed from it."""
    n, occ, K = 120, 21, 8
    sq, La, _ = O.rankk_square(99, n, K=K)
    Cm = O.random_orthonormal(n, 33)
    T.set_species(0, Cm)
    L = T.L
    T._ck(L.lowdin_it_ao_begin(T.h, 0, 0, 0))
    for p, q, r, s, v in O.list_from_pair_matrix_intra(sq, n, rows_per_block=512):
        T._ck(L.lowdin_it_ao_push_stacks(T.h, p, q, r, s, v, len(v)))
    T._ck(L.lowdin_it_ao_end(T.h))
    win = O.windows_e_intra("MP2", n, occ)          # p,r virtual; q,s occupied
    ij, kl, v = T.transform(0, 0, win, ol.CONV_E)
    virt, oc = np.arange(occ, n), np.arange(occ)
    ref = O.rankk_mo_block(La, Cm, virt, oc)         # [a][i][b][j]
    i1, i2 = np.triu_indices(n)                      # pair id -> (smaller, larger) orbital, 0-based
    a, i = i2[ij - 1] - occ, i1[ij - 1]
    b, j = i2[kl - 1] - occ, i1[kl - 1]
    assert np.abs(v - ref[a, i, b, j]).max() <= TOL
    assert len(v) >= ref.size - np.count_nonzero(np.abs(ref) <= 1.1e-10)
    eps = O.synthetic_eps(occ, n)
    den = -eps[occ:, None, None, None] + eps[None, :occ, None, None] - eps[None, None, occ:, None] + eps[None, None, None, :occ]
    e_ref = np.sum(ref * (2.0 * ref - ref.transpose(2, 1, 0, 3)) / den)
    sums = T.transform_stream(0, 0, win, ol.CONV_E, occ_batch=8, epsA=eps, lam=2.0)
    assert abs(sums[3] - e_ref) <= 1e-9 * max(1.0, abs(e_ref))
    # the same with many chunks of AO-pair rows and another occupied batch: identical up to summation order
    T.set_option(T.OPT_CHUNK_COLS, 3000)
    try:
        sums2 = T.transform_stream(0, 0, win, ol.CONV_E, occ_batch=5, epsA=eps, lam=2.0)
    finally:
        T.set_option(T.OPT_CHUNK_COLS, 0)
    assert sums2[0] == sums[0] and abs(sums2[3] - sums[3]) <= 1e-9 * max(1.0, abs(e_ref))


def test_full_size_properties_n500(T):
    """BASELINE full-size check through size-independent properties (no CPU oracle can run N=500): the streamed sums of
    one occupied batch do not depend on how the AO-pair rows are chunked, on the occupied batching, or on which variant of
    the fused first-quarter kernel produced them."""
    n, occ = 500, 50
    rng = np.random.default_rng(4)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    T.set_species(0, np.asfortranarray(q))
    T.set_generator(0, 0, 77, ol.GEN_FOLD)
    win = [occ + 1, n, 1, occ, occ + 1, n, 1, occ]
    eps = np.concatenate([np.linspace(-2.0, -0.5, occ), np.linspace(0.2, 3.0, n - occ)])
    base = T.transform_stream(0, 0, win, ol.CONV_E, occ_batch=8, first_pass=1, n_passes=1, epsA=eps)
    T.set_option(T.OPT_CHUNK_COLS, 20000)
    T.set_option(T.OPT_Q1_VARIANT, 2)
    try:
        a = T.transform_stream(0, 0, win, ol.CONV_E, occ_batch=8, first_pass=1, n_passes=1, epsA=eps)
        b1 = T.transform_stream(0, 0, win, ol.CONV_E, occ_batch=4, first_pass=2, n_passes=2, epsA=eps)
    finally:
        T.set_option(T.OPT_CHUNK_COLS, 0)
        T.set_option(T.OPT_Q1_VARIANT, T.DEFAULT_Q1_VARIANT)
    for other in (a, b1):
        assert other[0] == base[0]
        assert abs(other[1] - base[1]) <= 1e-9 * max(1.0, abs(base[1]))
        assert abs(other[2] - base[2]) <= 1e-9 * max(1.0, abs(base[2]))
        assert abs(other[3] - base[3]) <= 1e-9 * max(1.0, abs(base[3]))
