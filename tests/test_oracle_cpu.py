"""CPU tests of the oracle itself (no GPU): the C restatement against the golden vectors produced by the
reference's own transformer D (tests/golden/, made by oracle/make_golden.py), against a dense numpy.einsum
third opinion, and the restated host logic (index maps, window tables, MP2 formula).
"""
import glob
import os

import numpy as np
import pytest

from helpers import assert_lists_match, dense_pairs, dense_quads

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-10


def _d_to_xy_perm(O, n):
    """perm[d_pair_id] = xy pair id (both 0-based): D numbers pairs i(i+1)/2+j (i>=j), C/E row-wise upper."""
    d, xy = O.d_pair_table(n), O.pair_table(n)
    perm = np.zeros(O.npairs(n), dtype=np.int64)
    perm[d[np.tril_indices(n)]] = xy[np.tril_indices(n)]
    return perm


def _d_unpack_intra(eris, M):
    sq = np.zeros((M, M))
    il = np.tril_indices(M)
    sq[il] = eris
    sq.T[il] = eris
    return sq


def golden(kind):
    files = sorted(glob.glob(os.path.join(GOLD, f"d_{kind}_*.npz")))
    assert files, "tests/golden is empty: run python oracle/make_golden.py in the build container"
    return files


# ---------------------------------------------------------------------------------------------
# golden vectors from the reference's own code
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", golden("intra"), ids=os.path.basename)
def test_restated_d_intra_matches_reference_golden(O, path):
    g = np.load(path)
    got = O.transform_d_intra(np.asfortranarray(g["C"]), g["eris_in"], use_reference=False)
    assert np.abs(got - g["eris_out"]).max() <= 1e-12


@pytest.mark.parametrize("path", golden("inter"), ids=os.path.basename)
def test_restated_d_inter_matches_reference_golden(O, path):
    g = np.load(path)
    got = O.transform_d_inter(np.asfortranarray(g["Ca"]), np.asfortranarray(g["Cb"]), g["eris_in"], use_reference=False)
    assert np.abs(got - g["eris_out"]).max() <= 1e-12


@pytest.mark.parametrize("path", golden("intra"), ids=os.path.basename)
def test_restated_e_and_c_full_window_match_reference_golden(O, path):
    """Transformers E and C (restated) on the full window reproduce what the reference's D computed."""
    g = np.load(path)
    Cm = np.asfortranarray(g["C"])
    n = Cm.shape[0]
    M = O.npairs(n)
    perm = _d_to_xy_perm(O, n)
    sq_in = np.zeros((M, M))
    sq_in[np.ix_(perm, perm)] = _d_unpack_intra(g["eris_in"], M)
    sq_ref = np.zeros((M, M))
    sq_ref[np.ix_(perm, perm)] = _d_unpack_intra(g["eris_out"], M)
    packed = O.square_to_packed(sq_in)
    win = [1, n, 1, n, 1, n, 1, n]
    ij, kl, v = O.transform_e_intra(Cm, packed, win)
    got = dense_pairs(ij, kl, v, M, M)
    assert np.abs(got - sq_ref).max() <= TOL          # dropped (|v|<=1e-10) entries count as 0
    p, q, r, s, v = O.transform_c_intra(Cm, packed, win, True)
    xy = O.pair_table(n)
    ref4 = sq_ref[xy[:, :, None, None], xy[None, None, :, :]]
    assert np.abs(v - ref4[p - 1, q - 1, r - 1, s - 1]).max() <= TOL
    # C's symmetric skip rules (C.f90:380,394,409): q>=p, r>=p, s>=r only
    assert np.all(q >= p) and np.all(r >= p) and np.all(s >= r)


@pytest.mark.parametrize("path", golden("inter"), ids=os.path.basename)
def test_restated_e_and_c_inter_match_reference_golden(O, path):
    g = np.load(path)
    Ca, Cb = np.asfortranarray(g["Ca"]), np.asfortranarray(g["Cb"])
    na, nb = Ca.shape[0], Cb.shape[0]
    Ma, Mb = O.npairs(na), O.npairs(nb)
    pa, pb = _d_to_xy_perm(O, na), _d_to_xy_perm(O, nb)
    sq_in = np.zeros((Ma, Mb))
    sq_in[np.ix_(pa, pb)] = g["eris_in"].reshape(Ma, Mb)      # D: ij*Mb+kl
    sq_ref = np.zeros((Ma, Mb))
    sq_ref[np.ix_(pa, pb)] = g["eris_out"].reshape(Ma, Mb)
    rect = np.ascontiguousarray(sq_in.T).ravel()               # C/E AO storage (rs-1)*M_a+pq
    win = [1, na, 1, na, 1, nb, 1, nb]
    ij, kl, v = O.transform_e_inter(Ca, Cb, rect, win)
    assert np.abs(dense_pairs(ij, kl, v, Ma, Mb) - sq_ref).max() <= TOL
    p, q, r, s, v = O.transform_c_inter(Ca, Cb, rect, win, True)
    xa, xb = O.pair_table(na), O.pair_table(nb)
    ref4 = sq_ref[xa[:, :, None, None], xb[None, None, :, :]]
    assert np.abs(v - ref4[p - 1, q - 1, r - 1, s - 1]).max() <= TOL


def test_reference_library_agrees_when_present(O):
    """In the build container oracle/_ref/libref_d.so exists: the goldens are reproducible."""
    if O.ref() is None:
        pytest.skip("oracle/_ref/libref_d.so not present on this box")
    g = np.load(golden("intra")[0])
    got = O.transform_d_intra(np.asfortranarray(g["C"]), g["eris_in"], use_reference=True)
    assert np.array_equal(got, g["eris_out"]) or np.abs(got - g["eris_out"]).max() <= 1e-14


# ---------------------------------------------------------------------------------------------
# third opinion: dense einsum
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,occ,mode", [(6, 2, "MP2"), (9, 3, "PT2"), (9, 3, "MP2-PT2"), (8, 3, "BOUNDS")])
def test_windowed_e_and_c_vs_einsum(O, n, occ, mode):
    packed = O.hash_packed_intra(n, n)
    Cm = O.random_orthonormal(n, n)
    M = O.npairs(n)
    xy = O.pair_table(n)
    mo4 = O.einsum_transform(O.dense4_from_square(O.packed_to_square(packed, M), xy), np.asarray(Cm))
    w = O.windows_e_intra(mode, n, occ)
    ij, kl, v = O.transform_e_intra(Cm, packed, w)
    x1, x2 = np.triu_indices(n)
    # E emits pair ids of (j,i) with j<=i: value must equal mo4[i,j,k,l]
    i_, j_, k_, l_ = x2[ij - 1], x1[ij - 1], x2[kl - 1], x1[kl - 1]
    # half-transformed drops (E.f90:1113) perturb by at most n_dropped*1e-10*|C|^2; here nothing is near the edge
    assert np.abs(v - mo4[i_, j_, k_, l_]).max() <= 1e-9
    wc, sym = O.windows_c_intra(mode, n, occ)
    p, q, r, s, vc = O.transform_c_intra(Cm, packed, wc, sym)
    assert np.abs(vc - mo4[p - 1, q - 1, r - 1, s - 1]).max() <= TOL
    # every kept candidate of the window is present unless |x|<=1e-10
    cnt = 0
    for P in range(wc[0], wc[1] + 1):
        for Q in range(wc[2], wc[3] + 1):
            if sym and Q < P:
                continue
            for R in range(wc[4], wc[5] + 1):
                if sym and R < P:
                    continue
                for S in range(wc[6], wc[7] + 1):
                    if sym and S < R:
                        continue
                    cnt += abs(mo4[P - 1, Q - 1, R - 1, S - 1]) > 1e-10
    assert cnt == len(vc)


def test_inter_e_vs_einsum(O):
    na, nb = 7, 5
    rect = O.hash_rect_inter(3, na, nb)
    Ca, Cb = O.random_orthonormal(na, 1), O.random_orthonormal(nb, 2)
    Ma, Mb = O.npairs(na), O.npairs(nb)
    ao4 = O.dense4_from_square(rect.reshape(Mb, Ma).T, O.pair_table(na), O.pair_table(nb))
    mo4 = O.einsum_transform(ao4, np.asarray(Ca), np.asarray(Cb))
    w = O.windows_e_inter("MP2", na, nb, 3, 1)
    ij, kl, v = O.transform_e_inter(Ca, Cb, rect, w)
    a1, a2 = np.triu_indices(na)
    b1, b2 = np.triu_indices(nb)
    assert np.abs(v - mo4[a2[ij - 1], a1[ij - 1], b2[kl - 1], b1[kl - 1]]).max() <= 1e-9
    assert len(v) == (na - 3) * 3 * (nb - 1) * 1


# ---------------------------------------------------------------------------------------------
# index maps, AO list loader
# ---------------------------------------------------------------------------------------------
def test_pair_id_is_tensorR2ToVectorB(O):
    n = 9
    k = 0
    for i in range(1, n + 1):
        for j in range(i, n + 1):
            k += 1
            assert O.pair_id(i, j, n) == k == O.pair_id(j, i, n)     # IndexMap.f90:259-263
    M = O.npairs(n)
    L = O.lib()
    off = 0
    for pq in range(1, M + 1):                                        # ioff recurrence, C.f90:223-226
        assert L.orc_ioff(pq, M) == off
        off += M - pq                                                 # ioff(pq+1) = ioff(pq) + M - (pq+1) + 1
        assert L.orc_packed_index(pq, pq, M) == L.orc_ioff(pq, M) + pq


def test_scatter_roundtrip_and_terminator(O):
    n = 6
    packed = O.hash_packed_intra(99, n)
    lst = O.canonical_list_intra(packed, n)
    assert np.array_equal(O.scatter_intra(*lst, n), packed)
    # canonical order of lowdin-ints: i>=j, k>=l, (ij)>=(kl) (Iterators.cpp:45-77)
    p, q, r, s, _ = lst
    assert np.all(p >= q) and np.all(r >= s) and np.all((p > r) | ((p == r) & (q >= s)))
    na, nb = 5, 4
    rect = O.hash_rect_inter(5, na, nb)
    li = O.canonical_list_inter(rect, na, nb)
    assert np.array_equal(O.scatter_inter(*li, na, nb), rect)
    # swapped-pair branch (C.f90:906-972): same file, transposed destination
    sw = O.scatter_inter(*li, nb, na, swapped=True)
    assert np.array_equal(sw.reshape(O.npairs(na), O.npairs(nb)), rect.reshape(O.npairs(nb), O.npairs(na)).T)


def test_hash_generator_is_a_pure_function_of_the_index(O):
    n = 5
    M = O.npairs(n)
    packed = O.hash_packed_intra(1234, n)
    L = O.lib()
    for lo in range(M):
        for hi in range(lo, M):
            assert packed[lo * M - lo * (lo + 1) // 2 + hi] == L.orc_hash_value(1234, hi * M + lo)
    assert np.all(np.abs(packed) <= 1.0)


# ---------------------------------------------------------------------------------------------
# host logic: partialTransform and window tables (literal expectations read off the reference)
# ---------------------------------------------------------------------------------------------
def test_partial_transform_choice(O):
    assert O.partial_transform(mp_correction=2) == "MP2"                       # IntegralTransformation.f90:106-110
    assert O.partial_transform(pt_order=2) == "PT2"
    assert O.partial_transform(mp_correction=2, pt_order=2) == "MP2-PT2"
    assert O.partial_transform(ci_level="CISD") == "ALL"
    assert O.partial_transform(en_correction=2) == "BOUNDS"
    assert O.partial_transform(pt_order=3) == "BOUNDS"


def test_window_tables(O):
    n, occ = 19, 5
    assert O.windows_c_intra("MP2", n, occ) == ([1, 5, 6, 19, 1, 5, 6, 19], True)        # C.f90:1489-1501
    assert O.windows_c_intra("ALL", n, occ) == ([1, 19] * 4, True)
    assert O.windows_c_intra("PT2", n, occ) == ([5, 6, 1, 19, 1, 5, 6, 19], False)       # C.f90:1509-1520
    assert O.windows_c_intra("PT2", n, occ, ionize_mo=3) == ([3, 3, 1, 19, 1, 5, 6, 19], False)
    assert O.windows_c_intra("PT2", n, occ, ionize_mo=3, pt_transition_operator=True) == ([3, 3, 1, 19, 1, 5, 1, 19], False)
    assert O.windows_c_intra("MP2", n, occ, core=1, active=15) == ([2, 5, 6, 15, 2, 5, 6, 15], True)
    assert O.windows_e_intra("MP2", n, occ) == [6, 19, 1, 5, 6, 19, 1, 5]                 # E.f90:1938-1949 (roles swapped)
    assert O.windows_e_intra("PT2", n, occ) == [1, 19, 1, 6, 6, 19, 1, 5]
    assert O.windows_e_intra("ALL", n, occ) == [1, 19] * 4                                 # no ALL case in E: default window
    assert O.windows_c_inter("MP2", 19, 50, 5, 1) == ([1, 5, 6, 19, 1, 1, 2, 50], True)  # C.f90:1672-1684
    assert O.windows_e_inter("MP2", 19, 50, 5, 1) == [6, 19, 1, 5, 2, 50, 1, 1]
    w, sym = O.windows_c_inter("PT2", 19, 50, 5, 1, ionize_species=("A",), name_a="A", name_b="B")
    assert (w, sym) == ([1, 6, 1, 19, 1, 1, 2, 50], False)


# ---------------------------------------------------------------------------------------------
# downstream energy
# ---------------------------------------------------------------------------------------------
def test_mp2_formula_is_closed_shell_mp2(O):
    """MPFunctions.f90:450-512 with lambda=2 equals sum (ia|jb)[2(ia|jb)-(ib|ja)]/(ei+ej-ea-eb)."""
    n, occ = 8, 3
    packed = O.hash_packed_intra(42, n)
    Cm = O.random_orthonormal(n, n)
    eps = O.synthetic_eps(occ, n)
    M = O.npairs(n)
    mo4 = O.einsum_transform(O.dense4_from_square(O.packed_to_square(packed, M), O.pair_table(n)), np.asarray(Cm))
    o, v = slice(0, occ), slice(occ, n)
    iajb = mo4[o, v, o, v]
    den = eps[o, None, None, None] - eps[None, v, None, None] + eps[None, None, o, None] - eps[None, None, None, v]
    e_ref = np.sum(iajb * (2.0 * iajb - iajb.transpose(0, 3, 2, 1)) / den)
    for conv in ("E", "C"):
        if conv == "E":
            ij, kl, val = O.transform_e_intra(Cm, packed, O.windows_e_intra("MP2", n, occ))
            e = O.mp2_intra_from_pairs(ij, kl, val, n, occ, eps, lam=2.0)
        else:
            w, sym = O.windows_c_intra("MP2", n, occ)
            e = O.mp2_intra_from_quads(*O.transform_c_intra(Cm, packed, w, sym), n, occ, eps, lam=2.0)
        assert abs(e - e_ref) <= 1e-9


def test_mp2_inter_formula(O):
    na, nb, oa, ob = 7, 5, 3, 1
    rect = O.hash_rect_inter(8, na, nb)
    Ca, Cb = O.random_orthonormal(na, 1), O.random_orthonormal(nb, 2)
    ea, eb = O.synthetic_eps(oa, na), O.synthetic_eps(ob, nb)
    Ma, Mb = O.npairs(na), O.npairs(nb)
    mo4 = O.einsum_transform(O.dense4_from_square(rect.reshape(Mb, Ma).T, O.pair_table(na), O.pair_table(nb)),
                             np.asarray(Ca), np.asarray(Cb))
    x = mo4[:oa, oa:, :ob, ob:]
    den = ea[:oa, None, None, None] - ea[None, oa:, None, None] + eb[None, None, :ob, None] - eb[None, None, None, ob:]
    e_ref = 2.0 * 1.0 * np.sum(x * x / den)           # lambda_a * lambda_b * sum (MPFunctions.f90:750-765, :878-879)
    ij, kl, v = O.transform_e_inter(Ca, Cb, rect, O.windows_e_inter("MP2", na, nb, oa, ob))
    e = O.mp2_inter_from_pairs(ij, kl, v, na, nb, oa, ob, ea, eb, charge_a=-1.0, charge_b=1.0, lam_a=2.0, lam_b=1.0)
    assert abs(e - e_ref) <= 1e-9


def test_list_helper_flags_edge_cases():
    k = (np.array([1, 2]), np.array([1, 1]))
    assert_lists_match(k, np.array([1.0, 9e-11]), (np.array([1]), np.array([1])), np.array([1.0]))
    with pytest.raises(AssertionError):
        assert_lists_match(k, np.array([1.0, 0.5]), (np.array([1]), np.array([1])), np.array([1.0]))
    assert dense_quads(np.array([1]), np.array([2]), np.array([1]), np.array([1]), np.array([3.0]), 2, 1)[0, 1, 0, 0] == 3.0
