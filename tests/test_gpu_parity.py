"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Integer / byte work (slab expansion, scatter, index lists) must be bit-exact.  Floating point:
MO integrals max|delta| <= 1e-10 (BASELINE.json), downstream MP2 energies <= 1e-9.
"""
import numpy as np
import pytest

import openlowdin_b200 as ol
from helpers import assert_lists_match, dense_pairs, dense_quads

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _intra_setup(O, T, n, seed, slot=0):
    packed = O.hash_packed_intra(seed, n)
    Cm = O.random_orthonormal(n, n)
    T.set_species(slot, Cm)
    p, q, r, s, v = O.canonical_list_intra(packed, n)
    T.upload_ao(slot, slot, p, q, r, s, v, stack=1024)
    return packed, Cm


def _inter_setup(O, T, na, nb, seed, sa=0, sb=1):
    rect = O.hash_rect_inter(seed, na, nb)
    Ca, Cb = O.random_orthonormal(na, na), O.random_orthonormal(nb, nb + 1)
    T.set_species(sa, Ca)
    T.set_species(sb, Cb)
    p, q, r, s, v = O.canonical_list_inter(rect, na, nb)
    T.upload_ao(sa, sb, p, q, r, s, v, stack=1024)
    return rect, Ca, Cb


@pytest.mark.parametrize("m,n,k", [(19, 5, 19), (300, 150, 120), (1000, 8, 333), (257, 129, 65), (128, 128, 16),
                                    (7, 3, 2), (513, 81, 1501), (64, 24, 50)])
def test_dmma_gemm_vs_numpy(T, m, n, k):
    rng = np.random.default_rng(m * 1000 + n)
    A, B = rng.uniform(-1, 1, (m, k)), rng.uniform(-1, 1, (n, k))
    got = T.debug_gemm(A, B)
    ref = A @ B.T
    # FP64 accumulation of k products of magnitude <= 1: |delta| <= 1e-15 * k, far inside 1e-10
    assert np.abs(got - ref).max() <= 4e-16 * k + 1e-14


@pytest.mark.parametrize("n", [5, 19, 30])
def test_expand_packed_bit_exact(O, T, n):
    packed, _ = _intra_setup(O, T, n, 11)
    M = O.npairs(n)
    sq = O.packed_to_square(packed, M)
    xy = O.pair_table(n)
    X = T.debug_expand(0, 0, 0, M)
    ref = sq[:, xy]  # [slab][mu][nu]
    assert np.array_equal(X, ref)


def test_expand_rect_and_generator_bit_exact(O, T):
    na, nb = 6, 4
    rect, _, _ = _inter_setup(O, T, na, nb, 3)
    Ma, Mb = O.npairs(na), O.npairs(nb)
    X = T.debug_expand(0, 1, 0, Mb)
    ref = rect.reshape(Mb, Ma)[:, O.pair_table(na)]
    assert np.array_equal(X, ref)
    # generated slabs == uploaded slabs, bit for bit
    T.set_generator(0, 1, 3)
    assert np.array_equal(T.debug_expand(0, 1, 0, Mb), ref)
    n = 9
    packed = O.hash_packed_intra(77, n)
    T.set_species(2, O.random_orthonormal(n, n))
    T.set_generator(2, 2, 77)
    M = O.npairs(n)
    assert np.array_equal(T.debug_expand(2, 2, 0, M), O.packed_to_square(packed, M)[:, O.pair_table(n)])


@pytest.mark.parametrize("n,occ,mode", [(7, 3, "MP2"), (7, 3, "ALLACTIVE"), (19, 5, "MP2"), (19, 5, "PT2"), (19, 5, "MP2-PT2"),
                                         (19, 5, "BOUNDS"), (12, 1, "MP2"), (12, 11, "MP2")])
def test_transform_e_intra(O, T, n, occ, mode):
    packed, Cm = _intra_setup(O, T, n, 100 + n)
    win = O.windows_e_intra(mode, n, occ)
    rij, rkl, rv = O.transform_e_intra(Cm, packed, win)
    ij, kl, v = T.transform(0, 0, win, ol.CONV_E)
    M = O.npairs(n)
    assert np.abs(dense_pairs(ij, kl, v, M, M) - dense_pairs(rij, rkl, rv, M, M)).max() <= TOL
    assert_lists_match((ij, kl), v, (rij, rkl), rv)
    if len(v) == len(rv):  # same order as the reference's loops
        assert np.array_equal(ij, rij) and np.array_equal(kl, rkl)


@pytest.mark.parametrize("n,occ,mode", [(7, 3, "MP2"), (7, 3, "ALL"), (19, 5, "MP2"), (19, 5, "PT2"), (19, 5, "MP2-PT2"),
                                         (19, 5, "ALL"), (19, 5, "BOUNDS")])
def test_transform_c_intra(O, T, n, occ, mode):
    packed, Cm = _intra_setup(O, T, n, 200 + n)
    win, sym = O.windows_c_intra(mode, n, occ)
    ref = O.transform_c_intra(Cm, packed, win, sym)
    got = T.transform(0, 0, win, ol.CONV_C, symmetric=sym)
    assert np.abs(dense_quads(*got, n, n) - dense_quads(*ref, n, n)).max() <= TOL
    assert_lists_match(got[:4], got[4], ref[:4], ref[4])
    if len(got[4]) == len(ref[4]):
        for a, b in zip(got[:4], ref[:4]):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("na,nb,occ_a,occ_b,mode", [(6, 4, 2, 1, "MP2"), (19, 12, 5, 1, "MP2"), (19, 12, 5, 1, "PT2"),
                                                     (10, 19, 1, 5, "MP2-PT2"), (9, 7, 3, 2, "BOUNDS")])
def test_transform_e_inter(O, T, na, nb, occ_a, occ_b, mode):
    rect, Ca, Cb = _inter_setup(O, T, na, nb, 300 + na)
    win = O.windows_e_inter(mode, na, nb, occ_a, occ_b)
    win[5] = min(win[5], nb)  # the MP2-PT2 quirk (E.f90:2283) can exceed the other basis; the reference would read out of bounds
    rij, rkl, rv = O.transform_e_inter(Ca, Cb, rect, win)
    ij, kl, v = T.transform(0, 1, win, ol.CONV_E)
    Ma, Mb = O.npairs(na), O.npairs(nb)
    assert np.abs(dense_pairs(ij, kl, v, Ma, Mb) - dense_pairs(rij, rkl, rv, Ma, Mb)).max() <= TOL
    assert_lists_match((ij, kl), v, (rij, rkl), rv)


@pytest.mark.parametrize("na,nb,occ_a,occ_b,mode", [(6, 4, 2, 1, "MP2"), (19, 12, 5, 1, "ALL"), (19, 12, 5, 1, "PT2"),
                                                     (10, 19, 1, 5, "MP2")])
def test_transform_c_inter(O, T, na, nb, occ_a, occ_b, mode):
    rect, Ca, Cb = _inter_setup(O, T, na, nb, 400 + na)
    win, sym = O.windows_c_inter(mode, na, nb, occ_a, occ_b)
    ref = O.transform_c_inter(Ca, Cb, rect, win, sym)
    got = T.transform(0, 1, win, ol.CONV_C, symmetric=sym)
    assert np.abs(dense_quads(*got, na, nb) - dense_quads(*ref, na, nb)).max() <= TOL
    assert_lists_match(got[:4], got[4], ref[:4], ref[4])


def test_swapped_pair_upload(O, T):
    """C.f90:906-972: the caller passes (B,A); the stacks on file are still (A A|B B)."""
    na, nb = 5, 8
    rect = O.hash_rect_inter(9, na, nb)            # stored rect[rs_B][pq_A]
    lst = O.canonical_list_inter(rect, na, nb)     # file order (A A | B B)
    Ca, Cb = O.random_orthonormal(na, na), O.random_orthonormal(nb, nb)
    # transformer called with first species = B, second = A
    T.set_species(0, Cb); T.set_species(1, Ca)
    T.upload_ao(0, 1, *lst, swapped=True, stack=64)
    ref_rect = O.scatter_inter(*lst, nb, na, swapped=True)
    win, sym = O.windows_c_inter("ALL", nb, na, 2, 1)
    ref = O.transform_c_inter(Cb, Ca, ref_rect, win, sym)
    got = T.transform(0, 1, win, ol.CONV_C, symmetric=sym)
    assert np.abs(dense_quads(*got, nb, na) - dense_quads(*ref, nb, na)).max() <= TOL


def test_transformer_d_dropin(O):
    """lowdin_it_transform_all / _inter_all against the reference's own transformer D (oracle/_ref)."""
    n = 11
    M = O.npairs(n)
    rng = np.random.default_rng(5)
    sq = rng.uniform(-1, 1, (M, M)); sq = sq + sq.T
    eris = O.d_pack_intra(sq)
    Cm = O.random_orthonormal(n, n)
    use_ref = O.ref() is not None
    want = O.transform_d_intra(Cm, eris, use_reference=use_ref)
    got = ol.transform_all(Cm, eris.copy())
    assert np.abs(got - want).max() <= TOL
    na, nb = 8, 5
    er = rng.uniform(-1, 1, O.npairs(na) * O.npairs(nb))
    Ca, Cb = O.random_orthonormal(na, 1), O.random_orthonormal(nb, 2)
    want = O.transform_d_inter(Ca, Cb, er, use_reference=use_ref)
    got = ol.transform_inter_all(Ca, Cb, er.copy())
    assert np.abs(got - want).max() <= TOL


def test_stream_matches_download_and_mp2_energy(O, T):
    n, occ = 19, 5
    packed, Cm = _intra_setup(O, T, n, 555)
    eps = O.synthetic_eps(occ, n)
    win = O.windows_e_intra("MP2", n, occ)
    ij, kl, v = T.transform(0, 0, win, ol.CONV_E)
    e_ref = O.mp2_intra_from_pairs(ij, kl, v, n, occ, eps, lam=2.0)
    rij, rkl, rv = O.transform_e_intra(Cm, packed, win)
    e_orc = O.mp2_intra_from_pairs(rij, rkl, rv, n, occ, eps, lam=2.0)
    assert abs(e_ref - e_orc) <= 1e-9
    for qb in (0, 2, 5):
        sums = T.transform_stream(0, 0, win, ol.CONV_E, occ_batch=qb, epsA=eps, lam=2.0)
        assert sums[0] == len(v)
        assert abs(sums[1] - v.sum()) <= 1e-9
        assert abs(sums[2] - (v * v).sum()) <= 1e-9
        assert abs(sums[3] - e_orc) <= 1e-9
    # generator source gives the same answer as the uploaded list
    T.set_generator(0, 0, 555)
    sums = T.transform_stream(0, 0, win, ol.CONV_E, occ_batch=3, epsA=eps, lam=2.0)
    assert abs(sums[3] - e_orc) <= 1e-9


def test_inter_mp2_energy(O, T):
    na, nb, oa, ob = 19, 12, 5, 1
    rect, Ca, Cb = _inter_setup(O, T, na, nb, 808)
    ea, eb = O.synthetic_eps(oa, na), O.synthetic_eps(ob, nb)
    win = O.windows_e_inter("MP2", na, nb, oa, ob)
    rij, rkl, rv = O.transform_e_inter(Ca, Cb, rect, win)
    e_orc = O.mp2_inter_from_pairs(rij, rkl, rv, na, nb, oa, ob, ea, eb, charge_a=1.0, charge_b=1.0, lam_a=1.0, lam_b=1.0)
    sums = T.transform_stream(0, 1, win, ol.CONV_E, occ_batch=2, epsA=ea, epsB=eb)
    assert abs(sums[3] - e_orc) <= 1e-9


def test_error_paths(T):
    with pytest.raises(ol.LowdinITError):
        T.transform(5, 5, [1, 1, 1, 1, 1, 1, 1, 1], ol.CONV_E)          # species never set
    T.set_species(3, np.eye(4))
    with pytest.raises(ol.LowdinITError):
        T.transform(3, 3, [1, 4, 1, 4, 1, 4, 1, 4], ol.CONV_E)          # no AO integrals uploaded
    T.set_generator(3, 3, 1)
    with pytest.raises(ol.LowdinITError):
        T.transform(3, 3, [1, 5, 1, 4, 1, 4, 1, 4], ol.CONV_E)          # window beyond the orbitals
    with pytest.raises(ol.LowdinITError):
        T.upload_ao(3, 3, np.array([9], np.int32), np.array([1], np.int32), np.array([1], np.int32),
                    np.array([1], np.int32), np.array([1.0]))            # AO index outside the basis


def test_empty_ao_list_gives_no_integrals(O, T):
    n = 6
    T.set_species(0, O.random_orthonormal(n, n))
    e = np.zeros(0, np.int32)
    T.upload_ao(0, 0, e, e, e, e, np.zeros(0))
    ij, kl, v = T.transform(0, 0, [1, n, 1, n, 1, n, 1, n], ol.CONV_E)
    assert len(v) == 0


# ---------------------------------------------------------------------------------------------
# chunked second half (third-quarter accumulation over chunks of AO-pair rows) and the fused
# generation + first-quarter kernel
# ---------------------------------------------------------------------------------------------
@pytest.fixture
def chunked(T):
    """Force many small chunks (the large-N code path) for the duration of one test."""
    def _set(cols):
        T.set_option(T.OPT_CHUNK_COLS, cols)
    yield _set
    T.set_option(T.OPT_CHUNK_COLS, 0)


@pytest.mark.parametrize("cols", [1, 40, 100])
@pytest.mark.parametrize("n,occ,mode", [(19, 5, "MP2"), (19, 5, "MP2-PT2"), (13, 4, "ALLACTIVE")])
def test_chunked_second_half_e_intra(O, T, chunked, n, occ, mode, cols):
    packed, Cm = _intra_setup(O, T, n, 900 + n)
    win = O.windows_e_intra(mode, n, occ)
    rij, rkl, rv = O.transform_e_intra(Cm, packed, win)
    chunked(cols)
    ij, kl, v = T.transform(0, 0, win, ol.CONV_E)
    M = O.npairs(n)
    assert np.abs(dense_pairs(ij, kl, v, M, M) - dense_pairs(rij, rkl, rv, M, M)).max() <= TOL
    assert_lists_match((ij, kl), v, (rij, rkl), rv)
    if len(v) == len(rv):
        assert np.array_equal(ij, rij) and np.array_equal(kl, rkl)


@pytest.mark.parametrize("cols", [1, 30])
def test_chunked_second_half_c_and_inter(O, T, chunked, cols):
    n, occ = 14, 4
    packed, Cm = _intra_setup(O, T, n, 77)
    win, sym = O.windows_c_intra("MP2", n, occ)
    ref = O.transform_c_intra(Cm, packed, win, sym)
    chunked(cols)
    got = T.transform(0, 0, win, ol.CONV_C, symmetric=sym)
    assert np.abs(dense_quads(*got, n, n) - dense_quads(*ref, n, n)).max() <= TOL
    assert_lists_match(got[:4], got[4], ref[:4], ref[4])
    na, nb = 11, 9
    rect, Ca, Cb = _inter_setup(O, T, na, nb, 31)
    win = O.windows_e_inter("MP2", na, nb, 3, 2)
    rij, rkl, rv = O.transform_e_inter(Ca, Cb, rect, win)
    ij, kl, v = T.transform(0, 1, win, ol.CONV_E)
    Ma, Mb = O.npairs(na), O.npairs(nb)
    assert np.abs(dense_pairs(ij, kl, v, Ma, Mb) - dense_pairs(rij, rkl, rv, Ma, Mb)).max() <= TOL


@pytest.mark.parametrize("n,win", [(19, [6, 19, 1, 5, 6, 19, 1, 5]), (23, [1, 23, 1, 23, 1, 23, 1, 23]), (37, [12, 37, 1, 11, 12, 37, 1, 11]),
                                   (70, [1, 70, 1, 1, 1, 70, 1, 70]), (70, [66, 70, 1, 65, 1, 3, 1, 2])])
def test_generated_source_fused_first_quarter(O, T, n, win):
    """Slabs generated inside the first-quarter kernel (no dense slab in memory) == uploaded list + expansion."""
    seed = 4242 + n
    packed = O.hash_packed_intra(seed, n)
    Cm = O.random_orthonormal(n, n)
    T.set_species(0, Cm)
    T.set_generator(0, 0, seed)
    rij, rkl, rv = O.transform_e_intra(Cm, packed, win)
    ij, kl, v = T.transform(0, 0, win, ol.CONV_E)
    M = O.npairs(n)
    assert np.abs(dense_pairs(ij, kl, v, M, M) - dense_pairs(rij, rkl, rv, M, M)).max() <= TOL
    assert_lists_match((ij, kl), v, (rij, rkl), rv)


def test_generated_source_inter_and_chunked_stream(O, T, chunked):
    na, nb, oa, ob = 21, 16, 5, 2
    seed = 99
    rect = O.hash_rect_inter(seed, na, nb)
    Ca, Cb = O.random_orthonormal(na, 3), O.random_orthonormal(nb, 4)
    T.set_species(0, Ca); T.set_species(1, Cb)
    T.set_generator(0, 1, seed)
    ea, eb = O.synthetic_eps(oa, na), O.synthetic_eps(ob, nb)
    win = O.windows_e_inter("MP2", na, nb, oa, ob)
    rij, rkl, rv = O.transform_e_inter(Ca, Cb, rect, win)
    e_orc = O.mp2_inter_from_pairs(rij, rkl, rv, na, nb, oa, ob, ea, eb, charge_a=1.0, charge_b=1.0, lam_a=1.0, lam_b=1.0)
    for cols, qb in ((0, 0), (25, 2), (1, 3)):
        chunked(cols)
        sums = T.transform_stream(0, 1, win, ol.CONV_E, occ_batch=qb, epsA=ea, epsB=eb)
        assert sums[0] == len(rv)
        assert abs(sums[3] - e_orc) <= 1e-9


def test_chunked_stream_mp2_energy_intra(O, T, chunked):
    n, occ = 24, 6
    seed = 31337
    packed = O.hash_packed_intra(seed, n)
    Cm = O.random_orthonormal(n, n)
    eps = O.synthetic_eps(occ, n)
    T.set_species(0, Cm)
    T.set_generator(0, 0, seed)
    win = O.windows_e_intra("MP2", n, occ)
    rij, rkl, rv = O.transform_e_intra(Cm, packed, win)
    e_orc = O.mp2_intra_from_pairs(rij, rkl, rv, n, occ, eps, lam=2.0)
    for cols, qb in ((0, 0), (60, 4), (1, 1), (200, 6)):
        chunked(cols)
        sums = T.transform_stream(0, 0, win, ol.CONV_E, occ_batch=qb, epsA=eps, lam=2.0)
        assert sums[0] == len(rv)
        assert abs(sums[2] - (rv * rv).sum()) <= 1e-9
        assert abs(sums[3] - e_orc) <= 1e-9


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("kind", [1, 2])
def test_generator_kinds_and_fused_kernel_variants(O, T, kind, variant):
    """Both synthetic generators (H: splitmix64, F: mul-fold-mul) are bit-identical on host and device, through the
    expansion kernel and through both variants of the fused generation + first-quarter kernel."""
    n = 21
    seed = 1000 + kind
    packed = O.hash_packed_intra(seed, n, kind)
    Cm = O.random_orthonormal(n, n)
    T.set_species(0, Cm)
    T.set_generator(0, 0, seed, kind)
    M = O.npairs(n)
    assert np.array_equal(T.debug_expand(0, 0, 0, M), O.packed_to_square(packed, M)[:, O.pair_table(n)])
    win = O.windows_e_intra("MP2", n, 6)
    rij, rkl, rv = O.transform_e_intra(Cm, packed, win)
    T.set_option(T.OPT_Q1_VARIANT, variant)
    try:
        ij, kl, v = T.transform(0, 0, win, ol.CONV_E)
    finally:
        T.set_option(T.OPT_Q1_VARIANT, T.DEFAULT_Q1_VARIANT)
    assert np.abs(dense_pairs(ij, kl, v, M, M) - dense_pairs(rij, rkl, rv, M, M)).max() <= TOL
    assert_lists_match((ij, kl), v, (rij, rkl), rv)
    na, nb = 8, 6
    rect = O.hash_rect_inter(seed, na, nb, kind)
    T.set_species(1, O.random_orthonormal(na, 1)); T.set_species(2, O.random_orthonormal(nb, 2))
    T.set_generator(1, 2, seed, kind)
    assert np.array_equal(T.debug_expand(1, 2, 0, O.npairs(nb)), rect.reshape(O.npairs(nb), O.npairs(na))[:, O.pair_table(na)])


@pytest.mark.gpu
@pytest.mark.parametrize("conv", [ol.CONV_C, ol.CONV_E])
def test_empty_windows_on_a_fresh_handle(O, conv):
    """No window pair at all (one-function species under MP2), and a first pair with slots but an empty second window:
    a valid transform with zero integrals, downloadable on a handle that has never produced a result."""
    for n, win in ((1, [2, 1, 1, 1, 2, 1, 1, 1]), (4, [3, 4, 1, 2, 5, 4, 1, 2])):
        T = ol.Transformer(0)
        try:
            T.set_species(0, O.random_orthonormal(n, n) if n > 1 else np.ones((1, 1)))
            T.set_generator(0, 0, 3)
            out = T.transform(0, 0, win, conv)
            assert all(len(x) == 0 for x in out)
            assert np.all(T.transform_stream(0, 0, win, conv) == 0.0)
        finally:
            T.close()


@pytest.mark.gpu
@pytest.mark.parametrize("qb", [0, 2])
def test_stream_sink_delivers_every_mo_integral(O, T, qb):
    """lowdin_it_transform_stream_sink: the dense blocks handed to the host (pinned ring, copy overlapped with the next group's
    fourth quarter) hold exactly the integrals lowdin_it_transform + download give, for intra and inter MP2 windows."""
    n, occ = 19, 5
    packed, Cm = _intra_setup(O, T, n, 77)
    M = O.npairs(n)
    pid = lambda x, y, nn: min(x, y) * nn - min(x, y) * (min(x, y) - 1) // 2 + abs(x - y)  # noqa: E731  0-based pair id of 0-based (x, y)
    cases = [((0, 0), O.windows_e_intra("MP2", n, occ), n, n, M, M)]
    na, nb = 11, 8
    _inter_setup(O, T, na, nb, 5, 1, 2)
    cases.append(((1, 2), O.windows_e_inter("MP2", na, nb, 3, 2), na, nb, O.npairs(na), O.npairs(nb)))
    for (a, b), win, n1, n2, M1, M2 in cases:
        ij, kl, v = T.transform(a, b, win, ol.CONV_E)
        ref = dense_pairs(ij, kl, v, M1, M2)
        got = np.zeros((M1, M2))
        nblocks = [0]

        def sink(sa, sb, vals, blk):
            nblocks[0] += 1
            for t in range(len(sa)):
                row = pid(sa[t] - 1, sb[t] - 1, n1)
                for ks in range(blk.n_second):
                    for kf in range(blk.n_first):
                        r, s = (blk.orb_second0 + ks, blk.orb_first0 + kf) if blk.second_is_conv_first else (blk.orb_first0 + kf, blk.orb_second0 + ks)
                        x = vals[t, ks, kf]
                        if abs(x) > 1e-10:
                            got[row, pid(r - 1, s - 1, n2)] = x
        T.set_option(T.OPT_SINK_BLOCK_BYTES, 1 << 12)      # one first-contracted index per block: several blocks, both ring slots reused
        try:
            sums = T.transform_stream_sink(a, b, win, ol.CONV_E, sink, occ_batch=qb)
        finally:
            T.set_option(T.OPT_SINK_BLOCK_BYTES, 256 << 20)
        assert nblocks[0] >= (3 if a == b else 1) and sums[0] == len(v)     # the inter case is smaller than one 4 KiB block
        assert np.abs(got - ref).max() <= 1e-12
