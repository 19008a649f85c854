"""Row f4 on the GPU: the AO two-particle integrals evaluated on the device (lowdin_it_set_basis / lowdin_it_ao_compute,
it_eri.cuh; replaces Libint2Iface.cpp:219-416 and :930-1110) against the CPU oracle (oracle/eri_oracle.c, an independent
algorithm) and the textbook H2 / STO-3G values; then the whole chain basis -> AO integrals -> MO integrals with no `.ints` stream."""
import numpy as np
import pytest

import openlowdin_b200 as ol
from openlowdin_b200 import capi
from eri_cases import SZABO_H2, h2_sto3g, nbf, nuclear_like, water_like
from helpers import dense_pairs

pytestmark = pytest.mark.gpu


def _cmp(got, ref):
    both = (got != 0) & (ref != 0)
    assert np.abs(got - ref)[both].max() < 2e-12, np.abs(got - ref)[both].max()
    edge = (got != 0) != (ref != 0)       # dropped on one side only: raw value at the reference's 1e-10 filter
    assert np.abs(got - ref)[edge].max(initial=0.0) < 1e-8


def test_h2_sto3g_textbook_values(O, T):
    sh = h2_sto3g(O)
    T.set_species(0, np.eye(2))
    assert T.set_basis(0, sh) == 2
    T.compute_ao(0, 0)
    packed = T.download_ao(0, 0)
    M, pid = 3, {(0, 0): 0, (0, 1): 1, (1, 0): 1, (1, 1): 2}
    for (i, j, k, l), ref in SZABO_H2.items():
        lo, hi = sorted((pid[(i, j)], pid[(k, l)]))
        assert abs(packed[lo * M - lo * (lo + 1) // 2 + hi] - ref) < 1e-4
    assert np.allclose(T.basis_norma(0), O.eri_norma(sh), rtol=1e-14)
    _cmp(packed, O.eri_packed_intra(sh))


@pytest.mark.parametrize("with_f", [False, True])
def test_spdf_basis_matches_oracle(O, T, with_f):
    sh = water_like(with_f)
    n = nbf(sh)
    T.set_species(0, O.random_orthonormal(n, 5))
    T.set_basis(0, sh)
    T.compute_ao(0, 0)
    ref = O.eri_packed_intra(sh)
    _cmp(T.download_ao(0, 0), ref)
    assert np.abs(ref).max() > 1.0


def test_inter_species_matches_oracle_and_transforms(O, T):
    """Electron-like basis x nucleus-like basis (compute_coupling_disk), then the inter-species MP2 transform of the computed tensor."""
    sa, sb = water_like(), nuclear_like()
    na, nb = nbf(sa), nbf(sb)
    Ca, Cb = O.random_orthonormal(na, 7), O.random_orthonormal(nb, 8)
    T.set_species(1, Ca); T.set_species(2, Cb)
    T.set_basis(1, sa); T.set_basis(2, sb)
    T.compute_ao(1, 2)
    rect = O.eri_rect_inter(sa, sb)
    _cmp(T.download_ao(1, 2), rect)
    win = O.windows_e_inter("MP2", na, nb, 5, 1)
    ij, kl, v = T.transform(1, 2, win, ol.CONV_E)
    rij, rkl, rv = O.transform_e_inter(Ca, Cb, rect, win)
    Ma, Mb = O.npairs(na), O.npairs(nb)
    assert np.abs(dense_pairs(ij, kl, v, Ma, Mb) - dense_pairs(rij, rkl, rv, Ma, Mb)).max() <= 1e-10


@pytest.mark.parametrize("conv", ["E", "C"])
def test_basis_to_mo_integrals_without_ints_stream(O, T, conv):
    """basis -> AO integrals (device) -> MO integrals: equal to the oracle's transform of the oracle's AO integrals, and to the
    path through a canonical list uploaded in the .ints stack layout (what the reference's two programs exchange)."""
    sh = water_like()
    n, occ = nbf(sh), 5
    Cm = O.random_orthonormal(n, 11)
    eps = O.synthetic_eps(occ, n)
    T.set_species(0, Cm)
    T.set_basis(0, sh)
    T.compute_ao(0, 0)
    packed = O.eri_packed_intra(sh)
    M = O.npairs(n)
    if conv == "E":
        win = O.windows_e_intra("MP2", n, occ)
        ij, kl, v = T.transform(0, 0, win, ol.CONV_E)
        rij, rkl, rv = O.transform_e_intra(Cm, packed, win)
        assert np.abs(dense_pairs(ij, kl, v, M, M) - dense_pairs(rij, rkl, rv, M, M)).max() <= 1e-10
        e2 = T.transform_stream(0, 0, win, ol.CONV_E, epsA=eps)[3]
        assert abs(e2 - O.mp2_intra_from_pairs(rij, rkl, rv, n, occ, eps)) <= 1e-9
        # the same integrals as a stored list in the .ints layout
        T.upload_ao(0, 0, *O.canonical_list_intra(T.download_ao(0, 0), n), stack=512)
        ij2, kl2, v2 = T.transform(0, 0, win, ol.CONV_E)
        assert np.array_equal(ij, ij2) and np.array_equal(kl, kl2) and np.abs(v - v2).max() <= 1e-13
    else:
        from helpers import dense_quads
        win, sym = O.windows_c_intra("MP2", n, occ)
        p, q, r, s, v = T.transform(0, 0, win, ol.CONV_C, symmetric=sym)
        rp, rq, rr, rs, rv = O.transform_c_intra(Cm, packed, win, sym)
        assert np.abs(dense_quads(p, q, r, s, v, n, n) - dense_quads(rp, rq, rr, rs, rv, n, n)).max() <= 1e-10


def test_computed_tensor_on_a_group_of_ranks(O):
    """On a communicator every rank evaluates the rows of the AO tensor it owns; the collective transform gives the 1-rank result."""
    sh = water_like()
    n, occ = nbf(sh), 5
    Cm = O.random_orthonormal(n, 13)
    eps = O.synthetic_eps(occ, n)
    win = O.windows_e_intra("MP2", n, occ)
    rij, rkl, rv = O.transform_e_intra(Cm, O.eri_packed_intra(sh), win)
    want = np.array([len(rv), rv.sum(), (rv * rv).sum(), O.mp2_intra_from_pairs(rij, rkl, rv, n, occ, eps, lam=2.0)])
    Ts = [ol.Transformer(0) for _ in range(3)]
    capi.local_group(Ts)
    try:
        def work(r, T):
            T.set_option(T.OPT_SLAB_BLOCK_LOG, 2)
            T.set_species(0, Cm)
            T.set_basis(0, sh)
            T.compute_ao(0, 0)
            T.set_option(T.OPT_CHUNK_COLS, 70)
            return T.transform_stream(0, 0, win, ol.CONV_E, epsA=eps, lam=2.0)
        got = np.sum(capi.run_ranks(Ts, work), axis=0)
        assert abs(got[0] - want[0]) <= 2 and np.abs(got[1:] - want[1:]).max() <= 1e-9, (got, want)
    finally:
        for t in Ts:
            t.close()


def test_basis_errors(T):
    T.set_species(3, np.eye(3))
    with pytest.raises(ol.LowdinITError, match="Cartesian functions"):
        T.set_basis(3, [(0, (0, 0, 0), [1.0], [1.0])])                 # 1 function for a species of 3
    with pytest.raises(ol.LowdinITError, match="l <= 3"):
        T.set_basis(3, [(4, (0, 0, 0), [1.0], [1.0])])
    with pytest.raises(ol.LowdinITError, match="basis not set"):
        T.compute_ao(3, 3)


def test_computed_ints_files_feed_the_file_to_file_transform(O, T, tmp_path):
    """The integrals program's drop-in: .ints stream files written from device-evaluated integrals (reference index conventions:
    p >= q, r >= s, (p,q) >= (r,s), 1-based; terminator), then the ordinary file-to-file transformer call on those files."""
    sh = water_like()
    n, occ, S = nbf(sh), 5, 300
    Cm = O.random_orthonormal(n, 17)
    ctl = capi.host_control("E", "MP2", stack=S, nfiles=3, scratch_dir=str(tmp_path))
    sp = capi.host_species("E-", 1, n, occ, coeff=Cm)
    T.set_species(0, Cm)
    T.set_basis(0, sh)
    cnt = capi.host_write_computed_ints(T, ctl, sp, None, 0)
    packed = O.eri_packed_intra(sh)
    assert abs(cnt - np.count_nonzero(packed)) <= 4          # entries at the 1e-10 raw-value filter may differ
    got = np.zeros_like(packed)
    M = O.npairs(n)
    total = 0
    for t in range(3):
        (p, q, r, s, v), nread = capi.host_read_ints_file(str(tmp_path / f"{t}E-.ints"), S, packed.size)
        assert nread == len(v)
        total += len(v)
        assert np.all(p >= q) and np.all(r >= s) and np.all((p > r) | ((p == r) & (q >= s)))
        p, q, r, s = (x.astype(np.int64) for x in (p, q, r, s))
        pid = lambda lo, hi: lo * n - lo * (lo - 1) // 2 + (hi - lo)     # noqa: E731  0-based row-wise upper pair id, lo <= hi
        lo_hi = np.sort(np.stack([pid(q - 1, p - 1), pid(s - 1, r - 1)]), axis=0)
        got[lo_hi[0] * M - lo_hi[0] * (lo_hi[0] + 1) // 2 + lo_hi[1]] = v
    assert total == cnt
    both = (got != 0) & (packed != 0)
    assert np.abs(got - packed)[both].max() < 2e-12
    # the reference host's transformer call on those files
    nmo = capi.host_transform_one_species(T, ctl, sp)
    rij, rkl, rv = O.transform_e_intra(Cm, packed, O.windows_e_intra("MP2", n, occ))
    assert abs(nmo - len(rv)) <= 2


def test_h2_minimal_basis_chain_reproduces_the_textbook(O, T):
    """The whole device chain on a real molecule whose every number is published (Szabo & Ostlund, minimal-basis H2, R = 1.4):
    basis -> AO integrals (it_eri.cuh) -> MO integrals J11, J12, J22, K12 (transformer-C conventions, full window) -> (ia|jb) of
    the MP2 window (transformer-E conventions) -> MP2 correlation energy -0.0132 Eh on the device."""
    from eri_cases import H2_MO, h2_mo_coefficients, sto3g_overlap
    sh = h2_sto3g(O)
    Cm = h2_mo_coefficients(sto3g_overlap(sh[0], 1.4))
    T.set_species(0, Cm)
    T.set_basis(0, sh)
    T.compute_ao(0, 0)
    win, sym = O.windows_c_intra("ALL", 2, 1)
    p, q, r, s, v = T.transform(0, 0, win, ol.CONV_C, symmetric=sym)
    mo = {(a, b, c, d): x for a, b, c, d, x in zip(p, q, r, s, v)}
    for key, name in (((1, 1, 1, 1), "J11"), ((1, 1, 2, 2), "J12"), ((2, 2, 2, 2), "J22"), ((1, 2, 1, 2), "K12")):
        assert abs(mo[key] - H2_MO[name]) < 1e-4, (key, mo[key])
    assert (1, 1, 1, 2) not in mo and (1, 2, 2, 2) not in mo
    winE = O.windows_e_intra("MP2", 2, 1)
    ij, kl, vv = T.transform(0, 0, winE, ol.CONV_E)
    assert len(vv) == 1 and abs(vv[0] - H2_MO["K12"]) < 1e-4
    e2 = T.transform_stream(0, 0, winE, ol.CONV_E, epsA=np.array(H2_MO["eps"]), lam=2.0)[3]
    assert abs(e2 + 0.0132) < 1e-4
    rij, rkl, rv = O.transform_e_intra(Cm, O.eri_packed_intra(sh), winE)
    assert abs(e2 - O.mp2_intra_from_pairs(rij, rkl, rv, 2, 1, np.array(H2_MO["eps"]))) <= 1e-12
