"""Multi-rank parity on ONE GPU: the ranks are handles of this process joined by lowdin_it_comm_init_local (peer copies
instead of NCCL, same collective semantics, same division of slabs and slots, same blocked layout after the exchange).
The driver's GPU test box has a single B200; this is where the N>1 path is checked against the oracle there.  The NCCL
transport itself is covered by tests/test_multi_gpu.py (needs >= 2 GPUs) and by bench.py's parity line at --gpus N."""
import numpy as np
import pytest

import openlowdin_b200 as ol
from openlowdin_b200 import capi


def _group(G):
    Ts = [ol.Transformer(0) for _ in range(G)]
    capi.local_group(Ts)
    return Ts


def _close(Ts):
    for t in Ts:
        t.close()


@pytest.mark.gpu
@pytest.mark.parametrize("G", [2, 3, 4])
def test_local_group_intra_mp2_matches_oracle(O, G):
    n, occ, seed = 24, 6, 31337
    packed = O.hash_packed_intra(seed, n)
    Cm = O.random_orthonormal(n, n)
    eps = O.synthetic_eps(occ, n)
    win = O.windows_e_intra("MP2", n, occ)
    rij, rkl, rv = O.transform_e_intra(Cm, packed, win)
    want = np.array([len(rv), rv.sum(), (rv * rv).sum(), O.mp2_intra_from_pairs(rij, rkl, rv, n, occ, eps, lam=2.0)])
    Ts = _group(G)
    try:
        for cols, qb in ((0, 0), (60, 4), (1, 2), (200, 3)):
            def work(r, T):
                T.set_species(0, Cm)
                T.set_generator(0, 0, seed)
                T.set_option(T.OPT_CHUNK_COLS, cols)
                return T.transform_stream(0, 0, win, ol.CONV_E, occ_batch=qb, epsA=eps, lam=2.0)
            got = np.sum(capi.run_ranks(Ts, work), axis=0)
            assert got[0] == want[0], (cols, qb, got, want)
            assert np.abs(got[1:] - want[1:]).max() <= 1e-9, (cols, qb, got, want)
    finally:
        _close(Ts)


@pytest.mark.gpu
@pytest.mark.parametrize("G", [2, 3])
def test_local_group_inter_and_stored_sources(O, G):
    """Inter-species MP2 and an UPLOADED (stored) intra tensor on G ranks."""
    (na, nb), (oa, ob), seed = (21, 16), (5, 2), 99
    rect = O.hash_rect_inter(seed, na, nb)
    Ca, Cb = O.random_orthonormal(na, 3), O.random_orthonormal(nb, 4)
    ea, eb = O.synthetic_eps(oa, na), O.synthetic_eps(ob, nb)
    win = O.windows_e_inter("MP2", na, nb, oa, ob)
    rij, rkl, rv = O.transform_e_inter(Ca, Cb, rect, win)
    want = np.array([len(rv), rv.sum(), (rv * rv).sum(),
                     O.mp2_inter_from_pairs(rij, rkl, rv, na, nb, oa, ob, ea, eb, charge_a=1.0, charge_b=1.0, lam_a=1.0, lam_b=1.0)])
    n, occ = 17, 4
    packed = O.hash_packed_intra(5, n)
    Cm = O.random_orthonormal(n, n)
    eps = O.synthetic_eps(occ, n)
    win2 = O.windows_e_intra("MP2", n, occ)
    sij, skl, sv = O.transform_e_intra(Cm, packed, win2)
    want2 = np.array([len(sv), sv.sum(), (sv * sv).sum(), O.mp2_intra_from_pairs(sij, skl, sv, n, occ, eps, lam=2.0)])
    lst = O.canonical_list_intra(packed, n)
    Ts = _group(G)
    try:
        def work(r, T):
            T.set_species(0, Ca); T.set_species(1, Cb)
            T.set_generator(0, 1, seed)
            T.set_option(T.OPT_CHUNK_COLS, 40)
            a = T.transform_stream(0, 1, win, ol.CONV_E, occ_batch=2, epsA=ea, epsB=eb)
            T.set_species(2, Cm)
            T.upload_ao(2, 2, *lst, stack=512)
            b = T.transform_stream(2, 2, win2, ol.CONV_E, occ_batch=3, epsA=eps, lam=2.0)
            return np.concatenate([a, b])
        got = np.sum(capi.run_ranks(Ts, work), axis=0)
        assert got[0] == want[0] and got[4] == want2[0], (got, want, want2)
        assert np.abs(got[1:4] - want[1:]).max() <= 1e-9 and np.abs(got[5:] - want2[1:]).max() <= 1e-9, (got, want, want2)
    finally:
        _close(Ts)


@pytest.mark.gpu
def test_two_handles_one_process_are_independent(O):
    """Two plain handles (no group) used one after the other: per-device kernel attributes and buffers are per handle."""
    n, occ = 19, 5
    packed = O.hash_packed_intra(7, n)
    Cm = O.random_orthonormal(n, n)
    win = O.windows_e_intra("MP2", n, occ)
    ref = O.transform_e_intra(Cm, packed, win)
    M = O.npairs(n)
    import torch
    devs = [0, 1] if torch.cuda.device_count() > 1 else [0, 0]
    Ts = [ol.Transformer(d) for d in devs]
    try:
        for T in Ts:
            T.set_species(0, Cm)
            T.set_generator(0, 0, 7)
            ij, kl, v = T.transform(0, 0, win, ol.CONV_E)
            assert np.abs(O.pairs_to_dense(ij, kl, v, M, M) - O.pairs_to_dense(*ref, M, M)).max() <= 1e-10
    finally:
        _close(Ts)
