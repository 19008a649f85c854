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
@pytest.mark.parametrize("G,logB,overlap", [(2, 5, 2), (3, 0, 2), (4, 1, 1), (2, 2, 0), (3, 1, 0)])
def test_local_group_intra_mp2_matches_oracle(O, G, logB, overlap):
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
                T.set_option(T.OPT_SLAB_BLOCK_LOG, logB)
                T.set_option(T.OPT_OVERLAP_EXCHANGE, overlap)     # exchange of chunk c under the first half of chunk c + 1 (2 = always, 1 = when chunks are wide), or on one stream
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


@pytest.mark.gpu
@pytest.mark.parametrize("G,logB", [(2, 0), (3, 2), (4, 5)])
def test_group_download_is_the_single_gpu_list(O, T, G, logB):
    """lowdin_it_transform on a group (collective; stored tensor sharded by rows at upload) + the merged download: the same
    entries in the same order as one GPU produces, for both record conventions."""
    n, occ = 15, 4
    packed = O.hash_packed_intra(44, n)
    Cm = O.random_orthonormal(n, n)
    lst = O.canonical_list_intra(packed, n)
    T.set_species(0, Cm)
    T.upload_ao(0, 0, *lst, stack=256)
    Ts = _group(G)
    try:
        for t in Ts:
            t.set_option(t.OPT_SLAB_BLOCK_LOG, logB)
            t.set_option(t.OPT_CHUNK_COLS, 37)
            t.set_species(0, Cm)
            t.upload_ao(0, 0, *lst, stack=256)          # every rank is pushed the whole list and keeps its own rows
        for conv, mode in ((ol.CONV_E, "MP2"), (ol.CONV_E, "ALL"), (ol.CONV_C, "MP2"), (ol.CONV_C, "PT2")):
            if conv == ol.CONV_E:
                win, sym = (O.windows_e_intra(mode, n, occ) if mode != "ALL" else [1, n] * 4), False
            else:
                win, sym = O.windows_c_intra(mode, n, occ)
            one = T.transform(0, 0, win, conv, symmetric=sym)
            got = capi.group_transform(Ts, 0, 0, win, conv, symmetric=sym)
            assert len(got[-1]) == len(one[-1]) > 0
            for a, b in zip(got[:-1], one[:-1]):
                assert np.array_equal(a, b)               # same index lists, same order
            assert np.abs(got[-1] - one[-1]).max() <= 1e-12
            # the per-rank segments partition the window pairs
            idx = np.concatenate([t.result_segments()[0] for t in Ts])
            assert len(set(idx)) == len(idx)
    finally:
        _close(Ts)


@pytest.mark.gpu
def test_group_list_mode_and_materialized_sources(O):
    """The list-driven first quarter and a device-materialised tensor on a group of 3 ranks."""
    n, occ = 14, 4
    packed = O.hash_packed_intra(12, n)
    Cm = O.random_orthonormal(n, n)
    lst = O.canonical_list_intra(packed, n)
    eps = O.synthetic_eps(occ, n)
    win = O.windows_e_intra("MP2", n, occ)
    rij, rkl, rv = O.transform_e_intra(Cm, packed, win)
    want = np.array([len(rv), rv.sum(), (rv * rv).sum(), O.mp2_intra_from_pairs(rij, rkl, rv, n, occ, eps)])
    Ts = _group(3)
    try:
        def work(r, T):
            T.set_option(T.OPT_SLAB_BLOCK_LOG, 1)
            T.set_option(T.OPT_CHUNK_COLS, 30)
            T.set_species(0, Cm)
            T.set_option(T.OPT_AO_LIST, 1)
            T.upload_ao(0, 0, *lst, stack=128)
            T.set_option(T.OPT_AO_LIST, 0)
            a = T.transform_stream(0, 0, win, ol.CONV_E, occ_batch=2, epsA=eps)
            T.set_generator(0, 0, 12)
            T.materialize(0, 0)
            b = T.transform_stream(0, 0, win, ol.CONV_E, occ_batch=3, epsA=eps)
            return np.concatenate([a, b])
        got = np.sum(capi.run_ranks(Ts, work), axis=0)
        for g in (got[:4], got[4:]):
            assert g[0] == want[0] and np.abs(g[1:] - want[1:]).max() <= 1e-9, (g, want)
    finally:
        _close(Ts)


@pytest.mark.gpu
def test_upload_before_group_is_rejected(O):
    n = 6
    Cm = O.random_orthonormal(n, n)
    lst = O.canonical_list_intra(O.hash_packed_intra(1, n), n)
    Ts = [ol.Transformer(0) for _ in range(2)]
    try:
        for t in Ts:
            t.set_species(0, Cm)
            t.upload_ao(0, 0, *lst)
        capi.local_group(Ts)
        with pytest.raises(ol.LowdinITError, match="before the communicator"):
            capi.group_transform(Ts, 0, 0, [1, n] * 4, ol.CONV_E)
    finally:
        _close(Ts)


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["C", "E"])
def test_host_mirror_group_file_to_file(O, T, tmp_path, method):
    """The reference host's call on a group of 3 GPUs (lowdin_host_group_atomic_to_molecular): the .ints streams are read once and
    pushed to every handle, ONE moint.dat comes out -- the same records as the single-handle call writes, and the oracle's integrals."""
    import shutil
    n, occ, S, nfiles = 13, 4, 64, 2
    packed = O.hash_packed_intra(61, n)
    Cm = O.random_orthonormal(n, n)
    lst = O.canonical_list_intra(packed, n)
    for t in range(nfiles):
        capi.host_write_ints_file(str(tmp_path / f"{t}E-.ints"), S, *[x[t::nfiles] for x in lst])
    ctl = capi.host_control(method, "MP2", stack=S, nfiles=nfiles, scratch_dir=str(tmp_path))
    sp = capi.host_species("E-", 1, n, occ, coeff=Cm)
    cnt1 = capi.host_transform_one_species(T, ctl, sp)
    shutil.copy(tmp_path / "E-moint.dat", tmp_path / "one.dat")
    Ts = _group(3)
    try:
        for t in Ts:
            t.set_option(t.OPT_SLAB_BLOCK_LOG, 1)
            t.set_option(t.OPT_CHUNK_COLS, 25)
        cnt = capi.host_group_transform(Ts, ctl, sp)
    finally:
        _close(Ts)
    assert cnt == cnt1 > 0
    a, b = open(tmp_path / "one.dat", "rb").read(), open(tmp_path / "E-moint.dat", "rb").read()
    assert len(a) == len(b)
    if method == "E":      # records: int64 ij[S], kl[S], float64 v[S] between 4-byte markers
        rec = 4 + 24 * S + 4
        for o in range(0, len(a), rec):
            assert a[o:o + 4 + 16 * S] == b[o:o + 4 + 16 * S]                       # same pair ids in the same order
            va, vb = np.frombuffer(a, np.float64, S, o + 4 + 16 * S), np.frombuffer(b, np.float64, S, o + 4 + 16 * S)
            assert np.abs(va - vb).max() <= 1e-13
    else:
        rec = 4 + 24 * S + 4
        for o in range(0, len(a), rec):
            assert a[o:o + 4 + 16 * S] == b[o:o + 4 + 16 * S]                       # same p,q,r,s in the same order
            va, vb = np.frombuffer(a, np.float64, S, o + 4 + 16 * S), np.frombuffer(b, np.float64, S, o + 4 + 16 * S)
            assert np.abs(va - vb).max() <= 1e-13
