"""The AO upload path (lowdin_it_ao_begin / _push_stacks / _push_blocks / _end): terminator, index check and scatter run on
the device (TransformIntegralsC.f90:247-298); raw .ints blocks and the five-array form must give the same packed tensor."""
import numpy as np
import pytest

import openlowdin_b200 as ol


def _raw_blocks(p, q, r, s, v, S):
    """The bytes lowdin-ints writes (Libint2Iface.cpp:3414-3426): blocks of int32 p[S],q[S],r[S],s[S]; float64 v[S];
    the last block carries p = -1 after its last entry, the rest of it is garbage."""
    n = len(v)
    nblk = n // S + 1
    out = bytearray()
    rng = np.random.default_rng(1)
    for t in range(nblk):
        a, b = t * S, min(n, (t + 1) * S)
        arrs = [rng.integers(1, 5, S).astype(np.int32) for _ in range(4)] + [rng.uniform(-1, 1, S)]   # garbage tail
        for dst, src in zip(arrs, (p, q, r, s, v)):
            dst[:b - a] = src[a:b]
        if b - a < S:
            arrs[0][b - a] = -1
        for x in arrs:
            out += x.tobytes()
    return np.frombuffer(bytes(out), np.uint8)


def _dense_slabs(T, n, M):
    return T.debug_expand(0, 0, 0, M)


@pytest.mark.gpu
@pytest.mark.parametrize("staging", [0, 1 << 16])
def test_raw_blocks_and_five_arrays_give_the_same_tensor(O, T, staging):
    n, S = 9, 100
    M = O.npairs(n)
    packed = O.hash_packed_intra(3, n)
    lst = O.canonical_list_intra(packed, n)
    T.set_species(0, O.random_orthonormal(n, n))
    if staging:
        T.set_option(T.OPT_STAGING_BYTES, staging)     # several pieces per call
    try:
        T.upload_ao(0, 0, *lst, stack=S)
        A = _dense_slabs(T, n, M)
        T.upload_ao_blocks(0, 0, _raw_blocks(*lst, S), S)
        B = _dense_slabs(T, n, M)
    finally:
        T.set_option(T.OPT_STAGING_BYTES, 96 << 20)
    sq = O.packed_to_square(packed, M)
    i1, i2 = np.triu_indices(n)
    ref = np.zeros((M, n, n)); ref[:, i1, i2] = sq; ref[:, i2, i1] = sq
    assert np.array_equal(A, ref) and np.array_equal(B, ref)


@pytest.mark.gpu
def test_terminator_ends_one_call_not_the_upload(O, T):
    """Two files = two calls, each ending in its own terminator stack (C.f90:247-296 loops over the per-thread files)."""
    n, S = 7, 64
    M = O.npairs(n)
    packed = O.hash_packed_intra(4, n)
    p, q, r, s, v = O.canonical_list_intra(packed, n)
    h = len(v) // 2
    T.set_species(0, O.random_orthonormal(n, n))
    L = T.L
    T._ck(L.lowdin_it_ao_begin(T.h, 0, 0, 0))
    for sl in (slice(0, h), slice(h, None)):
        raw = _raw_blocks(p[sl], q[sl], r[sl], s[sl], v[sl], S)
        T._ck(L.lowdin_it_ao_push_blocks(T.h, raw.ctypes.data, raw.size // (24 * S), S))
    T._ck(L.lowdin_it_ao_end(T.h))
    sq = O.packed_to_square(packed, M)
    i1, i2 = np.triu_indices(n)
    ref = np.zeros((M, n, n)); ref[:, i1, i2] = sq; ref[:, i2, i1] = sq
    assert np.array_equal(_dense_slabs(T, n, M), ref)


@pytest.mark.gpu
def test_bad_index_is_reported_at_ao_end(O, T):
    n = 5
    T.set_species(0, np.eye(n))
    one = lambda x: np.array([x], np.int32)  # noqa: E731
    with pytest.raises(ol.LowdinITError, match="outside the basis"):
        T.upload_ao(0, 0, one(1), one(2), one(6), one(1), np.array([1.0]))
    with pytest.raises(ol.LowdinITError):      # the set stays unusable
        T.transform(0, 0, [1, n] * 4, ol.CONV_E)
    with pytest.raises(ol.LowdinITError, match="outside the basis"):
        T.upload_ao(0, 0, one(0), one(2), one(1), one(1), np.array([1.0]))
    # entries after the terminator are garbage and must not be checked
    p = np.array([1, -1, 99], np.int32); o = np.array([1, 1, 1], np.int32)
    T._ck(T.L.lowdin_it_ao_begin(T.h, 0, 0, 0))
    T._ck(T.L.lowdin_it_ao_push_stacks(T.h, p, o, o, o, np.array([0.5, 7.0, 7.0]), 3))
    T._ck(T.L.lowdin_it_ao_end(T.h))
    X = T.debug_expand(0, 0, 0, 1)
    assert X[0, 0, 0] == 0.5 and np.count_nonzero(X) == 1


@pytest.mark.gpu
def test_async_push_option(O, T):
    n, S = 8, 50
    M = O.npairs(n)
    packed = O.hash_packed_intra(9, n)
    lst = O.canonical_list_intra(packed, n)
    T.set_species(0, O.random_orthonormal(n, n))
    T.set_option(T.OPT_ASYNC_PUSH, 1)
    try:
        T.upload_ao(0, 0, *lst, stack=len(lst[4]) + 1)
    finally:
        T.set_option(T.OPT_ASYNC_PUSH, 0)
    sq = O.packed_to_square(packed, M)
    i1, i2 = np.triu_indices(n)
    ref = np.zeros((M, n, n)); ref[:, i1, i2] = sq; ref[:, i2, i1] = sq
    assert np.array_equal(_dense_slabs(T, n, M), ref)
