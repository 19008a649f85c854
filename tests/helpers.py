"""Shared helpers for the parity tests."""
import numpy as np


def dense_pairs(ij, kl, v, Ma, Mb):
    d = np.zeros((Ma, Mb))
    d[np.asarray(ij) - 1, np.asarray(kl) - 1] = v
    return d


def dense_quads(p, q, r, s, v, na, nb):
    d = np.zeros((na, na, nb, nb))
    d[np.asarray(p) - 1, np.asarray(q) - 1, np.asarray(r) - 1, np.asarray(s) - 1] = v
    return d


def assert_lists_match(got_keys, got_v, ref_keys, ref_v, tol=1e-10, edge=1e-9):
    """Entry lists agree: same values within tol; key sets may differ only for |v| within `edge`
    of the 1e-10 drop threshold (the reference's own summation order is unspecified, SURVEY.md hard part 4)."""
    g = {tuple(k): x for k, x in zip(zip(*got_keys), got_v)}
    r = {tuple(k): x for k, x in zip(zip(*ref_keys), ref_v)}
    for k in set(g) | set(r):
        a, b = g.get(k, 0.0), r.get(k, 0.0)
        assert abs(a - b) <= tol, (k, a, b)
        if (k in g) != (k in r):
            assert abs(a) < edge and abs(b) < edge, (k, a, b)


def rankk_stream_sums(T_vo, eps, occ, i_range, lam=2.0, tol=1e-10, a_block=128):
    """What lowdin_it_transform_stream returns (count, sum x, sum x^2, pair energy) for the MP2 window in transformer-E roles
    of a kind-K tensor, from its closed form: x = (a i|b j) = sum_k T^k[a,i] T^k[b,j], T^k = (C^T L^k C)[virt, occ];
    restricted to occupied i in i_range (0-based) = the occupied batch of one pass.  O(K V^2 O) per occupied orbital, in
    blocks of a_block virtuals to bound memory."""
    T_vo = np.asarray(T_vo)
    K, V, O = T_vo.shape
    eo, ev = np.asarray(eps[:occ]), np.asarray(eps[occ:])
    cnt = s1 = s2 = e2 = 0.0
    for i in i_range:
        for a0 in range(0, V, a_block):
            a1 = min(V, a0 + a_block)
            X = np.einsum("ka,kbj->abj", T_vo[:, a0:a1, i], T_vo, optimize=True)       # (a i|b j)
            Xe = np.einsum("kb,kaj->abj", T_vo[:, :, i], T_vo[:, a0:a1, :], optimize=True)  # (b i|a j)
            den = eo[i] + eo[None, None, :] - ev[a0:a1, None, None] - ev[None, :, None]
            cnt += float((np.abs(X) > tol).sum()); s1 += float(X.sum()); s2 += float((X * X).sum())
            e2 += float((X * (lam * X - Xe) / den).sum())
    return np.array([cnt, s1, s2, e2])
