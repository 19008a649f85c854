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
