"""oracle/oracle.py -- Python face of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  It binds

  * oracle/liboracle.so   -- our C restatement of the reference's transformers C, E, D,
                             the AO-list loader, the MO-integral reader addressing and the
                             APMO-MP2 formula (it_oracle.c, energy_oracle.c);
  * oracle/_ref/libref_d.so -- the reference's OWN transformer D (IntTransfD.cpp) compiled in
                             place by oracle/Makefile, when present;

and restates in plain Python the integer-only host logic (window tables of
TransformIntegralsC.f90:1436-1963 and TransformIntegralsE.f90:1899-2418, the
partialTransform choice of IntegralTransformation.f90:106-126) plus a dense
numpy.einsum transform as a third opinion.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f64pf = np.ctypeslib.ndpointer(np.float64, flags="F_CONTIGUOUS")


def build(verbose: bool = False) -> None:
    """Compile liboracle.so (and _ref/libref_d.so when /root/reference is present)."""
    r = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("oracle build failed")


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        L.orc_pair_id.restype = C.c_int64
        L.orc_pair_id.argtypes = [C.c_int64] * 3
        L.orc_ioff.restype = C.c_int64
        L.orc_ioff.argtypes = [C.c_int64] * 2
        L.orc_packed_index.restype = C.c_int64
        L.orc_packed_index.argtypes = [C.c_int64] * 3
        L.orc_pairmap.restype = C.c_int64
        L.orc_pairmap.argtypes = [C.c_int] * 5 + [C.c_void_p]
        L.orc_scatter_intra.restype = C.c_int64
        L.orc_scatter_intra.argtypes = [_i32p] * 4 + [_f64p, C.c_int64, C.c_int, _f64p]
        L.orc_scatter_inter.restype = C.c_int64
        L.orc_scatter_inter.argtypes = [_i32p] * 4 + [_f64p, C.c_int64, C.c_int, C.c_int, C.c_int, _f64p]
        L.orc_transform_e_intra.restype = C.c_int64
        L.orc_transform_e_intra.argtypes = [C.c_int, _f64pf, C.c_int, _f64p, _i32p, _i64p, _i64p, _f64p, C.c_int64]
        L.orc_transform_e_inter.restype = C.c_int64
        L.orc_transform_e_inter.argtypes = [C.c_int, C.c_int, _f64pf, C.c_int, _f64pf, C.c_int, _f64p, _i32p,
                                            _i64p, _i64p, _f64p, C.c_int64]
        L.orc_transform_c_intra.restype = C.c_int64
        L.orc_transform_c_intra.argtypes = [C.c_int, _f64pf, C.c_int, _f64p, _i32p, C.c_int] + [_i32p] * 4 + [
            _f64p, C.c_int64]
        L.orc_transform_c_inter.restype = C.c_int64
        L.orc_transform_c_inter.argtypes = [C.c_int, C.c_int, _f64pf, C.c_int, _f64pf, C.c_int, _f64p, _i32p,
                                            C.c_int] + [_i32p] * 4 + [_f64p, C.c_int64]
        L.orc_d_multi_index.restype = C.c_int64
        L.orc_d_multi_index.argtypes = [C.c_int64] * 4
        L.orc_transform_d_intra.restype = None
        L.orc_transform_d_intra.argtypes = [_f64pf, _f64p, C.c_int]
        L.orc_transform_d_inter.restype = None
        L.orc_transform_d_inter.argtypes = [_f64pf, _f64pf, _f64p, C.c_int, C.c_int]
        L.orc_hash_value.restype = C.c_double
        L.orc_hash_value.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_fill_hash_intra.restype = None
        L.orc_fill_hash_intra.argtypes = [C.c_uint64, C.c_int, _f64p]
        L.orc_fill_hash_inter.restype = None
        L.orc_fill_hash_inter.argtypes = [C.c_uint64, C.c_int, C.c_int, _f64p]
        L.orc_fill_gen_intra.restype = None
        L.orc_fill_gen_intra.argtypes = [C.c_int, C.c_uint64, C.c_int, _f64p]
        L.orc_fill_gen_inter.restype = None
        L.orc_fill_gen_inter.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int, _f64p]
        L.orc_gen_value.restype = C.c_double
        L.orc_gen_value.argtypes = [C.c_int, C.c_uint64, C.c_uint64]
        L.orc_e_first_half_sample.restype = C.c_double
        L.orc_e_first_half_sample.argtypes = [C.c_uint64, C.c_int, _f64pf, C.c_int, _i32p, C.c_int64, C.c_int64, C.c_int]
        L.orc_direct_first_quarter.restype = None
        L.orc_direct_first_quarter.argtypes = [C.c_int, _f64p, _i32p, _i32p, _i32p, _i32p, _f64p, C.c_int64, _f64p]
        L.orc_e_first_half_values.restype = C.c_int64
        L.orc_e_first_half_values.argtypes = [C.c_int, C.c_uint64, C.c_int, _f64pf, C.c_int, _i32p, C.c_int64, C.c_int64, C.c_int, C.c_void_p]
        L.orc_reader_pairs_intra.restype = None
        L.orc_reader_pairs_intra.argtypes = [_i64p, _i64p, _f64p, C.c_int64, C.c_int, _f64p]
        L.orc_reader_quads_intra.restype = None
        L.orc_reader_quads_intra.argtypes = [_i32p] * 4 + [_f64p, C.c_int64, C.c_int, _f64p]
        L.orc_reader_pairs_inter.restype = None
        L.orc_reader_pairs_inter.argtypes = [_i64p, _i64p, _f64p, C.c_int64, C.c_int, C.c_int, _f64p]
        L.orc_reader_quads_inter.restype = None
        L.orc_reader_quads_inter.argtypes = [_i32p] * 4 + [_f64p, C.c_int64, C.c_int, C.c_int, _f64p]
        L.orc_mp2_intra.restype = C.c_double
        L.orc_mp2_intra.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _f64p]
        L.orc_mp2_intra_scale.restype = C.c_double
        L.orc_mp2_intra_scale.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double]
        L.orc_mp2_inter.restype = C.c_double
        L.orc_mp2_inter.argtypes = [_f64p] + [C.c_int] * 8 + [C.c_double] * 4 + [_f64p, _f64p]
        L.orc_blas_bind.restype = C.c_int
        L.orc_blas_bind.argtypes = [C.c_char_p, C.c_char_p]
        L.orc_blas_is_external.restype = C.c_int
        _lib = L
    return _lib


def bind_openblas() -> bool:
    """Point the dgemm_ shim at the OpenBLAS bundled with opencv (plain `dgemm_` symbol)."""
    L = lib()
    if L.orc_blas_is_external():
        return True
    for sp in sys.path:
        d = os.path.join(sp, "opencv_python_headless.libs")
        blas = sorted(glob.glob(os.path.join(d, "libopenblas*.so*")))
        gf = sorted(glob.glob(os.path.join(d, "libgfortran*.so*")))
        if blas:
            try:  # libgfortran's own dependency lives in scipy.libs in this image
                for qm in sorted(glob.glob(os.path.join(sp, "scipy.libs", "libquadmath*.so*"))):
                    C.CDLL(qm, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
            if L.orc_blas_bind((gf[0] if gf else "").encode(), blas[0].encode()) == 0:
                return True
    return False


def ref():
    """The reference's own transformer D (None when oracle/_ref/libref_d.so is absent)."""
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libref_d.so")
        if not os.path.exists(path):
            return None
        lib()  # provides dgemm_ (RTLD_GLOBAL)
        R = C.CDLL(path)
        R.c_integrals_transform_all.restype = None
        R.c_integrals_transform_all.argtypes = [_f64pf, _f64p, C.c_int]
        R.c_integrals_transform_inter_all.restype = None
        R.c_integrals_transform_inter_all.argtypes = [_f64pf, _f64pf, _f64p, C.c_int, C.c_int]
        _ref = R
    return _ref


# ----------------------------------------------------------------------------------------
# index helpers
# ----------------------------------------------------------------------------------------
def pair_id(i, j, n):
    return lib().orc_pair_id(i, j, n)


def npairs(n):
    return n * (n + 1) // 2


def pair_table(n):
    """xy(p,q) as a 0-based [n,n] array holding 0-based pair ids (row-wise upper triangle)."""
    xy = np.zeros((n, n), dtype=np.int64)
    iu = np.triu_indices(n)
    xy[iu] = np.arange(npairs(n))
    xy.T[iu] = np.arange(npairs(n))
    return xy


def packed_to_square(packed, M):
    """Symmetric-packed (C/E layout, row lo holds hi=lo..M) -> dense symmetric [M,M]."""
    sq = np.zeros((M, M))
    iu = np.triu_indices(M)
    sq[iu] = packed
    sq.T[iu] = packed
    return sq


def square_to_packed(sq):
    return np.ascontiguousarray(sq[np.triu_indices(sq.shape[0])])


def dense4_from_square(sq, xy_a, xy_b=None):
    """[M_a, M_b] pair matrix -> dense (mu nu | lam sig) [na,na,nb,nb]."""
    xy_b = xy_a if xy_b is None else xy_b
    return sq[xy_a[:, :, None, None], xy_b[None, None, :, :]]


def einsum_transform(ao4, Ca, Cb=None):
    """(pq|rs) = sum C(mu,p) C(nu,q) (mu nu|lam sig) C'(lam,r) C'(sig,s) -- third opinion."""
    Cb = Ca if Cb is None else Cb
    t = np.einsum("mnls,mp->pnls", ao4, Ca, optimize=True)
    t = np.einsum("pnls,nq->pqls", t, Ca, optimize=True)
    t = np.einsum("pqls,lr->pqrs", t, Cb, optimize=True)
    return np.einsum("pqrs,st->pqrt", t, Cb, optimize=True)


def random_orthonormal(n, seed):
    """Q factor of qr(default_rng(seed).standard_normal((n,n))), column-major C(mu,p)."""
    q, _ = np.linalg.qr(np.random.default_rng(seed).standard_normal((n, n)))
    return np.asfortranarray(q)


def synthetic_eps(occ, n):
    return np.concatenate([np.linspace(-2.0, -0.5, occ), np.linspace(0.2, 3.0, n - occ)])


def hash_packed_intra(seed, n, kind=1):
    """Synthetic AO tensor, packed C/E layout.  kind 1 = H (splitmix64), 2 = F (mul-fold-mul)."""
    M = npairs(n)
    out = np.empty(M * (M + 1) // 2)
    lib().orc_fill_gen_intra(kind, seed, n, out)
    return out


def hash_rect_inter(seed, na, nb, kind=1):
    out = np.empty(npairs(na) * npairs(nb))
    lib().orc_fill_gen_inter(kind, seed, na, nb, out)
    return out


def rankk_square(seed, na, nb=None, K=8):
    """kind K: (mu nu|lam sig) = sum_k L^k_{mu nu} R^k_{lam sig}; returns the [M_a,M_b] pair matrix."""
    rng = np.random.default_rng(seed)
    La = rng.uniform(-1, 1, (K, na, na))
    La = 0.5 * (La + La.transpose(0, 2, 1))
    iu = np.triu_indices(na)
    la = La[:, iu[0], iu[1]]
    if nb is None:
        return la.T @ la, La, La
    Lb = rng.uniform(-1, 1, (K, nb, nb))
    Lb = 0.5 * (Lb + Lb.transpose(0, 2, 1))
    iub = np.triu_indices(nb)
    return la.T @ Lb[:, iub[0], iub[1]], La, Lb


# ----------------------------------------------------------------------------------------
# AO list (the .ints stack content) helpers
# ----------------------------------------------------------------------------------------
def canonical_list_intra(packed, n, drop=1e-10):
    """Canonical AO list as lowdin-ints writes it: i>=j, k>=l, (ij)>=(kl) (Iterators.cpp:45-77),
    1-based, |v|>1e-10 only (Libint2Iface.cpp:369).  Built from a packed (C/E layout) tensor."""
    xy = pair_table(n)
    M = npairs(n)
    sq = packed_to_square(packed, M)
    P, Q, R, S, V = [], [], [], [], []
    for i in range(n):
        for j in range(i + 1):
            for k in range(i + 1):
                for l in range(k + 1):
                    if k == i and l > j:
                        continue
                    v = sq[xy[i, j], xy[k, l]]
                    if abs(v) > drop:
                        P.append(i + 1); Q.append(j + 1); R.append(k + 1); S.append(l + 1); V.append(v)
    return (np.array(P, np.int32), np.array(Q, np.int32), np.array(R, np.int32), np.array(S, np.int32),
            np.array(V, np.float64))


def canonical_list_inter(rect, na, nb, drop=1e-10):
    """Inter list: p<=q of species A, r<=s of species B (Libint2Iface.cpp:1052-1059), 1-based."""
    Ma, Mb = npairs(na), npairs(nb)
    sq = rect.reshape(Mb, Ma)
    ia, ib = np.triu_indices(na), np.triu_indices(nb)
    P, Q, R, S, V = [], [], [], [], []
    for pq in range(Ma):
        for rs in range(Mb):
            v = sq[rs, pq]
            if abs(v) > drop:
                P.append(ia[0][pq] + 1); Q.append(ia[1][pq] + 1); R.append(ib[0][rs] + 1); S.append(ib[1][rs] + 1)
                V.append(v)
    return (np.array(P, np.int32), np.array(Q, np.int32), np.array(R, np.int32), np.array(S, np.int32),
            np.array(V, np.float64))


def scatter_intra(p, q, r, s, v, n):
    M = npairs(n)
    packed = np.zeros(M * (M + 1) // 2)
    lib().orc_scatter_intra(p, q, r, s, v, len(v), n, packed)
    return packed


def scatter_inter(p, q, r, s, v, na, nb, swapped=False):
    rect = np.zeros(npairs(na) * npairs(nb))
    lib().orc_scatter_inter(p, q, r, s, v, len(v), na, nb, int(swapped), rect)
    return rect


# ----------------------------------------------------------------------------------------
# transformers
# ----------------------------------------------------------------------------------------
def _win(win):
    w = np.ascontiguousarray(win, dtype=np.int32)
    assert w.shape == (8,)
    return w


def transform_e_intra(Cm, packed, win):
    n = Cm.shape[0]
    w = _win(win)
    cap = max(1, lib().orc_pairmap(w[0], w[1], w[2], w[3], n, None) * lib().orc_pairmap(w[4], w[5], w[6], w[7], n, None))
    ij = np.zeros(cap, np.int64); kl = np.zeros(cap, np.int64); v = np.zeros(cap)
    m = lib().orc_transform_e_intra(n, np.asfortranarray(Cm), Cm.shape[0], packed, w, ij, kl, v, cap)
    return ij[:m], kl[:m], v[:m]


def transform_e_inter(Ca, Cb, rect, win):
    na, nb = Ca.shape[0], Cb.shape[0]
    w = _win(win)
    cap = max(1, lib().orc_pairmap(w[0], w[1], w[2], w[3], na, None) * lib().orc_pairmap(w[4], w[5], w[6], w[7], nb, None))
    ij = np.zeros(cap, np.int64); kl = np.zeros(cap, np.int64); v = np.zeros(cap)
    m = lib().orc_transform_e_inter(na, nb, np.asfortranarray(Ca), na, np.asfortranarray(Cb), nb, rect, w, ij, kl, v, cap)
    return ij[:m], kl[:m], v[:m]


def _cap4(w):
    return max(1, int(max(0, w[1] - w[0] + 1)) * int(max(0, w[3] - w[2] + 1)) * int(max(0, w[5] - w[4] + 1))
               * int(max(0, w[7] - w[6] + 1)))


def transform_c_intra(Cm, packed, win, symmetric):
    n = Cm.shape[0]
    w = _win(win)
    cap = _cap4(w)
    o = [np.zeros(cap, np.int32) for _ in range(4)]
    v = np.zeros(cap)
    m = lib().orc_transform_c_intra(n, np.asfortranarray(Cm), n, packed, w, int(symmetric), *o, v, cap)
    return tuple(x[:m] for x in o) + (v[:m],)


def transform_c_inter(Ca, Cb, rect, win, symmetric):
    na, nb = Ca.shape[0], Cb.shape[0]
    w = _win(win)
    cap = _cap4(w)
    o = [np.zeros(cap, np.int32) for _ in range(4)]
    v = np.zeros(cap)
    m = lib().orc_transform_c_inter(na, nb, np.asfortranarray(Ca), na, np.asfortranarray(Cb), nb, rect, w,
                                    int(symmetric), *o, v, cap)
    return tuple(x[:m] for x in o) + (v[:m],)


def d_pack_intra(sq_lower_pairs):
    """[m,m] pair matrix in D's 0-based lower-triangular pair numbering -> ERIS[ij(ij+1)/2+kl]."""
    m = sq_lower_pairs.shape[0]
    il = np.tril_indices(m)
    return np.ascontiguousarray(sq_lower_pairs[il])


def d_pair_table(n):
    """D / ReadIntegrals_index2: 0-based lower triangle i(i+1)/2+j (ReadIntegrals.f90:177-188)."""
    t = np.zeros((n, n), dtype=np.int64)
    il = np.tril_indices(n)
    t[il] = np.arange(npairs(n))
    t.T[il] = np.arange(npairs(n))
    return t


def transform_d_intra(Cm, eris, use_reference=False):
    out = np.array(eris, dtype=np.float64, copy=True)
    n = Cm.shape[0]
    if use_reference:
        ref().c_integrals_transform_all(np.asfortranarray(Cm), out, n)
    else:
        lib().orc_transform_d_intra(np.asfortranarray(Cm), out, n)
    return out


def transform_d_inter(Ca, Cb, eris, use_reference=False):
    out = np.array(eris, dtype=np.float64, copy=True)
    if use_reference:
        ref().c_integrals_transform_inter_all(np.asfortranarray(Ca), np.asfortranarray(Cb), out, Ca.shape[0], Cb.shape[0])
    else:
        lib().orc_transform_d_inter(np.asfortranarray(Ca), np.asfortranarray(Cb), out, Ca.shape[0], Cb.shape[0])
    return out


# ----------------------------------------------------------------------------------------
# dense comparison helpers ("absent == 0", SURVEY.md hard part 4)
# ----------------------------------------------------------------------------------------
def pairs_to_dense(ij, kl, v, Ma, Mb):
    d = np.zeros((Ma, Mb))
    d[ij - 1, kl - 1] = v
    return d


def quads_to_dense(p, q, r, s, v, na, nb):
    d = np.zeros((na, na, nb, nb))
    d[p - 1, q - 1, r - 1, s - 1] = v
    return d


# ----------------------------------------------------------------------------------------
# host logic restated: partialTransform and the window tables
# ----------------------------------------------------------------------------------------
def partial_transform(mp_correction=0, pt_order=0, en_correction=0, ci_level="NONE"):
    """IntegralTransformation.f90:106-126."""
    if mp_correction == 2 and pt_order == 0 and en_correction == 0 and ci_level == "NONE":
        return "MP2"
    if pt_order == 2 and mp_correction == 0 and en_correction == 0 and ci_level == "NONE":
        return "PT2"
    if pt_order == 2 and mp_correction == 2 and en_correction == 0 and ci_level == "NONE":
        return "MP2-PT2"
    if ci_level != "NONE":
        return "ALL"
    return "BOUNDS"


def windows_c_intra(mode, n, occ, core=0, active=0, ionize_mo=0, pt_transition_operator=False):
    """TransformIntegralsC.f90:1436-1622 -> (win[8], symmetric)."""
    c = core if core != 0 else 0
    a = active if active != 0 else n
    sym = True
    w = dict(p=(c + 1, a), q=(c + 1, a), r=(c + 1, a), s=(c + 1, a))
    if mode == "ALL":
        w = dict(p=(1, n), q=(1, n), r=(1, n), s=(1, n))
    if mode == "ALLACTIVE":
        w = dict(p=(1, a), q=(1, a), r=(1, a), s=(1, a))
    if mode == "MP2":
        w = dict(p=(c + 1, occ), q=(occ + 1, a), r=(c + 1, occ), s=(occ + 1, a))
    if mode == "PT2":
        sym = False
        if ionize_mo == 0:
            w = dict(p=(occ, occ + 1), q=(c + 1, a), r=(c + 1, occ), s=(occ + 1, a))
        elif pt_transition_operator:
            w = dict(p=(ionize_mo, ionize_mo), q=(c + 1, a), r=(c + 1, occ), s=(c + 1, a))
        else:
            w = dict(p=(ionize_mo, ionize_mo), q=(c + 1, a), r=(c + 1, occ), s=(occ + 1, a))
    if mode == "MP2-PT2":
        sym = False
        if ionize_mo == 0:
            w = dict(p=(c + 1, occ + 1), q=(c + 1, a), r=(c + 1, occ), s=(occ + 1, a))
        elif pt_transition_operator:
            w = dict(p=(c + 1, max(ionize_mo, occ)), q=(c + 1, a), r=(c + 1, occ), s=(c + 1, a))
        else:
            w = dict(p=(c + 1, max(ionize_mo, occ)), q=(c + 1, a), r=(c + 1, occ), s=(occ + 1, a))
    return [*w["p"], *w["q"], *w["r"], *w["s"]], sym


def windows_e_intra(mode, n, occ, core=0, active=0, ionize_mo=0, pt_transition_operator=False):
    """TransformIntegralsE.f90:1899-2073 -> win[8] (no ALL case: falls to the default window)."""
    c = core if core != 0 else 0
    a = active if active != 0 else n
    w = dict(p=(c + 1, a), q=(c + 1, a), r=(c + 1, a), s=(c + 1, a))
    if mode == "ALLACTIVE":
        w = dict(p=(1, a), q=(1, a), r=(1, a), s=(1, a))
    if mode == "MP2":
        w = dict(p=(occ + 1, a), q=(c + 1, occ), r=(occ + 1, a), s=(c + 1, occ))
    if mode == "PT2":
        if ionize_mo == 0:
            w = dict(q=(c + 1, occ + 1), p=(c + 1, a), s=(c + 1, occ), r=(occ + 1, a))
        elif pt_transition_operator:
            w = dict(q=(ionize_mo, ionize_mo), p=(c + 1, a), s=(c + 1, occ), r=(c + 1, a))
        else:
            w = dict(q=(ionize_mo, ionize_mo), p=(c + 1, a), s=(c + 1, occ), r=(occ + 1, a))
    if mode == "MP2-PT2":
        if ionize_mo == 0:
            w = dict(q=(c + 1, occ + 1), p=(c + 1, a), s=(c + 1, occ), r=(occ + 1, a))
        elif pt_transition_operator:
            w = dict(q=(c + 1, max(ionize_mo, occ)), p=(c + 1, a), s=(c + 1, occ), r=(c + 1, a))
        else:
            w = dict(q=(c + 1, max(ionize_mo, occ)), p=(c + 1, a), s=(c + 1, occ), r=(occ + 1, a))
    return [*w["p"], *w["q"], *w["r"], *w["s"]]


def _ionize_flags(name_a, name_b, ionize_species):
    ia = any(name_a == s.strip() for s in ionize_species)
    ib = any(name_b == s.strip() for s in ionize_species)
    return ia, ib


def windows_c_inter(mode, na, nb, occ_a, occ_b, core_a=0, core_b=0, active_a=0, active_b=0, ionize_mo=0,
                    ionize_species=("NONE",), name_a="A", name_b="B"):
    """TransformIntegralsC.f90:1625-1963 -> (win[8], symmetric)."""
    ca, cb = core_a, core_b
    aa = active_a if active_a != 0 else na
    ab = active_b if active_b != 0 else nb
    sym = True
    p, q, r, s = (ca + 1, aa), (ca + 1, aa), (cb + 1, ab), (cb + 1, ab)
    if mode == "ALL":
        p, q, r, s = (1, na), (1, na), (1, nb), (1, nb)
    if mode == "ALLACTIVE":
        p, q, r, s = (1, aa), (1, aa), (1, ab), (1, ab)
    if mode == "MP2":
        p, q, r, s = (ca + 1, occ_a), (occ_a + 1, aa), (cb + 1, occ_b), (occ_b + 1, ab)
    if mode in ("PT2", "MP2-PT2"):
        p, q, r, s = (ca + 1, occ_a + 1), (ca + 1, aa), (cb + 1, occ_b + 1), (cb + 1, ab)
        sym = True
        if ionize_species[0] != "NONE":
            sym = False
            iA, iB = _ionize_flags(name_a, name_b, ionize_species)
            if ionize_mo == 0:
                if iA and iB:
                    p, q, r, s = (ca + 1, occ_a + 1), (ca + 1, aa), (cb + 1, occ_b + 1), (cb + 1, ab)
                elif iA and not iB:
                    p, q, r, s = (ca + 1, occ_a + 1), (ca + 1, aa), (cb + 1, occ_b), (occ_b + 1, ab)
                elif iB and not iA:
                    p, q, r, s = (ca + 1, occ_a), (occ_a + 1, aa), (cb + 1, occ_b + 1), (cb + 1, ab)
            else:
                if iA and iB:
                    if ionize_mo <= occ_a and ionize_mo <= occ_b:
                        p, q, r, s = (ca + 1, occ_a), (ca + 1, aa), (cb + 1, occ_b), (cb + 1, ab)
                    elif ionize_mo > occ_a and ionize_mo > occ_b:
                        p, q, r, s = (ca + 1, ionize_mo), (ca + 1, aa), (cb + 1, ionize_mo), (cb + 1, ab)
                elif iA and not iB:
                    if mode == "PT2":
                        p = (ionize_mo, ionize_mo)
                    else:
                        p = (ca + 1, max(ionize_mo, occ_a))
                    q, r, s = (ca + 1, aa), (cb + 1, occ_b), (occ_b + 1, ab)
                elif iB and not iA:
                    p, q = (ca + 1, occ_a), (occ_a + 1, aa)
                    if mode == "PT2":
                        r = (ionize_mo, ionize_mo)
                    else:
                        r = (cb + 1, max(ionize_mo, occ_b))
                    s = (cb + 1, ab)
    return [*p, *q, *r, *s], sym


def windows_e_inter(mode, na, nb, occ_a, occ_b, core_a=0, core_b=0, active_a=0, active_b=0, ionize_mo=0,
                    ionize_species=("NONE",), name_a="A", name_b="B", pt_transition_operator=False):
    """TransformIntegralsE.f90:2076-2418 -> win[8].  Quirks kept: PT2 default uses the FIRST
    species' core for s_l and r_l (E.f90:2142-2145); MP2-PT2 default sets r_u to the first
    species' active count (E.f90:2283)."""
    ca, cb = core_a, core_b
    aa = active_a if active_a != 0 else na
    ab = active_b if active_b != 0 else nb
    p, q, r, s = (ca + 1, aa), (ca + 1, aa), (cb + 1, ab), (cb + 1, ab)
    if mode == "ALLACTIVE":
        p, q, r, s = (1, aa), (1, aa), (1, ab), (1, ab)
    if mode == "MP2":
        p, q, r, s = (occ_a + 1, aa), (ca + 1, occ_a), (occ_b + 1, ab), (cb + 1, occ_b)
    if mode in ("PT2", "MP2-PT2"):
        if mode == "PT2":
            q, p, s, r = (ca + 1, occ_a + 1), (ca + 1, aa), (ca + 1, occ_b + 1), (ca + 1, ab)
        else:
            q, p, s, r = (ca + 1, occ_a + 1), (ca + 1, aa), (cb + 1, occ_b + 1), (cb + 1, aa)
        if ionize_species[0] != "NONE":
            iA, iB = _ionize_flags(name_a, name_b, ionize_species)
            if ionize_mo == 0:
                if iA and iB:
                    q, p, s, r = (ca + 1, occ_a + 1), (ca + 1, aa), (cb + 1, occ_b + 1), (cb + 1, ab)
                elif iA and not iB:
                    q, p, s, r = (ca + 1, occ_a + 1), (ca + 1, aa), (cb + 1, occ_b), (occ_b + 1, ab)
                elif iB and not iA:
                    q, p, s, r = (ca + 1, occ_a), (occ_a + 1, aa), (cb + 1, occ_b + 1), (cb + 1, ab)
            else:
                if iA and iB:
                    if ionize_mo <= occ_a and ionize_mo <= occ_b:
                        q, p, s, r = (ca + 1, occ_a), (ca + 1, aa), (cb + 1, occ_b), (cb + 1, ab)
                    elif ionize_mo > occ_a and ionize_mo > occ_b:
                        q, p, s, r = (ca + 1, aa), (ca + 1, aa), (cb + 1, ab), (cb + 1, ab)
                elif iA and not iB:
                    if mode == "PT2":
                        q = (ionize_mo, ionize_mo)
                    else:
                        q = (ca + 1, max(ionize_mo, occ_a))
                    if pt_transition_operator:
                        q = (ca + 1, aa)
                    p, s, r = (ca + 1, aa), (cb + 1, occ_b), (occ_b + 1, ab)
                elif iB and not iA:
                    q, p, s, r = (ca + 1, occ_a), (occ_a + 1, aa), (cb + 1, ab), (cb + 1, ab)
    return [*p, *q, *r, *s]


# ----------------------------------------------------------------------------------------
# downstream energies
# ----------------------------------------------------------------------------------------
def mp2_intra_from_pairs(ij, kl, v, n, occ, eps, lam=2.0, frozen=0, active=0):
    M = npairs(n)
    packed = np.zeros(M * (M + 1) // 2)
    lib().orc_reader_pairs_intra(np.ascontiguousarray(ij, np.int64), np.ascontiguousarray(kl, np.int64),
                                 np.ascontiguousarray(v), len(v), n, packed)
    return lib().orc_mp2_intra(packed, n, occ, frozen, active or n, lam, np.ascontiguousarray(eps))


def mp2_intra_from_quads(p, q, r, s, v, n, occ, eps, lam=2.0, frozen=0, active=0):
    M = npairs(n)
    packed = np.zeros(M * (M + 1) // 2)
    lib().orc_reader_quads_intra(p, q, r, s, np.ascontiguousarray(v), len(v), n, packed)
    return lib().orc_mp2_intra(packed, n, occ, frozen, active or n, lam, np.ascontiguousarray(eps))


def mp2_inter_from_pairs(ij, kl, v, na, nb, occ_a, occ_b, eps_a, eps_b, charge_a=-1.0, charge_b=1.0,
                         lam_a=2.0, lam_b=1.0):
    rect = np.zeros(npairs(na) * npairs(nb))
    lib().orc_reader_pairs_inter(np.ascontiguousarray(ij, np.int64), np.ascontiguousarray(kl, np.int64),
                                 np.ascontiguousarray(v), len(v), na, nb, rect)
    return lib().orc_mp2_inter(rect, na, nb, occ_a, occ_b, 0, 0, na, nb, charge_a, charge_b, lam_a, lam_b,
                               np.ascontiguousarray(eps_a), np.ascontiguousarray(eps_b))


def mp2_inter_from_quads(p, q, r, s, v, na, nb, occ_a, occ_b, eps_a, eps_b, charge_a=-1.0, charge_b=1.0,
                         lam_a=2.0, lam_b=1.0):
    rect = np.zeros(npairs(na) * npairs(nb))
    lib().orc_reader_quads_inter(p, q, r, s, np.ascontiguousarray(v), len(v), na, nb, rect)
    return lib().orc_mp2_inter(rect, na, nb, occ_a, occ_b, 0, 0, na, nb, charge_a, charge_b, lam_a, lam_b,
                               np.ascontiguousarray(eps_a), np.ascontiguousarray(eps_b))


def e_first_half_sample(seed, Cm, win, pq0, npq, nthreads=1):
    """CPU baseline sample: first half of transformer E on npq slabs of the kind-H tensor."""
    n = Cm.shape[0]
    return lib().orc_e_first_half_sample(seed, n, np.asfortranarray(Cm), n, _win(win), pq0, npq, nthreads)


def e_first_half_values(seed, Cm, win, pq0, npq, nthreads=1, gen_kind=1):
    """First half of transformer E (E.f90:1043-1132) on slabs [pq0, pq0+npq) of the kind-H/F tensor: array [n_ij, npq]
    in ijmap order, |t| <= 1e-10 zeroed (E.f90:1113)."""
    n = Cm.shape[0]
    Cf = np.asfortranarray(Cm)
    nij = lib().orc_e_first_half_values(gen_kind, seed, n, Cf, n, _win(win), pq0, npq, nthreads, None)
    out = np.zeros((nij, npq))
    lib().orc_e_first_half_values(gen_kind, seed, n, Cf, n, _win(win), pq0, npq, nthreads, out.ctypes.data)
    return out


def direct_first_quarter(coef, p, q, r, s, v):
    """The DIRECT first quarter of the reference (Libint2Iface.cpp:793-868) fed from a canonical list:
    GG[nu, lam, sig] = sum_mu (mu nu|lam sig) coef[mu]."""
    n = len(coef)
    GG = np.zeros((n, n, n))
    lib().orc_direct_first_quarter(n, np.ascontiguousarray(coef, dtype=np.float64), *[np.ascontiguousarray(x, dtype=np.int32) for x in (p, q, r, s)],
                                   np.ascontiguousarray(v, dtype=np.float64), len(v), GG)
    return GG


def rankk_factors(seed, n, K=8):
    """The K symmetric N x N factor matrices of a kind-K tensor and their pair vectors [K, M] in xy order (mu <= nu)."""
    rng = np.random.default_rng(seed)
    L = rng.uniform(-1, 1, (K, n, n))
    L = 0.5 * (L + L.transpose(0, 2, 1))
    iu = np.triu_indices(n)
    return L, np.ascontiguousarray(L[:, iu[0], iu[1]])


def rankk_mo_factors(L, Cm, rows, cols):
    """T^k[p,q] = (C^T L^k C)[rows, cols]: the closed-form factors of the MO integrals of a kind-K tensor, [K, len(rows), len(cols)]."""
    Cm = np.asarray(Cm)
    return np.stack([Cm[:, rows].T @ (Lk @ Cm[:, cols]) for Lk in L])


# ----------------------------------------------------------------------------------------
# large-N checks: vectorised AO list of a dense pair matrix, closed-form MO integrals of a rank-K tensor
# ----------------------------------------------------------------------------------------
def list_from_pair_matrix_intra(sq, n, rows_per_block=256):
    """Every unique (pq|rs), pq<=rs in xy numbering, of a symmetric [M,M] pair matrix as 1-based AO quartets
    (the loader is order independent, C.f90:262-273).  Yields blocks (p,q,r,s,v) to bound memory."""
    M = npairs(n)
    i1, i2 = np.triu_indices(n)
    for a0 in range(0, M, rows_per_block):
        a1 = min(M, a0 + rows_per_block)
        a = np.repeat(np.arange(a0, a1), M)
        b = np.tile(np.arange(M), a1 - a0)
        keep = b >= a
        a, b = a[keep], b[keep]
        yield ((i2[a] + 1).astype(np.int32), (i1[a] + 1).astype(np.int32), (i2[b] + 1).astype(np.int32),
               (i1[b] + 1).astype(np.int32), np.ascontiguousarray(sq[a, b]))


def rankk_mo_block(La, Ca, rows_a, cols_a, Lb=None, Cb=None, rows_b=None, cols_b=None):
    """Closed form of the MO integrals of (mu nu|lam sig) = sum_k La^k[mu,nu] Lb^k[lam,sig]:
    (p q|r s) = sum_k (Ca^T La^k Ca)[p,q] (Cb^T Lb^k Cb)[r,s] for p in rows_a, q in cols_a, r in rows_b, s in cols_b
    (0-based orbital index arrays) -> array [len(rows_a), len(cols_a), len(rows_b), len(cols_b)].  O(K N^3)."""
    Lb = La if Lb is None else Lb
    Cb = Ca if Cb is None else Cb
    rows_b = rows_a if rows_b is None else rows_b
    cols_b = cols_a if cols_b is None else cols_b
    Ta = np.einsum("mp,kmn,nq->kpq", np.asarray(Ca)[:, rows_a], La, np.asarray(Ca)[:, cols_a], optimize=True)
    Tb = np.einsum("mp,kmn,nq->kpq", np.asarray(Cb)[:, rows_b], Lb, np.asarray(Cb)[:, cols_b], optimize=True)
    return np.einsum("kpq,krs->pqrs", Ta, Tb, optimize=True)


# ----------------------------------------------------------------------------------------
# second-order propagator (P2) poles: the consumer of the PT2-window MO integrals
# ----------------------------------------------------------------------------------------
def read_pairs_intra(ij, kl, v, n):
    """ReadTransformedIntegrals_readOneSpecies, method E layout (ReadTransformedIntegrals.f90:302-311): packed array
    addressed by IndexMapAA (PropagatorTheory.f90:178-193) minus one."""
    M = npairs(n)
    packed = np.zeros(M * (M + 1) // 2)
    lib().orc_reader_pairs_intra(np.ascontiguousarray(ij, np.int64), np.ascontiguousarray(kl, np.int64),
                                 np.ascontiguousarray(v), len(v), n, packed)
    return packed


def read_pairs_inter(ij, kl, v, na, nb, reversed_pair=False):
    """ReadTransformedIntegrals_readTwoSpecies, method E layout.  The file holds (A A|B B) pair ids of the pair in
    molecular-system order; reversed_pair=True is the reader's branch for a caller whose first species is B
    (ReadTransformedIntegrals.f90:911-922): the result is then addressed M_a*(pair_b-1)+pair_a."""
    Ma, Mb = npairs(na), npairs(nb)
    rect = np.zeros(Ma * Mb)
    lib().orc_reader_pairs_inter(np.ascontiguousarray(ij, np.int64), np.ascontiguousarray(kl, np.int64),
                                 np.ascontiguousarray(v), len(v), na, nb, rect)
    return np.ascontiguousarray(rect.reshape(Ma, Mb).T).reshape(-1) if reversed_pair else rect


def p2_poles(a, species, aux, ionize_mo=0, factor_ss=1.0, factor_os=1.0, max_iter=50):
    """PropagatorTheory_secondOrderCorrection (src/PT/PropagatorTheory.f90:459-1177) for species index `a`, without the
    transition-operator branch: numerators / denominators of the 2ph and 2hp terms (:736-811 intra, :848-915 inter),
    Newton-Raphson pole search from Koopmans' value until |delta omega| <= 1e-4 (:971-1073), pole strength (:1075).

    species[j] = dict(name, n, occ, charge, lam, eps[, active]); aux[j] = MO integrals as the reader leaves them BEFORE the
    charge scaling of :664-680: j == a the packed intra array (read_pairs_intra), j != a the rectangular array addressed
    M_j*(pair_a-1)+pair_j (read_pairs_inter).  Returns [(orbital, koopmans, omega, pole_strength, iterations)] in Hartree for
    orbital = ionize_mo, or HOMO and LUMO when ionize_mo == 0 (:603-621).
    The reference's loop condition `residual>1e-4 .or. limit<ni` (:971) never ends once 50 iterations are exceeded; here
    the search stops and raises instead."""
    A = species[a]
    na, oa, la = A["n"], A["occ"], float(A["lam"])
    acta = A.get("active") or na
    ea = np.asarray(A["eps"], dtype=np.float64)
    xya = pair_table(na) + 1                                   # 1-based pair ids, PropagatorTheory.f90:153-160
    Ma = npairs(na)
    if ionize_mo:
        orbitals = [ionize_mo]
    else:
        orbitals = list(range(oa if oa else 1, oa + 2))
    occs = np.arange(1, oa + 1)
    virs = np.arange(oa + 1, acta + 1)
    out = []
    for pa in orbitals:
        terms = []                                             # (species j, numerators, denominators) for 2ph then 2hp
        for j, B in enumerate(species):
            if j == a:
                vals = np.asarray(aux[j]) * A["charge"] * A["charge"]

                def g(i1, i2, i3, i4):                          # IndexMapAA: ioff(min)+max over pair ids (:178-193)
                    p1, p2 = xya[i1 - 1, i2 - 1], xya[i3 - 1, i4 - 1]
                    lo, hi = np.minimum(p1, p2), np.maximum(p1, p2)
                    return vals[(lo - 1) * Ma - (lo - 1) * lo // 2 + hi - 1]
                ia, aa, ba = np.meshgrid(occs, virs, virs, indexing="ij")                  # :736-758
                va, vb = g(pa, aa, ia, ba), g(pa, ba, ia, aa)
                terms.append((j, "2ph", (va * (la * va - vb)).ravel(), (ea[ia - 1] - ea[aa - 1] - ea[ba - 1]).ravel()))
                if oa > 1:                                                                   # :760-795
                    aa, ia, ja = np.meshgrid(virs, occs, occs, indexing="ij")
                    va, vb = g(pa, ia, ja, aa), g(pa, ja, ia, aa)
                    terms.append((j, "2hp", (va * (la * va - vb)).ravel(), (ea[aa - 1] - ea[ia - 1] - ea[ja - 1]).ravel()))
            else:
                nb, ob, lb = B["n"], B["occ"], float(B["lam"])
                actb = B.get("active") or nb
                eb = np.asarray(B["eps"], dtype=np.float64)
                xyb = pair_table(nb) + 1
                Mb = npairs(nb)
                vals = np.asarray(aux[j]) * A["charge"] * B["charge"]

                def gab(i1, i2, i3, i4):                        # IndexMapAB: (ij-1)*M_b+kl (:195-212)
                    return vals[(xya[i1 - 1, i2 - 1] - 1) * Mb + xyb[i3 - 1, i4 - 1] - 1]
                occb, virb = np.arange(1, ob + 1), np.arange(ob + 1, actb + 1)
                ib, aa, ab = np.meshgrid(occb, virs, virb, indexing="ij")                  # diagram A, :866-884
                va = gab(pa, aa, ib, ab)
                terms.append((j, "2ph", (la * lb * va ** 2).ravel(), (eb[ib - 1] - ea[aa - 1] - eb[ab - 1]).ravel()))
                ab, ia, ib = np.meshgrid(virb, occs, occb, indexing="ij")                  # diagram B, :888-915
                va = gab(pa, ia, ib, ab)
                terms.append((j, "2hp", (la * lb * va ** 2).ravel(), (eb[ab - 1] - ea[ia - 1] - eb[ib - 1]).ravel()))
        koop = ea[pa - 1]
        same_spin = A["name"] in ("E-ALPHA", "E-BETA")
        omega, it, residual = koop, 0, 1.0
        while residual > 1.0e-4:                                                             # :971-1071
            it += 1
            if it > max_iter:
                raise RuntimeError("P2 pole search did not converge")
            last = omega
            sigma, dsigma = last - koop, 1.0
            for j, _, num, den in terms:
                b = den + last
                e, de = np.sum(num / b), np.sum(num / b ** 2)
                nameb = species[j]["name"]
                f = 1.0
                if same_spin and j == a:
                    f = factor_ss
                elif {A["name"], nameb} == {"E-ALPHA", "E-BETA"}:
                    f = factor_os
                sigma -= f * e
                dsigma += f * de
            omega = last - sigma / dsigma
            residual = abs(omega - last)
        out.append((pa, koop, omega, 1.0 / dsigma, it))
    return out


# ---- row f4: two-particle AO integrals over contracted Cartesian Gaussians (eri_oracle.c; Libint2Iface.cpp:83-130, :219-416) ----
class OrcShell(C.Structure):
    _fields_ = [("l", C.c_int), ("nprim", C.c_int), ("first_prim", C.c_int), ("origin", C.c_double * 3)]


def _eri_pack(shells):
    """[(l, origin(3), exponents, coefficients), ...] -> (OrcShell array, exponents, coefficients, number of Cartesian functions)"""
    arr = (OrcShell * len(shells))()
    ex, co, nbf = [], [], 0
    for i, (l, origin, e, c) in enumerate(shells):
        arr[i] = OrcShell(l, len(e), len(ex), (C.c_double * 3)(*origin))
        ex += list(e); co += list(c)
        nbf += (l + 1) * (l + 2) // 2
    return arr, np.array(ex, dtype=np.float64), np.array(co, dtype=np.float64), nbf


def _eri_lib():
    L = lib()
    if not getattr(L, "_eri_bound", False):
        PS = C.POINTER(OrcShell)
        L.orc_eri_norma.argtypes = [C.c_int, PS, _f64p, _f64p, _f64p]
        L.orc_eri_packed_intra.argtypes = [C.c_int, PS, _f64p, _f64p, _f64p]
        L.orc_eri_rect_inter.argtypes = [C.c_int, PS, _f64p, _f64p, C.c_int, PS, _f64p, _f64p, _f64p]
        L.orc_eri_one.argtypes = [C.c_int, PS, _f64p, _f64p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_eri_one.restype = C.c_double
        L._eri_bound = True
    return L


def eri_norma(shells):
    arr, ex, co, nbf = _eri_pack(shells)
    out = np.zeros(nbf)
    _eri_lib().orc_eri_norma(len(shells), arr, ex, co, out)
    return out


def eri_packed_intra(shells):
    """The packed AO tensor the transformer stores (row lo: hi = lo..M-1), |raw| <= 1e-10 dropped, values * norma^4."""
    arr, ex, co, nbf = _eri_pack(shells)
    M = nbf * (nbf + 1) // 2
    out = np.zeros(M * (M + 1) // 2)
    _eri_lib().orc_eri_packed_intra(len(shells), arr, ex, co, out)
    return out


def eri_rect_inter(shells_a, shells_b):
    aa, ea, ca, na = _eri_pack(shells_a)
    ab, eb, cb, nb = _eri_pack(shells_b)
    out = np.zeros((nb * (nb + 1) // 2, na * (na + 1) // 2))
    _eri_lib().orc_eri_rect_inter(len(shells_a), aa, ea, ca, len(shells_b), ab, eb, cb, out)
    return out


def eri_one(shells, i, j, k, l):
    arr, ex, co, _ = _eri_pack(shells)
    return _eri_lib().orc_eri_one(len(shells), arr, ex, co, i, j, k, l)


def sto3g_1s(zeta, origin):
    """STO-3G 1s shell of Slater exponent zeta (Hehre, Stewart & Pople 1969): exponents scale with zeta^2."""
    return (0, tuple(origin), [2.227660584 * zeta ** 2, 0.405771156 * zeta ** 2, 0.109818 * zeta ** 2], [0.154328967, 0.535328142, 0.444634542])
