"""Row f4 checker (oracle/eri_oracle.c) pinned WITHOUT libint2: textbook values, derivative identities of the s-type integral,
permutational symmetry, translation / rotation invariance.  CPU only."""
import numpy as np
import pytest

from eri_cases import SZABO_H2, h2_sto3g, nbf, water_like


def test_szabo_ostlund_h2_sto3g(O):
    sh = h2_sto3g(O)
    for q, ref in SZABO_H2.items():
        assert abs(O.eri_one(sh, *q) - ref) < 1e-4, (q, O.eri_one(sh, *q), ref)   # four printed decimals
    # HeH+ (ibid. section 3.5.3): zeta(He) = 2.0925, zeta(H) = 1.24, R = 1.4632
    sh = [O.sto3g_1s(2.0925, (0, 0, 0)), O.sto3g_1s(1.24, (0, 0, 1.4632))]
    for q, ref in {(0, 0, 0, 0): 1.3072, (1, 0, 0, 0): 0.4373, (1, 0, 1, 0): 0.1773, (1, 1, 0, 0): 0.6057, (1, 1, 1, 0): 0.3118,
                   (1, 1, 1, 1): 0.7746}.items():
        assert abs(O.eri_one(sh, *q) - ref) < 1e-4, (q, O.eri_one(sh, *q), ref)   # four printed decimals


def test_norma_makes_unit_self_overlap(O):
    # single normalised primitives: libint2's renorm already normalises the (l,0,0) component, the others need (2l-1)!!-type factors
    n = O.eri_norma([(2, (0, 0, 0), [0.7], [1.0])])
    assert np.allclose(n, [1, np.sqrt(3), np.sqrt(3), 1, np.sqrt(3), 1], rtol=1e-13)   # xx xy xz yy yz zz
    n = O.eri_norma([(3, (0, 0, 0), [0.7], [1.0])])
    assert np.allclose(n[[0, 1, 4]], [1, np.sqrt(5), np.sqrt(15)], rtol=1e-13)         # xxx xxy xyz


def _s(alpha, A):
    return (0, tuple(A), [alpha], [1.0])


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_p_functions_are_centre_derivatives_of_s(O, axis):
    """x exp(-a r_A^2) = (1/2a) d/dA_x exp(-a r_A^2)  =>  (p_x b|c d) = a^(-1/2) d/dA_x (s_A b|c d) for normalised functions."""
    a, A = 0.9, np.array([0.1, -0.2, 0.3])
    rest = [_s(1.3, (0.7, 0.2, -0.4)), (1, (-0.5, 0.6, 0.1), [0.6, 1.9], [0.4, 0.7]), _s(0.5, (0.0, -0.9, 0.8))]
    got = O.eri_one([(1, tuple(A), [a], [1.0])] + rest, axis, 3, 4 + 1, 7)       # (p_axis s | p_y s)

    def f(h):
        Ap = A.copy(); Ap[axis] += h
        Am = A.copy(); Am[axis] -= h
        return (O.eri_one([_s(a, Ap)] + rest, 0, 1, 2 + 1, 5) - O.eri_one([_s(a, Am)] + rest, 0, 1, 2 + 1, 5)) / (2 * h)
    d = (4 * f(5e-4) - f(1e-3)) / 3   # Richardson
    assert abs(got - d / np.sqrt(a)) < 2e-9, (got, d / np.sqrt(a))


def test_d_functions_are_second_derivatives_of_s(O):
    a, A = 1.1, np.array([0.2, 0.1, -0.3])
    rest = [_s(0.8, (0.9, -0.2, 0.4)), _s(1.7, (-0.4, 0.5, 0.2)), _s(0.6, (0.1, 0.8, -0.7))]
    d_sh = [(2, tuple(A), [a], [1.0])] + rest           # functions: xx xy xz yy yz zz s s s
    s0 = O.eri_one([_s(a, A)] + rest, 0, 1, 2, 3)
    h = 2e-3

    def s_at(dx, dy):
        return O.eri_one([_s(a, A + np.array([dx, dy, 0.0]))] + rest, 0, 1, 2, 3)
    dxx = (s_at(h, 0) - 2 * s0 + s_at(-h, 0)) / h ** 2
    dxy = (s_at(h, h) - s_at(h, -h) - s_at(-h, h) + s_at(-h, -h)) / (4 * h * h)
    assert abs(O.eri_one(d_sh, 0, 6, 7, 8) - (dxx + 2 * a * s0) / (np.sqrt(3) * a)) < 2e-6     # (d_xx s|s s)
    assert abs(O.eri_one(d_sh, 1, 6, 7, 8) - dxy / a) < 2e-6                                   # (d_xy s|s s)


def test_permutational_symmetry_and_packed_layout(O):
    sh = water_like()
    n = nbf(sh)
    rng = np.random.default_rng(3)
    for _ in range(40):
        i, j, k, l = rng.integers(0, n, 4)
        v = O.eri_one(sh, i, j, k, l)
        for q in ((j, i, k, l), (i, j, l, k), (k, l, i, j), (l, k, j, i)):
            assert abs(O.eri_one(sh, *q) - v) < 1e-13
    small = water_like()[:3]   # s s p on one centre: 5 functions
    packed = O.eri_packed_intra(small)
    M = 15
    pairs = [(i, j) for i in range(5) for j in range(i, 5)]
    for lo in range(M):
        for hi in range(lo, M):
            v = O.eri_one(small, *pairs[lo], *pairs[hi])
            x = packed[lo * M - lo * (lo + 1) // 2 + hi]
            assert x == 0.0 or abs(x - v) < 1e-14
            assert x != 0.0 or abs(v) < 1e-9   # only integrals whose raw value is below the reference's 1e-10 filter are dropped


def test_translation_and_rotation_invariance(O):
    sh = water_like()
    c, s = np.cos(0.7), np.sin(0.7)
    Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    moved = [(l, tuple(Rz @ np.array(o) + np.array([0.3, -1.1, 2.0])), e, co) for l, o, e, co in sh]
    # s-type functions only are invariant one by one: shells 0, 1 (O), 4 (H1), 6 (H2) -> function indices 0, 1, 11, 15
    for q in ((0, 1, 11, 15), (0, 0, 11, 11), (1, 15, 1, 11)):
        assert abs(O.eri_one(sh, *q) - O.eri_one(moved, *q)) < 1e-13
    # a p shell rotates like a vector: sum over components of (p_i s|p_i s) is invariant
    tot = sum(O.eri_one(sh, 2 + i, 0, 2 + i, 11) for i in range(3))
    tot_m = sum(O.eri_one(moved, 2 + i, 0, 2 + i, 11) for i in range(3))
    assert abs(tot - tot_m) < 1e-13


def _mp_reference(prims):
    """(ab|cd) over NORMALISED primitive Cartesian Gaussians from 50-digit derivatives of the closed-form (ss|ss) integral:
    x^n exp(-a x^2) is a combination of centre derivatives of exp(-a x^2) (n = 1: d/2a; n = 2: d2/4a^2 + 1/2a; n = 3: d3/8a^3 +
    3 d/4a^2), so any class up to f follows from mixed partials of one analytic function -- no recurrences, no quadrature.
    prims: four tuples (exponent, centre(3), (lx, ly, lz))."""
    import itertools
    import mpmath as mp
    mp.mp.dps = 50
    expo = [mp.mpf(p[0]) for p in prims]

    def ssss(*c):   # 12 coordinates: A, B, C, D
        A, B, Cc, D = (c[0:3], c[3:6], c[6:9], c[9:12])
        a, b, cc, d = expo
        p, q = a + b, cc + d
        AB2 = sum((x - y) ** 2 for x, y in zip(A, B)); CD2 = sum((x - y) ** 2 for x, y in zip(Cc, D))
        P = [(a * x + b * y) / p for x, y in zip(A, B)]; Q = [(cc * x + d * y) / q for x, y in zip(Cc, D)]
        xx = p * q / (p + q) * sum((x - y) ** 2 for x, y in zip(P, Q))
        f0 = mp.sqrt(mp.pi / xx) / 2 * mp.erf(mp.sqrt(xx)) if xx > mp.mpf("1e-30") else mp.mpf(1)
        return 2 * mp.pi ** 2.5 / (p * q * mp.sqrt(p + q)) * mp.exp(-a * b / p * AB2 - cc * d / q * CD2) * f0

    def terms(n, a):   # x^n g = sum_k coef_k d^k g / dA^k
        return {0: [(0, mp.mpf(1))], 1: [(1, 1 / (2 * a))], 2: [(2, 1 / (4 * a * a)), (0, 1 / (2 * a))],
                3: [(3, 1 / (8 * a ** 3)), (1, 3 / (4 * a * a))]}[n]
    per_dim, point = [], []
    for (a, centre, l), am in zip(prims, expo):
        for d in range(3):
            per_dim.append(terms(l[d], am)); point.append(mp.mpf(centre[d]))
    total = mp.mpf(0)
    for choice in itertools.product(*per_dim):
        orders = tuple(k for k, _ in choice)
        coef = mp.mpf(1)
        for _, c in choice:
            coef *= c
        total += coef * (mp.diff(ssss, tuple(point), orders) if any(orders) else ssss(*point))

    def dfact(n):
        r = 1
        while n > 1:
            r *= n; n -= 2
        return r
    norm = mp.mpf(1)
    for (a, _, l), am in zip(prims, expo):
        L = sum(l)
        norm *= (2 * am / mp.pi) ** mp.mpf("0.75") * (4 * am) ** (mp.mpf(L) / 2) / mp.sqrt(dfact(2 * l[0] - 1) * dfact(2 * l[1] - 1) * dfact(2 * l[2] - 1))
    return float(total * norm)


def _fn_index(l, comp):
    """index of Cartesian component comp = (lx, ly, lz) inside a shell of angular momentum l (libint2 / CCA order)"""
    k = 0
    for lx in range(l, -1, -1):
        for ly in range(l - lx, -1, -1):
            if (lx, ly, l - lx - ly) == tuple(comp):
                return k
            k += 1
    raise ValueError(comp)


@pytest.mark.parametrize("case", [
    ((3, (1, 1, 1)), (1, (0, 1, 0)), (2, (1, 0, 1)), (0, (0, 0, 0))),      # (f_xyz p_y | d_xz s)
    ((3, (3, 0, 0)), (0, (0, 0, 0)), (0, (0, 0, 0)), (0, (0, 0, 0))),      # (f_xxx s | s s)
    ((2, (2, 0, 0)), (2, (0, 2, 0)), (1, (0, 0, 1)), (1, (0, 0, 1))),      # (d_xx d_yy | p_z p_z)
    ((3, (2, 1, 0)), (0, (0, 0, 0)), (3, (0, 1, 2)), (0, (0, 0, 0))),      # (f_xxy s | f_yzz s)
    ((1, (1, 0, 0)), (2, (0, 1, 1)), (3, (0, 3, 0)), (0, (0, 0, 0))),      # (p_x d_yz | f_yyy s)
])
def test_high_angular_momentum_classes_against_analytic_derivatives(O, case):
    """Every class up to f pinned WITHOUT any integral library: the oracle (Rys-form) against 50-digit mixed derivatives of the
    closed-form (ss|ss) integral (mpmath), single primitives on four different centres.  (Checked once in the same way, too slow
    to keep in the suite: (f_xyy f_yzz|s d_xy) and (p_x p_x|f_xxz f_yyy), 3 and 1 minutes of 50-digit differentiation.)"""
    pytest.importorskip("mpmath")
    expo = (0.9, 1.3, 0.7, 1.6)
    centres = ((0.1, -0.2, 0.3), (0.7, 0.2, -0.4), (-0.5, 0.6, 0.1), (0.0, -0.9, 0.8))
    shells, idx, off = [], [], 0
    for (l, comp), a, c in zip(case, expo, centres):
        shells.append((l, c, [a], [1.0]))
        idx.append(off + _fn_index(l, comp))
        off += (l + 1) * (l + 2) // 2
    got = O.eri_one(shells, *idx)
    ref = _mp_reference([(a, c, comp) for (l, comp), a, c in zip(case, expo, centres)])
    assert abs(got - ref) <= 1e-12 * max(1.0, abs(ref)), (got, ref)
    assert abs(ref) > 1e-6     # a non-trivial value


def test_h2_minimal_basis_chain_reproduces_the_textbook(O):
    """basis -> AO integrals -> MO integrals -> MP2 energy on the CHECKER's side, against Szabo & Ostlund's minimal-basis H2:
    the one real-molecule chain whose every number is published (the reference's own test energies need libint2 + an SCF)."""
    from eri_cases import H2_MO, h2_mo_coefficients, sto3g_overlap
    sh = h2_sto3g(O)
    S = sto3g_overlap(sh[0], 1.4)
    assert abs(S - H2_MO["S12"]) < 1e-4
    Cm = h2_mo_coefficients(S)
    packed = O.eri_packed_intra(sh)
    win, sym = O.windows_c_intra("ALL", 2, 1)
    p, q, r, s, v = O.transform_c_intra(Cm, packed, win, sym)
    mo = {(a, b, c, d): x for a, b, c, d, x in zip(p, q, r, s, v)}
    for key, name in (((1, 1, 1, 1), "J11"), ((1, 1, 2, 2), "J12"), ((2, 2, 2, 2), "J22"), ((1, 2, 1, 2), "K12")):
        assert abs(mo[key] - H2_MO[name]) < 1e-4, (key, mo[key])
    assert (1, 1, 1, 2) not in mo and (1, 2, 2, 2) not in mo          # zero by symmetry: below the 1e-10 output filter
    ij, kl, vv = O.transform_e_intra(Cm, packed, O.windows_e_intra("MP2", 2, 1))
    assert len(vv) == 1 and abs(vv[0] - H2_MO["K12"]) < 1e-4
    e2 = O.mp2_intra_from_pairs(ij, kl, vv, 2, 1, np.array(H2_MO["eps"]))
    assert abs(e2 - H2_MO["K12"] ** 2 / (2 * (H2_MO["eps"][0] - H2_MO["eps"][1]))) < 2e-5 and abs(e2 + 0.0132) < 1e-4
