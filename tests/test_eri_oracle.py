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
