"""Basis sets of the f4 (AO-integral producer) tests: [(l, origin, exponents, coefficients), ...] as LibintInterface::add_shell takes them."""
import numpy as np


def h2_sto3g(O, R=1.4):
    """Szabo & Ostlund, Modern Quantum Chemistry, section 3.5.2: H2, R = 1.4 a0, STO-3G with zeta = 1.24."""
    return [O.sto3g_1s(1.24, (0, 0, 0)), O.sto3g_1s(1.24, (0, 0, R))]


SZABO_H2 = {(0, 0, 0, 0): 0.7746, (0, 0, 1, 1): 0.5697, (1, 0, 0, 0): 0.4441, (1, 0, 1, 0): 0.2970}


def water_like(with_f=False):
    """A bent triatomic with contracted s and p shells, a d shell on the heavy atom, optionally an f shell: every
    angular-momentum class up to (dd|dd) / (ff|ff) appears, on three centres, with primitives of very different exponents."""
    Oc, H1, H2 = (0.0, 0.0, 0.2217), (0.0, 1.4309, -0.8867), (0.1, -1.4309, -0.8867)
    sh = [
        (0, Oc, [130.70932, 23.808861, 6.4436083], [0.15432897, 0.53532814, 0.44463454]),
        (0, Oc, [5.0331513, 1.1695961, 0.3803890], [-0.09996723, 0.39951283, 0.70011547]),
        (1, Oc, [5.0331513, 1.1695961, 0.3803890], [0.15591627, 0.60768372, 0.39195739]),
        (2, Oc, [1.2], [1.0]),
        (0, H1, [3.42525091, 0.62391373, 0.16885540], [0.15432897, 0.53532814, 0.44463454]),
        (1, H1, [0.8], [1.0]),
        (0, H2, [3.42525091, 0.62391373, 0.16885540], [0.15432897, 0.53532814, 0.44463454]),
    ]
    if with_f:
        sh.append((3, Oc, [0.9, 0.35], [0.6, 0.5]))
    return sh


def nuclear_like():
    """A protonic-style basis (tight s, p, d shells on one centre), the second species of an inter-species pair."""
    Hc = (0.0, 1.4309, -0.8867)
    return [(0, Hc, [24.0, 9.0], [0.5, 0.6]), (1, Hc, [16.0], [1.0]), (2, Hc, [12.0], [1.0])]


def nbf(shells):
    return sum((l + 1) * (l + 2) // 2 for l, *_ in shells)


def packed_to_full(packed, M):
    sq = np.zeros((M, M))
    iu = np.triu_indices(M)
    sq[iu] = packed
    return sq + np.triu(sq, 1).T


# Minimal-basis H2 at R = 1.4 a0 (Szabo & Ostlund, section 3.5.2 and eq. 6.77): overlap, MO two-electron integrals, orbital energies,
# second-order (MP2) correlation energy K12^2 / (2 (eps1 - eps2)).
H2_MO = {"S12": 0.6593, "J11": 0.6746, "J12": 0.6636, "J22": 0.6975, "K12": 0.1813, "eps": (-0.5782, 0.6703)}


def h2_mo_coefficients(S):
    """sigma_g, sigma_u of the minimal basis: (phi1 +- phi2) / sqrt(2 (1 +- S))"""
    return np.array([[1 / np.sqrt(2 * (1 + S)), 1 / np.sqrt(2 * (1 - S))], [1 / np.sqrt(2 * (1 + S)), -1 / np.sqrt(2 * (1 - S))]])


def sto3g_overlap(shell, R):
    """<phi1|phi2> of two copies of one contracted s shell a distance R apart, unit-normalised"""
    _, _, e, c = shell
    e, c = np.array(e), np.array(c)
    w = c * (2 * e / np.pi) ** 0.75
    p = e[:, None] + e[None, :]
    s12 = (w[:, None] * w[None, :] * (np.pi / p) ** 1.5 * np.exp(-e[:, None] * e[None, :] / p * R * R)).sum()
    s11 = (w[:, None] * w[None, :] * (np.pi / p) ** 1.5).sum()
    return s12 / s11
