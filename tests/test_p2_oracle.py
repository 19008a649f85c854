"""The second-order propagator (P2) consumer restated in the oracle (PropagatorTheory.f90:459-1177): it must find every
integral it needs inside the PT2 window tables of transformer E, agree with a dense numpy.einsum evaluation of the
self-energy, and behave as the reference does for one-particle species (no intra 2hp term)."""
import numpy as np
import pytest

SCALE = 0.05  # keeps the synthetic self-energy a small correction so that the Newton-Raphson search stays near Koopmans


def _system(O, ne=8, npz=5, oe=3, op=1):
    Ce, Cp = O.random_orthonormal(ne, 1), O.random_orthonormal(npz, 2)
    pe, pp = O.hash_packed_intra(5, ne) * SCALE, O.hash_packed_intra(6, npz) * SCALE
    rect = O.hash_rect_inter(7, ne, npz) * SCALE
    sp = [dict(name="E-", n=ne, occ=oe, charge=-1.0, lam=2, eps=O.synthetic_eps(oe, ne)),
          dict(name="POSITRON", n=npz, occ=op, charge=1.0, lam=1, eps=O.synthetic_eps(op, npz))]
    return Ce, Cp, pe, pp, rect, sp


def _aux(O, a, Ce, Cp, pe, pp, rect, sp, we, wp, wi):
    ne, npz = sp[0]["n"], sp[1]["n"]
    ii = O.transform_e_inter(Ce, Cp, rect, wi)
    if a == 0:
        return [O.read_pairs_intra(*O.transform_e_intra(Ce, pe, we), ne), O.read_pairs_inter(*ii, ne, npz)]
    return [O.read_pairs_inter(*ii, ne, npz, reversed_pair=True), O.read_pairs_intra(*O.transform_e_intra(Cp, pp, wp), npz)]


@pytest.mark.parametrize("a,ion", [(0, ("E-",)), (1, ("POSITRON",)), (0, ()), (1, ())])
def test_pt2_windows_hold_every_integral_p2_reads(O, a, ion):
    Ce, Cp, pe, pp, rect, sp = _system(O)
    ne, npz, oe, op = sp[0]["n"], sp[1]["n"], sp[0]["occ"], sp[1]["occ"]
    we, wp = O.windows_e_intra("PT2", ne, oe), O.windows_e_intra("PT2", npz, op)
    wi = O.windows_e_inter("PT2", ne, npz, oe, op, ionize_species=ion or ("NONE",), name_a="E-", name_b="POSITRON")
    got = O.p2_poles(a, sp, _aux(O, a, Ce, Cp, pe, pp, rect, sp, we, wp, wi))
    full = lambda n, m: [1, n, 1, n, 1, m, 1, m]
    ref = O.p2_poles(a, sp, _aux(O, a, Ce, Cp, pe, pp, rect, sp, full(ne, ne), full(npz, npz), full(ne, npz)))
    assert [g[0] for g in got] == [sp[a]["occ"], sp[a]["occ"] + 1]          # HOMO and LUMO (PropagatorTheory.f90:612-620)
    for g, r in zip(got, ref):
        assert g[0] == r[0] and g[4] == r[4]
        assert abs(g[2] - r[2]) <= 1e-12 and abs(g[3] - r[3]) <= 1e-12


def test_p2_pole_satisfies_dyson_equation_with_einsum_integrals(O):
    """omega = eps_p + Sigma_pp(omega) with Sigma built from a dense einsum transform (no reader, no pair addressing)."""
    Ce, Cp, pe, pp, rect, sp = _system(O)
    ne, npz, oe, op = sp[0]["n"], sp[1]["n"], sp[0]["occ"], sp[1]["occ"]
    full = lambda n, m: [1, n, 1, n, 1, m, 1, m]
    poles = O.p2_poles(0, sp, _aux(O, 0, Ce, Cp, pe, pp, rect, sp, full(ne, ne), full(npz, npz), full(ne, npz)))
    xe, xp = O.pair_table(ne), O.pair_table(npz)
    mo_ee = O.einsum_transform(O.dense4_from_square(O.packed_to_square(pe, O.npairs(ne)), xe), Ce)
    mo_ep = -1.0 * O.einsum_transform(O.dense4_from_square(rect.reshape(O.npairs(npz), O.npairs(ne)).T, xe, xp), Ce, Cp)
    ee, ep = sp[0]["eps"], sp[1]["eps"]
    o, v = slice(0, oe), slice(oe, ne)
    ob, vb = slice(0, op), slice(op, npz)
    for pa, koop, omega, strength, _ in poles:
        p = pa - 1
        x = mo_ee[p, v, o, v]                                        # (p a|i b) as [a,i,b]
        s = np.sum(x * (2 * x - x.transpose(2, 1, 0)) / (omega + ee[None, o, None] - ee[v, None, None] - ee[None, None, v]))
        y = mo_ee[p, o, o, v]                                        # (p i|j a) as [i,j,a]
        s += np.sum(y * (2 * y - y.transpose(1, 0, 2)) / (omega + ee[None, None, v] - ee[o, None, None] - ee[None, o, None]))
        z = mo_ep[p, v, ob, vb]                                      # (p a|I A)
        s += np.sum(2 * 1 * z ** 2 / (omega + ep[None, ob, None] - ee[v, None, None] - ep[None, None, vb]))
        w = mo_ep[p, o, ob, vb]                                      # (p i|I A)
        s += np.sum(2 * 1 * w ** 2 / (omega + ep[None, None, vb] - ee[o, None, None] - ep[None, ob, None]))
        # the search stops when a Newton step is below 1e-4 (PropagatorTheory.f90:971): the Dyson residual is of that size
        assert abs(omega - koop - s) <= 2e-4
        assert 0.5 < strength <= 1.0


def test_one_particle_species_has_no_intra_2hp_term(O):
    """occupation 1: the intra 2hp block is skipped (PropagatorTheory.f90:760, :996): doubling only (p i|j a) integrals
    of the positron species must not move its pole."""
    Ce, Cp, pe, pp, rect, sp = _system(O)
    ne, npz = sp[0]["n"], sp[1]["n"]
    full = lambda n, m: [1, n, 1, n, 1, m, 1, m]
    aux = _aux(O, 1, Ce, Cp, pe, pp, rect, sp, full(ne, ne), full(npz, npz), full(ne, npz))
    base = O.p2_poles(1, sp, aux, ionize_mo=1)
    xy = O.pair_table(npz) + 1
    M = O.npairs(npz)
    packed = aux[1].copy()
    for a_ in range(2, npz + 1):                                     # (1 1|1 a): only the 2hp term would read these
        lo, hi = sorted((xy[0, 0], xy[0, a_ - 1]))
        packed[(lo - 1) * M - (lo - 1) * lo // 2 + hi - 1] *= 2.0
    moved = O.p2_poles(1, sp, [aux[0], packed], ionize_mo=1)
    assert base[0][0] == 1 and abs(base[0][2] - moved[0][2]) == 0.0
