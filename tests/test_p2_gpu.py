"""Downstream P2 energies from GPU integrals (BASELINE.json: <= 1e-9 Eh): the second-order propagator poles at the HCN.e+
shapes.  The consumer itself (PropagatorTheory.f90:459-1177, restated in oracle.p2_poles) is tested on CPU in
tests/test_p2_oracle.py."""
import numpy as np
import pytest

import openlowdin_b200 as ol

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("a,ion", [(0, ("E-",)), (1, ("POSITRON",))])
def test_hcn_positron_p2_poles_downstream(O, T, a, ion):
    """Downstream P2 energies (BASELINE: <= 1e-9 Eh) at the HCN.e+ shapes, method E, PT2 windows: the second-order
    propagator poles (PropagatorTheory.f90:459-1177, restated in the oracle) evaluated from the GPU's MO integrals and from
    the oracle's.  The coefficients carry a factor s so that the synthetic self-energy is a small correction
    (MO integrals scale with s^4); the transform itself does not care whether C is orthonormal."""
    ne, npos, oe, op = 53, 30, 7, 1
    s = 2.0e-3 ** 0.25
    Ce, Cp = np.asfortranarray(O.random_orthonormal(ne, 21) * s), np.asfortranarray(O.random_orthonormal(npos, 22) * s)
    T.set_species(0, Ce); T.set_species(1, Cp)
    T.set_generator(0, 0, 5); T.set_generator(0, 1, 6); T.set_generator(1, 1, 7)
    pe, pp, rect = O.hash_packed_intra(5, ne), O.hash_packed_intra(7, npos), O.hash_rect_inter(6, ne, npos)
    we, wp = O.windows_e_intra("PT2", ne, oe), O.windows_e_intra("PT2", npos, op)
    wi = O.windows_e_inter("PT2", ne, npos, oe, op, ionize_species=ion, name_a="E-", name_b="POSITRON")
    sp = [dict(name="E-", n=ne, occ=oe, charge=-1.0, lam=2, eps=O.synthetic_eps(oe, ne)),
          dict(name="POSITRON", n=npos, occ=op, charge=1.0, lam=1, eps=O.synthetic_eps(op, npos))]

    def aux(intra_e, intra_p, inter):
        if a == 0:
            return [O.read_pairs_intra(*intra_e, ne), O.read_pairs_inter(*inter, ne, npos)]
        return [O.read_pairs_inter(*inter, ne, npos, reversed_pair=True), O.read_pairs_intra(*intra_p, npos)]

    got = O.p2_poles(a, sp, aux(T.transform(0, 0, we, ol.CONV_E), T.transform(1, 1, wp, ol.CONV_E), T.transform(0, 1, wi, ol.CONV_E)))
    ref = O.p2_poles(a, sp, aux(O.transform_e_intra(Ce, pe, we), O.transform_e_intra(Cp, pp, wp), O.transform_e_inter(Ce, Cp, rect, wi)))
    assert len(got) == len(ref) == 2
    for g, r in zip(got, ref):
        assert g[0] == r[0] and g[4] == r[4]
        assert abs(g[2] - r[2]) <= 1e-9 and abs(g[3] - r[3]) <= 1e-9
        assert abs(g[2] - g[1]) > 1e-4            # the correction is not trivially zero
