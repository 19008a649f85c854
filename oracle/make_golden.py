#!/usr/bin/env python
"""oracle/make_golden.py -- generate tests/golden/*.npz from the REFERENCE ITSELF.  TEST INFRASTRUCTURE ONLY.

The reference ships no MO-integral golden vectors, so the fixtures are outputs of the
reference's own transformer D (src/integralsTransformation/IntTransfD.cpp: c_integrals_transform_all,
c_integrals_transform_inter_all) compiled in place into oracle/_ref/libref_d.so by oracle/Makefile and run
HERE, in the build container, on seeded inputs.  Each fixture stores inputs and outputs so that the
GPU box (which has no /root/reference) can check both the oracle restatement and the CUDA path against
what the reference code computed.

  python oracle/make_golden.py        # rewrites tests/golden/d_intra_*.npz, d_inter_*.npz

Fixture content (all float64 unless noted):
  d_intra_n{N}.npz : C [N,N] column-major coefficients, eris_in [M(M+1)/2] D-packed AO integrals
                     (ERIS[ij(ij+1)/2+kl], ij = i(i+1)/2+j 0-based lower triangle, IntTransfD.cpp:25-45),
                     eris_out = the same array after c_integrals_transform_all (in-place MO integrals)
  d_inter_{NA}x{NB}.npz : Ca, Cb, eris_in [Ma*Mb] (ij*Mb+kl, IntTransfD.cpp:48-65), eris_out
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path = [ROOT] + [p for p in sys.path if os.path.abspath(p or '.') != HERE]
from oracle import oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
INTRA = [(3, 31), (5, 32), (7, 33), (11, 34), (19, 35)]      # (N, seed): up to the H2O/6-311G shape
INTER = [(4, 3, 41), (6, 4, 42), (8, 5, 43), (19, 7, 44)]   # (N_a, N_b, seed)


def main() -> None:
    O.build()
    R = O.ref()
    if R is None:
        raise SystemExit("oracle/_ref/libref_d.so missing: the reference tree is needed to make golden vectors")
    os.makedirs(GOLD, exist_ok=True)
    for n, seed in INTRA:
        rng = np.random.default_rng(seed)
        M = O.npairs(n)
        sq = rng.uniform(-1, 1, (M, M))
        sq = sq + sq.T
        eris = O.d_pack_intra(sq)
        Cm = O.random_orthonormal(n, seed)
        out = O.transform_d_intra(Cm, eris, use_reference=True)
        np.savez_compressed(os.path.join(GOLD, f"d_intra_n{n}.npz"), C=np.asarray(Cm), eris_in=eris, eris_out=out,
                            source="IntTransfD.cpp c_integrals_transform_all via oracle/_ref/libref_d.so")
        print(f"d_intra_n{n}: {eris.size} values")
    for na, nb, seed in INTER:
        rng = np.random.default_rng(seed)
        er = rng.uniform(-1, 1, O.npairs(na) * O.npairs(nb))
        Ca, Cb = O.random_orthonormal(na, seed), O.random_orthonormal(nb, seed + 100)
        out = O.transform_d_inter(Ca, Cb, er, use_reference=True)
        np.savez_compressed(os.path.join(GOLD, f"d_inter_{na}x{nb}.npz"), Ca=np.asarray(Ca), Cb=np.asarray(Cb), eris_in=er,
                            eris_out=out, source="IntTransfD.cpp c_integrals_transform_inter_all via oracle/_ref/libref_d.so")
        print(f"d_inter_{na}x{nb}: {er.size} values")


if __name__ == "__main__":
    main()
