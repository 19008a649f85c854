"""Diagnostic: fused stored first quarter vs the two-kernel form over many slabs; prints where they differ."""
import sys

import numpy as np

sys.path.insert(0, ".")
import openlowdin_b200 as ol  # noqa: E402

n, nf, ns = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (300, 50, 2000)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
q, _ = np.linalg.qr(np.random.default_rng(n).standard_normal((n, n)))
T = ol.Transformer(0)
T.set_species(0, np.asfortranarray(q))
T.set_generator(0, 0, 5)
T.materialize(0, 0)
M = n * (n + 1) // 2
ns = min(ns, M)
T.set_option(T.OPT_STORED_FUSED, 0)
ref = T.debug_first_quarter(0, 0, 1, nf, 0, ns)
T.set_option(T.OPT_STORED_FUSED, 1)
rb = (n + 255) // 256
for rep in range(reps):
    got = T.debug_first_quarter(0, 0, 1, nf, 0, ns)
    bad = np.abs(got - ref) > 1e-11
    print(f"rep {rep}: n={n} nf={nf} slabs={ns}: {bad.sum()} of {bad.size} elements differ, max |d| = {np.abs(got - ref).max():.3e}", flush=True)
    if bad.any():
        f, z, m = np.nonzero(bad)
        print("  bad f:", np.unique(f))
        print("  bad rows m: min", m.min(), "max", m.max(), "count", len(np.unique(m)))
        tile = z * rb + m // 256
        ordinal = tile // 148
        hist = np.bincount(ordinal, minlength=int(ordinal.max()) + 1)
        print("  bad elements by tile ordinal within its CTA:", hist[:40])
        # per (slab,row-block) tile: is the whole column f wrong or only some rows?
        t0 = tile[0]
        sel = tile == t0
        print("  first bad tile", t0, "slab", z[0], "f values", np.unique(f[sel]), "rows", np.unique(m[sel])[:20], "n rows", len(np.unique(m[sel])))
        zz, ff = z[0], f[0]
        g, r = got[ff, zz, :], ref[ff, zz, :]
        print("  got[:6]", g[:6], "\n  ref[:6]", r[:6])
        # does the bad column equal another column of the reference?
        for f2 in range(nf):
            if np.allclose(g, ref[f2, zz, :], atol=1e-9):
                print("  -> got column", ff, "equals reference column", f2)
        print("  ratio got/ref (first 6):", g[:6] / r[:6])
