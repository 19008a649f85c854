"""Diagnostic: fused stored first quarter vs the two-kernel form over ranges of slabs; prints where they differ.
usage: diag_fused.py n nf slab_count [range starts as fractions of M ...]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import openlowdin_b200 as ol  # noqa: E402

n, nf, ns = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
fracs = [float(x) for x in sys.argv[4:]] or [0.0]
q, _ = np.linalg.qr(np.random.default_rng(n).standard_normal((n, n)))
T = ol.Transformer(0)
T.set_species(0, np.asfortranarray(q))
T.set_generator(0, 0, 5)
T.materialize(0, 0)
M = n * (n + 1) // 2
ns = min(ns, M)
rb = (n + 255) // 256
for fr in fracs:
    s0 = min(int(fr * M), M - ns)
    T.set_option(T.OPT_STORED_FUSED, 0)
    ref = T.debug_first_quarter(0, 0, 1, nf, s0, ns)
    T.set_option(T.OPT_STORED_FUSED, 1)
    for rep in range(2):
        got = T.debug_first_quarter(0, 0, 1, nf, s0, ns)
        bad = got != ref
        print(f"slabs [{s0}, {s0 + ns}) rep {rep}: {bad.sum()} of {bad.size} elements differ, max |d| = {np.abs(got - ref).max():.3e}", flush=True)
        if bad.any():
            f, z, m = np.nonzero(bad)
            print("  bad f:", np.unique(f)[:60])
            print("  bad slabs (relative):", np.unique(z)[:40], "count", len(np.unique(z)))
            print("  bad rows m:", np.unique(m)[:64], "count", len(np.unique(m)))
            print("  first:", (f[0], z[0], m[0]), "got", got[f[0], z[0], m[0]], "ref", ref[f[0], z[0], m[0]])
            zz = z[0]
            X = T.debug_expand(0, 0, s0 + int(zz), 1)[0]
            want = X[m[0], :] @ q[:, f[0]]
            print("  numpy value of that element:", want)
            # is the wrong value what a neighbouring k-tile / row would give?
            d = got[f[0], zz, :] - ref[f[0], zz, :]
            print("  nonzero diffs in that (f, slab) column:", np.nonzero(d)[0][:40], d[np.nonzero(d)[0][:6]])
