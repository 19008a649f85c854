"""Timing of the device AO-integral producer (row f4): a benzene-like ring of 12 centres, contracted s / p shells on every centre,
d shells on the six heavy ones (N = 102 Cartesian functions), then the MP2 transform of the computed tensor.
usage: python scripts/eri_probe.py <tag>"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import openlowdin_b200 as ol  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "probe"
s3 = ([130.70932, 23.808861, 6.4436083], [0.15432897, 0.53532814, 0.44463454])
sp = ([5.0331513, 1.1695961, 0.3803890], [-0.09996723, 0.39951283, 0.70011547])
pp = ([5.0331513, 1.1695961, 0.3803890], [0.15591627, 0.60768372, 0.39195739])
hs = ([3.42525091, 0.62391373, 0.16885540], [0.15432897, 0.53532814, 0.44463454])
shells = []
for k in range(6):
    a = np.pi / 3 * k
    C = (2.63 * np.cos(a), 2.63 * np.sin(a), 0.0)
    H = (4.69 * np.cos(a), 4.69 * np.sin(a), 0.0)
    shells += [(0, C, *s3), (0, C, *sp), (1, C, *pp), (2, C, [0.8], [1.0]), (0, H, *hs), (1, H, [0.75], [1.0])]
n = sum((l + 1) * (l + 2) // 2 for l, *_ in shells)
occ = 21
q, _ = np.linalg.qr(np.random.default_rng(5).standard_normal((n, n)))
T = ol.Transformer(0)
T.set_species(0, np.asfortranarray(q))
T.set_basis(0, shells)
T.compute_ao(0, 0)          # warm-up (module load, local-memory reservation)
t0 = time.perf_counter()
T.compute_ao(0, 0)
wall = time.perf_counter() - t0
dev = T.timers()["ao_upload"]      # the device time of the fill kernel (reported like an upload)
M = n * (n + 1) // 2
nint = M * (M + 1) // 2
win = [occ + 1, n, 1, occ, occ + 1, n, 1, occ]
t1 = time.perf_counter()
ij, kl, v = T.transform(0, 0, win, ol.CONV_E)
tw = time.perf_counter() - t1
out = {"nbf": n, "shells": len(shells), "canonical_integrals": nint, "compute_wall_s": wall, "compute_device_s": dev,
       "integrals_per_s": nint / (dev or wall), "transform_wall_s": tw, "mo_integrals": int(len(v))}
print(json.dumps(out))
json.dump(out, open(f"gpurun_out/{tag}_eri_probe.json", "w"), indent=1)
T.close()
