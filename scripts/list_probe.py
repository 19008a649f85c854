"""Timing of the list-driven first quarter (row a14/f3, LOWDIN_IT_OPT_AO_LIST) against the packed-tensor path on the same input:
N=120 MP2 window O=21, the whole kind-H canonical list (26 M entries) pushed as raw .ints blocks.
usage: python scripts/list_probe.py <tag>"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402  (canonical_ao_list, raw_ints_blocks, window helpers: host-side input builders, no oracle)
import openlowdin_b200 as ol  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "probe"
n, occ, S = 120, 21, 30000
M = n * (n + 1) // 2
total = M * (M + 1) // 2
lst = bench.canonical_ao_list(bench.SEED, n)
raw = np.zeros((total // S + 1) * 24 * S, np.uint8)
nblk = bench.raw_ints_blocks(lst, total, S, raw)
win = bench.mp2_window_e(n, occ)
Cm = bench.random_orthonormal(n, n)
out = {"workload": f"N_bf={n} MP2 O={occ}, {total} canonical integrals"}
ref = None
for mode in ("packed", "list"):
    T = ol.Transformer(0)
    T.set_option(T.OPT_AO_LIST, 1 if mode == "list" else 0)
    T.set_species(0, Cm)
    T.upload_ao_blocks(0, 0, raw[:nblk * 24 * S], S)
    T.transform(0, 0, win, ol.CONV_E)                     # warm-up
    T.set_profiling(True)                                 # resets the per-category timers
    t0 = time.perf_counter()
    ij, kl, v = T.transform(0, 0, win, ol.CONV_E)
    wall = time.perf_counter() - t0
    tm, st = T.timers(), T.kernel_stats()
    out[mode] = {"transform_wall_ms": wall * 1e3, "first_half_ms": tm["first_half"] * 1e3, "second_half_ms": tm["second_half"] * 1e3,
                 "kernels_ms": {k: round(x["ms"], 3) for k, x in st.items() if x["launches"]}, "mo_integrals": int(len(v))}
    if ref is None:
        ref = (ij, kl, v)
    else:
        out["list_vs_packed"] = {"same_index_lists": bool(np.array_equal(ij, ref[0]) and np.array_equal(kl, ref[1])),
                                 "max_abs_diff": float(np.abs(v - ref[2]).max()) if len(v) == len(ref[2]) else None}
    T.close()
print(json.dumps(out))
json.dump(out, open(f"gpurun_out/{tag}_list_probe.json", "w"), indent=1)
