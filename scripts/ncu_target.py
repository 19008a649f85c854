"""Light target for ncu captures: one occupied-batch pass of the bench workload through the C ABI (no torch import, so
that the profiler attaches to a short process).  Not a bench: numbers printed under ncu are never bench values.
usage: python scripts/ncu_target.py [nbf] [passes] [gen] [q1_variant] [gemm_variant]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openlowdin_b200 as ol  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
gen = int(sys.argv[3]) if len(sys.argv) > 3 else 1
q1v = int(sys.argv[4]) if len(sys.argv) > 4 else 0
gv = int(sys.argv[5]) if len(sys.argv) > 5 else 0
occ = n // 10
win = [occ + 1, n, 1, occ, occ + 1, n, 1, occ]
q, _ = np.linalg.qr(np.random.default_rng(n).standard_normal((n, n)))
eps = np.concatenate([np.linspace(-2.0, -0.5, occ), np.linspace(0.2, 3.0, n - occ)])
T = ol.Transformer(0)
T.set_species(0, np.asfortranarray(q))
if q1v:
    T.set_option(T.OPT_Q1_VARIANT, q1v)
if gv and hasattr(T, "OPT_GEMM_VARIANT"):
    T.set_option(T.OPT_GEMM_VARIANT, gv)
T.set_generator(0, 0, 20261017, gen)
npass, qb = T.num_passes(0, 0, win, ol.CONV_E, 0)
for i in range(passes):
    s = T.transform_stream(0, 0, win, ol.CONV_E, occ_batch=qb, first_pass=i % npass, n_passes=1, epsA=eps)
    tm = T.timers()
    dev = tm["first_half"] + tm["exchange"] + tm["second_half"] + tm["consume"]
    print(f"pass {i}: occ_batch {qb} of {npass} passes, {dev:.3f} s, {tm['flops'] / max(dev, 1e-9) / 1e12:.2f} TF/s, sums {s}", flush=True)
T.close()
