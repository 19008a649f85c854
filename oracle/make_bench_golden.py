"""tests/golden/bench_sums.json: reference values of the streamed sums (count, sum, sum of squares, MP2 pair energy) that
bench.py compares its own results with.

  "small"    -- N=24, O=6 kind-H tensor, computed HERE by the CPU oracle (restated transformer E + the MP2 reader): the case
                bench.py --gpus N>1 runs over its communicator before timing anything.
  "n<N>_gen<g>" -- whole MP2-window transform at the benchmark sizes.  No CPU can run those; the entries are recorded from a
                1-GPU run of bench.py (the `parity.sums` of its line) whose kernels are checked against the oracle in the same
                round (first-half values at that size in bench.py's cpu_baseline leg; kind-K closed form at N=500/1500 in
                tests/test_gpu_rankk.py), and serve as the cross-check of the 2/4/8-GPU runs.  This script keeps them as they are.
Usage: python oracle/make_bench_golden.py [bench_line.json ...]   (bench lines add/replace their n<N>_gen<g> entry)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path = [ROOT] + [x for x in sys.path if os.path.abspath(x or ".") != os.path.join(ROOT, "oracle")]
import bench  # noqa: E402  (the same input generators the bench uses)
from oracle import oracle as O  # noqa: E402

PATH = os.path.join(ROOT, "tests", "golden", "bench_sums.json")


def main():
    try:
        out = json.load(open(PATH))
    except (OSError, ValueError):
        out = {}
    n, occ, seed = 24, 6, 31337
    Cm = bench.random_orthonormal(n, n)
    packed = O.hash_packed_intra(seed, n)
    win = bench.mp2_window_e(n, occ)
    ij, kl, v = O.transform_e_intra(Cm, packed, win)
    e2 = O.mp2_intra_from_pairs(ij, kl, v, n, occ, bench.synthetic_eps(occ, n), lam=2.0)
    out["small"] = {"n": n, "occ": occ, "seed": seed, "sums": [float(len(v)), float(v.sum()), float((v * v).sum()), float(e2)],
                    "source": "oracle: orc_transform_e_intra + orc_mp2_intra (oracle/make_bench_golden.py)"}
    for path in sys.argv[1:]:
        line = json.loads(open(path).read().strip().splitlines()[-1])
        cfg, par = line["config"], line["parity"]
        if par["passes_covered"].split()[0] != par["passes_covered"].split()[-1]:
            raise SystemExit(f"{path}: not every pass of the transform was run")
        s = par["sums"]
        out[f"n{cfg['nbf']}_gen{cfg.get('gen', 1)}"] = {
            "sums": [s["count"], s["sum"], s["sum_sq"], s["mp2_pair_energy"]],
            "source": f"bench.py on {line['n_gpus']} GPU(s), occ_batch {cfg['occ_batch']}, {os.path.basename(path)}"}
    json.dump(out, open(PATH, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
