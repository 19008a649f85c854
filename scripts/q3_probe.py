"""Kernel-alone timing of the third-quarter accumulation (read-modify-write of T3[slot][150][1500]) with one 8-warp CTA per SM
(default) and with two 4-warp CTAs per SM (LOWDIN_IT_OPT_Q3_TWO_CTA) -> gpurun_out/<tag>_q3_probe.json."""
import json
import sys

sys.path.insert(0, ".")
import openlowdin_b200 as ol  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "probe"
T = ol.Transformer(0)
out = {}
for two in (0, 4096):
    T.set_option(T.OPT_Q3_TWO_CTA, two)
    for slots, k in ((256, 32), (256, 48), (256, 64), (256, 96), (256, 128), (256, 256), (128, 1024)):
        m = slots * 1500
        ms, _ = T.kernel_bench(4, m, 150, k, iters=5)
        tf = 2.0 * m * 150 * k / (ms * 1e-3) / 1e12
        gb = m * 150 * 16 / (ms * 1e-3) / 1e9
        out[f"q3_two{int(bool(two))}_s{slots}_k{k}"] = {"ms": ms, "TFLOP/s": tf, "T3_GB/s": gb}
        print("q3 two_cta", int(bool(two)), "slots", slots, "k", k, "ms", round(ms, 3), "TF/s", round(tf, 2), "T3 read+write GB/s", round(gb), flush=True)
T.set_option(T.OPT_Q3_TWO_CTA, 0)
json.dump(out, open(f"gpurun_out/{tag}_q3_probe.json", "w"), indent=1)
T.close()
