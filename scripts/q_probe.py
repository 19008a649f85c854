"""Kernel-alone timing: fused first quarter variants 3 / 5, third-quarter accumulation as staged read-modify-write (kind 4) and as
red.global.add.f64 (kind 5) -> gpurun_out/<tag>_q_probe.json."""
import json
import sys

sys.path.insert(0, ".")
import openlowdin_b200 as ol  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "probe"
T = ol.Transformer(0)
out = {}
for variant in (3, 5):
    T.set_option(T.OPT_Q1_VARIANT, variant)
    for nc, nfb, bc in [(1500, 56, 512), (1500, 48, 512), (1500, 40, 512), (1500, 64, 512), (1500, 32, 512), (1500, 16, 1024), (500, 50, 2048), (1000, 32, 1024), (2000, 56, 256)]:
        try:
            ms, _ = T.kernel_bench(2, nc, nfb, bc, iters=3)
            tf = 2.0 * bc * nc * nc * nfb / (ms * 1e-3) / 1e12
            out[f"q1_v{variant}_n{nc}_f{nfb}"] = {"ms": ms, "TFLOP/s": tf}
            print("q1 variant", variant, nc, nfb, bc, "ms", round(ms, 3), "TF/s", round(tf, 2), flush=True)
        except Exception as e:  # noqa: BLE001
            print("q1 variant", variant, nc, nfb, bc, "FAILED", e, flush=True)
T.set_option(T.OPT_Q1_VARIANT, T.DEFAULT_Q1_VARIANT)
for kind, name in ((4, "rmw"), (5, "red")):
    for slots, k in ((64, 32), (64, 64), (64, 128), (64, 512), (64, 1500), (256, 32), (256, 128)):
        m = slots * 1500
        try:
            ms, _ = T.kernel_bench(kind, m, 150, k, iters=5)
            tf = 2.0 * m * 150 * k / (ms * 1e-3) / 1e12
            gb = m * 150 * 16 / (ms * 1e-3) / 1e9
            out[f"q3_{name}_s{slots}_k{k}"] = {"ms": ms, "TFLOP/s": tf, "T3_GB/s": gb}
            print("q3", name, "slots", slots, "k", k, "ms", round(ms, 3), "TF/s", round(tf, 2), "T3 read+write GB/s", round(gb), flush=True)
        except Exception as e:  # noqa: BLE001
            print("q3", name, slots, k, "FAILED", e, flush=True)
json.dump(out, open(f"gpurun_out/{tag}_q_probe.json", "w"), indent=1)

