"""Kernel-alone timing of the second / fourth quarter shapes: 128 x 128 tiles + row-tail launch against 192 x 64 tiles
(LOWDIN_IT_OPT_GEMM_TALL) -> gpurun_out/<tag>_gemm_tall_probe.json."""
import json
import sys

sys.path.insert(0, ".")
import openlowdin_b200 as ol  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "probe"
T = ol.Transformer(0)
out = {}
for tall in (0, 1):
    T.set_option(T.OPT_GEMM_TALL, tall)
    for (m, n, k) in [(1350, 89440, 1500), (1350, 28000, 1500), (1800, 40000, 2000), (900, 51200, 1000), (450, 53550, 500)]:
        try:
            ms, _ = T.kernel_bench(1, m, n, k, iters=3)
            tf = 2.0 * m * n * k / (ms * 1e-3) / 1e12
            out[f"gemm_tall{tall}_{m}x{n}x{k}"] = {"ms": ms, "TFLOP/s": tf}
            print("gemm tall", tall, m, n, k, "ms", round(ms, 3), "TF/s", round(tf, 2), flush=True)
        except Exception as e:  # noqa: BLE001
            print("gemm tall", tall, m, n, k, "FAILED", e, flush=True)
T.set_option(T.OPT_GEMM_TALL, 0)
json.dump(out, open(f"gpurun_out/{tag}_gemm_tall_probe.json", "w"), indent=1)
T.close()
