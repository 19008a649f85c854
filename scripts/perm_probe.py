"""Kernel-alone timing of the TMA GEMM and the fused first quarter with and without the fragment-row permutation
(LOWDIN_IT_OPT_FRAG_PERM) -> gpurun_out/<tag>_perm_probe.json."""
import json
import sys

sys.path.insert(0, ".")
import openlowdin_b200 as ol  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "probe"
T = ol.Transformer(0)
out = {}
for perm in (0, 1):
    T.set_option(T.OPT_FRAG_PERM, perm)
    for (m, n, k) in [(8192, 8192, 8192), (1350, 89440, 1500), (28672, 56, 1500), (60000, 150, 128), (60000, 150, 32), (450, 21000, 500)]:
        try:
            ms, _ = T.kernel_bench(1, m, n, k, iters=3)
            tf = 2.0 * m * n * k / (ms * 1e-3) / 1e12
            out[f"perm{perm}_gemm_{m}x{n}x{k}"] = {"ms": ms, "TFLOP/s": tf}
            print("perm", perm, "gemm", m, n, k, "ms", round(ms, 3), "TF/s", round(tf, 2), flush=True)
        except Exception as e:  # noqa: BLE001
            print("perm", perm, "gemm", m, n, k, "FAILED", e, flush=True)
    for variant in (3, 4):
        T.set_option(T.OPT_Q1_VARIANT, variant)
        for nc, nfb, bc in [(1500, 56, 512), (1500, 48, 512), (1500, 40, 512), (1500, 64, 512), (500, 50, 2048)]:
            try:
                ms, _ = T.kernel_bench(2, nc, nfb, bc, iters=3)
                tf = 2.0 * bc * nc * nc * nfb / (ms * 1e-3) / 1e12
                out[f"perm{perm}_q1_v{variant}_n{nc}_f{nfb}"] = {"ms": ms, "TFLOP/s": tf}
                print("perm", perm, "q1 variant", variant, nc, nfb, bc, "ms", round(ms, 3), "TF/s", round(tf, 2), flush=True)
            except Exception as e:  # noqa: BLE001
                print("perm", perm, "q1 variant", variant, nc, nfb, bc, "FAILED", e, flush=True)
T.set_option(T.OPT_FRAG_PERM, 0)
json.dump(out, open(f"gpurun_out/{tag}_perm_probe.json", "w"), indent=1)
