"""Kernel-alone timing of the GEMM variants (1 = cp.async ring, 2 = TMA + mbarrier persistent) and of the fused
generation + first-quarter variants (1 = single-role, 3 = warp-specialised) -> gpurun_out/variant_probe.json."""
import json
import sys

sys.path.insert(0, ".")
import openlowdin_b200 as ol  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "probe"
T = ol.Transformer(0)
out = {}
shapes = [(8192, 8192, 8192), (1350, 28672, 1500), (1350, 89440, 1500), (1350, 8400, 1500), (28672, 56, 1500), (60000, 150, 32),
          (60000, 150, 128), (450, 21000, 500), (4096, 4096, 4096)]
for gv, split in ((2, 0), (2, 1)):
    T.set_option(T.OPT_GEMM_VARIANT, gv)
    T.set_option(T.OPT_SPLIT_ROW_TAIL, split)
    for (m, n, k) in shapes:
        try:
            ms, _ = T.kernel_bench(1, m, n, k, iters=3)
        except Exception as e:  # noqa: BLE001
            print("gemm variant", gv, "split", split, m, n, k, "FAILED", e, flush=True)
            out[f"gemm_v{gv}_split{split}_{m}x{n}x{k}"] = {"error": str(e)}
            continue
        tf = 2.0 * m * n * k / (ms * 1e-3) / 1e12
        out[f"gemm_v{gv}_split{split}_{m}x{n}x{k}"] = {"ms": ms, "TFLOP/s": tf}
        print("gemm variant", gv, "split", split, m, n, k, "ms", round(ms, 3), "TF/s", round(tf, 2), flush=True)
T.set_option(T.OPT_GEMM_VARIANT, T.DEFAULT_GEMM_VARIANT)
for variant, gen in [(3, 1), (4, 1), (3, 2), (4, 2)]:
    T.set_option(T.OPT_Q1_VARIANT, variant)
    T.set_option(T.OPT_BENCH_GEN, gen)
    for nc, nfb, bc in [(1500, 56, 512), (1500, 40, 512), (1500, 64, 512), (1500, 32, 512), (1500, 16, 1024), (500, 50, 2048), (1000, 32, 1024)]:
        try:
            ms, _ = T.kernel_bench(2, nc, nfb, bc, iters=3)
        except Exception as e:  # noqa: BLE001
            print("q1 variant", variant, "gen", gen, nc, nfb, bc, "FAILED", e, flush=True)
            out[f"q1_v{variant}_g{gen}_n{nc}_f{nfb}_b{bc}"] = {"error": str(e)}
            continue
        tf = 2.0 * bc * nc * nc * nfb / (ms * 1e-3) / 1e12
        out[f"q1_v{variant}_g{gen}_n{nc}_f{nfb}_b{bc}"] = {"ms": ms, "TFLOP/s": tf}
        print("q1 variant", variant, "gen", gen, nc, nfb, bc, "ms", round(ms, 3), "TF/s", round(tf, 2), flush=True)
# fragment-row permutation (LOWDIN_IT_OPT_FRAG_PERM): conflict-free 128-bit fragment loads, same kernels otherwise
T.set_option(T.OPT_GEMM_VARIANT, T.DEFAULT_GEMM_VARIANT)
T.set_option(T.OPT_SPLIT_ROW_TAIL, 1)
T.set_option(T.OPT_BENCH_GEN, 1)
for perm in (0, 1):
    T.set_option(T.OPT_FRAG_PERM, perm)
    for (m, n, k) in [(8192, 8192, 8192), (1350, 89440, 1500), (28672, 56, 1500), (60000, 150, 128)]:
        try:
            ms, _ = T.kernel_bench(1, m, n, k, iters=3)
            tf = 2.0 * m * n * k / (ms * 1e-3) / 1e12
            out[f"perm{perm}_gemm_{m}x{n}x{k}"] = {"ms": ms, "TFLOP/s": tf}
            print("perm", perm, "gemm", m, n, k, "ms", round(ms, 3), "TF/s", round(tf, 2), flush=True)
        except Exception as e:  # noqa: BLE001
            print("perm", perm, "gemm", m, n, k, "FAILED", e, flush=True)
    for variant in (3, 4):
        T.set_option(T.OPT_Q1_VARIANT, variant)
        for nc, nfb, bc in [(1500, 56, 512), (1500, 48, 512), (1500, 40, 512), (500, 50, 2048)]:
            try:
                ms, _ = T.kernel_bench(2, nc, nfb, bc, iters=3)
                tf = 2.0 * bc * nc * nc * nfb / (ms * 1e-3) / 1e12
                out[f"perm{perm}_q1_v{variant}_n{nc}_f{nfb}"] = {"ms": ms, "TFLOP/s": tf}
                print("perm", perm, "q1 variant", variant, nc, nfb, bc, "ms", round(ms, 3), "TF/s", round(tf, 2), flush=True)
            except Exception as e:  # noqa: BLE001
                print("perm", perm, "q1 variant", variant, nc, nfb, bc, "FAILED", e, flush=True)
T.set_option(T.OPT_FRAG_PERM, 0)
json.dump(out, open(f"gpurun_out/{tag}_variant_probe.json", "w"), indent=1)
