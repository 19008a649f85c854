"""Micro-benchmark of the fused generation + first-quarter kernel alone (gpurun_out/q1_probe.json)."""
import json, sys
sys.path.insert(0, ".")
import openlowdin_b200 as ol
T = ol.Transformer(0)
out = {}
for variant, gen in [(1, 1), (2, 1), (1, 2), (2, 2)]:
  T.set_option(T.OPT_Q1_VARIANT, variant); T.set_option(T.OPT_BENCH_GEN, gen)
  for nc, nfb, bc in [(1500, 56, 512), (1500, 40, 512), (1500, 64, 512), (1500, 32, 512), (1500, 16, 1024), (500, 50, 2048)]:
    ms, _ = T.kernel_bench(2, nc, nfb, bc, iters=3)
    tf = 2.0 * bc * nc * nc * nfb / (ms * 1e-3) / 1e12
    out[f"q1_gen_v{variant}_g{gen}_n{nc}_f{nfb}_b{bc}"] = {"ms": ms, "TFLOP/s": tf}
    print("variant", variant, "gen", gen, nc, nfb, bc, "ms", round(ms, 3), "TF/s", round(tf, 2), flush=True)
json.dump(out, open("gpurun_out/q1_probe.json", "w"), indent=1)
