"""Where does the fused first quarter lose its tensor-pipe time?  Variants 3 and 5 with the probe switches of Q1WsArgs::dbg:
0 = normal, 1 = generator warps store zeros (no hashing), 2 = DMMA warps skip loads and DMMAs, 3 = both,
4 / 5 (variant 3) = roles assigned by scheduler (schedulers 0-1 run the DMMA warps, 2-3 the generators) -> gpurun_out/<tag>_q1_split.json"""
import json
import sys

sys.path.insert(0, ".")
import openlowdin_b200 as ol  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "probe"
T = ol.Transformer(0)
out = {}
for variant, modes in ((3, (0, 1, 2, 3, 4, 5)), (5, (0, 1, 2, 3))):
    T.set_option(T.OPT_Q1_VARIANT, variant)
    for gen in (1, 2):
        T.set_option(T.OPT_BENCH_GEN, gen)
        for dbg in modes:
            T.set_option(T.OPT_Q1_DEBUG, dbg)
            for nc, nfb, bc in [(1500, 56, 512), (1500, 40, 512)]:
                ms, _ = T.kernel_bench(2, nc, nfb, bc, iters=3)
                tf = 2.0 * bc * nc * nc * nfb / (ms * 1e-3) / 1e12
                out[f"v{variant}_g{gen}_dbg{dbg}_f{nfb}"] = {"ms": ms, "TFLOP/s_equivalent": tf}
                print("q1 variant", variant, "gen", gen, "dbg", dbg, nc, nfb, bc, "ms", round(ms, 3), "TF/s-equivalent", round(tf, 2), flush=True)
T.set_option(T.OPT_Q1_DEBUG, 0)
json.dump(out, open(f"gpurun_out/{tag}_q1_split.json", "w"), indent=1)
