#!/bin/bash
# Round-2 GPU call y (NG GPUs, meant for 8): parity over the communicator with the DMA exchange, then the default bench line.
TAG=${1:-r02y}; NG=${2:-8}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
( [ "$SKIPCHECK" = "1" ] || timeout 300 $TR --master-port 29511 scripts/mgpu_check.py > $O/${TAG}_mgpu_check_g$NG.log 2>&1; echo "exit $?" >> $O/${TAG}_mgpu_check_g$NG.log ); grep -v "^W\|^\[W\|warn" $O/${TAG}_mgpu_check_g$NG.log | tail -6 | cut -c1-300
( timeout 400 $TR --master-port 29512 bench.py --gpus $NG --steps ${STEPS:-3} --warmup 1 --no-cpu-baseline ${EXTRA} > $O/${TAG}_bench_n1500_g${NG}.json 2> $O/${TAG}_bench_n1500_g${NG}.err; echo "exit $?" >> $O/${TAG}_bench_n1500_g${NG}.err )
python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench_n1500_g${NG}.json").read().strip().splitlines()[-1])
    print("G=$NG", d["config"]["exchange"][:20], round(d["value"]), "GFLOP/s", round(d["ms_per_step"]), "ms/step", d["config"].get("occ_batch"), "occ/pass x", d["config"]["passes_per_transform"], "e2e", round(d["e2e"]["value"] or 0), {k: (round(v["ms"]), round(v.get("TFLOP/s", v.get("GB/s", 0)), 1)) for k, v in d["kernels"].items()}, d["parity"]["whole_transform_vs_reference_sums"]["ok"], d["parity"]["small_collective_transform"])
except Exception as e:
    print("G=$NG failed:", e)
PY
tail -3 $O/${TAG}_bench_n1500_g${NG}.err | cut -c1-300
