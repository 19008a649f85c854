#!/bin/bash
# Round-2 GPU call v (NG GPUs): one bench line with the occupied batch chosen by the full cost model (first quarter + third quarter).
TAG=${1:-r02v}; NG=${2:-2}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
( timeout 600 $TR --master-port 29512 bench.py --gpus $NG --steps 3 --warmup 1 --no-cpu-baseline > $O/${TAG}_bench_n1500_g${NG}.json 2> $O/${TAG}_bench_n1500_g${NG}.err; echo "exit $?" >> $O/${TAG}_bench_n1500_g${NG}.err )
python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench_n1500_g${NG}.json").read().strip().splitlines()[-1])
    print("G=$NG", round(d["value"]), "GFLOP/s", round(d["ms_per_step"]), "ms/step", d["config"].get("occ_batch"), "occ/pass x", d["config"]["passes_per_transform"], "e2e", round(d["e2e"]["value"] or 0), {k: (round(v["ms"]), round(v.get("TFLOP/s", v.get("GB/s", 0)), 1)) for k, v in d["kernels"].items()}, d["parity"])
except Exception as e:
    print("G=$NG failed:", e)
PY
tail -2 $O/${TAG}_bench_n1500_g${NG}.err
