#!/bin/bash
# Multi-GPU call of the next round: parity over NCCL, then the bench line at NG GPUs (strong scaling of the N=1500 transform).
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_round2_scaling.sh r02b 8'
TAG=${1:-r02b}; NG=${2:-8}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
( timeout 300 $TR scripts/mgpu_check.py > $O/${TAG}_mgpu_check_g$NG.log 2>&1; echo "exit $?" >> $O/${TAG}_mgpu_check_g$NG.log ); tail -8 $O/${TAG}_mgpu_check_g$NG.log
for G in 1 2 4 8; do
  [ $G -gt $NG ] && break
  if [ $G -eq 1 ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512"; fi
  ( timeout 500 $RUN bench.py --gpus $G --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_n1500_g$G.json 2> $O/${TAG}_bench_n1500_g$G.err; echo "exit $?" >> $O/${TAG}_bench_n1500_g$G.err )
  python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench_n1500_g$G.json").read().strip().splitlines()[-1])
    print("G=$G", round(d["value"]), "GFLOP/s", round(d["ms_per_step"]), "ms/step", d["config"].get("occ_batch"), "occ/pass", {k: round(v["ms"]) for k, v in d["kernels"].items()})
except Exception as e:
    print("G=$G failed:", e)
PY
  tail -2 $O/${TAG}_bench_n1500_g$G.err
done
