#!/bin/bash
# Multi-GPU call: parity over NCCL (stream sums, collective download -> one moint.dat, species pairs), then bench lines at NG GPUs.
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_r02_mgpu.sh r02m 2'
TAG=${1:-r02m}; NG=${2:-2}; STEPS=${3:-1}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
( [ "$SKIPCHECK" = "1" ] || timeout 400 $TR scripts/mgpu_check.py > $O/${TAG}_mgpu_check_g$NG.log 2>&1; echo "exit $?" >> $O/${TAG}_mgpu_check_g$NG.log ); grep -v "^W\|^\[W\|warn" $O/${TAG}_mgpu_check_g$NG.log | tail -22
( [ "$SKIPCHECK" = "1" ] || timeout 300 python -m pytest tests/test_multi_gpu.py tests/test_gpu_local_group.py -m gpu -q -p timeout --timeout 250 > $O/${TAG}_pytest_mgpu.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_mgpu.log ); tail -4 $O/${TAG}_pytest_mgpu.log
for OV in 1 0; do
  ( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps $STEPS --warmup 1 --no-cpu-baseline --overlap $OV > $O/${TAG}_bench_n1500_g${NG}_ov$OV.json 2> $O/${TAG}_bench_n1500_g${NG}_ov$OV.err; echo "exit $?" >> $O/${TAG}_bench_n1500_g${NG}_ov$OV.err )
  python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench_n1500_g${NG}_ov$OV.json").read().strip().splitlines()[-1])
    print("G=$NG overlap=$OV", round(d["value"]), "GFLOP/s", round(d["ms_per_step"]), "ms/step", d["config"].get("occ_batch"), "occ/pass", "e2e", round(d["e2e"]["value"] or 0), {k: (round(v["ms"]), round(v.get("TFLOP/s", v.get("GB/s", 0)), 1)) for k, v in d["kernels"].items()}, d["parity"])
except Exception as e:
    print("G=$NG failed:", e)
PY
  tail -2 $O/${TAG}_bench_n1500_g${NG}_ov$OV.err
done
