#!/bin/bash
# Round-2 GPU call x (NG GPUs): the peer-to-peer DMA exchange: parity over the communicator (mgpu_check), bench line with it and with NCCL send/recv.
TAG=${1:-r02x}; NG=${2:-2}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
( timeout 400 $TR --master-port 29511 scripts/mgpu_check.py > $O/${TAG}_mgpu_check_g$NG.log 2>&1; echo "exit $?" >> $O/${TAG}_mgpu_check_g$NG.log ); grep -v "^W\|^\[W\|warn" $O/${TAG}_mgpu_check_g$NG.log | tail -12 | cut -c1-300
for DMA in 1 0; do
( timeout 500 $TR --master-port 2951$DMA bench.py --gpus $NG --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --exchange-dma $DMA > $O/${TAG}_bench_n1500_g${NG}_dma$DMA.json 2> $O/${TAG}_bench_n1500_g${NG}_dma$DMA.err; echo "exit $?" >> $O/${TAG}_bench_n1500_g${NG}_dma$DMA.err )
python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench_n1500_g${NG}_dma$DMA.json").read().strip().splitlines()[-1])
    print("G=$NG dma=$DMA", d["config"]["exchange"][:20], round(d["value"]), "GFLOP/s", round(d["ms_per_step"]), "ms/step", d["config"].get("occ_batch"), "occ/pass x", d["config"]["passes_per_transform"], {k: (round(v["ms"]), round(v.get("TFLOP/s", v.get("GB/s", 0)), 1)) for k, v in d["kernels"].items()}, d["parity"]["whole_transform_vs_reference_sums"]["ok"], d["parity"]["small_collective_transform"])
except Exception as e:
    print("G=$NG dma=$DMA failed:", e)
PY
tail -3 $O/${TAG}_bench_n1500_g${NG}_dma$DMA.err | cut -c1-300
done
