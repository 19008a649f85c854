#!/bin/bash
# Multi-GPU call: parity check over NCCL (first half sharded over AO-pair slabs, all-to-all, second half on the slot owners)
# and one short bench line at NG GPUs.
TAG=${1:-r01g}; NG=${2:-2}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $O/${TAG}_smi.txt 2>&1
( timeout 300 $TR scripts/mgpu_check.py > $O/${TAG}_mgpu_check.log 2>&1; echo "exit $?" >> $O/${TAG}_mgpu_check.log ); tail -16 $O/${TAG}_mgpu_check.log
( timeout 400 $TR bench.py --gpus $NG --steps 1 --warmup 1 --no-e2e > $O/${TAG}_bench_n1500_g$NG.json 2> $O/${TAG}_bench_n1500_g$NG.err; echo "exit $?" >> $O/${TAG}_bench_n1500_g$NG.err ); tail -c 1800 $O/${TAG}_bench_n1500_g$NG.json; tail -5 $O/${TAG}_bench_n1500_g$NG.err
