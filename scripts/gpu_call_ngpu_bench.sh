#!/bin/bash
# N-GPU bench only (one warm-up pass + one timed pass of the N=1500 MP2 transform), no end-to-end leg.
TAG=${1:-r01h}; NG=${2:-8}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
( timeout 300 $TR bench.py --gpus $NG --steps 1 --warmup 1 --no-e2e > $O/${TAG}_bench_n1500_g$NG.json 2> $O/${TAG}_bench_n1500_g$NG.err; echo "exit $?" >> $O/${TAG}_bench_n1500_g$NG.err ); tail -c 2000 $O/${TAG}_bench_n1500_g$NG.json; tail -5 $O/${TAG}_bench_n1500_g$NG.err
