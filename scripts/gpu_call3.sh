#!/bin/bash
# GPU call 3 (1 GPU): full parity suite on the new defaults, ncu --set full of the new hot kernels, one bench pass.
TAG=${1:-r01e}
mkdir -p gpurun_out
O=gpurun_out
( timeout 500 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_gpu.log ); tail -4 $O/${TAG}_pytest_gpu.log
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
timeout 200 $NCU -k regex:q1_gen_ws -s 6 -c 1 --kill 1 -o $O/${TAG}_full_q1ws_n1500 python scripts/ncu_target.py 1500 > $O/${TAG}_ncu_q1.log 2>&1; tail -1 $O/${TAG}_ncu_q1.log
timeout 200 $NCU -k regex:EpiScatterH -s 6 -c 1 --kill 1 -o $O/${TAG}_full_q2tma_n1500 python scripts/ncu_target.py 1500 > $O/${TAG}_ncu_q2.log 2>&1; tail -1 $O/${TAG}_ncu_q2.log
timeout 200 $NCU -k regex:EpiAccT -s 6 -c 2 --kill 1 -o $O/${TAG}_full_q3tma_n1500 python scripts/ncu_target.py 1500 > $O/${TAG}_ncu_q3.log 2>&1; tail -1 $O/${TAG}_ncu_q3.log
timeout 400 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/${TAG}_bench_n1500.json 2> $O/${TAG}_bench_n1500.err; tail -c 1300 $O/${TAG}_bench_n1500.json; tail -3 $O/${TAG}_bench_n1500.err
ls -la $O | tail -8
