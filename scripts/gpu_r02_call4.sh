#!/bin/bash
# Round-2 GPU call 4 (1 GPU): whole parity suite (per-test timeout), stored-AO resident leg (fused vs two-kernel first quarter),
# ncu of the fused stored first quarter, the default bench line (q1 variant 5, e2e through the host sink).
TAG=${1:-r02d}
mkdir -p gpurun_out
O=gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -p timeout --timeout 200 --durations=12 > $O/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_gpu.log ); tail -45 $O/${TAG}_pytest_gpu.log
timeout 200 python bench.py --resident-only > $O/${TAG}_resident_n500.json 2> $O/${TAG}_resident_n500.err; tail -c 3500 $O/${TAG}_resident_n500.json; tail -3 $O/${TAG}_resident_n500.err
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
timeout 300 $NCU -k regex:q1_load_ws5 -s 3 -c 1 -o $O/${TAG}_full_q1load_n500 python bench.py --resident-only > $O/${TAG}_ncu_q1load.log 2>&1; tail -1 $O/${TAG}_ncu_q1load.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/${TAG}_bench_n1500.json 2> $O/${TAG}_bench_n1500.err; tail -c 7000 $O/${TAG}_bench_n1500.json; tail -3 $O/${TAG}_bench_n1500.err
ls -la $O | tail -8
