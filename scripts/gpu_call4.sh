#!/bin/bash
# GPU call 4 (1 GPU): q1 variant 4 (pipelined DMMA warps + 4 vectorised generator warps), row-tail split, full-width RMW
# epilogue, templated expansion kernel: full parity suite, kernels alone, one bench pass, one ncu --set full capture.
TAG=${1:-r01f}
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_gpu.log ); tail -6 $O/${TAG}_pytest_gpu.log
timeout 300 python scripts/variant_probe.py $TAG > $O/${TAG}_variant_probe.log 2>&1; cat $O/${TAG}_variant_probe.log
timeout 400 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --q1-variant 4 > $O/${TAG}_bench_n1500_q4.json 2> $O/${TAG}_bench_n1500_q4.err; tail -c 1300 $O/${TAG}_bench_n1500_q4.json; tail -3 $O/${TAG}_bench_n1500_q4.err
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
timeout 200 $NCU -k regex:q1_gen_ws2 -s 6 -c 1 --kill 1 -o $O/${TAG}_full_q1ws2_n1500 python scripts/ncu_target.py 1500 1 1 4 > $O/${TAG}_ncu_q1.log 2>&1; tail -1 $O/${TAG}_ncu_q1.log
ls -la $O | tail -6
