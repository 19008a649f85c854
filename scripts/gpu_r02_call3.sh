#!/bin/bash
# Round-2 GPU call 3 (1 GPU): the suites touched since call 1 (per-test timeout), q1 split probe, ncu of the packed-row expansion.
TAG=${1:-r02c}
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -p timeout --timeout 150 --durations=12 -x > $O/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_gpu.log ); tail -40 $O/${TAG}_pytest_gpu.log
timeout 300 python scripts/q1_split_probe.py $TAG > $O/${TAG}_q1_split.log 2>&1; cat $O/${TAG}_q1_split.log
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
timeout 300 $NCU -k regex:"expand_block_kernel<.int.0>" -s 2 -c 1 -o $O/${TAG}_full_expand1_packed_n500 python bench.py --resident-only > $O/${TAG}_ncu_expand1.log 2>&1; tail -1 $O/${TAG}_ncu_expand1.log
ls -la $O | tail -8
