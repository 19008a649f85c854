#!/bin/bash
# Round-2 GPU call 2 (1 GPU): list-driven first quarter, q1 variant 5, third-quarter red epilogue, ncu of the packed-row expansion.
TAG=${1:-r02b}
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_list.py tests/test_gpu_variants.py tests/test_gpu_parity.py -m gpu -q --durations=8 > $O/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_gpu.log ); tail -30 $O/${TAG}_pytest_gpu.log
timeout 300 python scripts/q_probe.py $TAG > $O/${TAG}_q_probe.log 2>&1; cat $O/${TAG}_q_probe.log
for CFG in "--q1-variant 5" "--q3-red 1" "--q1-variant 5 --q3-red 1"; do
  NAME=$(echo $CFG | tr -d ' -')
  timeout 200 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline $CFG > $O/${TAG}_bench_n1500_$NAME.json 2> $O/${TAG}_bench_n1500_$NAME.err
  python -c "
import json,sys
d=json.loads(open('$O/${TAG}_bench_n1500_$NAME.json').read().strip().splitlines()[-1]); print('$CFG', round(d['value']), {k:(round(v['ms']), round(v.get('TFLOP/s', v.get('GB/s',0)),2)) for k,v in d['kernels'].items()}, d['parity']['sums'])"
  tail -2 $O/${TAG}_bench_n1500_$NAME.err
done
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
timeout 300 $NCU -k regex:"expand_block_kernel<0>" -s 2 -c 1 -o $O/${TAG}_full_expand1_packed_n500 python bench.py --resident-only > $O/${TAG}_ncu_expand1.log 2>&1; tail -1 $O/${TAG}_ncu_expand1.log
timeout 300 $NCU -k regex:q1_gen_ws5 -s 4 -c 1 -o $O/${TAG}_full_q1ws5_n1500 python scripts/ncu_target.py 1500 1 1 5 > $O/${TAG}_ncu_q1ws5.log 2>&1; tail -1 $O/${TAG}_ncu_q1ws5.log
ls -la $O | tail -12
