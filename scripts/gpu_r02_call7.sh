#!/bin/bash
# Round-2 GPU call 7 (1 GPU): whole parity suite, resident leg, sweep lines N=500/1000/2000, ncu of the fused stored first quarter and of the row completion.
TAG=${1:-r02g}
mkdir -p gpurun_out
O=gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -p timeout --timeout 200 --durations=8 > $O/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_gpu.log ); tail -30 $O/${TAG}_pytest_gpu.log
timeout 200 python bench.py --resident-only --resident-all > $O/${TAG}_resident_n500.json 2> $O/${TAG}_resident_n500.err; tail -c 5000 $O/${TAG}_resident_n500.json; tail -3 $O/${TAG}_resident_n500.err
for N in 500 1000 2000; do
  timeout 700 python bench.py --nbf $N --steps 3 --warmup 3 --no-cpu-baseline --stored-nbf 0 --resident-nbf 0 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err
  python -c "
import json
d=json.loads(open('$O/${TAG}_bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N', round(d['value']), 'GFLOP/s', round(d['ms_per_step'],1), 'ms/step', d['config']['occ_batch'], 'occ/pass x', d['config']['passes_per_transform'], 'e2e', round(d['e2e']['value'] or 0), {k:(round(v['ms']), round(v.get('TFLOP/s', v.get('GB/s',0)),2)) for k,v in d['kernels'].items()}, d['parity'])"
  tail -2 $O/${TAG}_bench_n$N.err
done
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
timeout 300 $NCU -k regex:q1_load_ws5 -s 3 -c 1 -o $O/${TAG}_full_q1load_n500 python bench.py --resident-only > $O/${TAG}_ncu_q1load.log 2>&1; tail -1 $O/${TAG}_ncu_q1load.log
timeout 300 $NCU -k regex:complete_rows -s 3 -c 1 -o $O/${TAG}_full_complete_rows_n500 python bench.py --resident-only > $O/${TAG}_ncu_complete.log 2>&1; tail -1 $O/${TAG}_ncu_complete.log
ls -la $O | tail -8
