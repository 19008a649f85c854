#!/bin/bash
TAG=${1:-r02i}
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python scripts/diag_fused.py 500 50 6000 0.0 0.5 0.95 > $O/${TAG}_diag_fused.log 2>&1; grep -c "0 of 150000000" $O/${TAG}_diag_fused.log; grep -v "^  " $O/${TAG}_diag_fused.log | cut -c1-200
timeout 200 python bench.py --resident-only --resident-all > $O/${TAG}_resident_n500.json 2> $O/${TAG}_resident_n500.err; python -c "
import json
d=json.loads(open('$O/${TAG}_resident_n500.json').read().strip().splitlines()[-1])['stored_ao_resident']
for k in ('mp2','mp2_unfused','all'):
    print(k, round(d[k]['value']), round(d[k]['ms_per_transform'],1), {c:(round(v['ms'],1), round(v.get('TFLOP/s', v.get('GB/s',0)),2)) for c,v in d[k]['kernels'].items()}, d[k].get('vs_generated_source'), d[k].get('vs_fused'))"
tail -3 $O/${TAG}_resident_n500.err
( timeout 1200 python -m pytest tests -m gpu -q -p timeout --timeout 200 --durations=8 > $O/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_gpu.log ); tail -40 $O/${TAG}_pytest_gpu.log | cut -c1-250
timeout 300 python scripts/q_probe.py $TAG > $O/${TAG}_q_probe.log 2>&1; grep "q1 variant 5\|k 32\|k 128" $O/${TAG}_q_probe.log
timeout 200 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/${TAG}_bench_n1500_1step.json 2> $O/${TAG}_bench_n1500_1step.err; python -c "
import json
d=json.loads(open('$O/${TAG}_bench_n1500_1step.json').read().strip().splitlines()[-1]); print('1step', round(d['value']), {k:(round(v['ms']), round(v.get('TFLOP/s', v.get('GB/s',0)),2)) for k,v in d['kernels'].items()})"
