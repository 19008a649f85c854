#!/bin/bash
TAG=${1:-r02h}
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python scripts/diag_fused.py 500 50 6000 0.0 0.3 0.6 0.95 > $O/${TAG}_diag_fused.log 2>&1; cat $O/${TAG}_diag_fused.log | cut -c1-400
timeout 200 python bench.py --resident-only > $O/${TAG}_resident_n500.json 2> $O/${TAG}_resident_n500.err; python -c "
import json
d=json.loads(open('$O/${TAG}_resident_n500.json').read().strip().splitlines()[-1])['stored_ao_resident']
for k in ('mp2','mp2_unfused'):
    print(k, round(d[k]['value']), d[k]['ms_per_transform'], {c:(round(v['ms'],1), round(v.get('TFLOP/s', v.get('GB/s',0)),2)) for c,v in d[k]['kernels'].items()}, d[k].get('vs_generated_source'), d[k].get('vs_fused'))"
tail -3 $O/${TAG}_resident_n500.err
( timeout 1200 python -m pytest tests -m gpu -q -p timeout --timeout 200 --durations=8 > $O/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_gpu.log ); tail -60 $O/${TAG}_pytest_gpu.log | cut -c1-250
