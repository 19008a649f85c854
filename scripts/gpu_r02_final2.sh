#!/bin/bash
# Round-2 closing call (1 GPU): whole parity suite, smoke (with the f4 check), a quick N=500 bench line through the final bench.py.
TAG=${1:-r02f}
mkdir -p gpurun_out
O=gpurun_out
( timeout 400 python -m pytest tests -m gpu -q -x -p timeout --timeout 250 --durations=4 > $O/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_gpu.log ); tail -9 $O/${TAG}_pytest_gpu.log | cut -c1-220
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "exit $?" >> $O/${TAG}_smoke.log ); tail -2 $O/${TAG}_smoke.log | cut -c1-400
timeout 200 python bench.py --nbf 500 --steps 3 --warmup 3 --no-cpu-baseline --stored-nbf 0 --resident-nbf 0 > $O/${TAG}_bench_n500.json 2> $O/${TAG}_bench_n500.err; python -c "
import json
d=json.loads(open('$O/${TAG}_bench_n500.json').read().strip().splitlines()[-1]); print('N=500', round(d['value']), round(d['ms_per_step'],1), d['config']['exchange'], 'e2e', round(d['e2e']['value'] or 0), d['parity'].get('whole_transform_vs_reference_sums',{}).get('ok'))"
tail -2 $O/${TAG}_bench_n500.err
