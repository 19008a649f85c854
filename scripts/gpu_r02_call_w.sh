#!/bin/bash
# Round-2 GPU call w (1 GPU): two-CTA third-quarter kernels: parity test, kernel-alone probe, one N=1500 pass with / without.
TAG=${1:-r02w}
mkdir -p gpurun_out
O=gpurun_out
( timeout 300 python -m pytest tests/test_gpu_variants.py -m gpu -q -k "two_cta" -p timeout --timeout 150 > $O/${TAG}_pytest_q3.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_q3.log ); tail -15 $O/${TAG}_pytest_q3.log | cut -c1-250
timeout 200 python scripts/q3_probe.py $TAG > $O/${TAG}_q3_probe.log 2>&1; cat $O/${TAG}_q3_probe.log | tail -16
for TWO in 0 256; do
  timeout 400 python bench.py --steps 1 --warmup 3 --q3-two-cta $TWO --no-cpu-baseline --no-e2e --stored-nbf 0 --resident-nbf 0 > $O/${TAG}_bench_n1500_q3two$TWO.json 2> $O/${TAG}_bench_n1500_q3two$TWO.err
  python -c "
import json
d=json.loads(open('$O/${TAG}_bench_n1500_q3two$TWO.json').read().strip().splitlines()[-1]); print('q3 two-cta $TWO', round(d['value']), round(d['ms_per_step'],1), {k:(round(v['ms']), round(v.get('TFLOP/s', v.get('GB/s',0)),2)) for k,v in d['kernels'].items()}, d['parity'].get('whole_transform_vs_reference_sums',{}).get('ok'))"
  tail -2 $O/${TAG}_bench_n1500_q3two$TWO.err
done
