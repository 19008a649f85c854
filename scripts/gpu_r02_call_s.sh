#!/bin/bash
# Round-2 GPU call s (1 GPU): whole parity suite (with the f4 tests), smoke, the default bench line + reference arm, ERI producer timing.
TAG=${1:-r02s}
mkdir -p gpurun_out
O=gpurun_out
( timeout 300 python -m pytest tests/test_gpu_eri.py -m gpu -q -x -p timeout --timeout 120 > $O/${TAG}_pytest_eri.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_eri.log ); tail -25 $O/${TAG}_pytest_eri.log | cut -c1-250
timeout 200 python scripts/eri_probe.py $TAG > $O/${TAG}_eri_probe.log 2>&1; tail -3 $O/${TAG}_eri_probe.log | cut -c1-400
( timeout 1200 python -m pytest tests -m gpu -q -p timeout --timeout 250 --durations=6 > $O/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_gpu.log ); tail -12 $O/${TAG}_pytest_gpu.log | cut -c1-220
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "exit $?" >> $O/${TAG}_smoke.log ); tail -2 $O/${TAG}_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/${TAG}_bench_n1500.json 2> $O/${TAG}_bench_n1500.err; python -c "
import json
d=json.loads(open('$O/${TAG}_bench_n1500.json').read().strip().splitlines()[-1]); print('N=1500', round(d['value']), 'GFLOP/s', round(d['ms_per_step'],1), 'ms/step e2e', round(d['e2e']['value']), 'd2h', d['e2e']['d2h_bytes_per_step'], d['roofline']['frac'], {k:(round(v['ms']), round(v.get('TFLOP/s', v.get('GB/s',0)),2)) for k,v in d['kernels'].items()}); print(d['parity']); print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','upload_gb_per_s','max_abs_diff','ok')}) for k,v in d['cpu_baseline'].items()}); print(d['e2e_stored_ao']['ms_per_step'], d['e2e_stored_ao']['value']); r=d['stored_ao_resident']; print({k:(round(r[k]['value']), round(r[k]['ms_per_transform'],1)) for k in ('mp2','mp2_unfused')})"
tail -3 $O/${TAG}_bench_n1500.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; cut -c1-300 $O/${TAG}_bench_reference.json
