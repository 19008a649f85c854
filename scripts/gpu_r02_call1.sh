#!/bin/bash
# Round-2 GPU call 1 (1 GPU): parity suite incl. the new kind-K / local-group / upload tests and the experimental
# fragment-permutation tests, perm probe, stored-AO legs, ncu captures of the expansion and scatter kernels, the default bench line.
TAG=${1:-r02a}
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --durations=15 > $O/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_gpu.log ); tail -40 $O/${TAG}_pytest_gpu.log
( LOWDIN_IT_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -k perm > $O/${TAG}_pytest_perm.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_perm.log ); tail -8 $O/${TAG}_pytest_perm.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "exit $?" >> $O/${TAG}_smoke.log ); tail -2 $O/${TAG}_smoke.log
timeout 300 python scripts/perm_probe.py $TAG > $O/${TAG}_perm_probe.log 2>&1; cat $O/${TAG}_perm_probe.log
timeout 200 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --frag-perm 1 > $O/${TAG}_bench_n1500_perm1.json 2> $O/${TAG}_bench_n1500_perm1.err; python -c "
import json,sys
d=json.loads(open('$O/${TAG}_bench_n1500_perm1.json').read().strip().splitlines()[-1]); print('perm1', round(d['value']), {k:(round(v['ms']), round(v.get('TFLOP/s', v.get('GB/s',0)),2)) for k,v in d['kernels'].items()}, d['parity']['sums'])"
timeout 120 python bench.py --resident-only > $O/${TAG}_resident_n500.json 2> $O/${TAG}_resident_n500.err; tail -c 2500 $O/${TAG}_resident_n500.json; tail -3 $O/${TAG}_resident_n500.err
timeout 120 python bench.py --stored-only --push-mode blocks --steps 3 > $O/${TAG}_stored_n120_blocks.json 2> $O/${TAG}_stored_n120_blocks.err; tail -c 1500 $O/${TAG}_stored_n120_blocks.json; tail -3 $O/${TAG}_stored_n120_blocks.err
timeout 120 python bench.py --stored-only --push-mode stacks --steps 3 > $O/${TAG}_stored_n120_stacks.json 2> $O/${TAG}_stored_n120_stacks.err; tail -c 1500 $O/${TAG}_stored_n120_stacks.json; tail -3 $O/${TAG}_stored_n120_stacks.err
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
timeout 300 $NCU -k regex:expand_block -s 2 -c 1 -o $O/${TAG}_full_expand1_stored_n500 python bench.py --resident-only > $O/${TAG}_ncu_expand1.log 2>&1; tail -1 $O/${TAG}_ncu_expand1.log
timeout 300 $NCU -k regex:scatter_stacks -s 2 -c 1 -o $O/${TAG}_full_scatter python bench.py --stored-only --push-mode blocks --steps 1 > $O/${TAG}_ncu_scatter.log 2>&1; tail -1 $O/${TAG}_ncu_scatter.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/${TAG}_bench_n1500.json 2> $O/${TAG}_bench_n1500.err; tail -c 6000 $O/${TAG}_bench_n1500.json; tail -3 $O/${TAG}_bench_n1500.err
ls -la $O | tail -20
