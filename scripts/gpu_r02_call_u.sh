#!/bin/bash
# Round-2 GPU call u (1 GPU): the f4 tests again after the grid fix, ERI / list-mode timing probes, stored-AO e2e with host-side phase timers.
TAG=${1:-r02u}
mkdir -p gpurun_out
O=gpurun_out
( timeout 400 python -m pytest tests/test_gpu_eri.py -m gpu -q -p timeout --timeout 150 > $O/${TAG}_pytest_eri.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_eri.log ); tail -30 $O/${TAG}_pytest_eri.log | cut -c1-250
timeout 200 python scripts/eri_probe.py $TAG > $O/${TAG}_eri_probe.log 2>&1; tail -3 $O/${TAG}_eri_probe.log | cut -c1-500
timeout 300 python scripts/list_probe.py $TAG > $O/${TAG}_list_probe.log 2>&1; tail -2 $O/${TAG}_list_probe.log | cut -c1-1500
for i in 1 2; do timeout 200 python bench.py --stored-only --steps 5 > $O/${TAG}_stored_n120_run$i.json 2> $O/${TAG}_stored_n120_run$i.err; cut -c1-1800 $O/${TAG}_stored_n120_run$i.json; tail -2 $O/${TAG}_stored_n120_run$i.err; done
