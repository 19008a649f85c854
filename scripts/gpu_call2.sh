#!/bin/bash
# GPU call 2: validate the TMA GEMM + warp-specialised first quarter (separate pytest processes so that a kernel fault in a
# new variant cannot poison the regression run), probe them alone, bench N=1500 with them, short ncu launch lists.
TAG=${1:-r01c}
mkdir -p gpurun_out
O=gpurun_out
( timeout 300 python -m pytest tests/test_gpu_variants.py -m gpu -q -x -k "tma_gemm" > $O/${TAG}_pytest_tma_gemm.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_tma_gemm.log ); tail -4 $O/${TAG}_pytest_tma_gemm.log
( timeout 400 python -m pytest tests/test_gpu_variants.py -m gpu -q -k "not tma_gemm" > $O/${TAG}_pytest_tma_rest.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_tma_rest.log ); tail -12 $O/${TAG}_pytest_tma_rest.log
( timeout 400 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_variants.py > $O/${TAG}_pytest_gpu_regress.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_gpu_regress.log ); tail -3 $O/${TAG}_pytest_gpu_regress.log
timeout 300 python scripts/variant_probe.py $TAG > $O/${TAG}_variant_probe.log 2>&1; cat $O/${TAG}_variant_probe.log
timeout 400 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --gemm-variant 2 --q1-variant 3 > $O/${TAG}_bench_n1500_g2q3.json 2> $O/${TAG}_bench_n1500_g2q3.err; tail -c 1500 $O/${TAG}_bench_n1500_g2q3.json; tail -3 $O/${TAG}_bench_n1500_g2q3.err
timeout 200 python bench.py --nbf 500 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --gemm-variant 2 --q1-variant 3 > $O/${TAG}_bench_n500_g2q3.json 2> $O/${TAG}_bench_n500_g2q3.err; tail -c 900 $O/${TAG}_bench_n500_g2q3.json
# launch list (every kernel launch with its device time) of one N=1000 pass with the new variants
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${TAG}_launches_n1000.csv \
  python bench.py --nbf 1000 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --gemm-variant 2 --q1-variant 3 > $O/${TAG}_ncu_launches.log 2>&1; tail -c 300 $O/${TAG}_ncu_launches.log; wc -l $O/${TAG}_launches_n1000.csv
ls -la $O | tail -12
