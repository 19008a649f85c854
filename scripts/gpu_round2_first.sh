#!/bin/bash
# First GPU call of the next round (1 GPU, ~25 min of box time): everything round 1 ended without being able to measure.
#   gpurun --timeout 1800 -- 'bash scripts/gpu_round2_first.sh r02a'
# 1. parity suite (new: P2 poles at the HCN.e+ shapes, lowdin_host_run_program file-to-file)
# 2. the default bench line (with e2e and e2e_stored_ao) and the reference arm
# 3. ncu launch list of the same bench command + --set full captures of the expansion and scatter kernels
TAG=${1:-r02a}
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_gpu.log ); tail -6 $O/${TAG}_pytest_gpu.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "exit $?" >> $O/${TAG}_smoke.log ); tail -2 $O/${TAG}_smoke.log
# the opt-in conflict-free fragment mapping (never run on a GPU in round 1): parity first, then kernels alone, then one pass
( LOWDIN_IT_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -k perm > $O/${TAG}_pytest_perm.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_perm.log ); tail -4 $O/${TAG}_pytest_perm.log
timeout 400 python scripts/variant_probe.py $TAG > $O/${TAG}_variant_probe.log 2>&1; grep "^perm" $O/${TAG}_variant_probe.log
timeout 200 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --frag-perm 1 > $O/${TAG}_bench_n1500_perm.json 2> $O/${TAG}_bench_n1500_perm.err; tail -c 1500 $O/${TAG}_bench_n1500_perm.json
timeout 600 python bench.py --steps 3 --warmup 3 > $O/${TAG}_bench_n1500.json 2> $O/${TAG}_bench_n1500.err; tail -c 3000 $O/${TAG}_bench_n1500.json; tail -3 $O/${TAG}_bench_n1500.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; cat $O/${TAG}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${TAG}_launches_n1500.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_launches.log 2>&1; tail -2 $O/${TAG}_ncu_launches.log
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
timeout 200 $NCU -k regex:expand_block -s 4 -c 1 --kill 1 -o $O/${TAG}_full_expand2_n1500 python scripts/ncu_target.py 1500 1 > $O/${TAG}_ncu_expand.log 2>&1; tail -1 $O/${TAG}_ncu_expand.log
timeout 200 $NCU -k regex:scatter_stacks -c 1 -o $O/${TAG}_full_scatter python bench.py --stored-only > $O/${TAG}_ncu_scatter.log 2>&1; tail -1 $O/${TAG}_ncu_scatter.log
ls -la $O | tail -12
