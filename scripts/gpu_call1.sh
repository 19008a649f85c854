#!/bin/bash
# Round-1 GPU call: parity tests, bench line, ncu launch list of the bench command, --set full captures of the hot kernels.
TAG=${1:-r01b}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/${TAG}_smi.txt 2>&1
( timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest_gpu.log ) 
tail -3 $O/${TAG}_pytest_gpu.log
timeout 300 python scripts/q1_probe.py > $O/${TAG}_q1_probe.log 2>&1; tail -30 $O/${TAG}_q1_probe.log
timeout 600 python bench.py --steps 2 --warmup 1 > $O/${TAG}_bench_n1500.json 2> $O/${TAG}_bench_n1500.err; tail -c 600 $O/${TAG}_bench_n1500.json
# full captures (light target; --kill ends the process once the capture is done)
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
timeout 300 $NCU -k regex:q1_gen -s 6 -c 1 --kill 1 -o $O/${TAG}_full_q1gen_n1500 python scripts/ncu_target.py 1500 > $O/${TAG}_ncu_q1.log 2>&1; tail -1 $O/${TAG}_ncu_q1.log
timeout 300 $NCU -k regex:EpiScatterH -s 6 -c 1 --kill 1 -o $O/${TAG}_full_q2_n1500 python scripts/ncu_target.py 1500 > $O/${TAG}_ncu_q2.log 2>&1; tail -1 $O/${TAG}_ncu_q2.log
timeout 300 $NCU -k regex:EpiAccT -s 6 -c 2 --kill 1 -o $O/${TAG}_full_q3_n1500 python scripts/ncu_target.py 1500 > $O/${TAG}_ncu_q3.log 2>&1; tail -1 $O/${TAG}_ncu_q3.log
timeout 300 $NCU -k regex:expand_block -s 6 -c 2 --kill 1 -o $O/${TAG}_full_expand2_n1500 python scripts/ncu_target.py 1500 > $O/${TAG}_ncu_ex.log 2>&1; tail -1 $O/${TAG}_ncu_ex.log
timeout 300 $NCU -k "regex:EpiOut|reduce_block" -s 2 -c 2 -o $O/${TAG}_full_q4_n1000 python scripts/ncu_target.py 1000 > $O/${TAG}_ncu_q4.log 2>&1; tail -1 $O/${TAG}_ncu_q4.log
# launch list of the bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/${TAG}_launches_n1500.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_launches.log 2>&1; tail -c 300 $O/${TAG}_ncu_launches.log
ls -la $O | tail -30
