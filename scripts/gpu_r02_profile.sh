#!/bin/bash
# Round-2 profiling call (1 GPU): ncu launch list of the bench command + one --set full capture per hot kernel, exported to text.
TAG=${1:-r02p}
mkdir -p gpurun_out
O=gpurun_out
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file $O/${TAG}_launches_n1500.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --stored-nbf 0 --resident-nbf 0 > $O/${TAG}_ncu_launches.log 2>&1; tail -2 $O/${TAG}_ncu_launches.log; wc -l $O/${TAG}_launches_n1500.csv
bash scripts/ncu_capture.sh ${TAG}_full_q1gen_ws5_n1500 q1_gen_ws5 4 python scripts/ncu_target.py 1500 1 1
bash scripts/ncu_capture.sh ${TAG}_full_q2_scatterH_n1500 "dgemm_tma_kernel<.int.128, .int.128.*EpiScatterH" 4 python scripts/ncu_target.py 1500 1 1
bash scripts/ncu_capture.sh ${TAG}_full_q3_accT_n1500 "dgemm_tma_kernel.*EpiAccT" 6 python scripts/ncu_target.py 1500 1 1
bash scripts/ncu_capture.sh ${TAG}_full_expand2_n1500 "expand_block_kernel<.int.1>" 6 python scripts/ncu_target.py 1500 1 1
bash scripts/ncu_capture.sh ${TAG}_full_q1load_ws5_n500 q1_load_ws5 3 python bench.py --resident-only
bash scripts/ncu_capture.sh ${TAG}_full_complete_rows_n500 complete_rows 3 python bench.py --resident-only
bash scripts/ncu_capture.sh ${TAG}_full_expand_packed_n500 "expand_block_kernel<.int.0>" 2 python bench.py --resident-only
bash scripts/ncu_capture.sh ${TAG}_full_scatter_stacks_n120 scatter_stacks 2 python bench.py --stored-only --push-mode blocks --steps 1
ls -la $O | grep $TAG | head -40
