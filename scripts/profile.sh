#!/bin/bash
# ncu evidence for profiles/: (1) launch list with device time of every launch of a short bench run,
# (2) one --set full capture of each hot kernel.  Run under gpurun (one GPU); numbers printed by bench.py
# under ncu are NOT bench values.  Usage: scripts/profile.sh <tag> [nbf]
TAG=${1:-r01}; NBF=${2:-1500}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --nbf $NBF --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches_n${NBF}.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_launches.log
for K in q1_gen_kernel dgemm_tn_kernel expand_block_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 3 -f -o gpurun_out/${TAG}_${K}_n${NBF} $CMD > gpurun_out/${TAG}_ncu_${K}.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_${K}.log
done
ls -la gpurun_out/
