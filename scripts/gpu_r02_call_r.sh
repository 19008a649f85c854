#!/bin/bash
# Round-2 GPU call r (1 GPU): parity suite, N_bf sweep lines (500/1000/2000), tall-tile GEMM probe (alone and in the N=1500 pass),
# launch list of an N=1000 pass, ncu --set full of the kernels new in this round (details + raw csv exported on the box).
TAG=${1:-r02r}
mkdir -p gpurun_out
O=gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -p timeout --timeout 250 --durations=6 > $O/${TAG}_pytest_gpu.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest_gpu.log ); tail -12 $O/${TAG}_pytest_gpu.log | cut -c1-220
for N in 500 1000 2000; do
  ST=3; WU=3; [ $N = 2000 ] && { ST=2; WU=1; }
  timeout 700 python bench.py --nbf $N --steps $ST --warmup $WU --no-cpu-baseline --stored-nbf 0 --resident-nbf 0 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err
  python -c "
import json
d=json.loads(open('$O/${TAG}_bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N', round(d['value']), 'GFLOP/s', round(d['ms_per_step'],1), 'ms/step', d['config']['occ_batch'], 'occ/pass x', d['config']['passes_per_transform'], 'e2e', round(d['e2e']['value'] or 0), {k:(round(v['ms']), round(v.get('TFLOP/s', v.get('GB/s',0)),2)) for k,v in d['kernels'].items()}, d['parity'].get('passes_covered'), d['e2e']['note'][-100:])"
  tail -2 $O/${TAG}_bench_n$N.err
done
timeout 200 python scripts/tall_probe.py $TAG > $O/${TAG}_gemm_tall_probe.log 2>&1; cat $O/${TAG}_gemm_tall_probe.log
for TALL in 0 1; do
  timeout 400 python bench.py --steps 1 --warmup 3 --gemm-tall $TALL --no-cpu-baseline --no-e2e --stored-nbf 0 --resident-nbf 0 > $O/${TAG}_bench_n1500_tall$TALL.json 2> $O/${TAG}_bench_n1500_tall$TALL.err
  python -c "
import json
d=json.loads(open('$O/${TAG}_bench_n1500_tall$TALL.json').read().strip().splitlines()[-1]); print('tall $TALL', round(d['value']), round(d['ms_per_step'],1), {k:(round(v['ms']), round(v.get('TFLOP/s', v.get('GB/s',0)),2)) for k,v in d['kernels'].items()})"
  tail -2 $O/${TAG}_bench_n1500_tall$TALL.err
done
# launch list (share of each kernel in a pass; per-launch times under ncu are serialised and cold)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/${TAG}_launches_n1000.csv python scripts/ncu_target.py 1000 1 > $O/${TAG}_launches_n1000.log 2>&1; tail -1 $O/${TAG}_launches_n1000.log | cut -c1-200
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
cap() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 400 $NCU -k regex:$rx -s $skip -c 1 -o $O/${TAG}_full_$name "$@" > $O/${TAG}_ncu_$name.log 2>&1
  if [ -f $O/${TAG}_full_$name.ncu-rep ]; then
    ncu -i $O/${TAG}_full_$name.ncu-rep --page details > $O/${TAG}_ncu_full_$name.details.txt 2>/dev/null
    ncu -i $O/${TAG}_full_$name.ncu-rep --page raw --csv > $O/${TAG}_ncu_full_$name.raw.csv 2>/dev/null
    ls -la $O/${TAG}_full_$name.ncu-rep
    [ $(stat -c %s $O/${TAG}_full_$name.ncu-rep) -gt 12000000 ] && rm -f $O/${TAG}_full_$name.ncu-rep
    python - $O/${TAG}_ncu_full_$name.raw.csv <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, val = rows[0], rows[-1]
want = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_pipe_fp64_op_dmma.sum")
for h, u, v in zip(hdr, rows[1], val):
    if h in want: print(" ", h[:70], v[:60], u)
PY
  else tail -3 $O/${TAG}_ncu_$name.log; fi
}
cap q1load_n500 q1_load_ws5 3 python bench.py --resident-only
cap complete_rows_n500 complete_rows 3 python bench.py --resident-only
cap q1gen5_n1000 q1_gen_ws5 2 python scripts/ncu_target.py 1000 1
cap q2_n1000 "dgemm_tma_kernel.*EpiScatterH" 2 python scripts/ncu_target.py 1000 1
cap q3_n1000 "dgemm_tma_kernel.*EpiAccT" 4 python scripts/ncu_target.py 1000 1
ls -la $O | grep $TAG | awk '{print $5, $9}'
