#!/bin/bash
# One `ncu --set full` capture of one kernel, exported to text ON THE BOX (the .ncu-rep files are 30-40 MB each and gpurun brings back
# at most 64 MiB): gpurun_out/<name>.details.txt (ncu --page details) and gpurun_out/<name>.raw.csv (every metric of the launch).
# usage: scripts/ncu_capture.sh <name> <kernel regex> <launches to skip> <command...>
NAME=$1; KRE=$2; SKIP=$3; shift 3
mkdir -p gpurun_out
REP=/tmp/${NAME}.ncu-rep
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f -k regex:"$KRE" -s $SKIP -c 1 -o $REP "$@" > gpurun_out/${NAME}.ncu.log 2>&1
tail -1 gpurun_out/${NAME}.ncu.log
ncu -i $REP --page details > gpurun_out/${NAME}.details.txt 2>/dev/null
ncu -i $REP --page raw --csv > gpurun_out/${NAME}.raw.csv 2>/dev/null
rm -f $REP
grep -E "^  [a-z].*\(|Duration|DRAM Throughput|Memory Throughput  |Compute \(SM\) Throughput|Registers Per|highest-utilized" gpurun_out/${NAME}.details.txt | head -8
